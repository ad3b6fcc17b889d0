"""CPU oracle: a plain-PyTorch (fp32, CPU) restatement of the reference's sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing in `diffusion_model_project_b200/` imports this
package.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and there only as the checker
or as the timed CPU baseline -- never as the product path.

Pinning: the reference ships no golden vectors or known-answer tests for this path
("parity unpinned" by any reference file), so the
oracle is pinned against outputs of the reference itself, generated in the build
container by importing /root/reference (tests/golden/make_golden.py) and committed
under tests/golden/*.npz, plus the scheduler known-answer values of SURVEY.md 8(c).
`oracle/train.py` restates one UNet training step (the next row; no kernels yet) and is pinned
the same way (tests/golden/make_train_golden.py -> train_step.npz).
"""
