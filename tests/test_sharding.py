"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and gather it back."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_model_project_b200 import sharding


def test_shard_range_covers_batch():
    for B in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_chunk_starts_cover_every_sample_with_full_windows():
    """The VAE micro-batching of predict / predict_ddim (B = 64 on one GPU in chunks of 8; ragged tails)."""
    assert sharding.chunk_starts(64, 8) == list(range(0, 64, 8))
    assert sharding.chunk_starts(3, 2) == [0, 1] and sharding.chunk_starts(3, 1) == [0, 1, 2] and sharding.chunk_starts(5, 5) == [0]
    assert sharding.chunk_starts(19, 8) == [0, 8, 11]
    for B in range(1, 70):
        for chunk in range(1, min(B, 9) + 1):
            st = sharding.chunk_starts(B, chunk)
            assert st[0] == 0 and st == sorted(set(st)) and all(0 <= c0 and c0 + chunk <= B for c0 in st)
            covered = set()
            for c0 in st:
                covered.update(range(c0, c0 + chunk))
            assert covered == set(range(B))
            assert len(st) == -(-B // chunk)                       # no more passes than a ragged split would need
    for bad in ((0, 1), (4, 0), (4, 5)):
        with pytest.raises(ValueError):
            sharding.chunk_starts(*bad)


class _FakePredictor:
    def predict_ddim(self, img, v2d, noise=None, **kw):
        if noise is None:
            return v2d * 2
        return v2d * 2 + noise.reshape(v2d.shape[0], v2d.shape[1], noise.shape[1:].numel()).sum(-1)[:, :, None, None, None]


def _worker(rank, world, port, B):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        img = torch.rand(B, 3, 1, 4, 4, generator=g)
        v2d = torch.rand(B, 3, 3, 4, 4, generator=g)
        noise = torch.rand(B * 3, 2, 2, 2, generator=g)
        full = sharding.predict_sharded(_FakePredictor(), img, v2d, noise)
        ref = _FakePredictor().predict_ddim(img, v2d, noise)
        assert full.shape == ref.shape and torch.allclose(full, ref)
        only0 = sharding.predict_sharded(_FakePredictor(), img, v2d, None, dst=0)
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert torch.allclose(only0, v2d * 2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_gloo_world2_shard_and_gather(B):
    """Even, ragged, and fewer samples than ranks (rank 1's shard is empty: the predictor hands back an empty field tensor
    and the gather pads it)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, B), nprocs=2, join=True)



def _allreduce_worker(rank, world, port):
    from diffusion_model_project_b200.train import FlatAdam
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        params = {"w": torch.randn(5, 3, generator=g), "b": torch.randn(7, generator=g)}
        opt = FlatAdam(params, device="cpu")
        for k in opt.names:
            opt.view(opt.grad, k).fill_(float(rank + 1))
        scale = opt.allreduce_gradients()
        assert scale == 0.5
        assert opt.offsets["w"][:2] == (0, 15) and opt.offsets["b"][:2] == (16, 7)               # 16-byte aligned segments
        assert torch.equal(opt.grad[0:15], torch.full((15,), 3.0)) and opt.grad[15] == 0         # 1 + 2; padding untouched
        assert torch.equal(opt.grad[16:23], torch.full((7,), 3.0))
        assert torch.equal(opt.state_dict()["w"], params["w"])
        with pytest.raises(RuntimeError):
            opt.step(scale)                                                                      # the update kernel is GPU-only
        # the training step's bucketed form: ranges in backward order, reduced one by one, equal one whole-buffer all-reduce
        names = ["time_mlp.0.weight", "encoder.0.w", "encoder.1.w", "bottleneck.a", "bottleneck.b", "decoder.0.w", "final_conv.weight"]
        gg = torch.Generator().manual_seed(7)
        big = FlatAdam({k: torch.randn(3 + 2 * i, 5, generator=gg) for i, k in enumerate(names)}, device="cpu")
        buckets = big.backward_order_buckets(("bottleneck.", "decoder."))
        assert buckets[0] == (big.offsets["decoder.0.w"][0], big.numel) and buckets[1][0] == big.offsets["bottleneck.a"][0]
        assert buckets[2] == (0, big.offsets["bottleneck.a"][0]) and [b[1] for b in buckets[1:]] == [b[0] for b in buckets[:-1]]
        big.grad.copy_(torch.arange(big.numel, dtype=torch.float32) * (rank + 1))
        whole = big.grad.clone()
        dist.all_reduce(whole)
        for b in buckets:
            big.allreduce_bucket(b)
        assert torch.equal(big.grad, whole)
        with pytest.raises(ValueError):
            big.backward_order_buckets(("decoder.", "bottleneck."))
        with pytest.raises(KeyError):
            big.backward_order_buckets(("nothing.",))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    """Training slice: every parameter's gradient lives in ONE flat buffer, summed over the ranks by one
    collective; the update then applies 1 / world.  Host logic on CPU with gloo (the Adam kernel itself is GPU-only)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_allreduce_worker, args=(2, port), nprocs=2, join=True)
