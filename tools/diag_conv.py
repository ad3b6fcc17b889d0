"""First-contact diagnostics for the tcgen05 conv engine on a real B200.  Each case runs in its own
process (a device trap must not take the other cases down) and prints an error map.
usage: python tools/diag_conv.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["gemm64", "gemm128", "gemm256", "gemm16", "gemm_k512", "conv2d_64", "conv2d_small", "conv2d_cat", "convT", "conv3d",
         "conv3d_s2", "residual_stats", "split", "out_modes"]


def run_case(name):
    import torch
    import torch.nn.functional as F
    from diffusion_model_project_b200 import engine
    from diffusion_model_project_b200.engine import ConvPlan, new_act
    from util import bf16_round, from_act, no_tf32, rel_err, stats_ref, to_act
    no_tf32()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1)
    s = torch.cuda.current_stream().cuda_stream

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev)

    def report(tag, got, ref):
        e = rel_err(got, ref)
        print(f"[{name}] {tag}: rel_err={e:.3e} ref_max={ref.abs().max().item():.3e} got_max={got.abs().max().item():.3e}", flush=True)
        if not (e < 2e-2):
            d = (got - ref).abs()
            # where are the errors: per-channel and per-position summaries
            dims = list(range(d.dim()))
            ch = d.amax(dim=[i for i in dims if i != 1])
            print("  worst channels:", torch.topk(ch, min(8, ch.numel())).indices.tolist())
            bad = (d > 1e-2 * ref.abs().max()).float()
            print("  bad fraction:", bad.mean().item(), "by channel block of 8:", bad.mean(dim=[i for i in dims if i != 1]).reshape(-1, 8).mean(1).tolist()[:32])
            pos = bad.mean(dim=1).reshape(bad.shape[0], -1)
            print("  bad fraction per image:", pos.mean(1).tolist()[:16])
        return e

    if name.startswith("gemm"):
        bn = {"gemm64": 64, "gemm128": 128, "gemm256": 256, "gemm16": 16, "gemm_k512": 128}[name]
        K = 512 if name == "gemm_k512" else 64
        cout = 8 if bn == 16 else bn * 2
        M_img, Hh, Ww = 2, 16, 16  # M = 512
        x = bf16_round(rnd(M_img, K, 1, Hh, Ww))
        w = bf16_round(rnd(cout, K, scale=K ** -0.5))
        b = rnd(cout)
        pw = engine.pack_linear(w, b, dev)
        xa = to_act(x)
        if bn == 16:
            out = torch.zeros(M_img, cout, Hh, Ww, device=dev)
            plan = ConvPlan([xa], pw, out, cout=cout, out_mode=1, out_cstride=cout, block_n=bn)
        else:
            out = new_act(M_img, 1, Hh, Ww, cout, dev)
            plan = ConvPlan([xa], pw, out, cout=cout, block_n=bn)
        print(plan.info(), flush=True)
        plan.run(s)
        torch.cuda.synchronize()
        ref = F.conv2d(x[:, :, 0], w[:, :, None, None], b)
        got = out if bn == 16 else from_act(out, cout)[:, :, 0]
        return report("gemm", got, ref)

    if name in ("conv2d_64", "conv2d_small", "conv2d_cat"):
        worst = 0.0
        shapes = [(3, 64, 128, 64, 64)] if name == "conv2d_64" else [(11, 128, 128, 4, 4), (5, 64, 64, 2, 2), (7, 64, 64, 1, 1), (3, 64, 64, 8, 8), (2, 17, 64, 32, 32)]
        if name == "conv2d_cat":
            shapes = [(3, 128, 64, 16, 16)]
        for (N, ci, co, Hh, Ww) in shapes:
            x = bf16_round(rnd(N, ci, 1, Hh, Ww))
            w = bf16_round(rnd(co, ci, 3, 3, scale=(9 * ci) ** -0.5))
            if name == "conv2d_cat":
                pw = engine.pack_conv2d(w, [ci // 2, ci // 2], None, dev)
                ins = [to_act(x[:, :ci // 2]), to_act(x[:, ci // 2:])]
            else:
                pw = engine.pack_conv2d(w, [ci], None, dev)
                ins = [to_act(x)]
            out = new_act(N, 1, Hh, Ww, co, dev)
            st = torch.zeros(N, 1, 2, dtype=torch.float64, device=dev)
            plan = ConvPlan(ins, pw, out, cout=co, stats=st, stats_cpg=co)
            print((N, ci, co, Hh, Ww), plan.info(), flush=True)
            plan.run(s)
            torch.cuda.synchronize()
            ref = F.conv2d(x[:, :, 0], w, None, padding=1)
            worst = max(worst, report(f"conv {N,ci,co,Hh,Ww}", from_act(out, co)[:, :, 0], ref))
            sref = stats_ref(ref[:, :, None], 1)
            print("   stats rel err", ((st - sref).abs().max() / sref.abs().max()).item(), flush=True)
        return worst

    if name == "convT":
        N, ci, co, Hh, Ww = 3, 128, 64, 8, 8
        x = bf16_round(rnd(N, ci, 1, Hh, Ww))
        w = bf16_round(rnd(ci, co, 2, 2, scale=ci ** -0.5))
        b = rnd(co)
        pw = engine.pack_convT2x2(w, b, dev)
        out = new_act(N, 1, 2 * Hh, 2 * Ww, co, dev)
        st = torch.zeros(N, 1, 2, dtype=torch.float64, device=dev)
        plan = ConvPlan([to_act(x)], pw, out, cout=co, nphase=4, stats=st, stats_cpg=co)
        print(plan.info(), flush=True)
        plan.run(s)
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(x[:, :, 0], w, b, stride=2)
        e = report("convT", from_act(out, co)[:, :, 0], ref)
        sref = stats_ref(ref[:, :, None], 1)
        print("   stats rel err", ((st - sref).abs().max() / sref.abs().max()).item(), flush=True)
        return e

    if name in ("conv3d", "conv3d_s2"):
        worst = 0.0
        down = name == "conv3d_s2"
        for (N, ci, co, D, Hh, Ww) in [(2, 128, 128, 3, 16, 16), (1, 64, 256, 11, 8, 8), (2, 128, 128, 3, 4, 4)]:
            x = bf16_round(rnd(N, ci, D, Hh, Ww))
            w = bf16_round(rnd(co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
            b = rnd(co)
            pw = engine.pack_conv3d(w, b, dev, down=down)
            st_ = 2 if down else 1
            out = new_act(N, D, Hh // st_, Ww // st_, co, dev)
            st = torch.zeros(N, 32, 2, dtype=torch.float64, device=dev)
            plan = ConvPlan([to_act(x)], pw, out, cout=co, stride=st_, stats=st, stats_cpg=co // 32)
            print((N, ci, co, D, Hh, Ww), plan.info(), flush=True)
            plan.run(s)
            torch.cuda.synchronize()
            if down:
                ref = F.conv3d(F.pad(x, (0, 1, 0, 1, 1, 1)), w, b, stride=(1, 2, 2))
            else:
                ref = F.conv3d(x, w, b, padding=1)
            worst = max(worst, report(f"{name} {N,ci,co,D,Hh,Ww}", from_act(out, co), ref))
            sref = stats_ref(ref, 32)
            print("   stats rel err", ((st - sref).abs().max() / sref.abs().max()).item(), flush=True)
        return worst

    if name == "residual_stats":
        N, c, D, Hh, Ww = 2, 128, 2, 8, 8
        x = bf16_round(rnd(N, c, D, Hh, Ww))
        r = bf16_round(rnd(N, c, D, Hh, Ww))
        w = bf16_round(rnd(c, c, 3, 3, 3, scale=(27 * c) ** -0.5))
        b = rnd(c)
        pw = engine.pack_conv3d(w, b, dev)
        out = new_act(N, D, Hh, Ww, c, dev)
        st = torch.zeros(N, 32, 2, dtype=torch.float64, device=dev)
        plan = ConvPlan([to_act(x)], pw, out, cout=c, residual=to_act(r), stats=st, stats_cpg=4)
        plan.run(s)
        torch.cuda.synchronize()
        ref = F.conv3d(x, w, b, padding=1) + r
        e = report("residual", from_act(out, c), ref)
        sref = stats_ref(ref, 32)
        print("   stats rel err", ((st - sref).abs().max() / sref.abs().max()).item(), flush=True)
        return e

    if name == "split":
        N, ci, co, Hh, Ww = 2, 128, 128, 16, 16
        x = rnd(N, ci, 1, Hh, Ww)
        w = rnd(co, ci, 3, 3, scale=(9 * ci) ** -0.5)
        pw = engine.pack_conv2d(w, [ci], None, dev, split=True)
        out = new_act(N, 1, Hh, Ww, co, dev, split=True)
        plan = ConvPlan([to_act(x, split=True)], pw, out, cout=co)
        print(plan.info(), flush=True)
        plan.run(s)
        torch.cuda.synchronize()
        ref = F.conv2d(x[:, :, 0].double(), w.double(), None, padding=1).float()
        got = from_act(out, co)[:, :, 0]
        e = rel_err(got, ref)
        print(f"[{name}] fp32x conv rel_err={e:.3e}", flush=True)
        return 0.0 if e < 2e-4 else e + 1

    if name == "out_modes":
        N, ci, co, D, Hh, Ww = 2, 128, 3, 2, 16, 16
        x = bf16_round(rnd(N, ci, D, Hh, Ww))
        w = bf16_round(rnd(co, ci, 3, 3, 3, scale=(27 * ci) ** -0.5))
        b = rnd(co)
        scale = torch.tensor([0.5, 2.0, 3.0], device=dev)
        mask = (torch.rand(N, D, Hh, Ww, generator=g) > 0.3).float().to(dev)
        pw = engine.pack_conv3d(w, b, dev)
        out = torch.zeros(N, D, co, Hh, Ww, device=dev)
        plan = ConvPlan([to_act(x)], pw, out, cout=co, out_mode=1, out_cstride=co, out_scale=scale, out_mask=mask)
        plan.run(s)
        torch.cuda.synchronize()
        ref = F.conv3d(x, w, b, padding=1) * scale.view(1, 3, 1, 1, 1) * mask[:, None]
        e = report("planar+scale+mask", out.permute(0, 2, 1, 3, 4), ref)
        out2 = torch.zeros(N, D, Hh, Ww, 4, device=dev)
        plan2 = ConvPlan([to_act(x)], pw, out2, cout=co, out_mode=2, out_cstride=4)
        plan2.run(s)
        torch.cuda.synchronize()
        ref2 = F.conv3d(x, w, b, padding=1)
        e2 = report("fp32 NDHWC", out2[..., :3].permute(0, 4, 1, 2, 3), ref2)
        return max(e, e2)
    raise SystemExit(f"unknown case {name}")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        e = run_case(sys.argv[2])
        print(f"RESULT {sys.argv[2]} {'OK' if e < 2e-2 else 'FAIL'} {e:.3e}", flush=True)
        sys.exit(0 if e < 2e-2 else 1)
    cases = sys.argv[1:] or CASES
    summary = []
    for c in cases:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", c], timeout=300, capture_output=True, text=True)
            out = r.stdout + r.stderr
            rc = r.returncode
        except subprocess.TimeoutExpired as ex:
            out, rc = (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else str(ex.stdout), "TIMEOUT"
        print(out[-6000:], flush=True)
        summary.append((c, rc))
    print("SUMMARY", summary, flush=True)
