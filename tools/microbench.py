#!/usr/bin/env python3
"""Kernel-level timing of the hot path (graph-replayed, so host launch overhead is out).

    python tools/microbench.py [--bm 8] [--n 1568] [--c 768] [--cm 64] [--r 100] [--dtype bf16]

Each kernel is captured `nrot` times in one CUDA graph over rotating buffers larger than L2
(cold HBM reads, like the first touch in a real forward) and replayed; time per launch =
replay time / nrot.  Also reports a "warm" figure (single buffer, L2-resident input)."""
import argparse, json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native

L2 = 126 * 2 ** 20


def graph_time(fns, reps=20):
    """fns: list of zero-arg callables (one launch group each). Returns us per callable."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        st.record(); g.replay(); en.record(); torch.cuda.synchronize()
        ts.append(st.elapsed_time(en) * 1e3 / len(fns))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bm", type=int, default=8); ap.add_argument("--n", type=int, default=1568)
    ap.add_argument("--c", type=int, default=768); ap.add_argument("--cm", type=int, default=64)
    ap.add_argument("--r", type=int, default=100); ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--cls", type=int, default=0); ap.add_argument("--sized", type=int, default=0)
    a = ap.parse_args()
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    e = 2 if dt == torch.bfloat16 else 4
    dev = torch.device("cuda")
    _native.device_check()
    bm, n, c, cm, r = a.bm, a.n, a.c, a.cm, a.r
    nrot = max(2, math.ceil(1.5 * L2 / (bm * n * c * e)))
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(bm, n, c, device=dev, dtype=dt, generator=g) for _ in range(nrot)]
    ms = [torch.randn(bm, n, cm, device=dev, dtype=dt, generator=g) for _ in range(nrot)]
    size = torch.randint(1, 4, (bm, n, 1), device=dev, generator=g).float() if a.sized else None
    nm, ni = _native.match(ms[0], bool(a.cls))
    plan = _native.select(nm, ni, n, r, bool(a.cls))
    na = (n + 1) // 2
    out = {"shape": dict(bm=bm, n=n, c=c, cm=cm, r=r, dtype=a.dtype, nrot=nrot)}
    peak = 6446.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    alg = bm * (n * c * e + n * 4 + (n - r) * c * e + (n - r) * 8 + na * 12)
    med, best = graph_time([lambda i=i: _native.merge(plan, xs[i], "wavg", size=size, want_size=True) for i in range(nrot)])
    out["merge_wavg_cold"] = dict(us=med, us_best=best, GBps=alg / med / 1e3, frac=alg / med / 1e3 / peak, alg_bytes=alg)
    med, best = graph_time([lambda: _native.merge(plan, xs[0], "wavg", size=size, want_size=True) for _ in range(8)])
    out["merge_wavg_warm"] = dict(us=med, us_best=best, GBps=alg / med / 1e3)
    # the block-level fusions around the merge: LayerNorm after, residual add before
    w = torch.ones(c, device=dev, dtype=dt); bb = torch.zeros(c, device=dev, dtype=dt)
    rs = [torch.randn(bm, n, c, device=dev, dtype=dt, generator=g) for _ in range(nrot)]
    med, best = graph_time([lambda i=i: _native.merge(plan, xs[i], "wavg", size=size, want_size=True, norm=(w, bb, 1e-6)) for i in range(nrot)])
    out["merge_wavg_norm_cold"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.merge(plan, xs[i], "wavg", size=size, want_size=True, norm=(w, bb, 1e-6), residual=rs[i]) for i in range(nrot)])
    out["merge_wavg_add_norm_cold"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.merge(plan, xs[i] + rs[i], "wavg", size=size, want_size=True, norm=(w, bb, 1e-6)) for i in range(nrot)])
    out["torch_add_then_merge_wavg_norm_cold"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.add_layernorm(xs[i], rs[i], (w, bb, 1e-6)) for i in range(nrot)])
    out["add_layernorm_cold"] = dict(us=med, us_best=best, GBps=3 * bm * n * c * e / med / 1e3)
    if dt == torch.bfloat16:      # the MLP's first half: library GEMM + GELU pass vs the fused tcgen05 kernel
        for toks in (n, 1068, 768, 468):
            xm = torch.randn(bm * toks, c, device=dev, dtype=dt, generator=g)
            w1 = (torch.randn(4 * c, c, device=dev, generator=g) * c ** -0.5).to(dt)
            b1 = torch.zeros(4 * c, device=dev, dtype=dt)
            fl = 2.0 * bm * toks * c * 4 * c
            med, best = graph_time([lambda: torch.nn.functional.gelu(torch.nn.functional.linear(xm, w1, b1)) for _ in range(4)])
            out[f"torch_linear_then_gelu_{toks}"] = dict(us=med, us_best=best, tflops=fl / med / 1e6)
            med, best = graph_time([lambda: torch.nn.functional.linear(xm, w1, b1) for _ in range(4)])
            out[f"torch_linear_{toks}"] = dict(us=med, us_best=best, tflops=fl / med / 1e6)
            med, best = graph_time([lambda: _native.linear_gelu(xm, w1, b1) for _ in range(4)])
            out[f"linear_gelu_fused_{toks}"] = dict(us=med, us_best=best, tflops=fl / med / 1e6)
            med, best = graph_time([lambda: _native.linear_gelu(xm, w1, b1, gelu=False) for _ in range(4)])
            out[f"linear_fused_nogelu_{toks}"] = dict(us=med, us_best=best, tflops=fl / med / 1e6)
    flops = 2.0 * bm * na * (n // 2) * cm
    for algo, name in ((2, "match_tc"), (1, "match_exact")):
        med, best = graph_time([lambda i=i: _native.match(ms[i % nrot], bool(a.cls), algo=algo) for i in range(8)])
        out[name] = dict(us=med, us_best=best, tflops_alg=flops / med / 1e6)
    ks = [torch.randn(bm, n, 3, 12, cm, device=dev, dtype=dt, generator=g).permute(2, 0, 3, 1, 4)[1] for _ in range(4)]
    med, best = graph_time([lambda i=i: _native.match_heads(_native.HeadMeanMetric(ks[i % 4]), bool(a.cls)) for i in range(8)])
    out["match_heads12"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.plan_build(_native.HeadMeanMetric(ks[i % 4]), r, bool(a.cls)) for i in range(8)])
    out["plan_build_heads12"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.plan_build(ms[i % nrot], r, bool(a.cls)) for i in range(8)])
    out["plan_build"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda i=i: _native.match(ks[i % 4].mean(1), bool(a.cls)) for i in range(8)])
    out["torch_mean_then_match"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda: _native.select(nm, ni, n, r, bool(a.cls)) for _ in range(8)])
    out["select"] = dict(us=med, us_best=best)
    med, best = graph_time([lambda: _native.merge_source(plan, None) for _ in range(4)])
    out["merge_source_identity"] = dict(us=med, GBps=bm * (n - r) * n * 4 / med / 1e3)
    # torch reference points on the same device
    med, best = graph_time([lambda i=i: xs[i].clone() for i in range(nrot)])
    out["torch_clone_same_bytes"] = dict(us=med, GBps=2 * bm * n * c * e / med / 1e3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
