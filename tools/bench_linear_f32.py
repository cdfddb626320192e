#!/usr/bin/env python3
"""tome_linear_f32 against the library fp32 GEMM (TF32 off) at the VideoMAE-B shapes: qkv, proj, fc1 (+GELU), fc2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import bench
from tome import _native

assert not torch.backends.cuda.matmul.allow_tf32
M = 8 * 1568
for name, n, k, gelu in (("qkv", 2304, 768, False), ("proj", 768, 768, False), ("fc1+gelu", 3072, 768, True), ("fc2", 768, 3072, False)):
    xs = [torch.randn(M, k, device="cuda") for _ in range(3)]
    w = torch.randn(n, k, device="cuda") * 0.02
    b = torch.randn(n, device="cuda")
    flop = 2.0 * M * n * k
    with torch.no_grad():
        lib = (lambda i: torch.nn.functional.gelu(torch.nn.functional.linear(xs[i % 3], w, b))) if gelu else (lambda i: torch.nn.functional.linear(xs[i % 3], w, b))
        t_lib, _ = bench.graph_time([lambda i=i: lib(i) for i in range(3)])
        row = [f"library {t_lib:7.1f} us ({flop / t_lib / 1e6:5.1f} TFLOP/s)"]
        for terms in (9, 8, 6):
            t, _ = bench.graph_time([lambda i=i: _native.linear_f32(xs[i % 3], w, b, gelu=gelu, terms=terms) for i in range(3)])
            x3 = [_native.split3(x) for x in xs]
            row.append(f"x{terms} {t:7.1f} us ({flop / t / 1e6:5.1f} TFLOP/s incl. split)")
        t_split, _ = bench.graph_time([lambda i=i: _native.split3(xs[i % 3]) for i in range(3)])
        row.append(f"split alone {t_split:5.1f} us")
    print(f"{name:9s} m={M} n={n} k={k}: " + " | ".join(row), flush=True)
