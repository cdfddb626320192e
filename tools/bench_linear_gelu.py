#!/usr/bin/env python3
"""tome_linear_gelu at the VideoMAE-B fc1 shape: erf GELU, HF gelu_fast, bias only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch, bench
from tome import _native
m, n, k = 8 * 1568, 3072, 768
xs = [torch.randn(m, k, device="cuda").to(torch.bfloat16) for _ in range(3)]
w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
b = torch.randn(n, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for g in (True, "gelu_fast", False):
        t, _ = bench.graph_time([lambda i=i: _native.linear_gelu(xs[i % 3], w, b, gelu=g) for i in range(3)])
        print(f"linear_gelu gelu={g}: {t:.1f} us ({2.0*m*n*k/t/1e6:.0f} TFLOP/s)")
