#!/usr/bin/env python3
"""tome_linear_f32 output modes at the QKV shape: fp32, planes, both, and a separate tome_split3 of the result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch, bench
from tome import _native
for m in (25096, 12544):
    xs = [torch.randn(m, 768, device="cuda") for _ in range(3)]
    w = torch.randn(2304, 768, device="cuda") * 0.03
    b = torch.randn(2304, device="cuda")
    with torch.no_grad():
        x3 = [_native.Planes(_native.split3(x), (m, 768)) for x in xs]
        t_f, _ = bench.graph_time([lambda i=i: _native.linear_f32(x3[i % 3], w, b, out="fp32") for i in range(3)])
        t_b, _ = bench.graph_time([lambda i=i: _native.linear_f32(x3[i % 3], w, b, out="both") for i in range(3)])
        t_p, _ = bench.graph_time([lambda i=i: _native.linear_f32(x3[i % 3], w, b, out="planes") for i in range(3)])
        ys = [_native.linear_f32(x3[i], w, b, out="fp32") for i in range(3)]
        t_s, _ = bench.graph_time([lambda i=i: _native.split3(ys[i % 3]) for i in range(3)])
    print(f"m={m}: qkv GEMM out=fp32 {t_f:.1f} us, out=both {t_b:.1f} us, out=planes {t_p:.1f} us, split3 of the result {t_s:.1f} us")
