#!/usr/bin/env python3
"""tome_attention_f32 against torch's fp32 attention at the VideoMAE-B shapes (8 clips, 12 heads)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import bench
from tome import _native

B, h, d = 8, 12, 64
for N in (1568, 1068, 468):
    qkvs = [torch.randn(B, N, 3 * h * d, device="cuda") for _ in range(3)]
    flop = 4.0 * B * h * N * N * d
    with torch.no_grad():
        def lib(i):
            q, k, v = qkvs[i % 3].reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
            return torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125)
        t_lib, _ = bench.graph_time([lambda i=i: lib(i) for i in range(3)])
        t_own, _ = bench.graph_time([lambda i=i: _native.attention_f32(qkvs[i % 3], h, 0.125) for i in range(3)])
        t_split, _ = bench.graph_time([lambda i=i: _native.split3(qkvs[i % 3].reshape(B * N, -1)) for i in range(3)])
    print(f"N={N}: torch fp32 attention {t_lib:7.1f} us ({flop / t_lib / 1e6:5.1f} TFLOP/s) | tome_attention_f32 {t_own:7.1f} us "
          f"({flop / t_own / 1e6:5.1f} TFLOP/s incl. split {t_split:.1f} us)", flush=True)
