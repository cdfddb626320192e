#!/usr/bin/env python3
"""tome_attention_bf16 against the library's bf16 attention at the VideoMAE-B / ViViT-B shapes (8 clips, 12 heads):
plain, and with the proportional-attention key bias (library: the padded-head fold of tome/attention.py, and the masked call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import torch.nn.functional as F
import bench
from tome import _native

B, h, d = 8, 12, 64
for N in (3137, 1568, 1068, 468):
    qkvs = [torch.randn(B, N, 3 * h * d, device="cuda").to(torch.bfloat16) for _ in range(3)]
    bias = torch.randint(1, 8, (B, N), device="cuda").float().log()
    qp = [torch.randn(B, h, N, d + 8, device="cuda").to(torch.bfloat16) for _ in range(3)]      # padded heads of the round-1 fold
    vp = [torch.randn(B, h, N, d, device="cuda").to(torch.bfloat16) for _ in range(3)]
    flop = 4.0 * B * h * N * N * d
    with torch.no_grad():
        def lib(i):
            q, k, v = qkvs[i % 3].view(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
            return F.scaled_dot_product_attention(q, k, v, scale=0.125)
        def lib_pad(i):
            return F.scaled_dot_product_attention(qp[i % 3], qp[(i + 1) % 3], vp[i % 3], scale=0.125)
        t_lib, _ = bench.graph_time([lambda i=i: lib(i) for i in range(3)])
        t_pad, _ = bench.graph_time([lambda i=i: lib_pad(i) for i in range(3)])
        t_own, _ = bench.graph_time([lambda i=i: _native.attention_bf16(qkvs[i % 3], h, 0.125) for i in range(3)])
        t_ownb, _ = bench.graph_time([lambda i=i: _native.attention_bf16(qkvs[i % 3], h, 0.125, bias) for i in range(3)])
    print(f"N={N}: library plain {t_lib:7.1f} us ({flop / t_lib / 1e6:6.1f} TFLOP/s) | library, bias folded into padded heads {t_pad:7.1f} us | "
          f"tome_attention_bf16 plain {t_own:7.1f} us ({flop / t_own / 1e6:6.1f} TFLOP/s), with key bias {t_ownb:7.1f} us", flush=True)
