#!/usr/bin/env python3
"""One tome_frames_attention launch at Motionformer's layer-0 shape (8 clips, 8 x 196 keys, 12 heads) for ncu, plus its
graph-replayed time.  python tools/run_frames_attn_once.py [P] [bias]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import bench
from tome import _native

P = int(sys.argv[1]) if len(sys.argv) > 1 else 196
with_bias = len(sys.argv) > 2 and sys.argv[2] == "bias"
B, h, Fr = 8, 12, 8
S = Fr * P
g = torch.Generator(device="cuda").manual_seed(0)
qkvs = [torch.randn(B, 1 + S, 3 * h * 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
bias = torch.rand(B, S, device="cuda", generator=g) if with_bias else None
xs, _ = _native.frames_attention(qkvs[0], h, Fr, 0.125, bias)
torch.cuda.synchronize()
mean, med = bench.graph_time([lambda i=i: _native.frames_attention(qkvs[i % 3], h, Fr, 0.125, bias) for i in range(6)])
flop = 4.0 * B * h * S * S * 64
print(f"frames_attention P={P} bias={with_bias}: {mean:.1f} us, {flop / mean / 1e6:.0f} TFLOP/s algorithmic")
