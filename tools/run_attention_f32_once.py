#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1568
qkv = torch.randn(8, N, 3 * 768, device="cuda")
with torch.no_grad():
    for _ in range(2):
        out = _native.attention_f32(qkv, 12, 0.125)
torch.cuda.synchronize()
print("done")
