#!/usr/bin/env python3
"""A few launches of linear_gelu_kernel at the bench's layer-0 MLP shape (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native

g = torch.Generator(device="cuda").manual_seed(0)
m, n, k = 8 * 1468, 3072, 768
x = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(n, k, device="cuda", generator=g) * k ** -0.5).to(torch.bfloat16)
b = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(4):
        y = _native.linear_gelu(x, w, b)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
