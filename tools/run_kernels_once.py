#!/usr/bin/env python3
"""Run each hot-path kernel a few times at the bench's layer-0 shape (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
bm = 8 if dt == torch.bfloat16 else 4
n, c, cm, r = 1568, 768, 64, 100
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.randn(bm, n, c, device="cuda", dtype=dt, generator=g) for _ in range(12)]
ms = [torch.randn(bm, n, cm, device="cuda", dtype=dt, generator=g) for _ in range(4)]
for i in range(6):
    nm, ni = _native.match(ms[i % 4])
    plan = _native.select(nm, ni, n, r)
    out = _native.merge(plan, xs[i], "wavg", want_size=True)
torch.cuda.synchronize()
print("ok")
