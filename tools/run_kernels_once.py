#!/usr/bin/env python3
"""Run each hot-path kernel a few times at the bench's layer-0 shape, the way the patched block calls them
(for ncu captures): tome_plan_build on the lazy head-mean of K, then the merge with the residual add
and the LayerNorm fused in; plus the plain merge_wavg the roofline line is quoted on."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
bm = 8 if dt == torch.bfloat16 else 4
n, c, cm, r, heads = 1568, 768, 64, 100, 12
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.randn(bm, n, c, device="cuda", dtype=dt, generator=g) for _ in range(12)]
rs = [torch.randn(bm, n, c, device="cuda", dtype=dt, generator=g) for _ in range(12)]
ks = [torch.randn(bm, n, 3, heads, cm, device="cuda", dtype=dt, generator=g).permute(2, 0, 3, 1, 4)[1] for _ in range(4)]
w = torch.ones(c, device="cuda", dtype=dt)
b = torch.zeros(c, device="cuda", dtype=dt)
for i in range(6):
    plan = _native.plan_build(_native.HeadMeanMetric(ks[i % 4]), r)
    out = _native.merge(plan, xs[i], "wavg", want_size=True)
    out2 = _native.merge(plan, xs[i + 6], "wavg", want_size=True, norm=(w, b, 1e-6), residual=rs[i])
# the MLP's first half (fc1 + bias + erf GELU, one tcgen05 GEMM) on the merged tokens
xm = torch.randn(bm * (n - r), c, device="cuda", dtype=dt, generator=g)
w1 = (torch.randn(4 * c, c, device="cuda", generator=g) * c ** -0.5).to(dt)
b1 = torch.zeros(4 * c, device="cuda", dtype=dt)
if dt == torch.bfloat16:
    for i in range(3):
        h = _native.linear_gelu(xm, w1, b1)
torch.cuda.synchronize()
print("ok")
