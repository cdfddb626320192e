#!/usr/bin/env python3
"""Phase timeline of plan_cluster_kernel (TOME_PC_TRACE): per-CTA %globaltimer stamps.
    python tools/trace_plan.py [n] [bm] [dtype]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import numpy as np
import torch
from tome import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1568
bm = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dtype = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "fp32") else torch.bfloat16
r = min(100, n // 2)
g = torch.Generator(device="cuda").manual_seed(0)
ks = [torch.randn(bm, n, 3, 12, 64, device="cuda", generator=g).to(dtype).permute(2, 0, 3, 1, 4)[1] for _ in range(3)]
for k in ks:
    _native.plan_build(_native.HeadMeanMetric(k), r)
torch.cuda.synchronize()
cs = _native.plan_cluster_describe(bm, n)
ncta = cs[0] * bm
print("cluster geometry (CS, RA, RB, BN, smem):", cs, "->", ncta, "CTAs")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cold in (True, False):
    tr = torch.zeros(ncta * 16 + 256, dtype=torch.int64, device="cuda")
    if cold:
        flush.zero_()
    torch.cuda.synchronize()
    os.environ["TOME_PC_TRACE"] = str(tr.data_ptr())
    _native.plan_build(_native.HeadMeanMetric(ks[1]), r)
    torch.cuda.synchronize()
    del os.environ["TOME_PC_TRACE"]
    full = tr.cpu().numpy().astype("float64")
    t = full[:ncta * 16].reshape(ncta, 16)
    tiles = full[ncta * 16:ncta * 16 + 128].reshape(16, 8)
    t0 = t[:, 0].min()
    names = ["start", "prep done", "barrier 1 passed", "first accumulator ready", "sweep done", "refine done", "barrier 2 passed",
             "rank done", "barrier 3 passed", "csr done"]
    print("--- %s L2 --- kernel span (first start -> last end): %.2f us; CTA starts p0/50/100: %s" % (
        "cold" if cold else "warm", (t[:, 9].max() - t0) / 1e3, np.percentile((t[:, 0] - t0) / 1e3, [0, 50, 100]).round(2)))
    for k in range(1, 10):
        d = (t[:, k] - t[:, k - 1]) / 1e3
        print("  %-24s median %.2f  p90 %.2f  max %.2f us" % (names[k], np.median(d), np.percentile(d, 90), d.max()))
    print("  setup (barriers, TMEM alloc, tables) before the prep loop: median %.2f us" % np.median((t[:, 12] - t[:, 0]) / 1e3))
    b1 = t[0, 2]
    print("  CTA (0,0) per tile, us after barrier 1: [MMA: accumulator free | operands landed | issued]  [worker: waiting | accumulator ready | tile done]")
    for it in range(16):
        if tiles[it, 2] == 0:
            break
        print("    tile %2d  " % it + "  ".join("%6.2f" % ((tiles[it, k] - b1) / 1e3) for k in range(6)))
    print("  MMA: first tile committed at +%.2f us after barrier 1, all issued at +%.2f (median)" % (
        np.median((t[:, 10] - t[:, 2]) / 1e3), np.median((t[:, 11] - t[:, 2]) / 1e3)))
