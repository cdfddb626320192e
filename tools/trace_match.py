#!/usr/bin/env python3
"""Phase timeline of match_tc_kernel (TOME_TC_TRACE): per-CTA %globaltimer stamps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native
bm, n, cm = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 1568, 64
g = torch.Generator(device="cuda").manual_seed(0)
ms = [torch.randn(bm, n, cm, device="cuda", generator=g) for _ in range(3)]
for m in ms: _native.match(m, algo=2)
torch.cuda.synchronize()
n_ct = _native.match_tc_describe(bm, n, cm)[0]
ncta = n_ct * (((n + 1) // 2 + 127) // 128) * bm
print("grid: %d column tiles x %d row strips x %d = %d CTAs, BN = %d" % (n_ct, ncta // n_ct // bm, bm, ncta, _native.match_tc_describe(bm, n, cm)[1]))
tr = torch.zeros(ncta * 16, dtype=torch.int64, device="cuda")
os.environ["TOME_TC_TRACE"] = str(tr.data_ptr())
_native.match(ms[1], algo=2)
torch.cuda.synchronize()
del os.environ["TOME_TC_TRACE"]
t = tr.view(ncta, 16).cpu().numpy().astype("float64")
t0 = t[:, 0].min()
names = ["start", "setup done", "stage0 landed", "last stage landed", "mma issued", "acc ready", "pass1", "pass2", "exact done", "all synced", "end"]
import numpy as np
print("kernel span (first start -> last end): %.2f us" % ((t[:, 10].max() - t0) / 1e3))
order = np.argsort(t[:, 0])
print("CTA start times (us) pct 0/50/90/100:", np.percentile((t[:, 0] - t0) / 1e3, [0, 50, 90, 100]).round(2))
for k in range(1, 11):
    d = (t[:, k] - t[:, k - 1]) / 1e3
    print("%-18s median %.2f  p90 %.2f  max %.2f us" % (names[k], np.median(d), np.percentile(d, 90), d.max()))
print("pass2 bit phase (6->11) median %.2f us, extraction (11->7) median %.2f us" % (np.median((t[:, 11] - t[:, 6]) / 1e3), np.median((t[:, 7] - t[:, 11]) / 1e3)))
print("per-CTA total median %.2f max %.2f" % (np.median((t[:, 10] - t[:, 0]) / 1e3), ((t[:, 10] - t[:, 0]) / 1e3).max()))
