#!/usr/bin/env python3
"""Per-key-block timeline of one CTA of attn_f32_ts_kernel (clock64 stamps): when the S / P V batches are issued, when the
softmax warps get S, finish the exponentials, get O and hand P over.  Needs a library built with -DTOME_ATTN_TRACE
(NVCC_EXTRA=-DTOME_ATTN_TRACE python video-how-do-your-tokens-merge_b200/build.py --force)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native
N = 1568
qkv = torch.randn(8, N, 3 * 768, device="cuda")
with torch.no_grad():
    for _ in range(3):
        out = _native.attention_f32(qkv, 12, 0.125)
torch.cuda.synchronize()
lib = _native.load_library()
buf = (ctypes.c_longlong * 256)()
lib.tome_debug_attn_trace.argtypes = [ctypes.c_void_p]
print("rc", lib.tome_debug_attn_trace(buf))
ev = [[buf[e * 32 + j] for j in range(25)] for e in range(8)]
t0 = min(x for e in ev for x in e if x > 0)
names = ["S issue start", "S issue end", "PV issue start", "PV issue end", "softmax got S", "softmax exps done", "softmax got O", "softmax P arrived"]
for j in range(25):
    print(j, " ".join(f"{(ev[e][j] - t0) if ev[e][j] else -1:7d}" for e in range(8)))
print(names)
