#!/usr/bin/env python3
"""Per-key-block timeline of one CTA of attn_bf16_kernel (clock64 stamps).  Needs a library built with -DTOME_ATTN_TRACE
(NVCC_EXTRA=-DTOME_ATTN_TRACE python video-how-do-your-tokens-merge_b200/build.py --force)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native
N = 1568
qkv = torch.randn(8, N, 3 * 768, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        out = _native.attention_bf16(qkv, 12, 0.125)
torch.cuda.synchronize()
lib = _native.load_library()
buf = (ctypes.c_longlong * 320)()
lib.tome_debug_attn_bf16_trace.argtypes = [ctypes.c_void_p]
print("rc", lib.tome_debug_attn_bf16_trace(buf))
nb = (N + 127) // 128
ev = [[buf[e * 32 + j] for j in range(nb)] for e in range(10)]
t0 = min(x for e in ev for x in e if x > 0)
names = ["S issue start", "S issue end", "PV issue start", "PV issue end", "softmax got S", "S in registers", "row max exchanged",
         "exps done", "P arrived", "iteration start"]
order = [0, 1, 2, 3, 9, 4, 5, 6, 7, 8]
print(" j " + " ".join(f"{names[e][:9]:>9s}" for e in order))
for j in range(nb):
    print(f"{j:2d} " + " ".join(f"{(ev[e][j] - t0) if ev[e][j] else -1:9d}" for e in order))
