#!/usr/bin/env python3
"""Which torch ops are left in a patched bf16 host model forward (torch.profiler, grouped by op and shape): python tools/profile_torch_ops.py [model]."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200"), os.path.join(ROOT, "tools")]
import torch, tome
from bench_models import CONFIGS
name = sys.argv[1] if len(sys.argv) > 1 else "timesformer"
build, frames, r, kw = CONFIGS[name]
torch.manual_seed(0)
model = build().eval().to("cuda", torch.bfloat16)
clip = torch.rand(8, 3, frames, 224, 224, device="cuda").to(torch.bfloat16)
getattr(tome.patch, name)(model, **kw); model.r = r
with torch.no_grad():
    model([clip]); model([clip])
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        model([clip]); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.name in ("aten::copy_", "aten::cat", "aten::add", "aten::layer_norm", "aten::native_layer_norm", "aten::mean", "aten::fill_", "aten::zeros", "aten::pad", "aten::constant_pad_nd") and e.device_time_total > 0:
        st = [s for s in (e.stack or []) if "video-how" in s or "hostmodels" in s or "tome/" in s]
        key = (e.name, tuple(str(x) for x in (e.input_shapes or [])[:2]), st[0] if st else "?")
        agg[key][0] += 1; agg[key][1] += e.device_time_total
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:9.1f} us {v[0]:4d}x {k[0]:24s} {k[1]} {k[2][-110:]}")
