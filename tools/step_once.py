#!/usr/bin/env python3
"""Run the patched eager forward of ONE host model a few times (for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/step_once.py --model timesformer)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200"), os.path.join(ROOT, "tools")]
import torch
import tome
from bench_models import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="timesformer")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--plain", action="store_true", help="unpatched model")
a = ap.parse_args()
build, frames, r, kw = CONFIGS[a.model]
torch.manual_seed(0)
model = build().eval()
if a.model == "motionformer":
    torch.nn.init.trunc_normal_(model.patch_embed_3d.proj.weight, std=0.02)
    torch.nn.init.trunc_normal_(model.temp_embed, std=0.02)
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
model = model.to("cuda", dt)
clip = torch.rand(a.batch, 3, frames, 224, 224, device="cuda").to(dt)
if not a.plain:
    getattr(tome.patch, a.model)(model, **kw)
    model.r = r
with torch.no_grad():
    for _ in range(a.steps):
        torch.cuda.synchronize()
        print("STEP", flush=True)
        out = model([clip])
torch.cuda.synchronize()
print("done", tuple(out.shape))
