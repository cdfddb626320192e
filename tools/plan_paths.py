#!/usr/bin/env python3
"""Time tome_plan_build (kernels 1 + 2) per path at the shapes of the four models: the multi-launch chain, the chain with
the one-launch select (threshold published through global memory), and the one-launch cluster plan kernel.  CUDA-graph replay, rotating inputs.
    python tools/plan_paths.py > profiles/r02_plan_paths.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import bench
from tome import _native

SHAPES = [  # label, bm, n, r, class token
    ("VideoMAE-B layer 0 (8 clips)", 8, 1568, 100, False),
    ("VideoMAE-B layer 11", 8, 468, 100, False),
    ("ViViT-B layer 0 (8 clips)", 8, 3137, 300, True),
    ("ViViT-B layer 9", 8, 437, 218, True),
    ("TimeSformer / Motionformer layer 0 (8 clips x 8 frames)", 64, 196, 18, False),
    ("TimeSformer layer 6", 64, 88, 18, False),
    ("TimeSformer layer 10", 64, 17, 8, False),
]
PATHS = [("chain (split + match_tc + rank + finish)", dict(TOME_PLAN_CLUSTER="0", TOME_SELECT_TWO="1")),
         ("chain with one-launch select (split + match_tc + select_one)", dict(TOME_PLAN_CLUSTER="0", TOME_SELECT_TWO="0")),
         ("one-launch plan kernel", dict(TOME_PLAN_CLUSTER="1", TOME_SELECT_TWO="0"))]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for label, bm, n, r, cls in SHAPES:
    ks = [torch.randn(bm, n, 3, 12, 64, device=dev, generator=g).to(torch.bfloat16).permute(2, 0, 3, 1, 4)[1] for _ in range(4)]
    row = []
    for name, env in PATHS:
        os.environ.update(env)
        before = _native.launch_count()
        _native.plan_build(_native.HeadMeanMetric(ks[0]), r, cls)
        launches = _native.launch_count() - before
        mean, med = bench.graph_time([lambda i=i: _native.plan_build(_native.HeadMeanMetric(ks[i % 4]), r, cls) for i in range(8)])
        row.append(f"{name}: {mean:6.2f} us ({launches} launches)")
    print(f"{label:58s} bm={bm:3d} n={n:5d} r={r:4d}  |  " + "  |  ".join(row), flush=True)
    nm, ni = _native.match(ks[0].float().mean(1))
    for name, env in PATHS[:2]:
        os.environ.update(env)
        mean, _ = bench.graph_time([lambda: _native.select(nm, ni, n, r, cls) for _ in range(8)])
        print(f"    select only, {name}: {mean:6.2f} us")
