#!/usr/bin/env python3
"""The fused merge launch of bench.py's roofline (MergeBench: inputs and outputs rotating through > L2) between
cudaProfilerStart / cudaProfilerStop, for ncu:

  per-kernel counters (cold cache per replay pass; writes may still sit in L2 when the kernel ends):
    ncu --set full --clock-control none --import-source on -k regex:merge_gather -c 3 -o gpurun_out/r02_merge python tools/merge_range.py
  steady-state DRAM traffic of the whole rotating sequence (evictions of earlier launches' outputs included):
    ncu --replay-mode range --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --csv --log-file gpurun_out/r02_merge_range_bf16.csv python tools/merge_range.py --dtype bf16
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
ap.add_argument("--n", type=int, default=1568)
ap.add_argument("--r", type=int, default=100)
ap.add_argument("--bm", type=int, default=8)
ap.add_argument("--rounds", type=int, default=2, help="passes over the rotating buffers inside the profiled range")
ap.add_argument("--plain", action="store_true", help="merge_wavg only (no residual, no LayerNorm)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
mb = bench.MergeBench(dev, dtype, a.bm, a.n, a.r, fused=not a.plain)
for i in range(mb.nrot):                      # warm-up: every buffer touched once, outputs of the last launches in L2
    mb.launch(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.rounds):
    for i in range(mb.nrot):
        mb.launch(i)
torch.cuda.cudart().cudaProfilerStop()
torch.cuda.synchronize()
print(f"launches_in_range={a.rounds * mb.nrot} algorithmic_bytes_per_launch={mb.bytes} bm={a.bm} n={a.n} r={a.r} dtype={a.dtype}")
