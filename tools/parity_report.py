#!/usr/bin/env python3
"""Full-size parity numbers of the CUDA path against the reference-made goldens (tests/golden/fullsize.npz):
teacher-forced fp32, free-running fp32, teacher-forced and free-running bf16, per BASELINE.json architecture.
Prints one JSON line per run; `python tools/parity_report.py > profiles/rNN_parity_fullsize.jsonl`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import fullsize  # noqa: E402

only = sys.argv[1:]
for case in fullsize.FULL_CASES:
    if only and case["name"] not in only:
        continue
    for dtype in (torch.float32, torch.bfloat16):
        for forced in (True, False):
            print(json.dumps(fullsize.run_case(case, dtype, forced)), flush=True)
