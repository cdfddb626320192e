"""bench.py command-line contract that can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_is_rank0_only():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm fails loudly instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    for name in ("r01e_bench_n1.json", "r01e_bench_n2.json", "r01e_bench_n4.json", "r01e_bench_n8.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches"):
            assert key in d, (name, key)
        assert d["config"]["workload"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    d = json.load(open(os.path.join(ROOT, "profiles", "r01e_bench_n1.json")))
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
