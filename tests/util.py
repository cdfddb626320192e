"""Shared helpers for the parity tests (test infrastructure only)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REFERENCE = os.environ.get("TOME_REFERENCE", "/root/reference")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_mg = _load("make_golden", os.path.join(GOLDEN, "make_golden.py"))
CASES = _mg.CASES
CASE_BY_NAME = {c["name"]: c for c in CASES}
make_inputs = _mg.make_inputs


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def case_arrays(case):
    """(metric, x, size_or_None) as numpy fp32, identical to what make_golden.py fed the reference."""
    metric, x, size = make_inputs(case)
    if case.get("metric_is_x"):
        metric = x
    return metric.numpy(), x.numpy(), None if size is None else size.numpy()


def have_reference():
    return os.path.exists(os.path.join(REFERENCE, "tome", "merge.py"))


def reference_merge_module():
    return _load("ref_merge", os.path.join(REFERENCE, "tome", "merge.py"))


def assert_plan_matches_golden(plan, g, case, ulps=8.0):
    """Index lists equal the reference's, except where the reference's own margins are
    within ``ulps`` fp32 ulp (its GEMM rounding / unstable argsort decide those)."""
    from oracle import tome_oracle as O
    tol = ulps * 2.0 ** -24
    nm = g["node_max"]
    src, unm, dst = g["src_idx"], g["unm_idx"], g["dst_idx"]
    assert plan.r == int(g["r_eff"])
    assert plan.src_idx.shape == src.shape and plan.unm_idx.shape == unm.shape
    stats = dict(src_swaps=0, unm_swaps=0, dst_diffs=0)
    bm = src.shape[0]
    for b in range(bm):
        for k in range(src.shape[1]):
            ia, ib = int(plan.src_idx[b, k]), int(src[b, k])
            if ia != ib:
                stats["src_swaps"] += 1
                assert abs(float(nm[b, ia]) - float(nm[b, ib])) <= tol, (case["name"], "src", b, k)
            elif int(plan.dst_idx[b, k]) != int(dst[b, k]):
                stats["dst_diffs"] += 1
                assert float(g["top2_gap"][b, ia]) <= tol, (case["name"], "dst", b, k)
        for k in range(unm.shape[1]):
            ia, ib = int(plan.unm_idx[b, k]), int(unm[b, k])
            if ia != ib:
                stats["unm_swaps"] += 1
                if not case.get("cls"):
                    assert abs(float(nm[b, ia]) - float(nm[b, ib])) <= tol, (case["name"], "unm", b, k)
        # as sets, src/unm partition the A tokens
        assert sorted(list(plan.src_idx[b]) + list(plan.unm_idx[b])) == list(range(nm.shape[1]))
    return stats


def plans_identical(plan, g):
    return (np.array_equal(plan.src_idx, g["src_idx"]) and np.array_equal(plan.unm_idx, g["unm_idx"])
            and np.array_equal(plan.dst_idx, g["dst_idx"]))
