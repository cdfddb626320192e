"""oracle/torch_port.py (the CPU baseline that bench.py times) against the reference-made
golden vectors, the live reference when present, and the numpy oracle."""
import numpy as np
import pytest
import torch

import util
from oracle import tome_oracle as O
from oracle import torch_port as P

SMALL = [c for c in util.CASES if not c.get("large")]


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c["name"])
def test_port_reproduces_reference_outputs(case):
    g = util.golden(case["name"])
    metric, x, size = util.make_inputs(case)
    cls, dis, thr = bool(case.get("cls")), bool(case.get("distill")), case.get("hybrid")
    torch.set_num_threads(1)
    if thr is None:
        merge, unmerge = P.bipartite_soft_matching(metric, case["r"], cls, dis)
    else:
        merge, unmerge = P.bipartite_soft_matching_hybrid(metric, case["r"], cls, dis, "hybrid", thr)
    if bool(g["identity"]):
        assert merge is P.do_nothing
        return
    m = merge.match
    # same ATen calls as the reference -> identical lists, including its tie choices
    np.testing.assert_array_equal(m.src_idx[..., 0].numpy(), g["src_idx"])
    np.testing.assert_array_equal(m.unm_idx[..., 0].numpy(), g["unm_idx"])
    np.testing.assert_array_equal(m.dst_idx[..., 0].numpy(), g["dst_idx"])
    xw, sz = P.merge_wavg(merge, x, size)
    np.testing.assert_array_equal(xw.numpy(), g["x_wavg"])
    np.testing.assert_array_equal(sz.numpy(), g["size_out"])
    np.testing.assert_array_equal(merge(x, mode="mean").numpy(), g["x_mean"])
    np.testing.assert_array_equal(P.merge_source(merge, x).numpy(), g["source1"])
    np.testing.assert_array_equal(unmerge(xw).numpy(), g["x_unmerge"])
    if "x_drop" in g:
        np.testing.assert_array_equal(P.bipartite_soft_matching_drop(metric, case["r"], cls, dis)(x).numpy(), g["x_drop"])


def test_port_stable_mode_equals_numpy_oracle_on_ties():
    case = util.CASE_BY_NAME["pm1_ties"]
    metric, x, size = util.make_inputs(case)
    P.STABLE = True
    try:
        merge, _ = P.bipartite_soft_matching(metric, case["r"])
    finally:
        P.STABLE = False
    plan = O.bipartite_soft_matching(metric.numpy(), case["r"])
    np.testing.assert_array_equal(merge.match.src_idx[..., 0].numpy(), plan.src_idx)
    np.testing.assert_array_equal(merge.match.unm_idx[..., 0].numpy(), plan.unm_idx)
    np.testing.assert_array_equal(merge.match.dst_idx[..., 0].numpy(), plan.dst_idx)
