"""Generate golden vectors by running the UNMODIFIED reference ``tome/merge.py``.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files are the pin
for ``oracle/tome_oracle.py`` and for the CUDA path.  The reference module is imported by
file path (it depends only on torch); nothing from it is copied into this repository.
Each ``<case>.npz`` holds the seeded inputs (small cases) or the seed (large cases) and
what the reference produced: index lists read out of the closures, node_max, merged
features / sizes / source maps, unmerge output.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TOME_REFERENCE", "/root/reference")


def load_reference_merge():
    spec = importlib.util.spec_from_file_location("ref_merge", os.path.join(REF, "tome", "merge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(case):
    """Seeded inputs; shared with tests/util.py::case_inputs (keep in sync)."""
    g = torch.Generator().manual_seed(case["seed"])
    bm, n, cm, c = case["bm"], case["n"], case["cm"], case["c"]
    if case["dist"] == "gauss":
        metric = torch.randn(bm, n, cm, generator=g)
    elif case["dist"] == "pm1":          # exact arithmetic: every normalised product is dyadic
        metric = (torch.randint(0, 2, (bm, n, cm), generator=g) * 2 - 1).float()
    elif case["dist"] == "tokens":       # correlated "video-like" tokens: strong neighbours
        base = torch.randn(bm, n // 7 + 1, cm, generator=g)
        metric = base.repeat_interleave(7, dim=1)[:, :n] + 0.35 * torch.randn(bm, n, cm, generator=g)
    else:
        raise ValueError(case["dist"])
    x = torch.randn(bm, n, c, generator=g)
    size = None
    if case.get("with_size"):
        size = torch.randint(1, 8, (bm, n, 1), generator=g).float()
    return metric, x, size


CASES = [
    # name, shape, options.  "small" cases store full tensors; large ones store subsamples.
    dict(name="gauss_small", bm=2, n=64, cm=16, c=24, r=10, seed=1, dist="gauss"),
    dict(name="gauss_odd", bm=3, n=77, cm=32, c=20, r=17, seed=2, dist="gauss", with_size=True),
    dict(name="gauss_cls", bm=2, n=197, cm=64, c=32, r=40, seed=3, dist="gauss", cls=True, with_size=True),
    dict(name="gauss_cls_distill", bm=2, n=198, cm=64, c=16, r=30, seed=4, dist="gauss", cls=True, distill=True),
    dict(name="gauss_rclamp", bm=2, n=50, cm=16, c=8, r=1000, seed=5, dist="gauss", with_size=True),
    dict(name="gauss_rclamp_cls", bm=2, n=51, cm=16, c=8, r=1000, seed=6, dist="gauss", cls=True),
    dict(name="tokens_tsf", bm=8, n=196, cm=64, c=32, r=18, seed=7, dist="tokens", with_size=True),
    dict(name="tokens_hybrid", bm=2, n=200, cm=64, c=24, r=60, seed=8, dist="tokens", with_size=True,
         hybrid=0.8),
    dict(name="tokens_hybrid_cls", bm=2, n=201, cm=64, c=24, r=60, seed=9, dist="tokens", cls=True,
         hybrid=0.4),
    dict(name="pm1_ties", bm=2, n=96, cm=64, c=16, r=20, seed=10, dist="pm1", with_size=True),
    dict(name="pm1_ties_cls", bm=2, n=97, cm=256, c=16, r=24, seed=11, dist="pm1", cls=True),
    dict(name="r_zero", bm=2, n=32, cm=16, c=8, r=0, seed=12, dist="gauss"),
    dict(name="n_one", bm=2, n=1, cm=16, c=8, r=4, seed=13, dist="gauss"),
    # BASELINE.json configs[0] shapes (SURVEY.md 8d "Config 1"): M1' (Cm=64) and M1 (metric = x)
    dict(name="config1_m1p", bm=4, n=1568, cm=64, c=768, r=100, seed=0, dist="gauss", large=True),
    dict(name="config1_m1p_size", bm=4, n=1568, cm=64, c=768, r=100, seed=20, dist="gauss", large=True,
         with_size=True),
    dict(name="config1_m1", bm=4, n=1568, cm=768, c=768, r=100, seed=21, dist="gauss", large=True,
         metric_is_x=True),
    dict(name="vivit_layer0", bm=2, n=3137, cm=64, c=64, r=300, seed=22, dist="tokens", large=True, cls=True),
]


def closure_vars(fn):
    return dict(zip(fn.__code__.co_freevars, (c.cell_contents for c in fn.__closure__)))


def run_case(ref, case):
    metric, x, size = make_inputs(case)
    if case.get("metric_is_x"):
        metric = x
    cls, distill = bool(case.get("cls")), bool(case.get("distill"))
    out = {}
    small = not case.get("large")
    if small:
        out["metric"] = metric.numpy()
        out["x"] = x.numpy()
        if size is not None:
            out["size"] = size.numpy()

    hyb = case.get("hybrid")
    # node_max / node_idx the way merge.py:51-64 computes them (needed by the gap-aware
    # comparator).  The hybrid closure exposes node_max itself; use it to cross-check.
    with torch.no_grad():
        mn = metric / metric.norm(dim=-1, keepdim=True)
        sc = mn[..., ::2, :] @ mn[..., 1::2, :].transpose(-1, -2)
        if cls:
            sc[..., 0, :] = -float("inf")
        if distill:
            sc[..., :, 0] = -float("inf")
        if sc.shape[-1] > 0 and sc.shape[-2] > 0:
            node_max, node_idx = sc.max(dim=-1)
            top2 = sc.topk(min(2, sc.shape[-1]), dim=-1).values
            out["node_max"] = node_max.numpy()
            out["node_idx"] = node_idx.numpy().astype(np.int32)
            out["top2_gap"] = (top2[..., 0] - top2[..., -1]).numpy()

    if hyb is not None:
        merge, unmerge = ref.bipartite_soft_matching_hybrid(metric, case["r"], cls, distill, "hybrid", hyb)
    else:
        merge, unmerge = ref.bipartite_soft_matching(metric, case["r"], cls, distill)
    identity = merge is ref.do_nothing
    out["identity"] = np.array(identity)
    if not identity:
        cv = closure_vars(merge)
        out["r_eff"] = np.array(cv["r"])
        for k in ("src_idx", "unm_idx", "dst_idx"):
            out[k] = cv[k][..., 0].numpy().astype(np.int32)
        if hyb is not None:
            assert torch.equal(cv["node_max"], node_max)

    xm, sz = ref.merge_wavg(merge, x, size)
    out["size_out"] = sz.numpy()
    if small:
        out["x_wavg"] = xm.numpy()
        out["x_mean"] = merge(x, mode="mean").numpy()
        out["x_amax"] = merge(x, mode="max").numpy()
        if x.shape[1] <= 256:
            src1 = ref.merge_source(merge, x, None)
            out["source1"] = src1.numpy()
        if not identity:
            out["x_unmerge"] = unmerge(xm).numpy()
            if hyb is None:
                drop = ref.bipartite_soft_matching_drop(metric, case["r"], cls, distill)
                dv = closure_vars(drop)
                assert torch.equal(dv["und_idx"], closure_vars(merge)["unm_idx"])
                out["x_drop"] = drop(x).numpy()
    else:
        out["x_wavg_sub"] = xm[:, ::37, ::5].numpy()
        out["x_wavg_sum"] = xm.double().sum(dim=(1, 2)).numpy()
    return out


def main():
    ref = load_reference_merge()
    torch.set_num_threads(1)   # deterministic MKL blocking for the recorded node_max
    for case in CASES:
        out = run_case(ref, case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']:24s} {os.path.getsize(path) / 1024:8.1f} KiB  keys={sorted(out)}")


if __name__ == "__main__":
    sys.exit(main())
