"""Goldens for the two upstream-ToMe matchers the reference keeps (no caller inside it):
``kth_bipartite_soft_matching`` (tome/merge.py:105-158) and ``random_bipartite_soft_matching`` (:161-212),
made by running the UNMODIFIED reference functions on CPU.

    python tests/golden/make_sets_golden.py         # build container only (/root/reference)

The random variant draws ``torch.rand(B, N, 1)``; the generator records the draw (``rand``) and the test feeds
the same numbers to the CUDA implementation by patching ``torch.rand``, so both sides split the tokens alike."""
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference_merge  # noqa: E402

SET_CASES = [
    dict(name="kth2", kind="kth", bm=2, n=64, cm=16, c=24, k=2, seed=31),
    dict(name="kth3_ragged", kind="kth", bm=3, n=77, cm=32, c=20, k=3, seed=32),     # 77 = 25 * 3 + 2: tail ignored
    dict(name="kth7", kind="kth", bm=2, n=196, cm=64, c=32, k=7, seed=33),
    dict(name="rand_small", kind="random", bm=2, n=64, cm=16, c=24, r=10, seed=34),
    dict(name="rand_tsf", kind="random", bm=4, n=196, cm=64, c=32, r=98, seed=35),
    dict(name="rand_most", kind="random", bm=2, n=50, cm=16, c=8, r=45, seed=36),
]


def inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    metric = torch.randn(case["bm"], case["n"], case["cm"], generator=g)
    x = torch.randn(case["bm"], case["n"], case["c"], generator=g)
    rand = torch.rand(case["bm"], case["n"], 1, generator=g)
    return metric, x, rand


def main():
    ref = load_reference_merge()
    torch.set_num_threads(1)
    out = {}
    for case in SET_CASES:
        metric, x, rand = inputs(case)
        if case["kind"] == "kth":
            merge, unmerge = ref.kth_bipartite_soft_matching(metric, case["k"])
        else:
            with mock.patch.object(torch, "rand", lambda *a, **k: rand.clone()):
                merge, unmerge = ref.random_bipartite_soft_matching(metric, case["r"])
        cells = dict(zip(merge.__code__.co_freevars, (c.cell_contents for c in merge.__closure__)))
        out[case["name"] + "/dst_idx"] = cells["dst_idx"][..., 0].numpy().astype(np.int32)
        for mode in ("mean", "sum", "amax"):
            out[f"{case['name']}/merge_{mode}"] = merge(x, mode=mode).numpy()
        out[case["name"] + "/unmerge"] = unmerge(merge(x, mode="mean")).numpy()
        print(case["name"], tuple(out[case["name"] + "/merge_mean"].shape), tuple(out[case["name"] + "/unmerge"].shape))
    np.savez_compressed(os.path.join(HERE, "sets.npz"), **out)


if __name__ == "__main__":
    main()
