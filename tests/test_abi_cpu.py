"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/tome_b200.h declares; the Python mirror refuses CPU tensors (no
fallback) and keeps the reference's API surface."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_path():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "tome_build", os.path.join(ROOT, "video-how-do-your-tokens-merge_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tome_b200.h")).read()
    declared = re.findall(r"TOME_API\s+[\w\s\*]+?\b(tome_\w+)\s*\(", hdr)
    assert len(declared) >= 11
    lib = ctypes.CDLL(_lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    from tome import _native
    assert sorted(declared) == sorted(_native.EXPORTS)
    lib.tome_abi_version.restype = ctypes.c_int
    assert lib.tome_abi_version() == _native.ABI_VERSION
    m = re.search(r"#define TOME_ABI_VERSION (\d+)", hdr)
    assert int(m.group(1)) == _native.ABI_VERSION


def test_ctypes_structs_match_header_layout():
    from tome import _native
    # tome_plan: 5 int32 (+4 pad) then 9 pointers; tome_view: 3 int64 + int32 (+pad)
    assert ctypes.sizeof(_native.TomePlanC) == 24 + 9 * 8
    assert _native.TomePlanC.node_max.offset == 24
    assert ctypes.sizeof(_native.TomeViewC) == 32


def test_no_cpu_fallback():
    _lib_path()
    import tome
    x = torch.randn(2, 16, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tome.merge.bipartite_soft_matching(x, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tome.merge.bipartite_soft_matching_drop(x, 4)


def test_r_clamp_returns_do_nothing_without_touching_the_device():
    import tome
    x = torch.randn(2, 1, 8)
    m, u = tome.merge.bipartite_soft_matching(x, 4)
    assert m is tome.merge.do_nothing and u is tome.merge.do_nothing
    m, u = tome.merge.bipartite_soft_matching(torch.randn(2, 16, 8), 0)
    assert m is tome.merge.do_nothing
    y, s = tome.merge.merge_wavg(m, x)
    assert torch.equal(y, x) and s.shape == (2, 1, 1) and torch.all(s == 1)
    # reference quirk kept: drop returns a tuple when r clamps to 0 (merge.py:232-233)
    assert tome.merge.bipartite_soft_matching_drop(x, 4) == (tome.merge.do_nothing, tome.merge.do_nothing)


def test_api_surface_matches_reference():
    import tome
    for name in ("bipartite_soft_matching", "kth_bipartite_soft_matching", "random_bipartite_soft_matching",
                 "bipartite_soft_matching_drop", "bipartite_soft_matching_hybrid", "merge_wavg", "merge_source",
                 "do_nothing"):
        assert callable(getattr(tome.merge, name))
    sig = inspect.signature(tome.merge.bipartite_soft_matching)
    assert list(sig.parameters) == ["metric", "r", "class_token", "distill_token", "mode"]
    sig = inspect.signature(tome.merge.bipartite_soft_matching_hybrid)
    assert list(sig.parameters) == ["metric", "r", "class_token", "distill_token", "mode", "threshold"]
    assert list(inspect.signature(tome.merge.merge_wavg).parameters) == ["merge", "x", "size"]
    assert list(inspect.signature(tome.merge.merge_source).parameters) == ["merge", "x", "source"]
    assert tome.utils.parse_r(12, (150, -1))[0] == 300
    import util
    if util.have_reference():
        ref = util.reference_merge_module()
        for name in ("bipartite_soft_matching", "bipartite_soft_matching_drop", "bipartite_soft_matching_hybrid",
                     "merge_wavg", "merge_source", "kth_bipartite_soft_matching", "random_bipartite_soft_matching"):
            assert (list(inspect.signature(getattr(ref, name)).parameters)
                    == list(inspect.signature(getattr(tome.merge, name)).parameters)), name


def test_benchmark_helper_keeps_the_reference_signature_and_protocol():
    """tome/utils.py:15-24 of the reference: same positional parameters (``use_fp16`` included), images/s for a 3-d
    input size, frames/s for a 4-d one, runs on any device; our additions are keyword-only."""
    import tome
    params = inspect.signature(tome.utils.benchmark).parameters
    assert list(params)[:8] == ["model", "device", "input_size", "batch_size", "runs", "throw_out", "use_fp16", "verbose"]
    assert all(params[k].kind is inspect.Parameter.KEYWORD_ONLY for k in list(params)[8:])
    assert params["input_size"].default == (3, 224, 224) and params["batch_size"].default == 64

    calls = []

    class Probe(torch.nn.Module):
        def forward(self, x):
            calls.append(type(x))
            return x[0].sum() if isinstance(x, list) else x.sum()

    thr = tome.utils.benchmark(Probe(), device="cpu", input_size=(3, 4, 8, 8), batch_size=2, runs=8, throw_out=0.25)
    assert thr > 0 and len(calls) == 8 and calls[0] is torch.Tensor
    tome.utils.benchmark(Probe(), device="cpu", input_size=(3, 8, 8), batch_size=2, runs=4, as_pathways=True)
    assert calls[-1] is list


def test_vivit_patch_rejects_what_the_fused_attention_cannot_honour():
    import hostmodels
    import tome
    m = hostmodels.ViViT(num_classes=4, num_frames=4, hidden_size=32, num_hidden_layers=1, num_attention_heads=2,
                         intermediate_size=64).eval()
    tome.patch.vivit(m)
    layer = m.vivit.encoder.layer[0]
    x = torch.zeros(1, 393, 32)
    with pytest.raises(NotImplementedError, match="head_mask"):
        layer(x, head_mask=torch.ones(2))
    with pytest.raises(NotImplementedError, match="output_attentions"):
        layer(x, output_attentions=True)
