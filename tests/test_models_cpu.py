"""Host models + tome.patch wiring against the UNMODIFIED reference models and patches (CPU, fp32).

The reference side runs from /root/reference through tests/refshim.py (only in the build container;
skipped where the reference is absent).  Our side runs hostmodels + our tome.patch with the merge
calls routed to oracle/torch_port.py -- the CUDA kernels cannot run here, and this test is about the
HOST logic: module matching, r schedules, size / prop-attn plumbing, token layouts, state-dict
compatibility.  Same weights, same clip -> logits must agree to fp32 round-off."""
import contextlib
import sys

import pytest
import torch

import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@contextlib.contextmanager
def port_backend(*module_names):
    """Route the patches' merge calls to the torch-CPU port (test-only)."""
    from oracle import torch_port as P
    names = ("bipartite_soft_matching", "bipartite_soft_matching_drop", "bipartite_soft_matching_hybrid",
             "merge_wavg", "merge_source")
    saved = []
    try:
        for mn in module_names:
            mod = sys.modules[mn]
            for n in names:
                if hasattr(mod, n):
                    saved.append((mod, n, getattr(mod, n)))
                    setattr(mod, n, getattr(P, n))
        yield
    finally:
        for mod, n, f in saved:
            setattr(mod, n, f)


def _reference_logits(builder, patch_name, clip, r, patch_kwargs, seed, **build_kwargs):
    with refshim.reference_modules():
        import tome as ref_tome
        torch.manual_seed(seed)
        ref = builder(**build_kwargs).eval()
        sd = {k: v.clone() for k, v in ref.state_dict().items()}
        with torch.no_grad():
            plain = ref([clip]).clone()
        getattr(ref_tome.patch, patch_name)(ref, **patch_kwargs)
        ref.r = r
        with torch.no_grad():
            merged = ref([clip]).clone()
        info = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ref._tome_info.items() if k in ("size", "source")}
    return sd, plain, merged, info


CASES = [
    # name, reference builder, our model factory, patch name, clip frames, r, patch kwargs
    ("videomae_merge", "build_reference_videomae", dict(num_classes=11, num_frames=4, depth=3), "videomae", 4, 40, dict()),
    ("videomae_propattn_source", "build_reference_videomae", dict(num_classes=11, num_frames=4, depth=3), "videomae", 4, (40, -0.5),
     dict(prop_attn=True, trace_source=True)),
    ("videomae_hybrid", "build_reference_videomae", dict(num_classes=11, num_frames=4, depth=3), "videomae", 4, 40,
     dict(mode="hybrid", threshold=0.9, prop_attn=True)),
    ("videomae_drop", "build_reference_videomae", dict(num_classes=11, num_frames=4, depth=3), "videomae", 4, 40, dict(mode="drop")),
    ("timesformer_merge", "build_reference_timesformer", dict(num_classes=11, num_frames=4, depth=3), "timesformer", 4, 18, dict()),
    ("timesformer_drop_source", "build_reference_timesformer", dict(num_classes=11, num_frames=4, depth=3), "timesformer", 4, 18,
     dict(mode="drop", trace_source=True)),
    ("motionformer_merge", "build_reference_motionformer", dict(num_classes=11, num_frames=8, depth=3), "motionformer", 8, 18, dict()),
    ("motionformer_hybrid", "build_reference_motionformer", dict(num_classes=11, num_frames=8, depth=3), "motionformer", 8, [18, 30, 10],
     dict(mode="hybrid", threshold=0.7)),
]


def _our_model(name, kwargs):
    import hostmodels
    if name.startswith("videomae"):
        from hostmodels.videomae import VideoMAE, VisionTransformer
        from functools import partial
        m = VideoMAE(num_classes=kwargs["num_classes"], num_frames=kwargs["num_frames"])
        m.model = VisionTransformer(patch_size=16, embed_dim=768, depth=kwargs["depth"], num_heads=12, mlp_ratio=4, qkv_bias=True,
                                    norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=kwargs["num_classes"],
                                    all_frames=kwargs["num_frames"], tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
        return m
    if name.startswith("timesformer"):
        from hostmodels.timesformer import TimeSformer, VisionTransformer
        from functools import partial
        m = TimeSformer(num_classes=kwargs["num_classes"], num_frames=kwargs["num_frames"])
        m.model = VisionTransformer(img_size=224, num_classes=kwargs["num_classes"], patch_size=16, embed_dim=768,
                                    depth=kwargs["depth"], num_heads=12, mlp_ratio=4, qkv_bias=True,
                                    norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=kwargs["num_frames"])
        return m
    if name.startswith("motionformer"):
        return hostmodels.Motionformer(num_classes=kwargs["num_classes"], num_frames=kwargs["num_frames"], depth=kwargs["depth"])
    raise KeyError(name)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_patched_host_model_matches_reference(case):
    name, builder, bkw, patch_name, frames, r, pkw = case
    torch.set_num_threads(4)
    g = torch.Generator().manual_seed(7)
    clip = torch.rand(2, 3, frames, 224, 224, generator=g)
    sd, ref_plain, ref_merged, ref_info = _reference_logits(getattr(refshim, builder), patch_name, clip, r, pkw, seed=3, **bkw)
    if name.startswith("motionformer"):
        # the reference zero-initialises these two (identical frames -> all-tie matching); use real values
        g2 = torch.Generator().manual_seed(11)
        for key in ("patch_embed_3d.proj.weight", "temp_embed"):
            sd[key] = torch.nn.init.trunc_normal_(torch.empty_like(sd[key]), std=0.02, generator=g2)
        with refshim.reference_modules():
            import tome as ref_tome
            torch.manual_seed(3)
            ref = getattr(refshim, builder)(**bkw).eval()
            ref.load_state_dict(sd)
            with torch.no_grad():
                ref_plain = ref([clip]).clone()
            ref_tome.patch.motionformer(ref, **pkw)
            ref.r = r
            with torch.no_grad():
                ref_merged = ref([clip]).clone()
            ref_info = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ref._tome_info.items() if k in ("size", "source")}

    import tome
    ours = _our_model(name, bkw).eval()
    missing, unexpected = ours.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)          # state dicts interchange
    with torch.no_grad():
        plain = ours([clip])
    torch.testing.assert_close(plain, ref_plain, rtol=2e-4, atol=2e-5)      # the host model itself
    getattr(tome.patch, patch_name)(ours, **pkw)
    ours.r = r
    with port_backend("tome.patch.videomae", "tome.patch.timesformer", "tome.patch.motionformer", "tome.patch.vivit"), \
            torch.no_grad():
        merged = ours([clip])
    torch.testing.assert_close(merged, ref_merged, rtol=5e-4, atol=5e-5)
    info = ours._tome_info
    assert info["size"].shape == ref_info["size"].shape
    torch.testing.assert_close(info["size"].float(), ref_info["size"].float())
    if pkw.get("trace_source"):
        # a near-tie may be decided differently (fused SDPA vs eager softmax perturb the metric in the last
        # ulp): allow a couple of swapped rows, nothing systematic
        assert info["source"].shape == ref_info["source"].shape
        assert int((info["source"] != ref_info["source"]).sum()) <= 16
    assert not torch.allclose(merged, plain, atol=1e-6) or r == 0          # the patch actually did something


def test_patch_api_surface():
    import inspect
    import tome
    for name in ("videomae", "timesformer", "motionformer", "vivit", "duplicate_videomae", "duplicate_timesformer",
                 "duplicate_motionformer", "duplicate_vivit"):
        assert callable(getattr(tome.patch, name))
    want = ["trace_source", "prop_attn", "mode", "head_aggregation", "threshold", "verbose"]
    for name, default_prop in (("videomae", False), ("timesformer", True), ("motionformer", True), ("vivit", True)):
        sig = inspect.signature(getattr(tome.patch, name))
        assert list(sig.parameters)[1:] == want
        assert sig.parameters["prop_attn"].default is default_prop
    with refshim.reference_modules():
        import tome as ref_tome
        for name in ("videomae", "timesformer", "motionformer"):
            ref_sig = inspect.signature(getattr(ref_tome.patch, name))
            assert list(ref_sig.parameters)[1:] == want
