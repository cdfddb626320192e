"""Whole-model golden logits from the UNMODIFIED reference models + reference tome patches (CPU).

    python tests/golden/make_model_golden.py          # build container only (/root/reference)

    python tests/golden/make_model_golden.py --full   # ViT-B / 12-layer cases -> fullsize.npz

Two sets.  ``MODEL_CASES`` (models.npz): tiny widths (embed 96, 3 heads, depth 3), every mode.  ``FULL_CASES``
(fullsize.npz): the BASELINE.json architectures at full size (ViT-B, 12 layers, 12 heads, 400 classes) with, per
layer, the index lists the reference's closures captured (``src_idx`` / ``dst_idx`` / ``unm_idx``, and ``node_max``
for the hybrid threshold) -- what a teacher-forced run replays and a free-running run is compared against.
The weights are NOT stored -- both sides fill every parameter from the same seeded stream in sorted-key
order (``seeded_fill``), so the GPU box can rebuild identical weights without the reference.  Stored per
case: plain (r = 0) and ToMe logits, final token sizes (+ the per-layer lists for FULL_CASES).
ViViT: the reference patch (tome/patch/vivit.py, unmodified) runs on the installed HuggingFace modules through
tests/refshim.py::build_reference_vivit, which only restores the old layer API the patch was written for."""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

DIM, HEADS, DEPTH, CLASSES = 96, 3, 3, 10

MODEL_CASES = [
    dict(name="videomae_merge", model="videomae", frames=4, r=40, kw=dict()),
    dict(name="videomae_propattn_dec", model="videomae", frames=4, r=(40, -1), kw=dict(prop_attn=True)),
    dict(name="videomae_hybrid", model="videomae", frames=4, r=40, kw=dict(mode="hybrid", threshold=0.6, prop_attn=True)),
    dict(name="videomae_drop", model="videomae", frames=4, r=40, kw=dict(mode="drop")),
    dict(name="videomae_concat", model="videomae", frames=4, r=40, kw=dict(head_aggregation="concat")),
    dict(name="timesformer_merge", model="timesformer", frames=4, r=18, kw=dict()),
    dict(name="timesformer_hybrid", model="timesformer", frames=4, r=18, kw=dict(mode="hybrid", threshold=0.6)),
    dict(name="timesformer_drop", model="timesformer", frames=4, r=18, kw=dict(mode="drop")),
    dict(name="motionformer_merge", model="motionformer", frames=8, r=18, kw=dict()),
    dict(name="motionformer_noprop", model="motionformer", frames=8, r=[18, 10, 30], kw=dict(prop_attn=False)),
    dict(name="vivit_merge", model="vivit", frames=8, r=60, kw=dict()),
    dict(name="vivit_drop", model="vivit", frames=8, r=(60, -1), kw=dict(mode="drop", prop_attn=False)),
    dict(name="vivit_hybrid", model="vivit", frames=8, r=60, kw=dict(mode="hybrid", threshold=0.4)),
    dict(name="vivit_concat_noprop", model="vivit", frames=8, r=[60, 0, 200], kw=dict(head_aggregation="concat", prop_attn=False)),
    # layer duplication ablation (tools/test_net.py:270-283: duplicate_<model>(model, layer, quantity), THEN the patch,
    # r = [0] * layer + [R] * quantity + [0] * rest)
    dict(name="videomae_duplicate", model="videomae", frames=4, r=[0, 40, 40, 0], kw=dict(), duplicate=(1, 2)),
    dict(name="timesformer_duplicate", model="timesformer", frames=4, r=[0, 18, 18, 0], kw=dict(), duplicate=(1, 2)),
    dict(name="motionformer_duplicate", model="motionformer", frames=8, r=[0, 18, 18, 18, 0], kw=dict(), duplicate=(1, 3)),
    dict(name="vivit_duplicate", model="vivit", frames=8, r=[0, 60, 60, 0], kw=dict(prop_attn=False), duplicate=(1, 2)),
]

# BASELINE.json configs 2-5 at full size: ViT-B, 12 layers, 12 heads, 400 classes; r / mode as experiments.sh uses them
FULL = dict(dim=768, heads=12, depth=12, classes=400, wstd=0.02)
FULL_CASES = [
    dict(name="videomae_b_r100", model="videomae", frames=16, r=(100, 0), kw=dict(), batch=2, **FULL),
    dict(name="videomae_b_r150dec_prop", model="videomae", frames=16, r=(150, -1), kw=dict(prop_attn=True), batch=1, **FULL),
    dict(name="videomae_b_hybrid08", model="videomae", frames=16, r=(150, 0), kw=dict(mode="hybrid", threshold=0.8), batch=1, **FULL),
    dict(name="timesformer_b_r18", model="timesformer", frames=8, r=(18, 0), kw=dict(), batch=2, **FULL),
    dict(name="motionformer_b_r18", model="motionformer", frames=16, r=(18, 0), kw=dict(), batch=1, **FULL),
    dict(name="vivit_b_r300", model="vivit", frames=32, r=(300, 0), kw=dict(), batch=1, **FULL),
    dict(name="vivit_b_hybrid04", model="vivit", frames=32, r=(300, 0), kw=dict(mode="hybrid", threshold=0.4), batch=1, **FULL),
]


def dims(case):
    return (case.get("dim", DIM), case.get("heads", HEADS), case.get("depth", DEPTH), case.get("classes", CLASSES))


def seeded_fill(module, seed=123, wstd=0.08):
    """Deterministic weights independent of construction order: sorted state-dict keys, one stream."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k in sorted(sd):
        v = sd[k]
        if not v.dtype.is_floating_point:
            continue
        if k.endswith("norm.weight") or "norm1.weight" in k or "norm2.weight" in k or "layernorm" in k and k.endswith("weight"):
            v.copy_(1.0 + 0.05 * torch.randn(v.shape, generator=g))
        elif v.dim() <= 1:
            v.copy_(0.02 * torch.randn(v.shape, generator=g))
        else:
            v.copy_(wstd * torch.randn(v.shape, generator=g))
    module.load_state_dict(sd)
    return module


def clip_for(case):
    g = torch.Generator().manual_seed(hash(case["name"]) % 1000 if False else sum(map(ord, case["name"])))
    return torch.rand(case.get("batch", 2), 3, case["frames"], 224, 224, generator=g)


def build_reference(case):
    import refshim
    dim, heads, depth, classes = dims(case)
    if case["model"] == "videomae":
        import slowfast.models.videomae_video_model_builder as vb
        m = vb.VisionTransformer(patch_size=16, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=classes,
                                 all_frames=case["frames"], tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
        return refshim._Wrap(m)
    if case["model"] == "timesformer":
        import slowfast.models.timesformer as tf
        m = tf.VisionTransformer(img_size=224, num_classes=classes, patch_size=16, embed_dim=dim, depth=depth,
                                 num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=case["frames"],
                                 attention_type='divided_space_time')
        return refshim._Wrap(m)
    if case["model"] == "motionformer":
        import slowfast.models.motionformer_video_model_builder as mb
        C = refshim._Cfg
        cfg = C(DATA=C(TRAIN_CROP_SIZE=224), MODEL=C(NUM_CLASSES=classes), EPICKITCHENS=C(NUM_CLASSES=None),
                MOTIONFORMER=C(PATCH_SIZE=16, PATCH_SIZE_TEMP=2, CHANNELS=3, EMBED_DIM=dim, DEPTH=depth, NUM_HEADS=heads,
                               MLP_RATIO=4, QKV_BIAS=True, VIDEO_INPUT=True, TEMPORAL_RESOLUTION=case["frames"] // 2,
                               USE_MLP=True, DROP=0.0, POS_DROPOUT=0.0, DROP_PATH=0.0, HEAD_DROPOUT=0.0, HEAD_ACT="tanh",
                               ATTN_DROPOUT=0.0, POS_EMBED="separate", ATTN_LAYER="trajectory",
                               USE_ORIGINAL_TRAJ_ATTN_CODE=True, APPROX_ATTN_TYPE="none", APPROX_ATTN_DIM=128))
        return mb.Motionformer(cfg)
    if case["model"] == "vivit":
        return refshim.build_reference_vivit(classes, num_frames=case["frames"], hidden_size=dim, num_hidden_layers=depth,
                                             num_attention_heads=heads, intermediate_size=4 * dim)
    raise KeyError(case["model"])


def build_ours(case):
    """The same architecture from hostmodels (used by the tests, CPU or GPU)."""
    from hostmodels.videomae import VideoMAE, VisionTransformer as VMAE
    from hostmodels.timesformer import TimeSformer, VisionTransformer as TSF
    import hostmodels
    ln = partial(torch.nn.LayerNorm, eps=1e-6)
    dim, heads, depth, classes = dims(case)
    if case["model"] == "videomae":
        w = VideoMAE.__new__(VideoMAE)
        torch.nn.Module.__init__(w)
        w.num_classes = classes
        w.model = VMAE(patch_size=16, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True, norm_layer=ln,
                       num_classes=classes, all_frames=case["frames"], tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
        return w
    if case["model"] == "timesformer":
        w = TimeSformer.__new__(TimeSformer)
        torch.nn.Module.__init__(w)
        w.num_classes, w.attention_type = classes, 'divided_space_time'
        w.model = TSF(img_size=224, num_classes=classes, patch_size=16, embed_dim=dim, depth=depth, num_heads=heads,
                      mlp_ratio=4, qkv_bias=True, norm_layer=ln, num_frames=case["frames"])
        return w
    if case["model"] == "motionformer":
        return hostmodels.Motionformer(num_classes=classes, num_frames=case["frames"], embed_dim=dim, depth=depth,
                                       num_heads=heads)
    if case["model"] == "vivit":
        return hostmodels.ViViT(num_classes=classes, num_frames=case["frames"], hidden_size=dim, num_hidden_layers=depth,
                                num_attention_heads=heads, intermediate_size=4 * dim)
    raise KeyError(case["model"])


_MATCHERS = ("bipartite_soft_matching", "bipartite_soft_matching_drop", "bipartite_soft_matching_hybrid")


class trace_reference_matching:
    """Record, per reduction step, what the reference's matching closures captured (merge.py:67-73 ``src_idx``,
    ``dst_idx``, ``unm_idx`` / ``und_idx``; hybrid: ``node_max``) by wrapping the three matchers where the reference
    patch module imported them."""

    def __init__(self, model_name):
        self.mod = sys.modules["tome.patch." + model_name]
        self.layers = []

    def _wrap(self, fn):
        def spy(metric, r, *a, **k):
            out = fn(metric, r, *a, **k)
            f = out[0] if isinstance(out, tuple) else out
            cells = dict(zip(f.__code__.co_freevars, (c.cell_contents for c in (f.__closure__ or ()))))
            rec = {}
            if "src_idx" in cells:
                rec["src"] = cells["src_idx"][..., 0].to(torch.int16).numpy()
                rec["unm"] = cells["unm_idx" if "unm_idx" in cells else "und_idx"][..., 0].to(torch.int16).numpy()
                if "dst_idx" in cells:
                    rec["dst"] = cells["dst_idx"][..., 0].to(torch.int16).numpy()
                # the reference's own margins (same ops as merge.py:51-64 on the same machine, so node_max is bitwise
                # what the closure sorted): the free-running comparison allows a difference only inside a near-tie
                cls, dist = (list(a) + [False, False])[:2] if a else (k.get("class_token", False), k.get("distill_token", False))
                m = metric / metric.norm(dim=-1, keepdim=True)
                scores = m[..., ::2, :] @ m[..., 1::2, :].transpose(-1, -2)
                if cls:
                    scores[..., 0, :] = -float("inf")
                if dist:
                    scores[..., :, 0] = -float("inf")
                top2 = scores.topk(min(2, scores.shape[-1]), dim=-1).values
                rec["node_max"] = top2[..., 0].float().numpy()
                rec["gap2"] = (top2[..., 0] - top2[..., -1]).float().numpy()
                if "node_max" in cells:
                    assert torch.equal(cells["node_max"], top2[..., 0])
            self.layers.append(rec)
            return out
        return spy

    def __enter__(self):
        self.saved = {n: getattr(self.mod, n) for n in _MATCHERS}
        for n, f in self.saved.items():
            setattr(self.mod, n, self._wrap(f))
        return self

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(self.mod, n, f)
        return False


def run_reference(case, trace=False):
    import refshim
    clip = clip_for(case)
    with refshim.reference_modules():
        import tome as ref_tome
        ref = seeded_fill(build_reference(case).eval(), wstd=case.get("wstd", 0.08))
        with torch.no_grad():
            plain = ref([clip]).clone()
        if case.get("duplicate"):
            getattr(ref_tome.patch, "duplicate_" + case["model"])(ref, *case["duplicate"])
        getattr(ref_tome.patch, case["model"])(ref, **case["kw"])
        ref.r = case["r"]
        layers = None
        with torch.no_grad():
            if trace:
                with trace_reference_matching(case["model"]) as t:
                    merged = ref([clip]).clone()
                layers = t.layers
            else:
                merged = ref([clip]).clone()
        size = ref._tome_info["size"].float().clone()
    return plain, merged, size, layers


def main():
    full = "--full" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    torch.set_num_threads(os.cpu_count() or 4)
    fname = os.path.join(HERE, "fullsize.npz" if full else "models.npz")
    out = dict(np.load(fname)) if (only and os.path.exists(fname)) else {}
    for case in (FULL_CASES if full else MODEL_CASES):
        if only and case["name"] not in only:
            continue
        plain, merged, size, layers = run_reference(case, trace=full)
        out[case["name"] + "/plain"] = plain.numpy()
        out[case["name"] + "/tome"] = merged.numpy()
        out[case["name"] + "/size"] = size.numpy()
        for i, rec in enumerate(layers or ()):
            for k, v in rec.items():
                out[f"{case['name']}/L{i}/{k}"] = v
        print(f"{case['name']:26s} plain {plain.abs().mean():.4f}  tome-plain {(merged - plain).abs().mean():.5f}  "
              f"tokens {tuple(size.shape)}", flush=True)
    np.savez_compressed(fname, **out)
    print("wrote", os.path.basename(fname), os.path.getsize(fname) // 1024, "KiB")


if __name__ == "__main__":
    main()
