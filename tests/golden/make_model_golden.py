"""Whole-model golden logits from the UNMODIFIED reference models + reference tome patches (CPU).

    python tests/golden/make_model_golden.py          # build container only (/root/reference)

Tiny widths (embed 96, 3 heads, depth 3) keep this fast; the weights are NOT stored -- both sides fill
every parameter from the same seeded stream in sorted-key order (``seeded_fill``), so the GPU box can
rebuild identical weights without the reference.  Stored per case: the clip seed, plain (r = 0) and
ToMe logits, final token sizes.  ViViT is absent: the reference's ViViT does not construct against the
installed transformers -- its model-level parity is unpinned, its merge path is pinned
by the class-token merge.py goldens."""
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
]


def seeded_fill(module, seed=123):
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
            v.copy_(0.08 * torch.randn(v.shape, generator=g))
    module.load_state_dict(sd)
    return module


def clip_for(case):
    g = torch.Generator().manual_seed(hash(case["name"]) % 1000 if False else sum(map(ord, case["name"])))
    return torch.rand(2, 3, case["frames"], 224, 224, generator=g)


def build_reference(case):
    import refshim
    if case["model"] == "videomae":
        import slowfast.models.videomae_video_model_builder as vb
        m = vb.VisionTransformer(patch_size=16, embed_dim=DIM, depth=DEPTH, num_heads=HEADS, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=CLASSES,
                                 all_frames=case["frames"], tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
        return refshim._Wrap(m)
    if case["model"] == "timesformer":
        import slowfast.models.timesformer as tf
        m = tf.VisionTransformer(img_size=224, num_classes=CLASSES, patch_size=16, embed_dim=DIM, depth=DEPTH,
                                 num_heads=HEADS, mlp_ratio=4, qkv_bias=True,
                                 norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=case["frames"],
                                 attention_type='divided_space_time')
        return refshim._Wrap(m)
    if case["model"] == "motionformer":
        import slowfast.models.motionformer_video_model_builder as mb
        C = refshim._Cfg
        cfg = C(DATA=C(TRAIN_CROP_SIZE=224), MODEL=C(NUM_CLASSES=CLASSES), EPICKITCHENS=C(NUM_CLASSES=None),
                MOTIONFORMER=C(PATCH_SIZE=16, PATCH_SIZE_TEMP=2, CHANNELS=3, EMBED_DIM=DIM, DEPTH=DEPTH, NUM_HEADS=HEADS,
                               MLP_RATIO=4, QKV_BIAS=True, VIDEO_INPUT=True, TEMPORAL_RESOLUTION=case["frames"] // 2,
                               USE_MLP=True, DROP=0.0, POS_DROPOUT=0.0, DROP_PATH=0.0, HEAD_DROPOUT=0.0, HEAD_ACT="tanh",
                               ATTN_DROPOUT=0.0, POS_EMBED="separate", ATTN_LAYER="trajectory",
                               USE_ORIGINAL_TRAJ_ATTN_CODE=True, APPROX_ATTN_TYPE="none", APPROX_ATTN_DIM=128))
        return mb.Motionformer(cfg)
    raise KeyError(case["model"])


def build_ours(case):
    """The same architecture from hostmodels (used by the tests, CPU or GPU)."""
    from hostmodels.videomae import VideoMAE, VisionTransformer as VMAE
    from hostmodels.timesformer import TimeSformer, VisionTransformer as TSF
    import hostmodels
    ln = partial(torch.nn.LayerNorm, eps=1e-6)
    if case["model"] == "videomae":
        w = VideoMAE(arch="vit_small_patch16_224", num_classes=CLASSES, num_frames=case["frames"])
        w.model = VMAE(patch_size=16, embed_dim=DIM, depth=DEPTH, num_heads=HEADS, mlp_ratio=4, qkv_bias=True, norm_layer=ln,
                       num_classes=CLASSES, all_frames=case["frames"], tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
        return w
    if case["model"] == "timesformer":
        w = TimeSformer.__new__(TimeSformer)
        torch.nn.Module.__init__(w)
        w.num_classes, w.attention_type = CLASSES, 'divided_space_time'
        w.model = TSF(img_size=224, num_classes=CLASSES, patch_size=16, embed_dim=DIM, depth=DEPTH, num_heads=HEADS,
                      mlp_ratio=4, qkv_bias=True, norm_layer=ln, num_frames=case["frames"])
        return w
    if case["model"] == "motionformer":
        return hostmodels.Motionformer(num_classes=CLASSES, num_frames=case["frames"], embed_dim=DIM, depth=DEPTH,
                                       num_heads=HEADS)
    raise KeyError(case["model"])


def main():
    import refshim
    torch.set_num_threads(4)
    out = {}
    for case in MODEL_CASES:
        clip = clip_for(case)
        with refshim.reference_modules():
            import tome as ref_tome
            ref = seeded_fill(build_reference(case).eval())
            with torch.no_grad():
                plain = ref([clip]).clone()
            getattr(ref_tome.patch, case["model"])(ref, **case["kw"])
            ref.r = case["r"]
            with torch.no_grad():
                merged = ref([clip]).clone()
            size = ref._tome_info["size"].float().clone()
        out[case["name"] + "/plain"] = plain.numpy()
        out[case["name"] + "/tome"] = merged.numpy()
        out[case["name"] + "/size"] = size.numpy()
        print(f"{case['name']:26s} plain {plain.abs().mean():.4f}  tome-plain {(merged - plain).abs().mean():.5f}  "
              f"tokens {tuple(size.shape)}")
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)
    print("wrote models.npz", os.path.getsize(os.path.join(HERE, "models.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
