"""Import shims that let the UNMODIFIED reference models + tome patches run on CPU in the build
container.  Test infrastructure only; nothing here ships, and nothing is
copied from the reference: the modules are imported from /root/reference where they lie."""
import os
import sys
import types

import torch

REFERENCE = os.environ.get("TOME_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REFERENCE, "slowfast", "models", "timesformer.py"))


class _Registry:
    def register(self, *a, **k):
        return lambda cls: cls


def _pkg(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path:
        m.__path__ = [path]
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_REF_MODULE_PREFIXES = ("slowfast", "timm", "tome")


class reference_modules:
    """Context manager: inside it ``import tome`` / ``import slowfast.models.x`` resolve to the reference;
    on exit the repo's own modules are restored."""

    def __enter__(self):
        import transformers  # noqa: F401  (must be imported BEFORE the timm stub: HF probes timm.__spec__)
        self.saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_MODULE_PREFIXES}
        for k in self.saved:
            del sys.modules[k]
        self.saved_path = list(sys.path)

        def drop_path(x, drop_prob=0., training=False):
            return x

        class DropPath(torch.nn.Module):
            def __init__(self, p=None):
                super().__init__()

            def forward(self, x):
                return x

        sys.modules["slowfast"] = _pkg("slowfast", os.path.join(REFERENCE, "slowfast"))
        sys.modules["slowfast.models"] = _pkg("slowfast.models", os.path.join(REFERENCE, "slowfast", "models"))
        sys.modules["slowfast.models.build"] = _pkg("slowfast.models.build", MODEL_REGISTRY=_Registry())
        sys.modules["timm"] = _pkg("timm", "/nonexistent")
        sys.modules["timm.models"] = _pkg("timm.models", "/nonexistent")
        sys.modules["timm.models.layers"] = _pkg(
            "timm.models.layers", drop_path=drop_path, DropPath=DropPath, to_2tuple=lambda v: (v, v),
            trunc_normal_=lambda t, std=1.0, **k: torch.nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0))
        sys.modules["timm.models.registry"] = _pkg("timm.models.registry", register_model=lambda f: f)
        sys.modules["timm.models.resnet"] = _pkg("timm.models.resnet", resnet26d=None, resnet50d=None)
        sys.modules["timm.data"] = _pkg("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                                        IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
        sys.path.insert(0, REFERENCE)
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_MODULE_PREFIXES]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        sys.path[:] = self.saved_path
        return False


class _Wrap(torch.nn.Module):
    """The reference patches expect ``wrapper.model`` (videomae.py:176, timesformer.py:191)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model(x)


def build_reference_videomae(num_classes, num_frames, depth=None):
    import slowfast.models.videomae_video_model_builder as vb
    from functools import partial
    kw = dict(num_classes=num_classes, all_frames=num_frames, tubelet_size=2, init_scale=0.001, use_mean_pooling=True)
    if depth is None:
        return _Wrap(vb.vit_base_patch16_224(**kw))
    return _Wrap(vb.VisionTransformer(patch_size=16, embed_dim=768, depth=depth, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw))


def build_reference_timesformer(num_classes, num_frames, depth=12):
    import slowfast.models.timesformer as tf
    from functools import partial
    return _Wrap(tf.VisionTransformer(img_size=224, num_classes=num_classes, patch_size=16, embed_dim=768, depth=depth,
                                      num_heads=12, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_path_rate=0.1,
                                      num_frames=num_frames, attention_type='divided_space_time'))


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def build_reference_motionformer(num_classes, num_frames, depth=12):
    import slowfast.models.motionformer_video_model_builder as mb
    cfg = _Cfg(
        DATA=_Cfg(TRAIN_CROP_SIZE=224),
        MODEL=_Cfg(NUM_CLASSES=num_classes),
        EPICKITCHENS=_Cfg(NUM_CLASSES=None),
        MOTIONFORMER=_Cfg(PATCH_SIZE=16, PATCH_SIZE_TEMP=2, CHANNELS=3, EMBED_DIM=768, DEPTH=depth, NUM_HEADS=12,
                          MLP_RATIO=4, QKV_BIAS=True, VIDEO_INPUT=True, TEMPORAL_RESOLUTION=num_frames // 2,
                          USE_MLP=True, DROP=0.0, POS_DROPOUT=0.0, DROP_PATH=0.0, HEAD_DROPOUT=0.0, HEAD_ACT="tanh",
                          ATTN_DROPOUT=0.0, POS_EMBED="separate", ATTN_LAYER="trajectory",
                          USE_ORIGINAL_TRAJ_ATTN_CODE=True, APPROX_ATTN_TYPE="none", APPROX_ATTN_DIM=128))
    return mb.Motionformer(cfg)


def build_reference_vivit(num_classes, num_frames=32, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                          intermediate_size=3072):
    """The reference's ViViT wrapper + the installed HuggingFace modules, given back the pre-4.4x layer API the
    reference patch (tome/patch/vivit.py:17-130) was written against.  Nothing of the reference is re-stated:
    ``ViViT.forward`` (vivit_video_model_builder.py:33-60) and the whole tome patch run as they lie; the shim only
      * builds ``VivitConfig`` by keyword (builder:17 passes it positionally, which transformers 5.5 rejects; the
        values are configs/vivit/kinetics/tome_vivit_8x32_224.json's),
      * gives ``VivitSelfAttention`` the ``transpose_for_scores`` / ``dropout`` members the patch calls (vivit.py:93-110),
      * runs the encoder as the old API did: ``layer(hidden_states, head_mask, output_attentions)[0]``.
    Call inside ``reference_modules()``."""
    import slowfast.models.vivit_video_model_builder as vb
    from transformers import VivitConfig, VivitModel
    from transformers.modeling_outputs import BaseModelOutput
    from transformers.models.vivit.modeling_vivit import VivitEncoder, VivitSelfAttention

    class OldApiEncoder(VivitEncoder):
        def forward(self, hidden_states, **kwargs):
            for layer_module in self.layer:
                if hasattr(layer_module, "_tome_info"):       # patched: the old tuple-returning signature
                    hidden_states = layer_module(hidden_states, None, False)[0]
                else:
                    hidden_states = layer_module(hidden_states)
            return BaseModelOutput(last_hidden_state=hidden_states)

    def transpose_for_scores(self, x):
        return x.view(x.size()[:-1] + (self.num_attention_heads, self.attention_head_size)).permute(0, 2, 1, 3)

    class ShimViViT(vb.ViViT):
        def __init__(self):
            config = VivitConfig(image_size=224, num_frames=num_frames, tubelet_size=[2, 16, 16], num_channels=3,
                                 hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                                 num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                                 hidden_act="gelu_fast", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                 initializer_range=0.02, layer_norm_eps=1e-6, qkv_bias=True)
            vb.VivitPreTrainedModel.__init__(self, config)
            self.num_labels = num_classes
            self.vivit = VivitModel(config, add_pooling_layer=False)
            self.classifier = torch.nn.Linear(config.hidden_size, num_classes)
            self.vivit.encoder.__class__ = OldApiEncoder
            for m in self.vivit.modules():
                if isinstance(m, VivitSelfAttention):
                    m.dropout = torch.nn.Dropout(m.dropout_prob)
                    m.transpose_for_scores = types.MethodType(transpose_for_scores, m)

    return ShimViViT()
