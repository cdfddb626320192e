"""ViViT-B/16x2 (what the reference wraps: HuggingFace ``VivitModel`` inside
slowfast/models/vivit_video_model_builder.py) as a self-contained module.

The reference ViViT cannot be built against the installed transformers 5.5 (positional
``VivitConfig`` arguments, tuple-returning layer API); this restatement needs no HF import and
keeps HF's parameter names so checkpoints interchange:
``vivit.embeddings.{cls_token,position_embeddings,patch_embeddings.projection}``,
``vivit.encoder.layer.{i}.{layernorm_before,attention.attention.{query,key,value},
attention.output.dense,layernorm_after,intermediate.dense,output.dense}``, ``vivit.layernorm``,
``classifier``.  Joint space-time attention over 1 + (T/2)(H/16)(W/16) tokens with a class token
(``class_token=True`` in the ToMe patch, tome/patch/vivit.py:242)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fastlinear


class VivitConfig:
    def __init__(self, num_frames=32, image_size=224, tubelet_size=(2, 16, 16), num_channels=3, hidden_size=768,
                 num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu_fast",
                 layer_norm_eps=1e-6, qkv_bias=True, initializer_range=0.02):
        self.num_frames, self.image_size, self.tubelet_size = num_frames, image_size, list(tubelet_size)
        self.num_channels, self.hidden_size, self.num_hidden_layers = num_channels, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size, self.hidden_act = num_attention_heads, intermediate_size, hidden_act
        self.layer_norm_eps, self.qkv_bias, self.initializer_range = layer_norm_eps, qkv_bias, initializer_range


def gelu_fast(x):                                        # HF FastGELUActivation
    return 0.5 * x * (1.0 + torch.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))


class VivitTubeletEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.patch_size = config.tubelet_size
        self.num_patches = ((config.image_size // self.patch_size[2]) * (config.image_size // self.patch_size[1])
                            * (config.num_frames // self.patch_size[0]))
        self.projection = nn.Conv3d(config.num_channels, config.hidden_size, kernel_size=tuple(config.tubelet_size),
                                    stride=tuple(config.tubelet_size))

    def forward(self, pixel_values):                    # (B, T, C, H, W)
        B, T, C, H, W = pixel_values.shape
        z, ph, pw = self.patch_size
        if self.training or not pixel_values.is_cuda:
            return self.projection(pixel_values.permute(0, 2, 1, 3, 4)).flatten(2).transpose(1, 2)
        x = pixel_values.reshape(B, T // z, z, C, H // ph, ph, W // pw, pw).permute(0, 1, 4, 6, 3, 2, 5, 7)
        x = x.reshape(B, (T // z) * (H // ph) * (W // pw), C * z * ph * pw)   # kernel == stride: one GEMM
        return fastlinear.linear(x, self.projection.weight.reshape(self.projection.out_channels, -1), self.projection.bias)


class VivitEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, config.hidden_size))
        self.patch_embeddings = VivitTubeletEmbeddings(config)
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.patch_embeddings.num_patches + 1, config.hidden_size))

    def forward(self, pixel_values):
        x = self.patch_embeddings(pixel_values)
        return torch.cat((self.cls_token.expand(x.size(0), -1, -1).to(x.dtype), x), dim=1) + self.position_embeddings


class VivitSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = config.hidden_size
        self.scaling = self.attention_head_size ** -0.5
        self.query = nn.Linear(config.hidden_size, self.all_head_size, bias=config.qkv_bias)
        self.key = nn.Linear(config.hidden_size, self.all_head_size, bias=config.qkv_bias)
        self.value = nn.Linear(config.hidden_size, self.all_head_size, bias=config.qkv_bias)

    def forward(self, hidden_states, **kwargs):
        B = hidden_states.shape[0]
        if hidden_states.is_cuda and hidden_states.dtype == torch.float32:
            from tome import attention as prop_attention
            if prop_attention.usable_f32(hidden_states, self, self.num_attention_heads, self.attention_head_size):
                ctx, _ = prop_attention.attention_f32(hidden_states, self, self.num_attention_heads, self.attention_head_size,
                                                      self.scaling, None, self.query.weight, self.key.weight, self.value.weight,
                                                      self.query.bias, self.key.bias, self.value.bias,
                                                      planes_for=getattr(self, "_tome_ctx_consumer", None))
                return ctx, None
        shp = (B, -1, self.num_attention_heads, self.attention_head_size)
        q = self.query(hidden_states).view(*shp).transpose(1, 2)
        k = self.key(hidden_states).view(*shp).transpose(1, 2)
        v = self.value(hidden_states).view(*shp).transpose(1, 2)
        ctx = F.scaled_dot_product_attention(q, k, v, scale=self.scaling)
        return ctx.transpose(1, 2).reshape(B, -1, self.all_head_size), None


class VivitSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)

    def forward(self, hidden_states, input_tensor):     # residual is added in the layer
        return self.dense(hidden_states)


class VivitAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = VivitSelfAttention(config)
        self.output = VivitSelfOutput(config)

    def forward(self, hidden_states, **kwargs):
        object.__setattr__(self.attention, "_tome_ctx_consumer", self.output.dense)   # a hint (see tome/attention.py), not a submodule
        return self.output(self.attention(hidden_states)[0], hidden_states)


class VivitIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.intermediate_act_fn = gelu_fast if config.hidden_act == "gelu_fast" else F.gelu

    def forward(self, hidden_states):
        x = hidden_states
        if (self.intermediate_act_fn is gelu_fast and not self.training and x.is_cuda and x.dtype == torch.bfloat16
                and x.numel() // x.shape[-1] >= 2048):
            # dense + bias + gelu_fast from ONE tcgen05 GEMM (tome_linear_gelu): the eager form is a library GEMM
            # plus seven elementwise passes over the (tokens, 4C) tensor, each rounding to bf16
            from tome import _native
            if _native.linear_gelu_supported(x, self.dense.weight, self.dense.bias):
                return _native.linear_gelu(x, self.dense.weight, self.dense.bias, gelu="gelu_fast")
        if (self.intermediate_act_fn is gelu_fast and not self.training and x.is_cuda and x.dtype == torch.float32
                and not torch.is_grad_enabled()):
            # fp32 inference: dense + bias + gelu_fast in the exact-split GEMM's epilogue (tome_linear_f32, gelu = 2) instead of
            # a GEMM plus seven elementwise passes (16 % of the fp32 ViViT step); the result goes to the output projection as
            # split planes when that is an exact-split GEMM too
            from tome import _native
            if _native.linear_f32_usable(x, self.dense.weight, self.dense.bias):
                return _native.linear_f32(x, self.dense.weight, self.dense.bias, gelu="gelu_fast",
                                          out="planes" if getattr(self, "_planes_to_output", False) else "fp32")
        return self.intermediate_act_fn(self.dense(hidden_states))


class VivitOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)

    def forward(self, hidden_states, input_tensor):
        return self.dense(hidden_states) + input_tensor


class VivitLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = VivitAttention(config)
        self.intermediate = VivitIntermediate(config)
        self.output = VivitOutput(config)
        self.layernorm_before = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.layernorm_after = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        hidden_states = self.attention(self.layernorm_before(hidden_states)) + hidden_states
        return self.output(self.intermediate(self.layernorm_after(hidden_states)), hidden_states)


class VivitEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([VivitLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states):
        for layer in self.layer:
            hidden_states = layer(hidden_states)
        return hidden_states


class VivitModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = VivitEmbeddings(config)
        self.encoder = VivitEncoder(config)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, pixel_values):
        return self.layernorm(self.encoder(self.embeddings(pixel_values)))


class ViViT(nn.Module):                                 # vivit_video_model_builder.py:12-60
    def __init__(self, num_classes=400, config=None, **cfg_kwargs):
        super().__init__()
        self.config = config or VivitConfig(**cfg_kwargs)
        self.num_labels = num_classes
        self.vivit = VivitModel(self.config)
        self.classifier = nn.Linear(self.config.hidden_size, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)
        fastlinear.install(self)                        # fp32 CUDA inference: linears on tome_linear_f32
        for layer in self.vivit.encoder.layer:          # fc1 may hand its result to fc2 as split planes
            layer.intermediate._planes_to_output = isinstance(layer.output.dense, fastlinear.TomeLinear)

    def _init_weights(self, m):                         # HF VivitPreTrainedModel._init_weights
        std = self.config.initializer_range
        if isinstance(m, (nn.Linear, nn.Conv3d)):
            nn.init.normal_(m.weight, mean=0.0, std=std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)
        elif isinstance(m, VivitEmbeddings):
            nn.init.zeros_(m.cls_token)
            nn.init.zeros_(m.position_embeddings)

    def forward(self, pixel_values):
        x = pixel_values[0].permute(0, 2, 1, 3, 4)      # builder:42: (B, C, T, H, W) -> (B, T, C, H, W)
        return self.classifier(self.vivit(x)[:, 0, :])
