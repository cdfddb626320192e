"""VideoMAE ViT (slowfast/models/videomae_video_model_builder.py) restated.

Module names follow the reference so ``state_dict`` keys are identical:
  model.patch_embed.proj, model.blocks.{i}.{norm1,attn.{qkv,q_bias,v_bias,proj},norm2,mlp.{fc1,fc2}},
  model.fc_norm, model.head        (videomae builder:180-292, 363-397)
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fastlinear


FUSED_MLP_MIN_ROWS = 2048


class Mlp(nn.Module):                                   # builder:40-56
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        h = fused_fc1_gelu(self, x)
        return self.drop(self.fc2(h if h is not None else self.act(self.fc1(x))))


def fused_fc1_gelu(mlp, x):
    """``act(fc1(x))`` of an fc1 -> nn.GELU (erf) -> fc2 MLP from ONE tcgen05 GEMM with bias and activation in its
    epilogue (``tome_linear_gelu``) instead of a library GEMM plus an elementwise pass over the (tokens, 4C)
    tensor.  CUDA bf16 inference only, from a few thousand rows (tools/microbench.py); None otherwise."""
    if (not mlp.training and x.is_cuda and x.dtype == torch.bfloat16 and isinstance(mlp.act, nn.GELU)
            and mlp.act.approximate == "none" and x.numel() // x.shape[-1] >= FUSED_MLP_MIN_ROWS):
        from tome import _native
        if _native.linear_gelu_supported(x, mlp.fc1.weight, mlp.fc1.bias):
            return _native.linear_gelu(x, mlp.fc1.weight, mlp.fc1.bias)
    if (not mlp.training and x.is_cuda and x.dtype == torch.float32 and isinstance(mlp.act, nn.GELU)
            and mlp.act.approximate == "none"):
        from tome import _native
        if _native.linear_f32_usable(x, mlp.fc1.weight, mlp.fc1.bias):       # fp32: exact-split tensor-core GEMM, erf GELU in its epilogue
            # fc2 is an exact-split GEMM too (fastlinear.TomeLinear): hand it the planes, skip the fp32 round trip
            planes = (isinstance(mlp.fc2, fastlinear.TomeLinear) and not torch.is_grad_enabled()
                      and _native.linear_f32_weight_ok(mlp.fc2.weight, mlp.fc2.bias))
            return _native.linear_f32(x, mlp.fc1.weight, mlp.fc1.bias, gelu=True, out="planes" if planes else "fp32")
    return None


class Attention(nn.Module):                             # builder:59-103
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 attn_head_dim=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = attn_head_dim if attn_head_dim is not None else dim // num_heads
        all_head_dim = head_dim * num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, all_head_dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(all_head_dim))
            self.v_bias = nn.Parameter(torch.zeros(all_head_dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(all_head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv_bias = None
        if self.q_bias is not None:
            qkv_bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias, requires_grad=False), self.v_bias))
        if x.is_cuda and x.dtype == torch.float32 and not self.training:
            from tome import _native
            if _native.linear_f32_usable(x, self.qkv.weight, qkv_bias) and _native.attention_f32_planes_ok(x, self.qkv.weight, self.num_heads):
                # fp32 inference: exact-split QKV GEMM and flash attention on tcgen05 (tome_linear_f32 / tome_attention_f32),
                # planes from one to the other and on to the projection
                qkv3 = _native.linear_f32(x, self.qkv.weight, qkv_bias, out="planes")
                keep = isinstance(self.proj, fastlinear.TomeLinear)
                ctx = _native.attention_f32(qkv3, self.num_heads, self.scale, out="planes" if keep else "fp32")
                return self.proj_drop(self.proj(ctx))
        qkv_flat = fastlinear.linear(x, self.qkv.weight, qkv_bias)
        qkv = qkv_flat.reshape(B, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=self.scale,
                                           dropout_p=self.attn_drop.p if self.training else 0.0)
        return self.proj_drop(self.proj(x.transpose(1, 2).reshape(B, N, -1)))


class Block(nn.Module):                                 # builder:106-135
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attn_head_dim=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, attn_head_dim=attn_head_dim)
        self.drop_path = nn.Identity()                  # stochastic depth is identity at inference
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if init_values is not None and init_values > 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1, self.gamma_2 = None, None

    def forward(self, x):
        if self.gamma_1 is None:
            x = x + self.drop_path(self.attn(self.norm1(x)))
            x = x + self.drop_path(self.mlp(self.norm2(x)))
        else:
            x = x + self.drop_path(self.gamma_1 * self.attn(self.norm1(x)))
            x = x + self.drop_path(self.gamma_2 * self.mlp(self.norm2(x)))
        return x


class PatchEmbed(nn.Module):                            # builder:138-160
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.tubelet_size = int(tubelet_size)
        self.num_patches = (img_size // patch_size) ** 2 * (num_frames // self.tubelet_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(self.tubelet_size, patch_size, patch_size),
                              stride=(self.tubelet_size, patch_size, patch_size))

    def forward(self, x):
        B, C, T, H, W = x.shape
        assert (H, W) == self.img_size, f"Input image size ({H}*{W}) doesn't match model {self.img_size}."
        if x.dtype == torch.uint8 and (self.training or not x.is_cuda or torch.is_grad_enabled()):
            x = (x.float() / 255).to(self.proj.weight.dtype)        # decoder-style frames; on the CUDA path tome_patchify converts
        if self.training or not x.is_cuda:
            return self.proj(x).flatten(2).transpose(1, 2)
        # kernel == stride, so the tubelet conv is one GEMM over non-overlapping patches: cuDNN's
        # implicit-GEMM Conv3d ran as an fp32 SIMT kernel (22% of the forward, profiles/r01_launches_v1).
        tt, (ph, pw) = self.tubelet_size, self.patch_size
        wdt = self.proj.weight.dtype
        if (not torch.is_grad_enabled() and pw % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16, torch.uint8)
                and wdt in (torch.float32, torch.bfloat16)):
            # one coalesced pass (tome_patchify, cast to the weight dtype included) instead of torch's generic
            # 8-d strided copy, which ran at a sixth of HBM speed (profiles/r01b)
            from tome import _native
            x = _native.patchify(x, tt, ph, pw, wdt)
        else:
            x = x.reshape(B, C, T // tt, tt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
            x = x.reshape(B, (T // tt) * (H // ph) * (W // pw), C * tt * ph * pw)
        return fastlinear.linear(x, self.proj.weight.reshape(self.proj.out_channels, -1), self.proj.bias)


def get_sinusoid_encoding_table(n_position, d_hid):    # builder:164-174
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float).unsqueeze(0)


def trunc_normal_(t, std=.02):
    return nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)


class VisionTransformer(nn.Module):                     # builder:177-305
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, fc_drop_rate=0., drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0.,
                 use_learnable_pos_emb=False, init_scale=0., all_frames=16, tubelet_size=2, use_checkpoint=False,
                 use_mean_pooling=True):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, all_frames, tubelet_size)
        num_patches = self.patch_embed.num_patches
        if use_learnable_pos_emb:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        else:   # plain tensor in the reference (builder:216), i.e. absent from the state dict
            self.register_buffer("pos_embed", get_sinusoid_encoding_table(num_patches, embed_dim), persistent=False)
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, norm_layer=norm_layer, init_values=init_values)
            for _ in range(depth)])
        self.norm = nn.Identity() if use_mean_pooling else norm_layer(embed_dim)
        self.fc_norm = norm_layer(embed_dim) if use_mean_pooling else None
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        if use_learnable_pos_emb:
            trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)
        if num_classes > 0:
            trunc_normal_(self.head.weight, std=.02)
            self.head.weight.data.mul_(init_scale)
            self.head.bias.data.mul_(init_scale)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    def forward_features(self, x):
        x = x[0]                                        # builder:273 -- input is a list of pathways
        x = self.patch_embed(x)
        if self.pos_embed is not None:
            pos = self.pos_embed.to(dtype=x.dtype, device=x.device)
            info = getattr(self.blocks[0], "_tome_info", None) if len(self.blocks) else None
            fused = None
            if info is not None and not self.training and isinstance(self.pos_drop, nn.Dropout):
                # ToMe-patched CUDA inference: the position-embedding add and the first block's norm1 in one
                # pass (tome_add_rows_layernorm); the block picks the normalised tokens up from _tome_info
                from tome.patch.videomae import fusable_norm
                fn = fusable_norm(self.blocks[0].norm1, x)
                if fn is not None and pos.shape[1:] == x.shape[1:]:
                    from tome import _native
                    x, normed = _native.add_layernorm(x, pos, fn)
                    info["normed1"] = (x, self.blocks[0].norm1, normed)
                    fused = True
            if fused is None:
                x = x + pos
        if self.training:                               # identity in eval (and keeps x the tensor norm1 was taken of)
            x = self.pos_drop(x)
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        if self.fc_norm is not None:
            return self.fc_norm(x.mean(1))
        return x[:, 0]

    def forward(self, x):
        return self.head(self.fc_dropout(self.forward_features(x)))


def videomae_vit_base_patch16_224(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def videomae_vit_small_patch16_224(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


class VideoMAE(nn.Module):                              # builder:363-397 (cfg replaced by keywords)
    def __init__(self, arch="vit_base_patch16_224", num_classes=400, num_frames=16, tubelet_size=2,
                 use_mean_pooling=True, init_scale=0.001, **kwargs):
        super().__init__()
        func = {"vit_base_patch16_224": videomae_vit_base_patch16_224,
                "vit_small_patch16_224": videomae_vit_small_patch16_224}[arch]
        self.num_classes = num_classes
        self.model = func(num_classes=num_classes, all_frames=num_frames, tubelet_size=tubelet_size,
                          use_mean_pooling=use_mean_pooling, init_scale=init_scale, **kwargs)
        fastlinear.install(self)                        # fp32 CUDA inference: linears on tome_linear_f32

    def forward(self, x):
        return self.model(x)
