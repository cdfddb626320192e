"""Motionformer (slowfast/models/motionformer_video_model_builder.py + motionformer_vit_helper.py)
restated for the configuration the reference's ToMe runs use: trajectory attention, 3-D
tubelet embedding (PATCH_SIZE_TEMP 2), separate space / time position embeddings, MLP head
(configs/motionformer/*/tome_motionformer_224_16x4.yaml).  Parameter names follow the reference:
``patch_embed.proj, patch_embed_3d.proj, cls_token, pos_embed, temp_embed,
blocks.{i}.{norm1,attn.{qkv,proj_q,proj_kv,proj},norm2,mlp.{fc1,fc2}}, norm, pre_logits.fc, head``.

Reference quirk kept: ``patch_embed_3d.proj.weight`` is zeroed and ``temp_embed`` is zero at
construction (builder:69-70, 94-95), so a freshly built model sees identical tokens in every frame;
``randomize_degenerate_init()`` re-draws those two tensors for synthetic benchmarks."""
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fastlinear


from .videomae import fused_fc1_gelu  # noqa: E402

class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        h = fused_fc1_gelu(self, x)                     # tome_linear_gelu on the CUDA bf16 inference path
        if h is not None and not torch.is_tensor(h):        # split planes straight into fc2 (eval: dropout is the identity)
            return self.drop(self.fc2(h))
        return self.drop(self.fc2(self.drop(h if h is not None else self.act(self.fc1(x)))))


from tome.patch.motionformer import trajectory_attention  # noqa: E402  (shared with the ToMe patch)


class TrajectoryAttention(nn.Module):                   # vit_helper.py:146-267
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0., use_original_code=True):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj_kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.use_original_code = use_original_code

    def forward(self, x, seq_len=196, num_frames=8, approx='none', num_landmarks=128):
        if approx != 'none':
            raise NotImplementedError("approximate attention variants are not used by any ToMe config")
        return trajectory_attention(self, x, num_frames)[0], None


class Block(nn.Module):                                 # vit_helper.py:286-318
    def __init__(self, dim=768, num_heads=12, attn_type='trajectory', mlp_ratio=4., qkv_bias=False, drop=0.,
                 attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, use_original_code=True):
        super().__init__()
        assert attn_type == 'trajectory'
        self.norm1 = norm_layer(dim)
        self.attn = TrajectoryAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                                        proj_drop=drop, use_original_code=use_original_code)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, seq_len=196, num_frames=8, approx='none', num_landmarks=128):
        x = x + self.attn(self.norm1(x), seq_len=seq_len, num_frames=num_frames, approx=approx)[0]
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class PatchEmbed3D(nn.Module):                          # vit_helper.py:410-432
    def __init__(self, img_size=224, temporal_resolution=4, in_chans=3, patch_size=16, z_block_size=2, embed_dim=768):
        super().__init__()
        self.patch_size, self.z = patch_size, z_block_size
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(z_block_size, patch_size, patch_size),
                              stride=(z_block_size, patch_size, patch_size))

    def forward(self, x):
        B, C, T, H, W = x.shape
        if self.training or not x.is_cuda:
            return self.proj(x).flatten(2).transpose(1, 2)
        z, p = self.z, self.patch_size                  # kernel == stride: one GEMM over tubelets
        wdt = self.proj.weight.dtype
        if (not torch.is_grad_enabled() and p % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16, torch.uint8)
                and wdt in (torch.float32, torch.bfloat16)):
            from tome import _native                    # one coalesced pass instead of torch's generic 8-d strided copy
            x = _native.patchify(x, z, p, p, wdt)
        else:
            x = x.reshape(B, C, T // z, z, H // p, p, W // p, p).permute(0, 2, 4, 6, 1, 3, 5, 7)
            x = x.reshape(B, (T // z) * (H // p) * (W // p), C * z * p * p)
        return fastlinear.linear(x, self.proj.weight.reshape(self.proj.out_channels, -1), self.proj.bias)


class Motionformer(nn.Module):                          # builder:23-282, cfg replaced by keywords
    def __init__(self, img_size=224, patch_size=16, patch_size_temp=2, num_frames=16, in_chans=3, num_classes=400,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True, use_mlp=True, head_act='tanh',
                 use_original_code=True):
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = self.num_features = embed_dim
        self.temporal_resolution = num_frames // patch_size_temp
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PatchEmbed(224, patch_size, in_chans, embed_dim)
        self.patch_embed_3d = PatchEmbed3D(img_size, num_frames, in_chans, patch_size, patch_size_temp, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(0.0)
        self.temp_embed = nn.Parameter(torch.zeros(1, self.temporal_resolution, embed_dim))
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                  use_original_code=use_original_code) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        if use_mlp:
            act = {'tanh': nn.Tanh, 'gelu': nn.GELU}.get(head_act, nn.ReLU)()
            self.pre_logits = nn.Sequential(OrderedDict([('fc', nn.Linear(embed_dim, embed_dim)), ('act', act)]))
        else:
            self.pre_logits = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.cls_token, std=.02)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)
        self.patch_embed_3d.proj.weight.data.zero_()    # builder:69-70
        fastlinear.install(self)                        # fp32 CUDA inference: linears on tome_linear_f32

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def randomize_degenerate_init(self, seed=0):
        """Synthetic benchmarks only: the zeroed tubelet weights / zero temp_embed make every frame's tokens
        identical (massive exact ties); draw them like the other weights."""
        g = torch.Generator().manual_seed(seed)
        w = torch.empty(self.patch_embed_3d.proj.weight.shape)
        t = torch.empty(self.temp_embed.shape)
        nn.init.trunc_normal_(w, std=.02, generator=g)
        nn.init.trunc_normal_(t, std=.02, generator=g)
        self.patch_embed_3d.proj.weight.data.copy_(w)
        self.temp_embed.data.copy_(t)
        return self

    def forward_features(self, x):
        x = x[0]
        B = x.shape[0]
        x = self.patch_embed_3d(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        npatch = self.patch_embed.num_patches
        cls_embed = self.pos_embed[:, 0, :].unsqueeze(1)
        tile_pos = self.pos_embed[:, 1:, :].repeat(1, self.temporal_resolution, 1)
        tile_tmp = self.temp_embed.repeat_interleave(npatch, 1)
        x = self.pos_drop(x + torch.cat([cls_embed, tile_pos + tile_tmp], dim=1))
        for blk in self.blocks:
            x = blk(x, seq_len=npatch, num_frames=self.temporal_resolution, approx='none', num_landmarks=128)
        return self.pre_logits(self.norm(x)[:, 0])

    def forward(self, x):
        x = self.head(self.head_drop(self.forward_features(x)))
        if not self.training:                           # builder:281-282
            x = torch.nn.functional.softmax(x, dim=-1)
        return x
