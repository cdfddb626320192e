"""TimeSformer (slowfast/models/timesformer.py) restated: divided space-time ViT-B.

Parameter names follow the reference (``model.{cls_token,pos_embed,time_embed,patch_embed.proj,
blocks.{i}.{norm1,attn.{qkv,proj},temporal_norm1,temporal_attn.{qkv,proj},temporal_fc,norm2,
mlp.{fc1,fc2}},norm,head}``).  The reference wrapper downloads ImageNet weights in its
constructor (timesformer.py:336-346); this one never touches the network."""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fastlinear


from .videomae import fused_fc1_gelu  # noqa: E402

class Mlp(nn.Module):                                   # timesformer.py:39-55
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


class Attention(nn.Module):                             # timesformer.py:57-88
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., with_qkv=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.with_qkv = with_qkv
        if with_qkv:
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)
            self.proj_drop = nn.Dropout(proj_drop)
        self.attn_drop = nn.Dropout(attn_drop)

    def forward(self, x):
        B, N, C = x.shape
        if self.with_qkv and x.is_cuda:
            from tome import _native
            if _native.attn_short_usable(x, self.num_heads):
                # the temporal attention's 8-token sequences: one streaming pass over the QKV GEMM's output
                # (tome_attn_short) instead of a tiled flash kernel at a few percent of HBM speed
                return self.proj_drop(self.proj(_native.attn_short(self.qkv(x), self.num_heads, self.scale)))
        if self.with_qkv:
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]
        else:
            q = k = v = x.reshape(B, N, self.num_heads, C // self.num_heads).permute(0, 2, 1, 3)
        x = F.scaled_dot_product_attention(q, k, v, scale=self.scale).transpose(1, 2).reshape(B, N, C)
        if self.with_qkv:
            x = self.proj_drop(self.proj(x))
        return x


class Block(nn.Module):                                 # timesformer.py:90-153
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0.1, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attention_type='divided_space_time'):
        super().__init__()
        assert attention_type in ['divided_space_time', 'space_only', 'joint_space_time']
        self.attention_type = attention_type
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        if attention_type == 'divided_space_time':
            self.temporal_norm1 = norm_layer(dim)
            self.temporal_attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                           attn_drop=attn_drop, proj_drop=drop)
            self.temporal_fc = nn.Linear(dim, dim)
        self.drop_path = nn.Identity()                  # eval
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, B, T, W):
        P = (x.size(1) - 1) // T
        C = x.size(2)
        if self.attention_type in ['space_only', 'joint_space_time']:
            x = x + self.attn(self.norm1(x))
            return x + self.mlp(self.norm2(x))
        # temporal: 'b (p t) m -> (b p) t m'
        xt = x[:, 1:, :].reshape(B * P, T, C)
        res_t = self.temporal_attn(self.temporal_norm1(xt)).reshape(B, P * T, C)
        xt = x[:, 1:, :] + self.temporal_fc(res_t)
        # spatial: 'b (p t) m -> (b t) p m' with the class token replicated per frame
        init_cls = x[:, 0, :].unsqueeze(1)
        cls = init_cls.repeat(1, T, 1).reshape(B * T, 1, C)
        xs = xt.reshape(B, P, T, C).transpose(1, 2).reshape(B * T, P, C)
        res_s = self.attn(self.norm1(torch.cat((cls, xs), 1)))
        cls = res_s[:, 0, :].reshape(B, T, C).mean(1, keepdim=True)
        res = res_s[:, 1:, :].reshape(B, T, P, C).transpose(1, 2).reshape(B, P * T, C)
        x = torch.cat((init_cls, xt), 1) + torch.cat((cls, res), 1)
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):                            # timesformer.py:155-175
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        B, C, T, H, W = x.shape
        ph, pw = self.patch_size
        if x.is_cuda and not self.training:             # kernel == stride: one GEMM over patches
            wdt = self.proj.weight.dtype
            if (not torch.is_grad_enabled() and pw % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16, torch.uint8)
                    and wdt in (torch.float32, torch.bfloat16) and x.is_contiguous()):
                from tome import _native                # one coalesced pass (a tubelet of one frame) instead of two strided copies
                x = _native.patchify(x, 1, ph, pw, wdt).view(B * T, (H // ph) * (W // pw), C * ph * pw)
            else:
                x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H // ph, ph, W // pw, pw).permute(0, 2, 4, 1, 3, 5)
                x = x.reshape(B * T, (H // ph) * (W // pw), C * ph * pw)
            return fastlinear.linear(x, self.proj.weight.reshape(self.proj.out_channels, -1), self.proj.bias), T, W // pw
        x = self.proj(x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W))
        Wp = x.size(-1)
        return x.flatten(2).transpose(1, 2), T, Wp


class VisionTransformer(nn.Module):                     # timesformer.py:178-321
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, num_frames=8, attention_type='divided_space_time',
                 dropout=0.):
        super().__init__()
        self.attention_type = attention_type
        self.depth = depth
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        if attention_type != 'space_only':
            self.time_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
            self.time_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, norm_layer=norm_layer, attention_type=attention_type)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        if attention_type == 'divided_space_time':      # timesformer.py:230-239: temporal_fc of blocks > 0 starts at zero
            for i, blk in enumerate(self.blocks):
                if i > 0:
                    nn.init.constant_(blk.temporal_fc.weight, 0)
                    nn.init.constant_(blk.temporal_fc.bias, 0)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        B = x.shape[0]
        x, T, W = self.patch_embed(x)                    # (B*T, P, C)
        if (self.attention_type == 'divided_space_time' and x.is_cuda and not self.training and not torch.is_grad_enabled()
                and x.dtype in (torch.float32, torch.bfloat16) and x.size(2) % 8 == 0 and x.is_contiguous()
                and self.pos_embed.size(1) == x.size(1) + 1 and self.time_embed.size(1) == T):
            x = self._embed_fused(x, B, T)
            for blk in self.blocks:
                x = blk(x, B, T, W)
            return self.norm(x)[:, 0]
        x = torch.cat((self.cls_token.expand(x.size(0), -1, -1), x), dim=1) + self.pos_embed
        x = self.pos_drop(x)
        if self.attention_type != 'space_only':
            cls_tokens = x[:B, 0, :].unsqueeze(1)
            P, C = x.size(1) - 1, x.size(2)
            x = x[:, 1:].reshape(B, T, P, C).transpose(1, 2)            # '(b t) n m -> b n t m'
            x = self.time_drop(x + self.time_embed[:, None])              # time_embed over t
            x = torch.cat((cls_tokens, x.reshape(B, P * T, C)), dim=1)    # 'b (n t) m'
        for blk in self.blocks:
            x = blk(x, B, T, W)
        if self.attention_type == 'space_only':
            x = x.reshape(B, T, x.size(1), x.size(2)).mean(1)
        return self.norm(x)[:, 0]

    def _embed_fused(self, x, B, T):
        """cat(cls, x) + pos_embed, '(b t) n m -> b (n t) m', + time_embed, cat(cls, ...) (timesformer.py:287-318 of the reference
        model) as ONE pass over the patch-embedding GEMM's output: out[b, 1 + p T + t] = x[(b t), p] + (pos[1 + p] + time[t])
        through (b, p, t) row views (tome_rows_add_layernorm), out[b, 0] = cls + pos[0].  Inference only (both dropouts are
        identities); the two embedding tables are summed once and cached."""
        from tome import _native
        P, C = x.size(1), x.size(2)
        key = (self.pos_embed._version, self.time_embed._version, self.cls_token._version, x.dtype, x.device, T)
        cached = self.__dict__.get("_tome_embed_sum")
        if cached is None or cached[0] != key:
            pos, tim = self.pos_embed.detach().to(x.dtype), self.time_embed.detach().to(x.dtype)
            emb = (pos[0, 1:, None, :] + tim[0, None, :, :]).contiguous()                  # (P, T, C)
            cls_row = (self.cls_token.detach().to(x.dtype)[0, 0] + pos[0, 0]).contiguous()    # (C,)
            cached = self.__dict__["_tome_embed_sum"] = (key, emb, cls_row)
        out = torch.empty(B, 1 + P * T, C, dtype=x.dtype, device=x.device)
        _native.rows_add_layernorm(x.view(B, T, P, C).permute(0, 2, 1, 3), cached[1].unsqueeze(0).expand(B, P, T, C), None,
                                   out[:, 1:].unflatten(1, (P, T)), None)
        out[:, 0] = cached[2]
        return out

    def forward(self, x):
        return self.head(self.forward_features(x[0]))


class TimeSformer(nn.Module):                           # timesformer.py:335-351, cfg replaced by keywords, no download
    def __init__(self, img_size=224, patch_size=16, num_classes=400, num_frames=8,
                 attention_type='divided_space_time', **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.attention_type = attention_type
        self.model = VisionTransformer(img_size=img_size, num_classes=num_classes, patch_size=patch_size,
                                       embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                       norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_rate=0., attn_drop_rate=0.,
                                       drop_path_rate=0.1, num_frames=num_frames, attention_type=attention_type,
                                       **kwargs)
        fastlinear.install(self)                        # fp32 CUDA inference: linears on tome_linear_f32

    def forward(self, x):
        return self.model(x)
