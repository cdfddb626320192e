"""``tome.patch.timesformer`` -- drop-in for the reference's tome/patch/timesformer.py.

Divided space-time attention: only the SPATIAL attention is patched (timesformer.py:224), the
matching runs per frame on the P patch tokens (batch (b t), class token excluded:
``class_token=False``, timesformer.py:203), the proportional-attention bias applies to
``attn[:, :, 1:, 1:]`` only (timesformer.py:74) and the metric is ``k.mean(1)[:, 1:, :]``
(timesformer.py:83).  Differences from the reference, all inside the hot path:
  * the two ``rearrange`` copies and the ``cat`` around the merge (timesformer.py:88-107) are
    addressing inside kernel 3 (``Merge.wavg_frames``);
  * modules are matched structurally, so the reference's slowfast model and
    ``hostmodels.timesformer`` both patch.
"""
import os

import torch
import torch.nn.functional as F

from tome.merge import (Drop, Merge, bipartite_soft_matching, bipartite_soft_matching_drop,
                        bipartite_soft_matching_hybrid, finish_source, merge_source, merge_wavg,
                        trace_source)
from tome import attention as prop_attention
from tome.patch.videomae import _normed_or, _swap, fusable_norm, lazy_head_mean
from tome.utils import parse_r


def _fused_block_ok(block, x):
    """The divided space-time block can run on the view-aware kernels: CUDA inference, fusable LayerNorms."""
    return (x.is_cuda and not torch.is_grad_enabled() and not block.training and x.is_contiguous()
            and os.environ.get("TOME_FUSED_BLOCKS", "1") != "0"
            and all(fusable_norm(getattr(block, n, None), x) is not None for n in ("temporal_norm1", "norm1", "norm2"))
            and isinstance(getattr(block, "temporal_fc", None), torch.nn.Linear))


def _ln(norm):
    return norm.weight, norm.bias, norm.eps


def _apply_spatial_residual(x, res_s, cls, B, T, P):
    """x (B, 1 + P*T, C) with a stale class row; res_s (B*T, 1 + P, C): the reference's rearrange + cat + add
    (timesformer.py:46-48) for the paths the merge kernel does not take (r == 0, foreign merge callables)."""
    C = x.size(2)
    out = torch.empty_like(x)
    out[:, 0] = cls
    out[:, 1:] = x[:, 1:] + res_s[:, 1:, :].reshape(B, T, P, C).transpose(1, 2).reshape(B, P * T, C)
    return out


def link_temporal(blocks, info):
    """id(block) -> temporal_norm1 of the block that follows it (the LayerNorm its closing add can feed)."""
    chain, seen = {}, set()
    blocks = list(blocks)
    for cur, nxt in zip(blocks[:-1], blocks[1:]):
        chain[id(cur)] = None if id(cur) in seen else getattr(nxt, "temporal_norm1", None)
        seen.add(id(cur))
    info["next_temporal_norm1"] = chain
    info["normed_t"] = None


class ToMeBlockMixin:
    """timesformer.py:12-57."""

    def forward(self, x, B, T, W):
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        P = (x.size(1) - 1) // T
        C = x.size(2)
        if self.attention_type in ['space_only', 'joint_space_time']:
            x = x + self.drop_path(self.attn(self.norm1(x))[0])
            return x + self.drop_path(self.mlp(self.norm2(x)))
        if _fused_block_ok(self, x):
            return self._forward_fused(x, B, T, P, C, info, attn_size, attn_bias)
        info["normed_t"] = None
        # temporal attention (un-patched): 'b (p t) m -> (b p) t m'
        xt = x[:, 1:, :].reshape(B * P, T, C)
        res_t = self.drop_path(self.temporal_attn(self.temporal_norm1(xt))).reshape(B, P * T, C)
        xt = x[:, 1:, :] + self.temporal_fc(res_t)
        # spatial attention: 'b (p t) m -> (b t) p m', class token replicated per frame
        init_cls = x[:, 0, :].unsqueeze(1)
        cls = init_cls.repeat(1, T, 1).reshape(B * T, 1, C)
        xs = xt.reshape(B, P, T, C).transpose(1, 2).reshape(B * T, P, C)
        res_s, metric = self.attn(self.norm1(torch.cat((cls, xs), 1)), attn_size, attn_bias)
        res_s = self.drop_path(res_s)
        cls = res_s[:, 0, :].reshape(B, T, C).mean(1, keepdim=True)          # averaged over frames
        res = res_s[:, 1:, :].reshape(B, T, P, C).transpose(1, 2).reshape(B, P * T, C)
        x = torch.cat((init_cls, xt), 1) + torch.cat((cls, res), 1)
        x = self.reduction_function(metric, x, info, B, T, P, norm=self.norm2)
        return x + self.drop_path(self.mlp(_normed_or(self.norm2, x, info)))

    def _forward_fused(self, x, B, T, P, C, info, attn_size, attn_bias):
        """The same block with every rearrange / cat / add / LayerNorm hop between the three row orders done as
        addressing inside tome_rows_add_layernorm and the merge kernel: per block the residual
        stream is read and written four times instead of ~fourteen."""
        from tome import _native
        x4 = x[:, 1:].unflatten(1, (P, T))                                   # (B, P, T, C) view of the patch tokens
        pre = info.pop("normed_t", None)
        if pre is not None and pre[0] is x and pre[1] is self.temporal_norm1:
            nt = pre[2]                                                      # made by the previous block's closing add
        else:
            nt = torch.empty(B, P, T, C, dtype=x.dtype, device=x.device)     # '(b p) t' rows for the temporal attention
            _native.rows_add_layernorm(x4, None, _ln(self.temporal_norm1), None, nt)
        res_t = self.temporal_attn(nt.view(B * P, T, C))
        tf = self.temporal_fc(res_t).view(B, P, T, C)
        # xt = x + temporal_fc(...) written into the residual stream's layout, norm1(xt) into '(b t) (1 + p)' rows
        xfull = torch.empty_like(x)                                          # class row filled by the merge (or below)
        ns = torch.empty(B * T, 1 + P, C, dtype=x.dtype, device=x.device)
        ns4 = ns.view(B, T, 1 + P, C)
        _native.rows_add_layernorm(x4, tf, _ln(self.norm1), xfull[:, 1:].unflatten(1, (P, T)), ns4[:, :, 1:].permute(0, 2, 1, 3))
        init_cls = x[:, 0]
        _native.cls_rows(init_cls, norm=_ln(self.norm1), normed_out=ns4[:, :, 0])   # class token replicated per frame
        res_s, metric = self.attn(ns, attn_size, attn_bias)
        cls = torch.empty(B, C, dtype=x.dtype, device=x.device)              # class token averaged over frames
        _native.cls_rows(init_cls, mean_src=res_s.view(B, T, 1 + P, C)[:, :, 0], sum_out=cls)
        x = self.reduction_function(metric, xfull, info, B, T, P, norm=self.norm2, residual=res_s, cls=cls)
        y = self.mlp(_normed_or(self.norm2, x, info))
        # x + mlp(...) and the NEXT block's temporal_norm1 in one pass
        nxt = (info.get("next_temporal_norm1") or {}).get(id(self))
        if nxt is None or fusable_norm(nxt, x) is None:
            return x + y
        Pn = (x.size(1) - 1) // T
        s = torch.empty_like(x)
        nt = torch.empty(B, Pn, T, C, dtype=x.dtype, device=x.device)
        _native.rows_add_layernorm(x[:, 1:].unflatten(1, (Pn, T)), y[:, 1:].unflatten(1, (Pn, T)), _ln(nxt),
                                   s[:, 1:].unflatten(1, (Pn, T)), nt)
        _native.cls_rows(x[:, 0], add=y[:, 0], sum_out=s[:, 0])
        info["normed_t"] = (s, nxt, nt)
        return s


class ToMeAttentionMixin:
    """timesformer.py:60-83."""

    def forward(self, x, size: torch.Tensor = None, log_size: torch.Tensor = None):
        B, N, C = x.shape
        if size is not None and self.with_qkv and prop_attention.usable(x, self):
            # attn[:, :, 1:, 1:] += log(size) (timesformer.py:72-74) folded into the contraction: the class key
            # carries bias 0 and the class query's bias channels are 0
            if log_size is None:
                log_size = size.log()
            wq, wk, wv = self.qkv.weight.chunk(3, 0)
            bq, bk, bv = self.qkv.bias.chunk(3, 0) if self.qkv.bias is not None else (None, None, None)
            early = {}
            x, k = prop_attention.attention(
                x, self, self.num_heads, C // self.num_heads, self.scale, log_size.float(), wq, wk, wv, bq, bk, bv, lead=1,
                on_keys=lambda keys: early.update(metric=prop_attention.early_metric(self, keys[:, :, 1:, :])))
            x = self.proj_drop(self.proj(x))
            return x, early["metric"]
        if self.with_qkv and prop_attention.usable_f32(x, self, self.num_heads, C // self.num_heads):
            # fp32 inference: exact-split tensor-core QKV GEMM + attention; the class query takes no bias (timesformer.py:74)
            if size is not None and log_size is None:
                log_size = size.log()
            wq, wk, wv = self.qkv.weight.chunk(3, 0)
            bq, bk, bv = self.qkv.bias.chunk(3, 0) if self.qkv.bias is not None else (None, None, None)
            early = {}
            x, k = prop_attention.attention_f32(
                x, self, self.num_heads, C // self.num_heads, self.scale, None if size is None else log_size, wq, wk, wv, bq, bk, bv,
                lead=1, on_keys=lambda keys: early.update(metric=prop_attention.early_metric(self, keys[:, :, 1:, :])),
                planes_for=self.proj)
            return self.proj_drop(self.proj(x)), early["metric"]
        if self.with_qkv:
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]
        else:
            q = k = v = x.reshape(B, N, self.num_heads, C // self.num_heads).permute(0, 2, 1, 3)
        metric = prop_attention.early_metric(self, k[:, :, 1:, :])          # k.mean(1)[:, 1:, :], matching started
        bias = None
        if size is not None:       # attn[:, :, 1:, 1:] += log(size): neither the cls query nor the cls key is biased
            if log_size is None:
                log_size = size.log()
            row = F.pad(log_size[..., 0].to(q.dtype), (1, 0))                  # (B, N), 0 for the cls key
            bias = row[:, None, None, :].expand(B, 1, N, N).clone()
            bias[:, :, 0, :] = 0
        drop = self.attn_drop.p if self.training else 0.0
        x = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=drop, scale=self.scale)
        x = x.transpose(1, 2).reshape(B, N, C)
        if self.with_qkv:
            x = self.proj_drop(self.proj(x))
        return x, metric


def _frames_view(x, B, T, P):
    """'b (p t) m -> (b t) p m' of the non-class tokens (only needed off the fused path)."""
    return x[:, 1:, :].reshape(B, P, T, -1).transpose(1, 2).reshape(B * T, P, -1)


def _frames_back(cls, y, B, T):
    Pn = y.size(1)
    return torch.cat((cls, y.reshape(B, T, Pn, -1).transpose(1, 2).reshape(B, Pn * T, -1)), dim=1)


def _merge_frames_generic(merge, x, info, B, T, P):
    cls, merged_x = x[:, 0:1, :], _frames_view(x, B, T, P)
    if info["trace_source"]:
        info["source"] = trace_source(merge, merged_x, info["source"])
    merged_x, info["size"] = merge_wavg(merge, merged_x, info["size"])
    info["log_size"] = None
    return _frames_back(cls, merged_x, B, T)


def timesformer_merge(metric, x, _tome_info, B, T, num_spatial_tokens, norm=None, residual=None, cls=None):
    """timesformer.py:85-109.  ``residual`` / ``cls``: the pending spatial-attention add (timesformer.py:46-48),
    taken inside the merge kernel when it runs."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                           _tome_info["mode"])
        pre_merge = num_spatial_tokens
        if isinstance(merge, Merge):
            if _tome_info["trace_source"]:
                _tome_info["source"] = trace_source(merge, None, _tome_info["source"])
            fn = fusable_norm(norm, x) if norm is not None else None
            res = merge.wavg_frames(x, T, _tome_info["size"], norm=fn, residual=residual, cls=cls)
            x, _tome_info["size"], _tome_info["log_size"] = res[0], res[1], res[2]
            _tome_info["normed"] = res[3] if fn is not None else None
        else:
            if residual is not None:
                x = _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens)
            x = _merge_frames_generic(merge, x, _tome_info, B, T, num_spatial_tokens)
        if _tome_info['verbose']:
            print(f'Merged {pre_merge} to {(x.size(1) - 1) // T} tokens')
    elif residual is not None:
        x = _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens)
    return x


def timesformer_drop(metric, x, _tome_info, B, T, num_spatial_tokens, norm=None, residual=None, cls=None):
    """timesformer.py:112-140."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r <= 0 and residual is not None:
        return _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens)
    if r > 0:
        drop = bipartite_soft_matching_drop(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                            _tome_info["mode"])
        if residual is not None and not isinstance(drop, Drop):
            x, residual = _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens), None
        if isinstance(drop, tuple):
            return x
        if _tome_info["trace_source"]:
            _tome_info["source"] = trace_source(drop, x.new_empty((B * T, num_spatial_tokens, 0)), _tome_info["source"], drop=True)
        pre_drop = num_spatial_tokens
        if isinstance(drop, Drop):
            x = drop.frames(x, T, residual=residual, cls=cls)
        else:
            x = _frames_back(x[:, 0:1, :], drop(_frames_view(x, B, T, num_spatial_tokens)), B, T)
        Pn = (x.size(1) - 1) // T
        _tome_info["size"] = torch.ones((B * T, Pn, 1), device=x.device)
        _tome_info["log_size"] = torch.zeros((B * T, Pn, 1), device=x.device)
        if _tome_info['verbose']:
            print(f'Dropped {pre_drop} to {Pn} tokens')
    return x


def timesformer_hybrid(metric, x, _tome_info, B, T, num_spatial_tokens, norm=None, residual=None, cls=None):
    """timesformer.py:143-167."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching_hybrid(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                                  _tome_info["mode"], _tome_info["threshold"])
        pre_merge = num_spatial_tokens
        if isinstance(merge, Merge):
            if _tome_info["trace_source"]:
                _tome_info["source"] = trace_source(merge, None, _tome_info["source"])
            fn = fusable_norm(norm, x) if norm is not None else None
            res = merge.wavg_frames(x, T, _tome_info["size"], norm=fn, residual=residual, cls=cls)
            x, _tome_info["size"], _tome_info["log_size"] = res[0], res[1], res[2]
            _tome_info["normed"] = res[3] if fn is not None else None
        else:
            if residual is not None:
                x = _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens)
            x = _merge_frames_generic(merge, x, _tome_info, B, T, num_spatial_tokens)
        if _tome_info['verbose']:
            print(f'Merged {pre_merge} to {(x.size(1) - 1) // T} tokens')
    elif residual is not None:
        x = _apply_spatial_residual(x, residual, cls, B, T, num_spatial_tokens)
    return x


def _is_block(m):
    return all(hasattr(m, a) for a in ("norm1", "attn", "norm2", "mlp", "attention_type")) and hasattr(m.attn, "with_qkv")


def apply_duplicate_patch(model, layer_to_duplicate, quantity):
    """timesformer.py:170-172: the SAME module is inserted again (shared weights)."""
    for i in range(layer_to_duplicate + 1, layer_to_duplicate + quantity):
        model.model.blocks.insert(index=i, module=model.model.blocks[layer_to_duplicate])


def make_tome_class(transformer_class):
    class ToMeVisionTransformer(transformer_class):
        def forward(self, *args, **kwdargs) -> torch.Tensor:
            self._tome_info["r"] = parse_r(len(self.model.blocks), self.r)
            self._tome_info["size"] = None
            self._tome_info["log_size"] = None
            self._tome_info["normed"] = None
            self._tome_info["source"] = None
            link_temporal(self.model.blocks, self._tome_info)
            out = super().forward(*args, **kwdargs)
            finish_source(self._tome_info)          # compact source map -> the reference's dense matrix, once
            return out

    return ToMeVisionTransformer


def apply_patch(model_wrapper, trace_source: bool = False, prop_attn: bool = True, mode: str = 'merge',
                head_aggregation: str = 'mean', threshold: float = 0.0, verbose: bool = False):
    """timesformer.py:187-224.  ``head_aggregation`` is accepted and ignored, as in the reference."""
    model = model_wrapper.model
    if not getattr(model_wrapper.__class__, "_tome_wrapper", False):
        cls = make_tome_class(model_wrapper.__class__)
        cls._tome_wrapper = True
        model_wrapper.__class__ = cls
    model_wrapper.r = 0
    model_wrapper._tome_info = {
        "r": model_wrapper.r,
        "size": None,
        "log_size": None,
        "normed": None,
        "source": None,
        "trace_source": trace_source,
        "prop_attn": prop_attn,
        "verbose": verbose,
        "class_token": False,
        "distill_token": False,
        "mode": mode,
        "threshold": threshold,
    }
    if hasattr(model, "dist_token") and model.dist_token is not None:
        model_wrapper._tome_info["distill_token"] = True

    if mode in ['merge', 'random_merge']:
        reduction_function = timesformer_merge
    elif mode in ['drop', 'random_drop']:
        reduction_function = timesformer_drop
    elif mode in ['hybrid']:
        reduction_function = timesformer_hybrid
    else:
        raise ValueError(f"unknown ToMe mode {mode!r}")

    for module in model.modules():
        if _is_block(module):
            _swap(module, ToMeBlockMixin, "ToMe")
            module._tome_info = model_wrapper._tome_info
            module.reduction_function = reduction_function
            _swap(module.attn, ToMeAttentionMixin, "ToMe")      # spatial attention only (timesformer.py:224)
            module.attn._tome_info = model_wrapper._tome_info
