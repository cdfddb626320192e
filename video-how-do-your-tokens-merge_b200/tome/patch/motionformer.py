"""``tome.patch.motionformer`` -- drop-in for the reference's tome/patch/motionformer.py.

Trajectory attention over the joint (frame, patch) keys with the proportional-attention bias laid
out ``(b f) s i -> b (s f) i`` (motionformer.py:107-111); the metric is the head-mean of K
regrouped ``(b h) (s f) d -> (b f) h s d`` (motionformer.py:143-144) and the merge runs per "frame"
on ``'b (s f) d -> (b f) s d'`` (motionformer.py:150-151).  Reference quirk kept verbatim: the
attention reads the token axis as (f n) while metric and merge read the same axis as (s f).
The rearranges around the merge are addressing inside kernel 3 (``Merge.wavg_frames``)."""
import torch
import torch.nn.functional as F

from tome.merge import (Drop, Merge, bipartite_soft_matching, bipartite_soft_matching_drop,
                        bipartite_soft_matching_hybrid, finish_source, trace_source)
from tome.patch.timesformer import _frames_back, _frames_view, _merge_frames_generic
from tome import attention as prop_attention
from tome.patch.videomae import _close_block, _norm1_or, _normed_or, _swap, fusable_norm, lazy_head_mean, link_blocks
from tome.utils import parse_r


def _weight_rows(mod, tag, w, lo, hi):
    """A row slice of ``w`` as one long-lived tensor object, so that tome_linear_f32's per-weight plane cache (keyed by the
    tensor's identity) holds: a fresh ``w[lo:hi]`` view per call would re-split the weight every time."""
    cache = mod.__dict__.setdefault("_tome_weight_rows", {})
    hit = cache.get(tag)
    if hit is None or hit[0] is not w or hit[1] != w._version:
        hit = cache[tag] = (w, w._version, w[lo:hi])
    return hit[2]


def trajectory_attention(mod, x, num_frames, log_size=None, on_keys=None):
    """vit_helper.py:146-267 (approx == 'none').  x (B, 1 + F*P, C), tokens '(f n)'.
    ``log_size`` (B, F*P) adds the proportional-attention key bias in the flat key order
    (tome/patch/motionformer.py:105-112).  Returns (out, k_) with k_ (B, h, F*P, d)."""
    B, N, C = x.shape
    Fr, h = num_frames, mod.num_heads
    P = (N - 1) // Fr
    d = C // h
    from tome import _native
    if d == 64 and _native.frames_attention_usable(x, h, P) and N == 1 + Fr * P and Fr <= 32:
        # bf16 inference: both stages on the library's kernels (tome_frames_attention on tcgen05, tome_traj_temporal),
        # q / k / v read in place from the QKV GEMM's output
        qkv = mod.qkv(x)
        q, k, v = qkv.view(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
        if on_keys is not None:
            on_keys(k[:, :, 1:])
        cls_out = F.scaled_dot_product_attention(q[:, :, 0:1], k, v, scale=mod.scale).transpose(1, 2).reshape(B, 1, C)
        xs, x_diag = _native.frames_attention(qkv, h, Fr, mod.scale, log_size)
        q2 = mod.proj_q(x_diag)
        wkv, bkv = mod.proj_kv.weight, mod.proj_kv.bias
        k2 = F.linear(xs, wkv[:C], None if bkv is None else bkv[:C])
        vals = xs if mod.use_original_code else F.linear(xs, wkv[C:], None if bkv is None else bkv[C:])
        out = _native.traj_temporal(q2, k2, vals, h, mod.scale)
        out = mod.proj_drop(mod.proj(torch.cat((cls_out, out), dim=1)))
        return out, k[:, :, 1:]
    if (d == 64 and _native.frames_attention_f32_usable(x, h) and N == 1 + Fr * P and Fr <= 32
            and _native.linear_f32_usable(x, mod.qkv.weight, mod.qkv.bias) and _native.linear_f32_weight_ok(mod.proj_kv.weight[:C], None)):
        # fp32 inference (the reference benchmark's arithmetic): both stages on the exact-split tensor-core kernels
        # (tome_frames_attention_f32: tome_attention_f32 with one problem per frame; fp32 tome_traj_temporal); the QKV GEMM
        # hands its result over as split planes, the space stage hands xs over as planes to the K projection
        qkv = _native.linear_f32(x, mod.qkv.weight, mod.qkv.bias)          # fp32 for the class row and the metric, and its split
        qkv3 = _native.Planes(_native.split3(qkv.view(B * N, 3 * C)), (B, N, 3 * C))    # (one epilogue writing both is slower)
        q, k, v = qkv.view(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
        if on_keys is not None:
            on_keys(k[:, :, 1:])
        cls_out = _native.cls_attention(qkv, h, mod.scale)                 # the class token attends to everything, unbiased
        xs, xs3, x_diag = _native.frames_attention_f32(qkv3, h, Fr, mod.scale, log_size)
        q2 = mod.proj_q(x_diag)
        wkv, bkv = mod.proj_kv.weight, mod.proj_kv.bias
        k2 = _native.linear_f32(xs3, _weight_rows(mod, "k", wkv, 0, C), None if bkv is None else bkv[:C])
        if mod.use_original_code:
            vals = xs
        else:
            vals = _native.linear_f32(xs3, _weight_rows(mod, "v", wkv, C, 2 * C), None if bkv is None else bkv[C:])
        out = _native.traj_temporal(q2, k2, vals, h, mod.scale)
        out = mod.proj_drop(mod.proj(torch.cat((cls_out, out), dim=1)))
        return out, k[:, :, 1:]
    q, k, v = mod.qkv(x).reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)        # each (B, h, N, d)
    if on_keys is not None:              # K exists: the matching can start beside the attention kernels
        on_keys(k[:, :, 1:])
    cls_out = F.scaled_dot_product_attention(q[:, :, 0:1], k, v, scale=mod.scale)  # cls attends to everything
    cls_out = cls_out.transpose(1, 2).reshape(B, 1, C)
    q_, k_, v_ = q[:, :, 1:], k[:, :, 1:], v[:, :, 1:]
    # space attention: every query against the P keys of each frame, softmax per frame
    kf = k_.reshape(B, h, Fr, P, d)
    vf = v_.reshape(B, h, Fr, P, d)
    qe = q_[:, :, None].expand(B, h, Fr, Fr * P, d)
    bias = None
    if log_size is not None:
        bias = log_size.reshape(B, 1, Fr, 1, P).to(q.dtype).expand(B, h, Fr, Fr * P, P)
    xs = F.scaled_dot_product_attention(qe, kf, vf, attn_mask=bias, scale=mod.scale)   # (B, h, F, S, d)
    xs = xs.permute(0, 3, 2, 1, 4).reshape(B, Fr * P, Fr, C)                            # 'b s f (h d)'
    # temporal attention: the query is the trajectory token of the query's own frame
    g = torch.arange(Fr, device=x.device).repeat_interleave(P)                          # frame of token s
    x_diag = xs[:, torch.arange(Fr * P, device=x.device), g]                             # (B, S, C)
    q2 = mod.proj_q(x_diag).reshape(B, Fr * P, h, d).transpose(1, 2) * mod.scale       # (B, h, S, d)
    k2, v2 = mod.proj_kv(xs).chunk(2, dim=-1)
    k2 = k2.reshape(B, Fr * P, Fr, h, d).permute(0, 3, 1, 2, 4)                        # (B, h, S, F, d)
    attn = torch.einsum('bhsd,bhsfd->bhsf', q2, k2).softmax(dim=-1)
    vals = xs if mod.use_original_code else v2                                          # the v = x "typo" kept by default
    vals = vals.reshape(B, Fr * P, Fr, h, d).permute(0, 3, 1, 2, 4)
    out = torch.einsum('bhsf,bhsfd->bhsd', attn, vals).transpose(1, 2).reshape(B, Fr * P, C)
    out = mod.proj_drop(mod.proj(torch.cat((cls_out, out), dim=1)))
    return out, k_


class ToMeBlockMixin:
    """motionformer.py:14-30."""

    def forward(self, x, seq_len=196, num_frames=8, approx='none', num_landmarks=128):
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        attn_out, _, metric = self.attn(_norm1_or(self, x, info), seq_len=seq_len, num_frames=num_frames, approx=approx,
                                        num_landmarks=num_landmarks, size=attn_size, log_size=attn_bias)
        # x = x + attn_out is taken inside the merge kernel, x + mlp(...) together with the next block's norm1
        # (tome_add_layernorm), when they can be (CUDA inference): SURVEY.md 8f-f2
        x = self.reduction_function(metric, x, info, num_frames, norm=self.norm2, residual=self.drop_path(attn_out))
        return _close_block(self, x, self.mlp(_normed_or(self.norm2, x, info)), info)


class ToMeTrajectoryAttentionMixin:
    """motionformer.py:33-144 (full attention branch; the approximations are unused by ToMe configs)."""

    def forward(self, x, seq_len=196, num_frames=8, approx='none', num_landmarks=128,
                size: torch.Tensor = None, log_size: torch.Tensor = None):
        if approx != 'none':
            raise NotImplementedError("tome.patch.motionformer: approximate attention is not supported")
        B, N, C = x.shape
        Fr = num_frames
        S = (N - 1) // Fr
        flat = None
        if size is not None:
            if log_size is None:
                log_size = size.log()
            # '(b f) s i -> b (s f) i': key j of the flat token axis gets log size[(b, j % F), j // F]
            flat = log_size[..., 0].reshape(B, Fr, S).transpose(1, 2).reshape(B, S * Fr)
        # metric: '(b h) (s f) d -> (b f) h s d' then mean over heads (motionformer.py:143-144)
        early = {}
        out, k_ = trajectory_attention(self, x, Fr, flat, on_keys=lambda keys: early.update(
            metric=prop_attention.early_metric(self, keys, frames=Fr)))
        return out, None, early["metric"]


def _residual_in_kernel(x, residual):
    return (residual is not None and x.is_cuda and not torch.is_grad_enabled() and residual.shape == x.shape
            and residual.dtype == x.dtype and x.dtype in (torch.float32, torch.bfloat16))


def motionformer_merge(metric, x, _tome_info, num_frames, norm=None, residual=None):
    """motionformer.py:147-170.  ``residual``: the pending ``x + attn_out`` (motionformer.py:24), added inside the merge kernel
    when it runs."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if residual is not None and not (r > 0 and _residual_in_kernel(x, residual)):
        x, residual = x + residual, None
    if r > 0:
        B, T = x.size(0), num_frames
        P = (x.size(1) - 1) // T
        merge, _ = bipartite_soft_matching(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                           _tome_info["mode"])
        if isinstance(merge, Merge):
            if _tome_info["trace_source"]:
                _tome_info["source"] = trace_source(merge, None, _tome_info["source"])
            fn = fusable_norm(norm, x) if norm is not None else None
            res = merge.wavg_frames(x, T, _tome_info["size"], norm=fn, residual=residual)
            x, _tome_info["size"], _tome_info["log_size"] = res[0], res[1], res[2]
            _tome_info["normed"] = res[3] if fn is not None else None
        else:
            if residual is not None:
                x = x + residual
            x = _merge_frames_generic(merge, x, _tome_info, B, T, P)
        if _tome_info['verbose']:
            print(f'Merged {P} to {(x.size(1) - 1) // T} tokens')
    return x


def motionformer_drop(metric, x, _tome_info, num_frames, norm=None, residual=None):
    """motionformer.py:173-200."""
    _tome_info["normed"] = None
    if residual is not None:
        x = x + residual
    r = _tome_info["r"].pop(0)
    if r > 0:
        B, T = x.size(0), num_frames
        P = (x.size(1) - 1) // T
        drop = bipartite_soft_matching_drop(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                            _tome_info["mode"])
        if isinstance(drop, tuple):
            return x
        if _tome_info["trace_source"]:
            _tome_info["source"] = trace_source(drop, x.new_empty((B * T, P, 0)), _tome_info["source"], drop=True)
        if isinstance(drop, Drop):
            x = drop.frames(x, T)
        else:
            x = _frames_back(x[:, 0:1, :], drop(_frames_view(x, B, T, P)), B, T)
        Pn = (x.size(1) - 1) // T
        _tome_info["size"] = torch.ones((B * T, Pn, 1), device=x.device)
        _tome_info["log_size"] = torch.zeros((B * T, Pn, 1), device=x.device)
        if _tome_info['verbose']:
            print(f'Dropped {P} to {Pn} tokens')
    return x


def motionformer_hybrid(metric, x, _tome_info, num_frames, norm=None, residual=None):
    """motionformer.py:203-227."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if residual is not None and not (r > 0 and _residual_in_kernel(x, residual)):
        x, residual = x + residual, None
    if r > 0:
        B, T = x.size(0), num_frames
        P = (x.size(1) - 1) // T
        merge, _ = bipartite_soft_matching_hybrid(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                                  _tome_info["mode"], _tome_info["threshold"])
        if isinstance(merge, Merge):
            if _tome_info["trace_source"]:
                _tome_info["source"] = trace_source(merge, None, _tome_info["source"])
            fn = fusable_norm(norm, x) if norm is not None else None
            res = merge.wavg_frames(x, T, _tome_info["size"], norm=fn, residual=residual)
            x, _tome_info["size"], _tome_info["log_size"] = res[0], res[1], res[2]
            _tome_info["normed"] = res[3] if fn is not None else None
        else:
            if residual is not None:
                x = x + residual
            x = _merge_frames_generic(merge, x, _tome_info, B, T, P)
        if _tome_info['verbose']:
            print(f'Merged {P} to {(x.size(1) - 1) // T} tokens')
    return x


def _is_block(m):
    return all(hasattr(m, a) for a in ("norm1", "attn", "norm2", "mlp")) and _is_traj_attention(m.attn)


def _is_traj_attention(m):
    return all(hasattr(m, a) for a in ("qkv", "proj_q", "proj_kv", "proj", "use_original_code"))


def apply_duplicate_patch(model, layer_to_duplicate, quantity):
    """motionformer.py:230-232."""
    for i in range(layer_to_duplicate + 1, layer_to_duplicate + quantity):
        model.blocks.insert(index=i, module=model.blocks[layer_to_duplicate])


def make_tome_class(transformer_class):
    class ToMeVisionTransformer(transformer_class):
        def forward(self, *args, **kwdargs) -> torch.Tensor:
            self._tome_info["r"] = parse_r(len(self.blocks), self.r)
            self._tome_info["size"] = None
            self._tome_info["log_size"] = None
            self._tome_info["normed"] = None
            self._tome_info["source"] = None
            link_blocks(self.blocks, self._tome_info)
            out = super().forward(*args, **kwdargs)
            finish_source(self._tome_info)          # compact source map -> the reference's dense matrix, once
            return out

    return ToMeVisionTransformer


def apply_patch(model, trace_source: bool = False, prop_attn: bool = True, mode: str = 'merge',
                head_aggregation: str = 'mean', threshold: float = 0.0, verbose: bool = False):
    """motionformer.py:247-284 -- the model itself is patched (no ``.model`` wrapper)."""
    if not getattr(model.__class__, "_tome_wrapper", False):
        cls = make_tome_class(model.__class__)
        cls._tome_wrapper = True
        model.__class__ = cls
    model.r = 0
    model._tome_info = {
        "r": model.r,
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
        model._tome_info["distill_token"] = True

    if mode in ['merge', 'random_merge']:
        reduction_function = motionformer_merge
    elif mode in ['drop', 'random_drop']:
        reduction_function = motionformer_drop
    elif mode in ['hybrid']:
        reduction_function = motionformer_hybrid
    else:
        raise ValueError(f"unknown ToMe mode {mode!r}")

    for module in model.modules():
        if _is_block(module):
            _swap(module, ToMeBlockMixin, "ToMe")
            module._tome_info = model._tome_info
            module.reduction_function = reduction_function
        elif _is_traj_attention(module):
            _swap(module, ToMeTrajectoryAttentionMixin, "ToMe")
            module._tome_info = model._tome_info
