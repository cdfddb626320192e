"""Proportional attention without a mask.

The reference adds ``size.log()`` of the key token to the attention logits
(tome/patch/videomae.py:62-63, vivit.py:103-104, timesformer.py:72-74).  Passed to fused attention as an
``attn_mask`` that costs 7x (a (B, 1, N, N) bias forces the masked kernels: 331 us instead of 46 us at
B=8, N=1568 on a B200).  A key-only bias is a rank-1 term, so it can ride inside the contraction:

    scale * q.k_j + b_j  ==  scale * [q, 1, 1] . [k_j, hi_j, lo_j],      hi_j + lo_j = b_j / scale

Each q/k head is padded from d to d + 8 channels by zero rows in a cached copy of the QKV weight (q's two
live spare channels are produced as 1 by the GEMM's bias), ``tome_attn_key_bias`` drops the two-term
split of ``log(size) / scale`` into k's spare channels, and cuDNN's flash attention runs unmasked
(d_qk = d + 8, d_v = d; 100 us at the shape above).  Inference, CUDA, bf16 only -- anything else takes
the reference's masked formulation in the callers.
"""
import os

import torch
import torch.nn.functional as F

PAD = 8          # spare channels per q/k head: keeps 16-byte alignment for bf16, two of them are used


def usable(x: torch.Tensor, module) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and not torch.is_grad_enabled() and not module.training)


def own_bf16_enabled() -> bool:
    return os.environ.get("TOME_ATTENTION_BF16", "0") == "1"


def _key_of(tensors):
    return tuple((t.data_ptr(), t._version, t.dtype, t.device) if t is not None else None for t in tensors)


def padded_qkv(owner, heads, d, wq, wk, wv, bq=None, bk=None, bv=None):
    """(weight, bias) of the padded QKV projection, cached on ``owner`` until a source tensor changes.
    Output columns: heads x (d + PAD) for q, the same for k, heads x d for v."""
    key = _key_of((wq, wk, wv, bq, bk, bv))
    cached = getattr(owner, "_tome_padded_qkv", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    c = wq.shape[1]
    da = d + PAD
    kw = dict(dtype=wq.dtype, device=wq.device)

    def pad_w(w):
        out = torch.zeros(heads, da, c, **kw)
        out[:, :d] = w.detach().reshape(heads, d, c)
        return out.reshape(heads * da, c)

    def pad_b(b, ones):
        out = torch.zeros(heads, da, **kw)
        if b is not None:
            out[:, :d] = b.detach().reshape(heads, d)
        if ones:
            out[:, d:d + 2] = 1
        return out.reshape(heads * da)

    weight = torch.cat((pad_w(wq), pad_w(wk), wv.detach()), 0).contiguous()
    bias = torch.cat((pad_b(bq, True), pad_b(bk, False),
                      bv.detach() if bv is not None else torch.zeros(heads * d, **kw)), 0).contiguous()
    owner._tome_padded_qkv = (key, weight, bias)
    return weight, bias


def attention(x, owner, heads, d, scale, log_size, wq, wk, wv, bq=None, bk=None, bv=None, lead=0, on_keys=None):
    """Attention over x (B, N, C) with ``log_size`` (B, N - lead[, 1]) fp32 added to the logits of the
    non-leading keys (and, when lead > 0, only for the non-leading queries).  ``on_keys(k)`` is called as
    soon as the key tensor exists.  Returns the context
    (B, N, heads * d) and the key tensor (B, heads, N, d) the matching metric is taken from."""
    from tome import _native
    B, N, _ = x.shape
    if d == 64 and N <= 256 and _native.frames_attention_usable(x, heads, N):
        # one key block: the library's own tcgen05 attention takes the key bias directly (tome_frames_attention with a
        # single "frame"), on the UNPADDED projection -- TimeSformer's 197-token spatial attention, the last ViViT layers
        key = _key_of((wq, wk, wv, bq, bk, bv))
        cached = getattr(owner, "_tome_plain_qkv", None)
        if cached is None or cached[0] != key:
            w = torch.cat((wq, wk, wv), 0).detach().contiguous()
            zb = lambda t, ref: t.detach() if t is not None else torch.zeros(ref.shape[0], dtype=ref.dtype, device=ref.device)
            bcat = None if (bq is None and bk is None and bv is None) else torch.cat((zb(bq, wq), zb(bk, wk), zb(bv, wv)), 0).contiguous()
            cached = owner._tome_plain_qkv = (key, w, bcat)
        qkv = F.linear(x, cached[1], cached[2])
        k = qkv[..., heads * d:2 * heads * d].view(B, N, heads, d).transpose(1, 2)
        if on_keys is not None:
            on_keys(k)
        kb = log_size.reshape(B, N - lead).float()
        if lead:
            kb = F.pad(kb, (lead, 0))
        ctx, _ = _native.frames_attention(qkv, heads, 1, scale, kb, want_diag=False, lead=0, unbiased_queries=lead)
        return ctx.view(B, N, heads * d), k
    if d == 64 and own_bf16_enabled():
        # opt-in (TOME_ATTENTION_BF16=1): the library's own bf16 flash attention takes the (B, N) key bias directly, on the
        # unpadded projection.  Correct at any length but slower than the folded-bias library call at these shapes (211 us
        # against 101 us at 8 x 12 x 1568: exp2-pipe bound), so the fold below stays the default.
        key = _key_of((wq, wk, wv, bq, bk, bv))
        cached = getattr(owner, "_tome_plain_qkv", None)
        if cached is None or cached[0] != key:
            w = torch.cat((wq, wk, wv), 0).detach().contiguous()
            zb = lambda t, ref: t.detach() if t is not None else torch.zeros(ref.shape[0], dtype=ref.dtype, device=ref.device)
            bcat = None if (bq is None and bk is None and bv is None) else torch.cat((zb(bq, wq), zb(bk, wk), zb(bv, wv)), 0).contiguous()
            cached = owner._tome_plain_qkv = (key, w, bcat)
        qkv = F.linear(x, cached[1], cached[2])
        if _native.attention_bf16_usable(qkv, heads):
            k = qkv[..., heads * d:2 * heads * d].view(B, N, heads, d).transpose(1, 2)
            if on_keys is not None:
                on_keys(k)
            kb = log_size.reshape(B, N - lead).float()
            if lead:
                kb = F.pad(kb, (lead, 0))
            return _native.attention_bf16(qkv, heads, scale, kb, unbiased_queries=lead), k
    da = d + PAD
    weight, bias = padded_qkv(owner, heads, d, wq, wk, wv, bq, bk, bv)
    qkv = F.linear(x, weight, bias)
    q = qkv[..., :heads * da].view(B, N, heads, da).transpose(1, 2)
    k = qkv[..., heads * da:2 * heads * da].view(B, N, heads, da).transpose(1, 2)
    v = qkv[..., 2 * heads * da:].view(B, N, heads, d).transpose(1, 2)
    if on_keys is not None:          # e.g. start the matching on K before the attention kernel is enqueued
        on_keys(k[..., :d])
    _native.attn_key_bias(log_size.reshape(B, N - lead), k, q if lead else None, d, scale, lead)
    ctx = F.scaled_dot_product_attention(q, k, v, scale=scale)
    return ctx.transpose(1, 2).reshape(B, N, heads * d), k[..., :d]


def usable_f32(x: torch.Tensor, module, heads: int, d: int) -> bool:
    """fp32 CUDA inference with 64-channel heads: QKV projection and attention on the exact-split tensor-core kernels."""
    from tome import _native
    return (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and not module.training and d == 64
            and x.dim() == 3 and x.shape[1] >= 64 and _native.linear_f32_usable(x, _PROBE.get(heads * d, x.device), None)
            and os.environ.get("TOME_ATTENTION_F32", "1") != "0")


class _Probe:
    """A (3C, C) fp32 meta-shaped stand-in so linear_f32_usable can vet shapes without a real weight."""
    def __init__(self):
        self.cache = {}

    def get(self, c, device):
        key = (c, str(device))
        if key not in self.cache:
            self.cache[key] = torch.empty(3 * c, c, dtype=torch.float32, device=device)
        return self.cache[key]


_PROBE = _Probe()


def attention_f32(x, owner, heads, d, scale, log_size, wq, wk, wv, bq=None, bk=None, bv=None, lead=0, on_keys=None, planes_for=None):
    """fp32 attention over x (B, N, C), ``log_size`` (B, N - lead[, 1]) or None added to the logits of the non-leading
    keys (for the non-leading queries only when lead > 0): one exact-split QKV GEMM (``tome_linear_f32`` on the cached
    concatenated weight, handing its result over as split planes) and ``tome_attention_f32``.  Returns (context (B, N,
    heads * d), keys (B, heads, N, d)); the context comes as ``Planes`` when ``planes_for`` -- the linear layer it feeds -- is
    an exact-split GEMM itself (no fp32 round trip, no second split)."""
    from tome import _native
    B, N, _ = x.shape
    key = _key_of((wq, wk, wv, bq, bk, bv))
    cached = getattr(owner, "_tome_plain_qkv", None)
    if cached is None or cached[0] != key:
        w = torch.cat((wq, wk, wv), 0).detach().contiguous()
        zb = lambda t, ref: t.detach() if t is not None else torch.zeros(ref.shape[0], dtype=ref.dtype, device=ref.device)
        bcat = None if (bq is None and bk is None and bv is None) else torch.cat((zb(bq, wq), zb(bk, wk), zb(bv, wv)), 0).contiguous()
        cached = owner._tome_plain_qkv = (key, w, bcat)
    qkv3 = None
    if _native.linear_f32_usable(x, cached[1], cached[2]):
        # planes only: writing the fp32 tensor from the same epilogue costs more than the exact sum of the K planes below
        qkv3 = _native.linear_f32(x, cached[1], cached[2], out="planes")
        k = _native.planes_to_f32(qkv3, heads * d, heads * d).view(B, N, heads, d).transpose(1, 2)
    else:
        qkv = F.linear(x, cached[1], cached[2])
        k = qkv[..., heads * d:2 * heads * d].view(B, N, heads, d).transpose(1, 2)
    if on_keys is not None:
        on_keys(k)
    kb = None
    if log_size is not None:
        kb = log_size.reshape(B, N - lead).float()
        if lead:
            kb = F.pad(kb, (lead, 0))
    as_planes = (planes_for is not None and type(planes_for).__name__ == "TomeLinear" and not planes_for.training
                 and _native.linear_f32_weight_ok(planes_for.weight, planes_for.bias))
    ctx = _native.attention_f32(qkv3 if qkv3 is not None else qkv, heads, scale, kb, unbiased_queries=lead if kb is not None else 0,
                                out="planes" if as_planes else "fp32")
    return ctx, k


def early_metric(module, k, head_aggregation="mean", frames=1):
    """The head-mean matching metric of K (videomae.py:72-73 ...), with match + select already started on
    a side stream when the module knows its block's reduction (``_tome_info``): the plan is then ready by
    the time the block asks for it (tome.merge.prefetch_matching)."""
    from tome.patch.videomae import lazy_head_mean
    from tome.merge import prefetch_matching
    metric = lazy_head_mean(k, frames)
    info = getattr(module, "_tome_info", None)
    if info is not None and head_aggregation == "mean" and info.get("mode") in ("merge", "drop", "hybrid"):
        rs = info.get("r")
        if isinstance(rs, list) and rs and rs[0] > 0:
            prefetch_matching(metric, rs[0], info["class_token"], info["distill_token"])
    return metric
