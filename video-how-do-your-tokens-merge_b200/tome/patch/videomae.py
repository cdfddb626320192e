"""``tome.patch.videomae`` -- drop-in for the reference's tome/patch/videomae.py.

Same entry points and semantics (apply_patch, apply_duplicate_patch, model.r, the shared
``_tome_info`` dict with the reference's keys), but
  * modules are matched structurally (class name + attributes) instead of by
    ``isinstance(slowfast...Block)``, so both the reference's slowfast VideoMAE and
    ``hostmodels.videomae`` can be patched;
  * the per-block reduction calls the sm_100a kernels: one match, one select, ONE fused
    merge_wavg pass that also emits the new token sizes and log(size)
    (reference: videomae.py:80-151 -> ~40 ATen launches);
  * proportional attention reads the kernel-emitted ``_tome_info['log_size']`` instead of
    recomputing ``size.log()`` (videomae.py:62-63) and goes through fused SDPA.
"""
import copy

import torch
import torch.nn.functional as F

from tome.merge import (Merge, bipartite_soft_matching, bipartite_soft_matching_drop,
                        bipartite_soft_matching_hybrid, finish_source, merge_source, merge_wavg,
                        trace_source)
from tome.utils import parse_r
from tome import attention as prop_attention


def lazy_head_mean(k, frames=1):
    """``k.mean(1)`` (videomae.py:72-73) -- on the CUDA inference path left to kernel 1's prologue
    (``tome_match_heads``), which reads K once and never writes the metric out."""
    from tome import _native
    if k.is_cuda and not torch.is_grad_enabled() and k.dtype in (torch.float32, torch.bfloat16) and k.stride(-1) == 1:
        return _native.HeadMeanMetric(k, frames)
    return _native.HeadMeanMetric(k, frames).materialize()


def fusable_norm(norm, x):
    """(weight, bias, eps) when ``norm`` is a LayerNorm kernel 3 can apply in its own pass (inference,
    affine, over the channel axis, same dtype/device as x); None otherwise."""
    if (isinstance(norm, torch.nn.LayerNorm) and not norm.training and not torch.is_grad_enabled()
            and norm.elementwise_affine and tuple(norm.normalized_shape) == (x.shape[-1],) and x.is_cuda
            and norm.weight.dtype == x.dtype and norm.weight.device == x.device
            and x.dtype in (torch.float32, torch.bfloat16)
            and x.shape[-1] % (8 if x.dtype == torch.bfloat16 else 4) == 0
            and x.shape[-1] <= (2048 if x.dtype == torch.bfloat16 else 1024)):
        return norm.weight, norm.bias, norm.eps
    return None


def _normed_or(norm, x, info):
    """LayerNorm(x): taken from the merge kernel's fused output when it produced one for this x."""
    y = info.pop("normed", None)
    return y if y is not None and y.shape == x.shape else norm(x)


def _norm1_or(block, x, info):
    """norm1(x): taken from the previous block's fused residual-add + LayerNorm when it was made for
    exactly this tensor and this LayerNorm."""
    pre = info.pop("normed1", None)
    if pre is not None and pre[0] is x and pre[1] is block.norm1:
        return pre[2]
    return block.norm1(x)


def _close_block(block, x, y, info):
    """``x + y`` at the end of a block (videomae.py:22).  On the CUDA inference path the LayerNorm that opens
    the NEXT block (its norm1) is applied in the same pass (``tome_add_layernorm``) and parked in
    ``info["normed1"]``; the chain block -> next block's norm1 is rebuilt at every model forward."""
    nxt = (info.get("next_norm1") or {}).get(id(block))
    fn = fusable_norm(nxt, x) if nxt is not None else None
    if fn is None or y.shape != x.shape or y.dtype != x.dtype or block.training:       # eval: drop_path is identity
        return x + block.drop_path(y)
    from tome import _native
    s, normed = _native.add_layernorm(x, y, fn)
    info["normed1"] = (s, nxt, normed)
    return s


def link_blocks(blocks, info):
    """id(block) -> the LayerNorm the following block opens with; blocks listed twice are left out."""
    chain, seen = {}, set()
    blocks = list(blocks)
    for cur, nxt in zip(blocks[:-1], blocks[1:]):
        if id(cur) in seen:
            chain[id(cur)] = None
        else:
            chain[id(cur)] = getattr(nxt, "norm1", None)
        seen.add(id(cur))
    info["next_norm1"] = chain
    info["normed1"] = None


class ToMeBlockMixin:
    """videomae.py:13-30."""

    def forward(self, x):
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        attn, metric = self.attn(_norm1_or(self, x, info), attn_size, info["head_aggregation"], attn_bias)
        if self.gamma_1 is None:
            # x = x + attn, then the reduction: the add is taken inside the merge kernel when it can be
            x = self.reduction_function(metric, x, info, norm=self.norm2, residual=self.drop_path(attn))
            x = _close_block(self, x, self.mlp(_normed_or(self.norm2, x, info)), info)
        else:
            x = x + self.drop_path(self.gamma_1 * attn)
            x = self.reduction_function(metric, x, info, norm=self.norm2)
            x = x + self.drop_path(self.gamma_2 * self.mlp(_normed_or(self.norm2, x, info)))
        return x


class ToMeDuplicateBlockMixin:
    """videomae.py:33-44: attention-only replica used by the layer-duplication ablation."""

    def forward(self, x):
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        _, metric = self.attn(_norm1_or(self, x, info), attn_size, info["head_aggregation"], attn_bias)
        return self.reduction_function(metric, x, info)


class ToMeAttentionMixin:
    """videomae.py:47-77: attention that also returns the matching metric."""

    def forward(self, x, size: torch.Tensor = None, head_aggregation: str = 'mean', log_size: torch.Tensor = None):
        B, N, C = x.shape
        qkv_bias = None
        if self.q_bias is not None:
            if self.training or torch.is_grad_enabled():
                qkv_bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias, requires_grad=False), self.v_bias))
            else:
                # inference: build (q_bias, 0, v_bias) once instead of a fill + cat per block per forward
                key = (self.q_bias.data_ptr(), self.v_bias.data_ptr(), self.q_bias._version, self.v_bias._version,
                       self.q_bias.dtype, self.q_bias.device)
                if getattr(self, "_tome_qkv_bias_key", None) != key:
                    self._tome_qkv_bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias), self.v_bias)).detach()
                    self._tome_qkv_bias_key = key
                qkv_bias = self._tome_qkv_bias
        early = {}

        def on_keys(keys):               # K exists: start match + select beside the attention kernel
            if head_aggregation == 'mean':
                early["metric"] = prop_attention.early_metric(self, keys)

        if size is not None and prop_attention.usable(x, self):
            # proportional attention (videomae.py:62-63) with the key bias folded into the contraction
            if log_size is None:
                log_size = size.log()
            wq, wk, wv = self.qkv.weight.chunk(3, 0)
            d = wq.shape[0] // self.num_heads
            x, k = prop_attention.attention(x, self, self.num_heads, d, self.scale, log_size.float(), wq, wk, wv,
                                            self.q_bias, None, self.v_bias, on_keys=on_keys)
        else:
            from tome import _native
            if size is not None and log_size is None:   # proportional attention (videomae.py:62-63)
                log_size = size.log()
            kb = None if size is None else log_size[..., 0]
            qkv3 = None
            if _native.linear_f32_usable(x, self.qkv.weight, qkv_bias) and _native.attention_f32_planes_ok(x, self.qkv.weight, self.num_heads):
                # fp32 inference: the QKV GEMM (tome_linear_f32) hands its result over as split planes ONLY; the fp32 keys the
                # matching metric needs are the exact sum of their three planes (tome_planes_sum on the K third)
                qkv3 = _native.linear_f32(x, self.qkv.weight, qkv_bias, out="planes")
                C1 = self.qkv.weight.shape[0] // 3
                k = _native.planes_to_f32(qkv3, C1, C1).view(B, N, self.num_heads, -1).transpose(1, 2)
                on_keys(k)
                keep = type(self.proj).__name__ == "TomeLinear"           # proj takes the planes directly
                x = _native.attention_f32(qkv3, self.num_heads, self.scale, kb, out="planes" if keep else "fp32")
                x = self.proj_drop(self.proj(x))
                if head_aggregation == 'mean':
                    return x, early["metric"]
                if head_aggregation == 'concat':
                    return x, k.transpose(1, 2).reshape(B, N, -1)
                raise ValueError(f"head_aggregation must be 'mean' or 'concat', got {head_aggregation!r}")
            qkv_flat = _native.linear(x, self.qkv.weight, qkv_bias) if x.is_cuda else F.linear(x, self.qkv.weight, qkv_bias)
            qkv = qkv_flat.reshape(B, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]
            on_keys(k)
            if q.shape[-1] == 64 and _native.attention_f32_usable(qkv_flat, self.num_heads, kb):
                # fp32 inference: exact-split flash attention on tcgen05, key bias taken directly
                keep = qkv3 is not None and type(self.proj).__name__ == "TomeLinear"      # proj takes the planes directly
                x = _native.attention_f32(qkv3 if qkv3 is not None else qkv_flat, self.num_heads, self.scale, kb,
                                          out="planes" if keep else "fp32")
            else:
                bias = None if size is None else log_size[:, None, None, :, 0].to(q.dtype).expand(B, 1, N, N)
                drop = self.attn_drop.p if self.training else 0.0
                x = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=drop, scale=self.scale)
                x = x.transpose(1, 2).reshape(B, N, -1)
        x = self.proj_drop(self.proj(x))

        if head_aggregation == 'mean':
            metric = early["metric"]
        elif head_aggregation == 'concat':
            metric = k.transpose(1, 2).reshape(B, N, -1)       # == cat(k.split(1, dim=1), -1).squeeze(1)
        else:
            raise ValueError(f"head_aggregation must be 'mean' or 'concat', got {head_aggregation!r}")
        return x, metric


def _fusable_residual(x, residual):
    return (residual is not None and x.is_cuda and not torch.is_grad_enabled() and residual.shape == x.shape
            and residual.dtype == x.dtype and x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
            and x.shape[-1] % (8 if x.dtype == torch.bfloat16 else 4) == 0
            and x.shape[-1] <= (2048 if x.dtype == torch.bfloat16 else 1024))


def _wavg(merge, x, info, norm, residual=None):
    """merge_wavg through the fused kernel: x', sizes, log sizes and (when the following LayerNorm is
    fusable) LayerNorm(x') stashed in info["normed"] for the block to pick up.  ``residual``: the tokens
    merged are x + residual, added inside the kernel."""
    fn = fusable_norm(norm, x) if norm is not None else None
    res = merge.wavg(x, info["size"], norm=fn, residual=residual)
    info["size"], info["log_size"] = res[1], res[2]
    info["normed"] = res[3] if fn is not None else None
    return res[0]


def videomae_merge(metric, x, _tome_info, norm=None, residual=None):
    """videomae.py:80-100.  ``residual``: the block's pending ``x = x + residual`` (videomae.py:19), folded
    into the merge kernel when that is possible, applied up front otherwise."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                           _tome_info["mode"])
        if residual is not None and not (isinstance(merge, Merge) and _fusable_residual(x, residual)):
            x, residual = x + residual, None
        if _tome_info["trace_source"]:
            _tome_info["source"] = trace_source(merge, x, _tome_info["source"])
        pre_merge = x.size(1)
        if isinstance(merge, Merge):
            x = _wavg(merge, x, _tome_info, norm, residual)
        else:
            x, _tome_info["size"] = merge_wavg(merge, x, _tome_info["size"])
            _tome_info["log_size"] = None
        if _tome_info['verbose']:
            print(f'Merged {pre_merge} to {x.size(1)} tokens')
    elif residual is not None:
        x = x + residual
    return x


def videomae_drop(metric, x, _tome_info, norm=None, residual=None):
    """videomae.py:103-126."""
    if residual is not None:
        x = x + residual
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        drop = bipartite_soft_matching_drop(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                            _tome_info["mode"])
        if isinstance(drop, tuple):              # r clamped to 0 (reference returns a tuple there)
            return x
        if _tome_info["trace_source"]:
            _tome_info["source"] = trace_source(drop, x, _tome_info["source"], drop=True)
        pre_drop = x.size(1)
        x = drop(x)
        _tome_info["size"] = torch.ones((x.size(0), x.size(1), 1), device=x.device)
        _tome_info["log_size"] = torch.zeros((x.size(0), x.size(1), 1), device=x.device)
        if _tome_info['verbose']:
            print(f'Dropped {pre_drop} to {x.size(1)} tokens')
    return x


def videomae_hybrid(metric, x, _tome_info, norm=None, residual=None):
    """videomae.py:129-151."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching_hybrid(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                                  _tome_info["mode"], _tome_info["threshold"])
        if residual is not None and not (isinstance(merge, Merge) and _fusable_residual(x, residual)):
            x, residual = x + residual, None
        if _tome_info["trace_source"]:
            _tome_info["source"] = trace_source(merge, x, _tome_info["source"])
        pre_merge = x.size(1)
        if isinstance(merge, Merge):
            x = _wavg(merge, x, _tome_info, norm, residual)
        else:
            x, _tome_info["size"] = merge_wavg(merge, x, _tome_info["size"])
            _tome_info["log_size"] = None
        if _tome_info['verbose']:
            print(f'Merged {pre_merge} to {x.size(1)} tokens')
    elif residual is not None:
        x = x + residual
    return x


# ---- structural matching of the modules to patch ----------------------------------------------
def _is_block(m):
    return all(hasattr(m, a) for a in ("norm1", "attn", "norm2", "mlp", "gamma_1")) and _is_attention(m.attn)


def _is_attention(m):
    return all(hasattr(m, a) for a in ("qkv", "q_bias", "v_bias", "proj", "num_heads", "scale"))


_CLASS_CACHE = {}


def _swap(module, mixin, tag):
    base = module.__class__
    if getattr(base, "_tome_mixin", None) is mixin:
        return
    if getattr(base, "_tome_mixin", None) is not None:       # re-patching: go back to the original class
        base = base._tome_base
    key = (base, mixin)
    if key not in _CLASS_CACHE:
        _CLASS_CACHE[key] = type(tag + base.__name__, (mixin, base), {"_tome_mixin": mixin, "_tome_base": base})
    module.__class__ = _CLASS_CACHE[key]


def apply_duplicate_patch(model, layer_to_duplicate, quantity):
    """videomae.py:154-157."""
    for i in range(layer_to_duplicate, layer_to_duplicate + quantity - 1):
        model.model.blocks.insert(index=i, module=copy.deepcopy(model.model.blocks[i]))
        _swap(model.model.blocks[i], ToMeDuplicateBlockMixin, "ToMeDuplicate")


def make_tome_class(transformer_class):
    """videomae.py:160-169: per-forward reset of the shared state."""
    class ToMeVisionTransformer(transformer_class):
        def forward(self, *args, **kwdargs) -> torch.Tensor:
            self._tome_info["r"] = parse_r(len(self.model.blocks), self.r)
            self._tome_info["size"] = None
            self._tome_info["log_size"] = None
            self._tome_info["normed"] = None
            self._tome_info["source"] = None
            link_blocks(self.model.blocks, self._tome_info)
            out = super().forward(*args, **kwdargs)
            finish_source(self._tome_info)          # compact source map -> the reference's dense matrix, once
            return out

    return ToMeVisionTransformer


def apply_patch(model_wrapper, trace_source: bool = False, prop_attn: bool = False, mode: str = 'merge',
                head_aggregation: str = 'mean', threshold: float = 0.0, verbose: bool = False):
    """videomae.py:172-214.  After this, set ``model_wrapper.r`` (int | (r, inflect) | list)."""
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
        "head_aggregation": head_aggregation,
        "threshold": threshold,
    }
    if hasattr(model, "dist_token") and model.dist_token is not None:
        model_wrapper._tome_info["distill_token"] = True     # (the reference has a typo here: videomae.py:195-196)

    if mode in ['merge', 'random_merge']:
        reduction_function = videomae_merge
    elif mode in ['drop', 'random_drop']:
        reduction_function = videomae_drop
    elif mode in ['hybrid']:
        reduction_function = videomae_hybrid
    else:
        raise ValueError(f"unknown ToMe mode {mode!r}")

    for module in model.modules():
        if _is_block(module):
            if getattr(module.__class__, "_tome_mixin", None) is not ToMeDuplicateBlockMixin:
                _swap(module, ToMeBlockMixin, "ToMe")
            module._tome_info = model_wrapper._tome_info
            module.reduction_function = reduction_function
        elif _is_attention(module):
            _swap(module, ToMeAttentionMixin, "ToMe")
            module._tome_info = model_wrapper._tome_info
