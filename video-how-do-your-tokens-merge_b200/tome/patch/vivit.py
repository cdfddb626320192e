"""``tome.patch.vivit`` -- drop-in for the reference's tome/patch/vivit.py.

The reference class-swaps HuggingFace's VivitLayer / VivitAttention / VivitSelfAttention and
relies on the pre-4.4x tuple-returning layer API (vivit.py:17-130); that code cannot run against
the installed transformers 5.5.  This patch matches the three module kinds structurally
(``layernorm_before``/``attention``/``intermediate``/``output``; ``attention``+``output``;
``query``/``key``/``value``) so it applies to ``hostmodels.vivit`` and to HF's own modules of either
API generation (a layer whose original ``forward`` takes ``head_mask`` gets tuple returns).
Class token kept out of the merge (``class_token = embeddings.cls_token is not None``,
vivit.py:242): score row 0 = -inf, kept tokens re-sorted ascending."""
import copy
import inspect

import torch
import torch.nn.functional as F

from tome.merge import (Merge, bipartite_soft_matching, bipartite_soft_matching_drop,
                        bipartite_soft_matching_hybrid, finish_source, merge_source, merge_wavg,
                        trace_source)
from tome import attention as prop_attention
from tome.patch.videomae import _fusable_residual, _normed_or, _swap, _wavg, fusable_norm, lazy_head_mean
from tome.utils import parse_r


def _reject_unsupported(head_mask, output_attentions):
    """The reference multiplies the attention probabilities by ``head_mask`` and can return them (vivit.py:113-128);
    the fused attention here never materialises them, so asking for either is an error, not a silent no-op."""
    if head_mask is not None:
        raise NotImplementedError("tome.patch.vivit: head_mask is not supported by the fused attention path")
    if output_attentions:
        raise NotImplementedError("tome.patch.vivit: output_attentions=True is not supported (probabilities are never materialised)")


class ToMeVivitLayerMixin:
    """vivit.py:17-47."""

    def forward(self, hidden_states, head_mask=None, output_attentions=False, **kwargs):
        _reject_unsupported(head_mask, output_attentions)
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        pre = info.pop("normed1", None)          # layernorm_before(hidden_states), made by the previous layer's closing add
        normed = (pre[2] if pre is not None and pre[0] is hidden_states and pre[1] is self.layernorm_before
                  else self.layernorm_before(hidden_states))
        attention_output, metric = self.attention(normed, attn_size, info["head_aggregation"], attn_bias)
        # first residual (attention_output + hidden_states), taken inside the merge kernel when it can be
        hidden_states = self.reduction_function(metric, hidden_states, info, norm=self.layernorm_after,
                                                residual=attention_output)
        h = self.intermediate(_normed_or(self.layernorm_after, hidden_states, info))
        layer_output = self._close(h, hidden_states, info)
        return (layer_output,) if self._tome_tuple_api else layer_output

    def _close(self, h, hidden_states, info):
        """output.dense(h) + hidden_states (vivit.py:44-45); on the CUDA inference path the add and the NEXT layer's
        layernorm_before are one pass (tome_add_layernorm), parked in info["normed1"]."""
        nxt = (info.get("next_norm1") or {}).get(id(self))
        fn = fusable_norm(nxt, hidden_states) if nxt is not None else None
        dense = getattr(self.output, "dense", None)
        if fn is None or dense is None or self.training or not hidden_states.is_contiguous():
            return self.output(h, hidden_states)
        y = dense(h)                                  # (the module's dropout is the identity in eval)
        if y.shape != hidden_states.shape or y.dtype != hidden_states.dtype:
            return y + hidden_states
        from tome import _native
        s, normed = _native.add_layernorm(hidden_states, y, fn)
        info["normed1"] = (s, nxt, normed)
        return s


class ToMeDuplicateVivitLayerMixin:
    """vivit.py:50-66."""

    def forward(self, hidden_states, head_mask=None, output_attentions=False, **kwargs):
        _reject_unsupported(head_mask, output_attentions)
        info = self._tome_info
        attn_size = info["size"] if info["prop_attn"] else None
        attn_bias = info.get("log_size") if info["prop_attn"] else None
        _, metric = self.attention(self.layernorm_before(hidden_states), attn_size, info["head_aggregation"], attn_bias)
        hidden_states = self.reduction_function(metric, hidden_states, info)
        return [hidden_states] if self._tome_tuple_api else hidden_states


class ToMeVivitAttentionMixin:
    """vivit.py:69-83."""

    def forward(self, hidden_states, size=None, head_aggregation='mean', log_size=None, **kwargs):
        # the context may come back as split planes when the output projection consumes them directly (fp32 inference);
        # kept out of the module registry: it is a hint, not a parameter
        object.__setattr__(self.attention, "_tome_ctx_consumer", getattr(self.output, "dense", None))
        ctx, metric = self.attention(hidden_states, size, head_aggregation, log_size)
        return self.output(ctx, hidden_states), metric


class ToMeVivitSelfAttentionMixin:
    """vivit.py:86-130."""

    def forward(self, hidden_states, size=None, head_aggregation='mean', log_size=None, **kwargs):
        B, N, _ = hidden_states.shape
        h, d = self.num_attention_heads, self.attention_head_size
        early = {}

        def on_keys(keys):               # K exists: start match + select beside the attention kernel
            if head_aggregation == 'mean':
                early["metric"] = prop_attention.early_metric(self, keys)

        if size is not None and prop_attention.usable(hidden_states, self):
            # proportional attention (vivit.py:103-104) with the key bias folded into the contraction
            if log_size is None:
                log_size = size.log()
            ctx, k = prop_attention.attention(hidden_states, self, h, d, d ** -0.5, log_size.float(),
                                              self.query.weight, self.key.weight, self.value.weight,
                                              self.query.bias, self.key.bias, self.value.bias, on_keys=on_keys)
        elif prop_attention.usable_f32(hidden_states, self, h, d):
            # fp32 inference: exact-split tensor-core QKV GEMM + flash attention with the key bias (tome_attention_f32)
            if size is not None and log_size is None:
                log_size = size.log()
            ctx, k = prop_attention.attention_f32(hidden_states, self, h, d, d ** -0.5, None if size is None else log_size,
                                                  self.query.weight, self.key.weight, self.value.weight,
                                                  self.query.bias, self.key.bias, self.value.bias, on_keys=on_keys,
                                                  planes_for=getattr(self, "_tome_ctx_consumer", None))
        else:
            q = self.query(hidden_states).view(B, N, h, d).transpose(1, 2)
            k = self.key(hidden_states).view(B, N, h, d).transpose(1, 2)
            on_keys(k)
            v = self.value(hidden_states).view(B, N, h, d).transpose(1, 2)
            bias = None
            if size is not None:                         # proportional attention (vivit.py:103-104)
                if log_size is None:
                    log_size = size.log()
                bias = log_size[:, None, None, :, 0].to(q.dtype).expand(B, 1, N, N)
            # attention-probability dropout in training (vivit.py:110: self.dropout(attention_probs))
            p_drop = self.dropout.p if isinstance(getattr(self, "dropout", None), torch.nn.Dropout) else float(getattr(self, "dropout_prob", 0.0))
            ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=d ** -0.5,
                                                 dropout_p=p_drop if self.training else 0.0)
            ctx = ctx.transpose(1, 2).reshape(B, N, h * d)
        if head_aggregation == 'mean':
            metric = early["metric"]
        elif head_aggregation == 'concat':
            metric = k.transpose(1, 2).reshape(B, N, h * d)
        else:
            raise ValueError(f"head_aggregation must be 'mean' or 'concat', got {head_aggregation!r}")
        return ctx, metric


def vivit_merge(metric, x, _tome_info, norm=None, residual=None):
    """vivit.py:133-153."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                           _tome_info["mode"])
        if residual is not None and not (isinstance(merge, Merge) and _fusable_residual(x, residual)):
            x, residual = residual + x, None
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
        x = residual + x
    return x


def vivit_drop(metric, x, _tome_info, norm=None, residual=None):
    """vivit.py:156-179."""
    if residual is not None:
        x = residual + x
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        drop = bipartite_soft_matching_drop(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                            _tome_info["mode"])
        if isinstance(drop, tuple):
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


def vivit_hybrid(metric, x, _tome_info, norm=None, residual=None):
    """vivit.py:182-204."""
    _tome_info["normed"] = None
    r = _tome_info["r"].pop(0)
    if r > 0:
        merge, _ = bipartite_soft_matching_hybrid(metric, r, _tome_info["class_token"], _tome_info["distill_token"],
                                                  _tome_info["mode"], _tome_info["threshold"])
        if residual is not None and not (isinstance(merge, Merge) and _fusable_residual(x, residual)):
            x, residual = residual + x, None
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
        x = residual + x
    return x


def _is_layer(m):
    return all(hasattr(m, a) for a in ("attention", "intermediate", "output", "layernorm_before", "layernorm_after"))


def _is_attention(m):
    return hasattr(m, "attention") and hasattr(m, "output") and _is_self_attention(getattr(m, "attention", None))


def _is_self_attention(m):
    return m is not None and all(hasattr(m, a) for a in ("query", "key", "value", "num_attention_heads", "attention_head_size"))


def _uses_tuple_api(layer):
    base = getattr(layer.__class__, "_tome_base", layer.__class__)
    try:
        return "head_mask" in inspect.signature(base.forward).parameters
    except (TypeError, ValueError):
        return False


def apply_duplicate_patch(model, layer_to_duplicate, quantity):
    """vivit.py:207-211."""
    for i in range(layer_to_duplicate, layer_to_duplicate + quantity - 1):
        model.vivit.encoder.layer.insert(index=i, module=copy.deepcopy(model.vivit.encoder.layer[i]))
        lyr = model.vivit.encoder.layer[i]
        lyr._tome_tuple_api = _uses_tuple_api(lyr)
        _swap(lyr, ToMeDuplicateVivitLayerMixin, "ToMeDuplicate")
    if hasattr(model.vivit, "config"):
        model.vivit.config.num_hidden_layers = model.vivit.config.num_hidden_layers + quantity


def make_tome_class(transformer_class):
    class ToMeVisionTransformer(transformer_class):
        def forward(self, *args, **kwdargs) -> torch.Tensor:
            self._tome_info["r"] = parse_r(len(self.vivit.encoder.layer), self.r)
            self._tome_info["size"] = None
            self._tome_info["log_size"] = None
            self._tome_info["normed"] = None
            self._tome_info["source"] = None
            layers, chain, seen = list(self.vivit.encoder.layer), {}, set()
            for cur, nxt in zip(layers[:-1], layers[1:]):          # layer -> the LayerNorm the following layer opens with
                chain[id(cur)] = None if id(cur) in seen else getattr(nxt, "layernorm_before", None)
                seen.add(id(cur))
            self._tome_info["next_norm1"], self._tome_info["normed1"] = chain, None
            out = super().forward(*args, **kwdargs)
            finish_source(self._tome_info)          # compact source map -> the reference's dense matrix, once
            return out

    return ToMeVisionTransformer


def apply_patch(model_wrapper, trace_source: bool = False, prop_attn: bool = True, mode: str = 'merge',
                head_aggregation: str = 'mean', threshold: float = 0.0, verbose: bool = False):
    """vivit.py:226-270."""
    model = model_wrapper.vivit
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
        "class_token": model.embeddings.cls_token is not None,
        "distill_token": False,
        "mode": mode,
        "head_aggregation": head_aggregation,
        "threshold": threshold,
    }
    if hasattr(model, "dist_token") and model.dist_token is not None:
        model_wrapper._tome_info["distill_token"] = True

    if mode in ['merge', 'random_merge']:
        reduction_function = vivit_merge
    elif mode in ['drop', 'random_drop']:
        reduction_function = vivit_drop
    elif mode in ['hybrid']:
        reduction_function = vivit_hybrid
    else:
        raise ValueError(f"unknown ToMe mode {mode!r}")

    for module in model.modules():
        if _is_layer(module):
            module._tome_tuple_api = _uses_tuple_api(module)
            if getattr(module.__class__, "_tome_mixin", None) is not ToMeDuplicateVivitLayerMixin:
                _swap(module, ToMeVivitLayerMixin, "ToMe")
            module._tome_info = model_wrapper._tome_info
            module.reduction_function = reduction_function
        elif _is_attention(module):
            _swap(module, ToMeVivitAttentionMixin, "ToMe")
        elif _is_self_attention(module):
            _swap(module, ToMeVivitSelfAttentionMixin, "ToMe")
            module._tome_info = model_wrapper._tome_info
