"""B200-native drop-in for the reference's ``tome/merge.py``.

Same public names, argument meaning and return conventions as
/root/reference/tome/merge.py; the work runs in the sm_100a kernels behind the C ABI
(include/tome_b200.h) instead of ~40 ATen launches per block:

  bipartite_soft_matching        (merge.py:17-102)   -> tome_match + tome_select
  bipartite_soft_matching_drop   (merge.py:215-271)  -> same matching, tome_merge(DROP)
  bipartite_soft_matching_hybrid (merge.py:274-352)  -> same matching, tome_merge(threshold)
  merge_wavg                     (merge.py:355-369)  -> ONE fused tome_merge(WAVG) pass
  merge_source                   (merge.py:372-384)  -> tome_merge_source
  kth_/random_bipartite_soft_matching (merge.py:105-212) -- upstream-ToMe leftovers with no
      caller in the reference; kept importable, built from device torch ops.

Differences that are deliberate and documented in DESIGN.md: ties are broken by the stable
rule (lowest index) where the reference's argsort is unstable; scores use the canonical
fp64-accumulated definition; token sizes are always fp32.  No CPU fallback: tensors must
be on an sm_100 CUDA device.
"""
import math
import os
from typing import Callable, Optional, Tuple

import torch

from . import _native


def do_nothing(x, mode=None):
    return x


def _effective_r(metric: torch.Tensor, r: int, class_token: bool, distill_token: bool) -> int:
    protected = int(bool(class_token)) + int(bool(distill_token))
    t = metric.shape[1]
    return min(r, (t - protected) // 2)          # merge.py:43-44


_SIDE_STREAMS = {}          # device index -> the stream early matching runs on
_PHILOX = {}                # device index -> _native.PhiloxStream, once philox_seed() has been called


def philox_seed(seed: Optional[int], clip_offset: int = 0) -> None:
    """Draw the scores of the random_* modes (merge.py:54-57, 235-238) from the library's counter-based Philox
    stream instead of ``torch.rand``: an edge's score then depends only on (seed, call number, clip_offset + clip,
    row, column) -- independent of the batch composition and of how clips are sharded over GPUs (give each rank its
    first clip's global index as ``clip_offset``) -- and the (bm, na, nb) score tensor is never materialised.
    ``seed=None`` goes back to the reference's own ``torch.rand`` call (the default)."""
    _PHILOX.clear()
    if seed is not None:
        _PHILOX["seed"] = (int(seed), int(clip_offset))


def prefetch_matching(metric, r: int, class_token: bool = False, distill_token: bool = False) -> None:
    """Start match + select for ``metric`` NOW, on a side stream, and park the plan on the metric.

    The matching needs only the attention keys, which exist right after the QKV projection, while the
    plan is consumed only after attention, projection and the residual add (videomae.py:13-30).  The
    patched attentions call this as soon as K is there, so the latency-bound match/select chain runs
    beside the attention kernel instead of after it; ``bipartite_soft_matching*`` picks the plan up
    (same r / tokens) and makes the consuming stream wait for it.  Under CUDA-graph capture the side
    stream becomes a parallel branch of the graph."""
    if not isinstance(metric, _native.HeadMeanMetric) or not metric.is_cuda or torch.is_grad_enabled():
        return
    if os.environ.get("TOME_PREFETCH", "0") != "1":
        # Measured on a B200 (VideoMAE-B bench, CUDA graph): 2.757 ms/step with the side branch vs 2.744
        # without (prop_attn: 2.78 vs 2.84; high stream priority: worse).  The attention kernel owns
        # every SM's shared memory, so the match CTAs only run as attention CTAs retire and the branch
        # buys nothing.  Kept behind this knob; off by default.
        return
    r = _effective_r(metric, r, class_token, distill_token)
    if r <= 0:
        return
    dev = metric.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(main)                       # K is complete in main-stream order
    with torch.cuda.stream(side):
        plan = _make_plan(metric, r, class_token, distill_token, False)
    metric.prefetched = (r, bool(class_token), bool(distill_token), plan, side)


def _make_plan(metric, r, class_token, distill_token, random_scores: bool) -> "_native.DevicePlan":
    _native._require_cuda(metric if not isinstance(metric, _native.HeadMeanMetric) else metric.keys, "metric")
    pre = getattr(metric, "prefetched", None)
    if pre is not None and not random_scores and pre[:3] == (r, bool(class_token), bool(distill_token)):
        plan, side = pre[3], pre[4]
        metric.prefetched = None
        main = torch.cuda.current_stream(plan.device)
        main.wait_stream(side)
        for t in (plan.node_max, plan.node_idx, plan._ints):     # allocated on the side stream, read on this one
            t.record_stream(main)
        return plan
    with torch.no_grad():
        if random_scores:                          # merge.py:54-57: same torch.rand call, same generator
            length = metric.size(1)
            len_a, len_b = (length + 1) // 2, length // 2
            if "seed" in _PHILOX:
                dev = metric.device
                stream = _PHILOX.get(dev.index)
                if stream is None:
                    stream = _PHILOX[dev.index] = _native.PhiloxStream(_PHILOX["seed"][0], dev, _PHILOX["seed"][1])
                node_max, node_idx = _native.random_rowmax(stream, metric.size(0), len_a, len_b, class_token, distill_token)
            else:
                scores = torch.rand(size=(metric.size(0), len_a, len_b), device=metric.device)
                node_max, node_idx = _native.rowmax(scores, class_token, distill_token)
            return _native.select(node_max, node_idx, metric.shape[1], r, class_token, distill_token)
        return _native.plan_build(metric, r, class_token, distill_token)        # kernels 1 + 2, one ABI call


class _PlanCallable:
    """Base of the merge/unmerge/drop callables: exposes what the reference closures
    capture (``unm_idx``, ``src_idx``, ``dst_idx`` as int64 (bm, k, 1); ``r``)."""

    def __init__(self, plan: "_native.DevicePlan", threshold: Optional[float] = None):
        self.plan = plan
        self.threshold = threshold

    r = property(lambda self: self.plan.r)
    src_idx = property(lambda self: self.plan.src_idx.long()[..., None])
    unm_idx = property(lambda self: self.plan.unm_idx.long()[..., None])
    dst_idx = property(lambda self: self.plan.dst_idx.long()[..., None])
    node_max = property(lambda self: self.plan.node_max)


class _MergeGrad(torch.autograd.Function):
    """Backward of the merge kernel: the reference trains through its gather /
    scatter_reduce closures (tools/train_net.py:727-741).  Every input token t lands in exactly one output
    slot, with weight 1 (sum / drop), 1 / count (mean) or size_t / size'_slot (merge_wavg), so the gradient is
    the kernel's own inverse map -- ``tome_unmerge`` of grad_out -- times that weight; tokens that do not
    reach the output (dropped sources; destinations a hybrid threshold zeroed, merge.py:326) get zero."""

    @staticmethod
    def forward(ctx, x, plan, mode, size, threshold):
        if plan.distill_token:
            raise NotImplementedError("tome_b200: merge backward with a distillation token is not supported")
        if mode in ("max", "amax"):
            raise NotImplementedError("tome_b200: merge backward is not defined for the amax mode")
        want = mode == "wavg"
        res = _native.merge(plan, x.detach(), mode, size=size, hybrid_threshold=threshold, want_size=want)
        out, size_out, logsize_out = (res if want else (res, None, None))
        ctx.plan, ctx.mode, ctx.threshold = plan, mode, threshold
        ctx.save_for_backward(size, size_out)
        if want:
            ctx.mark_non_differentiable(size_out, logsize_out)
            return out, size_out, logsize_out
        return out

    @staticmethod
    def backward(ctx, grad_out, *unused):
        plan, mode, thr = ctx.plan, ctx.mode, ctx.threshold
        size, size_out = ctx.saved_tensors
        bm, n, r = plan.bm, plan.n, plan.r
        na, nb = (n + 1) // 2, n // 2
        g = _native.unmerge(plan, grad_out.contiguous())                     # (bm, n, c): grad of the slot each token fed
        dev = g.device
        if mode == "wavg":
            s_in = torch.ones(bm, n, 1, device=dev) if size is None else size.reshape(bm, n, 1).float()
            s_slot = _native.unmerge(plan, size_out.reshape(bm, n - r, 1).contiguous())
            g = g * (s_in / s_slot).to(g.dtype)
        elif mode == "mean":
            cnt = torch.ones(bm, n - r, 1, device=dev)
            cnt[:, na - r:, 0] += plan.b_head[:, :, 0].float()
            g = g / _native.unmerge(plan, cnt).to(g.dtype)
        if mode == "drop":                                                   # sources are discarded (merge.py:260-269)
            g[torch.arange(bm, device=dev)[:, None], 2 * plan.src_idx.long()] = 0
        elif thr is not None and thr == thr:                                 # hybrid: destinations hit by an under-threshold edge
            low = (plan.node_max.gather(1, plan.src_idx.long()) < thr)
            hit = torch.zeros(bm, nb, device=dev).scatter_add_(1, plan.dst_idx.long(), low.float()) > 0
            g[:, 1::2][hit] = 0
        return g, None, None, None, None


def _needs_grad(x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and x.requires_grad


class Merge(_PlanCallable):
    def __call__(self, x: torch.Tensor, mode="mean") -> torch.Tensor:      # merge.py:75-85 / 316-334
        if _needs_grad(x):
            return _MergeGrad.apply(x, self.plan, mode, None, self.threshold)
        return _native.merge(self.plan, x, mode, hybrid_threshold=self.threshold)

    def wavg(self, x, size=None, norm=None, residual=None):
        """Fused merge_wavg: (x', size' (bm, n', 1) fp32, log size' (bm, n', 1) fp32); with
        ``norm=(weight, bias, eps)`` also LayerNorm(x') from the same pass as a 4th result; with
        ``residual`` the tokens merged are ``x + residual`` (the block's residual add, same pass)."""
        if _needs_grad(x):                         # training: kernel forward, unmerge-based backward; no fusions
            if residual is not None:
                x = x + residual
            out, s, ls = _MergeGrad.apply(x, self.plan, "wavg", size, self.threshold)
            res = (out, s, ls) + ((torch.nn.functional.layer_norm(out, (out.shape[-1],), norm[0], norm[1], norm[2]),)
                                  if norm is not None else ())
            return (res[0], s[..., None], ls[..., None]) + tuple(res[3:])
        res = _native.merge(self.plan, x, "wavg", size=size, hybrid_threshold=self.threshold, want_size=True, norm=norm,
                            residual=residual)
        out, s, ls = res[:3]
        return (out, s[..., None], ls[..., None]) + tuple(res[3:])

    def source(self, source=None):
        if isinstance(source, _native.SourceMap):      # compact form stays compact
            return _native.source_compose(self.plan, source, False, self.threshold)
        return _native.merge_source(self.plan, source, self.threshold)

    def source_map(self, source=None):
        """merge_source on the compact form: ``source`` is a SourceMap or None (identity)."""
        return _native.source_compose(self.plan, source, False, self.threshold)

    def wavg_frames(self, x, frames, size=None, norm=None, residual=None, cls=None):
        """merge_wavg on a (B, 1 + P*T, C) class-token + '(p t)' tensor whose matching batch is (b t):
        the TimeSformer / Motionformer case, rearranges folded into addressing.
        Returns (x' (B, 1 + P'*T, C), size' (B*T, P', 1), log size' (B*T, P', 1)).
        ``residual`` ((B*T, 1 + P, C), the spatial attention's output in its own layout) is added to the patch
        tokens inside the kernel and ``cls`` (B, C) becomes the class row (inference path only)."""
        if _needs_grad(x):                         # training: the reference's rearranges, differentiable merge
            B, L, C = x.shape
            T = int(frames)
            P = (L - 1) // T
            xs = x[:, 1:].reshape(B, P, T, C).transpose(1, 2).reshape(B * T, P, C)
            out, s, ls = self.wavg(xs, size)[:3]
            Pn = out.size(1)
            y = torch.cat((x[:, :1], out.reshape(B, T, Pn, C).transpose(1, 2).reshape(B, Pn * T, C)), 1)
            extra = (torch.nn.functional.layer_norm(y, (C,), norm[0], norm[1], norm[2]),) if norm is not None else ()
            return (y, s, ls) + extra
        res = _native.merge_frames(self.plan, x, frames, "wavg", size=size, hybrid_threshold=self.threshold, norm=norm,
                                   residual=residual, cls=cls)
        out, s, ls = res[:3]
        return (out, s[..., None], ls[..., None]) + tuple(res[3:])


class _UnmergeGrad(torch.autograd.Function):
    """unmerge copies merged slot s to every token that fed it (merge.py:87-100: gather + scatter), so its adjoint
    sums the gradients of those tokens back into the slot: the merge kernel in 'sum' mode."""

    @staticmethod
    def forward(ctx, x, plan):
        ctx.plan = plan
        return _native.unmerge(plan, x.detach())

    @staticmethod
    def backward(ctx, grad_out):
        return _native.merge(ctx.plan, grad_out.contiguous(), "sum"), None


class Unmerge(_PlanCallable):
    def __call__(self, x: torch.Tensor) -> torch.Tensor:                    # merge.py:87-100
        if _needs_grad(x):
            if self.plan.distill_token:
                raise NotImplementedError("tome_b200: unmerge backward with a distillation token is not supported")
            return _UnmergeGrad.apply(x, self.plan)
        return _native.unmerge(self.plan, x)


class Drop(_PlanCallable):
    und_idx = property(lambda self: self.plan.unm_idx.long()[..., None])

    def __call__(self, x: torch.Tensor) -> torch.Tensor:                    # merge.py:260-269
        if _needs_grad(x):
            return _MergeGrad.apply(x, self.plan, "drop", None, None)
        return _native.merge(self.plan, x, "drop")

    def source_map(self, source=None):
        """The drop closure applied to the compact source form (dropped tokens get group -1)."""
        return _native.source_compose(self.plan, source, True, None)

    def frames(self, x, frames, residual=None, cls=None):
        """drop on the (B, 1 + P*T, C) layout (see Merge.wavg_frames)."""
        if _needs_grad(x):
            B, L, C = x.shape
            T = int(frames)
            P = (L - 1) // T
            out = self(x[:, 1:].reshape(B, P, T, C).transpose(1, 2).reshape(B * T, P, C))
            Pn = out.size(1)
            return torch.cat((x[:, :1], out.reshape(B, T, Pn, C).transpose(1, 2).reshape(B, Pn * T, C)), 1)
        return _native.merge_frames(self.plan, x, frames, "drop", residual=residual, cls=cls)[0]


def bipartite_soft_matching(
    metric: torch.Tensor, r: int, class_token: bool = False, distill_token: bool = False, mode: str = 'merge'
) -> Tuple[Callable, Callable]:
    """Balanced (50/50) bipartite soft matching; see the reference docstring (merge.py:24-35).

    metric: [batch, tokens, channels]; r: tokens to remove (clamped to 50% of unprotected)."""
    r = _effective_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    if mode not in ('merge', 'random_merge'):
        raise ValueError(f"bipartite_soft_matching: mode must be 'merge' or 'random_merge', got {mode!r}")
    plan = _make_plan(metric, r, class_token, distill_token, mode == 'random_merge')
    return Merge(plan), Unmerge(plan)


def bipartite_soft_matching_drop(
    metric: torch.Tensor, r: int, class_token: bool = False, distill_token: bool = False, mode: str = 'drop'
):
    r = _effective_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing      # sic: the reference returns a tuple here (merge.py:232-233)
    if mode not in ('drop', 'random_drop'):
        raise ValueError(f"bipartite_soft_matching_drop: mode must be 'drop' or 'random_drop', got {mode!r}")
    plan = _make_plan(metric, r, class_token, distill_token, mode == 'random_drop')
    return Drop(plan)


def bipartite_soft_matching_hybrid(
    metric: torch.Tensor, r: int, class_token: bool = False, distill_token: bool = False, mode: str = 'merge',
    threshold: float = 0.0
) -> Tuple[Callable, Callable]:
    r = _effective_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    if mode not in ('merge', 'hybrid', 'random_merge'):
        raise ValueError(f"bipartite_soft_matching_hybrid: bad mode {mode!r}")
    plan = _make_plan(metric, r, class_token, distill_token, mode == 'random_merge')
    return Merge(plan, threshold=float(threshold)), Unmerge(plan)


def merge_wavg(merge: Callable, x: torch.Tensor, size: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Size-weighted average merge; returns (merged x, new sizes) like merge.py:355-369."""
    if isinstance(merge, Merge):
        out, size_out, _ = merge.wavg(x, size)
        return out, size_out
    if size is None:
        size = torch.ones_like(x[..., 0, None])
    x = merge(x * size, mode="sum")
    size = merge(size, mode="sum")
    x = x / size
    return x, size


def merge_source(merge: Callable, x: torch.Tensor, source: torch.Tensor = None) -> torch.Tensor:
    """Source adjacency tracking (merge.py:372-384)."""
    if isinstance(merge, Merge):
        return merge.source(source)
    if isinstance(source, _native.SourceMap):
        return source                      # do_nothing: no tokens moved
    if source is None:
        n, t, _ = x.shape
        source = torch.eye(t, device=x.device)[None, ...].expand(n, t, t)
    return merge(source, mode="max")


# ---- compact source tracing used by the patches ---------------------------------------
def trace_source(op, x: torch.Tensor, source, drop: bool = False):
    """One block's ``merge_source`` (merge.py:372-384) / ``drop(source)`` (videomae.py:118-122) on the compact
    form: returns a ``_native.SourceMap``.  The patched model expands it to the reference's dense fp32 matrix
    once, at the end of the forward (``finish_source``), instead of re-reducing a (bm, n, n0) matrix per block.
    A dense ``source`` tensor (a caller that started in the reference's form) stays dense."""
    if isinstance(op, (Merge, Drop)):
        if torch.is_tensor(source):
            return op(source.contiguous()) if isinstance(op, Drop) else op.source(source)
        return op.source_map(source)
    if op is do_nothing or isinstance(op, tuple):        # r clamped to 0: nothing moves
        if source is None:
            n, t = x.shape[0], x.shape[1]
            source = _native.SourceMap(torch.arange(t, device=x.device, dtype=torch.int32).expand(n, t).contiguous(), t)
        return source
    # a foreign callable (the tests route the patches to a CPU port of the reference): the reference's dense form
    if drop:
        if source is None:
            n, t, _ = x.shape
            source = torch.eye(t, device=x.device)[None, ...].expand(n, t, t)
        return op(source.contiguous())
    return merge_source(op, x, source)


def finish_source(info: dict) -> None:
    """``_tome_info['source']`` as the reference leaves it -- dense (bm, tokens, n0) fp32 -- with the compact map
    kept beside it as ``_tome_info['source_map']``."""
    src = info.get("source")
    if isinstance(src, _native.SourceMap):
        info["source_map"] = src
        info["source"] = src.dense()


# ---- upstream-ToMe variants with no caller in the reference (merge.py:105-212) -------------
# Both match two ARBITRARY token sets (sources -> destinations) and merge EVERY source.  Here a set is an index
# list into the token axis (``_native.TokenSets``) and the work runs in three kernels of the C ABI:
# ``tome_match_sets`` (canonical scores, best destination per source, lowest index on ties), ``tome_group_reduce``
# (the out-of-place scatter_reduce with include_self, reference CPU order) and ``tome_gather_rows`` (unmerge).
class _SetMerge:
    def __init__(self, sets, dst_row):
        self.sets, self.dst_row = sets, dst_row
        self.dst_idx = dst_row.long()[..., None]          # what the reference closure captures

    def __call__(self, x: torch.Tensor, mode="mean") -> torch.Tensor:
        return _native.group_reduce(x, self.sets, self.dst_row, mode)


class _SetUnmerge:
    def __init__(self, row_of_token):
        self.row_of_token = row_of_token                  # (bm, tokens out): merged row each original token reads

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return _native.gather_rows(x, self.row_of_token)


def kth_bipartite_soft_matching(metric: torch.Tensor, k: int) -> Tuple[Callable, Callable]:
    """Every k-th token is a destination, the k - 1 before it are sources (merge.py:105-158); tokens / k remain.
    A ragged tail (tokens % k) is ignored, as in the reference."""
    if k <= 1:
        return do_nothing, do_nothing
    bm, n, _ = metric.shape
    groups = n // k
    dev = metric.device
    token = torch.arange(groups * k, device=dev, dtype=torch.int32).view(groups, k)
    sets = _native.TokenSets(token[:, :k - 1].reshape(-1), token[:, k - 1])
    with torch.no_grad():
        _, dst_row = _native.match_sets(metric, sets)
    own = torch.arange(groups, device=dev, dtype=torch.int32).expand(bm, groups)
    row_of_token = torch.cat((dst_row.view(bm, groups, k - 1), own[..., None]), dim=2).reshape(bm, groups * k)
    return _SetMerge(sets, dst_row), _SetUnmerge(row_of_token)


def random_bipartite_soft_matching(metric: torch.Tensor, r: int) -> Tuple[Callable, Callable]:
    """r randomly chosen sources, everything else a destination (merge.py:161-212); tokens - r remain.  The
    draw is the reference's own call (``torch.rand(B, N, 1)`` on metric's device), so the same generator state
    gives the same partition."""
    if r <= 0:
        return do_nothing, do_nothing
    bm, n, _ = metric.shape
    dev = metric.device
    with torch.no_grad():
        order = torch.rand(bm, n, 1, device=dev).argsort(dim=1)[..., 0]
        sets = _native.TokenSets(order[:, :r], order[:, r:])
        _, dst_row = _native.match_sets(metric, sets)
        rows = torch.cat((dst_row, torch.arange(n - r, device=dev, dtype=torch.int32).expand(bm, n - r)), dim=1)
        row_of_token = torch.empty(bm, n, device=dev, dtype=torch.int32).scatter_(1, order, rows)
    return _SetMerge(sets, dst_row), _SetUnmerge(row_of_token)
