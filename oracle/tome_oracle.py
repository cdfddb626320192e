"""CPU oracle for the token-merging hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this module.  The shipped path
(``video-how-do-your-tokens-merge_b200/tome``) never imports it and fails loudly when the
CUDA extension is missing.

It is a numpy restatement (written from SURVEY.md Appendix A, not copied) of the
reference algorithm in ``/root/reference/tome/merge.py``:

  * ``match``        <- merge.py:49-64   (normalise, A.B^T, cls/distill mask, row max/argmax)
  * ``select``       <- merge.py:65-73   (rank "a" tokens, split src/unm, gather dst, cls sort)
  * ``merge``        <- merge.py:75-85   (gather unm/src, scatter_reduce into dst, concat)
  * ``merge_hybrid`` <- merge.py:316-334 (threshold 'prod' mask before the reduce)
  * ``drop``         <- merge.py:260-269
  * ``unmerge``      <- merge.py:87-100
  * ``merge_wavg``   <- merge.py:355-369
  * ``merge_source`` <- merge.py:372-384
  * ``parse_r``      <- tome/utils.py:83-108

Pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py``
imports ``/root/reference/tome/merge.py`` in the build container, runs it on seeded
inputs and commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against those fixtures (and, when ``/root/reference`` is present, against
the live reference).

Canonical arithmetic (the "reference's stable rule" of BASELINE.json made explicit; the
reference's own float rounding differs between MKL, cuBLAS and any other GEMM, and its
``argsort`` is not stable, so it is not self-consistent across devices):

  norm[n]   = fp32( sqrt( sum_k fp64(M[n,k])^2 ) )            (no eps: zero row -> NaN)
  mhat[n,k] = fp32( M[n,k] / norm[n] )                        (IEEE fp32 division)
  S[i,j]    = fp32( sum_k fp64(mhat[2i,k]) * fp64(mhat[2j+1,k]) ) + 0.0f
  node_idx  = lowest j attaining the row max; edge order = stable descending
              (lowest i first among equal node_max); NaN sorts above +inf (torch rule).

Decisions made with this arithmetic equal the reference's wherever the reference's own
margins exceed a few fp32 ulp; tests use a gap-aware comparator for the rest.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np

F32 = np.float32
NEG_INF = F32(-np.inf)


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def orderable_u32(v: np.ndarray) -> np.ndarray:
    """Map fp32 -> uint32 so that unsigned order == float order (NaN above +inf).

    Same bit trick the CUDA kernels use (csrc/common.cuh: ``orderable_key``)."""
    v = np.asarray(v, dtype=F32) + F32(0.0)  # -0.0 -> +0.0
    u = v.view(np.uint32).copy()
    neg = (u >> np.uint32(31)).astype(bool)
    u = np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    u = np.where(np.isnan(v), np.uint32(0xFFFFFFFF), u).astype(np.uint32)
    return u


def effective_r(n_tokens: int, r: int, class_token: bool, distill_token: bool) -> int:
    """merge.py:36-44 -- at most half of the unprotected tokens can go."""
    protected = int(bool(class_token)) + int(bool(distill_token))
    return max(min(int(r), (n_tokens - protected) // 2), 0)


def parse_r(num_layers: int, r) -> List[int]:
    """tome/utils.py:83-108."""
    inflect = 0
    if isinstance(r, list):
        if len(r) < num_layers:
            r = r + [0] * (num_layers - len(r))
        return list(r)
    if isinstance(r, tuple):
        r, inflect = r
    lo = int(r * (1.0 - inflect))
    hi = 2 * r - lo
    step = (hi - lo) / (num_layers - 1)
    return [int(lo + step * i) for i in range(num_layers)]


# ----------------------------------------------------------------------------------------
# kernel 1: match
# ----------------------------------------------------------------------------------------
def normalise(metric: np.ndarray) -> np.ndarray:
    m = np.asarray(metric, dtype=F32)
    with np.errstate(all="ignore"):
        ss = np.sum(m.astype(np.float64) ** 2, axis=-1)
        norm = np.sqrt(ss).astype(F32)
        return (m / norm[..., None]).astype(F32)


def scores(metric: np.ndarray, class_token=False, distill_token=False) -> np.ndarray:
    """Full (Bm, Na, Nb) canonical score matrix -- only the oracle materialises it."""
    mh = normalise(metric).astype(np.float64)
    a, b = mh[:, 0::2, :], mh[:, 1::2, :]
    with np.errstate(all="ignore"):
        s = np.matmul(a, b.transpose(0, 2, 1)).astype(F32) + F32(0.0)
    if class_token:
        s[:, 0, :] = NEG_INF
    if distill_token:
        s[:, :, 0] = NEG_INF
    return s


def rowmax(s: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """merge.py:64 -- max + first index attaining it (NaN wins, first NaN)."""
    key = orderable_u32(s)
    idx = np.argmax(key, axis=-1).astype(np.int32)  # first occurrence of the max key
    val = np.take_along_axis(s, idx[..., None].astype(np.int64), axis=-1)[..., 0]
    return val.astype(F32), idx


def match(metric: np.ndarray, class_token=False, distill_token=False):
    return rowmax(scores(metric, class_token, distill_token))


# ----------------------------------------------------------------------------------------
# kernel 2: select
# ----------------------------------------------------------------------------------------
@dataclass
class Plan:
    """What the reference's closures capture (merge.py:75: unm_idx, src_idx, dst_idx, r)."""
    n_tokens: int
    r: int
    class_token: bool
    distill_token: bool
    src_idx: np.ndarray   # (Bm, r)      index into the A (even) tokens
    unm_idx: np.ndarray   # (Bm, Na - r)
    dst_idx: np.ndarray   # (Bm, r)      index into the B (odd) tokens
    node_max: Optional[np.ndarray] = None   # (Bm, Na)
    node_idx: Optional[np.ndarray] = None   # (Bm, Na)

    @property
    def na(self):
        return (self.n_tokens + 1) // 2

    @property
    def nb(self):
        return self.n_tokens // 2


def select(node_max, node_idx, n_tokens, r, class_token=False, distill_token=False) -> Optional[Plan]:
    r = effective_r(n_tokens, r, class_token, distill_token)
    if r <= 0:
        return None
    key = orderable_u32(node_max).astype(np.int64)
    edge = np.argsort(-key, axis=-1, kind="stable").astype(np.int32)   # stable descending
    src = edge[:, :r]
    unm = edge[:, r:]
    dst = np.take_along_axis(node_idx, src.astype(np.int64), axis=-1).astype(np.int32)
    if class_token:
        unm = np.sort(unm, axis=-1)
    return Plan(n_tokens, r, bool(class_token), bool(distill_token),
                np.ascontiguousarray(src), np.ascontiguousarray(unm), np.ascontiguousarray(dst),
                np.asarray(node_max, F32), np.asarray(node_idx, np.int32))


def bipartite_soft_matching(metric, r, class_token=False, distill_token=False,
                            given_scores: Optional[np.ndarray] = None) -> Optional[Plan]:
    """merge.py:17-73.  ``given_scores`` stands in for torch.rand in the random_* modes."""
    n = metric.shape[1]
    if effective_r(n, r, class_token, distill_token) <= 0:
        return None
    if given_scores is None:
        s = scores(metric, class_token, distill_token)
    else:
        s = np.array(given_scores, dtype=F32, copy=True)
        if class_token:
            s[:, 0, :] = NEG_INF
        if distill_token:
            s[:, :, 0] = NEG_INF
    nm, ni = rowmax(s)
    return select(nm, ni, n, r, class_token, distill_token)


# ----------------------------------------------------------------------------------------
# kernel 3: merge family
# ----------------------------------------------------------------------------------------
def _assemble(plan: Plan, unm: np.ndarray, dst: np.ndarray) -> np.ndarray:
    if plan.distill_token:
        return np.concatenate([unm[:, :1], dst[:, :1], unm[:, 1:], dst[:, 1:]], axis=1)
    return np.concatenate([unm, dst], axis=1)


def merge(plan: Optional[Plan], x: np.ndarray, mode: str = "mean",
          hybrid_threshold: Optional[float] = None) -> np.ndarray:
    """merge.py:75-85 (and 316-334 when ``hybrid_threshold`` is given).

    Reduction order is the reference CPU order: the dst token itself, then its sources in
    ascending k (sorted-edge order); every step rounds to x.dtype (fp32)."""
    if plan is None:
        return x
    x = np.asarray(x)
    a, b = x[:, 0::2, :], x[:, 1::2, :]
    bm = x.shape[0]
    unm = np.take_along_axis(a, plan.unm_idx[..., None].astype(np.int64), axis=1)
    dst = np.array(b, copy=True)
    if hybrid_threshold is not None:
        # merge.py:326 -- multiply each hit dst row by (node_max_sorted[k] >= thr)
        keep = np.take_along_axis(plan.node_max, plan.src_idx.astype(np.int64), axis=1) >= F32(hybrid_threshold)
        for bi in range(bm):
            for k in range(plan.r):
                dst[bi, plan.dst_idx[bi, k]] = dst[bi, plan.dst_idx[bi, k]] * x.dtype.type(keep[bi, k])
    cnt = np.ones((bm, dst.shape[1]), dtype=np.int64)
    with np.errstate(all="ignore"):
        for bi in range(bm):
            for k in range(plan.r):
                j = plan.dst_idx[bi, k]
                s = a[bi, plan.src_idx[bi, k]]
                if mode in ("sum", "mean"):
                    dst[bi, j] = dst[bi, j] + s
                elif mode in ("max", "amax"):
                    dst[bi, j] = np.maximum(dst[bi, j], s)
                else:
                    raise ValueError(mode)
                cnt[bi, j] += 1
        if mode == "mean":
            dst = (dst / cnt[..., None].astype(x.dtype)).astype(x.dtype)   # include_self=True
    return _assemble(plan, unm, dst)


def drop(plan: Optional[Plan], x: np.ndarray) -> np.ndarray:
    """merge.py:260-269 -- src tokens are discarded."""
    if plan is None:
        return x
    a, b = x[:, 0::2, :], x[:, 1::2, :]
    und = np.take_along_axis(a, plan.unm_idx[..., None].astype(np.int64), axis=1)
    return _assemble(plan, und, b)


def unmerge(plan: Optional[Plan], x: np.ndarray) -> np.ndarray:
    """merge.py:87-100."""
    if plan is None:
        return x
    bm, _, c = x.shape
    ul = plan.unm_idx.shape[1]
    unm, dst = x[:, :ul, :], x[:, ul:, :]
    out = np.zeros((bm, plan.n_tokens, c), dtype=x.dtype)
    out[:, 1::2, :] = dst
    for bi in range(bm):
        out[bi, 2 * plan.unm_idx[bi]] = unm[bi]
        out[bi, 2 * plan.src_idx[bi]] = dst[bi, plan.dst_idx[bi]]
    return out


def merge_wavg(plan, x, size=None, hybrid_threshold=None):
    """merge.py:355-369: x*size, sum-merge both, divide."""
    x = np.asarray(x)
    if size is None:
        size = np.ones_like(x[..., :1])
    with np.errstate(all="ignore"):
        xs = merge(plan, (x * size).astype(x.dtype), "sum", hybrid_threshold)
        sz = merge(plan, size, "sum", hybrid_threshold)
        return (xs / sz).astype(x.dtype), sz


def merge_source(plan, x, source=None, hybrid_threshold=None):
    """merge.py:372-384 (fp32 identity when source is None)."""
    if source is None:
        n, t, _ = x.shape
        source = np.broadcast_to(np.eye(t, dtype=F32)[None], (n, t, t)).copy()
    return merge(plan, source, "max", hybrid_threshold)


# ----------------------------------------------------------------------------------------
# comparison helper used by the parity tests (gap-aware; SURVEY.md section 7, hard part 1)
# ----------------------------------------------------------------------------------------

# ---- random-mode scores: counter-based Philox4x32-10 ---------------------------------
# The reference draws torch.rand((bm, na, nb)) (tome/merge.py:54-57, 235-238); the CUDA path can instead draw
# from a Philox stream (include/tome_b200.h: tome_random_rowmax).  Algorithm: Salmon, Moraes, Dror, Shaw,
# "Parallel random numbers: as easy as 1, 2, 3" (SC'11), philox4x32 with 10 rounds; pinned by the Random123
# known-answer vectors in tests/test_oracle_golden.py.
_PH_M0, _PH_M1, _PH_W0, _PH_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """counter: (..., 4) uint32-valued array, key: (k0, k1).  Returns (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c[..., 0]
        p1 = np.uint64(_PH_M1) * c[..., 2]
        c = np.stack(((p1 >> np.uint64(32)) ^ c[..., 1] ^ np.uint64(k0), p1 & mask,
                      (p0 >> np.uint64(32)) ^ c[..., 3] ^ np.uint64(k1), p0 & mask), axis=-1)
        k0, k1 = (k0 + _PH_W0) & 0xFFFFFFFF, (k1 + _PH_W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox_scores(seed: int, call: int, clip0: int, bm: int, na: int, nb: int) -> np.ndarray:
    """(bm, na, nb) fp32 scores of the stream: key = seed, counter = (column // 4, A row, clip0 + b, call),
    score = (word >> 8) * 2^-24."""
    quads = (nb + 3) // 4
    ctr = np.zeros((bm, na, quads, 4), dtype=np.uint64)
    ctr[..., 0] = np.arange(quads, dtype=np.uint64)[None, None, :]
    ctr[..., 1] = np.arange(na, dtype=np.uint64)[None, :, None]
    ctr[..., 2] = (np.uint64(clip0) + np.arange(bm, dtype=np.uint64))[:, None, None] & np.uint64(0xFFFFFFFF)
    ctr[..., 3] = np.uint64(call & 0xFFFFFFFF)
    words = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(bm, na, quads * 4)[..., :nb]
    return ((words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def source_dense(group: np.ndarray, tokens: int) -> np.ndarray:
    """Dense fp32 (bm, tokens, n0) matrix of a compact source map (group[b, t] = merged token of original t, -1 = gone)."""
    return (group[:, None, :] == np.arange(tokens)[None, :, None]).astype(np.float32)


def decisions_equivalent(ref_scores: np.ndarray, plan_a: Plan, plan_b: Plan, ulps: float = 8.0):
    """True when two plans differ only where the score margins are within ``ulps`` fp32 ulp.

    Returns (ok, n_idx_mismatch, n_rank_mismatch)."""
    tol = ulps * 2.0 ** -24
    nm = ref_scores.max(-1)
    bad = 0
    n_idx = n_rank = 0
    for bi in range(ref_scores.shape[0]):
        # src membership and order
        for k in range(plan_a.r):
            ia, ib = plan_a.src_idx[bi, k], plan_b.src_idx[bi, k]
            if ia != ib:
                n_rank += 1
                if abs(float(nm[bi, ia]) - float(nm[bi, ib])) > tol:
                    bad += 1
            ja, jb = plan_a.dst_idx[bi, k], plan_b.dst_idx[bi, k]
            if ia == ib and ja != jb:
                n_idx += 1
                if abs(float(ref_scores[bi, ia, ja]) - float(ref_scores[bi, ia, jb])) > tol:
                    bad += 1
        for k in range(plan_a.unm_idx.shape[1]):
            ia, ib = plan_a.unm_idx[bi, k], plan_b.unm_idx[bi, k]
            if ia != ib:
                n_rank += 1
                if abs(float(nm[bi, ia]) - float(nm[bi, ib])) > tol and not plan_a.class_token:
                    bad += 1
    return bad == 0, n_idx, n_rank
