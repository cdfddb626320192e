"""Torch-CPU port of the reference merge path  --  TEST / BASELINE INFRASTRUCTURE ONLY.

A restatement of /root/reference/tome/merge.py with the same ATen operator mix the reference
runs on CPU (norm, div, strided bmm, max, argsort, gather, scatter_reduce, cat), so that
timing it on the GPU box's host cores is a fair stand-in for the reference's own CPU path
(`/root/reference` does not exist on the GPU box).  Used only by:
  * ``bench.py``: the ``cpu_baseline`` leg and ``--impl reference`` (kind "port");
  * ``tests/``: CPU host-logic tests that need a merge backend without a GPU.
The product path never imports it.  Checked against the live reference and the golden
vectors in tests/test_torch_port.py.

API mirrors the reference module: bipartite_soft_matching(metric, r, class_token,
distill_token, mode) -> (merge, unmerge); bipartite_soft_matching_drop -> drop;
bipartite_soft_matching_hybrid(..., threshold); merge_wavg; merge_source.
``stable=True`` switches argsort to the stable rule the CUDA path implements.
"""
import math

import torch

STABLE = False          # tests flip this to compare with the canonical (stable) tie rule


def do_nothing(x, mode=None):
    return x


class _Match:
    """merge.py:36-73 once, shared by merge/unmerge/drop."""

    def __init__(self, metric, r, class_token, distill_token, random_scores):
        self.t = metric.shape[1]
        self.r = r
        self.class_token, self.distill_token = class_token, distill_token
        with torch.no_grad():
            if random_scores:
                na, nb = (self.t + 1) // 2, self.t // 2
                scores = torch.rand(size=(metric.size(0), na, nb), device=metric.device)
            else:
                unit = metric / metric.norm(dim=-1, keepdim=True)
                scores = unit[..., ::2, :] @ unit[..., 1::2, :].transpose(-1, -2)
            if class_token:
                scores[..., 0, :] = -math.inf
            if distill_token:
                scores[..., :, 0] = -math.inf
            self.node_max, node_idx = scores.max(dim=-1)
            self.edge_idx = self.node_max.argsort(dim=-1, descending=True, stable=STABLE)[..., None]
            self.unm_idx = self.edge_idx[..., r:, :]
            self.src_idx = self.edge_idx[..., :r, :]
            self.dst_idx = node_idx[..., None].gather(dim=-2, index=self.src_idx)
            if class_token:
                self.unm_idx = self.unm_idx.sort(dim=1)[0]

    def _cat(self, unm, dst):
        if self.distill_token:
            return torch.cat([unm[:, :1], dst[:, :1], unm[:, 1:], dst[:, 1:]], dim=1)
        return torch.cat([unm, dst], dim=1)

    def merge(self, x, mode="mean", threshold=None):
        a, b = x[..., ::2, :], x[..., 1::2, :]
        n, t1, c = a.shape
        r = self.r
        if threshold is not None:       # merge.py:326
            keep = (self.node_max[..., None].gather(dim=-2, index=self.edge_idx) >= threshold).type(b.type())
            b = b.scatter_reduce(-2, self.dst_idx.expand(n, r, c), keep.expand(n, -1, c), reduce='prod')
        unm = a.gather(dim=-2, index=self.unm_idx.expand(n, t1 - r, c))
        src = a.gather(dim=-2, index=self.src_idx.expand(n, r, c))
        b = b.scatter_reduce(-2, self.dst_idx.expand(n, r, c), src, reduce=mode)
        return self._cat(unm, b)

    def drop(self, x):
        a, b = x[..., ::2, :], x[..., 1::2, :]
        n, t1, c = a.shape
        return self._cat(a.gather(dim=-2, index=self.unm_idx.expand(n, t1 - self.r, c)), b)

    def unmerge(self, x):
        ul = self.unm_idx.shape[1]
        unm, dst = x[..., :ul, :], x[..., ul:, :]
        n, _, c = unm.shape
        src = dst.gather(dim=-2, index=self.dst_idx.expand(n, self.r, c))
        out = torch.zeros(n, self.t, c, device=x.device, dtype=x.dtype)
        out[..., 1::2, :] = dst
        out.scatter_(dim=-2, index=(2 * self.unm_idx).expand(n, ul, c), src=unm)
        out.scatter_(dim=-2, index=(2 * self.src_idx).expand(n, self.r, c), src=src)
        return out


def _clamp_r(metric, r, class_token, distill_token):
    return min(r, (metric.shape[1] - int(bool(class_token)) - int(bool(distill_token))) // 2)


def bipartite_soft_matching(metric, r, class_token=False, distill_token=False, mode='merge'):
    r = _clamp_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    m = _Match(metric, r, class_token, distill_token, mode == 'random_merge')
    merge = lambda x, mode="mean": m.merge(x, mode)          # noqa: E731
    merge.match = m
    return merge, m.unmerge


def bipartite_soft_matching_drop(metric, r, class_token=False, distill_token=False, mode='drop'):
    r = _clamp_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    m = _Match(metric, r, class_token, distill_token, mode == 'random_drop')
    drop = lambda x: m.drop(x)                               # noqa: E731
    drop.match = m
    return drop


def bipartite_soft_matching_hybrid(metric, r, class_token=False, distill_token=False, mode='merge', threshold=0.0):
    r = _clamp_r(metric, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    m = _Match(metric, r, class_token, distill_token, mode == 'random_merge')
    merge = lambda x, mode="mean": m.merge(x, mode, threshold)   # noqa: E731
    merge.match = m
    return merge, m.unmerge


def merge_wavg(merge, x, size=None):
    if size is None:
        size = torch.ones_like(x[..., 0, None])
    x = merge(x * size, mode="sum")
    size = merge(size, mode="sum")
    return x / size, size


def merge_source(merge, x, source=None):
    if source is None:
        n, t, _ = x.shape
        source = torch.eye(t, device=x.device)[None, ...].expand(n, t, t)
    return merge(source, mode="max")
