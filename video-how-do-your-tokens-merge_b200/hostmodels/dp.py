"""Data-parallel plumbing for the patched models: one process per GPU, clips split evenly over ranks,
weights replicated, no collective inside the model; logits (and small counters) are all-gathered
after the forward -- what the reference does in tools/test_net.py:159 with
slowfast/utils/distributed.py:25-44.  Works with NCCL (GPU) and gloo (CPU tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``global_batch`` clips: the first ``global_batch % world`` ranks get one
    extra clip (ragged batches are allowed; an empty shard is (k, k))."""
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_logits(local_logits: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather per-rank logits into (global_batch, classes), restoring clip order; shards may be ragged."""
    rank, ws = world()
    if ws == 1:
        return local_logits
    sizes = [shard_range(global_batch, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, local_logits.shape[1], dtype=local_logits.dtype, device=local_logits.device)
    pad[: local_logits.shape[0]] = local_logits
    out = torch.empty(ws * mx, local_logits.shape[1], dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def gather_counters(values) -> torch.Tensor:
    """Sum small integer counters (clips done, tokens merged ...) over ranks."""
    rank, ws = world()
    t = torch.as_tensor(values, dtype=torch.int64)
    if ws == 1:
        return t
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu()
