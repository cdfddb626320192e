"""The multi-GPU path is pure data parallelism (no collective inside the model): cover the sharding and
the logits / counter gathers with world_size-2 and -3 gloo groups on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")


def _worker(rank, world, port, global_batch, q):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        import test_models_golden as T           # tiny model factory + CPU port backend
        from hostmodels import dp
        import tome
        torch.set_num_threads(1)
        case = next(c for c in T.G.MODEL_CASES if c["name"] == "videomae_merge")
        model = T.G.seeded_fill(T.G.build_ours(case).eval())       # identical replicas: same seeded weights
        tome.patch.videomae(model)
        model.r = 40
        g = torch.Generator().manual_seed(5)
        clips = torch.rand(global_batch, 3, 4, 224, 224, generator=g)
        lo, hi = dp.shard_range(global_batch, rank, world)
        with T.port_backend(), torch.no_grad():
            local = model([clips[lo:hi]]) if hi > lo else torch.zeros(0, 10)
        full = dp.gather_logits(local, global_batch)
        counts = dp.gather_counters([hi - lo, int(model._tome_info["size"].shape[1]) if hi > lo else 0])
        if rank == 0:
            with T.port_backend(), torch.no_grad():
                want = model([clips])
            q.put((full.numpy(), want.numpy(), counts.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,global_batch", [(2, 4), (3, 4)])
def test_data_parallel_gather_matches_single_process(world, global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + global_batch
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, want, counts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import numpy as np
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)     # clip order restored, ragged shards handled
    assert counts[0] == global_batch


def test_shard_range_is_a_partition():
    from hostmodels import dp
    for gb in (0, 1, 7, 8, 64):
        for ws in (1, 2, 3, 8):
            spans = [dp.shard_range(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
