"""Full-size (ViT-B, 12 layers) model parity helpers shared by tests/test_fullsize_parity.py and
tools/parity_report.py.  Test infrastructure only.

Goldens: tests/golden/fullsize.npz, made by tests/golden/make_model_golden.py --full from the UNMODIFIED reference
models + reference tome patches on CPU (fp32): logits, final sizes and, per layer, the index lists the reference's
matching closures captured.  Two kinds of run of the CUDA path against them:

  * teacher-forced: every matching step replays the reference's own ``src/unm/dst`` lists (the plan is built by
    kernel 2 from keys that encode exactly that order), so only kernel 3 + the host model arithmetic differ from
    the reference: logits must agree to fp32 round-off (north_star: 1e-5 relative).
  * free-running: the CUDA path makes its own decisions (kernels 1 + 2 on the GPU-computed keys); reported: the
    first layer whose lists differ from the reference's, how many entries differ, logits error, top-1."""
import contextlib
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "golden", "make_model_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
FULL_CASES = G.FULL_CASES
_GOLD = None


def gold():
    global _GOLD
    if _GOLD is None:
        _GOLD = dict(np.load(os.path.join(HERE, "golden", "fullsize.npz")))
    return _GOLD


def layer_lists(name):
    g, out, i = gold(), [], 0
    while f"{name}/L{i}/src" in g:
        out.append({k: g[f"{name}/L{i}/{k}"] for k in ("src", "unm", "dst", "node_max", "gap2") if f"{name}/L{i}/{k}" in g})
        i += 1
    return out


def plan_from_lists(device, n, class_token, distill_token, rec):
    """A DevicePlan that reproduces the given lists exactly: kernel 2 ranks synthetic keys (src[k] -> 2 na - k,
    unm[k] -> na - k: distinct, descending in list order) and reads dst out of node_idx; for the hybrid threshold
    the reference's own node_max then replaces the keys."""
    from tome import _native
    src, unm = torch.from_numpy(rec["src"].astype(np.int64)), torch.from_numpy(rec["unm"].astype(np.int64))
    bm, r = src.shape
    na = (n + 1) // 2
    keys = torch.zeros(bm, na)
    idx = torch.zeros(bm, na, dtype=torch.int32)
    keys.scatter_(1, src, (2 * na - torch.arange(r, dtype=torch.float32)).expand(bm, r))
    keys.scatter_(1, unm, (na - torch.arange(unm.shape[1], dtype=torch.float32)).expand(bm, -1))
    if "dst" in rec:
        idx.scatter_(1, src, torch.from_numpy(rec["dst"].astype(np.int32)))
    plan = _native.select(keys.to(device), idx.to(device), n, r, class_token, distill_token)
    if "node_max" in rec:
        plan.node_max.copy_(torch.from_numpy(rec["node_max"]).to(device))
    return plan


@contextlib.contextmanager
def matching_hook(forced=None, record=None):
    """Replace tome.merge._make_plan: replay ``forced`` lists step by step and/or append every plan to ``record``."""
    import tome.merge as M
    real, step = M._make_plan, [0]

    def hooked(metric, r, class_token, distill_token, random_scores):
        if forced is not None:
            rec = forced[step[0]]
            dev = metric.keys.device if hasattr(metric, "keys") else metric.device
            assert rec["src"].shape[1] == r, (step[0], rec["src"].shape, r)
            plan = plan_from_lists(dev, metric.shape[1], class_token, distill_token, rec)
        else:
            plan = real(metric, r, class_token, distill_token, random_scores)
        step[0] += 1
        if record is not None:
            record.append(plan)
        return plan

    M._make_plan = hooked
    try:
        yield
    finally:
        M._make_plan = real


def compare_plans(plans, lists):
    """(first divergent layer or None, total differing entries, widest reference margin overturned at that first
    layer).  The margin is measured on the REFERENCE's own numbers: for two A tokens that swapped rank (or side of
    the src / unm boundary) the difference of their ``node_max``, for a changed destination the gap between the
    row's best and second-best score -- i.e. how close to a tie the reference's decision was."""
    first, diffs, margin = None, 0, 0.0
    for i, (p, rec) in enumerate(zip(plans, lists)):
        src, unm = p.src_idx.cpu().numpy(), p.unm_idx.cpu().numpy()
        ds, du = src != rec["src"], unm != rec["unm"]
        dd = (p.dst_idx.cpu().numpy() != rec["dst"]) & ~ds if "dst" in rec else np.zeros_like(ds)
        d = int(ds.sum() + du.sum() + dd.sum())
        diffs += d
        if d and first is None:
            first = i
            nm = rec["node_max"]
            for b in range(src.shape[0]):
                for ours, ref, bad, by_index in ((src[b], rec["src"][b], ds[b], False), (unm[b], rec["unm"][b], du[b], bool(p.class_token))):
                    if not bad.any():
                        continue
                    if by_index:
                        # ascending-index lists (merge.py:71-73): compare as sets -- every token that changed side of
                        # the src / unm boundary against the boundary value (the reference's r-th best node_max)
                        moved = np.setxor1d(ours, ref)
                        moved = moved[moved != 0]                       # the class token (key -inf) never moves
                        edge = nm[b][rec["src"][b][-1]]
                        if len(moved):
                            margin = max(margin, float(np.abs(nm[b][moved] - edge).max()))
                    else:
                        margin = max(margin, float(np.abs(nm[b][ours[bad]] - nm[b][ref[bad]]).max()))
                if dd[b].any():
                    margin = max(margin, float(rec["gap2"][b][src[b][dd[b]]].max()))
    return first, diffs, margin


def run_case(case, dtype=torch.float32, forced=False, device="cuda"):
    """One forward of the patched host model at full size on ``device``; returns a dict of parity numbers."""
    import tome
    g = gold()
    name = case["name"]
    model = G.seeded_fill(G.build_ours(case).eval(), wstd=case["wstd"]).to(device=device, dtype=dtype)
    clip = G.clip_for(case).to(device=device, dtype=dtype)
    with torch.no_grad():
        plain = model([clip]).float().cpu().numpy()
    plain_err = float(np.abs(plain - g[name + "/plain"]).max() / np.abs(g[name + "/plain"]).max())
    getattr(tome.patch, case["model"])(model, **case["kw"])
    model.r = case["r"]
    lists, plans = layer_lists(name), []
    with matching_hook(forced=lists if forced else None, record=plans), torch.no_grad():
        logits = model([clip]).float().cpu().numpy()
    want = g[name + "/tome"]
    first, diffs, margin = compare_plans(plans, lists)
    size = model._tome_info["size"].float().cpu().numpy()
    return dict(name=name, dtype=str(dtype).replace("torch.", ""), forced=forced,
                err=float(np.abs(logits - want).max() / np.abs(want).max()), unpatched_err=plain_err,
                top1_same=bool((logits.argmax(-1) == want.argmax(-1)).all()),
                top1_margin=float(np.sort(want, -1)[:, -1].min() - np.sort(want, -1)[:, -2].max()) / float(np.abs(want).max()),
                first_divergent_layer=first, differing_entries=diffs, overturned_margin=margin, layers=len(lists),
                size_shape_ok=size.shape == g[name + "/size"].shape,
                size_sum_ok=float(size.sum()) == float(g[name + "/size"].sum()))


@contextlib.contextmanager
def port_forced_backend(lists):
    """CPU stand-in for the teacher-forced run: the patches' matchers return oracle/torch_port.py closures built
    from the reference's lists instead of matching (host-model arithmetic at full size, no GPU)."""
    import sys
    from oracle import torch_port as P
    import tome  # noqa: F401
    step = [0]

    def match_from(metric, r, class_token, distill_token):
        rec = lists[step[0]]
        step[0] += 1
        m = P._Match.__new__(P._Match)
        m.t, m.r, m.class_token, m.distill_token = metric.shape[1], rec["src"].shape[1], class_token, distill_token
        as_idx = lambda a: torch.from_numpy(a.astype(np.int64))[..., None]          # noqa: E731
        m.src_idx, m.unm_idx = as_idx(rec["src"]), as_idx(rec["unm"])
        m.dst_idx = as_idx(rec["dst"]) if "dst" in rec else torch.zeros_like(m.src_idx)
        m.edge_idx = torch.cat((m.src_idx, m.unm_idx), 1)
        m.node_max = torch.from_numpy(rec["node_max"])
        return m

    def merge_fn(metric, r, class_token=False, distill_token=False, mode='merge', threshold=None):
        m = match_from(metric, r, class_token, distill_token)
        f = lambda x, mode="mean": m.merge(x, mode, threshold)                      # noqa: E731
        f.match = m
        return f, m.unmerge

    def drop_fn(metric, r, class_token=False, distill_token=False, mode='drop'):
        m = match_from(metric, r, class_token, distill_token)
        return lambda x: m.drop(x)

    repl = {"bipartite_soft_matching": merge_fn, "bipartite_soft_matching_drop": drop_fn,
            "bipartite_soft_matching_hybrid": merge_fn, "merge_wavg": P.merge_wavg, "merge_source": P.merge_source}
    saved = []
    try:
        for mn in ("tome.patch.videomae", "tome.patch.timesformer", "tome.patch.motionformer", "tome.patch.vivit"):
            mod = sys.modules[mn]
            for n, f in repl.items():
                if hasattr(mod, n):
                    saved.append((mod, n, getattr(mod, n)))
                    setattr(mod, n, f)
        yield
    finally:
        for mod, n, f in saved:
            setattr(mod, n, f)


def run_case_cpu_forced(case):
    import tome
    g, name = gold(), case["name"]
    model = G.seeded_fill(G.build_ours(case).eval(), wstd=case["wstd"])
    clip = G.clip_for(case)
    getattr(tome.patch, case["model"])(model, **case["kw"])
    model.r = case["r"]
    with port_forced_backend(layer_lists(name)), torch.no_grad():
        logits = model([clip]).numpy()
    want = g[name + "/tome"]
    return float(np.abs(logits - want).max() / np.abs(want).max()), model._tome_info["size"].float().numpy()
