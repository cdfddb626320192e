"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against what
the reference produced (tests/golden).  Bit-exact for indices, sizes and fp32 features."""
import numpy as np
import pytest
import torch

import util
from oracle import tome_oracle as O

pytestmark = pytest.mark.gpu

SMALL = [c for c in util.CASES if not c.get("large")]
ALL_ACTIVE = [c for c in util.CASES if c["name"] not in ("r_zero", "n_one")]


@pytest.fixture(scope="module")
def native():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from tome import _native
    _native.device_check()
    return _native


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _oracle_plan(case, metric):
    return O.bipartite_soft_matching(metric, case["r"], bool(case.get("cls")), bool(case.get("distill")))


ALGOS = ["exact", "tc", "tc_streamed", "tc_nct5", "tc_pair_loads"]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("case", ALL_ACTIVE, ids=lambda c: c["name"])
def test_match_bit_exact_vs_oracle(native, case, algo, monkeypatch):
    """Every variant of kernel 1 gives the oracle's bits: fp64 SIMT tiles; tcgen05 with the
    exact refine fused in the epilogue (operands resident in smem); tcgen05 with streamed
    k-blocks + separate refine kernel; and a different column tiling."""
    metric, _, _ = util.case_arrays(case)
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    if algo == "tc_streamed":
        monkeypatch.setenv("TOME_TC_NO_FUSED_REFINE", "1")
    if algo == "tc_nct5":
        monkeypatch.setenv("TOME_TC_NCT", "5")
    if algo == "tc_pair_loads":             # the 4-byte-per-lane normalisation kernel (unaligned views take it)
        monkeypatch.setenv("TOME_SPLIT_PAIRS", "1")
    nm, ni = native.match(_dev(metric), cls, dis, algo=1 if algo == "exact" else 2)
    onm, oni = O.match(metric, cls, dis)
    np.testing.assert_array_equal(ni.cpu().numpy(), oni)
    np.testing.assert_array_equal(nm.cpu().numpy().view(np.uint32), onm.view(np.uint32))


@pytest.mark.parametrize("case", ALL_ACTIVE, ids=lambda c: c["name"])
def test_select_bit_exact_vs_oracle(native, case):
    metric, _, _ = util.case_arrays(case)
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    plan = _oracle_plan(case, metric)
    dp = native.select(_dev(plan.node_max), _dev(plan.node_idx), case["n"], plan.r, cls, dis)
    np.testing.assert_array_equal(dp.src_idx.cpu().numpy(), plan.src_idx)
    np.testing.assert_array_equal(dp.unm_idx.cpu().numpy(), plan.unm_idx)
    np.testing.assert_array_equal(dp.dst_idx.cpu().numpy(), plan.dst_idx)
    # derived maps: a_map inverse of unm/src; CSR groups sources by dst in ascending k
    a_map = dp.a_map.cpu().numpy(); b_off = dp.b_off.cpu().numpy(); b_src = dp.b_src.cpu().numpy()
    for b in range(plan.src_idx.shape[0]):
        for p, i in enumerate(plan.unm_idx[b]):
            assert a_map[b, i] == p
        for k, i in enumerate(plan.src_idx[b]):
            assert a_map[b, i] == -(plan.dst_idx[b, k] + 1)
        assert b_off[b, 0] == 0 and b_off[b, -1] == plan.r
        for j in range(plan.nb):
            want = [plan.src_idx[b, k] for k in range(plan.r) if plan.dst_idx[b, k] == j]
            assert list(b_src[b, b_off[b, j]:b_off[b, j + 1]]) == want
            head = dp.b_head[b, j].tolist()
            assert head[0] == len(want) and head[3] == b_off[b, j]
            assert head[1] == (want[0] if want else 0) and head[2] == (want[1] if len(want) > 1 else 0)


@pytest.mark.parametrize("case", ALL_ACTIVE, ids=lambda c: c["name"])
def test_select_reproduces_reference_lists_when_teacher_forced(native, case):
    """Fed the reference's own node_max/node_idx, kernel 2 must give the reference's index
    lists exactly (tie-free cases) -- the 'bit-exact merge assignment' bar."""
    if case["dist"] == "pm1":
        pytest.skip("reference argsort is unstable on exact ties; covered by the oracle test")
    g = util.golden(case["name"])
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    dp = native.select(_dev(g["node_max"]), _dev(g["node_idx"]), case["n"], int(g["r_eff"]), cls, dis)
    plan = O.Plan(case["n"], dp.r, cls, dis, dp.src_idx.cpu().numpy(), dp.unm_idx.cpu().numpy(),
                  dp.dst_idx.cpu().numpy(), g["node_max"], g["node_idx"])
    # ulps=0: only EXACTLY equal node_max values may be ordered differently (unstable argsort)
    stats = util.assert_plan_matches_golden(plan, g, case, ulps=0.0)
    assert stats["dst_diffs"] == 0
    if len(np.unique(g["node_max"][0])) == g["node_max"].shape[1] and all(
            len(np.unique(row)) == len(row) for row in g["node_max"]):
        assert util.plans_identical(plan, g)


def _device_plan_like_reference(native, case, g):
    """DevicePlan whose lists equal the golden (reference) lists."""
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    return native.select(_dev(g["node_max"]), _dev(g["node_idx"]), case["n"], int(g["r_eff"]), cls, dis)


@pytest.mark.parametrize("case", [c for c in SMALL if c["name"] not in ("r_zero", "n_one")], ids=lambda c: c["name"])
def test_merge_family_vs_reference_and_oracle(native, case):
    g = util.golden(case["name"])
    metric, x, size = util.case_arrays(case)
    thr = case.get("hybrid")
    tie = case["dist"] == "pm1"
    if tie:
        plan = _oracle_plan(case, metric)
        cls, dis = bool(case.get("cls")), bool(case.get("distill"))
        dp = native.select(_dev(plan.node_max), _dev(plan.node_idx), case["n"], plan.r, cls, dis)
        want_w, want_s = O.merge_wavg(plan, x, size, thr)
        want_mean, want_amax = O.merge(plan, x, "mean", thr), O.merge(plan, x, "max", thr)
        want_src, want_unm, want_drop = O.merge_source(plan, x, None, thr), O.unmerge(plan, want_w), O.drop(plan, x)
    else:
        dp = _device_plan_like_reference(native, case, g)
        want_w, want_s, want_mean, want_amax = g["x_wavg"], g["size_out"], g["x_mean"], g["x_amax"]
        want_src, want_unm, want_drop = g.get("source1"), g.get("x_unmerge"), g.get("x_drop")
    xd = _dev(x)
    sd = None if size is None else _dev(size)
    out, so, lo = native.merge(dp, xd, "wavg", size=sd, hybrid_threshold=thr, want_size=True)
    np.testing.assert_array_equal(so.cpu().numpy()[..., None], want_s)
    np.testing.assert_array_equal(out.cpu().numpy(), want_w)                       # bit-exact fp32
    np.testing.assert_allclose(lo.cpu().numpy(), np.log(want_s[..., 0]), rtol=2e-7, atol=1e-7)
    np.testing.assert_array_equal(native.merge(dp, xd, "mean", hybrid_threshold=thr).cpu().numpy(), want_mean)
    np.testing.assert_array_equal(native.merge(dp, xd, "max", hybrid_threshold=thr).cpu().numpy(), want_amax)
    np.testing.assert_array_equal(native.merge(dp, xd, "sum", hybrid_threshold=thr).cpu().numpy(),
                                  O.merge(O.Plan(case["n"], dp.r, dp.class_token, dp.distill_token,
                                                 dp.src_idx.cpu().numpy(), dp.unm_idx.cpu().numpy(),
                                                 dp.dst_idx.cpu().numpy(), dp.node_max.cpu().numpy(),
                                                 dp.node_idx.cpu().numpy()), x, "sum", thr))
    if want_src is not None:
        s1 = native.merge_source(dp, None, thr)
        np.testing.assert_array_equal(s1.cpu().numpy(), want_src)
        # explicit-identity path must agree with the implicit one
        eye = torch.eye(case["n"], device="cuda")[None].expand(case["bm"], -1, -1)
        np.testing.assert_array_equal(native.merge_source(dp, eye, thr).cpu().numpy(), want_src)
    if want_unm is not None:
        np.testing.assert_array_equal(native.unmerge(dp, out).cpu().numpy(), want_unm)
    if want_drop is not None and thr is None:
        np.testing.assert_array_equal(native.merge(dp, xd, "drop").cpu().numpy(), want_drop)


@pytest.mark.parametrize("name", ["config1_m1p", "config1_m1p_size", "config1_m1", "vivit_layer0"])
def test_full_size_end_to_end_vs_reference(native, name):
    """BASELINE configs[0] shapes: whole chain match -> select -> merge_wavg on the GPU,
    compared with what the reference produced (indices bit-exact unless the reference's own
    margin is within 8 ulp; sizes exact; features bit-exact on the committed subsample)."""
    import tome
    case = util.CASE_BY_NAME[name]
    g = util.golden(name)
    metric, x, size = util.case_arrays(case)
    cls = bool(case.get("cls"))
    md, xd = _dev(metric), _dev(x)
    sd = None if size is None else _dev(size)
    merge, unmerge = tome.merge.bipartite_soft_matching(md, case["r"], cls, False)
    plan = O.Plan(case["n"], merge.r, cls, False, merge.plan.src_idx.cpu().numpy(), merge.plan.unm_idx.cpu().numpy(),
                  merge.plan.dst_idx.cpu().numpy(), merge.plan.node_max.cpu().numpy(), merge.plan.node_idx.cpu().numpy())
    stats = util.assert_plan_matches_golden(plan, g, case)
    assert stats["dst_diffs"] == 0 and stats["src_swaps"] == 0 and stats["unm_swaps"] <= 4, stats
    out, so = tome.merge.merge_wavg(merge, xd, sd)
    assert out.shape == (case["bm"], case["n"] - case["r"], case["c"]) and so.shape == (case["bm"], case["n"] - case["r"], 1)
    if util.plans_identical(plan, g):
        np.testing.assert_array_equal(so.cpu().numpy(), g["size_out"])
        np.testing.assert_array_equal(out[:, ::37, ::5].cpu().numpy(), g["x_wavg_sub"])
        np.testing.assert_allclose(out.double().sum(dim=(1, 2)).cpu().numpy(), g["x_wavg_sum"], rtol=1e-9)
    # size-independent properties
    assert float(so.sum()) == case["bm"] * (case["n"] if size is None else 0) or size is not None
    if size is not None:
        assert float(so.sum()) == float(size.sum())
    back = unmerge(out)
    assert back.shape == xd.shape
    np.testing.assert_array_equal(back[:, 1::2].cpu().numpy()[:, :8], out[:, plan.unm_idx.shape[1]:][:, :8].cpu().numpy())
    src = tome.merge.merge_source(merge, xd, None)
    assert torch.all(src.sum(dim=1) == 1) and src.shape == (case["bm"], case["n"] - case["r"], case["n"])
    assert torch.equal(src.sum(dim=2), so[..., 0]) or size is not None


def test_bf16_merge_within_tolerance(native):
    case = util.CASE_BY_NAME["tokens_tsf"]
    g = util.golden(case["name"])
    metric, x, size = util.case_arrays(case)
    dp = _device_plan_like_reference(native, case, g)
    xb = _dev(x).bfloat16()
    out, so, _ = native.merge(dp, xb, "wavg", size=_dev(size), want_size=True)
    assert out.dtype == torch.bfloat16
    ref = O.merge_wavg(O.Plan(case["n"], dp.r, False, False, g["src_idx"], g["unm_idx"], g["dst_idx"], g["node_max"],
                              g["node_idx"]), xb.float().cpu().numpy(), size)[0]
    np.testing.assert_allclose(out.float().cpu().numpy(), ref, rtol=1e-2, atol=1e-2)   # north_star bf16 tolerance
    np.testing.assert_array_equal(so.cpu().numpy()[..., None], g["size_out"])
    # bf16 metric: matching runs in fp32 on the upcast values
    mb = _dev(metric).bfloat16()
    nm, ni = native.match(mb)
    onm, oni = O.match(mb.float().cpu().numpy())
    np.testing.assert_array_equal(ni.cpu().numpy(), oni)
    np.testing.assert_array_equal(nm.cpu().numpy(), onm)


def test_strided_views_match_contiguous(native):
    """TimeSformer layout 'b (p t) m -> (b t) p m' (timesformer.py:89-90) as addressing."""
    import ctypes
    torch.manual_seed(0)
    B, T, P, C, r = 2, 4, 50, 32, 9
    x = torch.randn(B, 1 + P * T, C, device="cuda")
    metric = torch.randn(B * T, P, 16, device="cuda")
    nm, ni = native.match(metric)
    dp = native.select(nm, ni, P, r)
    xr = x[:, 1:].reshape(B, P, T, C).permute(0, 2, 1, 3).reshape(B * T, P, C).contiguous()
    want = native.merge(dp, xr, "wavg", want_size=True)[0]                       # (B*T, P-r, C)
    want_full = torch.cat([x[:, :1], want.reshape(B, T, P - r, C).permute(0, 2, 1, 3).reshape(B, (P - r) * T, C)], 1)
    out = torch.empty(B, 1 + (P - r) * T, C, device="cuda")
    out[:, 0] = x[:, 0]
    lib = native.load_library()
    xv = native.TomeViewC(x.stride(0), x.stride(1), T * x.stride(1), T)
    ov = native.TomeViewC(out.stride(0), out.stride(1), T * out.stride(1), T)
    rc = lib.tome_merge(dp.c_ptr(), x[:, 1:].data_ptr(), 0, C, ctypes.byref(xv), None, 0, float("nan"),
                        out[:, 1:].data_ptr(), ctypes.byref(ov), None, None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.tome_last_error()
    torch.cuda.synchronize()
    assert torch.equal(out, want_full)


def test_random_modes_and_rowmax(native):
    import tome
    torch.manual_seed(3)
    s = torch.rand(3, 40, 39, device="cuda")
    nm, ni = native.rowmax(s, True, True)
    onm, oni = O.rowmax(np.where(np.arange(39)[None, None] == 0, -np.inf,
                                 np.where(np.arange(40)[None, :, None] == 0, -np.inf, s.cpu().numpy())).astype(np.float32))
    np.testing.assert_array_equal(ni.cpu().numpy(), oni)
    np.testing.assert_array_equal(nm.cpu().numpy(), onm)
    x = torch.randn(2, 60, 8, device="cuda")
    torch.manual_seed(5)
    merge, _ = tome.merge.bipartite_soft_matching(x, 7, mode="random_merge")
    torch.manual_seed(5)
    want_scores = torch.rand(2, 30, 30, device="cuda")     # the reference consumes the generator identically
    plan = O.bipartite_soft_matching(x.cpu().numpy(), 7, given_scores=want_scores.cpu().numpy())
    np.testing.assert_array_equal(merge.plan.src_idx.cpu().numpy(), plan.src_idx)
    np.testing.assert_array_equal(merge.plan.dst_idx.cpu().numpy(), plan.dst_idx)
    torch.manual_seed(5)
    drop = tome.merge.bipartite_soft_matching_drop(x, 7, mode="random_drop")
    np.testing.assert_array_equal(drop(x).cpu().numpy(), O.drop(plan, x.cpu().numpy()))


def test_errors_are_loud(native):
    import tome
    x = torch.randn(2, 16, 8, device="cuda")
    merge, _ = tome.merge.bipartite_soft_matching(x, 3)
    with pytest.raises(RuntimeError):
        merge(torch.randn(2, 18, 8, device="cuda"))           # wrong token count
    with pytest.raises(RuntimeError):
        merge(x.double())                                       # unsupported dtype
    with pytest.raises(RuntimeError):
        merge(x.cpu())                                          # CPU tensor: no fallback


def _tc_layout(native, bm, n, cm):
    """Geometry and workspace carving of the tensor-core pass, as the library reports it."""
    na, nb = (n + 1) // 2, n // 2
    n_ct, bn, off_max, off_cnt, _ = native.match_tc_describe(bm, n, cm)
    return na, nb, n_ct, bn, bm * na * n_ct, off_max, off_cnt


@pytest.mark.parametrize("name", ["config1_m1p", "config1_m1", "tokens_tsf", "vivit_layer0"])
def test_tensor_core_pass_really_prunes(native, name, monkeypatch):
    """The exact refine would hide a broken MMA pass (everything would overflow into the
    fallback).  Check the tcgen05 pass itself: its per-tile maxima equal the true maxima to
    within the error window, almost every tile keeps exactly one candidate, none overflow."""
    case = util.CASE_BY_NAME[name]
    metric, _, _ = util.case_arrays(case)
    cls = bool(case.get("cls"))
    bm, n, cm = metric.shape
    monkeypatch.setenv("TOME_TC_NO_FUSED_REFINE", "1")      # the streamed path leaves its pruning record in the workspace
    nm, ni, ws = native.match(_dev(metric), cls, False, algo=2, _return_workspace=True)
    na, nb, n_ct, bn, rows, off_max, off_cnt = _tc_layout(native, bm, n, cm)
    tile_max = ws[off_max:off_max + rows * 4].view(torch.float32).view(bm, na, n_ct).cpu().numpy()
    tile_cnt = ws[off_cnt:off_cnt + rows * 4].view(torch.int32).view(bm, na, n_ct).cpu().numpy()
    s = O.scores(metric, cls, False)
    eps = 5e-5 + 2e-7 * cm
    lo = 1 if cls else 0
    for c in range(n_ct):
        true = s[:, lo:, c * bn:min(nb, (c + 1) * bn)].max(-1)
        np.testing.assert_allclose(tile_max[:, lo:, c], true, rtol=0, atol=eps)
    cnt = tile_cnt[:, lo:]
    assert (cnt == 255).sum() == 0, "no overflow expected on tie-free data"
    assert cnt.min() >= 1
    assert (cnt == 1).mean() > (0.97 if cm <= 64 else 0.9), (cnt == 1).mean()
    # measured error of the 3-product bf16 pass, for DESIGN.md
    err = 0.0
    for c in range(n_ct):
        true = s[:, lo:, c * bn:min(nb, (c + 1) * bn)].max(-1)
        err = max(err, float(np.abs(tile_max[:, lo:, c] - true).max()))
    print(f"[tc-pass] {name}: cm={cm} max |approx-true| = {err:.3e} (bound {eps:.3e}), "
          f"single-candidate tiles {(cnt == 1).mean() * 100:.2f}%")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_layernorm_matches_separate_pass(native, dtype):
    """tome_merge_norm: the merged rows are bit-identical to tome_merge, and the fused LayerNorm equals
    torch's LayerNorm of those rows (fp32: round-off; bf16: one output ulp)."""
    case = util.CASE_BY_NAME["tokens_tsf"]
    g = util.golden(case["name"])
    metric, x, size = util.case_arrays(case)
    dp = _device_plan_like_reference(native, case, g)
    gen = torch.Generator(device="cuda").manual_seed(1)
    c = 768
    xd = torch.randn(case["bm"], case["n"], c, device="cuda", generator=gen).to(dtype)
    w = (1 + 0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    b = (0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    plain, s0, l0 = native.merge(dp, xd, "wavg", size=_dev(size), want_size=True)
    out, s1, l1, normed = native.merge(dp, xd, "wavg", size=_dev(size), want_size=True, norm=(w, b, 1e-6))
    assert torch.equal(out, plain) and torch.equal(s0, s1) and torch.equal(l0, l1)
    want = torch.nn.functional.layer_norm(out.float(), (c,), w.float(), b.float(), 1e-6)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    torch.testing.assert_close(normed.float(), want, **tol)
    # frame layout (class token row normalised on the side)
    B, T, P = 2, 4, case["n"]
    xf = torch.randn(B, 1 + P * T, c, device="cuda", generator=gen).to(dtype)
    o2, _, _, n2 = native.merge_frames(dp, xf, T, "wavg", norm=(w, b, 1e-6))
    want2 = torch.nn.functional.layer_norm(o2.float(), (c,), w.float(), b.float(), 1e-6)
    torch.testing.assert_close(n2.float(), want2, **tol)
    assert torch.equal(o2, native.merge_frames(dp, xf, T, "wavg")[0])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_match_heads_equals_match_of_mean(native, dtype):
    """tome_match_heads (head-mean taken in kernel 1's prologue) against tome_match on k.mean(1)."""
    gen = torch.Generator(device="cuda").manual_seed(2)
    B, H, N, d = 3, 12, 197, 64
    qkv = torch.randn(B, N, 3, H, d, device="cuda", generator=gen).to(dtype)
    k = qkv.permute(2, 0, 3, 1, 4)[1]                                  # (B, H, N, d) strided view, as in the patches
    nm, ni = native.match_heads(native.HeadMeanMetric(k), True, False)
    if dtype == torch.float32:
        # the kernel adds heads 0..H-1 sequentially in fp32: same values as an explicit sequential sum
        seq = torch.zeros(B, N, d, device="cuda")
        for h in range(H):
            seq = seq + k[:, h]
        ref = seq * (1.0 / H)                                           # ATen's mean multiplies by 1/H
    else:
        ref = k.float().mean(1).to(dtype)
    onm, oni = O.match(ref.float().cpu().numpy(), True, False)
    agree = (ni.cpu().numpy() == oni).mean()
    assert agree > 0.999, agree                                         # bf16: the mean's rounding may differ in rare ties
    if dtype == torch.float32:
        np.testing.assert_array_equal(ni.cpu().numpy(), oni)
        np.testing.assert_array_equal(nm.cpu().numpy(), onm)
    # Motionformer regrouping: (B, H, S*F, d) with tokens '(s f)' -> batch (b f)
    F_, S = 4, 49
    k2 = torch.randn(2, H, S * F_, d, device="cuda", generator=gen).to(dtype)
    m2 = native.HeadMeanMetric(k2, frames=F_)
    nm2, ni2 = native.match_heads(m2)
    nm3, ni3 = native.match(m2.materialize().contiguous())
    assert (ni2 == ni3).float().mean() > 0.999


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(8, 1568, 768), (3, 197, 384), (2, 50, 1024)])
def test_add_layernorm_matches_two_passes(native, dtype, shape):
    """tome_add_layernorm: the sum is bit-identical to torch's a + b, the LayerNorm equals torch's
    LayerNorm of that sum (fp32: round-off; bf16: one output ulp)."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    c = shape[-1]
    a = torch.randn(*shape, device="cuda", generator=gen).to(dtype)
    b = torch.randn(*shape, device="cuda", generator=gen).to(dtype)
    w = (1 + 0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    bias = (0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    s, y = native.add_layernorm(a, b, (w, bias, 1e-6))
    assert torch.equal(s, a + b)
    want = torch.nn.functional.layer_norm((a + b).float(), (c,), w.float(), bias.float(), 1e-6)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    torch.testing.assert_close(y.float(), want, **tol)
    s2, y2 = native.add_layernorm(a, b, (w, None, 1e-6))
    want2 = torch.nn.functional.layer_norm((a + b).float(), (c,), w.float(), None, 1e-6)
    torch.testing.assert_close(y2.float(), want2, **tol)
    with pytest.raises(RuntimeError):
        native.add_layernorm(a, b[:, :-1], (w, bias, 1e-6))


@pytest.mark.parametrize("lead", [0, 1])
@pytest.mark.parametrize("shape", [(2, 12, 392, 64), (3, 3, 197, 32)])
def test_prop_attention_key_bias_equals_masked_attention(native, shape, lead):
    """tome.attention: log(size) folded into two spare q/k channels (tome_attn_key_bias) + unmasked fused
    attention == the reference's masked formulation (videomae.py:62-63; timesformer.py:72-74 with lead=1),
    to bf16 round-off, and the key tensor handed to the matching is the plain projection."""
    from tome import attention as A
    B, H, N, d = shape
    C = H * d
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(B, N, C, device="cuda", generator=gen).to(torch.bfloat16)
    w = (0.05 * torch.randn(3 * C, C, device="cuda", generator=gen)).to(torch.bfloat16)
    bias = (0.1 * torch.randn(3 * C, device="cuda", generator=gen)).to(torch.bfloat16)
    size = torch.randint(1, 30, (B, N - lead, 1), device="cuda", generator=gen).float()
    owner = torch.nn.Module().eval()
    wq, wk, wv = w.chunk(3, 0)
    bq, bk, bv = bias.chunk(3, 0)
    with torch.no_grad():
        ctx, k = A.attention(x, owner, H, d, d ** -0.5, size.log(), wq, wk, wv, bq, bk, bv, lead=lead)
        qkv = torch.nn.functional.linear(x, w, bias)
        q0, k0, v0 = qkv.reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        assert torch.equal(k, k0)
        mask = torch.zeros(B, 1, N, N, device="cuda")
        mask[:, :, lead:, lead:] = size.log()[:, None, None, :, 0]
        want = torch.nn.functional.scaled_dot_product_attention(q0.float(), k0.float(), v0.float(), attn_mask=mask, scale=d ** -0.5)
        want = want.transpose(1, 2).reshape(B, N, C)
        base = torch.nn.functional.scaled_dot_product_attention(q0, k0, v0, attn_mask=mask.to(torch.bfloat16), scale=d ** -0.5)
        base = base.transpose(1, 2).reshape(B, N, C).float()
    err = (ctx.float() - want).abs().max().item()
    err_masked_bf16 = (base - want).abs().max().item()
    print(f"[prop-attn] {shape} lead={lead}: folded {err:.2e} vs masked bf16 {err_masked_bf16:.2e}")
    assert err < max(2 * err_masked_bf16, 4e-3)
    # the cached padded weight is rebuilt when a source tensor changes in place
    w.mul_(2.0)
    with torch.no_grad():
        ctx2, _ = A.attention(x, owner, H, d, d ** -0.5, size.log(), wq, wk, wv, bq, bk, bv, lead=lead)
    assert not torch.equal(ctx2, ctx)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["tokens_tsf", "tokens_hybrid", "gauss_cls_distill"])
def test_fused_residual_equals_merging_the_sum(native, name, dtype):
    """tome_merge_add_norm: merging (x, residual) in one pass is bit-identical to merging the materialised
    x + residual -- merged rows, sizes, log sizes and the fused LayerNorm."""
    case = util.CASE_BY_NAME[name]
    g = util.golden(case["name"])
    metric, x, size = util.case_arrays(case)
    dp = _device_plan_like_reference(native, case, g)
    gen = torch.Generator(device="cuda").manual_seed(5)
    c = 768
    xd = torch.randn(case["bm"], case["n"], c, device="cuda", generator=gen).to(dtype)
    rd = torch.randn(case["bm"], case["n"], c, device="cuda", generator=gen).to(dtype)
    w = (1 + 0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    b = (0.1 * torch.randn(c, device="cuda", generator=gen)).to(dtype)
    sz = _dev(size) if size is not None else None
    thr = case.get("hybrid")
    want = native.merge(dp, xd + rd, "wavg", size=sz, want_size=True, norm=(w, b, 1e-6), hybrid_threshold=thr)
    got = native.merge(dp, xd, "wavg", size=sz, want_size=True, norm=(w, b, 1e-6), residual=rd, hybrid_threshold=thr)
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)
    got2 = native.merge(dp, xd, "wavg", size=sz, want_size=True, residual=rd, hybrid_threshold=thr)
    for a_, b_ in zip(got2, want[:3]):
        assert torch.equal(a_, b_)
    with pytest.raises(RuntimeError):
        native.merge(dp, xd, "wavg", residual=rd[:, :-1])


@pytest.mark.timeout(180)
@pytest.mark.parametrize("path", ["auto", "chain", "chain_two_select", "one_launch", "exact_simt"])
@pytest.mark.parametrize("case", ALL_ACTIVE, ids=lambda c: c["name"])
def test_plan_build_bit_exact_vs_oracle(native, case, path, monkeypatch):
    """tome_plan_build (kernels 1 + 2 in one ABI call) gives the oracle's node_max / node_idx / src / unm / dst
    bits: as dispatched by default, as the three-launch tcgen05 chain (normalise, match, one-launch select on the packed
    keys), as the same chain with the rank + finish pair, as the one-launch cluster kernel (forced for every shape it
    supports) and with the exact SIMT matching."""
    metric, _, _ = util.case_arrays(case)
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    plan = _oracle_plan(case, metric)
    if path in ("chain", "chain_two_select", "one_launch"):
        monkeypatch.setenv("TOME_PLAN_CLUSTER", "1" if path == "one_launch" else "0")
    if path == "chain_two_select":
        monkeypatch.setenv("TOME_SELECT_TWO", "1")
    dp = native.plan_build(_dev(metric), plan.r, cls, dis, algo=1 if path == "exact_simt" else 0)
    np.testing.assert_array_equal(dp.node_idx.cpu().numpy(), plan.node_idx)
    np.testing.assert_array_equal(dp.node_max.cpu().numpy().view(np.uint32), plan.node_max.view(np.uint32))
    np.testing.assert_array_equal(dp.src_idx.cpu().numpy(), plan.src_idx)
    np.testing.assert_array_equal(dp.unm_idx.cpu().numpy(), plan.unm_idx)
    np.testing.assert_array_equal(dp.dst_idx.cpu().numpy(), plan.dst_idx)
    # the CSR the merge kernel gathers through is the same whichever select built it
    monkeypatch.setenv("TOME_SELECT_TWO", "0" if path == "chain_two_select" else "1")
    ref = native.select(_dev(plan.node_max), _dev(plan.node_idx), case["n"], plan.r, cls, dis)
    for name in ("a_map", "b_off", "b_src", "b_head"):
        assert torch.equal(getattr(dp, name), getattr(ref, name)), name


@pytest.mark.parametrize("dtypes", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
def test_patchify_equals_permute_and_conv(native, dtypes):
    """tome_patchify: the tubelet rows equal torch's reshape/permute (bit for bit, cast included), and the
    GEMM over them equals the Conv3d the reference runs (videomae builder:138-160)."""
    din, dout = dtypes
    gen = torch.Generator(device="cuda").manual_seed(6)
    B, C, T, H, W, tt, ph, pw = 2, 3, 4, 32, 48, 2, 16, 16
    x = torch.rand(B, C, T, H, W, device="cuda", generator=gen).to(din)
    got = native.patchify(x, tt, ph, pw, dout)
    want = x.reshape(B, C, T // tt, tt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
    want = want.reshape(B, (T // tt) * (H // ph) * (W // pw), C * tt * ph * pw).to(dout)
    assert torch.equal(got, want)
    if din == dout == torch.float32:
        conv = torch.nn.Conv3d(C, 8, (tt, ph, pw), (tt, ph, pw)).cuda()
        ref = conv(x).flatten(2).transpose(1, 2)
        out = torch.nn.functional.linear(got, conv.weight.reshape(8, -1), conv.bias)
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dout", [torch.float32, torch.bfloat16])
def test_patchify_takes_uint8_frames(native, dout):
    """Decoder-style uint8 clips: tome_patchify converts value / 255 in the same pass, bit-identical to
    `x.float() / 255` followed by the fp32 / bf16 patchify (19 MB instead of 77 MB over PCIe per 8 clips)."""
    gen = torch.Generator(device="cuda").manual_seed(8)
    B, C, T, H, W, tt, ph, pw = 2, 3, 4, 32, 48, 2, 16, 16
    x8 = torch.randint(0, 256, (B, C, T, H, W), device="cuda", generator=gen, dtype=torch.uint8)
    got = native.patchify(x8, tt, ph, pw, dout)
    # correctly rounded division, i.e. torch's CPU `x.float() / 255` (its CUDA kernel multiplies by fp32(1 / 255) instead)
    want = native.patchify((x8.cpu().float() / 255).cuda(), tt, ph, pw, dout)
    assert torch.equal(got, want)
    import hostmodels
    m = hostmodels.VideoMAE(arch="vit_small_patch16_224", num_classes=7, num_frames=4).eval().cuda().to(dout)
    clip = torch.randint(0, 256, (1, 3, 4, 224, 224), device="cuda", generator=gen, dtype=torch.uint8)
    with torch.no_grad():
        a, b = m([clip]), m([(clip.float() / 255).to(dout)])
    torch.testing.assert_close(a, b, rtol=1e-5 if dout == torch.float32 else 2e-2, atol=1e-5 if dout == torch.float32 else 2e-2)


def test_merge_writes_into_caller_owned_buffers(native):
    case = util.CASE_BY_NAME["gauss_odd"]
    metric, x, size = util.case_arrays(case)
    dp = native.plan_build(_dev(metric), case["r"])
    xd = _dev(x)
    want = native.merge(dp, xd, "wavg", size=_dev(size), want_size=True)
    bufs = (torch.empty_like(want[0]), torch.empty_like(want[1]), torch.empty_like(want[2]))
    got = native.merge(dp, xd, "wavg", size=_dev(size), want_size=True, out=bufs)
    assert all(g.data_ptr() == b.data_ptr() for g, b in zip(got, bufs))
    assert all(torch.equal(g, w) for g, w in zip(got, want))
    with pytest.raises(RuntimeError, match="out buffer"):
        native.merge(dp, xd, "wavg", out=(torch.empty(1, 2, 3, device="cuda"),))


def test_add_layernorm_broadcasts_the_position_embedding(native):
    gen = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(4, 50, 768, device="cuda", generator=gen).to(torch.bfloat16)
    pos = torch.randn(1, 50, 768, device="cuda", generator=gen).to(torch.bfloat16)
    w = torch.ones(768, device="cuda", dtype=torch.bfloat16)
    s, y = native.add_layernorm(a, pos, (w, None, 1e-6))
    assert torch.equal(s, a + pos)
    want = torch.nn.functional.layer_norm((a + pos).float(), (768,), w.float(), None, 1e-6)
    torch.testing.assert_close(y.float(), want, rtol=1.6e-2, atol=1.6e-2)


@pytest.mark.parametrize("name", ["gauss_small", "gauss_odd", "gauss_cls", "tokens_tsf", "tokens_hybrid", "tokens_hybrid_cls"])
def test_merge_backward_matches_reference_autograd(native, name):
    """Training through the merge: gradients of merge_wavg / merge(sum, mean) / drop from the
    kernel's unmerge-based backward equal autograd through the reference's gather / scatter_reduce closures
    (the CPU port, pinned to tome/merge.py by tests/test_torch_port.py) on the same plan."""
    import tome
    from oracle import torch_port as P
    case = util.CASE_BY_NAME[name]
    metric, x, size = util.case_arrays(case)
    cls, dis = bool(case.get("cls")), bool(case.get("distill"))
    thr = case.get("hybrid")
    mt = torch.from_numpy(metric)
    if thr is None:
        ref_merge, _ = P.bipartite_soft_matching(mt, case["r"], cls, dis)
        our_merge, _ = tome.merge.bipartite_soft_matching(mt.cuda(), case["r"], cls, dis)
    else:
        ref_merge, _ = P.bipartite_soft_matching_hybrid(mt, case["r"], cls, dis, "hybrid", thr)
        our_merge, _ = tome.merge.bipartite_soft_matching_hybrid(mt.cuda(), case["r"], cls, dis, "hybrid", thr)
    gen = torch.Generator().manual_seed(8)
    xs = torch.from_numpy(x)
    sz = None if size is None else torch.from_numpy(size)
    wgt = torch.randn(xs.shape[0], xs.shape[1] - our_merge.r, xs.shape[2], generator=gen)

    def grads(merge_wavg, merge, xin, szin, w):
        out = []
        a = xin.clone().requires_grad_(True)
        y, s = merge_wavg(merge, a, szin)
        (y * w).sum().backward()
        out.append(a.grad.detach().cpu())
        for mode in ("sum", "mean"):
            a = xin.clone().requires_grad_(True)
            (merge(a, mode=mode) * w).sum().backward()
            out.append(a.grad.detach().cpu())
        return out

    got = grads(tome.merge.merge_wavg, our_merge, xs.cuda(), None if sz is None else sz.cuda(), wgt.cuda())
    if thr is None:
        want = grads(P.merge_wavg, ref_merge, xs, sz, wgt)
        for g, w_ in zip(got, want):
            torch.testing.assert_close(g, w_, rtol=1e-5, atol=1e-6)
    else:
        # the reference cannot differentiate its hybrid closure (scatter_reduce 'prod' with an (na)-row mask
        # against an r-row index, merge.py:326, fails in autograd); the merge is linear in x, so check the
        # adjoint identity <L u, w> == <u, L^T w> instead
        u = torch.randn(xs.shape, generator=gen)
        with torch.no_grad():
            lu = [tome.merge.merge_wavg(our_merge, u.cuda(), None if sz is None else sz.cuda())[0],
                  our_merge(u.cuda(), mode="sum"), our_merge(u.cuda(), mode="mean")]
        for y, g in zip(lu, got):
            lhs, rhs = float((y.cpu().double() * wgt.double()).sum()), float((u.double() * g.double()).sum())
            assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), (lhs, rhs)
    if thr is None:                                    # drop mode shares the matching
        ref_drop = P.bipartite_soft_matching_drop(mt, case["r"], cls, dis)
        our_drop = tome.merge.bipartite_soft_matching_drop(mt.cuda(), case["r"], cls, dis)
        a = xs.clone().requires_grad_(True)
        (ref_drop(a) * wgt).sum().backward()
        b = xs.cuda().requires_grad_(True)
        (our_drop(b) * wgt.cuda()).sum().backward()
        torch.testing.assert_close(b.grad.cpu(), a.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape", [(300, 256, 64), (12544, 3072, 768), (3744, 3072, 768), (129, 512, 200)])
@pytest.mark.parametrize("gelu", [True, False, "gelu_fast"])
def test_linear_gelu_matches_linear_then_gelu(native, shape, gelu):
    """tome_linear_gelu (tcgen05 GEMM, bias + erf GELU in the epilogue) against F.linear followed by F.gelu on
    the same bf16 tensors: the two GEMMs accumulate in different orders, so outputs may differ by one bf16
    ulp of the pre-activation; against the fp32 reference both are equally close."""
    m, n, k = shape
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(m, k, device="cuda", generator=gen).to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda", generator=gen) * k ** -0.5).to(torch.bfloat16)
    b = (0.1 * torch.randn(n, device="cuda", generator=gen)).to(torch.bfloat16)
    with torch.no_grad():
        assert native.linear_gelu_supported(x, w, b)
    from hostmodels.vivit import gelu_fast                   # HF FastGELUActivation, ViViT's hidden_act
    act = {True: torch.nn.functional.gelu, False: (lambda t: t), "gelu_fast": gelu_fast}[gelu]
    got = native.linear_gelu(x, w, b, gelu=gelu).float()
    pre = torch.nn.functional.linear(x, w, b)
    lib = act(pre).float()
    ref = act(x.float() @ w.float().t() + b.float())
    err_ours, err_lib = (got - ref).abs().max().item(), (lib - ref).abs().max().item()
    print(f"[linear-gelu] {shape} gelu={gelu}: max err vs fp32 ours {err_ours:.3e} library {err_lib:.3e}")
    assert err_ours <= 1.5 * err_lib + 1e-3
    assert (got - lib).abs().max().item() <= 4e-2
    got3 = native.linear_gelu(x.reshape(1, m, k), w, None, gelu=gelu)
    assert got3.shape == (1, m, n)


def test_side_stream_matching_gives_the_same_plan(native, monkeypatch):
    """tome.merge.prefetch_matching (TOME_PREFETCH=1: kernels 1 + 2 on a side stream as soon as K exists) hands
    bipartite_soft_matching the same plan as the in-line path, for the same (r, class token) request only."""
    import tome
    gen = torch.Generator(device="cuda").manual_seed(10)
    k = torch.randn(2, 197, 3, 12, 64, device="cuda", generator=gen).to(torch.bfloat16).permute(2, 0, 3, 1, 4)[1]
    base, _ = tome.merge.bipartite_soft_matching(native.HeadMeanMetric(k), 40, True)
    monkeypatch.setenv("TOME_PREFETCH", "1")
    with torch.no_grad():
        m = native.HeadMeanMetric(k)
        tome.merge.prefetch_matching(m, 40, True)
        assert m.prefetched is not None
        pre, _ = tome.merge.bipartite_soft_matching(m, 40, True)
        assert m.prefetched is None
        m2 = native.HeadMeanMetric(k)
        tome.merge.prefetch_matching(m2, 40, True)
        other, _ = tome.merge.bipartite_soft_matching(m2, 30, True)      # different r: the parked plan is not used
    torch.cuda.synchronize()
    for name in ("src_idx", "unm_idx", "dst_idx", "b_head"):
        assert torch.equal(getattr(pre.plan, name), getattr(base.plan, name)), name
    assert other.plan.r == 30


@pytest.mark.parametrize("cm", [64, 256])
@pytest.mark.parametrize("cls", [False, True])
def test_zero_norm_tokens_give_nan_like_the_reference(native, cls, cm):
    """The reference has no eps in its normalisation (merge.py:51): a zero metric row becomes NaN, NaN wins every
    max it takes part in (first NaN column) and sorts above +inf.  The tensor-core path used to HANG on such a
    row (a lane-divergent tcgen05.ld); it must give the oracle's bits, as the exact SIMT path does.  Seen in the
    wild: HuggingFace-initialised ViViT has a zero class token and zero position embeddings."""
    import warnings
    gen = torch.Generator().manual_seed(11)
    metric = torch.randn(2, 300, cm, generator=gen)
    metric[0, 0] = 0          # an A token (the class token when cls)
    metric[0, 7] = 0          # a B token: every A row of batch 0 sees a NaN in column 3
    metric[1, 10] = 0         # an A token that is not the class token
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        onm, oni = O.match(metric.numpy(), cls, False)
    for algo in (1, 2):
        nm, ni = native.match(metric.cuda(), cls, False, algo=algo)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(ni.cpu().numpy(), oni)
        np.testing.assert_array_equal(np.isnan(nm.cpu().numpy()), np.isnan(onm))
        ok = ~np.isnan(onm)
        np.testing.assert_array_equal(nm.cpu().numpy()[ok], onm[ok])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = O.bipartite_soft_matching(metric.numpy(), 40, cls, False)
    dp = native.plan_build(metric.cuda(), plan.r, cls, False)
    np.testing.assert_array_equal(dp.src_idx.cpu().numpy(), plan.src_idx)
    np.testing.assert_array_equal(dp.unm_idx.cpu().numpy(), plan.unm_idx)
    np.testing.assert_array_equal(dp.dst_idx.cpu().numpy(), plan.dst_idx)
    # and through the lazy head-mean path (bf16 K with a zero token)
    k = torch.randn(2, 197, 3, 12, 64, generator=gen).to(torch.bfloat16)
    k[:, 0, 1] = 0
    kd = k.cuda().permute(2, 0, 3, 1, 4)[1]
    nm2, ni2 = native.match_heads(native.HeadMeanMetric(kd), cls, False)
    nm3, ni3 = native.match(kd.float().mean(1).to(torch.bfloat16).contiguous(), cls, False, algo=1)
    torch.cuda.synchronize()
    assert torch.equal(ni2, ni3) and torch.equal(torch.isnan(nm2), torch.isnan(nm3))


@pytest.mark.parametrize("cls", [False, True])
def test_tiny_token_counts_down_to_two(native, cls):
    """The ViViT r sweep (experiments.sh:395-428) drives the token count down to 2-3 in the last layers: every n
    from 2 up, r clamped to its maximum, through tome_plan_build and the merge, against the oracle."""
    gen = torch.Generator().manual_seed(12)
    for n in (2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 65):
        metric = torch.randn(3, n, 64, generator=gen)
        x = torch.randn(3, n, 32, generator=gen)
        plan = O.bipartite_soft_matching(metric.numpy(), 10 ** 6, cls, False)
        if plan is None or plan.r <= 0:
            continue
        dp = native.plan_build(metric.cuda(), plan.r, cls, False)
        np.testing.assert_array_equal(dp.node_idx.cpu().numpy(), plan.node_idx, err_msg=f"n={n}")
        np.testing.assert_array_equal(dp.src_idx.cpu().numpy(), plan.src_idx, err_msg=f"n={n}")
        np.testing.assert_array_equal(dp.unm_idx.cpu().numpy(), plan.unm_idx, err_msg=f"n={n}")
        np.testing.assert_array_equal(dp.dst_idx.cpu().numpy(), plan.dst_idx, err_msg=f"n={n}")
        out, so, _ = native.merge(dp, x.cuda(), "wavg", want_size=True)
        want, want_s = O.merge_wavg(plan, x.numpy(), None)
        np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg=f"n={n}")
        np.testing.assert_array_equal(so.cpu().numpy(), want_s.reshape(so.shape), err_msg=f"n={n}")


# ---- the upstream-ToMe set matchers (merge.py:105-212) against goldens made by the reference functions ----
def _sets_golden():
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_sets_golden", os.path.join(here, "make_sets_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    import sys
    sys.path.insert(0, here)
    spec.loader.exec_module(mod)
    return mod, dict(np.load(os.path.join(here, "sets.npz")))


_SETS_MOD, _SETS_GOLD = _sets_golden()


@pytest.mark.parametrize("case", _SETS_MOD.SET_CASES, ids=lambda c: c["name"])
def test_kth_and_random_matchers_vs_reference(native, case, monkeypatch):
    """kth_ / random_bipartite_soft_matching on tome_match_sets / tome_group_reduce / tome_gather_rows: the
    destination of every source, merge(sum / mean / amax) and unmerge equal what the reference functions
    produced on CPU, fp32 bit for bit (same reduction order: destination first, then sources ascending)."""
    import tome.merge as M
    metric, x, rand = _SETS_MOD.inputs(case)
    g, name = _SETS_GOLD, case["name"]
    if case["kind"] == "kth":
        merge, unmerge = M.kth_bipartite_soft_matching(metric.cuda(), case["k"])
    else:
        real_rand = torch.rand
        monkeypatch.setattr(torch, "rand", lambda *a, **k: rand.clone().to(k.get("device", "cpu")))
        merge, unmerge = M.random_bipartite_soft_matching(metric.cuda(), case["r"])
        monkeypatch.setattr(torch, "rand", real_rand)
    np.testing.assert_array_equal(merge.dst_idx[..., 0].cpu().numpy(), g[name + "/dst_idx"])
    for mode in ("mean", "sum", "amax"):
        got = merge(x.cuda(), mode=mode).cpu().numpy()
        np.testing.assert_array_equal(got.view(np.uint32), g[f"{name}/merge_{mode}"].view(np.uint32), err_msg=mode)
    got = unmerge(merge(x.cuda(), mode="mean")).cpu().numpy()
    np.testing.assert_array_equal(got.view(np.uint32), g[name + "/unmerge"].view(np.uint32))
    # bf16 rows: fp32 accumulation, one rounding
    half = merge(x.cuda().bfloat16(), mode="mean").float().cpu().numpy()
    np.testing.assert_allclose(half, g[name + "/merge_mean"], rtol=2e-2, atol=2e-2)
    assert M.kth_bipartite_soft_matching(metric.cuda(), 1) == (M.do_nothing, M.do_nothing)
    assert M.random_bipartite_soft_matching(metric.cuda(), 0) == (M.do_nothing, M.do_nothing)


def test_unmerge_backward_is_merge_sum(native):
    """The reference's unmerge (gather + scatter, merge.py:87-100) is differentiable; ours runs tome_unmerge forward
    and the merge kernel in 'sum' mode backward.  Checked against autograd through the CPU port's unmerge."""
    import tome.merge as M
    from oracle import torch_port as P
    case = util.CASE_BY_NAME["gauss_odd"]
    metric, x, _ = util.case_arrays(case)
    merge, unmerge = M.bipartite_soft_matching(_dev(metric), case["r"])
    pm, pu = P.bipartite_soft_matching(torch.from_numpy(metric), case["r"])
    P_STABLE = P.STABLE
    y = torch.from_numpy(x)[:, :case["n"] - merge.r].clone()
    w = torch.randn(case["bm"], case["n"], case["c"], generator=torch.Generator().manual_seed(5))
    yc = y.clone().requires_grad_(True)
    (pu(yc) * w).sum().backward()
    yg = y.cuda().requires_grad_(True)
    out = unmerge(yg)
    assert out.requires_grad
    (out * w.cuda()).sum().backward()
    assert P_STABLE == P.STABLE
    same_plan = (np.array_equal(merge.src_idx.cpu().numpy(), pm.match.src_idx.numpy())
                 and np.array_equal(merge.unm_idx.cpu().numpy(), pm.match.unm_idx.numpy()))
    if same_plan:
        torch.testing.assert_close(yg.grad.cpu(), yc.grad, rtol=1e-6, atol=1e-6)
    # adjoint identity holds whatever the plan: <unmerge(y), w> == <y, merge_sum(w)>
    lhs = float((unmerge(y.cuda()) * w.cuda()).double().sum())
    rhs = float((y.cuda() * merge(w.cuda(), mode="sum")).double().sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_batch_beyond_the_grid_limit_is_a_clear_error(native):
    metric = torch.zeros(65536, 4, 8, device="cuda")
    with pytest.raises(RuntimeError, match="split the batch"):
        native.match(metric)
    with pytest.raises(RuntimeError, match="split the batch"):
        native.plan_build(metric, 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
def test_second_device_in_the_same_process(native):
    """Function attributes (> 48 KB dynamic shared memory opt-in) are per device: after running on cuda:0 the same
    process must be able to run the tcgen05 matching, the select and the fused MLP GEMM on cuda:1."""
    case = util.CASE_BY_NAME["config1_m1p"]
    metric, _, _ = util.case_arrays(case)
    plans = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            native.device_check(dev)
            plans.append(native.plan_build(torch.from_numpy(metric).cuda(dev), 100))
            a = torch.randn(4096, 768, device=f"cuda:{dev}", dtype=torch.bfloat16)
            w = torch.randn(3072, 768, device=f"cuda:{dev}", dtype=torch.bfloat16)
            native.linear_gelu(a, w, None)
            qkv = torch.randn(1, 300, 3 * 128, device=f"cuda:{dev}")                     # the attention and fp32 GEMM kernels as well
            native.attention_f32(qkv, 2, 0.125)
            native.attention_bf16(qkv.to(torch.bfloat16), 2, 0.125)
            native.frames_attention(torch.randn(1, 1 + 2 * 50, 3 * 128, device=f"cuda:{dev}").to(torch.bfloat16), 2, 2, 0.125)
            native.linear_f32(torch.randn(256, 64, device=f"cuda:{dev}"), torch.randn(256, 64, device=f"cuda:{dev}"), None)
            torch.cuda.synchronize(dev)
    assert torch.equal(plans[0].src_idx.cpu(), plans[1].src_idx.cpu())
    assert torch.equal(plans[0].dst_idx.cpu(), plans[1].dst_idx.cpu())


# ---- kernels 1 + 2 as one cluster launch (csrc/plan_cluster.cu) against the multi-launch chain -------------------------
_PC_SHAPES = [
    # bm, n, cm, heads, dtype, cls, distill, r, forced cluster size (0 = the kernel's own choice)
    (8, 1568, 64, 12, torch.bfloat16, False, False, 100, 0),      # the bench shape: VideoMAE-B layer 0
    (8, 468, 64, 12, torch.bfloat16, False, False, 100, 0),       # ... and layer 11
    (4, 1568, 64, 1, torch.float32, False, False, 100, 0),        # BASELINE config 1 (M1')
    (2, 3137, 64, 12, torch.float32, True, False, 300, 0),        # ViViT-B layer 0: class token, odd n, 16-CTA clusters
    (16, 196, 64, 12, torch.bfloat16, False, False, 18, 0),       # TimeSformer / Motionformer frames
    (3, 77, 32, 1, torch.float32, False, False, 17, 0),
    (2, 198, 64, 3, torch.float32, True, True, 30, 0),            # class + distillation token
    (2, 198, 64, 3, torch.float32, True, True, 30, 4),
    (2, 97, 16, 1, torch.float32, True, False, 24, 2),
    (2, 9, 8, 1, torch.float32, False, False, 3, 4),              # more CTAs than row blocks: some own nothing
    (2, 2, 8, 1, torch.float32, False, False, 1, 0),
    (2, 3, 8, 1, torch.bfloat16, False, False, 1, 0),
    (1, 1568, 64, 12, torch.bfloat16, False, False, 784, 8),      # r = every A token
    (5, 600, 40, 6, torch.bfloat16, False, False, 150, 0),        # cm < 64: zero-padded channels
]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", _PC_SHAPES, ids=lambda s: f"bm{s[0]}_n{s[1]}_cm{s[2]}_h{s[3]}_{'bf16' if s[4] == torch.bfloat16 else 'f32'}_cls{int(s[5])}{int(s[6])}_r{s[7]}_cs{s[8]}")
def test_plan_cluster_equals_the_multi_launch_chain(native, shape, monkeypatch):
    """tome_plan_build as ONE cluster launch gives, bit for bit, every output of the split_rows -> match_tc -> rank ->
    finish chain (itself bit-exact against the oracle): node_max / node_idx, src / unm / dst, a_map and the CSR."""
    bm, n, cm, heads, dtype, cls, dis, r, cs = shape
    gen = torch.Generator(device="cuda").manual_seed(n * 7 + cm)
    if heads > 1:
        keys = torch.randn(bm, n, 3, heads, cm, device="cuda", generator=gen).to(dtype).permute(2, 0, 3, 1, 4)[1]
        metric = native.HeadMeanMetric(keys)
    else:
        metric = torch.randn(bm, n, cm, device="cuda", generator=gen).to(dtype)
    monkeypatch.setenv("TOME_PLAN_CLUSTER", "0")
    want = native.plan_build(metric, r, cls, dis)
    monkeypatch.setenv("TOME_PLAN_CLUSTER", "1")
    if cs:
        monkeypatch.setenv("TOME_PLAN_CS", str(cs))
    geom = native.plan_cluster_describe(bm, n)
    assert geom[0] >= max(cs, 1), geom                      # the shape really takes the cluster kernel
    before = native.launch_count()
    got = native.plan_build(metric, r, cls, dis)
    assert native.launch_count() - before == 1              # one launch
    torch.cuda.synchronize()
    assert torch.equal(got.node_idx, want.node_idx)
    assert torch.equal(got.node_max.view(torch.int32), want.node_max.view(torch.int32))
    for name in ("src_idx", "unm_idx", "dst_idx", "a_map", "b_off", "b_src", "b_head"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name


@pytest.mark.timeout(180)
def test_plan_cluster_frames_view_and_zero_norm_rows(native, monkeypatch):
    """Motionformer's '(b f)' regrouping of K (tome_view with inner = frames) and a zero-norm token (NaN scores, no eps:
    merge.py:51) through the cluster kernel, against the chain."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    B, F, S, H, d = 2, 8, 196, 12, 64
    keys = torch.randn(B, H, S * F, d, device="cuda", generator=gen).to(torch.bfloat16)
    keys[0, :, 5 * F + 3] = 0          # token s=5 of frame 3: an odd token -> NaN column in that frame's matching
    keys[1, :, 4 * F + 1] = 0          # token s=4 of frame 1: an even token -> NaN row
    for frames in (F,):
        metric = native.HeadMeanMetric(keys, frames)
        monkeypatch.setenv("TOME_PLAN_CLUSTER", "0")
        want = native.plan_build(metric, 18)
        monkeypatch.setenv("TOME_PLAN_CLUSTER", "1")
        got = native.plan_build(native.HeadMeanMetric(keys, frames), 18)
        torch.cuda.synchronize()
        assert torch.equal(got.node_idx, want.node_idx)
        assert torch.equal(got.node_max.view(torch.int32), want.node_max.view(torch.int32))
        for name in ("src_idx", "unm_idx", "dst_idx", "a_map", "b_off", "b_src", "b_head"):
            assert torch.equal(getattr(got, name), getattr(want, name)), name


# ---- compact source map and Philox random scores --------------------------------------
@pytest.mark.parametrize("mode", ["merge", "hybrid", "drop"])
@pytest.mark.parametrize("flags", [(False, False), (True, False), (True, True)], ids=["plain", "cls", "cls_distill"])
def test_compact_source_map_equals_dense_source_chain(native, mode, flags):
    """Three blocks of source tracking on the compact (bm, n0) int32 form; its dense expansion must be the matrix
    the dense kernel path AND the oracle's merge_source (merge.py:372-384, pinned by the reference goldens) build."""
    import tome
    cls, dis = flags
    g = torch.Generator().manual_seed(11)
    bm, n, c = 3, 61, 16
    x = torch.randn(bm, n, c, generator=g)
    compact, dense, want = None, None, None
    xs = x.numpy()
    for layer, r in enumerate((9, 7, 30)):
        metric = torch.randn(bm, xs.shape[1], 8, generator=g)
        plan = O.bipartite_soft_matching(metric.numpy(), r, cls, dis)
        xd = torch.from_numpy(xs).cuda()
        if mode == "drop":
            op = tome.merge.bipartite_soft_matching_drop(metric.cuda(), r, cls, dis)
            compact = op.source_map(compact)
            if dense is None:
                dense = torch.eye(xs.shape[1], device="cuda")[None].expand(bm, -1, -1).contiguous()
                want = np.broadcast_to(np.eye(xs.shape[1], dtype=np.float32), (bm, xs.shape[1], xs.shape[1]))
            dense = op(dense)
            want = O.drop(plan, want)
            xs = O.drop(plan, xs)
        else:
            thr = 0.85 if mode == "hybrid" else None
            if thr is None:
                op, _ = tome.merge.bipartite_soft_matching(metric.cuda(), r, cls, dis)
            else:
                op, _ = tome.merge.bipartite_soft_matching_hybrid(metric.cuda(), r, cls, dis, threshold=thr)
            compact = tome.merge.trace_source(op, xd, compact)
            dense = tome.merge.merge_source(op, xd, dense)
            want = O.merge_source(plan, xs, want, hybrid_threshold=thr)
            xs = O.merge_wavg(plan, xs, None, hybrid_threshold=thr)[0]
        assert isinstance(compact, native.SourceMap) and compact.tokens == xs.shape[1]
        got = compact.dense().cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg=f"layer {layer}")
        np.testing.assert_array_equal(dense.cpu().numpy(), want)
        np.testing.assert_array_equal(O.source_dense(compact.group.cpu().numpy(), compact.tokens), want)
        # what tome/vis.py:55 asks of the matrix
        np.testing.assert_array_equal(compact.argmax(dim=1).cpu().numpy(), want.argmax(axis=1))
    if mode == "hybrid":
        assert (compact.group < 0).any(), "the threshold should have dropped some destinations in this case"


def test_philox_random_scores_match_oracle_and_do_not_depend_on_sharding(native):
    """tome_random_rowmax: the raw draw equals the numpy Philox4x32-10 stream bit for bit, the fused row max equals
    the masked max of the materialised draw, `call` advances on the stream, and a clip's scores do not depend on
    which batch (rank) it rides in."""
    seed, bm, na, nb = 0x1234_5678_9ABC_DEF0, 5, 37, 35
    st = native.PhiloxStream(seed, "cuda")
    nm, ni, sc = native.random_rowmax(st, bm, na, nb, True, True, want_scores=True)
    want = O.philox_scores(seed, 0, 0, bm, na, nb)
    np.testing.assert_array_equal(sc.cpu().numpy().view(np.uint32), want.view(np.uint32))
    masked = want.copy()
    masked[:, 0, :] = -np.inf
    masked[:, :, 0] = -np.inf
    onm, oni = O.rowmax(masked)
    np.testing.assert_array_equal(ni.cpu().numpy(), oni)
    np.testing.assert_array_equal(nm.cpu().numpy(), onm)
    assert st.calls() == 1
    _, _, sc1 = native.random_rowmax(st, bm, na, nb, want_scores=True)
    np.testing.assert_array_equal(sc1.cpu().numpy(), O.philox_scores(seed, 1, 0, bm, na, nb))
    # a rank holding clips 3..4 of the global batch draws what a single GPU holding all five draws for them
    shard = native.PhiloxStream(seed, "cuda", clip_offset=3)
    _, _, sc_shard = native.random_rowmax(shard, 2, na, nb, want_scores=True)
    np.testing.assert_array_equal(sc_shard.cpu().numpy(), want[3:5])


def test_random_modes_on_the_philox_stream(native):
    import tome
    x = torch.randn(4, 90, 8, device="cuda")
    try:
        tome.merge.philox_seed(99)
        merge, _ = tome.merge.bipartite_soft_matching(x, 11, mode="random_merge")
        scores = O.philox_scores(99, 0, 0, 4, 45, 45)
        plan = O.bipartite_soft_matching(x.cpu().numpy(), 11, given_scores=scores)
        np.testing.assert_array_equal(merge.plan.src_idx.cpu().numpy(), plan.src_idx)
        np.testing.assert_array_equal(merge.plan.dst_idx.cpu().numpy(), plan.dst_idx)
        drop = tome.merge.bipartite_soft_matching_drop(x, 11, mode="random_drop")          # second call of the stream
        plan2 = O.bipartite_soft_matching(x.cpu().numpy(), 11, given_scores=O.philox_scores(99, 1, 0, 4, 45, 45))
        np.testing.assert_array_equal(drop(x).cpu().numpy(), O.drop(plan2, x.cpu().numpy()))
        # re-seeding replays the stream
        tome.merge.philox_seed(99)
        again, _ = tome.merge.bipartite_soft_matching(x, 11, mode="random_merge")
        assert torch.equal(again.plan.src_idx, merge.plan.src_idx)
    finally:
        tome.merge.philox_seed(None)


# ---- the callers either side of the path in the divided space-time blocks -------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_rows_add_layernorm_through_views(native, dtype):
    """tome_rows_add_layernorm on the three row orders of a TimeSformer block (timesformer.py:38-56): same bits as the
    contiguous tome_add_layernorm on materialised copies, and torch's add + layer_norm within rounding."""
    g = torch.Generator().manual_seed(5)
    B, P, T, C = 2, 7, 4, 64
    x = torch.randn(B, 1 + P * T, C, generator=g).to("cuda", dtype)
    tf = torch.randn(B, P, T, C, generator=g).to("cuda", dtype)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    b = (0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    norm = (w, b, 1e-6)
    x4 = x[:, 1:].unflatten(1, (P, T))
    # (1) LayerNorm only, '(b p) t' output
    nt = torch.empty(B, P, T, C, device="cuda", dtype=dtype)
    native.rows_add_layernorm(x4, None, norm, None, nt)
    want = torch.nn.functional.layer_norm(x[:, 1:].float(), (C,), w.float(), b.float(), 1e-6).reshape(B, P, T, C)
    torch.testing.assert_close(nt.float(), want, rtol=2e-2 if dtype == torch.bfloat16 else 1e-5, atol=2e-2 if dtype == torch.bfloat16 else 1e-5)
    # (2) add, sum into the residual-stream layout, LayerNorm into '(b t) (1 + p)' rows
    xfull = torch.zeros_like(x)
    ns = torch.zeros(B * T, 1 + P, C, device="cuda", dtype=dtype)
    native.rows_add_layernorm(x4, tf, norm, xfull[:, 1:].unflatten(1, (P, T)), ns.view(B, T, 1 + P, C)[:, :, 1:].permute(0, 2, 1, 3))
    s_ref, n_ref = native.add_layernorm(x[:, 1:].contiguous(), tf.reshape(B, P * T, C), norm)          # the contiguous kernel
    assert torch.equal(xfull[:, 1:], s_ref)
    assert torch.equal(xfull[:, 0], torch.zeros_like(xfull[:, 0]))                                        # class row untouched
    got = ns.view(B, T, 1 + P, C)[:, :, 1:].permute(0, 2, 1, 3).reshape(B, P * T, C)
    assert torch.equal(got, n_ref)
    assert torch.equal(ns[:, 0], torch.zeros_like(ns[:, 0]))
    torch.testing.assert_close(s_ref.float(), (x[:, 1:].float() + tf.reshape(B, P * T, C).float()).to(dtype).float())
    # (3) sum only
    only = torch.empty(B, P, T, C, device="cuda", dtype=dtype)
    native.rows_add_layernorm(x4, tf, None, only, None)
    assert torch.equal(only.reshape(B, P * T, C), s_ref)


@pytest.mark.parametrize("C", [64, 96, 768])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_cls_rows_match_the_separate_torch_ops(native, dtype, C):
    """tome_cls_rows on the three class-token steps of a divided space-time block (timesformer.py:41-48, 56) against the torch
    ops it replaces: the adds bit for bit (each step rounds to the dtype), the frame mean and LayerNorm within rounding."""
    g = torch.Generator().manual_seed(13)
    B, T, P = 3, 8, 5
    x = torch.randn(B, 1 + P * T, C, generator=g).to("cuda", dtype)
    y = torch.randn(B, 1 + P * T, C, generator=g).to("cuda", dtype)
    res_s = torch.randn(B * T, 1 + P, C, generator=g).to("cuda", dtype)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    b = (0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    bf = dtype == torch.bfloat16
    # (1) LayerNorm of the class row replicated into the '(b t) (1 + p)' buffer's class rows
    ns = torch.zeros(B * T, 1 + P, C, device="cuda", dtype=dtype)
    native.cls_rows(x[:, 0], norm=(w, b, 1e-6), normed_out=ns.view(B, T, 1 + P, C)[:, :, 0])
    want = torch.nn.functional.layer_norm(x[:, 0].float(), (C,), w.float(), b.float(), 1e-6)
    got = ns.view(B, T, 1 + P, C)[:, :, 0]
    torch.testing.assert_close(got.float(), want[:, None].expand(B, T, C), rtol=2e-2 if bf else 1e-5, atol=2e-2 if bf else 1e-5)
    assert torch.equal(got[:, 0], got[:, T - 1])
    assert torch.equal(ns[:, 1:], torch.zeros_like(ns[:, 1:]))                                           # patch rows untouched
    # (2) class row + mean over the frames of the spatial attention's class outputs
    cls = torch.empty(B, C, device="cuda", dtype=dtype)
    m = res_s.view(B, T, 1 + P, C)[:, :, 0]
    native.cls_rows(x[:, 0], mean_src=m, sum_out=cls)
    want = x[:, 0] + m.mean(1)
    if bf:
        assert torch.equal(cls, want)                                 # fp32 accumulation, one rounding: no order dependence left
    else:
        torch.testing.assert_close(cls, want, rtol=1e-6, atol=1e-6)
    # (3) the closing residual add of the class row, written into the next residual stream
    s = torch.zeros_like(x)
    native.cls_rows(x[:, 0], add=y[:, 0], sum_out=s[:, 0])
    assert torch.equal(s[:, 0], x[:, 0] + y[:, 0])
    assert torch.equal(s[:, 1:], torch.zeros_like(s[:, 1:]))
    # all three at once
    both = torch.empty(B, C, device="cuda", dtype=dtype)
    nb = torch.empty(B, 2, C, device="cuda", dtype=dtype)
    native.cls_rows(x[:, 0], add=y[:, 0], mean_src=m, sum_out=both, norm=(w, b, 1e-6), normed_out=nb)
    want = (x[:, 0] + m.mean(1)) + y[:, 0]
    torch.testing.assert_close(both, want, rtol=0 if bf else 1e-6, atol=0 if bf else 1e-6)
    torch.testing.assert_close(nb[:, 1].float(), torch.nn.functional.layer_norm(both.float(), (C,), w.float(), b.float(), 1e-6),
                               rtol=2e-2 if bf else 1e-5, atol=2e-2 if bf else 1e-5)


@pytest.mark.parametrize("mode", ["merge", "hybrid", "drop"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_merge_frames_with_the_spatial_residual_in_its_own_layout(native, mode, dtype):
    """merge on 'b (1 + p t)' tokens with the spatial attention's '(b t) (1 + p)' output added inside the kernel
    (tome_merge_add_norm_rv) == the reference's rearrange + cat + add (timesformer.py:46-48) followed by the merge."""
    import tome
    g = torch.Generator().manual_seed(9)
    B, T, P, C = 2, 4, 22, 64
    x = torch.randn(B, 1 + P * T, C, generator=g).to("cuda", dtype)
    res_s = torch.randn(B * T, 1 + P, C, generator=g).to("cuda", dtype)
    cls = torch.randn(B, C, generator=g).to("cuda", dtype)
    metric = torch.randn(B * T, P, 16, generator=g).cuda()
    size = torch.randint(1, 4, (B * T, P, 1), generator=g).float().cuda()
    w = (1 + 0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    bb = (0.1 * torch.randn(C, generator=g)).to("cuda", dtype)
    full = torch.empty_like(x)
    full[:, 0] = cls
    full[:, 1:] = x[:, 1:] + res_s[:, 1:].reshape(B, T, P, C).transpose(1, 2).reshape(B, P * T, C)
    if mode == "drop":
        op = tome.merge.bipartite_soft_matching_drop(metric, 6)
        got = op.frames(x, T, residual=res_s, cls=cls)
        want = op.frames(full, T)
        assert torch.equal(got, want)
        return
    if mode == "hybrid":
        op, _ = tome.merge.bipartite_soft_matching_hybrid(metric, 6, threshold=0.7)
    else:
        op, _ = tome.merge.bipartite_soft_matching(metric, 6)
    got = op.wavg_frames(x, T, size, norm=(w, bb, 1e-6), residual=res_s, cls=cls)
    want = op.wavg_frames(full, T, size, norm=(w, bb, 1e-6))
    for a, b_ in zip(got, want):
        assert torch.equal(a, b_)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("tn", [8, 5, 32])
def test_attn_short_matches_softmax_attention(native, dtype, tn):
    """tome_attn_short (TimeSformer's temporal attention, slowfast/models/timesformer.py:76-84 on '(b p) t' rows) against
    the eager softmax formulation in fp64."""
    g = torch.Generator().manual_seed(tn)
    seqs, H, d = 37, 3, 64
    qkv = torch.randn(seqs, tn, 3 * H * d, generator=g).to("cuda", dtype)
    out = native.attn_short(qkv, H, d ** -0.5)
    q, k, v = qkv.double().reshape(seqs, tn, 3, H, d).permute(2, 0, 3, 1, 4)
    want = ((q @ k.transpose(-1, -2)) * d ** -0.5).softmax(-1) @ v
    want = want.transpose(1, 2).reshape(seqs, tn, H * d)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-6
    torch.testing.assert_close(out.double(), want, rtol=tol, atol=tol)
    with torch.no_grad():
        assert native.attn_short_usable(qkv[..., :H * d], H) and not native.attn_short_usable(torch.zeros(2, 33, H * d, device="cuda"), H)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("with_bias", [False, True], ids=["plain", "prop_attn"])
@pytest.mark.parametrize("P", [196, 178, 128, 37, 9])
def test_frames_attention_matches_per_frame_softmax(native, P, with_bias):
    """tome_frames_attention (tcgen05) against the eager formulation of Motionformer's space attention with the
    proportional-attention key bias (tome/patch/motionformer.py:105-115) in fp64; bf16 inputs, probabilities rounded to
    bf16 before P V as every fused attention kernel does."""
    g = torch.Generator().manual_seed(P)
    B, h, Fr, d = 2, 3, 4, 64
    S, C = Fr * P, h * d
    qkv = torch.randn(B, 1 + S, 3 * C, generator=g).to("cuda", torch.bfloat16)
    bias = (torch.randint(1, 6, (B, S), generator=g).float().log()).cuda() if with_bias else None
    with torch.no_grad():
        assert native.frames_attention_usable(qkv[..., :C], h, P)
    xs, diag = native.frames_attention(qkv, h, Fr, d ** -0.5, bias)
    q, k, v = qkv.double().reshape(B, 1 + S, 3, h, d).permute(2, 0, 3, 1, 4)          # (B, h, N, d)
    q, k, v = q[:, :, 1:], k[:, :, 1:].reshape(B, h, Fr, P, d), v[:, :, 1:].reshape(B, h, Fr, P, d)
    sc = torch.einsum("bhsd,bhfpd->bhsfp", q, k) * d ** -0.5
    if bias is not None:
        sc = sc + bias.double().reshape(B, 1, 1, Fr, P)
    want = torch.einsum("bhsfp,bhfpd->bsfhd", sc.softmax(-1), v).reshape(B, S, Fr, C)
    torch.testing.assert_close(xs.double(), want, rtol=2e-2, atol=6e-3)
    frame = torch.arange(S, device="cuda") // P
    assert torch.equal(diag, xs[:, torch.arange(S, device="cuda"), frame])


@pytest.mark.timeout(120)
@pytest.mark.parametrize("with_bias", [False, True], ids=["plain", "prop_attn"])
@pytest.mark.parametrize("P", [196, 178, 64, 37, 9])
def test_frames_attention_f32_matches_per_frame_softmax(native, P, with_bias):
    """tome_frames_attention_f32 (the exact-split attention kernel, one problem per frame) against the eager formulation of
    Motionformer's space attention (tome/patch/motionformer.py:105-115) in fp64: fp32-class accuracy; the planes output is
    the exact split of xs, x_diag its own-frame slice."""
    g = torch.Generator().manual_seed(P)
    B, h, Fr, d = 2, 3, 4, 64
    S, C = Fr * P, h * d
    qkv = torch.randn(B, 1 + S, 3 * C, generator=g).cuda()
    bias = (torch.randint(1, 6, (B, S), generator=g).float().log()).cuda() if with_bias else None
    with torch.no_grad():
        assert native.frames_attention_f32_usable(qkv[..., :C], h)
        xs, xs3, diag = native.frames_attention_f32(qkv, h, Fr, d ** -0.5, bias)
    q, k, v = qkv.double().reshape(B, 1 + S, 3, h, d).permute(2, 0, 3, 1, 4)          # (B, h, N, d)
    q, k, v = q[:, :, 1:], k[:, :, 1:].reshape(B, h, Fr, P, d), v[:, :, 1:].reshape(B, h, Fr, P, d)
    sc = torch.einsum("bhsd,bhfpd->bhsfp", q, k) * d ** -0.5
    if bias is not None:
        sc = sc + bias.double().reshape(B, 1, 1, Fr, P)
    want = torch.einsum("bhsfp,bhfpd->bsfhd", sc.softmax(-1), v).reshape(B, S, Fr, C)
    err = (xs.double() - want).abs().max().item() / want.abs().max().item()
    print(f"[frames_attention_f32] P={P} bias={with_bias}: max err / max|y| = {err:.2e}")
    assert err <= 2e-6, err
    frame = torch.arange(S, device="cuda") // P
    assert torch.equal(diag, xs[:, torch.arange(S, device="cuda"), frame])
    assert torch.equal(xs3.float(), xs)                                               # h + m + l == xs, bit for bit


def test_frames_attention_f32_with_one_frame_is_plain_attention(native):
    """Plain attention is the one-frame, no-lead case of the frames kernel (same template, FRAMES = false vs true paths): the
    two entry points must agree bit for bit, with and without the key bias."""
    g = torch.Generator().manual_seed(77)
    B, h, d, N = 2, 3, 64, 333
    qkv = torch.randn(B, N, 3 * h * d, generator=g).cuda()
    bias = torch.rand(B, N, generator=g).cuda()
    with torch.no_grad():
        for kb in (None, bias):
            want = native.attention_f32(qkv, h, d ** -0.5, kb)
            xs, xs3, diag = native.frames_attention_f32(qkv, h, 1, d ** -0.5, kb, lead=0)
            assert torch.equal(xs.view(B, N, h * d), want)
            assert torch.equal(diag, want)                    # every query's own frame is frame 0
            assert torch.equal(xs3.float().view(B, N, h * d), want)


def test_traj_temporal_fp32_matches_einsum_formulation(native):
    """fp32 tome_traj_temporal against vit_helper.py:232-243 in fp64."""
    g = torch.Generator().manual_seed(4)
    B, S, Fr, h, d = 2, 77, 8, 3, 64
    C = h * d
    q2 = torch.randn(B, S, C, generator=g).cuda()
    k2 = torch.randn(B, S, Fr, C, generator=g).cuda()
    xs = torch.randn(B, S, Fr, C, generator=g).cuda()
    out = native.traj_temporal(q2, k2, xs, h, d ** -0.5)
    assert out.dtype == torch.float32
    qd = q2.double().reshape(B, S, h, d).transpose(1, 2) * d ** -0.5
    kd = k2.double().reshape(B, S, Fr, h, d).permute(0, 3, 1, 2, 4)
    vd = xs.double().reshape(B, S, Fr, h, d).permute(0, 3, 1, 2, 4)
    attn = torch.einsum("bhsd,bhsfd->bhsf", qd, kd).softmax(-1)
    want = torch.einsum("bhsf,bhsfd->bhsd", attn, vd).transpose(1, 2).reshape(B, S, C)
    torch.testing.assert_close(out.double(), want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("N", [1569, 197, 41])
def test_cls_attention_matches_single_query_softmax(native, N, dtype):
    """tome_cls_attention (Motionformer's class-token row, vit_helper.py:181-189) against fp64."""
    g = torch.Generator().manual_seed(N)
    B, h, d = 2, 3, 64
    C = h * d
    qkv = torch.randn(B, N, 3 * C, generator=g).to("cuda", dtype)
    out = native.cls_attention(qkv, h, d ** -0.5)
    q, k, v = qkv.double().reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
    want = (((q[:, :, 0:1] @ k.transpose(-1, -2)) * d ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B, 1, C)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-6
    assert out.shape == (B, 1, C) and out.dtype == dtype
    torch.testing.assert_close(out.double(), want, rtol=tol, atol=tol)


def test_traj_temporal_matches_einsum_formulation(native):
    """tome_traj_temporal against vit_helper.py:232-243 (softmax over frames of q2 . k2, values = xs) in fp64."""
    g = torch.Generator().manual_seed(3)
    B, S, Fr, h, d = 2, 77, 8, 3, 64
    C = h * d
    q2 = torch.randn(B, S, C, generator=g).to("cuda", torch.bfloat16)
    k2 = torch.randn(B, S, Fr, C, generator=g).to("cuda", torch.bfloat16)
    xs = torch.randn(B, S, Fr, C, generator=g).to("cuda", torch.bfloat16)
    out = native.traj_temporal(q2, k2, xs, h, d ** -0.5)
    qd = q2.double().reshape(B, S, h, d).transpose(1, 2) * d ** -0.5
    kd = k2.double().reshape(B, S, Fr, h, d).permute(0, 3, 1, 2, 4)
    vd = xs.double().reshape(B, S, Fr, h, d).permute(0, 3, 1, 2, 4)
    attn = torch.einsum("bhsd,bhsfd->bhsf", qd, kd).softmax(-1)
    want = torch.einsum("bhsf,bhsfd->bhsd", attn, vd).transpose(1, 2).reshape(B, S, C)
    torch.testing.assert_close(out.double(), want, rtol=1e-2, atol=1e-2)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shape", [(300, 256, 64), (1000, 768, 768), (2100, 3072, 768), (777, 768, 3072)], ids=str)
@pytest.mark.parametrize("terms", [9, 8, 6])
def test_linear_f32_matches_fp64(native, shape, terms):
    """tome_linear_f32: fp32 GEMM on tcgen05 through the exact three-way bf16 split (nine products, fp32 accumulation).
    Against fp64 it must be in the accuracy class of torch's fp32 GEMM with TF32 off (the reference's arithmetic,
    slowfast/utils/model_benchmark.py:21-45): within 4x the library's own error on the same operands (measured 1-2.2x: the
    tensor core's fp32 accumulator rounds differently from an FMA chain) and below 1.5e-6 of max |y| -- TF32 would be
    5e-4 -- and the split itself is exact."""
    m, n, k = shape
    g = torch.Generator().manual_seed(m + n + k)
    x = (torch.randn(m, k, generator=g) * torch.logspace(-2, 2, k)).cuda()          # a wide dynamic range across channels
    w = (torch.randn(n, k, generator=g) * 0.05).cuda()
    b = torch.randn(n, generator=g).cuda()
    x3 = native.split3(x).float()
    assert torch.equal(x3[:, :k] + x3[:, k:2 * k] + x3[:, 2 * k:], x)              # h + m + l == x, bit for bit
    assert not torch.backends.cuda.matmul.allow_tf32
    want = torch.nn.functional.linear(x.double(), w.double(), b.double())
    lib_out = torch.nn.functional.linear(x, w, b)
    scale = want.abs().max().item()
    lib_err = (lib_out.double() - want).abs().max().item() / scale
    with torch.no_grad():
        assert native.linear_f32_usable(x, w, b)
        out = native.linear_f32(x, w, b, terms=terms)
        err = (out.double() - want).abs().max().item() / scale
        print(f"[linear_f32] {shape} terms={terms}: max err / max|y| = {err:.2e} (library fp32 GEMM: {lib_err:.2e})")
        assert err <= min(max(4.0 * lib_err, 3e-7), 1.5e-6), (err, lib_err)
        act = native.linear_f32(x, w, b, gelu=True, terms=terms)
        torch.testing.assert_close(act, torch.nn.functional.gelu(out), rtol=1e-6, atol=1e-6)
        nob = native.linear_f32(x, w, None, terms=terms)
        torch.testing.assert_close(nob + b, out, rtol=1e-6, atol=1e-5)
        from hostmodels.vivit import gelu_fast                   # HF FastGELUActivation, ViViT's hidden_act, in the epilogue
        torch.testing.assert_close(native.linear_f32(x, w, b, gelu="gelu_fast", terms=terms), gelu_fast(out), rtol=2e-6, atol=2e-6)


def test_planes_round_trip_is_exact(native):
    """tome_split3 -> tome_planes_sum returns the fp32 tensor bit for bit (h + m + l == x), on a column slice too; the planes
    output of tome_linear_f32 sums to its fp32 output."""
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(515, 2304, generator=g) * torch.logspace(-6, 6, 2304)).cuda()
    p3 = native.Planes(native.split3(x), (515, 2304))
    assert torch.equal(native.planes_to_f32(p3), x)
    assert torch.equal(native.planes_to_f32(p3, 768, 768), x[:, 768:1536])
    w = (torch.randn(768, 2304, generator=g) * 0.02).cuda()
    with torch.no_grad():
        y, y3 = native.linear_f32(p3, w, None, out="both")
    assert torch.equal(native.planes_to_f32(y3), y)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        native.planes_to_f32(p3, 4, 768)


@pytest.mark.parametrize("shape", [(1000, 768, 768), (2100, 3072, 768), (777, 768, 3072)], ids=str)
def test_linear_f32_eight_products_equal_nine_at_fp32_resolution(native, shape):
    """The default drops ONE of the nine plane products, l.l (<= 2^-32 of |a||b| per product: 2^-8 of the fp32 accumulator's own
    rounding unit).  Against the nine-product result the outputs must agree to within one fp32 epsilon of the output scale,
    and against fp64 the two must be equally far away."""
    m, n, k = shape
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) * k ** -0.5).cuda()
    b = torch.randn(n, generator=g).cuda()
    want = x.double() @ w.double().t() + b.double()
    scale = want.abs().max().item()
    with torch.no_grad():
        y9 = native.linear_f32(x, w, b, terms=9)
        y8 = native.linear_f32(x, w, b, terms=8)
    diff = (y8.double() - y9.double()).abs().max().item() / scale
    e9 = (y9.double() - want).abs().max().item() / scale
    e8 = (y8.double() - want).abs().max().item() / scale
    print(f"[linear_f32] {shape}: |x8 - x9| / max|y| = {diff:.2e} (2^-24 = 5.96e-08); error vs fp64: x9 {e9:.3e}, x8 {e8:.3e}")
    assert diff <= 2.0 ** -24, diff                      # below one fp32 epsilon of the output scale (measured 1.6e-8 .. 3.4e-8): single flipped roundings
    assert e8 <= e9 * 1.02 + 1e-9


@pytest.mark.timeout(180)
@pytest.mark.parametrize("with_bias", [False, True], ids=["plain", "prop_attn"])
@pytest.mark.parametrize("N", [3137, 1568, 470, 129, 128, 65])
def test_attention_bf16_matches_fp64(native, N, with_bias):
    """tome_attention_bf16 (tcgen05 flash attention with the key bias, q / k / v in place from the QKV output) against fp64
    softmax attention on the same bf16 operands: the bf16 bar of the north star (1e-2), and no worse than 2x the library's
    bf16 attention with the reference's masked formulation (tome/patch/videomae.py:58-68)."""
    g = torch.Generator().manual_seed(N)
    B, h, d = 2, 3, 64
    C = h * d
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 1.5).to(torch.bfloat16).cuda()
    bias = (torch.randint(1, 8, (B, N), generator=g).float().log()).cuda() if with_bias else None
    q, k, v = qkv.double().reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * d ** -0.5
    if bias is not None:
        sc = sc + bias.double()[:, None, None, :]
    want = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    qf, kf, vf = qkv.reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
    mask = None if bias is None else bias.to(torch.bfloat16)[:, None, None, :].expand(B, 1, N, N)
    lib = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask, scale=d ** -0.5).transpose(1, 2).reshape(B, N, C)
    scale = want.abs().max().item()
    lib_err = (lib.double() - want).abs().max().item() / scale
    with torch.no_grad():
        assert native.attention_bf16_usable(qkv, h, bias)
        out = native.attention_bf16(qkv, h, d ** -0.5, bias)
    assert out.dtype == torch.bfloat16 and out.shape == (B, N, C)
    err = (out.double() - want).abs().max().item() / scale
    print(f"[attention_bf16] N={N} bias={with_bias}: max err / max|y| = {err:.2e} (library bf16 attention: {lib_err:.2e})")
    assert err <= 1e-2 and err <= max(2.0 * lib_err, 4e-3), (err, lib_err)
    if with_bias:                                             # unbiased leading query (TimeSformer's class token)
        with torch.no_grad():
            out1 = native.attention_bf16(qkv, h, d ** -0.5, bias, unbiased_queries=1)
        sc0 = (q @ k.transpose(-1, -2)) * d ** -0.5
        want0 = (sc0.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
        torch.testing.assert_close(out1[:, 0].double(), want0[:, 0], rtol=2e-2, atol=2e-2)
        assert torch.equal(out1[:, 1:], out[:, 1:])


@pytest.mark.timeout(180)
@pytest.mark.parametrize("with_bias", [False, True], ids=["plain", "prop_attn"])
@pytest.mark.parametrize("N", [1568, 468, 196, 100])
def test_attention_f32_matches_fp64(native, N, with_bias):
    """tome_attention_f32 (exact-split tcgen05 flash attention) against fp64 softmax attention; its error must be in the class of
    torch's own fp32 attention on the same operands (tome/patch/videomae.py:58-68 in the fp32 models)."""
    g = torch.Generator().manual_seed(N)
    B, h, d = 2, 3, 64
    C = h * d
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 2.0).cuda()
    bias = (torch.randint(1, 8, (B, N), generator=g).float().log()).cuda() if with_bias else None
    q, k, v = qkv.double().reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * d ** -0.5
    if bias is not None:
        sc = sc + bias.double()[:, None, None, :]
    want = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    qf, kf, vf = qkv.reshape(B, N, 3, h, d).permute(2, 0, 3, 1, 4)
    mask = None if bias is None else bias[:, None, None, :].expand(B, 1, N, N)
    lib = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask, scale=d ** -0.5).transpose(1, 2).reshape(B, N, C)
    scale = want.abs().max().item()
    lib_err = (lib.double() - want).abs().max().item() / scale
    with torch.no_grad():
        assert native.attention_f32_usable(qkv, h, bias)
        out = native.attention_f32(qkv, h, d ** -0.5, bias)
    err = (out.double() - want).abs().max().item() / scale
    print(f"[attention_f32] N={N} bias={with_bias}: max err / max|y| = {err:.2e} (torch fp32 attention: {lib_err:.2e})")
    assert err <= max(4.0 * lib_err, 1e-6), (err, lib_err)
    if with_bias:                                             # unbiased leading query (TimeSformer's class token)
        with torch.no_grad():
            out1 = native.attention_f32(qkv, h, d ** -0.5, bias, unbiased_queries=1)
        sc0 = (q @ k.transpose(-1, -2)) * d ** -0.5
        want0 = (sc0.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
        torch.testing.assert_close(out1[:, 0].double(), want0[:, 0], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out1[:, 1:], out[:, 1:])
