"""ctypes binding of libtome_b200.so (the C ABI in include/tome_b200.h).

PyTorch is only the plumbing here: it owns device memory and the stream; every kernel on the
token-merging path is ours.  There is NO CPU fallback and no alternative backend: if the
library is missing, the device is not sm_100, or a tensor is not on a CUDA device, these
functions raise.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG, "lib", "libtome_b200.so")
ABI_VERSION = 26

TOME_F32, TOME_BF16, TOME_U8 = 0, 1, 2
MATCH_AUTO, MATCH_EXACT_SIMT, MATCH_TCGEN05 = 0, 1, 2
MODE_WAVG, MODE_SUM, MODE_MEAN, MODE_AMAX, MODE_DROP = 0, 1, 2, 3, 4
_MODES = {"wavg": MODE_WAVG, "sum": MODE_SUM, "mean": MODE_MEAN, "max": MODE_AMAX, "amax": MODE_AMAX,
          "drop": MODE_DROP}

EXPORTS = (
    "tome_abi_version", "tome_last_error", "tome_launch_count", "tome_device_check", "tome_match_workspace_bytes", "tome_match", "tome_match_heads",
    "tome_plan_build_workspace_bytes", "tome_plan_build", "tome_match_tc_describe", "tome_plan_cluster_describe",
    "tome_rowmax", "tome_select_workspace_bytes", "tome_select", "tome_merge", "tome_merge_norm", "tome_merge_add_norm", "tome_add_layernorm", "tome_add_rows_layernorm",
    "tome_merge_source", "tome_attn_key_bias", "tome_patchify", "tome_linear_gelu", "tome_unmerge",
    "tome_match_sets_workspace_bytes", "tome_match_sets", "tome_group_reduce", "tome_gather_rows",
    "tome_source_compose", "tome_source_dense", "tome_random_rowmax", "tome_merge_add_norm_rv", "tome_rows_add_layernorm", "tome_attn_short",
    "tome_frames_attention", "tome_traj_temporal", "tome_split3", "tome_linear_f32", "tome_attention_f32", "tome_cls_rows", "tome_attention_bf16", "tome_frames_attention_f32", "tome_cls_attention", "tome_planes_sum",
)


class TomePlanC(ctypes.Structure):
    _fields_ = [
        ("bm", ctypes.c_int32), ("n", ctypes.c_int32), ("r", ctypes.c_int32),
        ("class_token", ctypes.c_int32), ("distill_token", ctypes.c_int32),
        ("node_max", ctypes.c_void_p), ("node_idx", ctypes.c_void_p),
        ("src_idx", ctypes.c_void_p), ("unm_idx", ctypes.c_void_p), ("dst_idx", ctypes.c_void_p),
        ("a_map", ctypes.c_void_p), ("b_off", ctypes.c_void_p), ("b_src", ctypes.c_void_p),
        ("b_head", ctypes.c_void_p),
    ]


class TomeViewC(ctypes.Structure):
    _fields_ = [("stride_bo", ctypes.c_int64), ("stride_bi", ctypes.c_int64), ("stride_n", ctypes.c_int64),
                ("inner", ctypes.c_int32)]


_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it is absent: build it with
    ``python video-how-do-your-tokens-merge_b200/build.py`` or ``__graft_entry__.build()``."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("TOME_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"tome_b200: CUDA extension not found at {p}; there is no CPU fallback. "
            "Build it with `python video-how-do-your-tokens-merge_b200/build.py`.")
    lib = ctypes.CDLL(p)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f"tome_b200: {p} does not export {name}")
    c_i32, c_f32, c_vp, c_sz = ctypes.c_int32, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
    lib.tome_abi_version.restype = c_i32
    lib.tome_last_error.restype = ctypes.c_char_p
    lib.tome_launch_count.restype = ctypes.c_ulonglong
    lib.tome_device_check.argtypes = [c_i32]
    lib.tome_match_workspace_bytes.restype = c_sz
    lib.tome_match_workspace_bytes.argtypes = [c_i32, c_i32, c_i32, c_i32]
    lib.tome_match.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_i32, c_i32, c_i32,
                               c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tome_match_heads.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(TomeViewC), ctypes.c_int64,
                                     c_i32, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tome_rowmax.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]
    lib.tome_select_workspace_bytes.restype = c_sz
    lib.tome_select_workspace_bytes.argtypes = [c_i32, c_i32]
    lib.tome_select.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_sz, c_vp]
    lib.tome_merge.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_vp, c_i32,
                               c_f32, c_vp, ctypes.POINTER(TomeViewC), c_vp, c_vp, c_vp]
    lib.tome_merge_norm.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_vp, c_i32,
                                    c_f32, c_vp, ctypes.POINTER(TomeViewC), c_vp, c_vp, c_vp, c_vp, c_f32, c_vp,
                                    ctypes.POINTER(TomeViewC), c_vp]
    lib.tome_plan_build_workspace_bytes.restype = ctypes.c_size_t
    lib.tome_plan_build_workspace_bytes.argtypes = [c_i32, c_i32, c_i32]
    lib.tome_plan_build.argtypes = [c_vp, c_i32, c_i32, ctypes.c_int64, ctypes.POINTER(TomeViewC), c_i32, c_i32,
                                    ctypes.POINTER(TomePlanC), c_vp, ctypes.c_size_t, c_vp]
    lib.tome_plan_build.restype = c_i32
    lib.tome_plan_cluster_describe.restype = None
    lib.tome_plan_cluster_describe.argtypes = [c_i32, c_i32, ctypes.POINTER(ctypes.c_int64)]
    lib.tome_match_tc_describe.restype = None
    lib.tome_match_tc_describe.argtypes = [c_i32, c_i32, c_i32, ctypes.POINTER(ctypes.c_int64)]
    lib.tome_merge_add_norm.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_vp, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_vp,
                                        c_i32, c_f32, c_vp, ctypes.POINTER(TomeViewC), c_vp, c_vp, c_vp, c_vp, c_f32, c_vp,
                                        ctypes.POINTER(TomeViewC), c_vp]
    lib.tome_add_layernorm.argtypes = [c_vp, c_vp, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp]
    lib.tome_add_rows_layernorm.argtypes = [c_vp, c_vp, ctypes.c_int64, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp, c_f32, c_vp,
                                            c_vp, c_vp]
    lib.tome_patchify.argtypes = [c_vp, c_i32] + [c_i32] * 8 + [c_vp, c_i32, c_vp]
    lib.tome_linear_gelu.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp]
    lib.tome_merge_source.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_i32, c_f32, c_vp, c_vp]
    c_i64 = ctypes.c_int64
    lib.tome_attn_key_bias.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp, c_i64, c_i64, c_i64,
                                       c_vp, c_i64, c_i64, c_i64, c_vp]
    lib.tome_unmerge.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_i32, c_i32, c_vp, c_vp]
    lib.tome_match_sets_workspace_bytes.restype = c_sz
    lib.tome_match_sets_workspace_bytes.argtypes = [c_i32, c_i32, c_i32, c_i32]
    lib.tome_match_sets.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_vp, c_i64, c_i32, c_i32,
                                    c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tome_group_reduce.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(TomeViewC), c_vp, c_i64, c_i32, c_i32,
                                      c_vp, c_i32, c_vp, c_vp]
    lib.tome_gather_rows.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp]
    lib.tome_source_compose.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_i32, c_i32, c_f32, c_vp, c_vp]
    lib.tome_source_dense.argtypes = [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]
    lib.tome_random_rowmax.argtypes = [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]
    lib.tome_merge_add_norm_rv.argtypes = [ctypes.POINTER(TomePlanC), c_vp, c_vp, ctypes.POINTER(TomeViewC), c_i32, c_i32,
                                           ctypes.POINTER(TomeViewC), c_vp, c_i32, c_f32, c_vp, ctypes.POINTER(TomeViewC), c_vp, c_vp,
                                           c_vp, c_vp, c_f32, c_vp, ctypes.POINTER(TomeViewC), c_vp]
    p_i64 = ctypes.POINTER(ctypes.c_int64)
    lib.tome_rows_add_layernorm.argtypes = [c_vp, p_i64, c_vp, p_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, p_i64,
                                            c_vp, p_i64, c_vp]
    lib.tome_attn_short.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i64, c_i64, c_f32, c_vp, c_vp]
    lib.tome_split3.argtypes = [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp]
    lib.tome_linear_f32.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]
    lib.tome_attention_f32.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_i32, c_vp, c_vp, c_vp]
    lib.tome_attention_bf16.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_i32, c_vp, c_vp]
    lib.tome_frames_attention_f32.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tome_planes_sum.argtypes = [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp]
    lib.tome_planes_sum.restype = c_i32
    lib.tome_cls_attention.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp]
    for name in ("tome_split3", "tome_linear_f32", "tome_attention_f32", "tome_attention_bf16", "tome_frames_attention_f32", "tome_cls_attention"):
        getattr(lib, name).restype = c_i32
    lib.tome_frames_attention.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp]
    lib.tome_traj_temporal.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp]
    lib.tome_cls_rows.argtypes = [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_f32,
                                  c_vp, c_i64, c_i64, c_i32, c_vp]
    for name in ("tome_source_compose", "tome_source_dense", "tome_random_rowmax", "tome_merge_add_norm_rv", "tome_rows_add_layernorm",
                 "tome_attn_short", "tome_frames_attention", "tome_traj_temporal", "tome_cls_rows"):
        getattr(lib, name).restype = c_i32
    for name in ("tome_device_check", "tome_match", "tome_match_heads", "tome_rowmax", "tome_select", "tome_merge", "tome_merge_norm", "tome_merge_add_norm", "tome_add_layernorm", "tome_add_rows_layernorm",
                 "tome_merge_source", "tome_attn_key_bias", "tome_patchify", "tome_linear_gelu", "tome_unmerge",
                 "tome_match_sets", "tome_group_reduce", "tome_gather_rows"):
        getattr(lib, name).restype = c_i32
    if lib.tome_abi_version() != ABI_VERSION:
        raise RuntimeError(f"tome_b200: ABI version {lib.tome_abi_version()} != expected {ABI_VERSION}; rebuild")
    if path is None:
        _lib = lib
    return lib


def _check(rc: int, lib) -> None:
    if rc != 0:
        raise RuntimeError(f"tome_b200 error {rc}: {lib.tome_last_error().decode()}")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"tome_b200: {what} must live on a CUDA (sm_100) device; got {t.device}. "
                           "The token-merging path has no CPU fallback.")


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return TOME_F32
    if t.dtype == torch.bfloat16:
        return TOME_BF16
    raise RuntimeError(f"tome_b200: unsupported dtype {t.dtype} (fp32 and bf16 only)")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _view_of(t: torch.Tensor) -> TomeViewC:
    """(bm, tokens, c) tensor with unit channel stride -> tome_view."""
    return TomeViewC(t.stride(0), 0, t.stride(1), 1)


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class DevicePlan:
    """Device buffers of one matching plan (mirrors ``tome_plan``)."""

    __slots__ = ("bm", "n", "r", "class_token", "distill_token", "node_max", "node_idx", "src_idx", "unm_idx",
                 "dst_idx", "a_map", "b_off", "b_src", "b_head", "_ints", "_c", "device")

    def __init__(self, bm, n, r, class_token, distill_token, node_max, node_idx):
        na, nb = (n + 1) // 2, n // 2
        self.bm, self.n, self.r = bm, n, r
        self.class_token, self.distill_token = bool(class_token), bool(distill_token)
        self.node_max, self.node_idx = node_max, node_idx
        self.device = node_max.device
        sizes = [bm * r, bm * (na - r), bm * r, bm * na, bm * (nb + 1), bm * r, bm * nb * 4]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += _align(s)
        self._ints = torch.empty(tot, dtype=torch.int32, device=self.device)
        v = [self._ints[o:o + s] for o, s in zip(offs, sizes)]
        self.src_idx = v[0].view(bm, r)
        self.unm_idx = v[1].view(bm, na - r)
        self.dst_idx = v[2].view(bm, r)
        self.a_map = v[3].view(bm, na)
        self.b_off = v[4].view(bm, nb + 1)
        self.b_src = v[5].view(bm, r)
        self.b_head = v[6].view(bm, nb, 4)
        self._c = TomePlanC(bm, n, r, int(self.class_token), int(self.distill_token),
                            node_max.data_ptr(), node_idx.data_ptr(), self.src_idx.data_ptr(),
                            self.unm_idx.data_ptr(), self.dst_idx.data_ptr(), self.a_map.data_ptr(),
                            self.b_off.data_ptr(), self.b_src.data_ptr(), self.b_head.data_ptr())

    @property
    def na(self):
        return (self.n + 1) // 2

    @property
    def nb(self):
        return self.n // 2

    def c_ptr(self):
        return ctypes.byref(self._c)


def launch_count() -> int:
    """Kernels enqueued by libtome_b200 since it was loaded."""
    return int(load_library().tome_launch_count())


def device_check(device: Optional[int] = None) -> None:
    lib = load_library()
    if not torch.cuda.is_available():
        raise RuntimeError("tome_b200: no CUDA device visible; the token-merging path has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else device
    _check(lib.tome_device_check(dev), lib)


def match(metric: torch.Tensor, class_token=False, distill_token=False, algo: int = MATCH_AUTO,
          _return_workspace: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernel 1.  metric (bm, n, cm) fp32/bf16 -> node_max (bm, na) f32, node_idx (bm, na) i32."""
    lib = load_library()
    _require_cuda(metric, "metric")
    if metric.dim() != 3:
        raise RuntimeError(f"tome_b200: metric must be (batch, tokens, channels); got {tuple(metric.shape)}")
    if metric.dtype not in (torch.float32, torch.bfloat16):
        metric = metric.float()
    if metric.stride(2) != 1:
        metric = metric.contiguous()
    bm, n, cm = metric.shape
    na = (n + 1) // 2
    with torch.cuda.device(metric.device):
        node_max = torch.empty(bm, na, dtype=torch.float32, device=metric.device)
        node_idx = torch.empty(bm, na, dtype=torch.int32, device=metric.device)
        ws_bytes = lib.tome_match_workspace_bytes(bm, n, cm, algo)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=metric.device)
        view = _view_of(metric)
        _check(lib.tome_match(metric.data_ptr(), _dtype_code(metric), bm, n, cm, ctypes.byref(view),
                              int(bool(class_token)), int(bool(distill_token)), algo, node_max.data_ptr(),
                              node_idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream(metric)), lib)
    if _return_workspace:          # tests only: lets them inspect the tensor-core pass's pruning
        return node_max, node_idx, ws
    return node_max, node_idx


class HeadMeanMetric:
    """A matching metric that is the mean over heads of an attention key tensor, kept lazy so kernel 1's
    prologue takes the mean itself (no separate reduction kernel, no metric tensor in HBM).

    ``keys`` is any (Bm, H, N, d) view with unit channel stride; ``frames`` > 1 says the Bm axis is
    (b f) over a (B, H, S*F, d) tensor whose token axis is '(s f)' -- Motionformer's regrouping
    (tome/patch/motionformer.py:143-144)."""

    def __init__(self, keys: torch.Tensor, frames: int = 1):
        self.keys, self.frames = keys, int(frames)
        if self.frames == 1:
            bm, h, n, d = keys.shape
        else:
            b, h, sf, d = keys.shape
            bm, n = b * self.frames, sf // self.frames
        self.shape = (bm, n, d)
        self.heads = h
        self.dtype, self.device = keys.dtype, keys.device
        self.is_cuda = keys.is_cuda
        self.prefetched = None        # (r, class_token, distill_token, plan, stream) from tome.merge.prefetch_matching

    def size(self, i):
        return self.shape[i]

    def materialize(self) -> torch.Tensor:
        k = self.keys
        if self.frames == 1:
            return k.mean(1)
        b, h, sf, d = k.shape
        f = self.frames
        return k.reshape(b, h, sf // f, f, d).permute(0, 3, 1, 2, 4).mean(2).reshape(b * f, sf // f, d)


def match_heads(metric: "HeadMeanMetric", class_token=False, distill_token=False):
    """Kernel 1 on a lazy head-mean metric."""
    lib = load_library()
    k = metric.keys
    _require_cuda(k, "keys")
    bm, n, cm = metric.shape
    if k.stride(3) != 1 or not _tc_friendly(k, cm, (k.stride(0), k.stride(1), k.stride(2))):
        return match(metric.materialize(), class_token, distill_token)
    view = _heads_view(metric)
    na = (n + 1) // 2
    with torch.cuda.device(k.device):
        node_max = torch.empty(bm, na, dtype=torch.float32, device=k.device)
        node_idx = torch.empty(bm, na, dtype=torch.int32, device=k.device)
        ws_bytes = lib.tome_match_workspace_bytes(bm, n, cm, MATCH_TCGEN05)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=k.device)
        _check(lib.tome_match_heads(k.data_ptr(), _dtype_code(k), bm, metric.heads, n, cm, ctypes.byref(view),
                                    k.stride(1), int(bool(class_token)), int(bool(distill_token)),
                                    node_max.data_ptr(), node_idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream(k)), lib)
    return node_max, node_idx


def _heads_view(metric: "HeadMeanMetric") -> "TomeViewC":
    k, f = metric.keys, metric.frames
    if f == 1:
        return TomeViewC(k.stride(0), 0, k.stride(2), 1)
    return TomeViewC(k.stride(0), k.stride(2), f * k.stride(2), f)   # batch (b f): b by stride(0), f by one token


def _tc_friendly(t: torch.Tensor, cm: int, strides) -> bool:
    """What the tensor-core path needs of its input: channel pairs aligned, 16-byte bf16 planes."""
    pair = 2 * t.element_size()
    return (t.dtype in (torch.float32, torch.bfloat16) and cm % 8 == 0 and t.data_ptr() % pair == 0
            and all(int(s) % 2 == 0 for s in strides))


def plan_build(metric, r: int, class_token=False, distill_token=False, algo: int = MATCH_AUTO) -> DevicePlan:
    """Kernels 1 + 2 in one call (``tome_plan_build``): ``metric`` is a (bm, n, cm) tensor or a lazy
    HeadMeanMetric; ``r`` the effective r (> 0).  Returns the plan with node_max / node_idx filled in."""
    lib = load_library()
    heads_mode = isinstance(metric, HeadMeanMetric)
    if heads_mode:
        k = metric.keys
        _require_cuda(k, "keys")
        bm, n, cm = metric.shape
        if k.stride(3) != 1 or not _tc_friendly(k, cm, (k.stride(0), k.stride(1), k.stride(2))) or algo == MATCH_EXACT_SIMT:
            return plan_build(metric.materialize(), r, class_token, distill_token, algo)
        src, view, heads, stride_h = k, _heads_view(metric), metric.heads, k.stride(1)
    else:
        _require_cuda(metric, "metric")
        if metric.dim() != 3:
            raise RuntimeError(f"tome_b200: metric must be (batch, tokens, channels); got {tuple(metric.shape)}")
        if metric.dtype not in (torch.float32, torch.bfloat16):
            metric = metric.float()
        if metric.stride(2) != 1:
            metric = metric.contiguous()
        bm, n, cm = metric.shape
        src, view, heads, stride_h = metric, _view_of(metric), 1, 0
    na = (n + 1) // 2
    dev = src.device
    with torch.cuda.device(dev):
        node_max = torch.empty(bm, na, dtype=torch.float32, device=dev)
        node_idx = torch.empty(bm, na, dtype=torch.int32, device=dev)
        plan = DevicePlan(bm, n, r, class_token, distill_token, node_max, node_idx)
        ws_bytes = lib.tome_plan_build_workspace_bytes(bm, n, cm)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        _check(lib.tome_plan_build(src.data_ptr(), _dtype_code(src), heads, stride_h, ctypes.byref(view), cm, algo,
                                   plan.c_ptr(), ws.data_ptr(), ws_bytes, _stream(src)), lib)
    return plan


def match_tc_describe(bm: int, n: int, cm: int):
    """(column tiles, BN, byte offset of tile_max, byte offset of tile_cnt, fused refine?) -- tests only."""
    out = (ctypes.c_int64 * 5)()
    load_library().tome_match_tc_describe(bm, n, cm, out)
    return tuple(int(v) for v in out)


def plan_cluster_describe(bm: int, n: int):
    """(CTAs per cluster, A rows per CTA, B rows per CTA, tile width, smem bytes) of the one-launch plan kernel; CTAs = 0
    when the shape takes the multi-launch chain -- tests only."""
    out = (ctypes.c_int64 * 5)()
    load_library().tome_plan_cluster_describe(bm, n, out)
    return tuple(int(v) for v in out)


def rowmax(scores: torch.Tensor, class_token=False, distill_token=False) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = load_library()
    _require_cuda(scores, "scores")
    scores = scores.float().contiguous()
    bm, na, nb = scores.shape
    with torch.cuda.device(scores.device):
        node_max = torch.empty(bm, na, dtype=torch.float32, device=scores.device)
        node_idx = torch.empty(bm, na, dtype=torch.int32, device=scores.device)
        _check(lib.tome_rowmax(scores.data_ptr(), bm, na, nb, int(bool(class_token)), int(bool(distill_token)),
                               node_max.data_ptr(), node_idx.data_ptr(), _stream(scores)), lib)
    return node_max, node_idx


def select(node_max: torch.Tensor, node_idx: torch.Tensor, n: int, r: int, class_token=False,
           distill_token=False) -> DevicePlan:
    """Kernel 2.  ``r`` must be the effective r (> 0)."""
    lib = load_library()
    _require_cuda(node_max, "node_max")
    bm = node_max.shape[0]
    with torch.cuda.device(node_max.device):
        plan = DevicePlan(bm, n, r, class_token, distill_token, node_max.contiguous(), node_idx.contiguous())
        ws_bytes = lib.tome_select_workspace_bytes(bm, n)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=node_max.device)
        _check(lib.tome_select(plan.c_ptr(), ws.data_ptr(), ws_bytes, _stream(node_max)), lib)
    return plan


def _norm_args(norm, x):
    """(weight, bias, eps) of a LayerNorm the kernel can fuse, as raw pointers in x's dtype."""
    w, b, eps = norm
    if w.dtype != x.dtype or w.device != x.device or not w.is_contiguous() or w.numel() != x.shape[-1]:
        raise RuntimeError("tome_b200: fused LayerNorm needs a contiguous weight of x's dtype/device and length c")
    if b is not None and (b.dtype != x.dtype or not b.is_contiguous() or b.numel() != w.numel()):
        raise RuntimeError("tome_b200: fused LayerNorm bias must match the weight")
    return w.data_ptr(), (None if b is None else b.data_ptr()), float(eps)


def merge(plan: DevicePlan, x: torch.Tensor, mode: str, size: Optional[torch.Tensor] = None,
          hybrid_threshold: Optional[float] = None, want_size: bool = False, norm=None,
          residual: Optional[torch.Tensor] = None, out=None):
    """Kernel 3.  x (bm, n, c) -> (bm, n - r, c).  mode in wavg/sum/mean/max/amax/drop.
    Returns out, or (out, size_out, logsize_out) when ``want_size``; with ``norm=(weight, bias, eps)``
    the LayerNorm of the merged rows is produced in the same pass and appended to the result; with
    ``residual`` (same shape as x) the rows merged are ``x + residual`` (rounded to x's dtype).
    ``out``: optional caller-owned (out, size_out, logsize_out, normed) buffers to write into instead of
    allocating (entries may be None); the caller keeps them alive (C-ABI ownership rule)."""
    lib = load_library()
    _require_cuda(x, "x")
    o_out, o_size, o_log, o_norm = (tuple(out) + (None,) * 4)[:4] if out is not None else (None,) * 4
    if x.dim() != 3 or x.shape[0] != plan.bm or x.shape[1] != plan.n:
        raise RuntimeError(f"tome_b200: merge expects x of shape ({plan.bm}, {plan.n}, c); got {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"tome_b200: unsupported dtype {x.dtype} (fp32 and bf16 only)")
    if x.stride(2) != 1:
        x = x.contiguous()
    m = _MODES[mode]
    bm, n, c = x.shape
    nout = n - plan.r
    thr = float("nan") if hybrid_threshold is None else float(hybrid_threshold)
    with torch.cuda.device(x.device):
        out = o_out if o_out is not None else torch.empty(bm, nout, c, dtype=x.dtype, device=x.device)
        if out.shape != (bm, nout, c) or out.dtype != x.dtype or out.stride(2) != 1:
            raise RuntimeError(f"tome_b200: out buffer must be ({bm}, {nout}, {c}) of x's dtype")
        size_out = logsize_out = None
        so = lo = None
        if want_size:
            size_out = o_size if o_size is not None else torch.empty(bm, nout, dtype=torch.float32, device=x.device)
            logsize_out = o_log if o_log is not None else torch.empty(bm, nout, dtype=torch.float32, device=x.device)
            so, lo = size_out.data_ptr(), logsize_out.data_ptr()
        sp = None
        if size is not None:
            if m != MODE_WAVG:
                raise RuntimeError("tome_b200: size is only used by the wavg mode")
            size = size.reshape(bm, n).to(dtype=torch.float32).contiguous()
            sp = size.data_ptr()
        xv, ov = _view_of(x), _view_of(out)
        normed = None
        if norm is None and residual is None:
            _check(lib.tome_merge(plan.c_ptr(), x.data_ptr(), _dtype_code(x), c, ctypes.byref(xv), sp, m, thr,
                                  out.data_ptr(), ctypes.byref(ov), so, lo, _stream(x)), lib)
        else:
            rp = None
            if residual is not None:
                if residual.shape != x.shape or residual.dtype != x.dtype or residual.device != x.device:
                    raise RuntimeError("tome_b200: residual must have x's shape, dtype and device")
                if residual.stride() != x.stride():
                    residual = residual.contiguous() if x.is_contiguous() else residual.clone(memory_format=torch.preserve_format)
                    if residual.stride() != x.stride():
                        raise RuntimeError("tome_b200: residual must be laid out like x")
                rp = residual.data_ptr()
            wp = bp = npz = None
            eps = 0.0
            nv = ov
            if norm is not None:
                wp, bp, eps = _norm_args(norm, x)
                normed = o_norm if o_norm is not None else torch.empty_like(out)
                nv, npz = _view_of(normed), normed.data_ptr()
            _check(lib.tome_merge_add_norm(plan.c_ptr(), x.data_ptr(), rp, _dtype_code(x), c, ctypes.byref(xv), sp, m, thr,
                                           out.data_ptr(), ctypes.byref(ov), so, lo, wp, bp, eps, npz,
                                           ctypes.byref(nv), _stream(x)), lib)
    res = (out, size_out, logsize_out) if want_size else (out,)
    if norm is not None:
        res = res + (normed,)
    return res if len(res) > 1 else res[0]


def merge_frames(plan: DevicePlan, x: torch.Tensor, frames: int, mode: str, size: Optional[torch.Tensor] = None,
                 hybrid_threshold: Optional[float] = None, norm=None, residual: Optional[torch.Tensor] = None,
                 cls: Optional[torch.Tensor] = None):
    """Kernel 3 on TimeSformer / Motionformer token layout, without the rearrange copies.

    x is (B, 1 + P*T, C): a class token followed by tokens ordered '(p t)'.  The plan's matching batch
    is (b t) with P tokens each -- the reference reaches that with
    ``rearrange(x[:, 1:], 'b (p t) m -> (b t) p m')`` before and ``'(b t) p m -> b (p t) m'`` + ``cat`` with
    the class token after (tome/patch/timesformer.py:88-107, motionformer.py:150-168).  Here both are
    addressing (tome_view with inner = T): returns (out (B, 1 + P'*T, C), size' (B*T, P'), log size').

    ``residual`` (B*T, 1 + P, C): the spatial attention's output in ITS layout ('(b t) (1 + p)'); the patch tokens
    merged are x + residual, added inside the kernel (the reference: rearrange, cat, add -- timesformer.py:46-48).
    ``cls`` (B, C): the class-token row of the output (default: x's)."""
    lib = load_library()
    _require_cuda(x, "x")
    T = int(frames)
    B, L, C = x.shape
    P = (L - 1) // T
    if (L - 1) != P * T or plan.bm != B * T or plan.n != P:
        raise RuntimeError(f"tome_b200: merge_frames expects x (B, 1 + {plan.n}*T, C) with B*T == {plan.bm}; got {tuple(x.shape)}, T={T}")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"tome_b200: unsupported dtype {x.dtype} (fp32 and bf16 only)")
    if x.stride(2) != 1:
        x = x.contiguous()
    m = _MODES[mode]
    Pn = P - plan.r
    thr = float("nan") if hybrid_threshold is None else float(hybrid_threshold)
    same_layout = residual is not None and tuple(residual.shape) == tuple(x.shape) and T > 1      # Motionformer: x + attn_out
    cls_ok = C % (8 if x.dtype == torch.bfloat16 else 4) == 0 and C <= (2048 if x.dtype == torch.bfloat16 else 1024)
    with torch.cuda.device(x.device):
        out = torch.empty(B, 1 + Pn * T, C, dtype=x.dtype, device=x.device)
        cls_src = x[:, 0] if cls is None else cls
        cls_add = residual[:, 0] if same_layout else None
        if not (cls_ok and norm is not None):                  # (with a fused LayerNorm the class row is written below, one launch)
            if cls_ok and cls_src.stride(-1) == 1 and (cls_add is None or cls_add.stride(-1) == 1):
                cls_rows(cls_src, add=cls_add, sum_out=out[:, 0])
            else:
                out[:, 0] = cls_src if cls_add is None else cls_src + cls_add
        size_out = torch.empty(B * T, Pn, dtype=torch.float32, device=x.device)
        logsize_out = torch.empty(B * T, Pn, dtype=torch.float32, device=x.device)
        sp = None
        if size is not None:
            size = size.reshape(B * T, P).to(dtype=torch.float32).contiguous()
            sp = size.data_ptr()
        xv = TomeViewC(x.stride(0), x.stride(1), T * x.stride(1), T)
        ov = TomeViewC(out.stride(0), out.stride(1), T * out.stride(1), T)
        if norm is None and residual is None:
            _check(lib.tome_merge(plan.c_ptr(), x[:, 1:].data_ptr(), _dtype_code(x), C, ctypes.byref(xv), sp, m, thr,
                                  out[:, 1:].data_ptr(), ctypes.byref(ov), size_out.data_ptr(), logsize_out.data_ptr(),
                                  _stream(x)), lib)
            return out, size_out, logsize_out
        rp, rv = None, None
        if same_layout:                                        # residual laid out like x: 'b (1 + p t)'
            if residual.dtype != x.dtype or residual.device != x.device:
                raise RuntimeError("tome_b200: merge_frames residual must have x's dtype and device")
            if residual.stride(2) != 1:
                residual = residual.contiguous()
                cls_add = residual[:, 0]
            rp = residual[:, 1:].data_ptr()
            rv = ctypes.byref(TomeViewC(residual.stride(0), residual.stride(1), T * residual.stride(1), T))
        elif residual is not None:
            if tuple(residual.shape) != (B * T, 1 + P, C) or residual.dtype != x.dtype or residual.device != x.device:
                raise RuntimeError(f"tome_b200: merge_frames residual must be ({B * T}, {1 + P}, {C}) or x's shape, of x's dtype; got {tuple(residual.shape)}")
            if residual.stride(2) != 1:
                residual = residual.contiguous()
            rp = residual[:, 1:].data_ptr()
            rv = ctypes.byref(TomeViewC(residual.stride(0), 0, residual.stride(1), 1))     # batch (b t), tokens p
        wp = bp = npz = None
        eps, normed, nv = 0.0, None, None
        if norm is not None:
            wp, bp, eps = _norm_args(norm, x)
            normed = torch.empty_like(out)
            if cls_ok and cls_src.stride(-1) == 1 and (cls_add is None or cls_add.stride(-1) == 1):
                cls_rows(cls_src, add=cls_add, sum_out=out[:, 0], norm=norm, normed_out=normed[:, 0:1])     # class token row, one launch
            else:
                out[:, 0] = cls_src if cls_add is None else cls_src + cls_add
                normed[:, 0] = torch.nn.functional.layer_norm(out[:, 0], (C,), norm[0], norm[1], eps)
            nv = ctypes.byref(TomeViewC(normed.stride(0), normed.stride(1), T * normed.stride(1), T))
            npz = normed[:, 1:].data_ptr()
        _check(lib.tome_merge_add_norm_rv(plan.c_ptr(), x[:, 1:].data_ptr(), rp, rv, _dtype_code(x), C, ctypes.byref(xv), sp, m,
                                          thr, out[:, 1:].data_ptr(), ctypes.byref(ov), size_out.data_ptr(),
                                          logsize_out.data_ptr(), wp, bp, eps, npz, nv, _stream(x)), lib)
    return (out, size_out, logsize_out) + ((normed,) if norm is not None else ())


def _strides3(t: torch.Tensor):
    return (ctypes.c_int64 * 3)(*[int(v) for v in t.stride()[:3]])


def rows_add_layernorm(a: torch.Tensor, b: Optional[torch.Tensor], norm, sum_out: Optional[torch.Tensor],
                       normed_out: Optional[torch.Tensor]):
    """sum_out = a + b, normed_out = LayerNorm(sum) through (B, P, T, C) VIEWS (include/tome_b200.h:
    tome_rows_add_layernorm): every argument is a 4-d view with unit channel stride -- ``x[:, 1:].unflatten(1, (P, T))``,
    a permuted '(b t) (1 + p)' buffer ... -- so the row-order hops of the divided space-time blocks cost nothing.
    ``b``, ``sum_out`` or ``normed_out`` may be None."""
    lib = load_library()
    _require_cuda(a, "a")
    B, P, T, C = a.shape
    for t in (b, sum_out, normed_out):
        if t is not None and (tuple(t.shape) != (B, P, T, C) or t.stride(3) != 1 or t.dtype != a.dtype or t.device != a.device):
            raise RuntimeError("tome_b200: rows_add_layernorm needs (B, P, T, C) views of one dtype with unit channel stride")
    if a.stride(3) != 1:
        raise RuntimeError("tome_b200: rows_add_layernorm needs unit channel stride")
    wp = bp = None
    eps = 0.0
    if normed_out is not None:
        wp, bp, eps = _norm_args(norm, a)
    with torch.cuda.device(a.device):
        _check(lib.tome_rows_add_layernorm(a.data_ptr(), _strides3(a), None if b is None else b.data_ptr(),
                                           None if b is None else _strides3(b), _dtype_code(a), B, P, T, C, wp, bp, eps,
                                           None if sum_out is None else sum_out.data_ptr(),
                                           None if sum_out is None else _strides3(sum_out),
                                           None if normed_out is None else normed_out.data_ptr(),
                                           None if normed_out is None else _strides3(normed_out), _stream(a)), lib)


def cls_rows(a: torch.Tensor, add: Optional[torch.Tensor] = None, mean_src: Optional[torch.Tensor] = None,
             sum_out: Optional[torch.Tensor] = None, norm=None, normed_out: Optional[torch.Tensor] = None) -> None:
    """The class-token rows of a divided space-time block in one launch (include/tome_b200.h: tome_cls_rows):
    ``sum = a (+ mean_src.mean(1)) (+ add)`` on (B, C) row views, each step rounded to the dtype as the separate torch ops
    round; ``sum_out`` (B, C) gets the sum, ``normed_out`` (B, reps, C) gets LayerNorm(sum) in every replica.  All arguments
    are views with unit channel stride; ``mean_src`` is (B, T, C)."""
    lib = load_library()
    _require_cuda(a, "a")
    B, C = a.shape
    for t, nd in ((add, 2), (mean_src, 3), (sum_out, 2), (normed_out, 3)):
        if t is not None and (t.dim() != nd or t.size(0) != B or t.size(-1) != C or t.stride(-1) != 1 or t.dtype != a.dtype
                              or t.device != a.device):
            raise RuntimeError("tome_b200: cls_rows needs (B, [T,] C) views of one dtype with unit channel stride")
    if a.stride(1) != 1:
        raise RuntimeError("tome_b200: cls_rows needs unit channel stride")
    wp = bp = None
    eps = 0.0
    if normed_out is not None:
        wp, bp, eps = _norm_args(norm, a)
    with torch.cuda.device(a.device):
        _check(lib.tome_cls_rows(
            a.data_ptr(), a.stride(0), None if add is None else add.data_ptr(), 0 if add is None else add.stride(0),
            None if mean_src is None else mean_src.data_ptr(), 0 if mean_src is None else mean_src.stride(0),
            0 if mean_src is None else mean_src.stride(1), 0 if mean_src is None else mean_src.size(1), _dtype_code(a), B, C,
            None if sum_out is None else sum_out.data_ptr(), 0 if sum_out is None else sum_out.stride(0), wp, bp, eps,
            None if normed_out is None else normed_out.data_ptr(), 0 if normed_out is None else normed_out.stride(0),
            0 if normed_out is None else normed_out.stride(1), 0 if normed_out is None else normed_out.size(1), _stream(a)), lib)


def add_layernorm(a: torch.Tensor, b: torch.Tensor, norm):
    """(a + b, LayerNorm(a + b)) in one pass; ``norm=(weight, bias, eps)``.  b has a's shape, or a's shape
    without (or with a unit) batch axis, in which case it is broadcast over the batch (position embedding)."""
    lib = load_library()
    _require_cuda(a, "a")
    c = a.shape[-1]
    if b.dim() == a.dim() and b.shape[0] == 1 and a.shape[0] != 1:
        b = b[0]
    if a.dtype != b.dtype or b.device != a.device or (b.shape != a.shape and tuple(b.shape) != tuple(a.shape[1:])):
        raise RuntimeError("tome_b200: add_layernorm needs b of a's shape, or of a's shape without the batch axis, same dtype")
    a, b = a.contiguous(), b.contiguous()
    wp, bp, eps = _norm_args(norm, a)
    with torch.cuda.device(a.device):
        s = torch.empty_like(a)
        y = torch.empty_like(a)
        _check(lib.tome_add_rows_layernorm(a.data_ptr(), b.data_ptr(), b.numel() // c, _dtype_code(a), a.numel() // c, c,
                                           wp, bp, eps, s.data_ptr(), y.data_ptr(), _stream(a)), lib)
    return s, y


def linear_gelu_supported(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.is_contiguous()
            and weight.shape[0] % 256 == 0 and weight.shape[1] % 8 == 0 and x.shape[-1] == weight.shape[1]
            and (bias is None or (bias.dtype == torch.bfloat16 and bias.is_contiguous())) and not torch.is_grad_enabled())


_ACTS = {False: 0, None: 0, "none": 0, True: 1, "erf": 1, "gelu": 1, "gelu_fast": 2, "tanh_fast": 2}


def linear_gelu(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], gelu=True) -> torch.Tensor:
    """act(x @ weight^T + bias) from one tcgen05 GEMM with the activation in its epilogue (bf16).
    ``gelu``: True / "erf" (nn.GELU), "gelu_fast" (HuggingFace FastGELUActivation, ViViT), False / None (bias only)."""
    lib = load_library()
    _require_cuda(x, "x")
    k = x.shape[-1]
    x2 = x.reshape(-1, k)
    if x2.stride(1) != 1 or x2.stride(0) % 8 != 0:
        x2 = x2.contiguous()
    m, n = x2.shape[0], weight.shape[0]
    with torch.cuda.device(x.device):
        out = torch.empty(m, n, dtype=x.dtype, device=x.device)
        _check(lib.tome_linear_gelu(x2.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(), m, n, k,
                                    x2.stride(0), _ACTS[gelu], out.data_ptr(), _stream(x)), lib)
    return out.reshape(*x.shape[:-1], n)


def attn_short_usable(x: torch.Tensor, heads: int) -> bool:
    """tome_attn_short serves this (seqs, n_tok, C) attention input: CUDA inference, <= 32 tokens, head dim 64."""
    return (x.is_cuda and not torch.is_grad_enabled() and x.dim() == 3 and x.shape[1] <= 32 and x.shape[2] == 64 * heads
            and x.dtype in (torch.float32, torch.bfloat16) and os.environ.get("TOME_ATTN_SHORT", "1") != "0")


def attn_short(qkv: torch.Tensor, heads: int, scale: float) -> torch.Tensor:
    """softmax(scale q k^T) v per sequence and head, straight from the QKV GEMM's output (seqs, n_tok, 3 * heads * 64)
    (channel order (3, heads, d), as ``qkv.reshape(B, N, 3, H, d)`` reads it); returns (seqs, n_tok, heads * 64)."""
    lib = load_library()
    _require_cuda(qkv, "qkv")
    qkv = qkv.contiguous()
    seqs, tn, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    esz = qkv.element_size()
    with torch.cuda.device(qkv.device):
        out = torch.empty(seqs, tn, c, dtype=qkv.dtype, device=qkv.device)
        base = qkv.data_ptr()
        _check(lib.tome_attn_short(base, base + c * esz, base + 2 * c * esz, _dtype_code(qkv), seqs, tn, heads, d, tn * c3, c3,
                                   float(scale), out.data_ptr(), _stream(qkv)), lib)
    return out


def frames_attention_usable(x: torch.Tensor, heads: int, keys_per_frame: int) -> bool:
    """tome_frames_attention / tome_traj_temporal serve this trajectory-attention input: CUDA bf16 inference, head
    dimension 64, at most 256 keys per frame."""
    return (x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.bfloat16 and x.dim() == 3
            and x.shape[2] == 64 * heads and 1 <= keys_per_frame <= 256 and os.environ.get("TOME_FRAMES_ATTN", "1") != "0")


def frames_attention(qkv: torch.Tensor, heads: int, frames: int, scale: float, key_bias: Optional[torch.Tensor] = None,
                     want_diag: bool = True, lead: int = 1, unbiased_queries: int = 0):
    """Space stage of the trajectory attention on the QKV GEMM's output (B, lead + F*P, 3*heads*64): returns
    xs (B, F*P, F, heads*64) and, with ``want_diag``, x_diag (B, F*P, heads*64) = xs[b, s, frame(s)].
    ``key_bias`` (B, F*P) fp32: log size per key in the token order (proportional attention).  ``frames=1, lead=0``:
    plain attention over <= 256 tokens with a key bias; the first ``unbiased_queries`` queries take no bias."""
    lib = load_library()
    _require_cuda(qkv, "qkv")
    qkv = qkv.contiguous()
    B, N, c3 = qkv.shape
    C = c3 // 3
    S = N - lead
    P = S // frames
    if S != P * frames:
        raise RuntimeError(f"tome_b200: frames_attention expects {lead} + frames * P tokens; got {N} with {frames} frames")
    bp = None
    if key_bias is not None:
        key_bias = key_bias.to(torch.float32).reshape(B, S).contiguous()
        bp = key_bias.data_ptr()
    with torch.cuda.device(qkv.device):
        xs = torch.empty(B, S, frames, C, dtype=qkv.dtype, device=qkv.device)
        diag = torch.empty(B, S, C, dtype=qkv.dtype, device=qkv.device) if want_diag else None
        _check(lib.tome_frames_attention(qkv.data_ptr(), _dtype_code(qkv), B, N, heads, C // heads, frames, P, int(lead),
                                         int(unbiased_queries), float(scale), bp, xs.data_ptr(),
                                         None if diag is None else diag.data_ptr(), _stream(qkv)), lib)
    return xs, diag


def traj_temporal(q2: torch.Tensor, k2: torch.Tensor, vals: torch.Tensor, heads: int, scale: float) -> torch.Tensor:
    """Temporal stage of the trajectory attention: q2 (B, S, C), k2 / vals (B, S, F, C) -> (B, S, C)."""
    lib = load_library()
    _require_cuda(q2, "q2")
    q2, k2, vals = q2.contiguous(), k2.contiguous(), vals.contiguous()
    B, S, F_, C = k2.shape
    with torch.cuda.device(q2.device):
        out = torch.empty(B, S, C, dtype=q2.dtype, device=q2.device)
        _check(lib.tome_traj_temporal(q2.data_ptr(), k2.data_ptr(), vals.data_ptr(), _dtype_code(q2), B * S, F_, heads, C // heads,
                                      float(scale), out.data_ptr(), _stream(q2)), lib)
    return out


_SPLIT_CACHE = {}          # id(weight) -> (weak reference to the weight, (version, address), its bf16 planes)


def _cached_planes(weight: torch.Tensor) -> torch.Tensor:
    """The weight's split planes, valid while this very tensor object is alive and unmodified.  Keyed by identity with a
    liveness check, never by address alone: a freed model's weight address is reused by the next model's weights."""
    k = id(weight)
    stamp = (weight._version, weight.data_ptr())
    hit = _SPLIT_CACHE.get(k)
    if hit is not None and hit[0]() is weight and hit[1] == stamp:
        return hit[2]
    w3 = split3(weight.detach())
    _SPLIT_CACHE[k] = (weakref.ref(weight, lambda _r, k=k: _SPLIT_CACHE.pop(k, None)), stamp, w3)
    return w3


def split3(x: torch.Tensor) -> torch.Tensor:
    """fp32 (rows, k) -> bf16 (rows, 3k) planes [h | m | l] with h + m + l == x exactly."""
    lib = load_library()
    _require_cuda(x, "x")
    k = x.shape[-1]
    x2 = x.reshape(-1, k)
    if x2.stride(1) != 1 or x2.stride(0) % 4 != 0:
        x2 = x2.contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty(x2.shape[0], 3 * k, dtype=torch.bfloat16, device=x.device)
        _check(lib.tome_split3(x2.data_ptr(), x2.shape[0], k, x2.stride(0), out.data_ptr(), _stream(x)), lib)
    return out


def linear_f32_weight_ok(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    """The shape / dtype conditions tome_linear_f32 puts on a layer (n % 256 == 0, k % 32 == 0, fp32, contiguous)."""
    return (weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 2 and weight.is_contiguous()
            and weight.shape[0] % 256 == 0 and weight.shape[1] % 32 == 0
            and (bias is None or (bias.dtype == torch.float32 and bias.is_contiguous()))
            and os.environ.get("TOME_LINEAR_F32", "1") != "0")


def linear_f32_usable(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> bool:
    """tome_linear_f32 serves this fp32 linear: CUDA inference, n % 256 == 0, k % 32 == 0."""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and not torch.is_grad_enabled()
            and weight.dim() == 2 and weight.is_contiguous() and weight.shape[0] % 256 == 0 and weight.shape[1] % 32 == 0
            and x.shape[-1] == weight.shape[1] and x.numel() // x.shape[-1] >= 128
            and (bias is None or (bias.dtype == torch.float32 and bias.is_contiguous()))
            and os.environ.get("TOME_LINEAR_F32", "1") != "0")


class Planes:
    """An fp32 tensor held as its exact three-way bf16 split: ``data`` (rows, 3 * k) = [h | m | l] planes with
    h + m + l == x, ``shape`` the logical fp32 shape (..., k).  What tome_linear_f32 / tome_attention_f32 consume, and what
    they can produce directly so that a GEMM -> GEMM / GEMM -> attention hand-over needs no fp32 round trip."""

    def __init__(self, data: torch.Tensor, shape):
        self.data, self.shape = data, tuple(shape)

    device = property(lambda self: self.data.device)
    is_cuda = True
    dtype = torch.float32

    def float(self) -> torch.Tensor:
        k = self.shape[-1]
        d = self.data.float()
        return ((d[:, :k] + d[:, k:2 * k]) + d[:, 2 * k:]).reshape(self.shape)


def planes_to_f32(planes: "Planes", col0: int = 0, ncols: Optional[int] = None) -> torch.Tensor:
    """The fp32 tensor (exact: h + m + l) of columns ``col0 .. col0 + ncols`` of the last axis of a ``Planes``
    (include/tome_b200.h: tome_planes_sum); shape planes.shape[:-1] + (ncols,)."""
    lib = load_library()
    n = planes.shape[-1]
    ncols = n - col0 if ncols is None else ncols
    rows = planes.data.shape[0]
    dev = planes.data.device
    with torch.cuda.device(dev):
        out = torch.empty(*planes.shape[:-1], ncols, dtype=torch.float32, device=dev)
        _check(lib.tome_planes_sum(planes.data.data_ptr(), rows, n, int(col0), int(ncols), out.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream), lib)
    return out


def linear_f32(x, weight: torch.Tensor, bias: Optional[torch.Tensor], gelu: bool = False, terms: Optional[int] = None,
               out: str = "fp32"):
    """act(x @ weight^T + bias) in fp32 on the tensor cores (exact bf16 three-way split; ``terms`` plane products: 8 by default =
    all but l.l, which lies below the fp32 accumulator's resolution, 9 = all, TOME_LINEAR_F32_TERMS overrides;
    include/tome_b200.h: tome_linear_f32).  ``x``: an fp32 tensor or a ``Planes``; the weight's planes are cached until the weight changes.
    ``out``: "fp32" (tensor), "planes" (``Planes`` only: the next exact-split kernel's operand) or "both"."""
    lib = load_library()
    n, k = weight.shape
    w3 = _cached_planes(weight)
    if isinstance(x, Planes):
        x3, lead = x.data, x.shape[:-1]
    else:
        _require_cuda(x, "x")
        x3, lead = split3(x), tuple(x.shape[:-1])
    m = x3.shape[0]
    if terms is None:
        terms = int(os.environ.get("TOME_LINEAR_F32_TERMS", "8"))
    dev = x3.device
    with torch.cuda.device(dev):
        res = torch.empty(m, n, dtype=torch.float32, device=dev) if out in ("fp32", "both") else None
        res3 = torch.empty(m, 3 * n, dtype=torch.bfloat16, device=dev) if out in ("planes", "both") else None
        gelu_code = 2 if gelu == "gelu_fast" else int(bool(gelu))          # HF FastGELUActivation (ViViT) / nn.GELU / none
        _check(lib.tome_linear_f32(x3.data_ptr(), w3.data_ptr(), None if bias is None else bias.data_ptr(), m, n, k, gelu_code,
                                   int(terms), None if res is None else res.data_ptr(), None if res3 is None else res3.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream), lib)
    t = None if res is None else res.reshape(*lead, n)
    p3 = None if res3 is None else Planes(res3, lead + (n,))
    return t if out == "fp32" else p3 if out == "planes" else (t, p3)


def attention_f32_usable(qkv: torch.Tensor, heads: int, key_bias: Optional[torch.Tensor] = None) -> bool:
    """tome_attention_f32 serves this (B, N, 3 * heads * 64) fp32 QKV tensor: CUDA inference, head dimension 64."""
    return (qkv.is_cuda and qkv.dtype == torch.float32 and not torch.is_grad_enabled() and qkv.dim() == 3
            and qkv.shape[2] == 3 * 64 * heads and qkv.shape[1] >= 64
            and os.environ.get("TOME_ATTENTION_F32", "1") != "0")


def attention_f32_planes_ok(x: torch.Tensor, qkv_weight: torch.Tensor, heads: int) -> bool:
    """The QKV GEMM of this input can hand its result to tome_attention_f32 as planes only (64-channel heads, >= 64 tokens)."""
    return (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and x.dim() == 3 and x.shape[1] >= 64
            and qkv_weight.shape[0] == 3 * 64 * heads and os.environ.get("TOME_ATTENTION_F32", "1") != "0")


def attention_f32(qkv, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None, unbiased_queries: int = 0,
                  out: str = "fp32"):
    """softmax(scale q k^T + key_bias) v in fp32 accuracy on the tensor cores, from the QKV GEMM's output
    (B, N, 3 * heads * 64) (channel order (3, heads, 64)) as an fp32 tensor or as ``Planes``; returns (B, N, heads * 64)
    as a tensor, as ``Planes`` (out="planes": the projection's operand) or both.  ``key_bias`` (B, N) fp32."""
    lib = load_library()
    if isinstance(qkv, Planes):
        B, N, c3 = qkv.shape
        x3 = qkv.data
    else:
        _require_cuda(qkv, "qkv")
        B, N, c3 = qkv.shape
        x3 = split3(qkv.reshape(B * N, c3))
    c = c3 // 3
    bp = None
    if key_bias is not None:
        key_bias = key_bias.to(torch.float32).reshape(B, N).contiguous()
        bp = key_bias.data_ptr()
    dev = x3.device
    with torch.cuda.device(dev):
        res = torch.empty(B, N, c, dtype=torch.float32, device=dev) if out in ("fp32", "both") else None
        res3 = torch.empty(B * N, 3 * c, dtype=torch.bfloat16, device=dev) if out in ("planes", "both") else None
        _check(lib.tome_attention_f32(x3.data_ptr(), B, N, heads, c // heads, float(scale), bp, int(unbiased_queries),
                                      None if res is None else res.data_ptr(), None if res3 is None else res3.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), lib)
    p3 = None if res3 is None else Planes(res3, (B, N, c))
    return res if out == "fp32" else p3 if out == "planes" else (res, p3)


def cls_attention(qkv: torch.Tensor, heads: int, scale: float, query_token: int = 0) -> torch.Tensor:
    """One query token against the whole sequence (include/tome_b200.h: tome_cls_attention): qkv (B, N, 3 * heads * 64)
    contiguous, bf16 or fp32 -> (B, 1, heads * 64)."""
    lib = load_library()
    _require_cuda(qkv, "qkv")
    if qkv.dim() != 3 or not qkv.is_contiguous() or qkv.shape[-1] != 3 * heads * 64:
        raise RuntimeError("tome_b200: cls_attention needs a contiguous (B, N, 3 * heads * 64) tensor")
    B, N, _ = qkv.shape
    with torch.cuda.device(qkv.device):
        out = torch.empty(B, 1, heads * 64, dtype=qkv.dtype, device=qkv.device)
        _check(lib.tome_cls_attention(qkv.data_ptr(), _dtype_code(qkv), B, N, heads, 64, int(query_token), float(scale), out.data_ptr(),
                                      _stream(qkv)), lib)
    return out


def frames_attention_f32_usable(x: torch.Tensor, heads: int) -> bool:
    """tome_frames_attention_f32 / fp32 tome_traj_temporal serve this trajectory-attention input: CUDA fp32 inference, head
    dimension 64."""
    return (x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.float32 and x.dim() == 3 and x.shape[2] == 64 * heads
            and os.environ.get("TOME_ATTENTION_F32", "1") != "0")


def frames_attention_f32(qkv, heads: int, frames: int, scale: float, key_bias: Optional[torch.Tensor] = None, lead: int = 1,
                         want_diag: bool = True):
    """Space stage of the trajectory attention in fp32 accuracy (include/tome_b200.h: tome_frames_attention_f32) from the QKV
    GEMM's output (B, lead + F*P, 3 * heads * 64) as an fp32 tensor or ``Planes``: returns xs (B, F*P, F, heads*64) fp32, the
    same values as ``Planes`` (the K projection's operand) and x_diag (B, F*P, heads*64) = xs[b, s, frame(s)] (or None).
    ``key_bias`` (B, F*P) fp32 in the token order."""
    lib = load_library()
    if isinstance(qkv, Planes):
        B, N, c3 = qkv.shape
        x3 = qkv.data
    else:
        _require_cuda(qkv, "qkv")
        B, N, c3 = qkv.shape
        x3 = split3(qkv.reshape(B * N, c3))
    c = c3 // 3
    S = N - lead
    P = S // frames
    if S != frames * P:
        raise RuntimeError(f"tome_b200: frames_attention_f32: {N} tokens != {lead} + {frames} frames x P")
    bp = None
    if key_bias is not None:
        key_bias = key_bias.to(torch.float32).reshape(B, S).contiguous()
        bp = key_bias.data_ptr()
    dev = x3.device
    with torch.cuda.device(dev):
        xs = torch.empty(B, S, frames, c, dtype=torch.float32, device=dev)
        xs3 = torch.empty(B * S * frames, 3 * c, dtype=torch.bfloat16, device=dev)
        xd = torch.empty(B, S, c, dtype=torch.float32, device=dev) if want_diag else None
        _check(lib.tome_frames_attention_f32(x3.data_ptr(), B, N, heads, c // heads, frames, P, lead, float(scale), bp, xs.data_ptr(),
                                             xs3.data_ptr(), None if xd is None else xd.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream), lib)
    return xs, Planes(xs3, (B, S, frames, c)), xd


def attention_bf16_usable(qkv: torch.Tensor, heads: int, key_bias: Optional[torch.Tensor] = None) -> bool:
    return (qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.dim() == 3 and qkv.is_contiguous() and qkv.shape[-1] == 3 * heads * 64
            and qkv.data_ptr() % 16 == 0 and qkv.shape[0] <= 65535 and not torch.is_grad_enabled()
            and (key_bias is None or (key_bias.is_cuda and tuple(key_bias.shape) == tuple(qkv.shape[:2]))))


def attention_bf16(qkv: torch.Tensor, heads: int, scale: float, key_bias: Optional[torch.Tensor] = None,
                   unbiased_queries: int = 0) -> torch.Tensor:
    """softmax(scale * q k^T + key_bias) v per head on bf16 tensor cores (include/tome_b200.h: tome_attention_bf16) from the
    QKV GEMM's contiguous (B, N, 3 * heads * 64) output; ``key_bias`` (B, N) fp32 = log size of the key token
    (tome/patch/videomae.py:62-63); the first ``unbiased_queries`` queries take no bias.  Returns (B, N, heads * 64) bf16."""
    lib = load_library()
    _require_cuda(qkv, "qkv")
    if qkv.dtype != torch.bfloat16 or qkv.dim() != 3 or not qkv.is_contiguous() or qkv.shape[-1] != 3 * heads * 64:
        raise RuntimeError("tome_b200: attention_bf16 needs a contiguous (B, N, 3 * heads * 64) bf16 tensor")
    B, N, _ = qkv.shape
    bp = None
    if key_bias is not None:
        key_bias = key_bias.reshape(B, N).float().contiguous()
        bp = key_bias.data_ptr()
    with torch.cuda.device(qkv.device):
        out = torch.empty(B, N, heads * 64, dtype=torch.bfloat16, device=qkv.device)
        _check(lib.tome_attention_bf16(qkv.data_ptr(), B, N, heads, 64, float(scale), bp, int(unbiased_queries), out.data_ptr(),
                                       _stream(qkv)), lib)
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """F.linear, through tome_linear_f32 when it applies (fp32 CUDA inference), else the library GEMM."""
    if linear_f32_usable(x, weight, bias):
        return linear_f32(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


def patchify(x: torch.Tensor, tubelet: int, ph: int, pw: int, out_dtype: torch.dtype) -> torch.Tensor:
    """(B, C, T, H, W) clip -> (B, tokens, C * tubelet * ph * pw) tubelet rows in ``out_dtype`` (one pass, cast
    included): the operand of the tubelet-embedding GEMM."""
    lib = load_library()
    _require_cuda(x, "x")
    codes = {torch.float32: TOME_F32, torch.bfloat16: TOME_BF16, torch.uint8: TOME_U8}
    if x.dtype not in codes or out_dtype not in (torch.float32, torch.bfloat16) or x.dim() != 5:
        raise RuntimeError("tome_b200: patchify takes a 5-d fp32 / bf16 / uint8 clip and writes fp32 or bf16")
    x = x.contiguous()
    B, C, T, H, W = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty(B, (T // tubelet) * (H // ph) * (W // pw), C * tubelet * ph * pw, dtype=out_dtype, device=x.device)
        _check(lib.tome_patchify(x.data_ptr(), codes[x.dtype], B, C, T, H, W, tubelet, ph, pw, out.data_ptr(),
                                 codes[out_dtype], _stream(x)), lib)
    return out


def attn_key_bias(log_size: torch.Tensor, k: torch.Tensor, q: Optional[torch.Tensor], d: int, scale: float, lead: int = 0):
    """Write the two-term split of log_size / scale into channels d, d+1 of every key head (in place);
    k, q: (B, H, N, d + pad) views with unit channel stride; log_size (B, N - lead) fp32."""
    lib = load_library()
    _require_cuda(k, "k")
    B, H, N, da = k.shape
    if da < d + 2 or k.stride(3) != 1 or (q is not None and (q.shape != k.shape or q.stride(3) != 1 or q.dtype != k.dtype)):
        raise RuntimeError("tome_b200: attn_key_bias needs (B, H, N, >= d + 2) q/k views with unit channel stride")
    if log_size.dtype != torch.float32 or log_size.numel() != B * (N - lead) or log_size.device != k.device:
        raise RuntimeError("tome_b200: attn_key_bias needs fp32 log_size of shape (B, N - lead) on k's device")
    log_size = log_size.contiguous()
    qp = (None, 0, 0, 0) if q is None else (q.data_ptr(), q.stride(0), q.stride(2), q.stride(1))
    with torch.cuda.device(k.device):
        _check(lib.tome_attn_key_bias(log_size.data_ptr(), B, N, int(lead), H, int(d), float(scale), _dtype_code(k),
                                      k.data_ptr(), k.stride(0), k.stride(2), k.stride(1), *qp, _stream(k)), lib)


def merge_source(plan: DevicePlan, source: Optional[torch.Tensor], hybrid_threshold: Optional[float] = None
                 ) -> torch.Tensor:
    lib = load_library()
    thr = float("nan") if hybrid_threshold is None else float(hybrid_threshold)
    dev = plan.device
    with torch.cuda.device(dev):
        if source is None:
            n0, sp = plan.n, None
        else:
            _require_cuda(source, "source")
            source = source.to(torch.float32).contiguous()
            n0, sp = source.shape[2], source.data_ptr()
        out = torch.empty(plan.bm, plan.n - plan.r, n0, dtype=torch.float32, device=dev)
        _check(lib.tome_merge_source(plan.c_ptr(), sp, n0, thr, out.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), lib)
    return out


class SourceMap:
    """Compact form of the reference's dense source matrix (merge.py:372-384): ``group`` (bm, n0) int32 names the
    merged token that holds each original token (-1: dropped).  ``dense()`` expands to the fp32 (bm, tokens, n0)
    0/1 matrix the reference API exposes; ``argmax(dim=1)`` -- what tome/vis.py:55,102,146 asks of it -- is the
    group map itself."""

    def __init__(self, group: torch.Tensor, tokens: int):
        self.group, self.tokens = group, int(tokens)

    shape = property(lambda self: torch.Size((self.group.shape[0], self.tokens, self.group.shape[1])))
    device = property(lambda self: self.group.device)

    def dense(self) -> torch.Tensor:
        lib = load_library()
        bm, n0 = self.group.shape
        with torch.cuda.device(self.group.device):
            out = torch.empty(bm, self.tokens, n0, dtype=torch.float32, device=self.group.device)
            _check(lib.tome_source_dense(self.group.data_ptr(), bm, self.tokens, n0, out.data_ptr(), _stream(self.group)), lib)
        return out

    def argmax(self, dim: int = 1) -> torch.Tensor:
        if dim != 1:
            return self.dense().argmax(dim=dim)
        return self.group.clamp(min=0).long()          # an all-zero column's argmax is 0, as torch's is

    def __getitem__(self, item):
        return self.dense()[item]


def source_compose(plan: DevicePlan, source: Optional["SourceMap"], drop: bool = False,
                   hybrid_threshold: Optional[float] = None) -> "SourceMap":
    """One block's update of the compact source map."""
    lib = load_library()
    thr = float("nan") if hybrid_threshold is None else float(hybrid_threshold)
    dev = plan.device
    with torch.cuda.device(dev):
        if source is None:
            n0, gp = plan.n, None
        else:
            if source.tokens != plan.n or source.group.shape[0] != plan.bm:
                raise RuntimeError(f"tome_b200: source map of {source.group.shape[0]} x {source.tokens} tokens does not match "
                                   f"the plan ({plan.bm} x {plan.n})")
            n0, gp = source.group.shape[1], source.group.data_ptr()
        out = torch.empty(plan.bm, n0, dtype=torch.int32, device=dev)
        _check(lib.tome_source_compose(plan.c_ptr(), gp, n0, int(bool(drop)), thr, out.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), lib)
    return SourceMap(out, plan.n - plan.r)


class PhiloxStream:
    """Device-resident (seed, call) pair of the random-score stream (include/tome_b200.h: tome_random_rowmax)."""

    def __init__(self, seed: int, device, clip_offset: int = 0):
        self.seed, self.clip_offset = int(seed) & (2 ** 64 - 1), int(clip_offset)
        words = [self.seed & 0x7FFFFFFFFFFFFFFF if self.seed < 2 ** 63 else self.seed - 2 ** 64, 0]
        self.state = torch.tensor(words, dtype=torch.int64, device=device)

    def calls(self) -> int:
        return int(self.state[1].item())


def random_rowmax(stream: "PhiloxStream", bm: int, na: int, nb: int, class_token=False, distill_token=False,
                  want_scores: bool = False, advance: bool = True):
    lib = load_library()
    dev = stream.state.device
    with torch.cuda.device(dev):
        node_max = torch.empty(bm, na, dtype=torch.float32, device=dev)
        node_idx = torch.empty(bm, na, dtype=torch.int32, device=dev)
        scores = torch.empty(bm, na, nb, dtype=torch.float32, device=dev) if want_scores else None
        _check(lib.tome_random_rowmax(stream.state.data_ptr(), stream.clip_offset, bm, na, nb, int(bool(class_token)),
                                      int(bool(distill_token)), node_max.data_ptr(), node_idx.data_ptr(),
                                      None if scores is None else scores.data_ptr(), int(bool(advance)),
                                      torch.cuda.current_stream(dev).cuda_stream), lib)
    return (node_max, node_idx, scores) if want_scores else (node_max, node_idx)


def unmerge(plan: DevicePlan, x: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    _require_cuda(x, "x")
    x = x.contiguous()
    bm, nout, c = x.shape
    if bm != plan.bm or nout != plan.n - plan.r:
        raise RuntimeError(f"tome_b200: unmerge expects ({plan.bm}, {plan.n - plan.r}, c); got {tuple(x.shape)}")
    with torch.cuda.device(x.device):
        out = torch.empty(bm, plan.n, c, dtype=x.dtype, device=x.device)
        _check(lib.tome_unmerge(plan.c_ptr(), x.data_ptr(), _dtype_code(x), c, out.data_ptr(), _stream(x)), lib)
    return out


class TokenSets:
    """Two token sets as index lists into the token axis (``tome_match_sets`` / ``tome_group_reduce``): ``a_tok`` the
    sources, ``b_tok`` the destinations; (ra,) / (nb,) shared by the batch, or (bm, ra) / (bm, nb) per element."""

    def __init__(self, a_tok: torch.Tensor, b_tok: torch.Tensor):
        self.ra, self.nb = a_tok.shape[-1], b_tok.shape[-1]
        self.per_batch = a_tok.dim() == 2
        self.rows = torch.cat((a_tok, b_tok), -1).to(torch.int32).contiguous()
        self.stride_b = self.ra + self.nb if self.per_batch else 0


def match_sets(metric: torch.Tensor, sets: TokenSets) -> Tuple[torch.Tensor, torch.Tensor]:
    """Best destination ROW (and its canonical score) for every source row: (bm, ra) fp32, (bm, ra) int32."""
    lib = load_library()
    _require_cuda(metric, "metric")
    if metric.dtype not in (torch.float32, torch.bfloat16):
        metric = metric.float()
    if metric.stride(2) != 1:
        metric = metric.contiguous()
    bm, n, cm = metric.shape
    dev = metric.device
    with torch.cuda.device(dev):
        node_max = torch.empty(bm, sets.ra, dtype=torch.float32, device=dev)
        node_idx = torch.empty(bm, sets.ra, dtype=torch.int32, device=dev)
        ws_bytes = lib.tome_match_sets_workspace_bytes(bm, sets.ra, sets.nb, cm)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        view = _view_of(metric)
        _check(lib.tome_match_sets(metric.data_ptr(), _dtype_code(metric), bm, n, cm, ctypes.byref(view), sets.rows.data_ptr(),
                                   sets.stride_b, sets.ra, sets.nb, node_max.data_ptr(), node_idx.data_ptr(), ws.data_ptr(),
                                   ws_bytes, _stream(metric)), lib)
    return node_max, node_idx


def group_reduce(x: torch.Tensor, sets: TokenSets, dst_idx: torch.Tensor, mode: str) -> torch.Tensor:
    """(bm, nb, c): every destination row reduced with the source rows ``dst_idx`` assigns to it (include_self)."""
    lib = load_library()
    _require_cuda(x, "x")
    if mode not in ("sum", "mean", "max", "amax"):
        raise RuntimeError(f"tome_b200: group_reduce mode must be sum / mean / amax, got {mode!r}")
    _dtype_code(x)
    if x.stride(2) != 1:
        x = x.contiguous()
    bm, n, c = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty(bm, sets.nb, c, dtype=x.dtype, device=x.device)
        view = _view_of(x)
        _check(lib.tome_group_reduce(x.data_ptr(), _dtype_code(x), bm, n, c, ctypes.byref(view), sets.rows.data_ptr(),
                                     sets.stride_b, sets.ra, sets.nb, dst_idx.data_ptr(), _MODES[mode], out.data_ptr(),
                                     _stream(x)), lib)
    return out


def gather_rows(x: torch.Tensor, row_map: torch.Tensor) -> torch.Tensor:
    """out[b, t] = x[b, row_map[b, t]] (zeros where the map is negative); row_map (bm, n_out) int32."""
    lib = load_library()
    _require_cuda(x, "x")
    x = x.contiguous()
    row_map = row_map.to(torch.int32).contiguous()
    bm, n_in, c = x.shape
    n_out = row_map.shape[1]
    with torch.cuda.device(x.device):
        out = torch.empty(bm, n_out, c, dtype=x.dtype, device=x.device)
        _check(lib.tome_gather_rows(x.data_ptr(), _dtype_code(x), bm, n_in, c, row_map.data_ptr(), n_out, out.data_ptr(),
                                    _stream(x)), lib)
    return out
