#!/usr/bin/env python3
"""Benchmark of the B200-native token-merging path on the workload BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1]): VideoMAE ViT-B 16x224, ToMe merge mode, constant r schedule
(model.r = (r, 0), r = 100), 8 synthetic clips per GPU, random-init weights.  A step = one forward of
one batch through the patched model.  Prints ONE JSON line (see the task contract):
  * `value` / `e2e`: clips/s of the fp32 model -- the reference's own precision (its benchmark,
    slowfast/utils/model_benchmark.py:21-45, runs fp32 without autocast; TF32 stays off like torch's
    default) -- with inputs resident in HBM, and through pinned host buffers (uint8 frames -> H2D ->
    forward -> logits D2H);
  * `bf16`: the same two numbers for the bf16 model (fp32 matching), a named extra;
  * `models`: clips/s of the other BASELINE.json configs at this run's GPU count -- TimeSformer bf16 r=18
    (config 3), Motionformer r=18 (config 4), ViViT r in {0, 300, 1568} and hybrid 0.4 (config 5);
  * `roofline`: the dominant hot-path kernel (the fused merge: residual add + merge_wavg + sizes + LayerNorm,
    HBM-bound) at the layer-0 shape with inputs AND outputs rotating through more than L2, plus the figure
    weighted over the 12 layer shapes of the token schedule;
  * `cpu_baseline`: the reference's CPU path (oracle/torch_port.py, kind "port") on this box's host cores:
    whole model, and the config-1 microbench (matching + merge_wavg at B=4, N=1568).

Multi-GPU: pure data parallelism, one process per GPU (torchrun), weights replicated from
the same seed, no collective on the data path; logits are all-gathered once per step over
NCCL (slowfast/utils/distributed.py:25-44, tools/test_net.py:159).  Timing = CUDA events,
barrier + synchronize on both sides, max over ranks.
"""
import argparse
import contextlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

NUM_CLASSES = 400
FRAMES, CROP = 16, 224
L2_BYTES = 126 * 2 ** 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU per step (experiments.sh:126)")
    ap.add_argument("--r", type=int, default=100)
    ap.add_argument("--schedule", type=float, default=0.0, help="r inflection: 0 const, -1 decreasing, +1 increasing")
    ap.add_argument("--mode", default="merge")
    ap.add_argument("--prop-attn", type=int, default=0, help="VideoMAE default False (videomae.py:173)")
    ap.add_argument("--dtype", default="fp32", choices=["bf16", "fp32"],
                    help="precision of the HEADLINE line (value / e2e); the other one is reported as a named extra")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--match-algo", type=int, default=0, help="0 auto, 1 exact SIMT, 2 tcgen05")
    ap.add_argument("--cpu-clips", type=int, default=8, help="clips per CPU-baseline step (default: the GPU arm's batch)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-micro", action="store_true")
    ap.add_argument("--skip-models", action="store_true", help="leave out the `models` key (configs 3-5)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled every few ms through
    NVML from a thread (the timed region of a default run is ~70 ms: nvidia-smi's own polling loop is too slow
    to land a sample in it); falls back to `nvidia-smi -lms` when NVML is unavailable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None
        self.samples, self.max_mhz, self.stop_flag, self.thread, self.source = [], None, False, None, None

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.samples.append((float(mhz), int(get(handle))))
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            try:
                handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *a):
        self.stop_flag = True
        if self.source == "nvml" and self.thread is not None:
            self.thread.join(timeout=1)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], self.max_mhz or 0, set()
        for mhz, mask in self.samples:
            sm.append(mhz)
            for name, bit in self.BITS.items():
                if mask & bit:
                    reasons.add(name)
        for row in self.rows:
            f = [s.strip() for s in row.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


MODELS = {
    # name: (frames, r, patch kwargs)  -- BASELINE.json configs 2-5, r / mode as experiments.sh:17-18,124-126,395-428
    "videomae": (16, (100, 0.0), dict(prop_attn=False)),
    "timesformer": (8, (18, 0.0), dict()),
    "motionformer": (16, (18, 0.0), dict()),
    "vivit": (32, (300, 0.0), dict()),
}


def build_model(name, device, dtype, r=None, patch_kw=None, args=None):
    """Host model (random init, seed 0) + tome.patch.<name>; ``r`` None = unpatched."""
    import hostmodels
    import tome
    torch.manual_seed(0)
    frames = MODELS[name][0]
    if name == "videomae":
        model = hostmodels.VideoMAE(arch="vit_base_patch16_224", num_classes=NUM_CLASSES, num_frames=frames,
                                    tubelet_size=2, use_mean_pooling=True, init_scale=0.001)
    elif name == "timesformer":
        model = hostmodels.TimeSformer(num_classes=NUM_CLASSES, num_frames=frames)
    elif name == "motionformer":
        model = hostmodels.Motionformer(num_classes=NUM_CLASSES, num_frames=frames)
        # as constructed every frame embeds identically (zeroed 3-D patch weight, zero temp_embed: SURVEY.md 8a quirks)
        torch.nn.init.trunc_normal_(model.patch_embed_3d.proj.weight, std=0.02)
        torch.nn.init.trunc_normal_(model.temp_embed, std=0.02)
    elif name == "vivit":
        model = hostmodels.ViViT(num_classes=NUM_CLASSES, num_frames=frames)
    else:
        raise KeyError(name)
    model = model.eval().to(device=device, dtype=dtype)
    if r is not None:
        getattr(tome.patch, name)(model, trace_source=False, **(patch_kw or {}))
        model.r = r
    return model


def build_videomae(device, dtype, args):
    return build_model("videomae", device, dtype, (args.r, args.schedule),
                       dict(prop_attn=bool(args.prop_attn), mode=args.mode, head_aggregation="mean", threshold=0.8))


def token_schedule(args, depth=12, n0=1568):
    from tome.utils import parse_r
    n, out = n0, []
    for r in parse_r(depth, (args.r, args.schedule)):
        r_eff = max(min(r, n // 2), 0)
        out.append((n, r_eff))
        n -= r_eff
    return out


@contextlib.contextmanager
def cpu_port_backend():
    """Route tome.patch.videomae's merge calls to oracle/torch_port.py (CPU baseline legs only)."""
    from oracle import torch_port as P
    mod = sys.modules["tome.patch.videomae"]
    names = ("bipartite_soft_matching", "bipartite_soft_matching_drop", "bipartite_soft_matching_hybrid",
             "merge_wavg", "merge_source")
    saved = {n: getattr(mod, n) for n in names}
    try:
        for n in names:
            setattr(mod, n, getattr(P, n))
        yield
    finally:
        for n, f in saved.items():
            setattr(mod, n, f)


def time_cpu_reference(args, steps, warmup):
    """The reference's CPU path (kind 'port'): fp32 VideoMAE-B on the host cores, merge path =
    oracle/torch_port.py, `cpu_clips` clips per step."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    a2 = argparse.Namespace(**vars(args))
    model = build_videomae(torch.device("cpu"), torch.float32, a2)
    x = torch.rand(args.cpu_clips, 3, FRAMES, CROP, CROP)
    times = []
    with cpu_port_backend(), torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model([x])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mean = sum(times) / len(times)
    return {"value": args.cpu_clips / mean, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps x {args.cpu_clips} clip(s) after {warmup} warm-up, fp32, "
                      f"oracle/torch_port.py merge path, {mean * 1e3:.0f} ms/step"}, mean


def time_cpu_config1():
    """SURVEY.md 8(d) config 1 on the host cores: bipartite_soft_matching + merge_wavg of the reference's ATen call
    mix (oracle/torch_port.py) at B=4, N=1568, C=768, r=100 -- M1: metric = x (Cm = 768, the literal BASELINE
    shape), M1': metric (4, 1568, 64) (the in-model k.mean(1) shape).  3 warm-up + 20 timed, median."""
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 1568, 768, generator=g)
    m64 = torch.randn(4, 1568, 64, generator=g)
    out = {"cores": cores, "torch_threads": torch.get_num_threads(), "shape": "B=4 N=1568 C=768 r=100 fp32",
           "protocol": "3 warm-up + 20 timed, median; oracle/torch_port.py (kind port)"}
    for key, metric in (("M1_metric_is_x_cm768", x), ("M1p_cm64", m64)):
        ts = []
        for i in range(23):
            t0 = time.perf_counter()
            merge, _ = P.bipartite_soft_matching(metric, 100)
            P.merge_wavg(merge, x)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[3:])
        med = ts[len(ts) // 2]
        out[key] = {"ms": med * 1e3, "clips_per_s": 4 / med}
    return out


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    # exactly K timed steps after W warm-up steps, each a bounded sample of the workload: as many clips per
    # step (up to the GPU arm's batch) as keeps the whole run near two minutes on this box's cores
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    probe, _ = time_cpu_reference(argparse.Namespace(**{**vars(args), "cpu_clips": 1}), steps=1, warmup=1)
    per_clip = 1.0 / probe["value"]
    args.cpu_clips = int(max(1, min(args.batch, 120.0 / ((steps + warmup) * per_clip))))
    cb, mean = time_cpu_reference(args, steps, warmup)
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": cb["value"], "unit": "clips/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cpu=False):
    return {
        "workload": "VideoMAE ViT-B 16x224 (tubelet 2x16x16, 1568 tokens, 12 layers, 400 classes), ToMe "
                    f"mode={args.mode} r=({args.r},{args.schedule:g}) prop_attn={bool(args.prop_attn)}, "
                    "synthetic torch.rand clips, random-init weights (seed 0)",
        "clips_per_gpu_per_step": args.batch,
        "token_schedule": [n for n, _ in token_schedule(args)],
        "parallelism": f"dp{args.gpus}",
        "cuda_graph": not args.no_graph,
        "fp32_arithmetic": "fp32 model, TF32 off everywhere (torch default, as slowfast/utils/model_benchmark.py runs it); the "
                           "host model's linear layers and attention run on tcgen05 through the exact three-way bf16 split of every fp32 "
                           "operand with fp32 accumulation (tome_linear_f32: " + os.environ.get("TOME_LINEAR_F32_TERMS", "8") + " of the nine plane "
                           "products -- the default leaves out l.l, <= 2^-32 of |x||w|, below the accumulator's resolution; measured equal to "
                           "the nine-product result at fp32 resolution and as close to fp64 as the library SGEMM: "
                           "tests/test_kernels_gpu.py::test_linear_f32_matches_fp64, ::test_linear_f32_eight_products_equal_nine_at_fp32_resolution; "
                           "TOME_LINEAR_F32_TERMS=9 runs all nine); TOME_LINEAR_F32=0 gives the library SGEMMs",
        "l2": "each step reads a different resident input batch (4 rotate) and the model's weights (344 MB fp32 / "
              "172 MB bf16): working set > 126 MB L2",
    }


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def graph_time(fns, reps=20):
    """Capture the callables back to back in ONE CUDA graph and replay it: time per callable
    without host launch overhead (CUDA events on the replaying stream, median of `reps`)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        st.record()
        g.replay()
        en.record()
        torch.cuda.synchronize()
        ts.append(st.elapsed_time(en) * 1e3 / len(fns))
    ts.sort()
    return sum(ts) / len(ts), ts[len(ts) // 2]


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def merge_alg_bytes(bm, n, r, c, e, fused):
    """Algorithmic bytes of one merge launch (DESIGN.md section 4 / SURVEY.md 8d kernel 3): read x (and the residual)
    once, write x' (and LayerNorm(x')) once, sizes in / out + log sizes, plan indices."""
    na = (n + 1) // 2
    k = 2 if fused else 1
    return bm * (k * n * c * e + k * (n - r) * c * e + (n - r) * 8 + na * 12)


class MergeBench:
    """The fused merge launch of the patched block (tome/patch/videomae.py: `x + attn` on the way in, merge_wavg +
    sizes + log sizes, norm2 on the way out -- ONE merge_gather_kernel<LN, RES>) at one (n, r), with inputs AND
    outputs rotating through more than the 126 MB L2: every launch reads cold HBM and its output lines are
    evicted to HBM by the launches that follow, so bytes / time is an HBM figure, not an L2 one."""

    def __init__(self, device, dtype, bm, n, r, c=768, cm=64, fused=True, seed=1):
        from tome import _native
        self.native, self.fused = _native, fused
        e = 2 if dtype == torch.bfloat16 else 4
        self.bytes = merge_alg_bytes(bm, n, r, c, e, fused)
        per_launch_out = bm * (n - r) * c * e
        self.nrot = max(3, int(math.ceil(1.5 * L2_BYTES / per_launch_out)))
        g = torch.Generator(device=device).manual_seed(seed)
        rnd = lambda *s: torch.randn(*s, device=device, dtype=dtype, generator=g)      # noqa: E731
        self.xs = [rnd(bm, n, c) for _ in range(self.nrot)]
        self.rs = [rnd(bm, n, c) for _ in range(self.nrot)] if fused else None
        self.outs = [(torch.empty(bm, n - r, c, device=device, dtype=dtype), torch.empty(bm, n - r, device=device),
                      torch.empty(bm, n - r, device=device),
                      torch.empty(bm, n - r, c, device=device, dtype=dtype) if fused else None) for _ in range(self.nrot)]
        self.plan = _native.plan_build(torch.randn(bm, n, cm, device=device, generator=g), r)
        self.lw = torch.ones(c, device=device, dtype=dtype)
        self.lb = torch.zeros(c, device=device, dtype=dtype)
        self.note = (f"{self.nrot} launches over rotating inputs and outputs ({self.nrot * per_launch_out >> 20} MiB of outputs > L2) "
                     "captured in one CUDA graph, CUDA events around the replay, / launches")

    def launch(self, i):
        if self.fused:
            return self.native.merge(self.plan, self.xs[i], "wavg", want_size=True, norm=(self.lw, self.lb, 1e-6),
                                     residual=self.rs[i], out=self.outs[i])
        return self.native.merge(self.plan, self.xs[i], "wavg", want_size=True, out=self.outs[i][:3])

    def time(self, reps=20):
        return graph_time([lambda i=i: self.launch(i) for i in range(self.nrot)], reps=reps)


def measured_traffic(bm, n, r, dtype_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused merge kernel, from THIS round's ncu
    capture if one was committed (profiles/r02_merge_traffic.json, written by tools/ncu_traffic.py); None otherwise."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_merge_traffic.json")))
    except Exception:
        return None, None
    for e in rec.get("entries", []):
        if (e.get("bm"), e.get("n"), e.get("r"), e.get("dtype")) == (bm, n, r, dtype_name):
            return e.get("dram_bytes_per_launch"), {k: e.get(k) for k in ("source", "mode", "dram_bytes_read", "dram_bytes_write", "launches")}
    return None, None


def micro_kernels(args, device, dtype):
    """Device time of the hot-path kernels at the workload's shapes.  Returns (roofline dict for the fused merge,
    per-kernel dict)."""
    from tome import _native
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dname = "bf16" if dtype == torch.bfloat16 else "fp32"
    bm, n, c, cm, r = args.batch, 1568, 768, 64, min(args.r, 784)
    na = (n + 1) // 2
    res = {}
    mb = MergeBench(device, dtype, bm, n, r, c, cm, fused=True)
    mean_us, med_us = mb.time()
    achieved = mb.bytes / (mean_us * 1e-6) / 1e9
    traffic, traffic_src = measured_traffic(bm, n, r, dname)
    roofline = {"kernel": "merge_gather_kernel<LN, RES> (residual add + merge_wavg + size + log size + LayerNorm, as the "
                          f"patched block launches it; layer-0 shape Bm={bm} N={n} C={c} r={r} {dname})",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": mb.bytes,
                "us_mean": mean_us, "us_median": med_us, "peak_source": peak_src, "timing": mb.note}
    del mb
    # the same launch over the 12 layer shapes of the token schedule: sum of bytes / sum of time
    tot_b, tot_us, per_layer = 0, 0.0, []
    for (ln, lr) in token_schedule(args):
        if lr <= 0:
            continue
        lb = MergeBench(device, dtype, bm, ln, lr, c, cm, fused=True, seed=ln)
        l_mean, _ = lb.time(reps=10)
        per_layer.append({"n": ln, "r": lr, "us": l_mean, "frac": lb.bytes / (l_mean * 1e-6) / 1e9 / hbm_peak})
        tot_b += lb.bytes
        tot_us += l_mean
        del lb
    roofline["schedule_weighted"] = {"achieved": tot_b / (tot_us * 1e-6) / 1e9, "frac": tot_b / (tot_us * 1e-6) / 1e9 / hbm_peak,
                                     "us_total": tot_us, "algorithmic_bytes_total": tot_b, "layers": per_layer}
    pb = MergeBench(device, dtype, bm, n, r, c, cm, fused=False)
    p_mean, p_med = pb.time()
    res["merge_wavg_plain"] = {"us_mean": p_mean, "us_median": p_med, "algorithmic_bytes": pb.bytes,
                               "GBps": pb.bytes / (p_mean * 1e-6) / 1e9, "frac": pb.bytes / (p_mean * 1e-6) / 1e9 / hbm_peak,
                               "kernels": "merge_gather_kernel (merge_wavg + size + log size only: tome.merge.merge_wavg)"}
    del pb
    g = torch.Generator(device=device).manual_seed(2)
    ms = [torch.randn(bm, n, cm, device=device, dtype=dtype, generator=g) for _ in range(8)]
    m_mean, m_med = graph_time([lambda i=i: _native.match(ms[i], algo=args.match_algo) for i in range(8)])
    flops = 2.0 * bm * na * (n // 2) * cm
    tf_sus = float(peaks.get("bf16_tflops_sustained", 1417.0))
    res["match"] = {"us_mean": m_mean, "us_median": m_med, "algorithmic_gflop": flops / 1e9,
                    "tflops_algorithmic": flops / (m_mean * 1e-6) / 1e12, "algo": args.match_algo or "auto",
                    "kernels": "kernel 1 (tome_match): normalise + bf16-split tcgen05 contraction + exact fp64 refine"}
    ks = [torch.randn(bm, n, 3, 12, cm, device=device, dtype=dtype, generator=g).permute(2, 0, 3, 1, 4)[1] for _ in range(4)]
    for (ln, lr) in ((n, r), (468, min(r, 234))):
        kl = [k[:, :, :ln] for k in ks]
        p_mean, p_med = graph_time([lambda i=i: _native.plan_build(_native.HeadMeanMetric(kl[i % 4]), lr) for i in range(8)])
        # bound of kernels 1 + 2 together: the K read (12 heads) + plan written, and the contraction on the tensor pipe
        kb = bm * ln * 12 * cm * (2 if dtype == torch.bfloat16 else 4) + bm * ((ln + 1) // 2) * 24
        fl = 2.0 * bm * ((ln + 1) // 2) * (ln // 2) * cm
        bound_us = max(kb / (hbm_peak * 1e9), fl / (tf_sus * 1e12)) * 1e6
        res[f"plan_build_heads12_n{ln}"] = {"us_mean": p_mean, "us_median": p_med, "bound_us": bound_us, "frac_of_bound": bound_us / p_mean,
                                            "kernels": "tome_plan_build on the lazy head-mean of K (12 heads): kernels 1 + 2"}
    nm, ni = _native.match(ms[0], algo=args.match_algo)
    s_mean, s_med = graph_time([lambda: _native.select(nm, ni, n, r) for _ in range(8)])
    res["select"] = {"us_mean": s_mean, "us_median": s_med, "kernels": "kernel 2 (tome_select)"}
    if dtype == torch.bfloat16:
        # caller-side tensor-core kernel: the MLP's fc1 + bias + erf GELU as one tcgen05 GEMM,
        # against the library GEMM + elementwise GELU it replaces, at the layer-0 shape
        tf_peak = float(peaks.get("bf16_tflops", 1662.0))
        xm = torch.randn(bm * n, c, device=device, dtype=dtype, generator=g)
        w1 = (torch.randn(4 * c, c, device=device, generator=g) * c ** -0.5).to(dtype)
        b1 = torch.zeros(4 * c, device=device, dtype=dtype)
        fl = 2.0 * bm * n * c * 4 * c
        f_mean, f_med = graph_time([lambda: _native.linear_gelu(xm, w1, b1) for _ in range(4)])
        t_mean, t_med = graph_time([lambda: torch.nn.functional.gelu(torch.nn.functional.linear(xm, w1, b1)) for _ in range(4)])
        res["linear_gelu"] = {"us_mean": f_mean, "us_median": f_med, "algorithmic_gflop": fl / 1e9,
                              "roofline": {"bound": "tensor", "achieved": fl / (f_mean * 1e-6) / 1e12, "peak": tf_peak,
                                           "unit": "TFLOP/s", "frac": fl / (f_mean * 1e-6) / 1e12 / tf_peak,
                                           "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)"},
                              "library_gemm_plus_gelu_us": t_mean,
                              "kernels": "linear_gelu_kernel (persistent tcgen05 GEMM 128x256x64, TMEM double-buffered, "
                                         "bias + erf GELU + TMA store in the epilogue)"}
    if dtype == torch.float32:
        # caller-side tensor-core kernel of the fp32 model: the QKV projection as an exact-split (nine bf16 plane products)
        # tcgen05 GEMM, against the library's fp32 GEMM (TF32 off) it replaces.  `achieved` counts the nine MMA products.
        tf_peak = float(peaks.get("bf16_tflops", 1662.0))
        xm = [torch.randn(bm * n, c, device=device, generator=g) for _ in range(3)]
        wq = torch.randn(3 * c, c, device=device, generator=g) * c ** -0.5
        bq = torch.zeros(3 * c, device=device)
        fl = 2.0 * bm * n * c * 3 * c
        with torch.no_grad():
            f_mean, f_med = graph_time([lambda i=i: _native.linear_f32(xm[i % 3], wq, bq) for i in range(3)])
            s_mean2, _ = graph_time([lambda i=i: _native.split3(xm[i % 3]) for i in range(3)])
            t_mean, _ = graph_time([lambda i=i: torch.nn.functional.linear(xm[i % 3], wq, bq) for i in range(3)])
        mma_us = f_mean - s_mean2
        res["linear_f32"] = {"us_mean": f_mean, "us_median": f_med, "split_us": s_mean2, "algorithmic_gflop": fl / 1e9,
                             "fp32_tflops_incl_split": fl / (f_mean * 1e-6) / 1e12, "library_fp32_gemm_us": t_mean,
                             "library_fp32_tflops": fl / (t_mean * 1e-6) / 1e12,
                             "roofline": {"bound": "tensor", "achieved": 9 * fl / (mma_us * 1e-6) / 1e12, "peak": tf_peak,
                                          "unit": "TFLOP/s", "frac": 9 * fl / (mma_us * 1e-6) / 1e12 / tf_peak,
                                          "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)",
                                          "note": "nine bf16 plane products per fp32 product; time = linear_f32 minus the activation split"},
                             "kernels": "split3_kernel + linear_f32_kernel (persistent tcgen05 GEMM 128x256x32 on the exact three-way "
                                        "bf16 split, 256-channel accumulation chunks summed in registers)"}
    if dtype == torch.bfloat16:
        # Motionformer's space attention (config 4) at its layer-0 shape: per-frame softmax attention with the key bias
        qk = [torch.randn(bm, 1 + 8 * 196, 3 * c, device=device, generator=g).to(dtype) for _ in range(3)]
        kb = torch.rand(bm, 8 * 196, device=device, generator=g)
        a_mean, a_med = graph_time([lambda i=i: _native.frames_attention(qk[i % 3], 12, 8, 0.125, kb) for i in range(3)])
        afl = 4.0 * bm * 12 * (8 * 196) ** 2 * 64
        res["frames_attention"] = {"us_mean": a_mean, "us_median": a_med, "algorithmic_gflop": afl / 1e9,
                                   "tflops_algorithmic": afl / (a_mean * 1e-6) / 1e12,
                                   "kernels": "frames_attn_kernel (tcgen05 / TMEM / TMA per-frame attention, Motionformer layer 0: "
                                              "8 frames x 196 keys, 12 heads)"}
    xs = [torch.randn(bm, n, c, device=device, dtype=dtype, generator=g) for _ in range(12)]
    ys = [torch.empty_like(xs[0]) for _ in range(12)]
    c_mean, _ = graph_time([lambda i=i: ys[i].copy_(xs[i]) for i in range(12)])
    e = 2 if dtype == torch.bfloat16 else 4
    res["torch_copy_same_shape"] = {"us_mean": c_mean, "GBps": 2 * bm * n * c * e / (c_mean * 1e-6) / 1e9}
    return roofline, res


def capture(fn):
    """Warm `fn` on a side stream, then capture it into a CUDA graph; returns (graph, static output)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    return graph, out


def measure_forward(model, frames, dtype, args, device, rank, world, steps, warmup, want_e2e, sample_clocks=False):
    """clips/s of one patched model: `value` with inputs resident in HBM (rotating through 4 batches) and, when
    ``want_e2e``, through pinned host buffers.  Exactly `steps` timed steps after `warmup` (>= 3) untimed ones,
    barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    from tome import _native
    B, nrot = args.batch, 4
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    resident = [torch.rand(B, 3, frames, CROP, CROP, device=device, generator=g).to(dtype) for _ in range(nrot)]
    static_in = torch.empty_like(resident[0])
    logits_all = torch.empty(world * B, NUM_CLASSES, device=device, dtype=torch.float32) if world > 1 else None

    def forward():
        return model([static_in]).float()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    res = {}
    with torch.no_grad():
        launches_before = _native.launch_count()
        static_in.copy_(resident[0])
        out = forward()
        res["launches_per_step"] = _native.launch_count() - launches_before
        for i in range(3):
            static_in.copy_(resident[i % nrot])
            out = forward()
        torch.cuda.synchronize()
        graph, static_out = (None, None) if args.no_graph else capture(forward)

        def step(i):
            static_in.copy_(resident[i % nrot], non_blocking=True)
            if graph is not None:
                graph.replay()
                o = static_out
            else:
                o = forward()
            if world > 1:
                torch.distributed.all_gather_into_tensor(logits_all, o)
            return o

        for i in range(max(warmup, 3)):
            step(i)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with (ClockSampler(device.index) if sample_clocks else contextlib.nullcontext()) as clocks:
            barrier()
            start.record()
            for i in range(steps):
                out = step(i)
            end.record()
            barrier()
        res["ms"] = start.elapsed_time(end)
        res["clocks"] = clocks.summary() if sample_clocks else None
        res["top1"] = out.argmax(-1)[:4].tolist()
        del graph, static_out

        if want_e2e:
            # pinned uint8 frames (what a decoder hands over) -> H2D on a copy stream, double-buffered -> forward
            # (tome_patchify converts value / 255 to the model dtype in its own pass) -> logits D2H
            host_in = [torch.randint(0, 256, (B, 3, frames, CROP, CROP), dtype=torch.uint8).pin_memory() for _ in range(2)]
            host_out = torch.empty(B, NUM_CLASSES, dtype=torch.float32).pin_memory()
            stage = [torch.empty(B, 3, frames, CROP, CROP, device=device, dtype=torch.uint8) for _ in range(2)]
            copy_stream = torch.cuda.Stream()
            main = torch.cuda.current_stream()
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            graphs, outs = [None, None], [None, None]
            if not args.no_graph:
                for j in range(2):
                    stage[j].copy_(host_in[j])
                    graphs[j], outs[j] = capture(lambda j=j: model([stage[j]]).float())
                torch.cuda.synchronize()

            def e2e_loop(k):
                for i in range(k + 1):
                    if i < k:                     # prefetch step i's clips on the copy stream
                        with torch.cuda.stream(copy_stream):
                            if i >= 2:
                                copy_stream.wait_event(consumed[i % 2])
                            stage[i % 2].copy_(host_in[i % 2], non_blocking=True)
                            copied[i % 2].record(copy_stream)
                    if i >= 1:                    # run step i-1
                        j = i - 1
                        main.wait_event(copied[j % 2])
                        if graphs[j % 2] is not None:
                            graphs[j % 2].replay()
                            o = outs[j % 2]
                        else:
                            o = model([stage[j % 2]]).float()
                        consumed[j % 2].record(main)
                        if world > 1:
                            torch.distributed.all_gather_into_tensor(logits_all, o)
                        host_out.copy_(o, non_blocking=True)

            e2e_loop(max(warmup, 3))
            barrier()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record()
            e2e_loop(steps)
            e2.record()
            barrier()
            res["ms_e2e"] = s2.elapsed_time(e2)
            res["h2d"] = B * 3 * frames * CROP * CROP
            res["d2h"] = B * NUM_CLASSES * 4
            del graphs, outs
    keys = [k for k in ("ms", "ms_e2e") if k in res]
    t = torch.tensor([res[k] for k in keys], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    for k, v in zip(keys, t.tolist()):
        res[k] = v
    del resident, static_in
    torch.cuda.empty_cache()
    return res


def other_models(args, device, rank, world):
    """BASELINE.json configs 3-5 at this run's GPU count (weak scaling, `batch` clips per GPU): clips/s through the
    CUDA-graph forward, inputs resident in HBM, 10 timed steps after 3 warm-up.  Config 3 is bf16 by BASELINE's own
    wording; the others are given at the reference's precision (fp32) and in bf16."""
    runs = [("timesformer", torch.bfloat16, MODELS["timesformer"][1], {}, "config 3: TimeSformer divST 8x224, r=18 per frame, bf16")]
    for dt in (torch.float32, torch.bfloat16):
        runs.append(("motionformer", dt, MODELS["motionformer"][1], {}, "config 4: Motionformer 16x224 trajectory attention, r=18 per frame"))
    for dt in (torch.float32, torch.bfloat16):
        runs.append(("vivit", dt, None, {}, "config 5: ViViT-B 32x224, r=0 (unpatched)"))
        runs.append(("vivit", dt, (300, 0.0), {}, "config 5: ViViT-B, merge r=300"))
        runs.append(("vivit", dt, (1568, 0.0), {}, "config 5: ViViT-B, merge r=max (1568)"))
        runs.append(("vivit", dt, (300, 0.0), dict(mode="hybrid", threshold=0.4), "config 5: ViViT-B, hybrid r=300 threshold 0.4"))
    out = []
    for name, dt, r, kw, label in runs:
        model = build_model(name, device, dt, r, kw)
        m = measure_forward(model, MODELS[name][0], dt, args, device, rank, world, steps=10, warmup=3, want_e2e=False)
        out.append({"config": label, "model": name, "dtype": "bf16" if dt == torch.bfloat16 else "fp32",
                    "r": list(r) if r else None, "mode": kw.get("mode", "merge") if r else None,
                    "clips_per_s": world * args.batch * 10 / (m["ms"] * 1e-3), "ms_per_step": m["ms"] / 10,
                    "clips_per_gpu_per_step": args.batch, "steps": 10, "warmup": 3})
        del model
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the token-merging path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from tome import _native
    _native.device_check(local)
    if args.match_algo:
        orig_match = _native.match
        _native.match = lambda metric, c=False, d=False, algo=0: orig_match(metric, c, d, algo=args.match_algo)
    args.warmup = max(args.warmup, 3)
    head_dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    other_dtype = torch.bfloat16 if args.dtype == "fp32" else torch.float32
    B = args.batch

    model = build_videomae(device, head_dtype, args)
    head = measure_forward(model, FRAMES, head_dtype, args, device, rank, world, args.steps, args.warmup, want_e2e=True,
                           sample_clocks=True)
    del model
    torch.cuda.empty_cache()
    model = build_videomae(device, other_dtype, args)
    other = measure_forward(model, FRAMES, other_dtype, args, device, rank, world, args.steps, args.warmup, want_e2e=True)
    del model
    torch.cuda.empty_cache()
    models = None if args.skip_models else other_models(args, device, rank, world)

    roofline = roofline_other = kernels = kernels_other = cpu_baseline = None
    if rank == 0 and not args.skip_micro:
        roofline, kernels = micro_kernels(args, device, head_dtype)
        roofline_other, kernels_other = micro_kernels(args, device, other_dtype)
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu_baseline, _ = time_cpu_reference(args, steps=8, warmup=2)       # ~10 s of host work
        cpu_baseline["config1_microbench"] = time_cpu_config1()
    if world > 1:
        torch.distributed.barrier()

    if rank == 0:
        total_clips = world * B * args.steps

        def pair(m):
            return {"value": total_clips / (m["ms"] * 1e-3), "ms_per_step": m["ms"] / args.steps,
                    "e2e": {"value": total_clips / (m["ms_e2e"] * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": m["h2d"],
                            "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["ms_e2e"] / args.steps,
                            "note": "pinned uint8 frames -> H2D on a copy stream (double-buffered) -> forward (tome_patchify converts "
                                    "value / 255 to the model dtype) -> logits D2H"}}

        hp, op = pair(head), pair(other)
        name_h, name_o = ("f32", "bf16") if args.dtype == "fp32" else ("bf16", "f32")
        kern_lg = (kernels or {}).get("linear_gelu") or (kernels_other or {}).get("linear_gelu")
        kern_lf = (kernels or {}).get("linear_f32") or (kernels_other or {}).get("linear_f32")
        line = {
            "metric": "clips_per_sec", "value": hp["value"], "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": hp["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": name_h, "data": "synthetic",
            "config": workload_config(args),
            "clocks": head["clocks"],
            "e2e": hp["e2e"],
            "gpu_launches": head["launches_per_step"] * args.steps,
            "gpu_launches_per_step": head["launches_per_step"],
            name_o: dict(op, gpu_launches_per_step=other["launches_per_step"],
                         note=f"the same workload with a {name_o} model (matching always runs in fp32)"),
            "models": models,
            "roofline": roofline,
            "roofline_" + name_o: roofline_other,
            # the caller-side tensor-core kernel (fc1 + GELU of the bf16 model), the largest single kernel of libtome_b200 there
            "roofline_tensor": (dict(kern_lg["roofline"], kernel=kern_lg["kernels"], us_mean=kern_lg["us_mean"]) if kern_lg else None),
            # the same for the fp32 model: its linear layers as exact-split tensor-core GEMMs
            "roofline_tensor_f32": (dict(kern_lf["roofline"], kernel=kern_lf["kernels"], us_mean=kern_lf["us_mean"]) if kern_lf else None),
            "kernels": kernels, "kernels_" + name_o: kernels_other, "cpu_baseline": cpu_baseline,
            "match_algo": "auto" if not args.match_algo else args.match_algo,
            "top1_sample": head["top1"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
