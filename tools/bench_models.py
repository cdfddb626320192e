#!/usr/bin/env python3
"""Full-size forward of the four patched host models on one GPU (BASELINE.json configs 2-5): clips/s with
and without ToMe, eager and CUDA-graph, bf16, synthetic clips, random-init weights.  Not the contract
benchmark (bench.py is) -- a scale check of the drop-in on every model family and numbers for DESIGN.md.

    python tools/bench_models.py [--batch 8] [--iters 10] [--models videomae,timesformer,motionformer,vivit]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
import hostmodels
import tome

CONFIGS = {
    # name: (builder, frames, ToMe r, patch kwargs)
    "videomae": (lambda: hostmodels.VideoMAE(num_classes=400, num_frames=16), 16, (100, 0), dict(prop_attn=False)),
    "timesformer": (lambda: hostmodels.TimeSformer(num_classes=400, num_frames=8), 8, (18, 0), dict()),
    "motionformer": (lambda: hostmodels.Motionformer(num_classes=400, num_frames=16), 16, (18, 0), dict()),
    "vivit": (lambda: hostmodels.ViViT(num_classes=400, num_frames=32), 32, (300, 0), dict()),
}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_MODELS_WATCHDOG", "240")), exit=True)   # a stuck phase names itself
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--models", default="videomae,timesformer,motionformer,vivit")
    ap.add_argument("--r", type=int, default=None, help="override the ToMe r of every model")
    ap.add_argument("--schedule", type=float, default=0.0, help="r inflection (-1 decreasing, 0 constant, +1 increasing)")
    ap.add_argument("--mode", default=None, help="merge | drop | hybrid | random_merge | random_drop")
    a = ap.parse_args()
    dev = torch.device("cuda")
    out = {}
    for name in a.models.split(","):
        build, frames, r, kw = CONFIGS[name]
        if a.r is not None:
            r = (a.r, a.schedule)
        if a.mode is not None:
            kw = dict(kw, mode=a.mode, threshold=0.6)
        torch.manual_seed(0)
        model = build().eval()
        if name == "motionformer":            # as constructed every frame embeds identically (zeroed 3-D patch weight,
            torch.nn.init.trunc_normal_(model.patch_embed_3d.proj.weight, std=0.02)      # zero temp_embed): SURVEY.md 8a quirks
            torch.nn.init.trunc_normal_(model.temp_embed, std=0.02)
        model = model.to(dev, torch.bfloat16)
        clip = torch.rand(a.batch, 3, frames, 224, 224, device=dev).to(torch.bfloat16)
        res = {}
        with torch.no_grad():
            res["plain_ms"] = timed(lambda: model([clip]), a.iters)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model([clip])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0):
                static0 = model([clip])
            res["plain_graph_ms"] = timed(g0.replay, a.iters)
            del g0, static0
            print(name, "plain", res["plain_ms"], res["plain_graph_ms"], flush=True)
            getattr(tome.patch, name)(model, **kw)
            model.r = r
            logits = model([clip]).float()
            assert torch.isfinite(logits).all(), name
            res["tokens_out"] = tuple(model._tome_info["size"].shape)
            res["tome_ms"] = timed(lambda: model([clip]), a.iters)
            print(name, "tome", res["tome_ms"], res["tokens_out"], flush=True)
            # CUDA graph of the patched forward
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model([clip])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static = model([clip])
            res["tome_graph_ms"] = timed(g.replay, a.iters)
            same = (static.float().argmax(-1) == logits.argmax(-1)).float().mean().item()
            res["graph_top1_agree"] = same
        for k in ("plain_ms", "plain_graph_ms", "tome_ms", "tome_graph_ms"):
            res[k.replace("_ms", "_clips_per_s")] = a.batch / res[k] * 1e3
        out[name] = res
        print(name, json.dumps(res), flush=True)
        del model, g, static
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
