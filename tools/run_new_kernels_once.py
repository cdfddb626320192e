#!/usr/bin/env python3
"""One launch of each round-2 kernel at its bench shape, for ncu (profiles/r02_new_kernels_ncu.txt)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-how-do-your-tokens-merge_b200")]
import torch
from tome import _native

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    # kernels 1 + 2: three-launch chain at VideoMAE layer 0 (12 bf16 heads)
    k = torch.randn(8, 1568, 3, 12, 64, device=dev, generator=g).to(torch.bfloat16).permute(2, 0, 3, 1, 4)[1]
    os.environ["TOME_PLAN_CLUSTER"] = "0"
    for _ in range(2):
        plan = _native.plan_build(_native.HeadMeanMetric(k), 100, False)
    # rows3 add + LayerNorm (TimeSformer layer 0: B 8, P 196, T 8)
    B, P, T, C = 8, 196, 8, 768
    x = torch.randn(B, 1 + P * T, C, device=dev, generator=g).to(torch.bfloat16)
    tf = torch.randn(B, P, T, C, device=dev, generator=g).to(torch.bfloat16)
    w = torch.ones(C, device=dev, dtype=torch.bfloat16); b = torch.zeros(C, device=dev, dtype=torch.bfloat16)
    xfull = torch.empty_like(x); ns = torch.empty(B * T, 1 + P, C, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        _native.rows_add_layernorm(x[:, 1:].unflatten(1, (P, T)), tf, (w, b, 1e-6), xfull[:, 1:].unflatten(1, (P, T)),
                                   ns.view(B, T, 1 + P, C)[:, :, 1:].permute(0, 2, 1, 3))
    # attn_short (TimeSformer temporal attention, layer 0)
    qkv = torch.randn(B * P, T, 3 * C, device=dev, generator=g).to(torch.bfloat16)
    for _ in range(2):
        _native.attn_short(qkv, 12, 0.125)
    # frames attention (Motionformer layer 0)
    q2 = torch.randn(8, 1 + 8 * 196, 3 * C, device=dev, generator=g).to(torch.bfloat16)
    kb = torch.rand(8, 8 * 196, device=dev, generator=g)
    for _ in range(2):
        _native.frames_attention(q2, 12, 8, 0.125, kb)
    # fp32 linear (QKV projection of VideoMAE layer 0)
    xm = torch.randn(8 * 1568, C, device=dev, generator=g)
    wq = torch.randn(3 * C, C, device=dev, generator=g) * C ** -0.5
    for _ in range(2):
        _native.linear_f32(xm, wq, None)
    # fp32 attention (VideoMAE layer 0), bf16 attention with the key bias (opt-in kernel), class-token rows (TimeSformer)
    qf = torch.randn(8, 1568, 3 * C, device=dev, generator=g)
    for _ in range(2):
        _native.attention_f32(qf, 12, 0.125)
    qb = qf.to(torch.bfloat16)
    bias = torch.rand(8, 1568, device=dev, generator=g)
    for _ in range(2):
        _native.attention_bf16(qb, 12, 0.125, bias)
    res_s = torch.randn(B * T, 1 + P, C, device=dev, generator=g).to(torch.bfloat16)
    cls = torch.empty(B, C, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        _native.cls_rows(x[:, 0], mean_src=res_s.view(B, T, 1 + P, C)[:, :, 0], sum_out=cls)
torch.cuda.synchronize()
print("done")
