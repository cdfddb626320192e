"""Probe (GPU): what does proportional attention's key bias cost through torch SDPA's backends, and
what do the alternatives cost (cuDNN native bias, augmented-K head_dim 72)?"""
import sys, torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3

B, H, N, D = 8, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 1568, 64
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3, H, D, device="cuda", generator=g).to(dt)
q, k, v = qkv.permute(2, 0, 3, 1, 4)
size = torch.randint(1, 9, (B, N, 1), device="cuda", generator=g).float()
ls = size.log()
bias_e = ls[:, None, None, :, 0].to(dt).expand(B, 1, N, N)
bias_c = bias_e.contiguous()
bias_k = ls[:, None, None, :, 0].to(dt)            # (B,1,1,N)
ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=ls[:, None, None, :, 0], scale=D ** -0.5)
print("no bias default", t(lambda: F.scaled_dot_product_attention(q, k, v, scale=D ** -0.5)))
for name, be in [("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)]:
    for bn, bb in [("none", None), ("expand", bias_e), ("contig", bias_c), ("b11n", bias_k)]:
        try:
            with sdpa_kernel(be):
                us = t(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=bb, scale=D ** -0.5))
                o = F.scaled_dot_product_attention(q, k, v, attn_mask=bb, scale=D ** -0.5)
            err = (o.float() - ref).abs().max().item() if bb is not None else float("nan")
            print(f"{name:10s} bias={bn:7s} {us:9.1f} us  maxerr {err:.3e}")
        except Exception as ex:
            print(f"{name:10s} bias={bn:7s} FAILED {str(ex)[:90]}")
print("default w/ expand bias", t(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=bias_e, scale=D ** -0.5)))
# augmented K: head dim 72 for q/k (extra channels: q = 1/scale, k = hi/lo split of log size), v stays 64
for DA in (72, 80, 96, 128):
    qa = torch.zeros(B, N, H, DA, device="cuda", dtype=dt); ka = torch.zeros_like(qa)
    qa[..., :D] = q.transpose(1, 2); ka[..., :D] = k.transpose(1, 2)
    hi = ls.to(dt); lo = (ls - hi.float()).to(dt)
    qa[..., D] = D ** 0.5; qa[..., D + 1] = D ** 0.5
    ka[..., D] = hi; ka[..., D + 1] = lo
    qq, kk = qa.transpose(1, 2), ka.transpose(1, 2)
    for name, be in [("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("default", None)]:
        try:
            def run():
                return F.scaled_dot_product_attention(qq, kk, v, scale=D ** -0.5)
            if be is None:
                us = t(run); o = run()
            else:
                with sdpa_kernel(be):
                    us = t(run); o = run()
            print(f"augK d={DA} {name:8s} {us:9.1f} us  maxerr {(o.float() - ref).abs().max().item():.3e}")
        except Exception as ex:
            print(f"augK d={DA} {name:8s} FAILED {str(ex)[:90]}")
    # v padded as well
    va = torch.zeros(B, N, H, DA, device="cuda", dtype=dt); va[..., :D] = v.transpose(1, 2); vv = va.transpose(1, 2)
    try:
        us = t(lambda: F.scaled_dot_product_attention(qq, kk, vv, scale=D ** -0.5))
        o = F.scaled_dot_product_attention(qq, kk, vv, scale=D ** -0.5)[..., :D]
        print(f"augK d={DA} v padded default {us:9.1f} us  maxerr {(o.float() - ref).abs().max().item():.3e}")
    except Exception as ex:
        print(f"augK d={DA} vpad FAILED {str(ex)[:90]}")
