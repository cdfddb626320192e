// Per-frame ("segmented") attention on tcgen05 / TMEM / TMA: the space stage of Motionformer's
// trajectory attention with the proportional-attention key bias (tome/patch/motionformer.py:105-115,
// slowfast/models/motionformer_vit_helper.py:196-214):
//     attn[b, h, s, f, :] = softmax_p( scale * q[b,h,s] . k[b,h,f,p] + log size[b, f, p] )      s: all F*P queries
//     xs[b, s, f, (h d)]  = sum_p attn[b,h,s,f,p] * v[b,h,f,p]
// Every query attends to the P keys of EACH frame separately (softmax per frame), so a (128-query tile, frame) pair is
// a self-contained problem: S = Q K_f^T (128 x P, one UMMA chain into TMEM), a complete softmax over the <= 256
// columns (no running maximum, no rescaling), O = P V_f (128 x 64).  The reference runs this stage as broadcast
// matmuls over a (B, h, F, S, P) score tensor; torch's fused attention falls back to its unfused path for the 5-D,
// masked call (10 ms of element-wise kernels per forward at the bench shape, profiles/r02_motionformer_launches_before.csv).
//
// One CTA per (128 queries, head, clip), looping over the F frames; 192 threads:
//   warp 0      TMA producer: Q once, then K_f / V_f boxes (PB = ceil16(P) rows x 64 channels, SWIZZLE_128B) read in
//               place from the QKV GEMM's output;
//   warp 1      TMEM allocation + single-thread MMA issue: S = Q K^T (kind::f16, both operands K-major), then
//               O = P V with P from shared memory (K-major) and V as the TMA left it (rows = keys: MN-major B);
//   warps 2-5   softmax, one thread per query row, straight out of TMEM: pass 1 row maximum (key bias added), pass 2
//               exp2, row sum, bf16 probabilities into the swizzled P tile -- in two halves of <= 128 keys so that the P
//               tile is 32 KB and two CTAs share an SM (one CTA's softmax runs beside the other's MMAs); then the
//               normalised output row, written once as 128 contiguous bytes.
// O reuses the TMEM columns of S (dead once the probabilities are out), so a CTA holds 256 TMEM columns.
#include "tc_ptx.cuh"

namespace tome {

constexpr int FA_BM = 128;          // queries per CTA == UMMA M == TMEM lanes
constexpr int FA_D = 64;            // head dimension (one 128-byte swizzle row of bf16)
constexpr int FA_THREADS = 192;
constexpr int FA_HALF_KS = 8;       // k-steps (16 keys each) per half of the P tile: 128 keys, 32 KB

struct FaParams {
  int B, N, S, F, P, PB, heads;     // N = tok0 + F * P tokens per clip, S = F * P queries (and keys)
  int tok0;                         // leading tokens that are neither queries nor keys (Motionformer's class token: 1)
  int nobias_q;                     // leading QUERIES whose logits take no key bias (TimeSformer's class token, timesformer.py:74)
  float scale_log2e;                // softmax scale * log2(e)
  const float* bias;                // (B, F * P) log size per key in the token order, or NULL
  __nv_bfloat16* xs;                // (B, S, F, heads * 64)
  __nv_bfloat16* x_diag;            // (B, S, heads * 64): xs[b, s, frame(s)], or NULL
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// shared-memory key bias as seen by one query row: padded keys stay -inf, unbiased rows (TimeSformer's class query) see 0
__device__ __forceinline__ float key_bias(float b, bool) { return b; }     // the row's bias array is chosen once per row
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// MN-major, SWIZZLE_128B descriptor of a (keys x 64 channels) tile as TMA writes it: 8-key groups 1024 bytes apart
// along K (SBO); a single 64-element chunk along N, so the leading offset is never used.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(FA_THREADS, 2)
frames_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C = p.heads * FA_D;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t kv_bytes = (uint32_t)p.PB * 128u;
  const uint32_t kv_pad = (kv_bytes + 1023u) & ~1023u;
  const uint32_t sm_q = base, sm_k = sm_q + FA_BM * 128u, sm_v = sm_k + kv_pad, sm_p = sm_v + kv_pad;   // P: 2 k-blocks x 16 KB
  const uint32_t sm_bias = sm_p + 2u * FA_BM * 128u;                 // 2 x 256 floats
  const uint32_t bars = sm_bias + 2048u;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16, bar_s = bars + 24, bar_p0 = bars + 32, bar_p1 = bars + 40,
                 bar_pfree = bars + 48, bar_o = bars + 56, bar_sfree = bars + 64, tmem_slot = bars + 72;
  float* bias_all = reinterpret_cast<float*>(gen + (sm_bias - base));
  uint8_t* p_gen = gen + (sm_p - base);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_kv);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1);
    mbar_init(bar_p0, 128); mbar_init(bar_p1, 128); mbar_init(bar_pfree, 1); mbar_init(bar_o, 1); mbar_init(bar_sfree, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int nks = p.PB >> 4;                                   // k-steps of 16 keys per frame
  const int h0 = nks < FA_HALF_KS ? nks : FA_HALF_KS;          // k-steps in the first half of the P tile
  const int row0 = b * p.N + p.tok0;                           // first query / key token of this clip in the (B*N, 3C) tensor

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, FA_BM * 128u);
      tma_load_2d(sm_q, &map_q, h * FA_D, row0 + qt * FA_BM, bar_q);
      for (int f = 0; f < p.F; ++f) {
        if (f > 0) mbar_wait_sleep(bar_s, (f - 1) & 1, 64);              // S(f-1) complete: the K buffer is free
        mbar_expect_tx(bar_k, kv_bytes);
        tma_load_2d(sm_k, &map_kv, C + h * FA_D, row0 + f * p.P, bar_k);
        if (f > 0) mbar_wait_sleep(bar_o, (f - 1) & 1, 64);              // O(f-1) complete: the V buffer is free
        mbar_expect_tx(bar_v, kv_bytes);
        tma_load_2d(sm_v, &map_kv, 2 * C + h * FA_D, row0 + f * p.P, bar_v);
      }
    }
  } else if (warp == 1) {
    // The MMA warp runs converged with warp-uniform operands and one elected lane issuing: under `if (lane == 0)` the
    // compiler wraps every tcgen05.mma in an ELECT / vote loop (~70 cycles per MMA, attn_f32.cu), and the 17 MMAs of a
    // frame sit on its serial path.
    // S = Q K^T: D fp32, A/B bf16, both K-major, M = 128, N = PB.   O = P V: B MN-major (bit 16), N = 64.
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.PB >> 3) << 17) | ((uint32_t)(FA_BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(FA_D >> 3) << 17) | ((uint32_t)(FA_BM >> 4) << 24);
    const uint64_t dq = make_sw128_desc(sm_q), dk = make_sw128_desc(sm_k), dp = make_sw128_desc(sm_p), dv = make_sw128_mn_desc(sm_v);
    constexpr uint64_t PBLK = (FA_BM * 128u) >> 4;               // one 64-key block of the P tile
    mbar_wait_sleep(bar_q, 0, 64);
    for (int f = 0; f < p.F; ++f) {
      const uint32_t ph = (uint32_t)(f & 1);
      if (f > 0) mbar_wait_sleep(bar_sfree, (f - 1) & 1, 64);          // O(f-1) has been read out of the columns S(f) lands in
      mbar_wait_sleep(bar_k, ph, 64);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) umma_bf16(tb, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc1, k ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait_sleep(bar_v, ph, 64);
      mbar_wait_sleep(bar_p0, ph, 64);
      tc_fence_after();
      if (leader) {
        for (int ks = 0; ks < h0; ++ks)                          // P: k-block ks / 4, +32 bytes per k-step; V: 16 keys = 2048 bytes
          umma_bf16(tb, dp + (uint64_t)(ks >> 2) * PBLK + (uint64_t)(2 * (ks & 3)), dv + (uint64_t)(128 * ks), idesc2, ks ? 1u : 0u);
        if (nks > h0) umma_commit(bar_pfree);
      }
      __syncwarp();
      if (nks > h0) {
        mbar_wait_sleep(bar_p1, ph, 64);
        tc_fence_after();
        if (leader) {
          for (int ks = h0; ks < nks; ++ks)
            umma_bf16(tb, dp + (uint64_t)((ks - h0) >> 2) * PBLK + (uint64_t)(2 * ((ks - h0) & 3)), dv + (uint64_t)(128 * ks), idesc2, 1u);
        }
        __syncwarp();
      }
      if (leader) umma_commit(bar_o);
      __syncwarp();
    }
  } else {
    const int q4 = warp & 3;                                   // TMEM lane quarter this warp may touch
    const int row = q4 * 32 + lane;                            // query row of the tile == TMEM lane
    const int st = (int)threadIdx.x - 64;                      // 0..127 among the softmax threads
    const int s = qt * FA_BM + row;                            // query index within the clip's S patch tokens
    const bool live = s < p.S;
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const bool biased = p.bias != nullptr && s >= p.nobias_q;  // per query row
    const float* bias_s = bias_all + (biased ? 0 : 256);
    for (int f = 0; f < p.F; ++f) {
      const uint32_t ph = (uint32_t)(f & 1);
      // key bias of this frame (times log2 e); padded keys get -inf so they vanish from max, sum and P
      asm volatile("bar.sync 1, 128;" ::: "memory");           // everyone is done with the previous frame's bias
      for (int j = st; j < 256; j += 128) {                     // [0, 256): with the key bias; [256, 512): without (padding only)
        bias_all[j] = j < p.P ? (p.bias ? __ldg(p.bias + (long long)b * p.S + f * p.P + j) * LOG2E : 0.f) : -INFINITY;
        bias_all[256 + j] = j < p.P ? 0.f : -INFINITY;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait_sleep(bar_s, ph, 32);
      tc_fence_after();
      // k-steps of 16 columns, two TMEM loads in flight per wait (a version with 32-column units and both halves of a
      // pair live spilled the parked probabilities to local memory and ran 1.5x slower: profiles/r02_frames_attn_ncu.txt)
      // pass 1: row maximum of t_j = s_j * scale * log2e + bias_j
      float m = -INFINITY;
      for (int ks = 0; ks < nks; ks += 2) {
        float va[16], vb[16];
        const bool two = ks + 1 < nks;                         // warp-uniform
        tmem_ld16_nowait(taddr + (uint32_t)(ks * 16), va);
        if (two) tmem_ld16_nowait(taddr + (uint32_t)(ks * 16 + 16), vb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float4* b4p = reinterpret_cast<const float4*>(bias_s + ks * 16);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 b4 = b4p[e >> 2];
          m = fmaxf(m, fmaxf(fmaxf(fmaf(va[e], p.scale_log2e, b4.x), fmaf(va[e + 1], p.scale_log2e, b4.y)),
                             fmaxf(fmaf(va[e + 2], p.scale_log2e, b4.z), fmaf(va[e + 3], p.scale_log2e, b4.w))));
        }
        if (two) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 b4 = b4p[4 + (e >> 2)];
            m = fmaxf(m, fmaxf(fmaxf(fmaf(vb[e], p.scale_log2e, b4.x), fmaf(vb[e + 1], p.scale_log2e, b4.y)),
                               fmaxf(fmaf(vb[e + 2], p.scale_log2e, b4.z), fmaf(vb[e + 3], p.scale_log2e, b4.w))));
          }
        }
      }
      // pass 2: probabilities of one k-step -> 8 packed bf16 pairs; the row sum in two chains
      float l0 = 0.f, l1 = 0.f;
      auto probs16 = [&](const float (&v)[16], int ks, uint32_t (&w)[8]) {
        const float4* b4p = reinterpret_cast<const float4*>(bias_s + ks * 16);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 b4 = b4p[e >> 2];
          const float p0 = ex2_approx(fmaf(v[e], p.scale_log2e, b4.x) - m);
          const float p1 = ex2_approx(fmaf(v[e + 1], p.scale_log2e, b4.y) - m);
          const float p2 = ex2_approx(fmaf(v[e + 2], p.scale_log2e, b4.z) - m);
          const float p3 = ex2_approx(fmaf(v[e + 3], p.scale_log2e, b4.w) - m);
          l0 += p0 + p1;
          l1 += p2 + p3;
          w[e >> 1] = pack_bf16(p0, p1);
          w[(e >> 1) + 1] = pack_bf16(p2, p3);
        }
      };
      // k-step `kp` of a half of the P tile: K-major SWIZZLE_128B, k-block = kp / 4 (64 keys, 16 KB), row = 128 bytes,
      // 16-byte chunk index (2 * (kp % 4) + i) XOR (row % 8)
      auto store16 = [&](int kp, const uint32_t (&w)[8]) {
        uint8_t* rowp = p_gen + (size_t)(kp >> 2) * (FA_BM * 128) + (size_t)row * 128;
        const int c0 = 2 * (kp & 3);
        *reinterpret_cast<uint4*>(rowp + (((c0) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(rowp + (((c0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
      };
      for (int ks = 0; ks < h0; ks += 2) {                     // first half straight into the P tile
        float va[16], vb[16];
        uint32_t w[8];
        const bool two = ks + 1 < h0;
        tmem_ld16_nowait(taddr + (uint32_t)(ks * 16), va);
        if (two) tmem_ld16_nowait(taddr + (uint32_t)(ks * 16 + 16), vb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        probs16(va, ks, w);
        store16(ks, w);
        if (two) { probs16(vb, ks + 1, w); store16(ks + 1, w); }
      }
      fence_async_smem();
      mbar_arrive(bar_p0);
      if (nks > h0) {
        // second half parked in registers until the MMAs of the first half have released the tile
        uint32_t park[FA_HALF_KS][8];
#pragma unroll
        for (int u = 0; u < FA_HALF_KS; ++u) {
          if (h0 + u < nks) {                                   // warp-uniform
            float va[16];
            tmem_ld16_nowait(taddr + (uint32_t)((h0 + u) * 16), va);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            probs16(va, h0 + u, park[u]);
          }
        }
        mbar_wait_sleep(bar_pfree, ph, 32);
#pragma unroll
        for (int u = 0; u < FA_HALF_KS; ++u)
          if (h0 + u < nks) store16(u, park[u]);
        fence_async_smem();
        mbar_arrive(bar_p1);
      }
      const float l = l0 + l1;
      // output row: O / l, written once (128 contiguous bytes of the (B, S, F, C) tensor)
      mbar_wait_sleep(bar_o, ph, 32);
      tc_fence_after();
      const float inv = 1.0f / l;
      float o0[32], o1[32];
      tmem_ld32(taddr, o0);
      tmem_ld32(taddr + 32u, o1);
      tc_fence_before();
      mbar_arrive(bar_sfree);                                   // the S / O columns may be overwritten
      if (live) {
        uint4 pk[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pk[e] = make_uint4(pack_bf16(o0[8 * e] * inv, o0[8 * e + 1] * inv), pack_bf16(o0[8 * e + 2] * inv, o0[8 * e + 3] * inv),
                             pack_bf16(o0[8 * e + 4] * inv, o0[8 * e + 5] * inv), pack_bf16(o0[8 * e + 6] * inv, o0[8 * e + 7] * inv));
          pk[4 + e] = make_uint4(pack_bf16(o1[8 * e] * inv, o1[8 * e + 1] * inv), pack_bf16(o1[8 * e + 2] * inv, o1[8 * e + 3] * inv),
                                 pack_bf16(o1[8 * e + 4] * inv, o1[8 * e + 5] * inv), pack_bf16(o1[8 * e + 6] * inv, o1[8 * e + 7] * inv));
        }
        uint4* dst = reinterpret_cast<uint4*>(p.xs + (((long long)b * p.S + s) * p.F + f) * C + h * FA_D);
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = pk[e];
        if (p.x_diag && s / p.P == f) {                        // the query's own frame: the trajectory "diagonal"
          uint4* dd = reinterpret_cast<uint4*>(p.x_diag + ((long long)b * p.S + s) * C + h * FA_D);
#pragma unroll
          for (int e = 0; e < 8; ++e) dd[e] = pk[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---- temporal stage of the trajectory attention (vit_helper.py:216-243): one query per (token, head) against the F
// per-frame keys k2 = proj_k(xs), values = xs itself (use_original_code) or v2.  One thread per (token, head); bf16 or
// fp32 tensors (the fp32 models: tome_frames_attention_f32 in front of it), fp32 arithmetic either way.
__device__ __forceinline__ void tt_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ void tt_load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void tt_store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ void tt_store8(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// FOUR threads per (token, head), 16 channels each (adjacent lanes: a quad reads one 128-byte head row per frame, partial dot
// products meet in two shuffles); MAXF = 8: loops unrolled, scores indexed statically.  (One thread per (token, head) with 64-
// channel q / accumulator rows and a 32-way select per frame to keep the scores in registers was 12 % of the bf16 Motionformer
// step; the same change took tome_attn_short from 35 us to a fraction.)
template <typename T, int MAXF>
__global__ void __launch_bounds__(128) traj_temporal_kernel(const T* __restrict__ q2, const T* __restrict__ k2, const T* __restrict__ vals,
                                                           long long rows, int F, int heads, float scale, T* __restrict__ out) {
  constexpr int DP = FA_D / 4;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * heads;
  const int part = (int)(gid & 3);
  const bool valid = (gid >> 2) < total;
  const long long quad = valid ? (gid >> 2) : total - 1;   // idle quads of the last warp recompute the last row (full-mask shuffles)
  const int h = (int)(quad % heads);
  const long long r = quad / heads;
  const int C = heads * FA_D;
  const int off = h * FA_D + part * DP;
  float q[DP];
#pragma unroll
  for (int c = 0; c < DP / 8; ++c) {
    float v[8];
    tt_load8(q2 + r * C + off + 8 * c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[8 * c + i] = v[i] * scale;
  }
  auto score = [&](int f) {
    const T* kp = k2 + (r * F + f) * C + off;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < DP / 8; ++c) {
      float v[8];
      tt_load8(kp + 8 * c, v);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        a0 = fmaf(q[8 * c + i], v[i], a0);
        a1 = fmaf(q[8 * c + i + 1], v[i + 1], a1);
      }
    }
    float d = a0 + a1;
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    return d;
  };
  float sc[MAXF];
  float m = -INFINITY;
  if constexpr (MAXF <= 8) {
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {                        // F is uniform: the shuffles stay converged
      sc[f] = f < F ? score(f) : -INFINITY;
      m = fmaxf(m, sc[f]);
    }
  } else {
#pragma unroll 1
    for (int f = 0; f < F; ++f) {
      const float d = score(f);
#pragma unroll
      for (int u = 0; u < MAXF; ++u) if (u == f) sc[u] = d;
      m = fmaxf(m, d);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int u = 0; u < MAXF; ++u) if (u < F) { sc[u] = __expf(sc[u] - m); l += sc[u]; }
  const float inv = 1.0f / l;
  float acc[DP];
#pragma unroll
  for (int c = 0; c < DP; ++c) acc[c] = 0.f;
  auto accumulate = [&](int f, float pw) {
    const T* vp = vals + (r * F + f) * C + off;
#pragma unroll
    for (int c = 0; c < DP / 8; ++c) {
      float v[8];
      tt_load8(vp + 8 * c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[8 * c + i] = fmaf(pw, v[i], acc[8 * c + i]);
    }
  };
  if constexpr (MAXF <= 8) {
#pragma unroll
    for (int f = 0; f < MAXF; ++f) if (f < F) accumulate(f, sc[f] * inv);
  } else {
#pragma unroll 1
    for (int f = 0; f < F; ++f) {
      float pw = 0.f;
#pragma unroll
      for (int u = 0; u < MAXF; ++u) if (u == f) pw = sc[u];
      accumulate(f, pw * inv);
    }
  }
  if (!valid) return;
#pragma unroll
  for (int c = 0; c < DP / 8; ++c) tt_store8(out + r * C + off + 8 * c, reinterpret_cast<const float(&)[8]>(acc[8 * c]));
}

// ---- host -------------------------------------------------------------------------------------------------------
int launch_frames_attention(const void* qkv, int B, int N, int heads, int F, int P, float scale, const float* bias, void* xs,
                            void* x_diag, int tok0, int nobias_q, cudaStream_t st) {
  if (tok0 < 0 || N != tok0 + F * P) return set_error(TOME_ERR_ARG, "tome_frames_attention: N=%d != lead + F*P (lead=%d F=%d P=%d)", N, tok0, F, P);
  if (P < 1 || P > 256) return set_error(TOME_ERR_UNSUPPORTED, "tome_frames_attention: %d keys per frame (1..256)", P);
  if (((uintptr_t)qkv & 15) || ((uintptr_t)xs & 15) || (x_diag && ((uintptr_t)x_diag & 15)))
    return set_error(TOME_ERR_ALIGN, "tome_frames_attention: buffers must be 16-byte aligned");
  FaParams p;
  p.B = B; p.N = N; p.S = F * P; p.F = F; p.P = P; p.PB = (P + 15) & ~15; p.heads = heads; p.tok0 = tok0; p.nobias_q = nobias_q;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias; p.xs = (__nv_bfloat16*)xs; p.x_diag = (__nv_bfloat16*)x_diag;
  const long long rows = (long long)B * N, cols = 3LL * heads * FA_D;
  alignas(64) CUtensorMap map_q, map_kv;
  int rc = make_bf16_map(&map_q, qkv, rows, cols, cols, FA_BM, "tome_frames_attention");
  if (rc) return rc;
  rc = make_bf16_map(&map_kv, qkv, rows, cols, cols, p.PB, "tome_frames_attention");
  if (rc) return rc;
  const size_t kv_pad = ((size_t)p.PB * 128 + 1023) & ~(size_t)1023;
  const size_t smem = 1024 + FA_BM * 128 + 2 * kv_pad + 2 * FA_BM * 128 + 2048 + 128;
  static PerDeviceOnce once;
  if (once.first_time()) TOME_CUDA(cudaFuncSetAttribute(frames_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((p.S + FA_BM - 1) / FA_BM, heads, B);
  if (grid.z > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_frames_attention: batch %d > 65535", B);
  frames_attn_kernel<<<grid, FA_THREADS, smem, st>>>(map_q, map_kv, p);
  TOME_LAUNCH_CHECK("frames_attn_kernel");
  return TOME_OK;
}

int launch_traj_temporal(const void* q2, const void* k2, const void* vals, int dtype, long long rows, int F, int heads, float scale,
                         void* out, cudaStream_t st) {
  if (F < 1 || F > 32) return set_error(TOME_ERR_UNSUPPORTED, "tome_traj_temporal: %d frames (1..32)", F);
  if (((uintptr_t)q2 & 15) || ((uintptr_t)k2 & 15) || ((uintptr_t)vals & 15) || ((uintptr_t)out & 15))
    return set_error(TOME_ERR_ALIGN, "tome_traj_temporal: buffers must be 16-byte aligned");
  const long long total = rows * heads * 4;              // four threads per (token, head)
  const unsigned grid = (unsigned)((total + 127) / 128);
#define TOME_TT(T_, M_) traj_temporal_kernel<T_, M_><<<grid, 128, 0, st>>>((const T_*)q2, (const T_*)k2, (const T_*)vals, rows, F, heads, scale, (T_*)out)
  if (dtype == TOME_BF16) { if (F <= 8) TOME_TT(__nv_bfloat16, 8); else TOME_TT(__nv_bfloat16, 32); }
  else if (dtype == TOME_F32) { if (F <= 8) TOME_TT(float, 8); else TOME_TT(float, 32); }
  else return set_error(TOME_ERR_DTYPE, "tome_traj_temporal: unsupported dtype %d", dtype);
#undef TOME_TT
  TOME_LAUNCH_CHECK("traj_temporal_kernel");
  return TOME_OK;
}

}  // namespace tome
