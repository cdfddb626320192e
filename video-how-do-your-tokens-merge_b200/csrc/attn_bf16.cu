// bf16 flash attention with the proportional-attention key bias for sequences beyond one key block:
//     out[b, s, (h d)] = softmax_j( scale * q[b,h,s] . k[b,h,j] + log size[b, j] ) v[b,h,j]
// (tome/patch/videomae.py:58-68, vivit.py:98-117: `attn + size.log()[:, None, None, :, 0]`).  The library's fused attention
// takes a bias only as a (B, 1, N, N) mask (331 us at 8 x 12 x 1568 against 46 us unmasked); round 1 folded the bias into two
// spare channels of padded q / k heads so that the unmasked library kernel could run (100 us).  This kernel takes the
// (B, N) bias as it is, reads q, k, v in place from the QKV GEMM's (B, N, 3C) output and writes (B, N, C).
//
// Structure = attn_f32.cu with one bf16 plane instead of three (see there for how it got its shape):
//   one CTA per (128 queries, head, clip) walks the keys in blocks of 128 with a running maximum;
//   warp 0      TMA producer: K / V tiles (128 keys x 64 channels, SWIZZLE_128B, 3-stage ring, separate K / V barriers);
//   warp 1      S issuer   S = Q K^T : 4 MMAs 128 x 128 x 16, A = Q from tensor memory (written once from global memory);
//   warp 18     P V issuer O_j = P V : 8 MMAs 128 x 64 x 16, A = P from tensor memory, B = V as TMA wrote it (MN-major);
//               both issuers are converged warps with one elected lane (one UTCHMMA per MMA);
//   warps 2-17  FOUR threads per query row, 32 key columns of the block and 16 channels of O each (four softmax warps per
//               scheduler: with two the dependent-issue stalls were not hidden): scores out of TMEM once ->
//               + key bias -> row maximum (quarters exchanged through shared memory) -> exp2 / sum -> bf16 P into TMEM
//               (tcgen05.st); O is read once, after the last block.
// O stays in tensor memory for the whole key loop (P V accumulates): reading a 128 x 128 fp32 S tile out of TMEM is 1024 cycles
// of the 64 B / cycle read port per block, and folding O through registers every block added 512 more -- the port, not the
// tensor or the exponential pipe, set the pace (156 us at 8 x 12 x 1568).  The running maximum is therefore lazy: it moves,
// and O and the row sum are rescaled in place (tcgen05.ld / st of the row's own columns), only when a block beats it by more
// than 2^8 -- softmax is shift-invariant, so the result is the same function; probabilities stay below 2^8 and the sums in
// fp32.  P is double-buffered so that the softmax of block j + 1 never waits for P V (j).
// TMEM columns: S0 0..127 | S1 128..255 | O 256..319 | P0 320..383 | P1 384..447 | Q 448..479 (512 allocated: one CTA per SM).
#include "tc_ptx.cuh"

namespace tome {

constexpr int AB_BM = 128, AB_BKV = 128, AB_D = 64, AB_NST = 3;
constexpr int AB_THREADS = 608, AB_SM = 512;          // TMA, S issuer, 4 x 4 softmax warps, P V issuer
constexpr uint32_t AB_TILE = AB_BKV * 128u;           // one K or V tile: 128 keys x 128 bytes
constexpr uint32_t AB_STAGE = 2u * AB_TILE;           // K | V
constexpr uint32_t AB_S1 = 128u, AB_O = 256u, AB_P = 320u, AB_P1 = 64u, AB_Q = 448u;
constexpr float AB_LAZY = 8.0f;                       // the running maximum moves only when it is beaten by more than 2^8

struct AbParams {
  int B, N, heads, nblk, nobias_q;
  float scale_log2e;
  const float* bias;                                  // (B, N) log size per key, or NULL
  __nv_bfloat16* out;                                 // (B, N, heads * 64)
};

__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ab_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr) {     // MN-major SWIZZLE_128B: 8-key groups 1024 bytes apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t ab_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

#ifdef TOME_ATTN_TRACE
__device__ long long g_ab_trace[10 * 32];
#define AB_TR(ev, j) do { if (tr && (j) < 32) g_ab_trace[(ev) * 32 + (j)] = clock64(); } while (0)
#else
#define AB_TR(ev, j) do { } while (0)
#endif

template <bool HAS_BIAS>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bf16_kernel(const __grid_constant__ CUtensorMap map_kv, const __nv_bfloat16* __restrict__ qkv, const AbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C = p.heads * AB_D;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sm_kv = base, bars = sm_kv + (uint32_t)AB_NST * AB_STAGE;
  const uint32_t bar_q = bars, bar_p = bars + 8, bar_o = bars + 16, tmem_slot = bars + 24, bar_s = bars + 32, bar_sfree = bars + 48,
                 bar_kfull = bars + 64, bar_kempty = bars + 96, bar_vfull = bars + 128, bar_vempty = bars + 160, bar_o1 = bars + 184,
                 sm_xch = bars + 192;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_kv);
    mbar_init(bar_q, AB_SM); mbar_init(bar_p, AB_SM); mbar_init(bar_o, 1); mbar_init(bar_o1, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_s + 8u * s, 1); mbar_init(bar_sfree + 8u * s, AB_SM); }
    for (int s = 0; s < AB_NST; ++s) {
      mbar_init(bar_kfull + 8u * s, 1); mbar_init(bar_kempty + 8u * s, 1);
      mbar_init(bar_vfull + 8u * s, 1); mbar_init(bar_vempty + 8u * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int row0 = b * p.N;
  const int nb = p.nblk;
#ifdef TOME_ATTN_TRACE
  const bool tr = blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < nb; ++j) {
        const uint32_t s = (uint32_t)(j % AB_NST), k = (uint32_t)(j / AB_NST);
        const uint32_t st = sm_kv + s * AB_STAGE;
        if (k >= 1) mbar_wait_sleep(bar_kempty + 8u * s, (k - 1) & 1u, 32);
        mbar_expect_tx(bar_kfull + 8u * s, AB_TILE);
        tma_load_2d(st, &map_kv, C + h * AB_D, row0 + j * AB_BKV, bar_kfull + 8u * s);
        if (k >= 1) mbar_wait_sleep(bar_vempty + 8u * s, (k - 1) & 1u, 32);
        mbar_expect_tx(bar_vfull + 8u * s, AB_TILE);
        tma_load_2d(st + AB_TILE, &map_kv, 2 * C + h * AB_D, row0 + j * AB_BKV, bar_vfull + 8u * s);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AB_BKV >> 3) << 17) | ((uint32_t)(AB_BM >> 4) << 24);
    const uint64_t dk0 = make_sw128_desc(sm_kv);
    constexpr uint64_t ST = AB_STAGE >> 4;
    mbar_wait(bar_q, 0);
    for (int j = 0; j < nb; ++j) {                                          // up to two blocks ahead of the softmax
      const uint32_t s = (uint32_t)(j % AB_NST), k = (uint32_t)(j / AB_NST), sb = (uint32_t)(j & 1), kb = (uint32_t)(j >> 1);
      mbar_wait(bar_kfull + 8u * s, k & 1u);
      if (kb >= 1) mbar_wait(bar_sfree + 8u * sb, (kb - 1) & 1u);           // the softmax has pulled the previous S out of this buffer
      tc_fence_after();
      AB_TR(0, j);
      if (leader) {
        const uint32_t d = tb + sb * AB_S1;
        const uint64_t dk = dk0 + (uint64_t)s * ST;
#pragma unroll
        for (int ks = 0; ks < AB_D / 16; ++ks)                               // A: 8 columns per k-step; K: +32 bytes inside the swizzle row
          umma_bf16_ts(d, tb + AB_Q + 8u * ks, dk + (uint64_t)(2 * ks), idesc_s, ks ? 1u : 0u);
        umma_commit(bar_s + 8u * sb);
        umma_commit(bar_kempty + 8u * s);
      }
      __syncwarp();
      AB_TR(1, j);
    }
  } else if (warp == 18) {
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AB_D >> 3) << 17) | ((uint32_t)(AB_BM >> 4) << 24);
    const uint64_t dv0 = ab_desc_mn(sm_kv + AB_TILE);
    constexpr uint64_t ST = AB_STAGE >> 4;
    for (int j = 0; j < nb; ++j) {
      const uint32_t s = (uint32_t)(j % AB_NST);
      mbar_wait(bar_vfull + 8u * s, (uint32_t)((j / AB_NST) & 1));
      mbar_wait(bar_p, (uint32_t)(j & 1));
      tc_fence_after();
      AB_TR(2, j);
      if (leader) {
        const uint32_t d = tb + AB_O;
        const uint64_t dv = dv0 + (uint64_t)s * ST;
#pragma unroll
        for (int ks = 0; ks < AB_BKV / 16; ++ks)                             // P: 8 columns per 16 keys; V: 16 keys = 2048 bytes
          umma_bf16_ts(d, tb + AB_P + (uint32_t)(j & 1) * AB_P1 + 8u * ks, dv + (uint64_t)(128 * ks), idesc_o, (j || ks) ? 1u : 0u);
        umma_commit((j & 1) ? bar_o1 : bar_o);               // one barrier per P buffer: with P double-buffered a single one
        umma_commit(bar_vempty + 8u * s);                    // could run two phases ahead of a waiter and alias its parity
      }
      __syncwarp();
      AB_TR(3, j);
    }
  } else if (warp >= 2 && warp < 18) {
    const int q4 = warp & 3;
    const int qr = (warp - 2) >> 2;                            // which 32 key columns of a block / 16 channels of O
    const int row = q4 * 32 + lane;
    const int s_idx = qt * AB_BM + row;                        // query token within the clip
    const bool live = s_idx < p.N;
    const bool biased = HAS_BIAS && s_idx >= p.nobias_q;
    const uint32_t tlane = (uint32_t)(q4 * 32) << 16;
    const float* brow = HAS_BIAS ? p.bias + (long long)b * p.N : nullptr;
    const int st = (int)threadIdx.x - 64;                                  // 0..511 among the softmax threads
    float* xch = reinterpret_cast<float*>(gen + (sm_xch - base));          // [2 blocks][4 quarters][128 rows]
    float* bias_s = xch + 8 * AB_BM;                                       // [2 blocks][128 keys]
    const float LOG2E = 1.4426950408889634f;
    if (HAS_BIAS) {
      if (st < AB_BKV) bias_s[st] = st < p.N ? __ldg(brow + st) * LOG2E : -INFINITY;
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    {   // this row's query into tensor memory: 16 channels (32 bytes) per thread; rows past the tensor are zero
      const bool in = (long long)row0 + s_idx < (long long)p.B * p.N;
      const uint4* src = reinterpret_cast<const uint4*>(qkv + ((long long)row0 + s_idx) * (3LL * C) + h * AB_D + 16 * qr);
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4 v = in ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
      }
      tmem_st8(tmem_base + tlane + AB_Q + 8u * qr, w);
      tmem_st_wait();
      tc_fence_before();
      ab_arrive(bar_q);
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nb; ++j) {
      const uint32_t sb = (uint32_t)(j & 1), k = (uint32_t)(j >> 1);
      if (warp == 2) AB_TR(9, j);
      mbar_wait(bar_s + 8u * sb, k & 1u);
      tc_fence_after();
      if (warp == 2) AB_TR(4, j);
      float t[32];
      tmem_ld32(tmem_base + tlane + sb * AB_S1 + 32u * qr, t);
      if (warp == 2) AB_TR(5, j);
      tc_fence_before();
      ab_arrive(bar_sfree + 8u * sb);
      // logits in log2 units, key bias, padding
      const int key0 = j * AB_BKV + 32 * qr;
      if (HAS_BIAS) {
        // the block's bias row from shared memory (times log2 e, -inf beyond the clip: padding needs no second pass); the
        // next block's row is staged now and becomes visible at this block's row-maximum barrier
        const float4* bs = reinterpret_cast<const float4*>(bias_s + sb * AB_BKV + 32 * qr);
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 b4 = bs[e >> 2];
          if (!biased) {                                       // unbiased query rows still must not see the padding
            b4.x = b4.x == -INFINITY ? b4.x : 0.f; b4.y = b4.y == -INFINITY ? b4.y : 0.f;
            b4.z = b4.z == -INFINITY ? b4.z : 0.f; b4.w = b4.w == -INFINITY ? b4.w : 0.f;
          }
          t[e] = fmaf(t[e], p.scale_log2e, b4.x);
          t[e + 1] = fmaf(t[e + 1], p.scale_log2e, b4.y);
          t[e + 2] = fmaf(t[e + 2], p.scale_log2e, b4.z);
          t[e + 3] = fmaf(t[e + 3], p.scale_log2e, b4.w);
        }
        if (st < AB_BKV && j + 1 < nb) {
          const int key = (j + 1) * AB_BKV + st;
          bias_s[(sb ^ 1u) * AB_BKV + st] = key < p.N ? __ldg(brow + key) * LOG2E : -INFINITY;
        }
      }                                                        // (without a bias the scale rides in the exponent's FMA: scale > 0)
      if (!HAS_BIAS && key0 + 32 > p.N) {                      // last block: keys beyond the clip
#pragma unroll
        for (int e = 0; e < 32; ++e) if (key0 + e >= p.N) t[e] = -INFINITY;
      }
      float bmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; e += 4) bmax = fmaxf(bmax, fmaxf(fmaxf(t[e], t[e + 1]), fmaxf(t[e + 2], t[e + 3])));
      if (!HAS_BIAS) bmax *= p.scale_log2e;
      float* xr = xch + (sb * 4) * AB_BM + row;                // the row's other three quarters
      xr[qr * AB_BM] = bmax;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (warp == 2) AB_TR(6, j);
      bmax = fmaxf(fmaxf(xr[0], xr[AB_BM]), fmaxf(xr[2 * AB_BM], xr[3 * AB_BM]));   // finite: the first quarter of every block holds a real key
      // lazy running maximum: the row's four threads see the same numbers and take the same decision
      const bool move = bmax - m > AB_LAZY;                    // always on the first block (m = -inf)
      const bool any_move = __any_sync(0xffffffffu, move) && j > 0;
      float alpha = 1.0f;
      if (move) { alpha = ab_ex2(m - bmax); m = bmax; }        // alpha = 0 on the first block
      float l0 = 0.f, l1 = 0.f;
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ab_ex2(HAS_BIAS ? t[2 * i] - m : fmaf(t[2 * i], p.scale_log2e, -m));
        const float p1 = ab_ex2(HAS_BIAS ? t[2 * i + 1] - m : fmaf(t[2 * i + 1], p.scale_log2e, -m));
        l0 += p0; l1 += p1;
        w[i] = ab_pack(p0, p1);
      }
      l = fmaf(l, alpha, l0 + l1);
      if (warp == 2) AB_TR(7, j);
      if (j >= 2) {                                            // P V (j - 2) has read this P buffer
        mbar_wait(sb ? bar_o1 : bar_o, (uint32_t)(((j >> 1) - 1) & 1));
        tc_fence_after();
      }
      if (any_move) {                                          // rare after the first blocks: rescale the row's O columns in place
        mbar_wait(sb ? bar_o : bar_o1, (uint32_t)(((j - 1) >> 1) & 1));    // P V (j - 1): everything accumulated so far is in
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + tlane + AB_O + 16u * qr, v);
        uint32_t u[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) u[e] = __float_as_uint(v[e] * alpha);
        tmem_st16(tmem_base + tlane + AB_O + 16u * qr, u);
      }
      tmem_st16(tmem_base + tlane + AB_P + sb * AB_P1 + 16u * qr, w);       // this thread's 32 keys: columns 16 * qr .. + 15 of P
      tmem_st_wait();
      tc_fence_before();
      ab_arrive(bar_p);
      if (warp == 2) AB_TR(8, j);
    }
    mbar_wait(((nb - 1) & 1) ? bar_o1 : bar_o, (uint32_t)(((nb - 1) >> 1) & 1));
    tc_fence_after();
    float oacc[16];
    tmem_ld16(tmem_base + tlane + AB_O + 16u * qr, oacc);
    asm volatile("bar.sync 1, 512;" ::: "memory");             // the row sum of the four quarters (same running maximum)
    xch[qr * AB_BM + row] = l;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    l = (xch[row] + xch[AB_BM + row]) + (xch[2 * AB_BM + row] + xch[3 * AB_BM + row]);
    if (live) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.N + s_idx) * C + h * AB_D + 16 * qr);
#pragma unroll
      for (int e = 0; e < 2; ++e)
        dst[e] = make_uint4(ab_pack(oacc[8 * e] * inv, oacc[8 * e + 1] * inv), ab_pack(oacc[8 * e + 2] * inv, oacc[8 * e + 3] * inv),
                            ab_pack(oacc[8 * e + 4] * inv, oacc[8 * e + 5] * inv), ab_pack(oacc[8 * e + 6] * inv, oacc[8 * e + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int launch_attention_bf16(const void* qkv, int B, int N, int heads, float scale, const float* bias, int nobias_q, void* out,
                          cudaStream_t st) {
  if (!(scale > 0.f)) return set_error(TOME_ERR_ARG, "tome_attention_bf16: scale must be positive");
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || (bias && ((uintptr_t)bias & 15)))
    return set_error(TOME_ERR_ALIGN, "tome_attention_bf16: buffers must be 16-byte aligned");
  AbParams p;
  p.B = B; p.N = N; p.heads = heads; p.nblk = (N + AB_BKV - 1) / AB_BKV; p.nobias_q = nobias_q;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias; p.out = (__nv_bfloat16*)out;
  const long long rows = (long long)B * N, cols = 3LL * heads * AB_D;
  alignas(64) CUtensorMap map_kv;
  int rc = make_bf16_map(&map_kv, qkv, rows, cols, cols, AB_BKV, "tome_attention_bf16");
  if (rc) return rc;
  dim3 grid((N + AB_BM - 1) / AB_BM, heads, B);
  if (grid.z > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_attention_bf16: batch %d > 65535", B);
  const size_t smem = 1024 + AB_NST * AB_STAGE + 192 + (8 * AB_BM + 2 * AB_BKV) * sizeof(float);
  static PerDeviceOnce once;
  if (once.first_time()) {
    TOME_CUDA(cudaFuncSetAttribute(attn_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOME_CUDA(cudaFuncSetAttribute(attn_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (bias) attn_bf16_kernel<true><<<grid, AB_THREADS, smem, st>>>(map_kv, (const __nv_bfloat16*)qkv, p);
  else attn_bf16_kernel<false><<<grid, AB_THREADS, smem, st>>>(map_kv, (const __nv_bfloat16*)qkv, p);
  TOME_LAUNCH_CHECK("attn_bf16_kernel");
  return TOME_OK;
}

}  // namespace tome
#ifdef TOME_ATTN_TRACE
extern "C" TOME_API int tome_debug_attn_bf16_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, tome::g_ab_trace, sizeof(long long) * 10 * 32);
}
#endif
