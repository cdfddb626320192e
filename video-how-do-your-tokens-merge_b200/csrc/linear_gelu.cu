// Caller-side fusion ("MLP right after the merge"): out = GELU(x @ W^T + bias) in one
// kernel -- the first half of every block's MLP (slowfast/models/videomae_video_model_builder.py:40-56:
// fc1 -> nn.GELU (erf) -> fc2).  With a library GEMM the activation is a separate elementwise pass over
// the (tokens, 4C) tensor: 154 MB of traffic per block at the bench shape, 11 % of the whole forward
// (profiles/r01d_launch_summary.txt).  Here it rides in the GEMM's epilogue.
//
// bf16 in / bf16 out, fp32 accumulation in TMEM.  Persistent CTAs (one per SM) walk 128 x 256 output tiles:
//   warp 0      TMA producer: A (128 x 64) and W (256 x 64) boxes, SWIZZLE_128B, 4-stage mbarrier ring that
//               keeps running across tiles
//   warp 1      TMEM allocator (all 512 columns = two 128 x 256 fp32 accumulators) + single-thread MMA issuer:
//               4 x tcgen05.mma kind::f16 (M 128, N 256, K 16) per stage, tcgen05.commit frees the stage /
//               publishes the accumulator
//   warps 2-17  epilogue, four warps per TMEM lane quarter (64 columns each): tcgen05.ld -> + bias -> round to
//               bf16 (what F.linear would have stored) -> erf GELU in fp32 (|error| <= 2e-7) -> bf16 -> a per-warp
//               swizzled shared-memory box -> TMA store (each lane owns a ROW of the tile, so direct stores were 32
//               scattered 16-byte pieces per instruction and capped the whole kernel at 68 us).
//               Accumulator t+1 is being filled while accumulator t is drained.
#include <math.h>

#include "tc_ptx.cuh"

namespace tome {

constexpr int LG_BM = 128, LG_BN = 256, LG_BK = 64, LG_STAGES = 4;
constexpr int LG_EPI_WARPS = 16;                      // four per TMEM lane quarter, 64 columns each
constexpr int LG_THREADS = 64 + 32 * LG_EPI_WARPS;    // TMA, MMA, epilogue warps
constexpr uint32_t LG_A_BYTES = LG_BM * 128u, LG_B_BYTES = LG_BN * 128u, LG_STAGE_BYTES = LG_A_BYTES + LG_B_BYTES;
constexpr uint32_t LG_OBOX_BYTES = 32u * 64u;         // one epilogue warp's output box: 32 rows x 32 bf16 columns (SWIZZLE_64B)

struct LinearGeluParams {
  int m, n, k, num_kb, tiles_n, tiles;
  const __nv_bfloat16* bias;                          // (n) or NULL
  __nv_bfloat16* out;                                 // (m, n) row-major
  int gelu;                                           // 0: plain linear (bias only), 1: erf GELU, 2: HF "gelu_fast" (tanh form)
};

// erf GELU.  libdevice's erff is two branches and ~50 instructions: with 32 k elements per tile the epilogue then
// takes longer than the tile's MMAs (90 us for the whole GEMM against 68 us without the activation).  Branch-free
// instead: Abramowitz & Stegun 7.1.26, erfc(|z|) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1 / (1 + p |z|),
// |error| <= 1.5e-7 -- four orders below the bf16 rounding of the result -- one MUFU.RCP, one MUFU.EX2, 8 FMAs.
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// u = 0.5 |x| erfc(|x| / sqrt 2) = 0.5 |x| poly t exp2(-x^2 log2(e) / 2);  GELU(x) = max(x, 0) - u.  Evaluated for a pair of
// values with packed f32x2 arithmetic (FFMA2 / FMUL2: two IEEE operations per issue slot, each lane rounded like its scalar
// form): the epilogue of the erf variant was short of issue slots (56.7 us against 50.2 without the activation), not of the
// XU pipe.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = fma2(make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f), ax, make_float2(1.0f, 1.0f));
  const float2 t = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
  float2 poly = fma2(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  poly = fma2(t, poly, make_float2(1.421413741f, 1.421413741f));
  poly = fma2(t, poly, make_float2(-0.284496736f, -0.284496736f));
  poly = fma2(t, poly, make_float2(0.254829592f, 0.254829592f));
  const float2 a = mul2(mul2(x, x), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  const float2 e = make_float2(exp2f_approx(a.x), exp2f_approx(a.y));
  const float2 u = mul2(mul2(mul2(make_float2(0.5f, 0.5f), ax), mul2(poly, t)), e);
  return sub2(make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)), u);
}

// HuggingFace FastGELUActivation (ViViT's hidden_act "gelu_fast", configs/vivit/kinetics/tome_vivit_8x32_224.json):
// 0.5 x (1 + tanh(u)), u = 0.7978845608 x (1 + 0.044715 x^2)  ==  x / (1 + exp(-2u)).  One MUFU.EX2 + one MUFU.RCP,
// evaluated in fp32 from the bf16-rounded pre-activation and rounded ONCE (the eager bf16 form rounds seven times).
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = x * 0.7978845608f * fmaf(0.044715f * x, x, 1.0f);
  const float e = exp2f_approx(u * -2.8853900817779268f);        // exp(-2u)
  return __fdividef(x, 1.0f + e);
}

__global__ void __launch_bounds__(LG_THREADS, 1)
linear_gelu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_o, const LinearGeluParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t obox = base + LG_STAGES * LG_STAGE_BYTES;             // one 2 KB box per epilogue warp
  const uint32_t bars = obox + (uint32_t)LG_EPI_WARPS * LG_OBOX_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * LG_STAGES;
  const uint32_t bar_tfull = bars + 16u * LG_STAGES, bar_tempty = bar_tfull + 16u;
  const uint32_t tmem_slot = bar_tempty + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_w);
    prefetch_tensormap(&map_o);
    for (int s = 0; s < LG_STAGES; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, LG_EPI_WARPS); }
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

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int mt = t / p.tiles_n, nt = t - mt * p.tiles_n;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % LG_STAGES, ph = (it / LG_STAGES) & 1u;
          mbar_wait(bar_empty + 8u * s, ph ^ 1u);
          const uint32_t st = base + s * LG_STAGE_BYTES, full = bar_full + 8u * s;
          mbar_expect_tx(full, LG_STAGE_BYTES);
          tma_load_2d(st, &map_a, kb * LG_BK, mt * LG_BM, full);
          tma_load_2d(st + LG_A_BYTES, &map_w, kb * LG_BK, nt * LG_BN, full);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LG_BN >> 3) << 17) | ((uint32_t)(LG_BM >> 4) << 24);
      uint32_t it = 0, tl = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        mbar_wait(bar_tempty + 8u * acc, aph ^ 1u);            // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)LG_BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % LG_STAGES, ph = (it / LG_STAGES) & 1u;
          mbar_wait(bar_full + 8u * s, ph);
          tc_fence_after();
          const uint32_t st = base + s * LG_STAGE_BYTES;
          const uint64_t a_desc = make_sw128_desc(st), b_desc = make_sw128_desc(st + LG_A_BYTES);
#pragma unroll
          for (int k = 0; k < LG_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);   // +32 bytes inside the swizzle row
            umma_bf16(d_tmem, a_desc + adv, b_desc + adv, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8u * s);
        }
        umma_commit(bar_tfull + 8u * acc);
      }
    }
  } else {
    const int q = warp & 3;                            // TMEM lane quarter this warp may touch
    const int part = (warp - 2) >> 2;                  // which 64 of the tile's 256 columns
    const uint32_t mybox = obox + (uint32_t)(warp - 2) * LG_OBOX_BYTES;
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++tl) {
      const int mt = t / p.tiles_n, nt = t - mt * p.tiles_n;
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      mbar_wait(bar_tfull + 8u * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (uint32_t)LG_BN + (uint32_t)(part * 64) + ((uint32_t)(q * 32) << 16);
      const __nv_bfloat16* brow = p.bias ? p.bias + nt * LG_BN + part * 64 : nullptr;
      float vv[2][32];                                 // both 32-column chunks in flight: one TMEM round trip per tile
      tmem_ld32_nowait(taddr, vv[0]);
      tmem_ld32_nowait(taddr + 32u, vv[1]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 2; ++c) {                    // two 32-column chunks, one 2 KB box and one TMA store each
        float (&v)[32] = vv[c];
        uint4 packed[4];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float bf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (brow) {
            const uint4 bw = __ldg(reinterpret_cast<const uint4*>(brow + c * 32 + g4 * 8));
            const uint32_t w[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) { bf[2 * i] = __uint_as_float(w[i] << 16); bf[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
          }
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x0 = v[g4 * 8 + 2 * i] + bf[2 * i], x1 = v[g4 * 8 + 2 * i + 1] + bf[2 * i + 1];
            if (p.gelu) {
              // round to bf16 first: the value F.linear would have stored and the activation would have read.  The packed
              // conversion is an ALU instruction and a bf16 is the upper half of its fp32; the single-value conversion goes
              // through the XU pipe (16 / clk / SM), which the activation's MUFU.RCP / EX2 already keep half busy
              const __nv_bfloat162 r2 = __floats2bfloat162_rn(x0, x1);
              const uint32_t rw = *reinterpret_cast<const uint32_t*>(&r2);
              x0 = __uint_as_float(rw << 16);
              x1 = __uint_as_float(rw & 0xFFFF0000u);
              if (p.gelu == 1) { const float2 y = gelu_erf2(make_float2(x0, x1)); x0 = y.x; x1 = y.y; }
              else { x0 = gelu_tanh_fast(x0); x1 = gelu_tanh_fast(x1); }
            }
            const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
          }
          packed[g4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // the previous TMA store out of this warp's box must have finished READING shared memory
        // (waited for only now: the arithmetic above ran beside it)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          // row = lane, 16-byte chunk g4 of the 64-byte box row, SWIZZLE_64B: chunk ^ ((row >> 1) & 3)
          const uint32_t dst = mybox + (uint32_t)lane * 64u + ((((uint32_t)g4) ^ (((uint32_t)lane >> 1) & 3u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[g4].x), "r"(packed[g4].y), "r"(packed[g4].z), "r"(packed[g4].w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          const int col = nt * LG_BN + part * 64 + c * 32, row0 = mt * LG_BM + q * 32;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(&map_o), "r"(col), "r"(row0), "r"(mybox) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8u * acc) : "memory");
    }
  }
  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store has landed
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int launch_linear_gelu(const void* x, const void* w, const void* bias, int m, int n, int k, long long x_row_stride, int gelu,
                       void* out, cudaStream_t st) {
  if (n % LG_BN != 0 || k % 8 != 0 || k < 8 || m < 1)
    return set_error(TOME_ERR_UNSUPPORTED, "tome_linear_gelu: needs n %% %d == 0 and k %% 8 == 0 (m=%d n=%d k=%d)", LG_BN, m, n, k);
  if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)out & 15) || (bias && ((uintptr_t)bias & 15)) || (x_row_stride % 8))
    return set_error(TOME_ERR_ALIGN, "tome_linear_gelu: 16-byte aligned tensors and row strides required");
  alignas(64) CUtensorMap map_a, map_w;
  int rc = make_bf16_map(&map_a, x, m, k, x_row_stride, LG_BM, "tome_linear_gelu");
  if (rc) return rc;
  rc = make_bf16_map(&map_w, w, n, k, k, LG_BN, "tome_linear_gelu");
  if (rc) return rc;
  alignas(64) CUtensorMap map_o;
  rc = make_bf16_map(&map_o, out, m, n, n, 32, "tome_linear_gelu", 32);
  if (rc) return rc;
  LinearGeluParams p;
  p.m = m; p.n = n; p.k = k; p.num_kb = (k + LG_BK - 1) / LG_BK;
  p.tiles_n = n / LG_BN; p.tiles = ((m + LG_BM - 1) / LG_BM) * p.tiles_n;
  p.bias = (const __nv_bfloat16*)bias; p.out = (__nv_bfloat16*)out; p.gelu = gelu;
  const size_t smem = (size_t)LG_STAGES * LG_STAGE_BYTES + LG_EPI_WARPS * LG_OBOX_BYTES + 16 * LG_STAGES + 32 + 16 + 1024;
  static PerDeviceOnce attr;
  if (attr.first_time())
    TOME_CUDA(cudaFuncSetAttribute(linear_gelu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
  linear_gelu_kernel<<<grid, LG_THREADS, smem, st>>>(map_a, map_w, map_o, p);
  TOME_LAUNCH_CHECK("linear_gelu_kernel");
  return TOME_OK;
}

}  // namespace tome
