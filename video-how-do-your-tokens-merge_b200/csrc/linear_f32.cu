// fp32 linear layers on the tensor cores, at fp32 accuracy (the GEMMs either side of the merge).
//
// The reference benchmark runs the models in fp32 with TF32 off (slowfast/utils/model_benchmark.py:21-45), and so does
// the headline line of bench.py.  There the library GEMMs are CUDA-core SGEMMs (cutlass simt_sgemm, ~65 TFLOP/s on a
// B200) and take 83 % of the patched VideoMAE step (profiles/r02_videomae_fp32_launches.csv); the merge path is 3 %.
// tcgen05 has no fp32 input kind, but an fp32 number is EXACTLY the sum of three bf16 numbers
//     x = h + m + l,   h = bf16(x), m = bf16(x - h), l = bf16(x - h - m)        (8 + 8 + 8 significand bits)
// and every bf16 x bf16 product is exact in the fp32 accumulator, so
//     x . w = sum over the nine plane pairs (h, m, l) x (h, m, l)
// with fp32 accumulation in TMEM reproduces an fp32 GEMM to fp32 rounding (the scheme cuBLAS ships as
// "BF16x9 FP32 emulation"): no TF32.  Measured against fp64 the result is as close as the SIMT SGEMM
// (tests/test_kernels_gpu.py::test_linear_f32_matches_fp64).  `terms` picks the products: 9 = all; 8 = without l.l, which is
// <= 2^-32 of |x||w| -- 2^-8 of the accumulator's own rounding unit, so it vanishes in the accumulation anyway: against the
// nine-product result the outputs agree to within one fp32 epsilon of the output scale (measured 0.3-0.6) and their distance from fp64 is the same to
// four digits (test_linear_f32_eight_products_equal_nine_at_fp32_resolution); 7 % faster, the host mirror's default; 6 = also
// without m.l and l.m (<= 2^-23 per product), a knob.
//
//   split3_kernel        x (rows, k) fp32 -> (rows, 3k) bf16 planes [h | m | l]; weights are split once and cached
//   linear_f32_kernel    persistent CTAs, 128 x 256 output tiles; per 32-channel k-block the producer warp TMA-loads the
//                        three A planes and the three W planes (72 KB, SWIZZLE_64B, 3-stage ring), the MMA warp issues
//                        2 x 9 tcgen05.mma kind::f16 into one of two TMEM accumulators.  The tensor core's fp32
//                        accumulator does not round like an FMA chain: the error of a single long accumulation grows
//                        linearly with k (4x the SGEMM's at k = 3072).  So an accumulator only ever holds 256 channels
//                        (8 k-blocks, 144 MMAs); the eight epilogue warps pull each finished chunk out of TMEM and add it,
//                        round-to-nearest, to the tile's running sum in registers while the next chunk accumulates in the
//                        other TMEM buffer; bias (and the erf GELU for fc1) at the end, fp32 rows stored directly.
#include <math.h>

#include "tc_ptx.cuh"

namespace tome {

constexpr int LF_BM = 128, LF_BN = 256, LF_BK = 32, LF_STAGES = 3;
constexpr int LF_EPI_WARPS = 8;                       // two per TMEM lane quarter, 128 columns each
constexpr int LF_CHUNK_KB = 8;                        // k-blocks (256 channels) accumulated in TMEM before the sum moves to registers
constexpr int LF_THREADS = 64 + 32 * LF_EPI_WARPS;
constexpr uint32_t LF_A_BYTES = LF_BM * 64u, LF_B_BYTES = LF_BN * 64u;            // one plane of one stage (64-byte rows)
constexpr uint32_t LF_STAGE_BYTES = 3u * (LF_A_BYTES + LF_B_BYTES);

struct LinearF32Params {
  int m, n, k, num_kb, tiles_n, tiles, terms;       // terms: 9 (every product), 8 (without l.l: <= 2^-32 relative) or 6 (without m.l, l.m, l.l: <= 2^-23)
  const float* bias;                                // (n) or NULL
  float* out;                                       // (m, n) row-major, or NULL
  __nv_bfloat16* out3;                              // (m, 3n) bf16 planes of the same values (the next GEMM's operand), or NULL
  int gelu;                                         // 0: bias only; 1: erf GELU (nn.GELU); 2: HF "gelu_fast" (tanh form)
};

// ---- split -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, long long row_stride, long long rows, int k,
                                                     __nv_bfloat16* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 channels
  const int k4 = k >> 2;
  if (idx >= rows * k4) return;
  const long long r = idx / k4;
  const int c = (int)(idx - r * k4) * 4;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * row_stride + c));
  const float f[4] = {v.x, v.y, v.z, v.w};
  uint32_t hw[2], mw[2], lw[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float h[2], mm[2], ll[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float xx = f[2 * i + e];
      h[e] = __bfloat162float(__float2bfloat16_rn(xx));
      const float r1 = xx - h[e];                                               // exact
      mm[e] = __bfloat162float(__float2bfloat16_rn(r1));
      ll[e] = r1 - mm[e];                                                       // exact; fits bf16 (<= 8 significant bits left)
    }
    const __nv_bfloat162 hh = __floats2bfloat162_rn(h[0], h[1]), m2 = __floats2bfloat162_rn(mm[0], mm[1]),
                         l2 = __floats2bfloat162_rn(ll[0], ll[1]);
    hw[i] = *reinterpret_cast<const uint32_t*>(&hh);
    mw[i] = *reinterpret_cast<const uint32_t*>(&m2);
    lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  __nv_bfloat16* o = out + r * 3LL * k + c;
  *reinterpret_cast<uint2*>(o) = make_uint2(hw[0], hw[1]);
  *reinterpret_cast<uint2*>(o + k) = make_uint2(mw[0], mw[1]);
  *reinterpret_cast<uint2*>(o + 2LL * k) = make_uint2(lw[0], lw[1]);
}

// K-major, SWIZZLE_64B descriptor: 64-byte rows, 8-row groups 512 bytes apart, layout type 4.
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

__global__ void __launch_bounds__(LF_THREADS, 1)
linear_f32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const LinearF32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + LF_STAGES * LF_STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * LF_STAGES;
  const uint32_t bar_tfull = bars + 16u * LF_STAGES, bar_tempty = bar_tfull + 16u;
  const uint32_t tmem_slot = bar_tempty + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_w);
    for (int s = 0; s < LF_STAGES; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, LF_EPI_WARPS); }
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

  // Producer and MMA warps run converged with one elected lane issuing ("Issuing tcgen05.mma"): under
  // `if (lane == 0)` every TMA / tcgen05.mma issue sits in an ELECT / vote loop of ~70 cycles plus a descriptor built in
  // ordinary registers -- 16 MMAs a stage came to ~1600 of the stage's 2048 tensor cycles, close enough for every barrier poll to
  // open a gap in the tensor pipe.
  if (warp == 0) {
    const bool leader = elect_one_sync();
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const int mt = t / p.tiles_n, nt = t - mt * p.tiles_n;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const uint32_t s = it % LF_STAGES, ph = (it / LF_STAGES) & 1u;
        mbar_wait_sleep(bar_empty + 8u * s, ph ^ 1u, 64);
        if (leader) {
          const uint32_t st = base + s * LF_STAGE_BYTES, full = bar_full + 8u * s;
          mbar_expect_tx(full, LF_STAGE_BYTES);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            tma_load_2d(st + pl * LF_A_BYTES, &map_a, pl * p.k + kb * LF_BK, mt * LF_BM, full);
            tma_load_2d(st + 3u * LF_A_BYTES + pl * LF_B_BYTES, &map_w, pl * p.k + kb * LF_BK, nt * LF_BN, full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LF_BN >> 3) << 17) | ((uint32_t)(LF_BM >> 4) << 24);
    const uint64_t da0 = make_sw64_desc(base), db0 = make_sw64_desc(base + 3u * LF_A_BYTES);   // stage 0, plane h; planes / stages
    constexpr uint64_t PA = LF_A_BYTES >> 4, PB = LF_B_BYTES >> 4, ST = LF_STAGE_BYTES >> 4;  // further are constants in the address field
    const int terms = p.terms;
    uint32_t it = 0, cl = 0;                                   // k-block and chunk counters across tiles
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      for (int kb0 = 0; kb0 < p.num_kb; kb0 += LF_CHUNK_KB, ++cl) {
        const uint32_t acc = cl & 1u, aph = (cl >> 1) & 1u;
        mbar_wait_sleep(bar_tempty + 8u * acc, aph ^ 1u, 64);   // the epilogue has pulled the previous chunk out of this buffer
        tc_fence_after();
        const uint32_t d_tmem = tb + acc * (uint32_t)LF_BN;
        const int kb1 = min(p.num_kb, kb0 + LF_CHUNK_KB);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % LF_STAGES, ph = (it / LF_STAGES) & 1u;
          mbar_wait_sleep(bar_full + 8u * s, ph, 20);
          tc_fence_after();
          if (leader) {
            const uint64_t da = da0 + (uint64_t)s * ST, db = db0 + (uint64_t)s * ST;
            uint32_t first = kb == kb0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < LF_BK / 16; ++k) {
              const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);        // +32 bytes inside the 64-byte swizzle row
              // smallest products first: (l, l) ... (h, h)
#pragma unroll
              for (int i = 2; i >= 0; --i) {
#pragma unroll
                for (int j = 2; j >= 0; --j) {
                  if ((terms == 6 && i + j >= 3) || (terms == 8 && i + j == 4)) continue;   // 6: m.l, l.m, l.l; 8: l.l
                  umma_bf16(d_tmem, da + (uint64_t)i * PA + adv, db + (uint64_t)j * PB + adv, idesc, first ? 0u : 1u);
                  first = 0u;
                }
              }
            }
            umma_commit(bar_empty + 8u * s);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bar_tfull + 8u * acc);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                  // which 128 of the tile's 256 columns
    uint32_t cl = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const int mt = t / p.tiles_n, nt = t - mt * p.tiles_n;
      float sum[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 32; ++e) sum[c][e] = 0.f;
      for (int kb0 = 0; kb0 < p.num_kb; kb0 += LF_CHUNK_KB, ++cl) {
        const uint32_t acc = cl & 1u, aph = (cl >> 1) & 1u;
        mbar_wait_sleep(bar_tfull + 8u * acc, aph, 128);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * (uint32_t)LF_BN + (uint32_t)(part * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          float va[32], vb[32];
          tmem_ld32_nowait(taddr + (uint32_t)(32 * c), va);
          tmem_ld32_nowait(taddr + (uint32_t)(32 * c + 32), vb);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 32; ++e) { sum[c][e] += va[e]; sum[c + 1][e] += vb[e]; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8u * acc) : "memory");
      }
      const int row = mt * LF_BM + q * 32 + lane, col0 = nt * LF_BN + part * 128;
      if (row < p.m) {
        float4* orow = p.out ? reinterpret_cast<float4*>(p.out + (long long)row * p.n + col0) : nullptr;
        __nv_bfloat16* prow = p.out3 ? p.out3 + (long long)row * 3 * p.n + col0 : nullptr;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int g = 0; g < 8; g += 2) {
            float o[8];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c * 32 + (g + u) * 4));
              o[4 * u] = sum[c][4 * (g + u)] + b4.x; o[4 * u + 1] = sum[c][4 * (g + u) + 1] + b4.y;
              o[4 * u + 2] = sum[c][4 * (g + u) + 2] + b4.z; o[4 * u + 3] = sum[c][4 * (g + u) + 3] + b4.w;
            }
            if (p.gelu == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = 0.5f * o[e] * (1.0f + erff(o[e] * 0.70710678118654752440f));     // nn.GELU, exact erf
            } else if (p.gelu == 2) {                       // HuggingFace FastGELUActivation (ViViT's hidden_act), libdevice tanhf
#pragma unroll
              for (int e = 0; e < 8; ++e)
                o[e] = 0.5f * o[e] * (1.0f + tanhf(o[e] * 0.7978845608f * (1.0f + 0.044715f * o[e] * o[e])));
            }
            if (orow) { orow[c * 8 + g] = make_float4(o[0], o[1], o[2], o[3]); orow[c * 8 + g + 1] = make_float4(o[4], o[5], o[6], o[7]); }
            if (prow) store_planes8(prow + c * 32 + g * 4, p.n, o);       // the next GEMM's operand, no fp32 round trip
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// The fp32 value of a column slice of a planes tensor: out (rows, ncols) = h + m + l, exact (h + m and (h + m) + l are both
// representable).  The QKV GEMM hands its result over as planes ONLY -- writing the fp32 tensor as well from the same epilogue
// costs more than this pass over the K third that the matching metric needs (329 us against 251 + 15 at 12544 x 2304:
// tools/bench_linear_f32_outputs.py).
__global__ void __launch_bounds__(256) planes_sum_kernel(const __nv_bfloat16* __restrict__ x3, long long rows, int n, int col0, int ncols,
                                                        float* __restrict__ out) {
  const int per_row = ncols >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * per_row) return;
  const long long r = gid / per_row;
  const int c = (int)(gid - r * per_row) * 8;
  const __nv_bfloat16* src = x3 + r * 3 * n + col0 + c;
  const uint4 vh = __ldg(reinterpret_cast<const uint4*>(src)), vm = __ldg(reinterpret_cast<const uint4*>(src + n)),
              vl = __ldg(reinterpret_cast<const uint4*>(src + 2 * n));
  const uint32_t wh[4] = {vh.x, vh.y, vh.z, vh.w}, wm[4] = {vm.x, vm.y, vm.z, vm.w}, wl[4] = {vl.x, vl.y, vl.z, vl.w};
  float f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = (__uint_as_float(wh[i] << 16) + __uint_as_float(wm[i] << 16)) + __uint_as_float(wl[i] << 16);
    f[2 * i + 1] = (__uint_as_float(wh[i] & 0xFFFF0000u) + __uint_as_float(wm[i] & 0xFFFF0000u)) + __uint_as_float(wl[i] & 0xFFFF0000u);
  }
  float4* dst = reinterpret_cast<float4*>(out + r * ncols + c);
  dst[0] = make_float4(f[0], f[1], f[2], f[3]);
  dst[1] = make_float4(f[4], f[5], f[6], f[7]);
}

int launch_planes_sum(const void* x3, long long rows, int n, int col0, int ncols, void* out, cudaStream_t st) {
  if (n % 8 || col0 % 8 || ncols % 8 || col0 < 0 || ncols < 8 || col0 + ncols > n)
    return set_error(TOME_ERR_ARG, "tome_planes_sum: columns %d..%d of %d (multiples of 8 inside the row)", col0, col0 + ncols, n);
  if (((uintptr_t)x3 & 15) || ((uintptr_t)out & 15)) return set_error(TOME_ERR_ALIGN, "tome_planes_sum: buffers must be 16-byte aligned");
  const long long total = rows * (ncols >> 3);
  planes_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)x3, rows, n, col0, ncols, (float*)out);
  TOME_LAUNCH_CHECK("planes_sum_kernel");
  return TOME_OK;
}

// ---- host -----------------------------------------------------------------------------------------------------
int launch_split3(const void* x, long long rows, int k, long long row_stride, void* out, cudaStream_t st) {
  if (k % 4 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)out & 7) || row_stride % 4)
    return set_error(TOME_ERR_ALIGN, "tome_split3: needs k %% 4 == 0 and 16-byte aligned rows");
  const long long total = rows * (k / 4);
  split3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)x, row_stride, rows, k, (__nv_bfloat16*)out);
  TOME_LAUNCH_CHECK("split3_kernel");
  return TOME_OK;
}

int launch_linear_f32(const void* x3, const void* w3, const void* bias, int m, int n, int k, int gelu, int terms, void* out,
                      void* out3, cudaStream_t st) {
  if (n % LF_BN != 0 || k % LF_BK != 0 || m < 1)
    return set_error(TOME_ERR_UNSUPPORTED, "tome_linear_f32: needs n %% %d == 0 and k %% %d == 0 (m=%d n=%d k=%d)", LF_BN, LF_BK, m, n, k);
  if (!out && !out3) return set_error(TOME_ERR_ARG, "tome_linear_f32: no output");
  if (((uintptr_t)x3 & 15) || ((uintptr_t)w3 & 15) || ((uintptr_t)out & 15) || ((uintptr_t)out3 & 15) || (bias && ((uintptr_t)bias & 15)))
    return set_error(TOME_ERR_ALIGN, "tome_linear_f32: 16-byte aligned tensors required");
  if (terms != 6 && terms != 8 && terms != 9) return set_error(TOME_ERR_ARG, "tome_linear_f32: terms must be 6, 8 or 9");
  alignas(64) CUtensorMap map_a, map_w;
  int rc = make_bf16_map(&map_a, x3, m, 3LL * k, 3LL * k, LF_BM, "tome_linear_f32", LF_BK);
  if (rc) return rc;
  rc = make_bf16_map(&map_w, w3, n, 3LL * k, 3LL * k, LF_BN, "tome_linear_f32", LF_BK);
  if (rc) return rc;
  LinearF32Params p;
  p.m = m; p.n = n; p.k = k; p.num_kb = k / LF_BK; p.terms = terms;
  p.tiles_n = n / LF_BN; p.tiles = ((m + LF_BM - 1) / LF_BM) * p.tiles_n;
  p.bias = (const float*)bias; p.out = (float*)out; p.out3 = (__nv_bfloat16*)out3; p.gelu = gelu;
  const size_t smem = (size_t)LF_STAGES * LF_STAGE_BYTES + 16 * LF_STAGES + 32 + 16 + 1024;
  static PerDeviceOnce attr;
  if (attr.first_time())
    TOME_CUDA(cudaFuncSetAttribute(linear_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
  linear_f32_kernel<<<grid, LF_THREADS, smem, st>>>(map_a, map_w, p);
  TOME_LAUNCH_CHECK("linear_f32_kernel");
  return TOME_OK;
}

}  // namespace tome
