// Kernel 1, tensor-core form: fused key-normalise + A.B^T cosine similarity + masked row
// max/argmax on tcgen05 / TMEM fed by TMA.  Replaces tome/merge.py:51-64 of the reference
// (norm, div, strided bmm that materialises a (bm, N/2, N/2) score tensor, two masked
// fills, max).  The score matrix never leaves TMEM.
//
// Three launches:
//   1. split_rows_kernel   one warp per token: fp64 sum of squares -> fp32 norm ->
//                          mhat = fp32(x / norm), written as hi (tf32-exact top bits) and
//                          lo = mhat - hi (exact) in the A-rows-then-B-rows layout TMA reads.
//   2. match_tc_kernel     per (128 A rows) x (BN <= 256 B rows) x batch tile: TMA
//                          (SWIZZLE_128B, K-major) -> 3xTF32 tcgen05.mma
//                          (hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM) ->
//                          epilogue straight out of TMEM: per row the approximate max and
//                          every column within a proven error window of it (<= KCAND, else
//                          "overflow").
//   3. refine_rows_kernel  per A row: exact canonical score (fp64 FMA over mhat, same order
//                          as match_exact.cu) of the few candidates -> node_max / node_idx.
// The tensor-core pass only prunes; every reported bit comes from step 3, so this path and
// the exact kernel agree bit for bit (tests/test_kernels_gpu.py).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace tome {

constexpr int TC_BM = 128;        // A rows per tile == UMMA M == TMEM lanes
constexpr int TC_BK = 32;         // fp32 per k-block == one 128-byte swizzle row
constexpr int TC_UK = 8;          // UMMA K for kind::tf32
constexpr int KCAND = 4;          // candidates kept per (row, column tile)
constexpr int TC_THREADS = 192;   // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int CNT_OVERFLOW = 255;

struct TcParams {
  int bm, n, na, nb, cm, cls, distill;
  int BN, n_ct, stages, num_kb, tmem_cols;
  int rows_total;          // bm * n : row offset of the "lo" half of the split buffer
  float window;            // 2 * error bound of the tensor-core pass
  float* tile_max;         // (bm, na, n_ct)            [streamed-K path only]
  int* tile_cnt;           // (bm, na, n_ct)
  int* tile_cand;          // (bm, na, n_ct, KCAND)
  int resident;            // all k-blocks stay in shared memory -> exact refine fused in the epilogue
  unsigned long long* keys;  // (bm, na) packed (score key, ~column), atomicMax across column tiles
  int* strip_count;        // (bm, row tiles) arrivals; last column tile of a strip decodes the keys
  float* node_max;         // (bm, na)
  int* node_idx;           // (bm, na)
  long long* trace;        // optional (TOME_TC_TRACE): per-CTA phase timestamps, 16 per CTA
};

// ---------------------------------------------------------------------------------------------
// 1. normalise + hi/lo split
// ---------------------------------------------------------------------------------------------
// `heads` > 1: the metric is the head-mean of K (tome/patch/videomae.py:72-73 `k.mean(1)`), taken here
// instead of in a separate reduction kernel: element (b, t, k) = mean_h keys[b, h, t, k], rounded to
// the input dtype first (what the reference's k.mean(1) tensor holds), then normalised.
template <typename T>
__global__ void __launch_bounds__(256) split_rows_kernel(const T* __restrict__ metric, View v, int heads, long long stride_h,
                                                         int bm, int n, int cm, float* __restrict__ split,
                                                         unsigned long long* __restrict__ keys,
                                                         int* __restrict__ strip_count, int n_strips) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= bm * n) return;
  const int b = warp / n, t = warp - b * n, na = na_of(n);
  if (lane == 0) {
    if (!(t & 1)) keys[(long long)b * na + (t >> 1)] = 0ull;
    if (warp < n_strips) strip_count[warp] = 0;
  }
  const T* src = metric + v.batch_offset(b) + (long long)t * v.sn;
  const int row = (t & 1) ? na + (t >> 1) : (t >> 1);
  float* hi = split + ((long long)b * n + row) * cm;
  float* lo = hi + (long long)bm * n * cm;
  auto value = [&](int k) -> float {
    if (heads == 1) return ld_as_float(src + k);
    float s = 0.f;                 // heads added in order; 1/H multiply like ATen's MeanOps
    int h = 0;
    for (; h + 4 <= heads; h += 4) {          // four loads in flight, adds kept sequential
      const float a0 = ld_as_float(src + (long long)(h + 0) * stride_h + k);
      const float a1 = ld_as_float(src + (long long)(h + 1) * stride_h + k);
      const float a2 = ld_as_float(src + (long long)(h + 2) * stride_h + k);
      const float a3 = ld_as_float(src + (long long)(h + 3) * stride_h + k);
      s = ((s + a0) + a1) + a2;
      s = s + a3;
    }
    for (; h < heads; ++h) s += ld_as_float(src + (long long)h * stride_h + k);
    s = s * (1.0f / (float)heads);
    if (sizeof(T) == 2) s = __bfloat162float(__float2bfloat16_rn(s));
    return s;
  };
  if (cm <= 128) {                 // whole row in registers: one pass over global memory
    float x[4];
    double ss = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = lane + 32 * q;
      x[q] = k < cm ? value(k) : 0.f;
      ss = fma((double)x[q], (double)x[q], ss);
    }
    ss = warp_sum(ss);
    const float norm = (float)sqrt(ss);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = lane + 32 * q;
      if (k < cm) {
        const float m = __fdiv_rn(x[q], norm);
        const float h = __uint_as_float(__float_as_uint(m) & 0xFFFFE000u);
        hi[k] = h;
        lo[k] = m - h;             // exact: fewer than 24 significant bits remain
      }
    }
    return;
  }
  double ss = 0.0;
  for (int k = lane; k < cm; k += 32) {
    const double x = (double)value(k);
    ss = fma(x, x, ss);
  }
  ss = warp_sum(ss);
  const float norm = (float)sqrt(ss);
  for (int k = lane; k < cm; k += 32) {
    const float m = __fdiv_rn(value(k), norm);
    const float h = __uint_as_float(__float_as_uint(m) & 0xFFFFE000u);
    hi[k] = h;
    lo[k] = m - h;
  }
}

// ---------------------------------------------------------------------------------------------
// PTX helpers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a lost arrival must abort the kernel (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE(slot) do { if (p.trace) p.trace[(((long long)b * gridDim.y + it) * gridDim.x + jt) * 16 + (slot)] = gtime(); } while (0)
__device__ __forceinline__ float fmax_nan(float a, float b) {     // NaN-propagating max
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// byte offset of the 16-byte chunk holding k..k+3 (k % 4 == 0, k < 32) of row r in a SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int k) { return (uint32_t)r * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)r & 7u)) << 4); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major),
//   [32,46) SBO >> 4 = 1024 B between 8-row groups, [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------
// 2. TMA -> tcgen05 -> TMEM epilogue
// ---------------------------------------------------------------------------------------------
// Exact canonical score of (tile row r, tile column c) from the SWIZZLE_128B stage buffers
// (all k-blocks resident): same fp64 FMA order as match_exact.cu / refine_rows_kernel.
__device__ __forceinline__ float exact_from_smem(uint32_t base, uint32_t stage_bytes, uint32_t a_bytes, uint32_t b_bytes,
                                                 int num_kb, int r, int c) {
  double acc = 0.0;
  for (int kb = 0; kb < num_kb; ++kb) {
    const uint32_t st = base + (uint32_t)kb * stage_bytes;
#pragma unroll
    for (int k = 0; k < TC_BK; k += 4) {
      const uint32_t oa = sw128_off(r, k), ob = sw128_off(c, k);
      const float4 ah = lds128(st + oa), al = lds128(st + a_bytes + oa);
      const float4 bh = lds128(st + 2u * a_bytes + ob), bl = lds128(st + 2u * a_bytes + b_bytes + ob);
      acc = fma((double)(ah.x + al.x), (double)(bh.x + bl.x), acc);
      acc = fma((double)(ah.y + al.y), (double)(bh.y + bl.y), acc);
      acc = fma((double)(ah.z + al.z), (double)(bh.z + bl.z), acc);
      acc = fma((double)(ah.w + al.w), (double)(bh.w + bl.w), acc);
    }
  }
  return (float)acc;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, it = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) TC_TRACE(0);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t a_bytes = TC_BM * 128u, b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = 2u * (a_bytes + b_bytes);
  const uint32_t bars = base + (uint32_t)p.stages * stage_bytes;    // full[stages] | empty[stages] | tmem_full | tmem_ptr
  const uint32_t bar_full = bars, bar_empty = bars + 8u * p.stages, bar_tmem = bars + 16u * p.stages;
  const uint32_t tmem_slot = bar_tmem + 8u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
  volatile int* is_last_ptr = reinterpret_cast<volatile int*>(gen_base + (tmem_slot + 4u - base));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TC_TRACE(1);

  const int row_a = b * p.n + it * TC_BM;               // rows of the split buffer (hi half)
  const int row_b = b * p.n + p.na + jt * p.BN;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(bar_empty + 8u * s, ((kb / p.stages) & 1) ^ 1);
        const uint32_t st = base + (uint32_t)s * stage_bytes, full = bar_full + 8u * s;
        mbar_expect_tx(full, stage_bytes);
        tma_load_2d(st, &map_a, kb * TC_BK, row_a, full);                                   // A hi
        tma_load_2d(st + a_bytes, &map_a, kb * TC_BK, row_a + p.rows_total, full);          // A lo
        tma_load_2d(st + 2u * a_bytes, &map_b, kb * TC_BK, row_b, full);                    // B hi
        tma_load_2d(st + 2u * a_bytes + b_bytes, &map_b, kb * TC_BK, row_b + p.rows_total, full);  // B lo
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(bar_full + 8u * s, (kb / p.stages) & 1);
        tc_fence_after();
        if (kb == 0) TC_TRACE(2);
        if (kb == p.num_kb - 1) TC_TRACE(3);
        const uint32_t st = base + (uint32_t)s * stage_bytes;
        const uint64_t a_hi = make_sw128_desc(st), a_lo = make_sw128_desc(st + a_bytes);
        const uint64_t b_hi = make_sw128_desc(st + 2u * a_bytes), b_lo = make_sw128_desc(st + 2u * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UK; ++k) {
          const uint64_t adv = (uint64_t)((k * TC_UK * 4) >> 4);     // +32 bytes inside the swizzle row
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
          umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        umma_commit(bar_empty + 8u * s);      // frees the stage when these MMAs retire
      }
      umma_commit(bar_tmem);                  // accumulator complete
      TC_TRACE(4);
    }
  } else {
    // ---- epilogue: thread <-> TMEM lane <-> A row ------------------------------------------
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane, i = it * TC_BM + row;
    const int j0 = jt * p.BN;
    const int ncol = min(p.BN, p.nb - j0);     // valid columns of this tile (>= 1)
    const bool mask0 = p.distill && jt == 0;   // column 0 is the distillation token
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TC_TRACE(5);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    // pass 1: approximate row max over the valid columns, remembering the max of every 32-column
    // chunk.  max.NaN propagates NaN, so a NaN anywhere (zero-norm row) surfaces in m.
    constexpr int MAXCH = 8;                   // BN <= 256
    float chmax[MAXCH];
#pragma unroll
    for (int q = 0; q < MAXCH; ++q) chmax[q] = -INFINITY;
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < MAXCH; ++q) {
      const int c = q * 32;
      if (c < ncol) {                          // warp-uniform
        float cm = -INFINITY;
        if (c + 32 <= ncol) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          if (mask0 && c == 0) v[0] = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) cm = fmax_nan(cm, v[e]);
        } else {
          for (int cc = c; cc < ncol; cc += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)cc, v);
            if (mask0 && cc == 0) v[0] = -INFINITY;
#pragma unroll
            for (int e = 0; e < 16; ++e) cm = fmax_nan(cm, (cc + e < ncol) ? v[e] : -INFINITY);
          }
        }
        chmax[q] = cm;
        m = fmax_nan(m, cm);
      }
    }
    const bool has_nan = (m != m);
    if (threadIdx.x == 64) TC_TRACE(6);
    // pass 2: every column within the error window of the max is a candidate.  Branch-free: one bit
    // per column (a divergent scan with a dynamically indexed candidate list cost 5.6 us per CTA,
    // see profiles/r01_match_notes.md), candidates are pulled out of the masks afterwards.
    const float thr = m - p.window;
    uint32_t bits[MAXCH];
#pragma unroll
    for (int q = 0; q < MAXCH; ++q) {
      const int c = q * 32;
      uint32_t bm_ = 0u;
      if (c < ncol && !has_nan) {              // warp-uniform
        if (c + 32 <= ncol) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          if (mask0 && c == 0) v[0] = -INFINITY;
          uint32_t part[4] = {0u, 0u, 0u, 0u};     // four independent OR chains (this phase is issue-latency bound)
#pragma unroll
          for (int e = 0; e < 32; ++e) part[e & 3] |= (v[e] >= thr ? 1u : 0u) << e;
          bm_ = (part[0] | part[1]) | (part[2] | part[3]);
        } else {
          for (int cc = c; cc < ncol; cc += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)cc, v);
            if (mask0 && cc == 0) v[0] = -INFINITY;
#pragma unroll
            for (int e = 0; e < 16; ++e) bm_ |= ((cc + e < ncol && v[e] >= thr) ? 1u : 0u) << (cc - c + e);
          }
        }
      }
      bits[q] = bm_;
    }
    if (threadIdx.x == 64) TC_TRACE(11);
    int cnt = 0, filled = 0, cand[KCAND];
#pragma unroll
    for (int e = 0; e < KCAND; ++e) cand[e] = 0;
#pragma unroll
    for (int q = 0; q < MAXCH; ++q) {
      uint32_t w = bits[q];
      cnt += __popc(w);
      while (w != 0u && filled < KCAND) {        // rare per (lane, chunk): a real branch beats a predicated chain
        const int col = q * 32 + __ffs(w) - 1;
        w &= w - 1u;
        if (filled == 0) cand[0] = col;
        else if (filled == 1) cand[1] = col;
        else if (filled == 2) cand[2] = col;
        else cand[3] = col;
        ++filled;
      }
    }
    if (threadIdx.x == 64) TC_TRACE(7);
    const bool overflow = has_nan || cnt > KCAND;
    if (!p.resident) {
      if (i < p.na) {
        const long long o = ((long long)b * p.na + i) * p.n_ct + jt;
        p.tile_max[o] = m;
        p.tile_cnt[o] = overflow ? CNT_OVERFLOW : cnt;
#pragma unroll
        for (int s = 0; s < KCAND; ++s) p.tile_cand[o * KCAND + s] = j0 + cand[s];
      }
    } else if (i < p.na) {
      // exact refine straight from the operand tiles still sitting in shared memory
      for (int s = 0; s < p.num_kb; ++s) mbar_wait(bar_full + 8u * s, 0);     // acquire the TMA writes
      unsigned long long best = 0ull;
      if (p.cls && i == 0) {
        best = pack_best(-INFINITY, 0);
      } else if (overflow) {
        for (int cc = 0; cc < ncol; ++cc) {
          const float sc = (mask0 && cc == 0) ? -INFINITY : exact_from_smem(base, stage_bytes, a_bytes, b_bytes, p.num_kb, row, cc);
          const unsigned long long k = pack_best(sc, j0 + cc);
          best = k > best ? k : best;
        }
      } else {
        for (int s = 0; s < cnt; ++s) {
          const float sc = exact_from_smem(base, stage_bytes, a_bytes, b_bytes, p.num_kb, row, cand[s]);
          const unsigned long long k = pack_best(sc, j0 + cand[s]);
          best = k > best ? k : best;
        }
      }
      atomicMax(p.keys + (long long)b * p.na + i, best);
    }
  }
  if (threadIdx.x == 64) TC_TRACE(8);
  tc_fence_before();
  if (p.resident) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(9);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
  if (p.resident) {
    // last column tile of this (batch, row strip) turns the packed keys into node_max / node_idx
    if (threadIdx.x == 0) {
      const int strip = b * gridDim.y + it;
      const int old = atomicAdd(p.strip_count + strip, 1);
      *is_last_ptr = (old == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (*is_last_ptr) {
      __threadfence();
      const int i = it * TC_BM + (int)threadIdx.x;
      if (threadIdx.x < TC_BM && i < p.na) {
        const long long o = (long long)b * p.na + i;
        const unsigned long long k = __ldcg(p.keys + o);
        p.node_max[o] = key_to_float((uint32_t)(k >> 32));
        p.node_idx[o] = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
      }
    }
  }
  if (threadIdx.x == 0) TC_TRACE(10);
}

// ---------------------------------------------------------------------------------------------
// 3. exact refine (streamed-K path): canonical fp64 score of the surviving candidates
// ---------------------------------------------------------------------------------------------
// One WARP per A row: lanes stride k, so every candidate costs four coalesced row reads (hi/lo of A and
// B) and a warp reduction.  (The first version used one thread per row: at Cm = 768 its uncoalesced,
// serial 768-step loops took 400 us.)  fp64 partial sums per lane, then a butterfly: the order of the
// fp64 additions is not part of the score definition (it can move the fp32-rounded result only when
// the fp64 sum lies within ~1e-16 of a rounding boundary).
__device__ __forceinline__ float exact_score_warp(const float* __restrict__ ah, const float* __restrict__ al,
                                                  const float* __restrict__ bh, const float* __restrict__ bl, int cm, int lane) {
  double acc = 0.0;
  for (int k = lane; k < cm; k += 32) acc = fma((double)(ah[k] + al[k]), (double)(bh[k] + bl[k]), acc);
  acc = warp_sum(acc);
  return (float)acc;
}

__global__ void __launch_bounds__(256) refine_rows_kernel(const float* __restrict__ split, TcParams p,
                                                          float* __restrict__ node_max, int* __restrict__ node_idx) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= p.bm * p.na) return;
  const int b = t / p.na, i = t - b * p.na;
  if (p.cls && i == 0) { if (lane == 0) { node_max[t] = -INFINITY; node_idx[t] = 0; } return; }
  const long long lo_off = (long long)p.rows_total * p.cm;
  const float* ah = split + ((long long)b * p.n + i) * p.cm;
  const float* bbase = split + ((long long)b * p.n + p.na) * p.cm;
  const long long o = (long long)t * p.n_ct;
  float big = -INFINITY;
  for (int c = 0; c < p.n_ct; ++c) big = fmaxf(big, p.tile_max[o + c]);
  const float thr = big - p.window;
  unsigned long long best = 0ull;
  for (int c = 0; c < p.n_ct; ++c) {
    const int cnt = p.tile_cnt[o + c];
    if (cnt == CNT_OVERFLOW) {            // too many near-ties (or NaN): score the whole column tile exactly
      const int je = min(p.nb, (c + 1) * p.BN);
      for (int j = c * p.BN; j < je; ++j) {
        const float s = (p.distill && j == 0) ? -INFINITY
                                              : exact_score_warp(ah, ah + lo_off, bbase + (long long)j * p.cm, bbase + (long long)j * p.cm + lo_off, p.cm, lane);
        const unsigned long long k = pack_best(s, j);
        best = k > best ? k : best;
      }
    } else if (p.tile_max[o + c] >= thr) {
      for (int s = 0; s < cnt; ++s) {
        const int j = p.tile_cand[(o + c) * KCAND + s];
        const float sc = exact_score_warp(ah, ah + lo_off, bbase + (long long)j * p.cm, bbase + (long long)j * p.cm + lo_off, p.cm, lane);
        const unsigned long long k = pack_best(sc, j);
        best = k > best ? k : best;
      }
    }
  }
  if (lane == 0) {
    node_max[t] = key_to_float((uint32_t)(best >> 32));
    node_idx[t] = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static void tc_geometry(int bm, int n, int cm, TcParams& p) {
  p.bm = bm; p.n = n; p.na = na_of(n); p.nb = nb_of(n); p.cm = cm;
  p.num_kb = (cm + TC_BK - 1) / TC_BK;
  int n_ct = (p.nb + 255) / 256;
  const int forced = env_int("TOME_TC_NCT", 0);          // tuning knob (bench sweeps)
  if (forced > 0) n_ct = forced > (p.nb + 15) / 16 ? (p.nb + 15) / 16 : forced;
  if (n_ct < (p.nb + 255) / 256) n_ct = (p.nb + 255) / 256;
  p.n_ct = n_ct;
  int bn = (p.nb + p.n_ct - 1) / p.n_ct;
  bn = (bn + 15) & ~15;
  p.BN = bn < 16 ? 16 : bn;
  p.n_ct = (p.nb + p.BN - 1) / p.BN;
  const int stage_bytes = 2 * (TC_BM + p.BN) * 128;
  int st = (216 * 1024) / stage_bytes;
  st = st > 4 ? 4 : st;
  p.stages = st > p.num_kb ? p.num_kb : st;
  p.resident = (p.num_kb <= p.stages) && !env_int("TOME_TC_NO_FUSED_REFINE", 0);
  p.tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  // error bound of hi.hi + hi.lo + lo.hi with fp32 accumulation on unit vectors:
  // 3 * 2^-20 (dropped lo.lo + truncated lo) + (3 cm / 8) accumulations * 2^-22, with margin
  const float eps = 4e-6f + 2e-7f * (float)cm;
  p.window = 2.0f * eps;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t match_tc_workspace(int bm, int n, int cm) {
  TcParams p;
  tc_geometry(bm, n, cm, p);
  const size_t rows = (size_t)bm * p.na * p.n_ct;
  const size_t strips = (size_t)bm * ((p.na + TC_BM - 1) / TC_BM);
  return align256((size_t)2 * bm * n * cm * sizeof(float)) + align256(rows * 4) + align256(rows * 4) +
         align256(rows * 4 * KCAND) + align256((size_t)bm * p.na * 8) + align256(strips * 4);
}

bool match_tc_supported(int dtype, int bm, int n, int cm, const View& v, const void* metric) {
  (void)metric; (void)v;
  if (dtype != TOME_F32 && dtype != TOME_BF16) return false;
  if (cm % 4 != 0 || cm < 4 || cm > 4096) return false;          // TMA: 16-byte row pitch
  if (n < 2 || (long long)bm * n * 2 > 0x7fffffffLL) return false;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int make_map(CUtensorMap* map, float* base, int rows, int cm, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(TOME_ERR_CUDA, "tome_match: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cm, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cm * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TOME_ERR_CUDA, "tome_match: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return TOME_OK;
}

int launch_match_tc(const void* metric, int dtype, int bm, int n, int cm, const View& v, int cls, int distill,
                    float* node_max, int* node_idx, void* ws, size_t ws_bytes, cudaStream_t st, int heads,
                    long long stride_h) {
  if (ws_bytes < match_tc_workspace(bm, n, cm))
    return set_error(TOME_ERR_WORKSPACE, "tome_match: workspace %zu < %zu bytes", ws_bytes, match_tc_workspace(bm, n, cm));
  TcParams p;
  tc_geometry(bm, n, cm, p);
  p.cls = cls; p.distill = distill; p.rows_total = bm * n;
  p.node_max = node_max; p.node_idx = node_idx;
  p.trace = getenv("TOME_TC_TRACE") ? (long long*)strtoull(getenv("TOME_TC_TRACE"), nullptr, 0) : nullptr;
  const int n_rt = (p.na + TC_BM - 1) / TC_BM;
  char* w = (char*)ws;
  float* split = (float*)w;                      w += align256((size_t)2 * bm * n * cm * sizeof(float));
  const size_t rows = (size_t)bm * p.na * p.n_ct;
  p.tile_max = (float*)w;                        w += align256(rows * 4);
  p.tile_cnt = (int*)w;                          w += align256(rows * 4);
  p.tile_cand = (int*)w;                         w += align256(rows * 4 * KCAND);
  p.keys = (unsigned long long*)w;               w += align256((size_t)bm * p.na * 8);
  p.strip_count = (int*)w;

  const int blocks = (int)(((long long)bm * n * 32 + 255) / 256);
  if (dtype == TOME_F32)
    split_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)metric, v, heads, stride_h, bm, n, cm, split, p.keys, p.strip_count, bm * n_rt);
  else
    split_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)metric, v, heads, stride_h, bm, n, cm, split, p.keys, p.strip_count, bm * n_rt);
  TOME_LAUNCH_CHECK("split_rows_kernel");

  alignas(64) CUtensorMap map_a, map_b;
  int rc = make_map(&map_a, split, 2 * bm * n, cm, TC_BM);
  if (rc) return rc;
  rc = make_map(&map_b, split, 2 * bm * n, cm, p.BN);
  if (rc) return rc;
  const size_t smem = (size_t)p.stages * 2 * (TC_BM + p.BN) * 128 + 16 * p.stages + 16 + 1024;
  static bool smem_set = false;
  if (!smem_set) {
    TOME_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    smem_set = true;
  }
  dim3 grid(p.n_ct, n_rt, bm);
  match_tc_kernel<<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
  TOME_LAUNCH_CHECK("match_tc_kernel");
  if (!p.resident) {
    const long long total = (long long)bm * p.na * 32;
    refine_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(split, p, node_max, node_idx);
    TOME_LAUNCH_CHECK("refine_rows_kernel");
  }
  return TOME_OK;
}

}  // namespace tome
