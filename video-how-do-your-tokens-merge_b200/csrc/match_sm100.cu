// Kernel 1, tensor-core form: fused key-normalise + A.B^T cosine similarity + masked row
// max/argmax on tcgen05 / TMEM fed by TMA.  Replaces tome/merge.py:51-64 of the reference
// (norm, div, strided bmm that materialises a (bm, N/2, N/2) score tensor, two masked
// fills, max).  The score matrix never leaves TMEM.
//
//   1. split_rows_kernel   one warp per token: (head-mean of K,) fp64 sum of squares -> fp32 norm ->
//                          mhat = fp32(x / norm), written as fp32 (the exact operand) and as a two-term
//                          bf16 split h = bf16(mhat), m = bf16(mhat - h) in the A-rows-then-B-rows
//                          layout TMA reads.
//   2. match_tc_kernel     per (128 A rows) x (BN <= 256 B rows) x batch tile: TMA (SWIZZLE_128B, K-major)
//                          -> tcgen05.mma kind::f16 on the bf16 split (h.h + h.m + m.h, fp32 accumulate
//                          in TMEM: an approximation with a PROVEN error bound) -> epilogue straight out
//                          of TMEM: per row the approximate max and every column within the error window
//                          of it (<= KCAND, else "overflow"), then the EXACT canonical score (fp64 FMA
//                          over the fp32 mhat rows) of those few candidates, atomicMax of the packed
//                          (score, ~column) key across column tiles.
//   3. refine_rows_kernel  wide metrics only (cm > 128): the exact scoring as its own warp-per-row pass.
// The tensor-core pass only prunes; every reported bit comes from the exact scoring, so this path and
// match_exact.cu agree bit for bit (tests/test_kernels_gpu.py).
//
// Why bf16 x 3 and not tf32 x 3 (the first version, profiles/r01_match_notes.md): half the operand bytes,
// one 128-byte swizzle row per 64 channels (Cm = 64 is ONE k-block), so a (128 x 112) tile is 60 KB of
// shared memory and 128 TMEM columns -- three CTAs per SM, the whole layer-0 grid (392 CTAs) resident at
// once instead of two waves of 172 KB CTAs.  The pruning window is 3.5x wider, still ~1 candidate/row.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace tome {

constexpr int TC_BM = 128;        // A rows per tile == UMMA M == TMEM lanes
constexpr int TC_BK = 64;         // bf16 per k-block == one 128-byte swizzle row
constexpr int TC_UK = 16;         // UMMA K for kind::f16
constexpr int KCAND = 4;          // candidates kept per (row, column tile)
constexpr int TC_THREADS = 192;   // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int CNT_OVERFLOW = 255;

struct TcParams {
  int bm, n, na, nb, cm, cls, distill;
  int BN, n_ct, stages, num_kb, tmem_cols;
  int rows_total;          // bm * n : row offset of the "m" plane of the bf16 split buffer
  float window;            // 2 * error bound of the tensor-core pass
  float* tile_max;         // (bm, na, n_ct)            [separate-refine path only]
  int* tile_cnt;           // (bm, na, n_ct)
  int* tile_cand;          // (bm, na, n_ct, KCAND)
  int fused_refine;        // exact scoring of the candidates in the epilogue (cm <= 128)
  const float* mhat;       // (bm * n, cm) fp32 normalised rows, A rows then B rows per batch element
  unsigned int* approx;    // (bm, na) orderable approximate row max over the tiles published so far (filter)
  unsigned long long* keys;  // (bm, na) packed (exact score key, ~column), atomicMax across column tiles
  int* strip_count;        // (bm, row tiles) arrivals; last column tile of a strip decodes the keys
  float* node_max;         // (bm, na)
  int* node_idx;           // (bm, na)
  long long* trace;        // optional (TOME_TC_TRACE): per-CTA phase timestamps, 16 per CTA
};

// ---------------------------------------------------------------------------------------------
// 1. normalise + split
// ---------------------------------------------------------------------------------------------
// `heads` > 1: the metric is the head-mean of K (tome/patch/videomae.py:72-73 `k.mean(1)`), taken here
// instead of in a separate reduction kernel: element (b, t, k) = mean_h keys[b, h, t, k], rounded to
// the input dtype first (what the reference's k.mean(1) tensor holds), then normalised.
// Each lane owns channel PAIRS (2 lane + 64 q): one 4-byte (bf16) / 8-byte (fp32) load per head covers
// a 64-channel head row per warp instruction.
__device__ __forceinline__ float2 ld_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* p) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

template <typename T>
__device__ __forceinline__ float2 metric_pair(const T* src, int heads, long long stride_h, int k) {
  if (heads == 1) return ld_pair(src + k);
  float2 s = make_float2(0.f, 0.f);       // heads added in order; 1/H multiply like ATen's MeanOps
  int h = 0;
  for (; h + 12 <= heads; h += 12) {      // ViT-B's 12 heads: every load in flight before the first add
    float2 a[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) a[u] = ld_pair(src + (long long)(h + u) * stride_h + k);
#pragma unroll
    for (int u = 0; u < 12; ++u) { s.x += a[u].x; s.y += a[u].y; }      // adds stay sequential in h
  }
  for (; h + 4 <= heads; h += 4) {        // four loads in flight, adds kept sequential
    const float2 a0 = ld_pair(src + (long long)(h + 0) * stride_h + k);
    const float2 a1 = ld_pair(src + (long long)(h + 1) * stride_h + k);
    const float2 a2 = ld_pair(src + (long long)(h + 2) * stride_h + k);
    const float2 a3 = ld_pair(src + (long long)(h + 3) * stride_h + k);
    s.x = (((s.x + a0.x) + a1.x) + a2.x) + a3.x;
    s.y = (((s.y + a0.y) + a1.y) + a2.y) + a3.y;
  }
  for (; h < heads; ++h) {
    const float2 a = ld_pair(src + (long long)h * stride_h + k);
    s.x += a.x; s.y += a.y;
  }
  const float inv = 1.0f / (float)heads;
  s.x *= inv; s.y *= inv;
  if (sizeof(T) == 2) {
    s.x = __bfloat162float(__float2bfloat16_rn(s.x));
    s.y = __bfloat162float(__float2bfloat16_rn(s.y));
  }
  return s;
}

__device__ __forceinline__ void store_split(float2 m, __nv_bfloat16* h_row, __nv_bfloat16* m_row, float* f_row, int k) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(m.x, m.y);
  const float2 hf = __bfloat1622float2(h);
  *reinterpret_cast<__nv_bfloat162*>(h_row + k) = h;
  *reinterpret_cast<__nv_bfloat162*>(m_row + k) = __floats2bfloat162_rn(m.x - hf.x, m.y - hf.y);   // differences are exact
  *reinterpret_cast<float2*>(f_row + k) = m;
}

template <typename T>
__global__ void __launch_bounds__(256) split_rows_kernel(const T* __restrict__ metric, View v, int heads, long long stride_h,
                                                         int bm, int n, int cm, __nv_bfloat16* __restrict__ hm,
                                                         float* __restrict__ mhat, unsigned long long* __restrict__ keys,
                                                         unsigned int* __restrict__ approx, int* __restrict__ strip_count,
                                                         int n_strips) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= bm * n) return;
  const int b = warp / n, t = warp - b * n, na = na_of(n);
  if (lane == 0) {
    if (!(t & 1)) { keys[(long long)b * na + (t >> 1)] = 0ull; approx[(long long)b * na + (t >> 1)] = 0u; }
    if (warp < n_strips) strip_count[warp] = 0;
  }
  const T* src = metric + v.batch_offset(b) + (long long)t * v.sn;
  const int row = (t & 1) ? na + (t >> 1) : (t >> 1);
  const long long ro = ((long long)b * n + row) * cm;
  __nv_bfloat16* h_row = hm + ro;
  __nv_bfloat16* m_row = h_row + (long long)bm * n * cm;
  float* f_row = mhat + ro;
  if (cm <= 256) {                 // whole row in registers: one pass over global memory
    float2 x[4];
    double ss = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 2 * lane + 64 * q;
      x[q] = k < cm ? metric_pair(src, heads, stride_h, k) : make_float2(0.f, 0.f);
      ss = fma((double)x[q].x, (double)x[q].x, ss);
      ss = fma((double)x[q].y, (double)x[q].y, ss);
    }
    ss = warp_sum(ss);
    const float norm = (float)sqrt(ss);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 2 * lane + 64 * q;
      if (k < cm) store_split(make_float2(__fdiv_rn(x[q].x, norm), __fdiv_rn(x[q].y, norm)), h_row, m_row, f_row, k);
    }
    return;
  }
  double ss = 0.0;
  for (int k = 2 * lane; k < cm; k += 64) {
    const float2 x = metric_pair(src, heads, stride_h, k);
    ss = fma((double)x.x, (double)x.x, ss);
    ss = fma((double)x.y, (double)x.y, ss);
  }
  ss = warp_sum(ss);
  const float norm = (float)sqrt(ss);
  for (int k = 2 * lane; k < cm; k += 64) {
    const float2 x = metric_pair(src, heads, stride_h, k);
    store_split(make_float2(__fdiv_rn(x.x, norm), __fdiv_rn(x.y, norm)), h_row, m_row, f_row, k);
  }
}


// The same normalisation, eight channels per lane: 16-byte loads, eight lanes per token and four tokens per warp, so a
// warp has 4 x 12 head rows of 128 bytes in flight and a quarter of the load / address instructions per token (the pair
// form above is issue-bound: 10.7 us under ncu at the bench shape, profiles/r01e_hotpath_ncu.txt).  Needs 16-byte aligned
// rows and cm <= 64; everything else takes split_rows_kernel.
template <typename T> struct Chunk8;
template <> struct Chunk8<float> {
  static constexpr int NB = 6;
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load(const float* p) {
    Raw r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[8]) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
  }
  static __device__ __forceinline__ float round_like_input(float x) { return x; }
};
template <> struct Chunk8<__nv_bfloat16> {
  static constexpr int NB = 12;
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void unpack(const Raw& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
  static __device__ __forceinline__ float round_like_input(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

template <typename T>
__global__ void __launch_bounds__(256) split_rows8_kernel(const T* __restrict__ metric, View v, int heads, long long stride_h,
                                                          int bm, int n, int cm, __nv_bfloat16* __restrict__ hm,
                                                          float* __restrict__ mhat, unsigned long long* __restrict__ keys,
                                                          unsigned int* __restrict__ approx, int* __restrict__ strip_count,
                                                          int n_strips) {
  const int lane = threadIdx.x & 31, l8 = lane & 7, k0 = 8 * l8;
  const long long token = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // eight lanes per token
  const bool on = token < (long long)bm * n;
  const int b = on ? (int)(token / n) : 0, t = on ? (int)(token - (long long)b * n) : 0, na = na_of(n);
  if (on && l8 == 0) {
    if (!(t & 1)) { keys[(long long)b * na + (t >> 1)] = 0ull; approx[(long long)b * na + (t >> 1)] = 0u; }
    if (token < n_strips) strip_count[token] = 0;
  }
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = 0.f;
  if (on && k0 < cm) {
    const T* src = metric + v.batch_offset(b) + (long long)t * v.sn + k0;
    if (heads == 1) {
      Chunk8<T>::unpack(Chunk8<T>::load(src), x);
    } else {
      constexpr int NB = Chunk8<T>::NB;
      int h = 0;
      for (; h + NB <= heads; h += NB) {        // every load of a batch in flight before the first add; adds sequential in h
        typename Chunk8<T>::Raw a[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) a[u] = Chunk8<T>::load(src + (long long)(h + u) * stride_h);
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          float f[8];
          Chunk8<T>::unpack(a[u], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] += f[e];
        }
      }
      for (; h < heads; ++h) {
        float f[8];
        Chunk8<T>::unpack(Chunk8<T>::load(src + (long long)h * stride_h), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] += f[e];
      }
      const float inv = 1.0f / (float)heads;    // 1/H multiply like ATen's MeanOps, rounded to the input dtype like k.mean(1)
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = Chunk8<T>::round_like_input(x[e] * inv);
    }
  }
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int e = 0; e < 8; e += 2) { s0 = fma((double)x[e], (double)x[e], s0); s1 = fma((double)x[e + 1], (double)x[e + 1], s1); }
  double ss = s0 + s1;
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  if (!on || k0 >= cm) return;
  const float norm = (float)sqrt(ss);
  float mh[8], hf[8], mf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mh[e] = __fdiv_rn(x[e], norm);
    hf[e] = __bfloat162float(__float2bfloat16_rn(mh[e]));
    mf[e] = mh[e] - hf[e];                      // exact
  }
  uint32_t hw[4], mw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(hf[2 * i], hf[2 * i + 1]), mm = __floats2bfloat162_rn(mf[2 * i], mf[2 * i + 1]);
    hw[i] = *reinterpret_cast<const uint32_t*>(&hh);
    mw[i] = *reinterpret_cast<const uint32_t*>(&mm);
  }
  const int row = (t & 1) ? na + (t >> 1) : (t >> 1);
  const long long ro = ((long long)b * n + row) * cm + k0;
  *reinterpret_cast<uint4*>(hm + ro) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(hm + (long long)bm * n * cm + ro) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  float4* fr = reinterpret_cast<float4*>(mhat + ro);
  fr[0] = make_float4(mh[0], mh[1], mh[2], mh[3]);
  fr[1] = make_float4(mh[4], mh[5], mh[6], mh[7]);
}

#define TC_TRACE(slot) do { if (p.trace) p.trace[(((long long)b * gridDim.y + it) * gridDim.x + jt) * 16 + (slot)] = gtime(); } while (0)

// ---------------------------------------------------------------------------------------------
// 2. TMA -> tcgen05 -> TMEM epilogue
// ---------------------------------------------------------------------------------------------
// Exact canonical scores from the fp32 normalised rows: every product is exact in fp64, the sum is
// fp64 (the order of the fp64 additions is not part of the score definition), one
// rounding to fp32 at the end.
//
// warp_exact32: lane r owns A row `a_blk + r * cm` and wants its score against B row `brows + col * cm`
// (`act`: this lane has work).  One thread per row reading its two 256-byte rows cost 5.9 us per CTA
// (32 scattered sectors per load instruction), and a warp walking the rows one at a time 12.5 us (32
// dependent L2 round trips) -- profiles/r01_match_notes.md.  Here the warp takes EIGHT work items per pass,
// four lanes per item: a lane loads every fourth 16-byte chunk of the row pair (the four lanes cover 64
// contiguous bytes per instruction), all loads of a pass are issued before the first use, and the sum
// needs two shuffles.  Items are the set bits of the ballot, compacted, so the usual 4-5 active rows of
// a warp (the cross-tile filter below removes the rest) take ONE pass.  cm <= 128, cm % 4 == 0.
__device__ __forceinline__ double pair_dot(const float* __restrict__ a, const float* __restrict__ b, int cm, int lane) {
  double acc = 0.0;
  for (int k = 2 * lane; k < cm; k += 64) {
    const float2 x = __ldg(reinterpret_cast<const float2*>(a + k)), y = __ldg(reinterpret_cast<const float2*>(b + k));
    acc = fma((double)x.x, (double)y.x, acc);
    acc = fma((double)x.y, (double)y.y, acc);
  }
  return acc;
}

__device__ __forceinline__ float warp_exact32(const float* __restrict__ a_blk, const float* __restrict__ brows, int cm,
                                              bool act, int col, int lane) {
  const unsigned full = 0xffffffffu;
  const unsigned actmask = __ballot_sync(full, act);
  float mine = 0.f;
  const int items = __popc(actmask);
  if (items == 0) return mine;
  const int sub = lane & 3;                      // which quarter of the 16-byte chunks of a row this lane reads
  const int nchunk = cm >> 2;                    // 16-byte chunks per row
  const int my_item = __popc(actmask & ((1u << lane) - 1u));     // position of this lane's own row in the work list
#pragma unroll 1
  for (int g = 0; 8 * g < items; ++g) {
    const int item = 8 * g + (lane >> 2);        // the work item this lane helps with in this pass
    const bool on = item < items;
    const int r = on ? (int)__fns(actmask, 0, item + 1) : 0;     // its row: the item-th set bit
    const int c = __shfl_sync(full, col, r);
    const float4* ap = reinterpret_cast<const float4*>(a_blk + (long long)r * cm);
    const float4* bp = reinterpret_cast<const float4*>(brows + (long long)c * cm);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll 1
    for (int f0 = 0; f0 < nchunk; f0 += 16) {    // 16 chunks (64 channels) per step: 4 chunks per lane
      float4 x[4], y[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int f = f0 + 4 * t + sub;
        const bool ok = on && f < nchunk;
        x[t] = ok ? __ldg(ap + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        y[t] = ok ? __ldg(bp + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc0 = fma((double)x[t].x, (double)y[t].x, acc0);
        acc1 = fma((double)x[t].y, (double)y[t].y, acc1);
        acc2 = fma((double)x[t].z, (double)y[t].z, acc2);
        acc3 = fma((double)x[t].w, (double)y[t].w, acc3);
      }
    }
    double sum = (acc0 + acc1) + (acc2 + acc3);
    sum += __shfl_xor_sync(full, sum, 1);
    sum += __shfl_xor_sync(full, sum, 2);
    // item 8g + u of the pass was summed by lanes 4u .. 4u+3; its owner fetches it
    const double res = __shfl_sync(full, sum, (my_item & 7) << 2);
    if (act && (my_item >> 3) == g) mine = (float)res;
  }
  return mine;
}

__global__ void __launch_bounds__(TC_THREADS, 3)
match_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int jt = blockIdx.x, it = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) TC_TRACE(0);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t a_bytes = TC_BM * 128u, b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = 2u * (a_bytes + b_bytes);            // A h | A m | B h | B m
  const uint32_t bars = base + (uint32_t)p.stages * stage_bytes;    // full[stages] | empty[stages] | tmem_full | tmem_ptr
  const uint32_t bar_full = bars, bar_empty = bars + 8u * p.stages, bar_tmem = bars + 16u * p.stages;
  const uint32_t tmem_slot = bar_tmem + 8u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
  volatile int* is_last_ptr = reinterpret_cast<volatile int*>(gen_base + (tmem_slot + 4u - base));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
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

  const int row_a = b * p.n + it * TC_BM;               // rows of the split buffer (h plane)
  const int row_b = b * p.n + p.na + jt * p.BN;

  // Producer and MMA warps run converged with one elected lane issuing: under `if (lane == 0)` the compiler cannot prove the
  // operands warp-uniform and wraps every TMA / tcgen05.mma issue in an ELECT / vote loop of ~70 cycles (attn_f32.cu), more
  // than the 56-cycle slot of a 128 x 112 x 16 MMA -- and this kernel is one serial chain of such latencies.
  if (warp == 0) {
    const bool leader = elect_one_sync();
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % p.stages;
      mbar_wait(bar_empty + 8u * s, ((kb / p.stages) & 1) ^ 1);
      if (leader) {
        const uint32_t st = base + (uint32_t)s * stage_bytes, full = bar_full + 8u * s;
        mbar_expect_tx(full, stage_bytes);
        tma_load_2d(st, &map_a, kb * TC_BK, row_a, full);                                   // A h
        tma_load_2d(st + a_bytes, &map_a, kb * TC_BK, row_a + p.rows_total, full);          // A m
        tma_load_2d(st + 2u * a_bytes, &map_b, kb * TC_BK, row_b, full);                    // B h
        tma_load_2d(st + 2u * a_bytes + b_bytes, &map_b, kb * TC_BK, row_b + p.rows_total, full);  // B m
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % p.stages;
      mbar_wait(bar_full + 8u * s, (kb / p.stages) & 1);
      tc_fence_after();
      if (leader) {
        if (kb == 0) TC_TRACE(2);
        if (kb == p.num_kb - 1) TC_TRACE(3);
        const uint32_t st = base + (uint32_t)s * stage_bytes;
        const uint64_t a_h = make_sw128_desc(st), a_m = make_sw128_desc(st + a_bytes);
        const uint64_t b_h = make_sw128_desc(st + 2u * a_bytes), b_m = make_sw128_desc(st + 2u * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UK; ++k) {
          const uint64_t adv = (uint64_t)((k * TC_UK * 2) >> 4);     // +32 bytes inside the swizzle row
          umma_bf16(tb, a_h + adv, b_h + adv, idesc, (kb | k) ? 1u : 0u);
          umma_bf16(tb, a_h + adv, b_m + adv, idesc, 1u);
          umma_bf16(tb, a_m + adv, b_h + adv, idesc, 1u);
        }
        umma_commit(bar_empty + 8u * s);      // frees the stage when these MMAs retire
      }
      __syncwarp();
    }
    if (leader) {
      umma_commit(bar_tmem);                  // accumulator complete
      TC_TRACE(4);
    }
    __syncwarp();
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
    // pass 1: approximate row max over the valid columns.  max.NaN propagates NaN, so a NaN anywhere
    // (zero-norm row) surfaces in m.
    constexpr int MAXCH = 8;                   // BN <= 256
    float m = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch) {
      const int c = ch * 32;
      if (c < ncol) {                          // warp-uniform
        float cm = -INFINITY;
        if (c + 32 <= ncol) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          if (mask0 && c == 0) v[0] = -INFINITY;
          float part[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four independent max chains
#pragma unroll
          for (int e = 0; e < 32; ++e) part[e & 3] = fmax_nan(part[e & 3], v[e]);
          cm = fmax_nan(fmax_nan(part[0], part[1]), fmax_nan(part[2], part[3]));
        } else {
          for (int cc = c; cc < ncol; cc += 16) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)cc, v);
            if (mask0 && cc == 0) v[0] = -INFINITY;
#pragma unroll
            for (int e = 0; e < 16; ++e) cm = fmax_nan(cm, (cc + e < ncol) ? v[e] : -INFINITY);
          }
        }
        m = fmax_nan(m, cm);
      }
    }
    const bool has_nan = (m != m);
    // publish this tile's approximate row max: tiles that later see a clearly better one skip their exact work
    if (p.fused_refine && i < p.na && !has_nan) atomicMax(p.approx + (long long)b * p.na + i, orderable_key(m));
    if (threadIdx.x == 64) TC_TRACE(6);
    // pass 2: every column within the error window of the max is a candidate.  Branch-free: one bit
    // per column (a divergent scan with a dynamically indexed candidate list cost 5.6 us per CTA,
    // see profiles/r01_match_notes.md), candidates are pulled out of the masks afterwards.
    const float thr = m - p.window;
    uint32_t bits[MAXCH];
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch) {
      const int c = ch * 32;
      uint32_t bm_ = 0u;
      // warp-uniform condition only: tcgen05.ld is .sync.aligned, every lane must execute it.  (A per-lane
      // `!has_nan` here hung the kernel as soon as one row of the strip was NaN -- a zero-norm token.)  With a NaN
      // threshold every comparison below is false, so a NaN row gets no candidates and goes the overflow way.
      if (c < ncol) {
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
      bits[ch] = bm_;
    }
    if (threadIdx.x == 64) TC_TRACE(11);
    // cross-tile filter (see below): the read-back is issued here so that its latency hides behind the extraction
    unsigned int peer_key = 0u;
    if (p.fused_refine && i < p.na && !has_nan) peer_key = __ldcg(p.approx + (long long)b * p.na + i);
    int cnt = 0, filled = 0, cand[KCAND];
#pragma unroll
    for (int e = 0; e < KCAND; ++e) cand[e] = 0;
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch) {
      uint32_t w = bits[ch];
      cnt += __popc(w);
      while (w != 0u && filled < KCAND) {        // rare per (lane, chunk): a real branch beats a predicated chain
        const int col = ch * 32 + __ffs(w) - 1;
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
    if (!p.fused_refine) {
      if (i < p.na) {                          // this tile's pruning record for the row (refine_rows_kernel reads it)
        const long long o = ((long long)b * p.na + i) * p.n_ct + jt;
        p.tile_max[o] = m;
        p.tile_cnt[o] = overflow ? CNT_OVERFLOW : cnt;
        *reinterpret_cast<int4*>(p.tile_cand + o * KCAND) = make_int4(j0 + cand[0], j0 + cand[1], j0 + cand[2], j0 + cand[3]);
      }
    } else {
      // Exact scoring of the survivors from the fp32 rows (L2-resident: split_rows just wrote them),
      // warp-cooperatively.  Cross-tile filter first: the other column tiles of this row strip run at the
      // same time and have published their approximate maxima; if one of them beats ours by more than the
      // window, this tile cannot hold the row's max and its exact work is skipped.  The filter is purely
      // opportunistic -- a tile that has not published yet just means some exact work that loses the
      // atomicMax below -- so the result does not depend on timing.
      const bool valid = i < p.na, cls_row = p.cls && i == 0;
      bool skip = false;
      if (valid && !has_nan) skip = key_to_float(peer_key) > m + p.window;
      const float* a_blk = p.mhat + ((long long)b * p.n + it * TC_BM + q * 32) * p.cm;
      const float* brows = p.mhat + ((long long)b * p.n + p.na + j0) * p.cm;
      unsigned long long best = cls_row ? pack_best(-INFINITY, 0) : 0ull;
      const bool need = valid && !cls_row && !overflow && !skip;
      const int rounds = __reduce_max_sync(0xffffffffu, need ? cnt : 0);
      for (int sidx = 0; sidx < rounds; ++sidx) {
        const bool act = need && sidx < cnt;
        const int col = sidx == 0 ? cand[0] : sidx == 1 ? cand[1] : sidx == 2 ? cand[2] : cand[3];
        const float sc = warp_exact32(a_blk, brows, p.cm, act, col, lane);
        if (act) {
          const unsigned long long k = pack_best(sc, j0 + col);
          best = k > best ? k : best;
        }
      }
      unsigned ov = __ballot_sync(0xffffffffu, valid && !cls_row && overflow && !skip);
      while (ov) {                               // too many near-ties (or NaN) in this row: score the whole tile
        const int r = __ffs(ov) - 1;
        ov &= ov - 1u;
        for (int cc = 0; cc < ncol; ++cc) {
          double acc = pair_dot(a_blk + (long long)r * p.cm, brows + (long long)cc * p.cm, p.cm, lane);
          acc = warp_sum(acc);
          if (lane == r) {
            const float sc = (mask0 && cc == 0) ? -INFINITY : (float)acc;
            const unsigned long long k = pack_best(sc, j0 + cc);
            best = k > best ? k : best;
          }
        }
      }
      if (valid && best != 0ull) atomicMax(p.keys + (long long)b * p.na + i, best);
    }
  }
  if (threadIdx.x == 64) TC_TRACE(8);
  tc_fence_before();
  if (p.fused_refine && p.node_max) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(9);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
  if (p.fused_refine && p.node_max) {      // node_max == NULL: the selection kernel consumes the packed keys directly
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
// 3. exact refine as its own pass (wide metrics): canonical fp64 score of the surviving candidates
// ---------------------------------------------------------------------------------------------
// One WARP per A row: lanes stride k, so every candidate costs two coalesced row reads and a warp
// reduction.  (The first version used one thread per row: at Cm = 768 its uncoalesced, serial 768-step
// loops took 400 us.)  fp64 partial sums per lane, then a butterfly: the order of the fp64 additions
// is not part of the score definition (it can move the fp32-rounded result only when the fp64 sum lies
// within ~1e-16 of a rounding boundary).
__device__ __forceinline__ float exact_score_warp(const float* __restrict__ a, const float* __restrict__ b, int cm, int lane) {
  double acc = 0.0;
  for (int k = lane; k < cm; k += 32) acc = fma((double)a[k], (double)b[k], acc);
  acc = warp_sum(acc);
  return (float)acc;
}

__global__ void __launch_bounds__(256) refine_rows_kernel(TcParams p, float* __restrict__ node_max, int* __restrict__ node_idx) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= p.bm * p.na) return;
  const int b = t / p.na, i = t - b * p.na;
  unsigned long long best = 0ull;
  if (p.cls && i == 0) {
    best = pack_best(-INFINITY, 0);
  } else {
    const float* arow = p.mhat + ((long long)b * p.n + i) * p.cm;
    const float* bbase = p.mhat + ((long long)b * p.n + p.na) * p.cm;
    const long long o = (long long)t * p.n_ct;
    float big = -INFINITY;
    for (int c = 0; c < p.n_ct; ++c) big = fmaxf(big, p.tile_max[o + c]);
    const float thr = big - p.window;
    for (int c = 0; c < p.n_ct; ++c) {
      const int cnt = p.tile_cnt[o + c];
      if (cnt == CNT_OVERFLOW) {            // too many near-ties (or NaN): score the whole column tile exactly
        const int je = min(p.nb, (c + 1) * p.BN);
        for (int j = c * p.BN; j < je; ++j) {
          const float s = (p.distill && j == 0) ? -INFINITY : exact_score_warp(arow, bbase + (long long)j * p.cm, p.cm, lane);
          const unsigned long long k = pack_best(s, j);
          best = k > best ? k : best;
        }
      } else if (p.tile_max[o + c] >= thr) {
        for (int s = 0; s < cnt; ++s) {
          const int j = p.tile_cand[(o + c) * KCAND + s];
          const float sc = exact_score_warp(arow, bbase + (long long)j * p.cm, p.cm, lane);
          const unsigned long long k = pack_best(sc, j);
          best = k > best ? k : best;
        }
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
  const int n_rt = (p.na + TC_BM - 1) / TC_BM;
  // column tiles: 128 wide by default (60-64 KB of operands: three CTAs per SM); narrower, down to 64,
  // while the whole grid would otherwise leave SM slots empty
  int n_ct = (p.nb + 127) / 128;
  const int want = (2 * 148) / (bm * n_rt > 0 ? bm * n_rt : 1);
  const int n_ct_max = (p.nb + 63) / 64;
  if (want > n_ct) n_ct = want < n_ct_max ? want : n_ct_max;
  const int forced = env_int("TOME_TC_NCT", 0);          // tuning knob (bench sweeps, tests)
  if (forced > 0) n_ct = forced > (p.nb + 15) / 16 ? (p.nb + 15) / 16 : forced;
  if (n_ct < (p.nb + 255) / 256) n_ct = (p.nb + 255) / 256;
  if (n_ct < 1) n_ct = 1;
  int bn = (p.nb + n_ct - 1) / n_ct;
  bn = (bn + 15) & ~15;
  p.BN = bn < 16 ? 16 : bn;
  p.n_ct = (p.nb + p.BN - 1) / p.BN;
  const int stage_bytes = 2 * (TC_BM + p.BN) * 128;
  int st = (200 * 1024) / stage_bytes;
  st = st > 4 ? 4 : st;
  p.stages = st > p.num_kb ? p.num_kb : st;
  p.fused_refine = (cm <= 128) && !env_int("TOME_TC_NO_FUSED_REFINE", 0);
  p.tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  // error bound of h.h + h.m + m.h on unit vectors: x = h + m + l with |l| <= 2^-16 |x|, so the
  // three dropped products (m.m, l.b, a.l) are <= 3 * 2^-16 = 4.6e-5 by Cauchy-Schwarz; the bf16 products
  // are exact in fp32 and the (3 cm / 16) fp32 accumulations add <= 2^-22 each; margin on top
  const float eps = 5e-5f + 2e-7f * (float)cm;
  p.window = 2.0f * eps;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TcLayout { size_t hm, mhat, tile_max, tile_cnt, tile_cand, keys, approx, strips, total; };

static TcLayout tc_layout(int bm, int n, const TcParams& p) {
  TcLayout l;
  const size_t elems = (size_t)bm * n * p.cm, rows = (size_t)bm * p.na * p.n_ct;
  const size_t strips = (size_t)bm * ((p.na + TC_BM - 1) / TC_BM);
  size_t o = 0;
  l.hm = o;        o += align256(2 * elems * sizeof(__nv_bfloat16));
  l.mhat = o;      o += align256(elems * sizeof(float));
  l.tile_max = o;  o += align256(rows * 4);
  l.tile_cnt = o;  o += align256(rows * 4);
  l.tile_cand = o; o += align256(rows * 4 * KCAND);
  l.keys = o;      o += align256((size_t)bm * p.na * 8);
  l.approx = o;    o += align256((size_t)bm * p.na * 4);
  l.strips = o;    o += align256((size_t)bm * 8 + strips * 4);     // [bm x u64 select flags][strip arrival counters]
  l.total = o;
  return l;
}

size_t match_tc_workspace(int bm, int n, int cm) {
  TcParams p;
  tc_geometry(bm, n, cm, p);
  return tc_layout(bm, n, p).total;
}

// Test/diagnostic hook: {n_ct, BN, byte offset of tile_max, byte offset of tile_cnt, fused_refine} so that
// tests can read the tensor-core pass's pruning record without mirroring the geometry.
void match_tc_describe(int bm, int n, int cm, long long out[5]) {
  TcParams p;
  tc_geometry(bm, n, cm, p);
  const TcLayout l = tc_layout(bm, n, p);
  out[0] = p.n_ct; out[1] = p.BN; out[2] = (long long)l.tile_max; out[3] = (long long)l.tile_cnt; out[4] = p.fused_refine;
}

bool match_tc_fused_refine(int bm, int n, int cm) {
  TcParams p;
  tc_geometry(bm, n, cm, p);
  return p.fused_refine != 0;
}

bool match_tc_supported(int dtype, int bm, int n, int cm, const View& v, const void* metric) {
  if (dtype != TOME_F32 && dtype != TOME_BF16) return false;
  if (cm % 8 != 0 || cm < 8 || cm > 4096) return false;          // TMA: 16-byte row pitch of the bf16 planes
  if (n < 2 || (long long)bm * n * 2 > 0x7fffffffLL) return false;
  // split_rows reads channel pairs: every row must start on a pair boundary
  const uintptr_t pair = dtype == TOME_F32 ? 8 : 4;
  if (((uintptr_t)metric % pair) || (v.sbo & 1) || (v.sbi & 1) || (v.sn & 1)) return false;
  return true;
}

int launch_match_tc(const void* metric, int dtype, int bm, int n, int cm, const View& v, int cls, int distill,
                    float* node_max, int* node_idx, void* ws, size_t ws_bytes, cudaStream_t st, int heads,
                    long long stride_h, unsigned long long** packed_out, unsigned long long** flags_out) {
  if (ws_bytes < match_tc_workspace(bm, n, cm))
    return set_error(TOME_ERR_WORKSPACE, "tome_match: workspace %zu < %zu bytes", ws_bytes, match_tc_workspace(bm, n, cm));
  if (((uintptr_t)ws & 255) != 0) return set_error(TOME_ERR_ALIGN, "tome_match: workspace must be 256-byte aligned");
  if (stride_h & 1) return set_error(TOME_ERR_ALIGN, "tome_match_heads: head stride must be even");
  TcParams p;
  tc_geometry(bm, n, cm, p);
  const TcLayout l = tc_layout(bm, n, p);
  p.cls = cls; p.distill = distill; p.rows_total = bm * n;
  p.node_max = node_max; p.node_idx = node_idx;
  p.trace = getenv("TOME_TC_TRACE") ? (long long*)strtoull(getenv("TOME_TC_TRACE"), nullptr, 0) : nullptr;
  const int n_rt = (p.na + TC_BM - 1) / TC_BM;
  char* w = (char*)ws;
  __nv_bfloat16* hm = (__nv_bfloat16*)(w + l.hm);
  float* mhat = (float*)(w + l.mhat);
  p.mhat = mhat;
  p.tile_max = (float*)(w + l.tile_max);
  p.tile_cnt = (int*)(w + l.tile_cnt);
  p.tile_cand = (int*)(w + l.tile_cand);
  p.keys = (unsigned long long*)(w + l.keys);
  p.approx = (unsigned int*)(w + l.approx);
  int* zero_base = (int*)(w + l.strips);                  // flags + counters, zeroed by the normalisation kernel
  p.strip_count = zero_base + 2 * bm;
  const int zero_ints = 2 * bm + bm * n_rt;
  if (packed_out) *packed_out = p.keys;
  if (flags_out) *flags_out = (unsigned long long*)zero_base;

  const long long vec = dtype == TOME_F32 ? 4 : 8;            // 16-byte loads of eight channels
  const bool wide = cm <= 64 && !((uintptr_t)metric & 15) && v.sbo % vec == 0 && v.sbi % vec == 0 && v.sn % vec == 0 &&
                    (heads == 1 || stride_h % vec == 0) && !env_int("TOME_SPLIT_PAIRS", 0);
  if (wide) {
    const int blocks8 = (int)(((long long)bm * n * 8 + 255) / 256);
    if (dtype == TOME_F32)
      split_rows8_kernel<float><<<blocks8, 256, 0, st>>>((const float*)metric, v, heads, stride_h, bm, n, cm, hm, mhat, p.keys, p.approx, zero_base, zero_ints);
    else
      split_rows8_kernel<__nv_bfloat16><<<blocks8, 256, 0, st>>>((const __nv_bfloat16*)metric, v, heads, stride_h, bm, n, cm, hm, mhat, p.keys, p.approx, zero_base, zero_ints);
    TOME_LAUNCH_CHECK("split_rows8_kernel");
  } else {
    const int blocks = (int)(((long long)bm * n * 32 + 255) / 256);
    if (dtype == TOME_F32)
      split_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)metric, v, heads, stride_h, bm, n, cm, hm, mhat, p.keys, p.approx, zero_base, zero_ints);
    else
      split_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)metric, v, heads, stride_h, bm, n, cm, hm, mhat, p.keys, p.approx, zero_base, zero_ints);
    TOME_LAUNCH_CHECK("split_rows_kernel");
  }

  alignas(64) CUtensorMap map_a, map_b;
  int rc = make_bf16_map(&map_a, hm, 2LL * bm * n, cm, cm, TC_BM, "tome_match");
  if (rc) return rc;
  rc = make_bf16_map(&map_b, hm, 2LL * bm * n, cm, cm, p.BN, "tome_match");
  if (rc) return rc;
  const size_t smem = (size_t)p.stages * 2 * (TC_BM + p.BN) * 128 + 16 * p.stages + 16 + 1024;
  static PerDeviceOnce smem_set;
  if (smem_set.first_time())
    TOME_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  dim3 grid(p.n_ct, n_rt, bm);
  match_tc_kernel<<<grid, TC_THREADS, smem, st>>>(map_a, map_b, p);
  TOME_LAUNCH_CHECK("match_tc_kernel");
  if (!p.fused_refine) {
    const long long total = (long long)bm * p.na * 32;
    refine_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p, node_max, node_idx);
    TOME_LAUNCH_CHECK("refine_rows_kernel");
  }
  return TOME_OK;
}

}  // namespace tome
