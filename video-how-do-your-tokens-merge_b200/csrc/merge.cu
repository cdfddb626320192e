// Kernel 3: merge.  Replaces the reference's merge()/drop()/unmerge() closures and
// merge_wavg / merge_source (tome/merge.py:75-100, 260-269, 316-334, 355-384).
//
// The reference does, per call: two strided slices, two gathers with stride-0 int64 index
// tensors, an out-of-place scatter_reduce (clone + atomics) and a cat -- and merge_wavg
// does all of that twice plus two elementwise passes.  Here every OUTPUT token row is
// produced by one warp in one pass: it gathers its own row and (for a B token) the rows
// of the A tokens merged into it through the dst-grouped CSR built by select.cu, reduces
// in fp32 in the reference CPU order (self, then ascending k; products and sums rounded
// separately, no FMA contraction), and writes the merged row, its new size and log(size)
// once.  HBM traffic = read x once + write x' once (kernel 3).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace tome {

struct MergeArgs {
  int bm, n, r, distill, c, mode, hybrid;
  float thr;
  const float* node_max;
  const int *unm_idx, *b_off, *b_src, *b_head;
  const void* x;
  const void* res;       // optional residual: every input row is round_T(x + res)
  View xv;
  View rv;               // addressing of the residual (== xv unless the caller gave it its own view)
  const float* size_in;
  void* out;
  View ov;
  float* size_out;
  float* logsize_out;
  // optional fused LayerNorm of the merged rows (the block's norm2, tome/patch/videomae.py:21-22):
  const void* ln_w;      // gamma (c) in x's dtype, or NULL
  const void* ln_b;      // beta  (c) or NULL
  float ln_eps;
  void* normed;          // (bm, n - r, c) LayerNorm(out), addressed like `out` through nv
  View nv;
};

// output slot -> (kind, index): merge.py:82-85 ordering
__device__ __forceinline__ void slot_to_token(int o, int nu, int distill, bool& is_unm, int& idx) {
  if (!distill) { is_unm = o < nu; idx = is_unm ? o : o - nu; return; }
  if (o == 0) { is_unm = true; idx = 0; }
  else if (o == 1) { is_unm = false; idx = 0; }
  else if (o <= nu) { is_unm = true; idx = o - 1; }
  else { is_unm = false; idx = o - nu; }
}
__device__ __forceinline__ int token_to_slot(bool is_unm, int idx, int nu, int distill) {
  if (!distill) return is_unm ? idx : nu + idx;
  if (is_unm) return idx == 0 ? 0 : idx + 1;
  return idx == 0 ? 1 : nu + idx;
}

template <typename T, int E> struct Vec;
template <> struct Vec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[1]) { f[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&f)[1]) { *p = f[0]; }
};
template <> struct Vec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[1]) { f[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[1]) { *p = __float2bfloat16_rn(f[0]); }
};

__device__ __forceinline__ float nanmax(float a, float b) { return (a != a || a > b) ? a : ((b != b) ? b : (a > b ? a : b)); }

// (packed fp32 arithmetic -- add2 / sub2 / mul2 / fma2 -- lives in common.cuh)

// One warp per output row; NV vectors of E elements per lane per chunk.
template <typename T, int E, int NV>
__global__ void __launch_bounds__(256) merge_rows_kernel(MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const int n = a.n, na = na_of(n), nb = nb_of(n), r = a.r, nu = na - r, nout = n - r;
  if (o >= nout) return;
  bool is_unm; int idx;
  slot_to_token(o, nu, a.distill, is_unm, idx);

  const T* xb = reinterpret_cast<const T*>(a.x) + a.xv.batch_offset(b);
  T* ob = reinterpret_cast<T*>(a.out) + a.ov.batch_offset(b) + (long long)o * a.ov.sn;
  const float* szb = a.size_in ? a.size_in + (long long)b * n : nullptr;

  int self_tok, beg = 0, end = 0;
  if (is_unm) {
    self_tok = 2 * __ldg(a.unm_idx + (long long)b * nu + idx);
  } else {
    self_tok = 2 * idx + 1;
    if (a.mode != TOME_MODE_DROP) {
      const int* off = a.b_off + (long long)b * (nb + 1) + idx;
      beg = __ldg(off); end = __ldg(off + 1);
    }
  }
  const int* bsrc = a.b_src + (long long)b * r;
  bool keep_self = true;
  if (a.hybrid) {   // merge.py:326: 'prod' with (node_max_sorted[:r] >= thr) over edges hitting this B token
    for (int q = beg; q < end; ++q)
      keep_self &= (__ldg(a.node_max + (long long)b * na + __ldg(bsrc + q)) >= a.thr);
  }
  const float s_self = szb ? __ldg(szb + self_tok) : 1.0f;
  const bool wavg = a.mode == TOME_MODE_WAVG;

  // new size (exact: sizes are small integers in fp32)
  float S = s_self;
  if (!is_unm) {
    if (!keep_self) S = __fmul_rn(S, 0.0f);
    if (wavg) for (int q = beg; q < end; ++q) S = __fadd_rn(S, szb ? __ldg(szb + 2 * __ldg(bsrc + q)) : 1.0f);
  }
  const float cnt = (float)(1 + end - beg);

  const int chunk = 32 * E * NV;
  for (int cb = 0; cb < a.c; cb += chunk) {
    float acc[NV][E];
    const T* row = xb + (long long)self_tok * a.xv.sn + cb;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c0 = (v * 32 + lane) * E;
      if (cb + c0 < a.c) Vec<T, E>::load(row + c0, acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float t = acc[v][e];
        if (wavg) t = __fmul_rn(t, s_self);
        if (!keep_self) t = __fmul_rn(t, 0.0f);
        acc[v][e] = t;
      }
    for (int q = beg; q < end; ++q) {
      const int tok = 2 * __ldg(bsrc + q);
      const float s = (wavg && szb) ? __ldg(szb + tok) : 1.0f;
      const T* srow = xb + (long long)tok * a.xv.sn + cb;
      float tmp[NV][E];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c0 = (v * 32 + lane) * E;
        if (cb + c0 < a.c) Vec<T, E>::load(srow + c0, tmp[v]);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (a.mode == TOME_MODE_AMAX) acc[v][e] = nanmax(acc[v][e], tmp[v][e]);
          else if (wavg) acc[v][e] = __fadd_rn(acc[v][e], __fmul_rn(tmp[v][e], s));
          else acc[v][e] = __fadd_rn(acc[v][e], tmp[v][e]);
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c0 = (v * 32 + lane) * E;
      if (cb + c0 < a.c) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (wavg) acc[v][e] = __fdiv_rn(acc[v][e], S);
          else if (a.mode == TOME_MODE_MEAN && !is_unm) acc[v][e] = __fdiv_rn(acc[v][e], cnt);
        }
        Vec<T, E>::store(ob + cb + c0, acc[v]);
      }
    }
  }
  if (lane == 0) {
    const long long so = (long long)b * nout + o;
    const float So = (a.mode == TOME_MODE_DROP) ? 1.0f : S;
    if (a.size_out) a.size_out[so] = So;
    if (a.logsize_out) a.logsize_out[so] = logf(So);
  }
}

// ---- fast path: 16-byte vectors -------------------------------------------------------------
// Most output rows are plain copies (kept tokens of size 1, B tokens nobody merged into):
// they stay packed in registers and go straight back out.  Only rows that need arithmetic
// ((x*s)/s with s != 1, or a reduction over merged sources) are unpacked to fp32.
template <typename T> struct Pack;
template <> struct Pack<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
  static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// One-shot gather, one output row per warp.
//
// What this kernel is bound by is not bandwidth but the length of its dependent chains: a
// plain device memcpy of the same bytes takes 6.1 us, and so does this kernel when no row
// has merged sources -- but the first versions took 11-13 us because the ~7% of warps that
// reduce merged sources chased plan lookup -> row -> lookup -> row chains (up to 7 dependent
// global round trips) and every other CTA slot waited behind them
// (profiles/r01_merge_notes.md).  Hence:
//   * hop 1 is ONE 16-byte lookup (plan.b_head = {count, source 0, source 1, CSR begin});
//   * hop 2 requests the token's own row (registers) and up to two source rows (cp.async into
//     a per-warp shared-memory slot pair, so they cost no registers) all at once;
//   * further sources (rare) go through the same slot pair two at a time: one hop per pair.
// Rows nobody merged into -- the vast majority -- are copies out of packed registers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// RES: the rows are the sum of two tensors (the block's residual add, tome/patch/videomae.py:19-20:
// x = x + attn(...) right before the reduction) -- added here instead of in a separate pass; each element
// is rounded to T first, so the result is bit-identical to merging the materialised sum.
template <typename T, int NV, int WARPS, int MINB, bool LN, bool RES>
__global__ void __launch_bounds__(WARPS * 32, MINB) merge_gather_kernel(MergeArgs a) {
  constexpr int E = Pack<T>::E;
  constexpr bool kCopyIfNoSrc = sizeof(T) == 2;   // bf16: round(fp32(x*s)/s) == x -> kept tokens are plain copies
  constexpr int SLOT = NV * 32;                   // 16-byte groups per staged row
  constexpr int NSLOT = RES ? 4 : 2;              // source pair of x (+ source pair of the residual)
  extern __shared__ uint4 stage[];                // WARPS x NSLOT x SLOT
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // grid = (batch, row blocks), row blocks walked from the END of the output: CTAs are dispatched in
  // blockIdx order, so the B-token rows (the only ones that may carry a reduction) start first and
  // the tail of the kernel is made of plain copies
  const int o = ((int)gridDim.y - 1 - (int)blockIdx.y) * WARPS + warp;
  const int b = blockIdx.x;
  const int n = a.n, na = na_of(n), nb = nb_of(n), r = a.r, nu = na - r, nout = n - r;
  if (o >= nout) return;
  const T* xb = reinterpret_cast<const T*>(a.x) + a.xv.batch_offset(b);
  const T* rb = RES ? reinterpret_cast<const T*>(a.res) + a.rv.batch_offset(b) : nullptr;
  const int nvec = a.c / E;
  const bool wavg = a.mode == TOME_MODE_WAVG;

  // hop 1: which input row is this, and what merges into it
  bool is_unm; int idx;
  slot_to_token(o, nu, a.distill, is_unm, idx);
  int tok;
  int4 head = make_int4(0, 0, 0, 0);
  if (is_unm) tok = 2 * __ldg(a.unm_idx + (long long)b * nu + idx);
  else {
    tok = 2 * idx + 1;
    if (a.mode != TOME_MODE_DROP) head = __ldg(reinterpret_cast<const int4*>(a.b_head) + (long long)b * nb + idx);
  }
  const int nsrc = head.x;
  uint4* mine = stage + (size_t)warp * NSLOT * SLOT;
  // hop 2: own row into registers, first two source rows into shared memory
  uint4 raw[NV];
  {
    const uint4* row = reinterpret_cast<const uint4*>(xb + (long long)tok * a.xv.sn);
    uint4 raw2[RES ? NV : 1];
#pragma unroll
    for (int v = 0; v < NV; ++v) { const int i = v * 32 + lane; if (i < nvec) raw[v] = ld_stream_u4(row + i); }
    if (RES) {
      const uint4* row2 = reinterpret_cast<const uint4*>(rb + (long long)tok * a.rv.sn);
#pragma unroll
      for (int v = 0; v < NV; ++v) { const int i = v * 32 + lane; if (i < nvec) raw2[RES ? v : 0] = ld_stream_u4(row2 + i); }
    }
    if (nsrc > 0) {
      const long long o0 = (long long)(2 * head.y) * a.xv.sn, o1 = (long long)(2 * head.z) * a.xv.sn;
      const long long q0 = (long long)(2 * head.y) * a.rv.sn, q1 = (long long)(2 * head.z) * a.rv.sn;
      const uint4* s0 = reinterpret_cast<const uint4*>(xb + o0);
      const uint4* s1 = reinterpret_cast<const uint4*>(xb + o1);
      for (int i = lane; i < nvec; i += 32) {
        cp_async16(mine + i, s0 + i);
        if (nsrc > 1) cp_async16(mine + SLOT + i, s1 + i);
        if (RES) {
          cp_async16(mine + 2 * SLOT + i, reinterpret_cast<const uint4*>(rb + q0) + i);
          if (nsrc > 1) cp_async16(mine + 3 * SLOT + i, reinterpret_cast<const uint4*>(rb + q1) + i);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (RES) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float f[E], g[E];
        Pack<T>::unpack(raw[v], f);
        Pack<T>::unpack(raw2[RES ? v : 0], g);
#pragma unroll
        for (int e = 0; e < E; e += 2) {
          const float2 t = add2(make_float2(f[e], f[e + 1]), make_float2(g[e], g[e + 1]));
          f[e] = t.x; f[e + 1] = t.y;
        }
        raw[v] = Pack<T>::pack(f);               // rounded to T: what x + attn holds in the reference
      }
    }
  }
  const float* szb = a.size_in ? a.size_in + (long long)b * n : nullptr;
  float S = (wavg && szb) ? __ldg(szb + tok) : 1.0f;
  uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.out) + a.ov.batch_offset(b) + (long long)o * a.ov.sn);

  if (nsrc == 0) {
    if (!(kCopyIfNoSrc || S == 1.0f)) {
      // kept token of size s != 1: the reference really computes (x*s)/s  (merge.py:365-368)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float f[E];
        Pack<T>::unpack(raw[v], f);
#pragma unroll
        for (int e = 0; e < E; ++e) f[e] = __fdiv_rn(__fmul_rn(f[e], S), S);
        raw[v] = Pack<T>::pack(f);
      }
    }
  } else {
    const int* bsrc = a.b_src + (long long)b * r + head.w;
    // lane k looks up source k (token, size, hybrid keep flag); k = 0, 1 come from the head word
    int k_ai = lane == 0 ? head.y : head.z;
    if (lane >= 2 && lane < nsrc) k_ai = __ldg(bsrc + lane);
    float k_s = 1.0f;
    bool k_keep = true;
    if (lane < nsrc) {
      if (wavg && szb) k_s = __ldg(szb + 2 * k_ai);
      if (a.hybrid) k_keep = (__ldg(a.node_max + (long long)b * na + k_ai) >= a.thr);     // merge.py:326
    }
    bool keep_self = __all_sync(0xffffffffu, k_keep);
    if (a.hybrid)
      for (int k = 32; k < nsrc; ++k) keep_self &= (__ldg(a.node_max + (long long)b * na + __ldg(bsrc + k)) >= a.thr);
    const float s_self = keep_self ? S : __fmul_rn(S, 0.0f);
    float Ssum = s_self;
    for (int k = 0; k < nsrc; ++k) {
      const float s = k < 32 ? __shfl_sync(0xffffffffu, k_s, k & 31) : ((wavg && szb) ? __ldg(szb + 2 * __ldg(bsrc + k)) : 1.0f);
      if (wavg) Ssum = __fadd_rn(Ssum, s);
    }
    const float cnt = (float)(1 + nsrc);
    // own row -> fp32 accumulators, kept packed-width: NV x E floats would be 24+ registers, so the
    // accumulation runs per 16-byte group and parks partial sums back in `raw` (as fp32 bit patterns
    // for fp32 rows; bf16 rows re-round only at the very end, see below)
    // pass over source pairs: slots hold sources [k0, k0 + 2)
    float accv[NV][E];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      Pack<T>::unpack(raw[v], accv[v]);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (wavg) accv[v][e] = __fmul_rn(accv[v][e], S);
        if (!keep_self) accv[v][e] = __fmul_rn(accv[v][e], 0.0f);
      }
    }
    for (int k0 = 0; k0 < nsrc; k0 += 2) {
      if (k0 > 0) {                              // rare: refill the slot pair with the next two sources
        __syncwarp();
        const int a0 = k0 < 32 ? __shfl_sync(0xffffffffu, k_ai, k0 & 31) : __ldg(bsrc + k0);
        const int a1 = (k0 + 1 < nsrc) ? ((k0 + 1) < 32 ? __shfl_sync(0xffffffffu, k_ai, (k0 + 1) & 31) : __ldg(bsrc + k0 + 1)) : a0;
        const long long o0 = (long long)(2 * a0) * a.xv.sn, o1 = (long long)(2 * a1) * a.xv.sn;
        const long long q0 = (long long)(2 * a0) * a.rv.sn, q1 = (long long)(2 * a1) * a.rv.sn;
        const uint4* s0 = reinterpret_cast<const uint4*>(xb + o0);
        const uint4* s1 = reinterpret_cast<const uint4*>(xb + o1);
        for (int i = lane; i < nvec; i += 32) {
          cp_async16(mine + i, s0 + i);
          if (k0 + 1 < nsrc) cp_async16(mine + SLOT + i, s1 + i);
          if (RES) {
            cp_async16(mine + 2 * SLOT + i, reinterpret_cast<const uint4*>(rb + q0) + i);
            if (k0 + 1 < nsrc) cp_async16(mine + 3 * SLOT + i, reinterpret_cast<const uint4*>(rb + q1) + i);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = k0 + kk;
        if (k < nsrc) {
          const float s = k < 32 ? __shfl_sync(0xffffffffu, k_s, k & 31) : ((wavg && szb) ? __ldg(szb + 2 * __ldg(bsrc + k)) : 1.0f);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int i = v * 32 + lane;
            if (i < nvec) {
              float f[E];
              Pack<T>::unpack(mine[kk * SLOT + i], f);
              if (RES) {
                float g[E];
                Pack<T>::unpack(mine[(2 + kk) * SLOT + i], g);
#pragma unroll
                for (int e = 0; e < E; ++e) f[e] = __fadd_rn(f[e], g[e]);
                Pack<T>::unpack(Pack<T>::pack(f), f);        // round the sum to T first
              }
#pragma unroll
              for (int e = 0; e < E; ++e) {
                if (a.mode == TOME_MODE_AMAX) accv[v][e] = nanmax(accv[v][e], f[e]);
                else if (wavg) accv[v][e] = __fadd_rn(accv[v][e], __fmul_rn(f[e], s));
                else accv[v][e] = __fadd_rn(accv[v][e], f[e]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (wavg) accv[v][e] = __fdiv_rn(accv[v][e], Ssum);
        else if (a.mode == TOME_MODE_MEAN) accv[v][e] = __fdiv_rn(accv[v][e], cnt);
      }
      raw[v] = Pack<T>::pack(accv[v]);
    }
    S = Ssum;
  }
  // the finished row sits in `raw` (packed, already rounded to the output dtype)
#pragma unroll
  for (int v = 0; v < NV; ++v) { const int i = v * 32 + lane; if (i < nvec) orow[i] = raw[v]; }
  if (LN) {
    // fused LayerNorm over the row this warp holds: fp32 statistics of the ROUNDED row (what a separate
    // LayerNorm kernel would read back), two-pass variance, one more row written instead of a whole
    // read-modify-write pass over x'.  The row is unpacked ONCE and stays in registers as centred values;
    // sums, centring and the affine map are packed f32x2 operations.
    float2 d[NV][E / 2];
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float f[E];
      Pack<T>::unpack(raw[v], f);
      const bool on = v * 32 + lane < nvec;
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        d[v][e / 2] = on ? make_float2(f[e], f[e + 1]) : make_float2(0.f, 0.f);
        s2 = add2(s2, d[v][e / 2]);
      }
    }
    float sum = s2.x + s2.y;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
    const float mean = sum / (float)a.c;
    const float2 mean2 = make_float2(mean, mean);
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const bool on = v * 32 + lane < nvec;
#pragma unroll
      for (int e = 0; e < E / 2; ++e) {
        d[v][e] = sub2(d[v][e], mean2);
        if (on) q2 = fma2(d[v][e], d[v][e], q2);
      }
    }
    float sq = q2.x + q2.y;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, of);
    const float rstd = rsqrtf(sq / (float)a.c + a.ln_eps);
    const float2 rstd2 = make_float2(rstd, rstd);
    uint4* nrow = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.normed) + a.nv.batch_offset(b) + (long long)o * a.nv.sn);
    const uint4* gw = reinterpret_cast<const uint4*>(a.ln_w);
    const uint4* gb = reinterpret_cast<const uint4*>(a.ln_b);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = v * 32 + lane;
      if (i < nvec) {
        float f[E], w[E], bb[E];
        Pack<T>::unpack(__ldg(gw + i), w);
        if (gb) Pack<T>::unpack(__ldg(gb + i), bb);
#pragma unroll
        for (int e = 0; e < E; e += 2) {
          const float2 g2 = mul2(make_float2(w[e], w[e + 1]), rstd2);
          const float2 y = fma2(d[v][e / 2], g2, gb ? make_float2(bb[e], bb[e + 1]) : make_float2(0.f, 0.f));
          f[e] = y.x; f[e + 1] = y.y;
        }
        nrow[i] = Pack<T>::pack(f);
      }
    }
  }
  if (lane == 0) {
    const long long so = (long long)b * nout + o;
    const float So = wavg ? S : 1.0f;
    if (a.size_out) a.size_out[so] = So;
    if (a.logsize_out) a.logsize_out[so] = logf(So);
  }
}

// merge_source with the implicit identity (merge.py:379-381): row of slot o is the OR of
// one-hot rows, generated without reading anything but the plan.
__global__ void __launch_bounds__(256) source_identity_kernel(MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const int n = a.n, na = na_of(n), nb = nb_of(n), r = a.r, nu = na - r, nout = n - r;
  if (o >= nout) return;
  bool is_unm; int idx;
  slot_to_token(o, nu, a.distill, is_unm, idx);
  float* orow = reinterpret_cast<float*>(a.out) + ((long long)b * nout + o) * n;
  int self_tok, beg = 0, end = 0;
  if (is_unm) self_tok = 2 * __ldg(a.unm_idx + (long long)b * nu + idx);
  else {
    self_tok = 2 * idx + 1;
    const int* off = a.b_off + (long long)b * (nb + 1) + idx;
    beg = __ldg(off); end = __ldg(off + 1);
  }
  const int* bsrc = a.b_src + (long long)b * r;
  bool keep_self = true;
  if (a.hybrid)
    for (int q = beg; q < end; ++q)
      keep_self &= (__ldg(a.node_max + (long long)b * na + __ldg(bsrc + q)) >= a.thr);
  for (int c = lane; c < n; c += 32) {
    float v = (c == self_tok && keep_self) ? 1.0f : 0.0f;
    for (int q = beg; q < end; ++q) v = (c == 2 * __ldg(bsrc + q)) ? 1.0f : v;
    orow[c] = v;
  }
}

template <typename T, int E>
__global__ void __launch_bounds__(256) unmerge_rows_kernel(const int* __restrict__ a_map, int n, int r, int distill,
                                                           int c, const T* __restrict__ x, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (t >= n) return;
  const int na = na_of(n), nu = na - r, nout = n - r;
  int slot;
  if (t & 1) slot = token_to_slot(false, t >> 1, nu, distill);
  else {
    const int m = __ldg(a_map + (long long)b * na + (t >> 1));
    slot = m >= 0 ? token_to_slot(true, m, nu, distill) : token_to_slot(false, -m - 1, nu, distill);
  }
  const T* src = x + ((long long)b * nout + slot) * c;
  T* dst = out + ((long long)b * n + t) * c;
  for (int c0 = lane * E; c0 < c; c0 += 32 * E) {
    float f[E];
    Vec<T, E>::load(src + c0, f);
    Vec<T, E>::store(dst + c0, f);
  }
}

// ---- host -----------------------------------------------------------------------------------
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static bool view_vec_ok(const View& v, int e) { return v.sbo % e == 0 && v.sbi % e == 0 && v.sn % e == 0; }

template <typename T>
static int launch_merge_scalar(const MergeArgs& a, cudaStream_t st) {
  const int nout = a.n - a.r;
  dim3 grid((nout + 7) / 8, a.bm);
  merge_rows_kernel<T, 1, 8><<<grid, 256, 0, st>>>(a);
  TOME_LAUNCH_CHECK("merge_rows_kernel");
  return TOME_OK;
}

template <typename T, int NV, int WARPS, int MINB>
static int launch_gather_inst(const MergeArgs& a, cudaStream_t st) {
  const int nout = a.n - a.r;
  dim3 grid(a.bm, (nout + WARPS - 1) / WARPS);
  if (grid.y > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_merge: too many output rows per batch element (%d)", nout);
  const size_t smem = (size_t)WARPS * (a.res ? 4 : 2) * NV * 32 * sizeof(uint4);
  static PerDeviceOnce attr;
  if (attr.first_time()) {
    const int big = (int)((size_t)WARPS * 4 * NV * 32 * sizeof(uint4));
    if (big > 48 * 1024) {
      TOME_CUDA(cudaFuncSetAttribute(merge_gather_kernel<T, NV, WARPS, MINB, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      TOME_CUDA(cudaFuncSetAttribute(merge_gather_kernel<T, NV, WARPS, MINB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      TOME_CUDA(cudaFuncSetAttribute(merge_gather_kernel<T, NV, WARPS, MINB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      TOME_CUDA(cudaFuncSetAttribute(merge_gather_kernel<T, NV, WARPS, MINB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    }
  }
  if (a.res) {
    if (a.normed) merge_gather_kernel<T, NV, WARPS, MINB, true, true><<<grid, WARPS * 32, smem, st>>>(a);
    else merge_gather_kernel<T, NV, WARPS, MINB, false, true><<<grid, WARPS * 32, smem, st>>>(a);
  } else {
    if (a.normed) merge_gather_kernel<T, NV, WARPS, MINB, true, false><<<grid, WARPS * 32, smem, st>>>(a);
    else merge_gather_kernel<T, NV, WARPS, MINB, false, false><<<grid, WARPS * 32, smem, st>>>(a);
  }
  TOME_LAUNCH_CHECK("merge_gather_kernel");
  return TOME_OK;
}

template <typename T, int WARPS, int MINB>
static int launch_merge_gather(const MergeArgs& a, cudaStream_t st) {
  constexpr int E = Pack<T>::E;
  const int nv = (a.c / E + 31) / 32;
  if (nv <= 1) return launch_gather_inst<T, 1, WARPS, MINB>(a, st);
  if (nv <= 2) return launch_gather_inst<T, 2, WARPS, MINB>(a, st);
  if (nv <= 3) return launch_gather_inst<T, 3, WARPS, MINB>(a, st);
  if (nv <= 4) return launch_gather_inst<T, 4, WARPS, MINB>(a, st);
  if (nv <= 6) return launch_gather_inst<T, 6, WARPS, MINB>(a, st);
  if (nv <= 8) return launch_gather_inst<T, 8, WARPS, MINB>(a, st);
  if (a.res) return set_error(TOME_ERR_UNSUPPORTED, "tome_merge: the fused residual needs c <= %d", 8 * 32 * E);
  return launch_merge_scalar<T>(a, st);     // very wide rows: chunked generic kernel
}

int launch_merge(const tome_plan* plan, const void* x, int dtype, int c, const View& xv, const float* size_in,
                 int mode, float thr, void* out, const View& ov, float* size_out, float* logsize_out,
                 cudaStream_t st, const void* ln_w, const void* ln_b, float ln_eps, void* normed, const View* nv,
                 const void* residual, const View* rv) {
  MergeArgs a;
  a.res = residual;
  a.rv = rv ? *rv : xv;
  a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps; a.normed = normed; a.nv = nv ? *nv : ov;
  a.bm = plan->bm; a.n = plan->n; a.r = plan->r; a.distill = plan->distill_token; a.c = c; a.mode = mode;
  a.hybrid = (thr == thr) ? 1 : 0; a.thr = thr; a.node_max = plan->node_max;
  a.unm_idx = plan->unm_idx; a.b_off = plan->b_off; a.b_src = plan->b_src; a.b_head = plan->b_head;
  a.x = x; a.xv = xv; a.size_in = size_in; a.out = out; a.ov = ov; a.size_out = size_out; a.logsize_out = logsize_out;
  if (dtype == TOME_F32) {
    const bool vec = c % 4 == 0 && aligned16(x) && aligned16(out) && view_vec_ok(xv, 4) && view_vec_ok(ov, 4);
    if (!vec && normed) return set_error(TOME_ERR_ALIGN, "tome_merge_norm: fused LayerNorm needs 16-byte aligned rows (c %% 4 == 0)");
    if (residual && (!vec || !aligned16(residual) || !view_vec_ok(a.rv, 4))) return set_error(TOME_ERR_ALIGN, "tome_merge: the fused residual needs 16-byte aligned rows (c %% 4 == 0)");
    if (normed && (c > 4 * 32 * 8 || !aligned16(ln_w) || (ln_b && !aligned16(ln_b)) || !aligned16(normed) || !view_vec_ok(a.nv, 4)))
      return set_error(TOME_ERR_UNSUPPORTED, "tome_merge_norm: c=%d too wide or LayerNorm buffers misaligned", c);
    if (!vec) return launch_merge_scalar<float>(a, st);
    if (residual) return launch_merge_gather<float, 4, 4>(a, st);      // two input rows per output row: more registers
    return launch_merge_gather<float, 4, 6>(a, st);
  } else if (dtype == TOME_BF16) {
    const bool vec = c % 8 == 0 && aligned16(x) && aligned16(out) && view_vec_ok(xv, 8) && view_vec_ok(ov, 8);
    if (!vec && normed) return set_error(TOME_ERR_ALIGN, "tome_merge_norm: fused LayerNorm needs 16-byte aligned rows (c %% 8 == 0)");
    if (residual && (!vec || !aligned16(residual) || !view_vec_ok(a.rv, 8))) return set_error(TOME_ERR_ALIGN, "tome_merge: the fused residual needs 16-byte aligned rows (c %% 8 == 0)");
    if (normed && (c > 8 * 32 * 8 || !aligned16(ln_w) || (ln_b && !aligned16(ln_b)) || !aligned16(normed) || !view_vec_ok(a.nv, 8)))
      return set_error(TOME_ERR_UNSUPPORTED, "tome_merge_norm: c=%d too wide or LayerNorm buffers misaligned", c);
    if (!vec) return launch_merge_scalar<__nv_bfloat16>(a, st);
    // 64-thread CTAs at <= 85 registers: a CTA retires with its slowest warp, so small CTAs keep the copy warps
    // from queueing behind the few reducing ones, and the two-input (residual) rows do not spill.
    // merge 7.6 -> 7.3 us, + LayerNorm 12.5 -> 11.7, + residual 20.6 -> 15.6 (profiles/r01_merge_notes.md)
    return launch_merge_gather<__nv_bfloat16, 2, 12>(a, st);
  }
  return set_error(TOME_ERR_DTYPE, "tome_merge: unsupported dtype %d", dtype);
}

// ---- caller-side fusion: residual add + LayerNorm in one pass -------------------------------------
// The patched blocks end with  x = x + mlp(norm2(x))  and the next block starts with norm1(x)
// (tome/patch/videomae.py:17-22).  torch runs that as an add kernel plus a LayerNorm kernel (5.5 + 14.4 us
// at the bench shape); here one warp per row adds, writes the sum, and normalises it while it is
// still in registers.
template <typename T, int NV>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const T* __restrict__ a, const T* __restrict__ b2, long long b_rows,
                                                           long long rows, int c, const T* __restrict__ w,
                                                           const T* __restrict__ bias, float eps, T* __restrict__ sum_out,
                                                           T* __restrict__ normed) {
  constexpr int E = Pack<T>::E;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = c / E;
  const uint4* ar = reinterpret_cast<const uint4*>(a + row * c);
  const uint4* br = reinterpret_cast<const uint4*>(b2 + (row % b_rows) * c);
  uint4 va[NV], vb[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) { const int i = v * 32 + lane; if (i < nvec) { va[v] = ld_stream_u4(ar + i); vb[v] = ld_stream_u4(br + i); } }
  // the sum is rounded to T (what x + y holds in the reference), written out, and kept in registers as fp32 pairs for
  // the statistics; packed f32x2 arithmetic throughout (see add2 / fma2 above)
  float2 d[NV][E / 2];
  float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    const bool on = i < nvec;
    float fa[E], fb[E];
    if (on) { Pack<T>::unpack(va[v], fa); Pack<T>::unpack(vb[v], fb); }
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      const float2 t = on ? add2(make_float2(fa[e], fa[e + 1]), make_float2(fb[e], fb[e + 1])) : make_float2(0.f, 0.f);
      fa[e] = t.x; fa[e + 1] = t.y;
    }
    if (on) {
      va[v] = Pack<T>::pack(fa);
      reinterpret_cast<uint4*>(sum_out + row * c)[i] = va[v];
      Pack<T>::unpack(va[v], fa);
    }
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      d[v][e / 2] = on ? make_float2(fa[e], fa[e + 1]) : make_float2(0.f, 0.f);
      s2 = add2(s2, d[v][e / 2]);
    }
  }
  float sum = s2.x + s2.y;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
  const float mean = sum / (float)c;
  const float2 mean2 = make_float2(mean, mean);
  float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bool on = v * 32 + lane < nvec;
#pragma unroll
    for (int e = 0; e < E / 2; ++e) {
      d[v][e] = sub2(d[v][e], mean2);
      if (on) q2 = fma2(d[v][e], d[v][e], q2);
    }
  }
  float sq = q2.x + q2.y;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, of);
  const float rstd = rsqrtf(sq / (float)c + eps);
  const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    if (i < nvec) {
      float f[E], wf[E], bf[E];
      Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(w) + i), wf);
      if (bias) Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(bias) + i), bf);
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        const float2 g2 = mul2(make_float2(wf[e], wf[e + 1]), rstd2);
        const float2 y = fma2(d[v][e / 2], g2, bias ? make_float2(bf[e], bf[e + 1]) : make_float2(0.f, 0.f));
        f[e] = y.x; f[e + 1] = y.y;
      }
      reinterpret_cast<uint4*>(normed + row * c)[i] = Pack<T>::pack(f);
    }
  }
}

template <typename T>
static int launch_add_ln_t(const void* a, const void* b, long long b_rows, long long rows, int c, const void* w, const void* bias,
                           float eps, void* sum_out, void* normed, cudaStream_t st) {
  constexpr int E = Pack<T>::E;
  const int nv = (c / E + 31) / 32;
  static const int threads = [] {                   // tuning knob; 256: 13.4 us, 64: 12.5 us.  Whole warps, 32..256.
    const char* e = getenv("TOME_ADDLN_THREADS");
    int t = e ? atoi(e) : 64;
    t = (t / 32) * 32;
    return t < 32 ? 32 : (t > 256 ? 256 : t);
  }();
  const int rows_per_cta = threads / 32;
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
#define TOME_ADDLN(NV_) add_layernorm_kernel<T, NV_><<<grid, threads, 0, st>>>((const T*)a, (const T*)b, b_rows, rows, c, (const T*)w, (const T*)bias, eps, (T*)sum_out, (T*)normed)
  if (nv <= 1) TOME_ADDLN(1); else if (nv <= 2) TOME_ADDLN(2); else if (nv <= 3) TOME_ADDLN(3); else if (nv <= 4) TOME_ADDLN(4);
  else if (nv <= 6) TOME_ADDLN(6); else if (nv <= 8) TOME_ADDLN(8);
  else return set_error(TOME_ERR_UNSUPPORTED, "tome_add_layernorm: c=%d too wide", c);
#undef TOME_ADDLN
  TOME_LAUNCH_CHECK("add_layernorm_kernel");
  return TOME_OK;
}

int launch_add_layernorm(const void* a, const void* b, long long b_rows, int dtype, long long rows, int c, const void* w,
                         const void* bias, float eps, void* sum_out, void* normed, cudaStream_t st) {
  const int e = dtype == TOME_F32 ? 4 : 8;
  if (c % e != 0 || !aligned16(a) || !aligned16(b) || !aligned16(w) || (bias && !aligned16(bias)) || !aligned16(sum_out) || !aligned16(normed))
    return set_error(TOME_ERR_ALIGN, "tome_add_layernorm: needs 16-byte aligned buffers and c %% %d == 0", e);
  if (dtype == TOME_F32) return launch_add_ln_t<float>(a, b, b_rows, rows, c, w, bias, eps, sum_out, normed, st);
  if (dtype == TOME_BF16) return launch_add_ln_t<__nv_bfloat16>(a, b, b_rows, rows, c, w, bias, eps, sum_out, normed, st);
  return set_error(TOME_ERR_DTYPE, "tome_add_layernorm: unsupported dtype %d", dtype);
}

// ---- caller-side fusion for the divided space-time blocks: add + LayerNorm through (b, p, t) row views --------
// TimeSformer / Motionformer keep tokens as 'b (p t) m' behind a class token and hop between three row orders per
// block: '(b p) t' for the temporal attention, '(b t) (1 + p)' for the spatial attention, 'b (1 + p t)' for the
// residual stream (tome/patch/timesformer.py:38-56, slowfast/models/timesformer.py:115-153).  The reference pays a
// rearrange / cat copy at every hop plus separate add and LayerNorm passes.  Here row (b, p, t) of each tensor is
// base + b*sb + p*sp + t*st, so one pass reads a (+ b), writes the sum in the layout of the residual stream and the
// LayerNorm in the layout of the NEXT consumer.  Arithmetic identical to add_layernorm_kernel.
struct Rows3 { long long sb, sp, st; };

template <typename T, int NV>
__global__ void __launch_bounds__(256) rows3_add_layernorm_kernel(const T* __restrict__ a, Rows3 av, const T* __restrict__ b2, Rows3 bv,
                                                                 int P, int Tn, long long rows, int c, const T* __restrict__ w,
                                                                 const T* __restrict__ bias, float eps, T* __restrict__ sum_out,
                                                                 Rows3 sv, T* __restrict__ normed, Rows3 nv) {
  constexpr int E = Pack<T>::E;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long bp = row / Tn;
  const int t = (int)(row - bp * Tn);
  const long long bb = bp / P;
  const int pp = (int)(bp - bb * P);
  const int nvec = c / E;
  const uint4* ar = reinterpret_cast<const uint4*>(a + bb * av.sb + pp * av.sp + t * av.st);
  const uint4* br = b2 ? reinterpret_cast<const uint4*>(b2 + bb * bv.sb + pp * bv.sp + t * bv.st) : nullptr;
  uint4 va[NV], vb[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    if (i < nvec) { va[v] = ld_stream_u4(ar + i); if (br) vb[v] = ld_stream_u4(br + i); }
  }
  float2 d[NV][E / 2];
  float2 s2 = make_float2(0.f, 0.f);
  uint4* srow = sum_out ? reinterpret_cast<uint4*>(sum_out + bb * sv.sb + pp * sv.sp + t * sv.st) : nullptr;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    const bool on = i < nvec;
    float fa[E], fb[E];
    if (on) {
      Pack<T>::unpack(va[v], fa);
      if (br) {
        Pack<T>::unpack(vb[v], fb);
#pragma unroll
        for (int e = 0; e < E; e += 2) {
          const float2 s = add2(make_float2(fa[e], fa[e + 1]), make_float2(fb[e], fb[e + 1]));
          fa[e] = s.x; fa[e + 1] = s.y;
        }
        va[v] = Pack<T>::pack(fa);               // the sum rounded to T, as x + y holds it
        Pack<T>::unpack(va[v], fa);
      }
      if (srow) srow[i] = va[v];
    }
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      d[v][e / 2] = on ? make_float2(fa[e], fa[e + 1]) : make_float2(0.f, 0.f);
      s2 = add2(s2, d[v][e / 2]);
    }
  }
  if (!normed) return;
  float sum = s2.x + s2.y;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
  const float mean = sum / (float)c;
  const float2 mean2 = make_float2(mean, mean);
  float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bool on = v * 32 + lane < nvec;
#pragma unroll
    for (int e = 0; e < E / 2; ++e) {
      d[v][e] = sub2(d[v][e], mean2);
      if (on) q2 = fma2(d[v][e], d[v][e], q2);
    }
  }
  float sq = q2.x + q2.y;
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, of);
  const float rstd = rsqrtf(sq / (float)c + eps);
  const float2 rstd2 = make_float2(rstd, rstd);
  uint4* nrow = reinterpret_cast<uint4*>(normed + bb * nv.sb + pp * nv.sp + t * nv.st);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    if (i < nvec) {
      float f[E], wf[E], bf[E];
      Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(w) + i), wf);
      if (bias) Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(bias) + i), bf);
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        const float2 g2 = mul2(make_float2(wf[e], wf[e + 1]), rstd2);
        const float2 y = fma2(d[v][e / 2], g2, bias ? make_float2(bf[e], bf[e + 1]) : make_float2(0.f, 0.f));
        f[e] = y.x; f[e + 1] = y.y;
      }
      nrow[i] = Pack<T>::pack(f);
    }
  }
}

template <typename T>
static int launch_rows3_t(const void* a, const Rows3& av, const void* b, const Rows3& bv, int P, int Tn, long long rows, int c,
                          const void* w, const void* bias, float eps, void* sum_out, const Rows3& sv, void* normed, const Rows3& nv,
                          cudaStream_t st) {
  constexpr int E = Pack<T>::E;
  const int nvv = (c / E + 31) / 32;
  const int threads = 64, rows_per_cta = threads / 32;
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
#define TOME_ROWS3(NV_) rows3_add_layernorm_kernel<T, NV_><<<grid, threads, 0, st>>>((const T*)a, av, (const T*)b, bv, P, Tn, rows, c, (const T*)w, (const T*)bias, eps, (T*)sum_out, sv, (T*)normed, nv)
  if (nvv <= 1) TOME_ROWS3(1); else if (nvv <= 2) TOME_ROWS3(2); else if (nvv <= 3) TOME_ROWS3(3); else if (nvv <= 4) TOME_ROWS3(4);
  else if (nvv <= 6) TOME_ROWS3(6); else if (nvv <= 8) TOME_ROWS3(8);
  else return set_error(TOME_ERR_UNSUPPORTED, "tome_rows_add_layernorm: c=%d too wide", c);
#undef TOME_ROWS3
  TOME_LAUNCH_CHECK("rows3_add_layernorm_kernel");
  return TOME_OK;
}

static bool rows3_ok(const long long* v, int e) { return v[0] % e == 0 && v[1] % e == 0 && v[2] % e == 0; }

int launch_rows_add_layernorm(const void* a, const long long* av, const void* b, const long long* bv, int dtype, int B, int P, int Tn, int c,
                              const void* w, const void* bias, float eps, void* sum_out, const long long* sv, void* normed,
                              const long long* nv, cudaStream_t st) {
  const int e = dtype == TOME_F32 ? 4 : 8;
  const long long zero[3] = {0, 0, 0};
  if (!b) bv = zero;
  if (!sum_out) sv = zero;
  if (!normed) nv = zero;
  if (c % e != 0 || !aligned16(a) || (b && !aligned16(b)) || (normed && (!aligned16(w) || (bias && !aligned16(bias)) || !aligned16(normed))) ||
      (sum_out && !aligned16(sum_out)) || !rows3_ok(av, e) || !rows3_ok(bv, e) || !rows3_ok(sv, e) || !rows3_ok(nv, e))
    return set_error(TOME_ERR_ALIGN, "tome_rows_add_layernorm: needs 16-byte aligned rows and c %% %d == 0", e);
  const Rows3 A{av[0], av[1], av[2]}, Bv{bv[0], bv[1], bv[2]}, S{sv[0], sv[1], sv[2]}, N{nv[0], nv[1], nv[2]};
  const long long rows = (long long)B * P * Tn;
  if (dtype == TOME_F32) return launch_rows3_t<float>(a, A, b, Bv, P, Tn, rows, c, w, bias, eps, sum_out, S, normed, N, st);
  if (dtype == TOME_BF16) return launch_rows3_t<__nv_bfloat16>(a, A, b, Bv, P, Tn, rows, c, w, bias, eps, sum_out, S, normed, N, st);
  return set_error(TOME_ERR_DTYPE, "tome_rows_add_layernorm: unsupported dtype %d", dtype);
}

// ---- class-token rows of the divided space-time blocks -----------------------------------------------------------
// The class token does not follow the (b, p, t) pattern of the patch tokens: per block the reference replicates its
// LayerNorm over the frames, averages the spatial attention's T class outputs back into one row, and carries it through
// the residual adds (tome/patch/timesformer.py:41-48, 56).  In torch that is ~10 launches on B rows per block; here one
// warp per clip does  sum = a (+ add) (+ mean over t of m[b, t]),  each step rounded to T as the separate ops round,
// writes it, and writes LayerNorm(sum) to `reps` destinations (the per-frame copies).
struct ClsArgs {
  const void *a, *add, *mean_src;
  long long a_sb, add_sb, m_sb, m_st;      // element strides: batch (and frame for mean_src)
  int mean_t;                              // frames averaged (0: no mean term)
  void *sum_out, *normed_out;
  long long s_sb, n_sb, n_sr;              // sum_out batch stride; normed_out batch / replica strides
  int reps, c;
  const void *ln_w, *ln_b;
  float eps;
};

template <typename T, int NV>
__global__ void __launch_bounds__(32) cls_rows_kernel(ClsArgs g) {
  constexpr int E = Pack<T>::E;
  const int lane = threadIdx.x, b = blockIdx.x, nvec = g.c / E;
  float f[NV][E];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    if (i < nvec) Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(g.a) + (long long)b * g.a_sb) + i), f[v]);
    else
#pragma unroll
      for (int e = 0; e < E; ++e) f[v][e] = 0.f;
  }
  auto round_t = [&]() {                               // what a T-typed intermediate tensor would hold
#pragma unroll
    for (int v = 0; v < NV; ++v) Pack<T>::unpack(Pack<T>::pack(f[v]), f[v]);
  };
  if (g.mean_t > 0) {                                  // mean over frames: fp32 accumulation, one rounding (ATen's MeanOps)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = v * 32 + lane;
      if (i < nvec) {
        float acc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = 0.f;
        for (int t = 0; t < g.mean_t; ++t) {
          float m[E];
          Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(g.mean_src) + (long long)b * g.m_sb + (long long)t * g.m_st) + i), m);
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] += m[e];
        }
        const float inv = 1.0f / (float)g.mean_t;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] *= inv;
        Pack<T>::unpack(Pack<T>::pack(acc), acc);      // the mean tensor, rounded to T
#pragma unroll
        for (int e = 0; e < E; ++e) f[v][e] += acc[e];
      }
    }
    round_t();
  }
  if (g.add) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = v * 32 + lane;
      if (i < nvec) {
        float m[E];
        Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(g.add) + (long long)b * g.add_sb) + i), m);
#pragma unroll
        for (int e = 0; e < E; ++e) f[v][e] += m[e];
      }
    }
    round_t();
  }
  if (g.sum_out) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int i = v * 32 + lane;
      if (i < nvec) reinterpret_cast<uint4*>(reinterpret_cast<T*>(g.sum_out) + (long long)b * g.s_sb)[i] = Pack<T>::pack(f[v]);
    }
  }
  if (!g.normed_out) return;
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < E; ++e) sum += f[v][e];
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
  const float mean = sum / (float)g.c;
  float sq = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bool on = v * 32 + lane < nvec;
#pragma unroll
    for (int e = 0; e < E; ++e) { f[v][e] -= mean; if (on) sq = fmaf(f[v][e], f[v][e], sq); }
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, of);
  const float rstd = rsqrtf(sq / (float)g.c + g.eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int i = v * 32 + lane;
    if (i < nvec) {
      float w[E], bb[E], y[E];
      Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(g.ln_w) + i), w);
      if (g.ln_b) Pack<T>::unpack(__ldg(reinterpret_cast<const uint4*>(g.ln_b) + i), bb);
#pragma unroll
      for (int e = 0; e < E; ++e) y[e] = fmaf(f[v][e], w[e] * rstd, g.ln_b ? bb[e] : 0.f);
      const uint4 packed = Pack<T>::pack(y);
      for (int rr = 0; rr < g.reps; ++rr)
        reinterpret_cast<uint4*>(reinterpret_cast<T*>(g.normed_out) + (long long)b * g.n_sb + (long long)rr * g.n_sr)[i] = packed;
    }
  }
}

int launch_cls_rows(const ClsArgs& g, int dtype, int batch, cudaStream_t st) {
  const int e = dtype == TOME_F32 ? 4 : 8;
  const int nv = (g.c / e + 31) / 32;
  if (g.c % e != 0 || nv > 8) return set_error(TOME_ERR_UNSUPPORTED, "tome_cls_rows: c=%d (needs c %% %d == 0, c <= %d)", g.c, e, 8 * 32 * e);
#define TOME_CLS(T_, NV_) cls_rows_kernel<T_, NV_><<<batch, 32, 0, st>>>(g)
  if (dtype == TOME_F32) { if (nv <= 2) TOME_CLS(float, 2); else if (nv <= 4) TOME_CLS(float, 4); else if (nv <= 6) TOME_CLS(float, 6); else TOME_CLS(float, 8); }
  else if (dtype == TOME_BF16) { if (nv <= 2) TOME_CLS(__nv_bfloat16, 2); else if (nv <= 3) TOME_CLS(__nv_bfloat16, 3); else if (nv <= 6) TOME_CLS(__nv_bfloat16, 6); else TOME_CLS(__nv_bfloat16, 8); }
  else return set_error(TOME_ERR_DTYPE, "tome_cls_rows: unsupported dtype %d", dtype);
#undef TOME_CLS
  TOME_LAUNCH_CHECK("cls_rows_kernel");
  return TOME_OK;
}

int launch_merge_source(const tome_plan* plan, const float* source, int n0, float thr, float* out, cudaStream_t st) {
  const int nout = plan->n - plan->r;
  if (source) {
    View xv{(long long)plan->n * n0, 0, n0, 1}, ov{(long long)nout * n0, 0, n0, 1};
    return launch_merge(plan, source, TOME_F32, n0, xv, nullptr, TOME_MODE_AMAX, thr, out, ov, nullptr, nullptr, st, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr);
  }
  MergeArgs a{};
  a.bm = plan->bm; a.n = plan->n; a.r = plan->r; a.distill = plan->distill_token; a.c = plan->n;
  a.mode = TOME_MODE_AMAX; a.hybrid = (thr == thr) ? 1 : 0; a.thr = thr; a.node_max = plan->node_max;
  a.unm_idx = plan->unm_idx; a.b_off = plan->b_off; a.b_src = plan->b_src; a.b_head = plan->b_head; a.out = out;
  dim3 grid((nout + 7) / 8, plan->bm);
  source_identity_kernel<<<grid, 256, 0, st>>>(a);
  TOME_LAUNCH_CHECK("source_identity_kernel");
  return TOME_OK;
}

int launch_unmerge(const tome_plan* plan, const void* x, int dtype, int c, void* out, cudaStream_t st) {
  // NB: the reference's unmerge (merge.py:87-100) splits x as [unm | dst] even when a
  // distill token made merge() interleave them (merge.py:82-83); kept as is, so the slot
  // mapping here never uses the distill layout.
  dim3 grid((plan->n + 7) / 8, plan->bm);
  if (dtype == TOME_F32) {
    if (c % 4 == 0 && aligned16(x) && aligned16(out))
      unmerge_rows_kernel<float, 4><<<grid, 256, 0, st>>>(plan->a_map, plan->n, plan->r, 0, c, (const float*)x, (float*)out);
    else
      unmerge_rows_kernel<float, 1><<<grid, 256, 0, st>>>(plan->a_map, plan->n, plan->r, 0, c, (const float*)x, (float*)out);
  } else if (dtype == TOME_BF16) {
    if (c % 8 == 0 && aligned16(x) && aligned16(out))
      unmerge_rows_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(plan->a_map, plan->n, plan->r, 0, c, (const __nv_bfloat16*)x, (__nv_bfloat16*)out);
    else
      unmerge_rows_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>(plan->a_map, plan->n, plan->r, 0, c, (const __nv_bfloat16*)x, (__nv_bfloat16*)out);
  } else return set_error(TOME_ERR_DTYPE, "tome_unmerge: unsupported dtype %d", dtype);
  TOME_LAUNCH_CHECK("unmerge_rows_kernel");
  return TOME_OK;
}

}  // namespace tome
