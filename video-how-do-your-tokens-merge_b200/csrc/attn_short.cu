// Attention over SHORT sequences (the caller right in front of the hot path): TimeSformer's temporal
// attention runs on '(b p) t' rows -- B*P = 1568 sequences of T = 8 tokens per clip batch, 12 heads
// (slowfast/models/timesformer.py:118-123; un-patched by ToMe, tome/patch/timesformer.py:224).  A flash kernel tiles
// queries and keys by 64-128 and spends 111 us per call on 18 816 problems of 8 x 8 (ncu launch list,
// profiles/r02_timesformer_launches.txt); the arithmetic is 0.3 GFLOP and the traffic 77 MB.
//
// Here ONE THREAD owns one query row of one (sequence, head): q stays in registers, the T keys and values of its
// group are read straight from the QKV GEMM's output (the T threads of a group read the same lines: L1 serves the
// repeats), scores / softmax / output are fp32 with a two-pass softmax over the <= 32 scores held in registers, and
// the row is written once in the '(b p) t (h d)' layout the projection GEMM reads.  No shared memory, no tensor
// cores: the kernel is a 77 MB streaming pass.
#include "common.cuh"

namespace tome {

constexpr int AS_MAX_T = 32;

template <typename T> struct Ld8;
template <> struct Ld8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
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
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Ld8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  static __device__ __forceinline__ float to_float(float v) { return v; }
  static __device__ __forceinline__ float from_float(float v) { return v; }
};

// q / k / v element (s, t, h, c) at base + s * ss + t * st + h * D + c; out (s, t, h, c) contiguous.
// FOUR threads per (sequence, head, query), 16 channels each (adjacent lanes: a quad reads one 128-byte head row per key,
// partial dot products meet in two shuffles).  With one thread per query the 64-channel q and accumulator rows cost ~150
// registers, twelve warps per SM were resident and the 77 MB streaming pass ran at a third of the HBM rate (35 us).
template <typename T, int D, int MAXT>                    // MAXT: 8 (TimeSformer's frames: every loop unrolled, scores indexed statically) or AS_MAX_T
__global__ void __launch_bounds__(128) attn_short_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                         long long ss, long long st, long long seqs, int heads, int tn, float scale,
                                                         T* __restrict__ out) {
  constexpr int DP = D / 4;                              // channels per thread
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = seqs * heads * tn;
  const int part = (int)(gid & 3);
  const bool valid = (gid >> 2) < total;
  const long long quad = valid ? (gid >> 2) : total - 1;   // idle quads of the last warp recompute the last row (shuffles stay full-mask)
  const int i = (int)(quad % tn);
  const long long sh = quad / tn;
  const int h = (int)(sh % heads);
  const long long s = sh / heads;
  const long long base = s * ss + (long long)h * D + part * DP;
  float qr[DP];
#pragma unroll
  for (int c = 0; c < DP; c += 8) {
    float f[8];
    Ld8<T>::load(q + base + (long long)i * st + c, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[c + e] = f[e] * scale;
  }
  float sc[MAXT];
  float m = -INFINITY;
  auto score = [&](int j) {                                // q . k_j over this thread's channels, summed over the quad
    const T* kr = k + base + (long long)j * st;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c = 0; c < DP; c += 8) {
      float f[8];
      Ld8<T>::load(kr + c, f);
      a0 = fmaf(qr[c], f[0], a0); a1 = fmaf(qr[c + 1], f[1], a1); a2 = fmaf(qr[c + 2], f[2], a2); a3 = fmaf(qr[c + 3], f[3], a3);
      a0 = fmaf(qr[c + 4], f[4], a0); a1 = fmaf(qr[c + 5], f[5], a1); a2 = fmaf(qr[c + 6], f[6], a2); a3 = fmaf(qr[c + 7], f[7], a3);
    }
    float d = (a0 + a1) + (a2 + a3);
    d += __shfl_xor_sync(0xffffffffu, d, 1);               // the quad's four partial sums: every lane ends with the same total
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    return d;
  };
  if constexpr (MAXT <= 8) {
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {                       // tn is warp-uniform: the shuffles stay converged
      sc[j] = j < tn ? score(j) : -INFINITY;
      m = fmaxf(m, sc[j]);
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < tn; ++j) {
      const float d = score(j);
#pragma unroll
      for (int u = 0; u < MAXT; ++u) if (u == j) sc[u] = d;      // static indexing keeps sc[] in registers
      m = fmaxf(m, d);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int u = 0; u < MAXT; ++u) {
    if (u < tn) { sc[u] = sizeof(T) == 4 ? expf(sc[u] - m) : __expf(sc[u] - m); l += sc[u]; }
  }
  const float inv = 1.0f / l;
  float acc[DP];
#pragma unroll
  for (int c = 0; c < DP; ++c) acc[c] = 0.f;
  auto accumulate = [&](int j, float p) {
    const T* vr = v + base + (long long)j * st;
#pragma unroll
    for (int c = 0; c < DP; c += 8) {
      float f[8];
      Ld8<T>::load(vr + c, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c + e] = fmaf(p, f[e], acc[c + e]);
    }
  };
  if constexpr (MAXT <= 8) {
#pragma unroll
    for (int j = 0; j < MAXT; ++j) if (j < tn) accumulate(j, sc[j] * inv);
  } else {
#pragma unroll 1
    for (int j = 0; j < tn; ++j) {
      float p = 0.f;
#pragma unroll
      for (int u = 0; u < MAXT; ++u) if (u == j) p = sc[u];
      accumulate(j, p * inv);
    }
  }
  if (!valid) return;
  T* orow = out + ((s * tn + i) * heads + h) * D + part * DP;
#pragma unroll
  for (int c = 0; c < DP; c += 8) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = acc[c + e];
    Ld8<T>::store(orow + c, f);
  }
}

int launch_attn_short(const void* q, const void* k, const void* v, int dtype, long long seqs, int tn, int heads, int d, long long ss,
                      long long st, float scale, void* out, cudaStream_t stm) {
  if (tn < 1 || tn > AS_MAX_T) return set_error(TOME_ERR_UNSUPPORTED, "tome_attn_short: %d tokens per sequence (1..%d)", tn, AS_MAX_T);
  if (d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_attn_short: head dim %d (64 only)", d);
  const int e = dtype == TOME_F32 ? 4 : 8;     // 16-byte loads
  if (((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15) || ss % e || st % e)
    return set_error(TOME_ERR_ALIGN, "tome_attn_short: q / k / v / out must be 16-byte aligned with 16-byte aligned strides");
  const long long total = seqs * heads * tn * 4;         // four threads per (sequence, head, query)
  const unsigned grid = (unsigned)((total + 127) / 128);
#define TOME_AS(T_, M_) attn_short_kernel<T_, 64, M_><<<grid, 128, 0, stm>>>((const T_*)q, (const T_*)k, (const T_*)v, ss, st, seqs, heads, tn, scale, (T_*)out)
  if (dtype == TOME_BF16) { if (tn <= 8) TOME_AS(__nv_bfloat16, 8); else TOME_AS(__nv_bfloat16, AS_MAX_T); }
  else if (dtype == TOME_F32) { if (tn <= 8) TOME_AS(float, 8); else TOME_AS(float, AS_MAX_T); }
  else return set_error(TOME_ERR_DTYPE, "tome_attn_short: unsupported dtype %d", dtype);
#undef TOME_AS
  TOME_LAUNCH_CHECK("attn_short_kernel");
  return TOME_OK;
}

// ---- one query against a whole sequence: Motionformer's class token (vit_helper.py:181-189: `cls_out`), which attends to all
// 1 + F * P tokens while the patch queries go through the per-frame kernel.  The library's fused attention spends 86 us per
// layer on that single row in fp32 (a 64 x 64-tile kernel with one live query); it is a 77 MB streaming pass: one CTA per
// (clip, head), scores of all keys in shared memory, then channel c of the output by thread c of each of four key groups.
constexpr int CA_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(CA_THREADS) cls_attn_kernel(const T* __restrict__ qkv, int n, int heads, int q_tok, float scale,
                                                             T* __restrict__ out) {
  extern __shared__ float ca_sm[];                     // [n scores][CA_THREADS reduction scratch]
  float* sc = ca_sm;
  float* red = ca_sm + n;
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads, tid = threadIdx.x;
  const int C = heads * 64;
  const long long row = 3LL * C;
  const T* base = qkv + (long long)b * n * row;
  float q[64];
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    float f[8];
    Ld8<T>::load(base + (long long)q_tok * row + h * 64 + c, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) q[c + e] = f[e] * scale;
  }
  float m = -INFINITY;
  for (int t = tid; t < n; t += CA_THREADS) {
    const T* kr = base + (long long)t * row + C + h * 64;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      float f[8];
      Ld8<T>::load(kr + c, f);
#pragma unroll
      for (int e = 0; e < 8; e += 2) { a0 = fmaf(q[c + e], f[e], a0); a1 = fmaf(q[c + e + 1], f[e + 1], a1); }
    }
    const float d = a0 + a1;
    sc[t] = d;
    m = fmaxf(m, d);
  }
  red[tid] = m;
  __syncthreads();
  for (int of = CA_THREADS / 2; of > 0; of >>= 1) { if (tid < of) red[tid] = fmaxf(red[tid], red[tid + of]); __syncthreads(); }
  m = red[0];
  __syncthreads();
  float l = 0.f;
  for (int t = tid; t < n; t += CA_THREADS) { const float e = __expf(sc[t] - m); sc[t] = e; l += e; }
  red[tid] = l;
  __syncthreads();
  for (int of = CA_THREADS / 2; of > 0; of >>= 1) { if (tid < of) red[tid] += red[tid + of]; __syncthreads(); }
  const float inv = 1.0f / red[0];
  __syncthreads();
  // out[c] = sum_t p_t v[t][c]: thread (c, g) takes the keys t = g mod 4: 64 consecutive channels per key row
  const int c = tid & 63, g = tid >> 6;
  float acc = 0.f;
  for (int t = g; t < n; t += CA_THREADS / 64) acc = fmaf(sc[t], Ld8<T>::to_float(base[(long long)t * row + 2 * C + h * 64 + c]), acc);
  red[tid] = acc;
  __syncthreads();
  if (tid < 64) {
    const float y = ((red[tid] + red[tid + 64]) + (red[tid + 128] + red[tid + 192])) * inv;
    out[(long long)b * C + h * 64 + tid] = Ld8<T>::from_float(y);
  }
}

int launch_cls_attention(const void* qkv, int dtype, int b, int n, int heads, int q_tok, float scale, void* out, cudaStream_t stm) {
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 3)) return set_error(TOME_ERR_ALIGN, "tome_cls_attention: qkv must be 16-byte aligned");
  if (q_tok < 0 || q_tok >= n) return set_error(TOME_ERR_ARG, "tome_cls_attention: query token %d outside 0..%d", q_tok, n - 1);
  const size_t smem = ((size_t)n + CA_THREADS) * sizeof(float);
  if (smem > 200 * 1024) return set_error(TOME_ERR_UNSUPPORTED, "tome_cls_attention: %d tokens", n);
  if (dtype == TOME_F32) {
    static PerDeviceOnce once;
    if (once.first_time()) TOME_CUDA(cudaFuncSetAttribute(cls_attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cls_attn_kernel<float><<<b * heads, CA_THREADS, smem, stm>>>((const float*)qkv, n, heads, q_tok, scale, (float*)out);
  } else if (dtype == TOME_BF16) {
    static PerDeviceOnce once;
    if (once.first_time()) TOME_CUDA(cudaFuncSetAttribute(cls_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cls_attn_kernel<__nv_bfloat16><<<b * heads, CA_THREADS, smem, stm>>>((const __nv_bfloat16*)qkv, n, heads, q_tok, scale, (__nv_bfloat16*)out);
  } else {
    return set_error(TOME_ERR_DTYPE, "tome_cls_attention: unsupported dtype %d", dtype);
  }
  TOME_LAUNCH_CHECK("cls_attn_kernel");
  return TOME_OK;
}

}  // namespace tome
