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
};

// q / k / v element (s, t, h, c) at base + s * ss + t * st + h * D + c; out (s, t, h, c) contiguous.
template <typename T, int D>
__global__ void __launch_bounds__(128) attn_short_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                         long long ss, long long st, long long seqs, int heads, int tn, float scale,
                                                         T* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = seqs * heads * tn;
  if (gid >= total) return;
  const int i = (int)(gid % tn);
  const long long sh = gid / tn;
  const int h = (int)(sh % heads);
  const long long s = sh / heads;
  const long long base = s * ss + (long long)h * D;
  float qr[D];
#pragma unroll
  for (int c = 0; c < D; c += 8) {
    float f[8];
    Ld8<T>::load(q + base + (long long)i * st + c, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[c + e] = f[e] * scale;
  }
  float sc[AS_MAX_T];
  float m = -INFINITY;
#pragma unroll 1
  for (int j = 0; j < tn; ++j) {
    const T* kr = k + base + (long long)j * st;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float f[8];
      Ld8<T>::load(kr + c, f);
      a0 = fmaf(qr[c], f[0], a0); a1 = fmaf(qr[c + 1], f[1], a1); a2 = fmaf(qr[c + 2], f[2], a2); a3 = fmaf(qr[c + 3], f[3], a3);
      a0 = fmaf(qr[c + 4], f[4], a0); a1 = fmaf(qr[c + 5], f[5], a1); a2 = fmaf(qr[c + 6], f[6], a2); a3 = fmaf(qr[c + 7], f[7], a3);
    }
    const float d = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int u = 0; u < AS_MAX_T; ++u) if (u == j) sc[u] = d;      // static indexing keeps sc[] in registers
    m = fmaxf(m, d);
  }
  float l = 0.f;
#pragma unroll
  for (int u = 0; u < AS_MAX_T; ++u) {
    if (u < tn) { sc[u] = sizeof(T) == 4 ? expf(sc[u] - m) : __expf(sc[u] - m); l += sc[u]; }
  }
  const float inv = 1.0f / l;
  float acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int j = 0; j < tn; ++j) {
    float p = 0.f;
#pragma unroll
    for (int u = 0; u < AS_MAX_T; ++u) if (u == j) p = sc[u];
    p *= inv;
    const T* vr = v + base + (long long)j * st;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float f[8];
      Ld8<T>::load(vr + c, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c + e] = fmaf(p, f[e], acc[c + e]);
    }
  }
  T* orow = out + ((s * tn + i) * heads + h) * D;
#pragma unroll
  for (int c = 0; c < D; c += 8) {
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
  const long long total = seqs * heads * tn;
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (dtype == TOME_BF16)
    attn_short_kernel<__nv_bfloat16, 64><<<grid, 128, 0, stm>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                                 ss, st, seqs, heads, tn, scale, (__nv_bfloat16*)out);
  else if (dtype == TOME_F32)
    attn_short_kernel<float, 64><<<grid, 128, 0, stm>>>((const float*)q, (const float*)k, (const float*)v, ss, st, seqs, heads, tn, scale,
                                                         (float*)out);
  else
    return set_error(TOME_ERR_DTYPE, "tome_attn_short: unsupported dtype %d", dtype);
  TOME_LAUNCH_CHECK("attn_short_kernel");
  return TOME_OK;
}

}  // namespace tome
