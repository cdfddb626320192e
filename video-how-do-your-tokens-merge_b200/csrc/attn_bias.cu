// Caller-side piece of proportional attention.
//
// The patched attentions add log(size) of the KEY token to every logit
// (tome/patch/videomae.py:62-63, vivit.py:103-104, timesformer.py:72-74).  A key-only bias b_j is a rank-1
// term:  scale * (q.k_j) + b_j  ==  scale * ([q, 1, 1] . [k_j, hi_j, lo_j])  with  hi_j + lo_j = b_j / scale.
// The host pads each q/k head with spare channels (zero weight rows in the cached QKV weight; q's two
// bias channels come out of the GEMM as 1), and this kernel drops hi/lo into k's two channels.  The
// fused flash attention then sees no mask at all.  hi/lo is an exact two-term split in the tensor's
// dtype (bf16: 16 mantissa bits of b_j / scale), so the bias is more accurate than the reference's own
// bf16 `attn + size.log()`.
#include "common.cuh"

namespace tome {

struct KeyBiasArgs {
  const float* log_size;       // (b, n - lead) fp32
  int b, n, lead, heads, d;
  float inv_scale;
  void* k; long long k_sb, k_sn, k_sh;
  void* q; long long q_sb, q_sn, q_sh;     // q may be NULL: channels d, d+1 already hold 1 (from the GEMM bias)
};

template <typename T> struct Two;
template <> struct Two<float> {
  static __device__ __forceinline__ void store(float* p, float hi, float lo) { *reinterpret_cast<float2*>(p) = make_float2(hi, lo); }
  static __device__ __forceinline__ void split(float t, float& hi, float& lo) { hi = t; lo = 0.f; }
};
template <> struct Two<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float hi, float lo) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(hi, lo);
  }
  static __device__ __forceinline__ void split(float t, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(t));
    lo = t - hi;                                   // exact; rounded to bf16 by the store
  }
};

template <typename T>
__global__ void __launch_bounds__(256) key_bias_kernel(KeyBiasArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.b * a.n * a.heads;
  if (i >= total) return;
  const int h = (int)(i % a.heads);
  const long long bt = i / a.heads;
  const int t = (int)(bt % a.n), b = (int)(bt / a.n);
  float hi = 0.f, lo = 0.f;
  if (t >= a.lead) Two<T>::split(__ldg(a.log_size + (long long)b * (a.n - a.lead) + (t - a.lead)) * a.inv_scale, hi, lo);
  Two<T>::store(reinterpret_cast<T*>(a.k) + b * a.k_sb + t * a.k_sn + h * a.k_sh + a.d, hi, lo);
  if (a.q) {
    const float one = t >= a.lead ? 1.f : 0.f;     // leading (class) queries take no bias: timesformer.py:74
    Two<T>::store(reinterpret_cast<T*>(a.q) + b * a.q_sb + t * a.q_sn + h * a.q_sh + a.d, one, one);
  }
}

int launch_key_bias(const float* log_size, int b, int n, int lead, int heads, int d, float inv_scale, int dtype, void* k,
                    long long k_sb, long long k_sn, long long k_sh, void* q, long long q_sb, long long q_sn, long long q_sh,
                    cudaStream_t st) {
  KeyBiasArgs a{log_size, b, n, lead, heads, d, inv_scale, k, k_sb, k_sn, k_sh, q, q_sb, q_sn, q_sh};
  const int e = dtype == TOME_F32 ? 4 : 2;
  const uintptr_t two = 2 * e;
  if ((d & 1) || ((uintptr_t)k % two) || (k_sb & 1) || (k_sn & 1) || (k_sh & 1) ||
      (q && (((uintptr_t)q % two) || (q_sb & 1) || (q_sn & 1) || (q_sh & 1))))
    return set_error(TOME_ERR_ALIGN, "tome_attn_key_bias: channel index and strides must be even");
  const long long total = (long long)b * n * heads;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == TOME_F32) key_bias_kernel<float><<<grid, 256, 0, st>>>(a);
  else if (dtype == TOME_BF16) key_bias_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else return set_error(TOME_ERR_DTYPE, "tome_attn_key_bias: unsupported dtype %d", dtype);
  TOME_LAUNCH_CHECK("key_bias_kernel");
  return TOME_OK;
}

}  // namespace tome
