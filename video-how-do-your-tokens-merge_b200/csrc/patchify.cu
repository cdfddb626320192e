// Caller-side data format: clip -> tubelet rows.
//
// The four models open with a Conv3d/Conv2d whose kernel equals its stride
// (slowfast/models/videomae_video_model_builder.py:138-160), i.e. a GEMM over non-overlapping tubelets.
// torch expresses the regrouping as an 8-d permute + copy, which its generic strided-copy kernel runs
// at a sixth of HBM speed (76 us for 8 clips, 5 % of the whole VideoMAE-B forward; profiles/r01b).  Here
// one thread moves 8 consecutive pixels of an image row (16 B of bf16 / 32 B of fp32, coalesced reads)
// to its place in the (tokens, c * tt * ph * pw) matrix (full 32-byte sectors), converting fp32 clips to
// the model dtype on the way, so the cast pass disappears too.
#include "common.cuh"

namespace tome {

struct PatchifyArgs {
  const void* x; void* out;
  int b, c, t, h, w, tt, ph, pw;
};

template <typename TI> struct In8;
template <> struct In8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(p)), b = ld_stream_f4(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <> struct In8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};
// uint8 frames (what a video decoder hands over): value / 255, the division rounded like torch's `x.float() / 255`
template <> struct In8<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* p, float (&f)[8]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[i] = __fdiv_rn((float)((v.x >> (8 * i)) & 0xFFu), 255.0f);
      f[4 + i] = __fdiv_rn((float)((v.y >> (8 * i)) & 0xFFu), 255.0f);
    }
  }
};
template <typename TO> struct Out8;
template <> struct Out8<float> {
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Out8<__nv_bfloat16> {
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

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) patchify_kernel(PatchifyArgs a) {
  const int w8 = a.w >> 3;
  const long long total = (long long)a.b * a.c * a.t * a.h * w8;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const int x8 = (int)(id % w8);
  long long rest = id / w8;
  const int y = (int)(rest % a.h); rest /= a.h;
  const int t = (int)(rest % a.t); rest /= a.t;
  const int c = (int)(rest % a.c);
  const int b = (int)(rest / a.c);
  float f[8];
  In8<TI>::load(reinterpret_cast<const TI*>(a.x) + id * 8, f);
  const int x = x8 * 8;
  const int tp = t / a.tt, ti = t - tp * a.tt, hp = y / a.ph, hi = y - hp * a.ph, wp = x / a.pw, wi = x - wp * a.pw;
  const int nh = a.h / a.ph, nw = a.w / a.pw;
  const long long token = ((long long)tp * nh + hp) * nw + wp;
  const long long tokens = (long long)(a.t / a.tt) * nh * nw;
  const int feat = ((c * a.tt + ti) * a.ph + hi) * a.pw + wi;
  const int width = a.c * a.tt * a.ph * a.pw;
  Out8<TO>::store(reinterpret_cast<TO*>(a.out) + ((long long)b * tokens + token) * width + feat, f);
}

int launch_patchify(const void* x, int in_dtype, int b, int c, int t, int h, int w, int tt, int ph, int pw, void* out, int out_dtype,
                    cudaStream_t st) {
  if (pw % 8 != 0 || w % pw != 0 || h % ph != 0 || t % tt != 0 || ((uintptr_t)x & 31) || ((uintptr_t)out & 31))
    return set_error(TOME_ERR_UNSUPPORTED, "tome_patchify: needs pw %% 8 == 0, sizes divisible by the tubelet, 32-byte aligned buffers");
  PatchifyArgs a{x, out, b, c, t, h, w, tt, ph, pw};
  const long long total = (long long)b * c * t * h * (w >> 3);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (in_dtype == TOME_F32 && out_dtype == TOME_F32) patchify_kernel<float, float><<<grid, 256, 0, st>>>(a);
  else if (in_dtype == TOME_F32 && out_dtype == TOME_BF16) patchify_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else if (in_dtype == TOME_BF16 && out_dtype == TOME_BF16) patchify_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else if (in_dtype == TOME_BF16 && out_dtype == TOME_F32) patchify_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(a);
  else if (in_dtype == TOME_U8 && out_dtype == TOME_BF16) patchify_kernel<uint8_t, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else if (in_dtype == TOME_U8 && out_dtype == TOME_F32) patchify_kernel<uint8_t, float><<<grid, 256, 0, st>>>(a);
  else return set_error(TOME_ERR_DTYPE, "tome_patchify: unsupported dtypes %d -> %d", in_dtype, out_dtype);
  TOME_LAUNCH_CHECK("patchify_kernel");
  return TOME_OK;
}

}  // namespace tome
