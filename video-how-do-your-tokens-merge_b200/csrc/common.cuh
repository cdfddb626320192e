// Shared device/host helpers for the tome_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tome_b200.h"

namespace tome {

// ---- error plumbing (host) --------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
#define TOME_CHECK_ARG(cond, ...)                         \
  do {                                                    \
    if (!(cond)) return tome::set_error(TOME_ERR_ARG, __VA_ARGS__); \
  } while (0)
#define TOME_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return tome::set_error(TOME_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
void count_launch();
#define TOME_LAUNCH_CHECK(name)                                                           \
  do {                                                                                    \
    tome::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return tome::set_error(TOME_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

int ensure_device_ok();   // TOME_OK if the current device is sm_100
int sm_count();

// Function attributes (the opt-in to > 48 KB of dynamic shared memory) are per DEVICE: a process that launches
// on cuda:0 and later on cuda:1 must set them on both.  One flag per device; setting an attribute twice is
// harmless, so a race between two host threads is benign.
struct PerDeviceOnce {
  unsigned char done[64] = {};
  bool first_time() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = 1;
    return true;
  }
};

// ---- geometry -----------------------------------------------------------------------------
__host__ __device__ inline int na_of(int n) { return (n + 1) >> 1; }
__host__ __device__ inline int nb_of(int n) { return n >> 1; }

struct View {            // device copy of tome_view (see include/tome_b200.h)
  long long sbo, sbi, sn;
  int inner;
  __host__ __device__ inline long long batch_offset(int b) const {
    return inner == 1 ? (long long)b * sbo : (long long)(b / inner) * sbo + (long long)(b % inner) * sbi;
  }
};
inline View make_view(const tome_view* v, long long n_tokens, long long c) {
  View o;
  if (v) { o.sbo = v->stride_bo; o.sbi = v->stride_bi; o.sn = v->stride_n; o.inner = v->inner > 0 ? v->inner : 1; }
  else   { o.sbo = n_tokens * c; o.sbi = 0; o.sn = c; o.inner = 1; }
  return o;
}

// ---- fp32 <-> totally ordered u32 (NaN above +inf, -0 == +0) ------------------------------
// Mirrors oracle/tome_oracle.py::orderable_u32.
__device__ __forceinline__ uint32_t orderable_key(float v) {
  v = v + 0.0f;                                    // -0 -> +0
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (v != v) ? 0xFFFFFFFFu : u;
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}
// (score key, lowest column wins) packed so one u64 max does max + first-argmax.
__device__ __forceinline__ unsigned long long pack_best(float s, int j) {
  return ((unsigned long long)orderable_key(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
}

// ---- loads ----------------------------------------------------------------------------------
__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Packed fp32 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE fp32 operations per issue slot, each lane rounded
// exactly like its scalar form, so every bit-exactness guarantee below is unchanged.  The fused bf16 launch executed
// 755 warp instructions per row and kept the issue slots 51 % busy with 31 % of the warps resident
// (profiles/r02_merge_ncu.txt): it was short of issue slots, not of HBM bandwidth.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

}  // namespace tome
