// Matching between two ARBITRARY token sets (upstream-ToMe variants kept by the reference:
// kth_bipartite_soft_matching, tome/merge.py:105-158, and random_bipartite_soft_matching, :161-212).
//
// The reference gathers the two sets into fresh tensors, normalises, runs a dense a @ b^T, takes
// argmax, and merges with an out-of-place scatter_reduce.  Here the sets are index lists into the
// token axis (`tok_of_row`: rows [0, ra) are the A / source tokens, rows [ra, ra + nb) the B /
// destination tokens) and three kernels do the work:
//   prep_set_rows_kernel  normalise the listed rows once (canonical arithmetic),
//   match_exact_kernel    (match_exact.cu) fp64 tiles + packed-key atomicMax: argmax per A row,
//   group_reduce_kernel   one warp per B row: itself, then every A row assigned to it in ascending
//                         k -- the reference CPU scatter_reduce order, include_self = True,
//   gather_rows_kernel    unmerge: out[t] = x[map[t]].
#include "common.cuh"

namespace tome {

int launch_match_exact_tiles(const float* mhat, int bm, int rows, int na, int nb, int cm, int cls, int distill,
                             unsigned long long* keys, float* node_max, int* node_idx, cudaStream_t st);

template <typename T>
__global__ void __launch_bounds__(256) prep_set_rows_kernel(const T* __restrict__ metric, View v, int bm, int rows, int ra,
                                                            int cm, const int* __restrict__ tok_of_row, long long tok_stride_b,
                                                            float* __restrict__ mhat, unsigned long long* __restrict__ keys) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= bm * rows) return;
  const int b = warp / rows, row = warp - b * rows;
  const int t = __ldg(tok_of_row + b * tok_stride_b + row);
  const T* src = metric + v.batch_offset(b) + (long long)t * v.sn;
  double ss = 0.0;
  for (int k = lane; k < cm; k += 32) {
    const double x = (double)ld_as_float(src + k);
    ss = fma(x, x, ss);
  }
  ss = warp_sum(ss);
  const float norm = (float)sqrt(ss);
  float* dst = mhat + ((long long)b * rows + row) * cm;
  for (int k = lane; k < cm; k += 32) dst[k] = __fdiv_rn(ld_as_float(src + k), norm);
  if (lane == 0 && row < ra) keys[(long long)b * ra + row] = 0ull;
}

__device__ __forceinline__ float nanmax2(float a, float b) { return (a != a || a > b) ? a : ((b != b) ? b : (a > b ? a : b)); }

template <typename T>
__global__ void __launch_bounds__(128) group_reduce_kernel(const T* __restrict__ x, View xv, int ra, int nb, int c,
                                                           const int* __restrict__ tok_of_row, long long tok_stride_b,
                                                           const int* __restrict__ dst_idx, int mode, T* __restrict__ out) {
  extern __shared__ int s_dst[];                      // [ra] destinations of this batch element's A rows
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < ra; k += blockDim.x) s_dst[k] = __ldg(dst_idx + (long long)b * ra + k);
  __syncthreads();
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= nb) return;
  const int* tok = tok_of_row + b * tok_stride_b;
  const T* xb = x + xv.batch_offset(b);
  T* orow = out + ((long long)b * nb + j) * c;
  const T* self = xb + (long long)__ldg(tok + ra + j) * xv.sn;
  for (int c0 = lane; c0 < c; c0 += 32) {
    float acc = ld_as_float(self + c0);
    int cnt = 1;
    for (int k = 0; k < ra; ++k) {
      if (s_dst[k] != j) continue;                    // warp-uniform
      const float v = ld_as_float(xb + (long long)__ldg(tok + k) * xv.sn + c0);
      acc = mode == TOME_MODE_AMAX ? nanmax2(acc, v) : __fadd_rn(acc, v);
      ++cnt;
    }
    if (mode == TOME_MODE_MEAN) acc = __fdiv_rn(acc, (float)cnt);
    if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16*>(orow)[c0] = __float2bfloat16_rn(acc);
    else reinterpret_cast<float*>(orow)[c0] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ x, int n_in, int n_out, int c,
                                                          const int* __restrict__ map, T* __restrict__ out) {
  const int lane = threadIdx.x & 31, b = blockIdx.y;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_out) return;
  const int s = __ldg(map + (long long)b * n_out + t);
  T* dst = out + ((long long)b * n_out + t) * c;
  if (s < 0) { for (int k = lane; k < c; k += 32) dst[k] = T(0.f); return; }
  const T* src = x + ((long long)b * n_in + s) * c;
  for (int k = lane; k < c; k += 32) dst[k] = src[k];
}

size_t match_sets_workspace(int bm, int rows, int ra, int cm) {
  const size_t mh = ((size_t)bm * rows * cm * sizeof(float) + 255) & ~(size_t)255;
  return mh + (size_t)bm * ra * sizeof(unsigned long long);
}

int launch_match_sets(const void* metric, int dtype, int bm, int cm, const View& v, const int* tok_of_row, long long tok_stride_b,
                      int ra, int nb, float* node_max, int* node_idx, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rows = ra + nb;
  if (ws_bytes < match_sets_workspace(bm, rows, ra, cm))
    return set_error(TOME_ERR_WORKSPACE, "tome_match_sets: workspace %zu < %zu bytes", ws_bytes, match_sets_workspace(bm, rows, ra, cm));
  const size_t mh = ((size_t)bm * rows * cm * sizeof(float) + 255) & ~(size_t)255;
  float* mhat = (float*)ws;
  unsigned long long* keys = (unsigned long long*)((char*)ws + mh);
  const int blocks = (int)(((long long)bm * rows * 32 + 255) / 256);
  if (dtype == TOME_F32)
    prep_set_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)metric, v, bm, rows, ra, cm, tok_of_row, tok_stride_b, mhat, keys);
  else
    prep_set_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)metric, v, bm, rows, ra, cm, tok_of_row, tok_stride_b, mhat, keys);
  TOME_LAUNCH_CHECK("prep_set_rows_kernel");
  return launch_match_exact_tiles(mhat, bm, rows, ra, nb, cm, 0, 0, keys, node_max, node_idx, st);
}

int launch_group_reduce(const void* x, int dtype, int bm, int c, const View& xv, const int* tok_of_row, long long tok_stride_b,
                        int ra, int nb, const int* dst_idx, int mode, void* out, cudaStream_t st) {
  const size_t smem = (size_t)ra * sizeof(int);
  if (smem > 48 * 1024) return set_error(TOME_ERR_UNSUPPORTED, "tome_group_reduce: %d source rows exceed the staging buffer", ra);
  dim3 grid((nb + 3) / 4, bm);
  if (grid.y > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_group_reduce: batch %d > 65535", bm);
  if (dtype == TOME_F32)
    group_reduce_kernel<float><<<grid, 128, smem, st>>>((const float*)x, xv, ra, nb, c, tok_of_row, tok_stride_b, dst_idx, mode, (float*)out);
  else
    group_reduce_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>((const __nv_bfloat16*)x, xv, ra, nb, c, tok_of_row, tok_stride_b, dst_idx, mode, (__nv_bfloat16*)out);
  TOME_LAUNCH_CHECK("group_reduce_kernel");
  return TOME_OK;
}

int launch_gather_rows(const void* x, int dtype, int bm, int n_in, int c, const int* map, int n_out, void* out, cudaStream_t st) {
  dim3 grid((n_out + 7) / 8, bm);
  if (grid.y > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_gather_rows: batch %d > 65535", bm);
  if (dtype == TOME_F32)
    gather_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n_in, n_out, c, map, (float*)out);
  else
    gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, n_in, n_out, c, map, (__nv_bfloat16*)out);
  TOME_LAUNCH_CHECK("gather_rows_kernel");
  return TOME_OK;
}

}  // namespace tome
