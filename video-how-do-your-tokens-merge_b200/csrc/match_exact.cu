// Kernel 1, exact form: canonical scores straight from fp64 CUDA-core tiles.
//
// Replaces tome/merge.py:51-64 (normalise, a @ b^T, cls/distill masks, max(dim=-1)) of the
// reference.  Works for every shape/dtype/stride; it is also the arbiter the tcgen05 path
// (match_sm100.cu) is tested against, since both must produce identical bits.
#include "common.cuh"

namespace tome {

// ---- prep: one warp per token row: fp64 sum of squares -> fp32 norm -> fp32 normalised
// row, written in split layout (A rows first, then B rows) so the tile kernel streams it. --
template <typename T>
__global__ void __launch_bounds__(256) prep_rows_kernel(const T* __restrict__ metric, View v, int bm,
                                                        int n, int cm, float* __restrict__ mhat,
                                                        unsigned long long* __restrict__ keys) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= bm * n) return;
  const int b = warp / n, t = warp - b * n;
  const int na = na_of(n);
  const T* src = metric + v.batch_offset(b) + (long long)t * v.sn;
  double ss = 0.0;
  for (int k = lane; k < cm; k += 32) {
    double x = (double)ld_as_float(src + k);
    ss = fma(x, x, ss);
  }
  ss = warp_sum(ss);
  const float norm = (float)sqrt(ss);
  const int row = (t & 1) ? na + (t >> 1) : (t >> 1);
  float* dst = mhat + ((long long)b * n + row) * cm;
  for (int k = lane; k < cm; k += 32) dst[k] = __fdiv_rn(ld_as_float(src + k), norm);
  if (lane == 0 && !(t & 1)) keys[(long long)b * na + (t >> 1)] = 0ull;
}

// ---- 64x64 score tile per CTA, 4x4 fp64 accumulators per thread ---------------------------
constexpr int TM = 64, TN = 64, KC = 32;

// rows = na + nb normalised rows per batch element, the A rows first (the even/odd split of merge.py:52, or any
// two token sets: sets.cu)
__global__ void __launch_bounds__(256) match_exact_kernel(const float* __restrict__ mhat, int n, int na, int nb, int cm,
                                                          int cls, int distill,
                                                          unsigned long long* __restrict__ keys) {
  __shared__ float As[KC][TM + 1];
  __shared__ float Bs[KC][TN + 1];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* Ab = mhat + (long long)b * n * cm;
  const float* Bb = Ab + (long long)na * cm;

  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) acc[u][w] = 0.0;

  for (int k0 = 0; k0 < cm; k0 += KC) {
#pragma unroll
    for (int q = 0; q < (TM * KC) / 256; ++q) {
      const int idx = tid + q * 256, row = idx / KC, kk = idx % KC;
      const int k = k0 + kk;
      As[kk][row] = (i0 + row < na && k < cm) ? Ab[(long long)(i0 + row) * cm + k] : 0.f;
      Bs[kk][row] = (j0 + row < nb && k < cm) ? Bb[(long long)(j0 + row) * cm + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KC; ++kk) {
      double a[4], bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = (double)As[kk][ty + 16 * u]; bb[u] = (double)Bs[kk][tx + 16 * u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) acc[u][w] = fma(a[u], bb[w], acc[u][w]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty + 16 * u;
    unsigned long long best = 0ull;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = j0 + tx + 16 * w;
      if (j < nb) {
        float s = (float)acc[u][w];
        if ((cls && i == 0) || (distill && j == 0)) s = -INFINITY;
        const unsigned long long p = pack_best(s, j);
        best = p > best ? p : best;
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (tx == 0 && i < na) atomicMax(keys + (long long)b * na + i, best);
  }
}

__global__ void decode_keys_kernel(const unsigned long long* __restrict__ keys, int total,
                                   float* __restrict__ node_max, int* __restrict__ node_idx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned long long k = keys[t];
  node_max[t] = key_to_float((uint32_t)(k >> 32));
  node_idx[t] = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

// ---- materialised-score row max (random_* modes, merge.py:54-57) -------------------------
__global__ void __launch_bounds__(128) rowmax_kernel(const float* __restrict__ scores, int na, int nb, int cls,
                                                     int distill, float* __restrict__ node_max,
                                                     int* __restrict__ node_idx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= na) return;
  const float* row = scores + ((long long)b * na + warp) * nb;
  unsigned long long best = 0ull;
  for (int j = lane; j < nb; j += 32) {
    float s = __ldg(row + j);
    if ((cls && warp == 0) || (distill && j == 0)) s = -INFINITY;
    const unsigned long long p = pack_best(s, j);
    best = p > best ? p : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0) {
    node_max[(long long)b * na + warp] = key_to_float((uint32_t)(best >> 32));
    node_idx[(long long)b * na + warp] = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
  }
}

// ---- host ---------------------------------------------------------------------------------
int launch_match_exact_tiles(const float* mhat, int bm, int rows, int na, int nb, int cm, int cls, int distill,
                             unsigned long long* keys, float* node_max, int* node_idx, cudaStream_t st) {
  dim3 grid((nb + TN - 1) / TN, (na + TM - 1) / TM, bm);
  if (grid.z > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_match: matching batch %d > 65535", bm);
  match_exact_kernel<<<grid, 256, 0, st>>>(mhat, rows, na, nb, cm, cls, distill, keys);
  TOME_LAUNCH_CHECK("match_exact_kernel");
  const int total = bm * na;
  decode_keys_kernel<<<(total + 255) / 256, 256, 0, st>>>(keys, total, node_max, node_idx);
  TOME_LAUNCH_CHECK("decode_keys_kernel");
  return TOME_OK;
}

size_t match_exact_workspace(int bm, int n, int cm) {
  size_t mh = (size_t)bm * n * cm * sizeof(float);
  mh = (mh + 255) & ~(size_t)255;
  return mh + (size_t)bm * na_of(n) * sizeof(unsigned long long);
}

int launch_match_exact(const void* metric, int dtype, int bm, int n, int cm, const View& v, int cls,
                       int distill, float* node_max, int* node_idx, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (ws_bytes < match_exact_workspace(bm, n, cm))
    return set_error(TOME_ERR_WORKSPACE, "tome_match: workspace %zu < %zu bytes", ws_bytes,
                     match_exact_workspace(bm, n, cm));
  const int na = na_of(n), nb = nb_of(n);
  size_t mh = ((size_t)bm * n * cm * sizeof(float) + 255) & ~(size_t)255;
  float* mhat = (float*)ws;
  unsigned long long* keys = (unsigned long long*)((char*)ws + mh);
  const long long rows = (long long)bm * n;
  const int blocks = (int)((rows * 32 + 255) / 256);
  if (dtype == TOME_F32)
    prep_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)metric, v, bm, n, cm, mhat, keys);
  else
    prep_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)metric, v, bm, n, cm, mhat, keys);
  TOME_LAUNCH_CHECK("prep_rows_kernel");
  return launch_match_exact_tiles(mhat, bm, n, na, nb, cm, cls, distill, keys, node_max, node_idx, st);
}

int launch_rowmax(const float* scores, int bm, int na, int nb, int cls, int distill, float* node_max,
                  int* node_idx, cudaStream_t st) {
  dim3 grid((na * 32 + 127) / 128, bm);
  rowmax_kernel<<<grid, 128, 0, st>>>(scores, na, nb, cls, distill, node_max, node_idx);
  TOME_LAUNCH_CHECK("rowmax_kernel");
  return TOME_OK;
}

}  // namespace tome
