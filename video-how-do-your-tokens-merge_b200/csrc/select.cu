// Kernel 2: select.  Replaces tome/merge.py:65-73 of the reference
//   edge_idx = node_max.argsort(descending=True); unm = edge[r:]; src = edge[:r];
//   dst = node_idx.gather(src); (class token) unm = unm.sort()
// with a rank-by-counting selection: rank(i) = #{j : key_j > key_i or (key_j == key_i and
// j < i)} is the position of A token i in a STABLE descending sort, every rank is
// independent (no sort network, no atomics), and the unique ranks scatter straight into
// src/unm/dst.  A second small kernel builds what needs the whole row: the ascending unm
// list for class-token models and the dst-grouped CSR the merge kernel gathers through.
#include "common.cuh"

namespace tome {

struct PlanDev {
  int bm, n, r, cls, distill;
  const float* node_max;
  const int* node_idx;
  int *src_idx, *unm_idx, *dst_idx, *a_map, *b_off, *b_src, *b_head;
};

constexpr int RANK_ROWS = 32;     // A tokens ranked per CTA (one per lane)
constexpr int RANK_WARPS = 8;     // j-range split across warps

__global__ void __launch_bounds__(RANK_ROWS * RANK_WARPS) rank_kernel(PlanDev p, int* __restrict__ rank_ws) {
  extern __shared__ uint32_t sk[];                       // [na] orderable keys
  __shared__ int partial[RANK_WARPS][RANK_ROWS];
  const int na = na_of(p.n), b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* nm = p.node_max + (long long)b * na;
  for (int j = threadIdx.x; j < na; j += blockDim.x) sk[j] = orderable_key(__ldg(nm + j));
  __syncthreads();
  const int i = blockIdx.x * RANK_ROWS + lane;
  const uint32_t mine = i < na ? sk[i] : 0u;
  const int chunk = (na + RANK_WARPS - 1) / RANK_WARPS;
  const int jb = warp * chunk, je = min(na, jb + chunk);
  int cnt = 0;
#pragma unroll 4
  for (int j = jb; j < je; ++j) {
    const uint32_t kj = sk[j];                           // warp-wide broadcast read
    cnt += (kj > mine) || (kj == mine && j < i);
  }
  partial[warp][lane] = cnt;
  __syncthreads();
  if (warp == 0 && i < na) {
    int rank = 0;
#pragma unroll
    for (int w = 0; w < RANK_WARPS; ++w) rank += partial[w][lane];
    rank_ws[(long long)b * na + i] = rank;
    const long long br = (long long)b * p.r, bu = (long long)b * (na - p.r), ba = (long long)b * na;
    if (rank < p.r) {
      const int d = __ldg(p.node_idx + ba + i);
      p.src_idx[br + rank] = i;
      p.dst_idx[br + rank] = d;
      p.a_map[ba + i] = -(d + 1);
    } else if (!p.cls) {
      p.unm_idx[bu + rank - p.r] = i;
      p.a_map[ba + i] = rank - p.r;
    }
  }
}

// exclusive scan of one int per thread across a 1024-thread block
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot /*[32]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const int res = warp_tot[warp] + inc - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(1024) finish_kernel(PlanDev p, const int* __restrict__ rank_ws) {
  extern __shared__ int sm[];
  __shared__ int warp_tot[32];
  __shared__ int total;
  const int na = na_of(p.n), nb = nb_of(p.n), r = p.r, b = blockIdx.x, tid = threadIdx.x;
  int* off = sm;                 // [nb + 1]
  int* cur = off + nb + 1;       // [nb]
  int* lst = cur + nb;           // [r]
  const long long br = (long long)b * r, bu = (long long)b * (na - r), ba = (long long)b * na;

  if (p.cls) {   // merge.py:71-73: kept tokens in ascending index order (cls, key -inf, first)
    const int per = (na + 1023) / 1024, lo = min(na, tid * per), hi = min(na, lo + per);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += rank_ws[ba + i] >= r;
    int pos = block_exclusive_scan(c, warp_tot, &total);
    for (int i = lo; i < hi; ++i)
      if (rank_ws[ba + i] >= r) { p.unm_idx[bu + pos] = i; p.a_map[ba + i] = pos; ++pos; }
  }

  // CSR of sources per dst token, ascending k inside a group (the reference CPU
  // scatter_reduce order), so the merge kernel is a pure gather.
  for (int j = tid; j < nb; j += 1024) cur[j] = 0;
  __syncthreads();
  for (int k = tid; k < r; k += 1024) atomicAdd(&cur[p.dst_idx[br + k]], 1);
  __syncthreads();
  {
    const int per = (nb + 1023) / 1024, lo = min(nb, tid * per), hi = min(nb, lo + per);
    int c = 0;
    for (int j = lo; j < hi; ++j) c += cur[j];
    int pos = block_exclusive_scan(c, warp_tot, &total);
    for (int j = lo; j < hi; ++j) { off[j] = pos; pos += cur[j]; }
    if (tid == 0) off[nb] = r;
  }
  __syncthreads();
  for (int j = tid; j < nb; j += 1024) cur[j] = off[j];
  __syncthreads();
  for (int k = tid; k < r; k += 1024) lst[atomicAdd(&cur[p.dst_idx[br + k]], 1)] = k;
  __syncthreads();
  for (int j = tid; j < nb; j += 1024) {
    const int s = off[j], e = off[j + 1];
    for (int a = s + 1; a < e; ++a) {           // insertion sort; groups are tiny
      const int v = lst[a];
      int q = a - 1;
      while (q >= s && lst[q] > v) { lst[q + 1] = lst[q]; --q; }
      lst[q + 1] = v;
    }
  }
  __syncthreads();
  int* boff = p.b_off + (long long)b * (nb + 1);
  for (int j = tid; j <= nb; j += 1024) boff[j] = off[j];
  for (int k = tid; k < r; k += 1024) p.b_src[br + k] = p.src_idx[br + lst[k]];
  // per-B-token head {count, source 0, source 1, CSR begin}: one 16-byte lookup lets the merge
  // kernel request the merged rows together with the token's own row
  int4* bhead = reinterpret_cast<int4*>(p.b_head) + (long long)b * nb;
  for (int j = tid; j < nb; j += 1024) {
    const int s0 = off[j], c = off[j + 1] - s0;
    bhead[j] = make_int4(c, c > 0 ? p.src_idx[br + lst[s0]] : 0, c > 1 ? p.src_idx[br + lst[s0 + 1]] : 0, s0);
  }
}

size_t select_workspace(int bm, int n) { return (size_t)bm * na_of(n) * sizeof(int); }

int launch_select(const tome_plan* plan, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int n = plan->n, na = na_of(n), nb = nb_of(n), r = plan->r, bm = plan->bm;
  if (ws_bytes < select_workspace(bm, n))
    return set_error(TOME_ERR_WORKSPACE, "tome_select: workspace %zu < %zu bytes", ws_bytes,
                     select_workspace(bm, n));
  PlanDev p{bm, n, r, plan->class_token, plan->distill_token, plan->node_max, plan->node_idx,
            plan->src_idx, plan->unm_idx, plan->dst_idx, plan->a_map, plan->b_off, plan->b_src, plan->b_head};
  const size_t sm_rank = (size_t)na * sizeof(uint32_t);
  const size_t sm_fin = ((size_t)2 * nb + 1 + r) * sizeof(int);
  if (sm_rank > 200 * 1024 || sm_fin > 200 * 1024)
    return set_error(TOME_ERR_UNSUPPORTED, "tome_select: n=%d needs more shared memory than one SM has", n);
  static PerDeviceOnce attr_set;
  if (attr_set.first_time()) {
    TOME_CUDA(cudaFuncSetAttribute(rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TOME_CUDA(cudaFuncSetAttribute(finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid((na + RANK_ROWS - 1) / RANK_ROWS, bm);
  rank_kernel<<<grid, RANK_ROWS * RANK_WARPS, sm_rank, st>>>(p, (int*)ws);
  TOME_LAUNCH_CHECK("rank_kernel");
  finish_kernel<<<bm, 1024, sm_fin, st>>>(p, (const int*)ws);
  TOME_LAUNCH_CHECK("finish_kernel");
  return TOME_OK;
}

}  // namespace tome
