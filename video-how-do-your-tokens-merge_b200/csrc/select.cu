// Kernel 2: select.  Replaces tome/merge.py:65-73 of the reference
//   edge_idx = node_max.argsort(descending=True); unm = edge[r:]; src = edge[:r];
//   dst = node_idx.gather(src); (class token) unm = unm.sort()
// with a rank-by-counting selection: rank(i) = #{j : key_j > key_i or (key_j == key_i and
// j < i)} is the position of A token i in a STABLE descending sort, every rank is
// independent (no sort network, no atomics), and the unique ranks scatter straight into
// src/unm/dst.  A second small kernel builds what needs the whole row: the ascending unm
// list for class-token models and the dst-grouped CSR the merge kernel gathers through.
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------
// Kernel 2 as ONE launch.
// ---------------------------------------------------------------------------------------------
// rank_kernel + finish_kernel cost two launches because the CSR needs to know, for EVERY A token of the batch element,
// whether it is a source -- i.e. the r-th largest rank key T -- while a CTA only ranks its own 32 tokens.  Here the one
// thread of the whole batch element whose rank is r - 1 publishes T through global memory (flag zeroed by the
// normalisation kernel / a memset node), the other CTAs of that batch element pick it up (acquire spin, bounded) and
// run a second counting scan over the same shared-memory keys:
//   source i        : CSR position = #{sources j : (col_j, rank order) before (col_i, mine)}
//   destination d   : begin = #{sources j : col_j < d}, count, best two sources  -> b_off / b_head
//   class token mode: ascending position of a kept token = #{kept j < i}
// Nothing depends on who publishes when; CTAs of a batch element are consecutive in dispatch order, so the publisher is
// resident (or already done) whenever a CTA of its batch element waits.  Reads either the packed keys kernel 1's atomicMax
// produced ((orderable score << 32) | ~column; then also writes node_max / node_idx) or plain node_max / node_idx.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct Top2 { unsigned long long a, b; };     // best and second-best rank key seen (0 = none)
__device__ __forceinline__ void top2_push(Top2& t, unsigned long long k) {
  if (k > t.a) { t.b = t.a; t.a = k; }
  else if (k > t.b) t.b = k;
}

__global__ void __launch_bounds__(RANK_ROWS * RANK_WARPS) select_one_kernel(PlanDev p, const unsigned long long* __restrict__ packed,
                                                                           float* __restrict__ node_max_out, int* __restrict__ node_idx_out,
                                                                           unsigned long long* __restrict__ flags) {
  extern __shared__ unsigned long long sk2[];               // [na] rank keys: (orderable score << 32) | ~row
  const int na = na_of(p.n), nb = nb_of(p.n), r = p.r, b = blockIdx.y;
  int* scol = reinterpret_cast<int*>(sk2 + na);             // [na] destination (B token) of each A token
  __shared__ int part[RANK_WARPS][RANK_ROWS][4];
  __shared__ Top2 part2[RANK_WARPS][RANK_ROWS];
  __shared__ unsigned long long s_thr;
  __shared__ int s_cnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ba = (long long)b * na;
  for (int j = threadIdx.x; j < na; j += blockDim.x) {
    uint32_t skey; int col;
    if (packed) {
      const unsigned long long k = __ldcg(packed + ba + j);
      skey = (uint32_t)(k >> 32);
      col = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
      if (j >= (int)blockIdx.x * RANK_ROWS && j < ((int)blockIdx.x + 1) * RANK_ROWS) {     // decode of this CTA's own rows
        node_max_out[ba + j] = key_to_float(skey);
        node_idx_out[ba + j] = col;
      }
    } else {
      skey = orderable_key(__ldg(p.node_max + ba + j));
      col = __ldg(p.node_idx + ba + j);
    }
    sk2[j] = ((unsigned long long)skey << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
    scol[j] = col;
  }
  __syncthreads();
  const int i = blockIdx.x * RANK_ROWS + lane;
  const unsigned long long mine = i < na ? sk2[i] : 0ull;
  const int mycol = i < na ? scol[i] : -1;
  const int chunk = (na + RANK_WARPS - 1) / RANK_WARPS;
  const int jb = warp * chunk, je = min(na, jb + chunk);
  // ---- scan 1: rank = number of keys above mine (keys are unique: score, then lower row first)
  int cnt = 0;
#pragma unroll 4
  for (int j = jb; j < je; ++j) cnt += sk2[j] > mine;
  part[warp][lane][0] = cnt;
  __syncthreads();
  int rank = 0;
#pragma unroll
  for (int w = 0; w < RANK_WARPS; ++w) rank += part[w][lane][0];
  const long long br = (long long)b * r, bu = (long long)b * (na - r);
  if (warp == 0 && i < na) {
    if (rank < r) {
      p.src_idx[br + rank] = i;
      p.dst_idx[br + rank] = mycol;
      p.a_map[ba + i] = -(mycol + 1);
    } else if (!p.cls) {
      p.unm_idx[bu + rank - r] = i;
      p.a_map[ba + i] = rank - r;
    }
    if (rank == r - 1) st_release_u64(flags + b, mine);     // the r-th largest key: everything >= it is a source
  }
  __syncwarp();
  // ---- the threshold of this batch element
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    for (unsigned spin = 0; (t = ld_acquire_u64(flags + b)) == 0ull; ++spin)
      if (spin > (1u << 24)) __trap();                     // never observed; a lost publisher must not hang the GPU
    s_thr = t;
  }
  __syncthreads();                                          // also: part[][][0] has been read by everyone
  const unsigned long long thr = s_thr;
  // ---- the r sources (keys >= threshold), compacted: the second scan walks r entries instead of na
  unsigned long long* lk = sk2 + na + (na + 1) / 2;         // [r] rank keys of the sources   (after scol, 8-byte aligned)
  int* lc = reinterpret_cast<int*>(lk + r);                // [r] their destinations
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int j0 = 0; j0 < na; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    const bool sj = j < na && sk2[j] >= thr;
    const unsigned m = __ballot_sync(0xffffffffu, sj);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sj) {
      const int at = base + __popc(m & ((1u << lane) - 1u));
      lk[at] = sk2[j];
      lc[at] = scol[j];
    }
  }
  __syncthreads();
  // ---- scan 2 over the sources: CSR position of my source, the CSR row of "my" destination d = i, and (class token)
  //      the ascending position of my kept token = i - #{sources below i}
  const int d = i;                                          // this lane also owns B token d (d < nb)
  const bool is_src = i < na && rank < r;
  int pos = 0, before = 0, ndst = 0, src_below = 0;
  Top2 t2{0ull, 0ull};
  for (int q = warp; q < r; q += RANK_WARPS) {
    const unsigned long long kj = lk[q];
    const int cj = lc[q];
    pos += (cj < mycol || (cj == mycol && kj > mine));
    before += cj < d;
    if (cj == d) { ++ndst; top2_push(t2, kj); }
    src_below += (int)(0xFFFFFFFFu - (uint32_t)(kj & 0xFFFFFFFFull)) < i;
  }
  const int kept_below = src_below;                         // combined below: K = i - sum
  part[warp][lane][0] = pos; part[warp][lane][1] = before; part[warp][lane][2] = ndst; part[warp][lane][3] = kept_below;
  part2[warp][lane] = t2;
  __syncthreads();
  if (warp == 0) {
    int P = 0, Bf = 0, N = 0, K = 0;
    Top2 m{0ull, 0ull};
#pragma unroll
    for (int w = 0; w < RANK_WARPS; ++w) {
      P += part[w][lane][0]; Bf += part[w][lane][1]; N += part[w][lane][2]; K += part[w][lane][3];
      const Top2 o = part2[w][lane];
      if (o.a) top2_push(m, o.a);
      if (o.b) top2_push(m, o.b);
    }
    if (is_src) p.b_src[br + P] = i;
    if (p.cls && i < na && !is_src) { p.unm_idx[bu + i - K] = i; p.a_map[ba + i] = i - K; }
    if (d < nb) {
      const int s0 = m.a ? (int)(0xFFFFFFFFu - (uint32_t)(m.a & 0xFFFFFFFFull)) : 0;
      const int s1 = m.b ? (int)(0xFFFFFFFFu - (uint32_t)(m.b & 0xFFFFFFFFull)) : 0;
      p.b_off[(long long)b * (nb + 1) + d] = Bf;
      reinterpret_cast<int4*>(p.b_head)[(long long)b * nb + d] = make_int4(N, s0, s1, Bf);
    }
    if (blockIdx.x == 0 && lane == 0) p.b_off[(long long)b * (nb + 1) + nb] = r;
  }
}

__global__ void zero_flags_kernel(unsigned long long* flags, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = 0ull;
}

size_t select_workspace(int bm, int n) { return (size_t)bm * na_of(n) * sizeof(int) + (size_t)bm * sizeof(unsigned long long) + 256; }

bool select_two_launches() {          // TOME_SELECT_TWO=1: the rank + finish pair (kept as a cross-check)
  const char* e = getenv("TOME_SELECT_TWO");
  return e && atoi(e) != 0;
}

// packed != nullptr: kernel 1's packed keys (plan->node_max / node_idx are then written here); flags: bm zeroed u64.
int launch_select_one(const tome_plan* plan, const unsigned long long* packed, unsigned long long* flags, cudaStream_t st) {
  const int n = plan->n, na = na_of(n), r = plan->r, bm = plan->bm;
  PlanDev p{bm, n, r, plan->class_token, plan->distill_token, plan->node_max, plan->node_idx,
            plan->src_idx, plan->unm_idx, plan->dst_idx, plan->a_map, plan->b_off, plan->b_src, plan->b_head};
  const size_t smem = ((size_t)na + (na + 1) / 2 + r + (r + 1) / 2 + 2) * sizeof(unsigned long long);     // keys, columns, source list
  if (smem > 200 * 1024) return set_error(TOME_ERR_UNSUPPORTED, "tome_select: n=%d needs more shared memory than one SM has", n);
  static PerDeviceOnce attr_set;
  if (attr_set.first_time())
    TOME_CUDA(cudaFuncSetAttribute(select_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((na + RANK_ROWS - 1) / RANK_ROWS, bm);
  select_one_kernel<<<grid, RANK_ROWS * RANK_WARPS, smem, st>>>(p, packed, const_cast<float*>(plan->node_max),
                                                                const_cast<int*>(plan->node_idx), flags);
  TOME_LAUNCH_CHECK("select_one_kernel");
  return TOME_OK;
}

int launch_select(const tome_plan* plan, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int n = plan->n, na = na_of(n), nb = nb_of(n), r = plan->r, bm = plan->bm;
  if (ws_bytes < select_workspace(bm, n))
    return set_error(TOME_ERR_WORKSPACE, "tome_select: workspace %zu < %zu bytes", ws_bytes,
                     select_workspace(bm, n));
  if (!select_two_launches()) {
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    zero_flags_kernel<<<(bm + 255) / 256, 256, 0, st>>>(flags, bm);
    TOME_LAUNCH_CHECK("zero_flags_kernel");
    return launch_select_one(plan, nullptr, flags, st);
  }
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
