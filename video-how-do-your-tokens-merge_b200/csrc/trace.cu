// the `source` consumers and the random modes.
//
// 1. Compact source map.  The reference tracks which original tokens ended up in which merged token as a
//    dense fp32 0/1 matrix (bm, n', n0) that every block re-reduces with scatter_reduce('amax')
//    (tome/merge.py:372-384; consumed by tome/vis.py:55,102,146 as `source.argmax(dim=1)`).  Every original
//    token belongs to AT MOST one merged token (the merges are disjoint unions of an identity start), so the
//    whole matrix is one int32 per original token: group[b, t] = output slot holding token t, -1 once the
//    token has been dropped (drop modes, merge.py:260-269; destinations a hybrid threshold zeroed,
//    merge.py:326).  A block's update is a composition with the plan's slot map -- bm * n0 * 8 bytes instead of
//    bm * (2n - r) * n0 * 4 -- and the dense matrix is expanded only when somebody asks for it.
// 2. Random scores (merge.py:54-57, 235-238: `torch.rand` of shape (bm, na, nb), then max).  Here a counter-based
//    Philox4x32-10 stream indexed by (clip, A row, B column): the score of an edge does not depend on the batch
//    composition or on how clips are sharded over GPUs, the (bm, na, nb) tensor is never written, and the
//    row max / first argmax come out of the same pass.
#include "common.cuh"

namespace tome {

// ---- 1. compact source ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) source_compose_kernel(int bm, int n, int r, int distill, int drop, int hybrid, float thr,
                                                             const float* __restrict__ node_max, const int* __restrict__ a_map,
                                                             const int* __restrict__ b_off, const int* __restrict__ b_src,
                                                             const int* __restrict__ group_in, int n0, int* __restrict__ group_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (t >= n0) return;
  const int na = na_of(n), nb = nb_of(n), nu = na - r;
  const int g = group_in ? __ldg(group_in + (long long)b * n0 + t) : t;     // token index before this merge
  int slot = -1;
  if (g >= 0 && g < n) {
    int j = -1;                       // B token the original token now lives in (if any)
    if (g & 1) j = g >> 1;
    else {
      const int m = __ldg(a_map + (long long)b * na + (g >> 1));
      if (m >= 0) slot = distill ? (m == 0 ? 0 : m + 1) : m;                 // kept A token: position in unm_idx
      else if (!drop) j = -m - 1;                                            // merged into B token -m-1 (dropped in drop modes)
    }
    if (j >= 0) {
      bool keep = true;
      if (hybrid && (g & 1)) {        // the destination's OWN tokens vanish when an under-threshold edge hits it
        const int* off = b_off + (long long)b * (nb + 1) + j;
        for (int q = __ldg(off); q < __ldg(off + 1); ++q)
          keep &= (__ldg(node_max + (long long)b * na + __ldg(b_src + (long long)b * r + q)) >= thr);
      }
      if (keep) slot = distill ? (j == 0 ? 1 : nu + j) : nu + j;
    }
  }
  group_out[(long long)b * n0 + t] = slot;
}

// dense (bm, n_tokens, n0) fp32 view of a group map: out[b, s, t] = (group[b, t] == s)
__global__ void __launch_bounds__(256) source_dense_kernel(const int* __restrict__ group, int n_tokens, int n0, float* __restrict__ out) {
  const int b = blockIdx.z, s = blockIdx.y;
  const int* g = group + (long long)b * n0;
  float* row = out + ((long long)b * n_tokens + s) * n0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n0; t += gridDim.x * blockDim.x) row[t] = (__ldg(g + t) == s) ? 1.0f : 0.0f;
}

int launch_source_compose(const tome_plan* p, const int* group_in, int n0, int drop, float thr, int* group_out, cudaStream_t st) {
  dim3 grid((n0 + 255) / 256, p->bm);
  source_compose_kernel<<<grid, 256, 0, st>>>(p->bm, p->n, p->r, p->distill_token, drop, (thr == thr) ? 1 : 0, thr, p->node_max,
                                               p->a_map, p->b_off, p->b_src, group_in, n0, group_out);
  TOME_LAUNCH_CHECK("source_compose_kernel");
  return TOME_OK;
}

int launch_source_dense(const int* group, int bm, int n_tokens, int n0, float* out, cudaStream_t st) {
  if (n_tokens > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_source_dense: %d tokens > 65535", n_tokens);
  int gx = (n0 + 255) / 256;
  gx = gx > 8 ? 8 : gx;
  dim3 grid(gx, n_tokens, bm);
  source_dense_kernel<<<grid, 256, 0, st>>>(group, n_tokens, n0, out);
  TOME_LAUNCH_CHECK("source_dense_kernel");
  return TOME_OK;
}

// ---- 2. Philox4x32-10 random scores ---------------------------------------------------------------------
// Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); constants of Random123's philox4x32.
// Stream layout (mirrored by oracle/tome_oracle.py::philox_scores):
//   key     = (seed lo, seed hi)
//   counter = (column quad j / 4, A row i, clip, call)      -> four 32-bit outputs = columns 4q .. 4q+3
//   score   = (output >> 8) * 2^-24   in [0, 1)   (24 bits: every value is an exact fp32, as torch.rand's are)
struct PhiloxState { unsigned long long seed, call; };      // device-resident, so a captured graph replays a fresh draw

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float philox_unit(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// One warp per (clip, A row): lanes stride the column quads; max + first argmax through the packed key.
__global__ void __launch_bounds__(256) random_rowmax_kernel(const PhiloxState* __restrict__ state, long long clip0, int na, int nb,
                                                            int cls, int distill, float* __restrict__ node_max,
                                                            int* __restrict__ node_idx, float* __restrict__ scores_out) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), b = blockIdx.y;
  if (i >= na) return;
  const unsigned long long seed = state->seed, call = state->call, clip = (unsigned long long)(clip0 + b);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  unsigned long long best = 0ull;
  for (int q = lane; 4 * q < nb; q += 32) {
    const uint4 o = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)i, (uint32_t)clip, (uint32_t)call), key);
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * q + e;
      if (j < nb) {
        float s = philox_unit(w[e]);
        if (scores_out) scores_out[((long long)b * na + i) * nb + j] = s;      // the raw draw (masks are applied to the max only)
        if ((cls && i == 0) || (distill && j == 0)) s = -INFINITY;             // merge.py:59-62
        const unsigned long long k = pack_best(s, j);
        best = k > best ? k : best;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0) {
    node_max[(long long)b * na + i] = key_to_float((uint32_t)(best >> 32));
    node_idx[(long long)b * na + i] = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
  }
}

__global__ void philox_advance_kernel(PhiloxState* state) { state->call += 1ull; }

int launch_random_rowmax(void* state, long long clip0, int bm, int na, int nb, int cls, int distill, float* node_max, int* node_idx,
                         float* scores_out, int advance, cudaStream_t st) {
  dim3 grid((na + 7) / 8, bm);
  random_rowmax_kernel<<<grid, 256, 0, st>>>((const PhiloxState*)state, clip0, na, nb, cls, distill, node_max, node_idx, scores_out);
  TOME_LAUNCH_CHECK("random_rowmax_kernel");
  if (advance) {
    philox_advance_kernel<<<1, 1, 0, st>>>((PhiloxState*)state);
    TOME_LAUNCH_CHECK("philox_advance_kernel");
  }
  return TOME_OK;
}

}  // namespace tome
