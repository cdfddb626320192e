// Kernels 1 + 2 as ONE launch: a thread-block CLUSTER per batch element (tome/merge.py:49-73 -- everything
// bipartite_soft_matching computes: normalise, a @ b^T, cls / distill masks, max(dim=-1), argsort, src / unm / dst,
// the class-token re-sort -- plus the dst-grouped CSR kernel 3 gathers through).
//
// The four-launch chain it replaces (split_rows -> match_tc -> rank -> finish: match_sm100.cu, select.cu) was a chain of
// latencies, 29 us at the bench shape against a 3 us bound (profiles/r01_match_notes.md): 6.4 MB of normalised rows
// written out and read back through TMA by 392 single-tile CTAs, cross-tile atomics, two more launches to sort 784
// keys.  Here the CTAs of a cluster split the A rows AND the B rows of one batch element between them:
//
//   prep     every CTA takes the head-mean of K for its tokens (tome/patch/videomae.py:72-73), normalises them
//            (canonical arithmetic) and writes the bf16 two-term split h + m of its A rows STRAIGHT
//            into the SWIZZLE_128B operand tile in its own shared memory (fp32 rows beside it for the exact refine);
//            only the B rows (a quarter of the old workspace traffic) go through L2, as plain rows TMA can fetch.
//   sweep    cluster barrier, then each CTA walks the cluster's B tiles: TMA -> tcgen05.mma kind::f16 (h.h + h.m + m.h,
//            fp32 in TMEM, two accumulators so the MMAs of tile t+1 run under the epilogue of tile t) -> ONE TMEM read
//            per tile into registers: row max and the bit mask of columns within the error window of it.  The running
//            row max lives in registers: no atomics, no strip counters.
//   refine   the few surviving candidates (about one per row) are scored exactly (fp64 FMA over the fp32 rows) by the
//            warp, four lanes per row; every reported bit comes from here, so the result equals match_exact.cu's.
//   rank     the CTAs exchange their rows' keys through distributed shared memory (st.shared::cluster), cluster
//            barrier, rank-by-counting of the own rows (stable descending order), src / unm / dst written directly.
//   csr      the ranked (src, dst) edges are exchanged the same way; each CTA builds b_off / b_src / b_head for its
//            own B tokens and the ascending kept list for class-token models.
//
// One launch, four cluster barriers, no global atomics.  Falls back to the multi-launch path for metrics wider than 64
// channels (head-concat), unaligned views, or more than 2048 A tokens.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace tome {

constexpr int PC_THREADS = 320;          // warp 0: TMA producer, warp 1: TMEM + MMA issuer, warps 2..9: workers
constexpr int PC_WORKERS = 256;
constexpr int PC_MAX_STAGES = 8;         // B tiles in flight (TMA latency ~1.2 us against ~0.1 us of work per tile)
constexpr int PC_MAXCS = 16;
constexpr int PC_TMEM_COLS = 256;        // two accumulators of up to 128 columns
constexpr int PC_K = 64;                 // channels per row as stored (one 128-byte swizzle row of bf16)

struct PcParams {
  int bm, n, na, nb, cm, cls, distill, heads, r;
  long long stride_h;
  View v;
  int CS, RA, RB, BN, stages;
  int rows_total;                        // bm * nb: row offset of the "m" plane in the (2 * bm * nb, 64) bf16 tensor
  float window;
  __nv_bfloat16* hmB;                    // workspace: B rows, h plane then m plane, plain row-major
  float* mhatB;                          // workspace: B rows, fp32
  float* node_max;
  int* node_idx;
  int *src_idx, *unm_idx, *dst_idx, *a_map, *b_off, *b_src, *b_head;
  long long* trace;                      // optional (TOME_PC_TRACE): per-CTA %globaltimer stamps, 16 per CTA
};

// per-tile stamps of CTA (0, 0) behind the per-CTA records: [ncta * 16 + tile * 8 + k]
#define PC_TILE(tid, it, k) do { if (p.trace && (int)threadIdx.x == (tid) && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 16) \
    p.trace[(long long)gridDim.x * gridDim.y * 16 + (it) * 8 + (k)] = gtime(); } while (0)
#define PC_TRACE(tid, slot) do { if (p.trace && (int)threadIdx.x == (tid)) p.trace[((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime(); } while (0)

// shared-memory layout (bytes from a 1024-aligned base); the same function sizes the launch on the host
struct PcSmem {
  uint32_t a_h, a_m, stage0, stage_bytes, a_f32, rec_tmax, rec_m0, rec_m1, rowmax, rowkey, cnt2, nidx, rankbuf, tile_pt,
      allkeys, edge_src, edge_dst, bars, total;
};
__host__ __device__ inline PcSmem pc_smem(int CS, int BN, int stages) {
  PcSmem s;
  uint32_t o = 0;
  s.a_h = o; o += 128 * 128;
  s.a_m = o; o += 128 * 128;
  s.stage_bytes = 2u * BN * 128u;
  s.stage0 = o; o += stages * s.stage_bytes;
  s.a_f32 = o; o += 128 * PC_K * 4;
  s.rec_tmax = o; o += CS * PC_WORKERS * 4;
  s.rec_m0 = o; o += CS * PC_WORKERS * 4;
  s.rec_m1 = o; o += CS * PC_WORKERS * 4;
  s.rowmax = o; o += 2 * 128 * 4;
  s.rowkey = o; o += 2 * 128 * 8;
  s.cnt2 = o; o += 128 * 4;
  s.nidx = o; o += 128 * 4;
  s.rankbuf = o; o += 128 * 4;
  s.tile_pt = o; o += PC_MAXCS * 4;
  s.allkeys = o; o += CS * 128 * 4 + 16;
  s.edge_src = o; o += CS * 128 * 4 + 16;
  s.edge_dst = o; o += CS * 128 * 4 + 16;
  s.bars = o; o += 256;                  // full[8] | empty[8] | tfull[2] | tempty[2] | tmem slot
  s.total = o;
  return s;
}

// ---- PTX: clusters and distributed shared memory ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {       // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// rows other CTAs of the cluster wrote to global memory during this kernel: a coherent load, not the read-only path
__device__ __forceinline__ float4 ld_global_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}

// ---- prep: eight channels of one token per lane ---------------------------------------------------------------------
template <typename T> struct Ld8;
template <> struct Ld8<float> {
  static constexpr int BATCH = 6;
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load(const float* p) {
    Raw r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[8]) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
  }
  static __device__ __forceinline__ float round_like_input(float x) { return x; }
};
template <> struct Ld8<__nv_bfloat16> {
  static constexpr int BATCH = 12;
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void unpack(const Raw& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
  static __device__ __forceinline__ float round_like_input(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

// channels [k0, k0 + 8) of TWO tokens at once (every load of both in flight before the first add): the value itself, or
// the mean over heads rounded to the input dtype like the reference's k.mean(1) tensor (heads added in order, one
// multiply by 1 / H: ATen's MeanOps)
template <typename T>
__device__ __forceinline__ void token_chunk2(const T* s0, const T* s1, bool on0, bool on1, int heads, long long stride_h,
                                             float (&x0)[8], float (&x1)[8]) {
  typedef typename Ld8<T>::Raw Raw;
#pragma unroll
  for (int e = 0; e < 8; ++e) { x0[e] = 0.f; x1[e] = 0.f; }
  if (heads == 1) {
    Raw r0, r1;
    if (on0) r0 = Ld8<T>::load(s0);
    if (on1) r1 = Ld8<T>::load(s1);
    if (on0) Ld8<T>::unpack(r0, x0);
    if (on1) Ld8<T>::unpack(r1, x1);
    return;
  }
  constexpr int NB = Ld8<T>::BATCH;
  int h = 0;
  for (; h + NB <= heads; h += NB) {
    Raw a[NB], c[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (on0) a[u] = Ld8<T>::load(s0 + (long long)(h + u) * stride_h);
      if (on1) c[u] = Ld8<T>::load(s1 + (long long)(h + u) * stride_h);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {               // adds stay sequential in h
      float f[8];
      if (on0) { Ld8<T>::unpack(a[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) x0[e] += f[e]; }
      if (on1) { Ld8<T>::unpack(c[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) x1[e] += f[e]; }
    }
  }
  for (; h < heads; ++h) {
    float f[8];
    if (on0) { Ld8<T>::unpack(Ld8<T>::load(s0 + (long long)h * stride_h), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) x0[e] += f[e]; }
    if (on1) { Ld8<T>::unpack(Ld8<T>::load(s1 + (long long)h * stride_h), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) x1[e] += f[e]; }
  }
  const float inv = 1.0f / (float)heads;
#pragma unroll
  for (int e = 0; e < 8; ++e) { x0[e] = Ld8<T>::round_like_input(x0[e] * inv); x1[e] = Ld8<T>::round_like_input(x1[e] * inv); }
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- exact canonical scores, warp-cooperative ---------------------------------------------------------------------------
// Lane L wants <A row (a_blk + L * 64), B row (brows + col * 64)> (`act`: it has a request).  Eight requests per pass,
// four lanes per request, every load of a pass issued before the first use (see match_sm100.cu::warp_exact32, whose
// scheme this is); the A rows sit in this CTA's shared memory, the B rows in L2.  Products are exact in fp64; the order of
// the fp64 additions is not part of the score definition.
__device__ __forceinline__ float warp_exact64(const float* a_blk, const float* brows, bool act, int col, int lane) {
  const unsigned full = 0xffffffffu;
  const unsigned actmask = __ballot_sync(full, act);
  float mine = 0.f;
  const int items = __popc(actmask);
  if (items == 0) return mine;
  const int sub = lane & 3, grp = lane >> 2;
  const int my_item = __popc(actmask & ((1u << lane) - 1u));
  // every B row of up to four passes (32 requests) requested before the first use: ONE L2 round trip for the whole warp
  float4 y[4][4];
  int rr[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int item = 8 * g + grp;
    const bool on = item < items;
    const int r = on ? (int)__fns(actmask, 0, item + 1) : 0;
    const int c = __shfl_sync(full, col, r);
    rr[g] = on ? r : -1;
    const float4* bp = reinterpret_cast<const float4*>(brows + (long long)c * PC_K);
#pragma unroll
    for (int t = 0; t < 4; ++t) y[g][t] = on ? ld_global_f4(bp + 4 * t + sub) : make_float4(0.f, 0.f, 0.f, 0.f);   // chunk 4 t + sub
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (8 * g < items) {                         // warp-uniform
      const float4* ap = reinterpret_cast<const float4*>(a_blk + max(rr[g], 0) * PC_K);
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 x = rr[g] >= 0 ? ap[4 * t + sub] : make_float4(0.f, 0.f, 0.f, 0.f);
        acc0 = fma((double)x.x, (double)y[g][t].x, acc0);
        acc1 = fma((double)x.y, (double)y[g][t].y, acc1);
        acc2 = fma((double)x.z, (double)y[g][t].z, acc2);
        acc3 = fma((double)x.w, (double)y[g][t].w, acc3);
      }
      double sum = (acc0 + acc1) + (acc2 + acc3);
      sum += __shfl_xor_sync(full, sum, 1);
      sum += __shfl_xor_sync(full, sum, 2);
      const double res = __shfl_sync(full, sum, (my_item & 7) << 2);
      if (act && (my_item >> 3) == g) mine = (float)res;
    }
  }
  return mine;
}

// ---- one tile of the sweep: TMEM -> registers once, row max + window mask ---------------------------------------------
__device__ __forceinline__ void or_if_ge(uint32_t& acc, float v, float thr, uint32_t bit) {      // two instructions per column
  asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(acc) : "f"(v), "f"(thr), "r"(bit));
}

template <int NCH>       // 8-column chunks per half tile
__device__ __forceinline__ void sweep_tile(uint32_t taddr, int ncol_half, bool mask_col0, float window, float& mrun,
                                           float& tmax_out, uint32_t& m0_out, uint32_t& m1_out, uint32_t bar_tempty, int lane) {
  float v[NCH * 8];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) tmem_ld8_nowait(taddr + (uint32_t)(ch * 8), v + ch * 8);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_tempty);          // the accumulator is in registers: the MMA warp may overwrite it
  if (ncol_half < NCH * 8 || mask_col0) {          // warp-uniform: only a boundary tile pays for the validity selects
#pragma unroll
    for (int e = 0; e < NCH * 8; ++e) v[e] = (e < ncol_half && !(mask_col0 && e == 0)) ? v[e] : -INFINITY;
  }
  float part[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int e = 0; e < NCH * 8; ++e) part[e & 3] = fmax_nan(part[e & 3], v[e]);
  const float mt = fmax_nan(fmax_nan(part[0], part[1]), fmax_nan(part[2], part[3]));
  const float thr = fmax_nan(mt, mrun) - window;   // NaN anywhere -> NaN threshold -> empty masks, the tile is flagged by mt
  uint32_t lo[2] = {0u, 0u}, hi[2] = {0u, 0u};
#pragma unroll
  for (int e = 0; e < NCH * 8; ++e) {
    if (e < 32) or_if_ge(lo[e & 1], v[e], thr, 1u << e); else or_if_ge(hi[e & 1], v[e], thr, 1u << (e - 32));
  }
  // columns past the end of the tile hold -inf: they pass a -inf threshold (a tile with no live column) -- cut them
  const uint32_t vlo = ncol_half >= 32 ? 0xffffffffu : ((1u << ncol_half) - 1u);
  const uint32_t vhi = ncol_half >= 64 ? 0xffffffffu : (ncol_half > 32 ? ((1u << (ncol_half - 32)) - 1u) : 0u);
  tmax_out = mt;
  m0_out = (lo[0] | lo[1]) & vlo;
  m1_out = (hi[0] | hi[1]) & vhi;
  mrun = fmax_nan(mrun, mt);
}

template <typename T>
__global__ void __launch_bounds__(PC_THREADS, 1)
plan_cluster_kernel(const __grid_constant__ CUtensorMap map_b, const T* __restrict__ metric, const PcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int c = (int)cluster_ctarank(), b = blockIdx.y;
  const int CS = p.CS, RA = p.RA, RB = p.RB, BN = p.BN, NST = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const PcSmem L = pc_smem(CS, BN, NST);
  const uint32_t bar_full = base + L.bars, bar_empty = bar_full + 64u, bar_tfull = bar_full + 128u, bar_tempty = bar_full + 144u;
  const uint32_t tmem_slot = bar_full + 160u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + L.bars + 160u);
  float* a_f32 = reinterpret_cast<float*>(gen + L.a_f32);
  float* rec_tmax = reinterpret_cast<float*>(gen + L.rec_tmax);
  uint32_t* rec_m0 = reinterpret_cast<uint32_t*>(gen + L.rec_m0);
  uint32_t* rec_m1 = reinterpret_cast<uint32_t*>(gen + L.rec_m1);
  float* rowmax = reinterpret_cast<float*>(gen + L.rowmax);
  unsigned long long* rowkey = reinterpret_cast<unsigned long long*>(gen + L.rowkey);
  int* cnt2 = reinterpret_cast<int*>(gen + L.cnt2);
  int* nidx = reinterpret_cast<int*>(gen + L.nidx);
  int* rankbuf = reinterpret_cast<int*>(gen + L.rankbuf);
  int* tile_pt = reinterpret_cast<int*>(gen + L.tile_pt);
  uint32_t* allkeys = reinterpret_cast<uint32_t*>(gen + L.allkeys);
  int* edge_src = reinterpret_cast<int*>(gen + L.edge_src);
  int* edge_dst = reinterpret_cast<int*>(gen + L.edge_dst);

  const int na = p.na, nb = p.nb, r = p.r;
  const int RAv = max(0, min(RA, na - c * RA));          // A rows (even tokens) this CTA owns: i = c * RA + row
  const int RBv = max(0, min(RB, nb - c * RB));          // B rows (odd tokens):              j = c * RB + row
  const int live_quarters = (RAv + 31) >> 5;             // TMEM lane quarters that hold rows of this CTA
  const bool sweeping = RAv > 0;

  // ---- setup ---------------------------------------------------------------------------------------------------------
  PC_TRACE(0, 0);
  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_b);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    // an accumulator is free again once every worker warp that owns rows has read it out (two warps per lane quarter)
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8u * a, 1); mbar_init(bar_tempty + 8u * a, 2 * max(live_quarters, 1)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)PC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // padding of the exchanged tables (vector reads run past the live entries): keys 0 rank below everything, destinations
  // INT_MAX match no token
  for (int k = threadIdx.x; k < CS * 128 + 4; k += PC_THREADS) {
    if (k >= na) allkeys[k] = 0u;
    if (k >= r) { edge_dst[k] = 0x7fffffff; edge_src[k] = 0x7fffffff; }
  }

  PC_TRACE(0, 12);
  // ---- prep: normalise this CTA's tokens ---------------------------------------------------------------------------
  // eight lanes per token (eight channels each), two tokens per lane group in flight: a warp has 8 tokens x 12 heads of
  // 16-byte loads outstanding before the first add
  {
    const int g4 = lane >> 3, l8 = lane & 7, k0 = 8 * l8;
    const T* mb = metric + p.v.batch_offset(b);
    const int ntok = RAv + RBv;
    for (int w0 = warp * 8; w0 < ntok; w0 += (PC_THREADS / 32) * 8) {
      int row[2], tok[2];
      bool on[2], isA[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int w = w0 + 4 * u + g4;
        on[u] = w < ntok;
        isA[u] = w < RAv;
        row[u] = isA[u] ? w : w - RAv;
        tok[u] = isA[u] ? 2 * (c * RA + row[u]) : 2 * (c * RB + row[u]) + 1;
      }
      float x[2][8];
      const bool ld = k0 < p.cm;
      token_chunk2<T>(mb + (long long)tok[0] * p.v.sn + k0, mb + (long long)tok[1] * p.v.sn + k0, on[0] && ld, on[1] && ld,
                      p.heads, p.stride_h, x[0], x[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        double ss = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fma((double)x[u][e], (double)x[u][e], ss);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        const float norm = (float)sqrt(ss);
        float mh[8], hf[8], mf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          mh[e] = (k0 + e < p.cm) ? __fdiv_rn(x[u][e], norm) : 0.f;    // no eps: a zero row gives NaN (merge.py:51)
          hf[e] = __bfloat162float(__float2bfloat16_rn(mh[e]));
          mf[e] = mh[e] - hf[e];                                        // exact
        }
        const uint4 hq = pack8_bf16(hf), mq = pack8_bf16(mf);
        if (on[u]) {
          if (isA[u]) {
            const uint32_t off = (uint32_t)row[u] * 128u + ((uint32_t)(l8 ^ (row[u] & 7)) << 4);   // SWIZZLE_128B, K-major
            *reinterpret_cast<uint4*>(gen + L.a_h + off) = hq;
            *reinterpret_cast<uint4*>(gen + L.a_m + off) = mq;
            float4* fr = reinterpret_cast<float4*>(a_f32 + row[u] * PC_K + k0);
            fr[0] = make_float4(mh[0], mh[1], mh[2], mh[3]);
            fr[1] = make_float4(mh[4], mh[5], mh[6], mh[7]);
          } else {
            const long long grow = (long long)b * nb + c * RB + row[u];
            *reinterpret_cast<uint4*>(p.hmB + grow * PC_K + k0) = hq;
            *reinterpret_cast<uint4*>(p.hmB + ((long long)p.rows_total + grow) * PC_K + k0) = mq;
            float4* fr = reinterpret_cast<float4*>(p.mhatB + grow * PC_K + k0);
            fr[0] = make_float4(mh[0], mh[1], mh[2], mh[3]);
            fr[1] = make_float4(mh[4], mh[5], mh[6], mh[7]);
          }
        }
      }
    }
  }
  PC_TRACE(0, 1);
  // generic-proxy writes (operand tile in shared memory, B rows in global memory) -> visible to the async proxy
  // (tcgen05.mma, TMA) of every CTA of the cluster once the barrier is passed
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  asm volatile("fence.proxy.async;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  PC_TRACE(0, 2);

  // ---- sweep over the cluster's B tiles ------------------------------------------------------------------------------
  // tile t of this CTA is the B rows of peer (c + t) % CS (its own first: L2-hot); peers that own no B rows are skipped by
  // every role alike
  int ntiles = 0;
  // producer and MMA warps converged, one elected lane issuing (one UTMALDG / UTCHMMA per issue instead of an ELECT / vote
  // loop of ~70 cycles around each: attn_f32.cu)
  if (warp == 0) {
    if (sweeping) {
      const bool leader = elect_one_sync();
      int it = 0;
      for (int t = 0; t < CS; ++t) {
        const int pt = (c + t) % CS;
        if (nb - pt * RB <= 0) continue;
        const int s = it % NST;
        mbar_wait(bar_empty + 8u * s, ((it / NST) & 1) ^ 1);
        if (leader) {
          const uint32_t st = base + L.stage0 + (uint32_t)s * L.stage_bytes, full = bar_full + 8u * s;
          mbar_expect_tx(full, L.stage_bytes);
          const int grow = b * nb + pt * RB;
          tma_load_2d(st, &map_b, 0, grow, full);                                       // h rows
          tma_load_2d(st + (uint32_t)BN * 128u, &map_b, 0, p.rows_total + grow, full);  // m rows
        }
        __syncwarp();
        ++it;
      }
    }
  } else if (warp == 1) {
    if (sweeping) {
      const bool leader = elect_one_sync();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_h = make_sw128_desc(base + L.a_h), a_m = make_sw128_desc(base + L.a_m);
      int it = 0;
      for (int t = 0; t < CS; ++t) {
        const int pt = (c + t) % CS;
        if (nb - pt * RB <= 0) continue;
        const int s = it % NST, acc = it & 1;
        mbar_wait(bar_tempty + 8u * acc, ((it >> 1) & 1) ^ 1);       // the workers have read this accumulator out
        if (leader) PC_TILE(32, it, 0);
        mbar_wait(bar_full + 8u * s, (it / NST) & 1);
        tc_fence_after();
        if (leader) {
          PC_TILE(32, it, 1);
          const uint32_t st = base + L.stage0 + (uint32_t)s * L.stage_bytes;
          const uint64_t b_h = make_sw128_desc(st), b_m = make_sw128_desc(st + (uint32_t)BN * 128u);
          const uint32_t d_tmem = tb + (uint32_t)acc * 128u;
#pragma unroll
          for (int k = 0; k < PC_K / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);        // +32 bytes inside the swizzle row
            umma_bf16(d_tmem, a_h + adv, b_h + adv, idesc, k ? 1u : 0u);
            umma_bf16(d_tmem, a_h + adv, b_m + adv, idesc, 1u);
            umma_bf16(d_tmem, a_m + adv, b_h + adv, idesc, 1u);
          }
          umma_commit(bar_empty + 8u * s);                             // the stage is free once these MMAs retire
          umma_commit(bar_tfull + 8u * acc);                           // and the accumulator is complete
          PC_TILE(32, it, 2);
          if (it == 0) PC_TRACE(32, 10);
        }
        __syncwarp();
        ++it;
      }
      if (leader) PC_TRACE(32, 11);
    }
  }
  // ---- workers: epilogue of every tile, refine, rank ---------------------------------------------------------------
  const int ew = warp - 2;                           // 0..7 for worker warps
  const int q = warp & 3;                            // TMEM lane quarter this warp may touch
  const int half = ew >> 2;                          // which half of a tile's columns
  const int row = q * 32 + lane;                     // A row of this CTA (TMEM lane)
  const int etid = half * 128 + row;                 // 0..255
  const int HB = BN >> 1;                            // columns per half tile (multiple of 8)
  const int gi = c * RA + row;                       // global A row
  if (warp >= 2) {
    float mrun = -INFINITY;
    if (sweeping) {
      int it = 0;
      for (int t = 0; t < CS; ++t) {
        const int pt = (c + t) % CS;
        const int ncol = min(RB, nb - pt * RB);
        if (ncol <= 0) continue;
        if (q < live_quarters) {                     // warp-uniform: quarters without rows of this CTA have nothing to read
          const int acc = it & 1;
          PC_TILE(128, it, 3);
          mbar_wait(bar_tfull + 8u * acc, (it >> 1) & 1);
          tc_fence_after();
          PC_TILE(128, it, 4);
          if (it == 0) PC_TRACE(128, 3);
          const uint32_t taddr = tmem_base + (uint32_t)acc * 128u + (uint32_t)(half * HB) + ((uint32_t)(q * 32) << 16);
          const int ncol_half = max(0, min(HB, ncol - half * HB));
          const bool mask_col0 = p.distill && pt == 0 && half == 0;      // column 0 is the distillation token
          float tm; uint32_t m0, m1;
          const uint32_t te = bar_tempty + 8u * acc;
          switch (HB >> 3) {
            case 1: sweep_tile<1>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 2: sweep_tile<2>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 3: sweep_tile<3>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 4: sweep_tile<4>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 5: sweep_tile<5>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 6: sweep_tile<6>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            case 7: sweep_tile<7>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
            default: sweep_tile<8>(taddr, ncol_half, mask_col0, p.window, mrun, tm, m0, m1, te, lane); break;
          }
          PC_TILE(128, it, 5);
          rec_tmax[it * PC_WORKERS + etid] = tm;
          rec_m0[it * PC_WORKERS + etid] = m0;
          rec_m1[it * PC_WORKERS + etid] = m1;
        }
        if (etid == 0) tile_pt[it] = pt;
        ++it;
      }
      ntiles = it;
    }
    PC_TRACE(128, 4);
    rowmax[half * 128 + row] = mrun;
    workers_sync();
    const float mrow = fmax_nan(rowmax[row], rowmax[128 + row]);
    // candidates: every column within the window of the row max (a superset: masks were cut against running maxima)
    const bool cls_row = p.cls && gi == 0;
    const bool live = sweeping && row < RAv && !cls_row;
    unsigned long long best = (row < RAv && cls_row) ? pack_best(-INFINITY, 0) : 0ull;
    if (q < live_quarters) {
      int ti = live ? -1 : ntiles;
      uint32_t b0 = 0u, b1 = 0u;
      int cur_pt = 0, colg = 0;
      bool has = false;
      const float* a_blk = a_f32 + (q * 32) * PC_K;
      const float* brows = p.mhatB + (long long)b * nb * PC_K;
      const float cut = mrow - p.window;                  // NaN row max: only NaN tiles qualify
      for (;;) {
        has = false;                                      // advance this lane to its next candidate
        while (ti < ntiles) {
          if (b0 | b1) {
            int e;
            if (b0) { e = __ffs(b0) - 1; b0 &= b0 - 1u; } else { e = 32 + __ffs(b1) - 1; b1 &= b1 - 1u; }
            colg = cur_pt * RB + half * HB + e;
            has = true;
            break;
          }
          if (++ti >= ntiles) break;
          const float tm = rec_tmax[ti * PC_WORKERS + etid];
          const bool nan_tile = tm != tm;
          if (!(nan_tile || tm >= cut)) continue;
          cur_pt = tile_pt[ti];
          if (nan_tile) {       // NaN in the tile (a zero-norm token): score every valid column of this half exactly
            const int cols = max(0, min(HB, min(RB, nb - cur_pt * RB) - half * HB));
            b0 = cols >= 32 ? 0xffffffffu : ((1u << cols) - 1u);
            b1 = cols > 32 ? (cols >= 64 ? 0xffffffffu : ((1u << (cols - 32)) - 1u)) : 0u;
          } else {
            b0 = rec_m0[ti * PC_WORKERS + etid];
            b1 = rec_m1[ti * PC_WORKERS + etid];
          }
        }
        if (!__any_sync(0xffffffffu, has)) break;
        float sc = warp_exact64(a_blk, brows, has, colg, lane);
        if (has) {
          if (p.distill && colg == 0) sc = -INFINITY;
          const unsigned long long k = pack_best(sc, colg);
          best = k > best ? k : best;
        }
      }
    }
    rowkey[half * 128 + row] = best;
    workers_sync();
    PC_TRACE(128, 5);
    if (half == 0 && row < RAv) {
      const unsigned long long k0 = rowkey[row], k1 = rowkey[128 + row];
      const unsigned long long final_key = k0 > k1 ? k0 : k1;
      const float nm = key_to_float((uint32_t)(final_key >> 32));
      const int ni = (int)(0xFFFFFFFFu - (uint32_t)(final_key & 0xFFFFFFFFull));
      p.node_max[(long long)b * na + gi] = nm;
      p.node_idx[(long long)b * na + gi] = ni;
      nidx[row] = ni;
      // this row's sort key to every CTA of the cluster (distributed shared memory)
      const uint32_t okey = (uint32_t)(final_key >> 32);
      const uint32_t slot = base + L.allkeys + 4u * (uint32_t)gi;
      for (int pr = 0; pr < CS; ++pr) st_cluster_u32(mapa_u32(slot, (uint32_t)pr), okey);
    }
  }
  __syncwarp();
  cluster_sync_all();
  PC_TRACE(128, 6);

  // ---- rank: position of every own row in a stable descending sort of the batch element's keys -----------------------
  // rank(i) = #{j : key_j > key_i, or key_j == key_i and j < i}.  Keys are read four at a time (warp-wide broadcast); a
  // chunk entirely below i counts key >= mine (as key > mine - 1), a chunk entirely above counts key > mine, and the one
  // chunk that contains i is corrected afterwards.
  if (warp >= 2 && q < live_quarters) {
    const bool valid = row < RAv;
    const uint32_t mine = valid ? allkeys[gi] : 0xFFFFFFFFu;
    const uint32_t tm1 = mine - 1u;                  // live keys are >= 0x007FFFFF (the key of -inf): no underflow
    const int na4 = (na + 3) & ~3;
    const int split = ((na4 >> 1) + 3) & ~3;
    const int jb = half ? split : 0, je = half ? na4 : split;
    int cnt = 0;
    const uint4* k4 = reinterpret_cast<const uint4*>(allkeys);
#pragma unroll 4
    for (int j = jb; j < je; j += 4) {
      const uint4 kk = k4[j >> 2];
      const uint32_t t = (j + 3 < gi) ? tm1 : mine;
      cnt += (int)(kk.x > t) + (int)(kk.y > t) + (int)(kk.z > t) + (int)(kk.w > t);
    }
    if (valid) {
      const int j0 = gi & ~3;
      if (j0 >= jb && j0 < je)
        for (int j = j0; j < gi; ++j) cnt += (int)(allkeys[j] == mine);
    }
    if (half) cnt2[row] = cnt;
    // the two halves of a row live in warps w and w + 4 of the same quarter: pair them up through a named barrier
    asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
    if (half == 0 && valid) {
      const int rank = cnt + cnt2[row];
      rankbuf[row] = rank;
      const long long br = (long long)b * r, ba = (long long)b * na;
      if (rank < r) {
        const int d = nidx[row];
        p.src_idx[br + rank] = gi;
        p.dst_idx[br + rank] = d;
        p.a_map[ba + gi] = -(d + 1);
        const uint32_t s_slot = base + L.edge_src + 4u * (uint32_t)rank, d_slot = base + L.edge_dst + 4u * (uint32_t)rank;
        for (int pr = 0; pr < CS; ++pr) {
          st_cluster_u32(mapa_u32(s_slot, (uint32_t)pr), (uint32_t)gi);
          st_cluster_u32(mapa_u32(d_slot, (uint32_t)pr), (uint32_t)d);
        }
      } else if (!p.cls) {
        p.unm_idx[(long long)b * (na - r) + rank - r] = gi;
        p.a_map[ba + gi] = rank - r;
      }
    }
  }
  PC_TRACE(128, 7);
  __syncwarp();
  cluster_sync_all();
  PC_TRACE(128, 8);

  // ---- csr: sources grouped by destination (ascending k inside a group: the reference CPU scatter_reduce order) -------
  if (warp >= 2) {
    const int r4 = (r + 3) & ~3;
    if (half == 0) {
      if (row < RBv) {                               // one thread per own B token: offset, count, first two sources
        const int j = c * RB + row;
        int before = 0, cnt = 0, s0 = 0, s1 = 0;
        const int4* d4 = reinterpret_cast<const int4*>(edge_dst);
#pragma unroll 2
        for (int k = 0; k < r4; k += 4) {
          const int4 dd = d4[k >> 2];
          before += (int)(dd.x < j) + (int)(dd.y < j) + (int)(dd.z < j) + (int)(dd.w < j);
          if (dd.x == j || dd.y == j || dd.z == j || dd.w == j) {        // rare
            const int dv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (dv[u] == j) {
                if (cnt == 0) s0 = edge_src[k + u]; else if (cnt == 1) s1 = edge_src[k + u];
                ++cnt;
              }
          }
        }
        p.b_off[(long long)b * (nb + 1) + j] = before;
        reinterpret_cast<int4*>(p.b_head)[(long long)b * nb + j] = make_int4(cnt, s0, s1, before);
      }
      if (c == 0 && row == 0) p.b_off[(long long)b * (nb + 1) + nb] = r;
    } else if (p.cls && row < RAv && rankbuf[row] >= r) {
      // merge.py:71-73: kept tokens in ascending index order -> position = own index minus the merged tokens before it
      int merged_before = 0;
      for (int k = 0; k < r; ++k) merged_before += (int)(edge_src[k] < gi);
      const int pos = gi - merged_before;
      p.unm_idx[(long long)b * (na - r) + pos] = gi;
      p.a_map[(long long)b * na + gi] = pos;
    }
  } else {
    // warps 0 and 1: b_src = the edges in a stable counting sort by destination; this CTA places its share of the edges
    const int RE = (r + CS - 1) / CS;
    const int ke = min(r, (c + 1) * RE);
    for (int k = c * RE + (int)threadIdx.x; k < ke; k += 64) {
      const int d = edge_dst[k];
      int pos = 0;
      for (int k2 = 0; k2 < r; ++k2) {
        const int d2 = edge_dst[k2];
        pos += (int)(d2 < d) + (int)(d2 == d && k2 < k);
      }
      p.b_src[(long long)b * r + pos] = edge_src[k];
    }
  }
  PC_TRACE(128, 9);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PC_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
struct PcGeom { int CS, RA, RB, BN, stages; size_t smem; };

static bool pc_geometry(int bm, int n, PcGeom& g) {
  const int na = na_of(n), nb = nb_of(n);
  if (nb < 1) return false;
  int cs = 1;
  while (cs < PC_MAXCS && (na + cs - 1) / cs > 128) cs *= 2;
  if ((na + cs - 1) / cs > 128) return false;
  // more CTAs per batch element while the grid still leaves SMs idle and the row blocks stay worth a tile: prep, epilogue
  // and rank work per CTA all shrink with the cluster size
  const char* e = getenv("TOME_PLAN_CS");
  if (e && atoi(e) > 0) {
    const int f = atoi(e);
    while (cs < f && cs < PC_MAXCS) cs *= 2;
  } else {
    while (cs < 8 && (long long)bm * cs * 2 <= 148 && (na + 2 * cs - 1) / (2 * cs) >= 24) cs *= 2;
  }
  g.CS = cs;
  g.RA = (na + cs - 1) / cs;
  g.RB = (nb + cs - 1) / cs;
  g.BN = ((g.RB + 15) / 16) * 16;
  if (g.BN < 16) g.BN = 16;
  if (g.BN > 128) return false;
  // as many B tiles in flight as shared memory holds (every CTA streams the whole B set of its batch element)
  const size_t limit = 227 * 1024;
  int st = cs < PC_MAX_STAGES ? cs : PC_MAX_STAGES;
  while (st > 1 && pc_smem(cs, g.BN, st).total + 1024 > limit) --st;
  g.stages = st;
  g.smem = pc_smem(cs, g.BN, st).total + 1024;
  return g.smem <= limit;
}

size_t plan_cluster_workspace(int bm, int n) {
  const size_t rows = (size_t)bm * nb_of(n);
  return ((2 * rows * PC_K * sizeof(__nv_bfloat16) + 255) & ~(size_t)255) + ((rows * PC_K * sizeof(float) + 255) & ~(size_t)255);
}

bool plan_cluster_supported(int dtype, int bm, int n, int cm, int heads, const View& v, long long stride_h, const void* metric) {
  // Measured on a B200 (profiles/r02_plan_paths.txt, graph replay, bf16 keys of 12 heads): the one launch beats the
  // four-launch chain while a batch element's A rows fit ONE 128-row tile (TimeSformer / Motionformer frames, the last
  // ViViT layers: 20.8 / 17.6 / 14.0 us against 22.9 / 21.7 / 18.7 at n = 196 / 88 / 17), and loses beyond (39 against
  // 26 us at n = 1568): a cluster per batch element reaches 64 of the 148 SMs, and both the normalisation (~800
  // instructions per token and lane) and the TMEM read-out of the sweep (64 B/clk/SM) scale with the SMs in use.
  // TOME_PLAN_CLUSTER=1 forces it for every shape it supports, =0 disables it.
  const char* sw = getenv("TOME_PLAN_CLUSTER");
  if (sw && atoi(sw) == 0) return false;
  if (!(sw && atoi(sw) == 1) && na_of(n) > 128) return false;
  if (dtype != TOME_F32 && dtype != TOME_BF16) return false;
  if (cm > PC_K || cm % 8 != 0 || n < 2 || bm > TOME_MAX_BATCH) return false;
  const long long vec = dtype == TOME_F32 ? 4 : 8;               // 16-byte loads of eight channels
  if (((uintptr_t)metric & 15) || (v.sbo % vec) || (v.sbi % vec) || (v.sn % vec) || (heads > 1 && (stride_h % vec))) return false;
  PcGeom g;
  return pc_geometry(bm, n, g);
}

template <typename T>
static int launch_pc_t(const CUtensorMap& map_b, const void* metric, const PcParams& p, const PcGeom& g, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first_time()) {
    TOME_CUDA(cudaFuncSetAttribute(plan_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TOME_CUDA(cudaFuncSetAttribute(plan_cluster_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.CS, p.bm, 1);
  cfg.blockDim = dim3(PC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = g.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (g.CS == 1) {          // a batch element per CTA: a plain launch (every CTA is its own cluster; no cluster scheduling cost)
    plan_cluster_kernel<T><<<cfg.gridDim, cfg.blockDim, cfg.dynamicSmemBytes, st>>>(map_b, (const T*)metric, p);
    e = cudaGetLastError();
  } else {
    e = cudaLaunchKernelEx(&cfg, plan_cluster_kernel<T>, map_b, (const T*)metric, p);
  }
  count_launch();
  if (e != cudaSuccess) return set_error(TOME_ERR_CUDA, "launch of plan_cluster_kernel failed: %s", cudaGetErrorString(e));
  return TOME_OK;
}

int launch_plan_cluster(const void* metric, int dtype, int heads, long long stride_h, const View& v, int cm, const tome_plan* plan,
                        void* ws, size_t ws_bytes, cudaStream_t st) {
  const int bm = plan->bm, n = plan->n;
  PcGeom g;
  if (!pc_geometry(bm, n, g)) return set_error(TOME_ERR_UNSUPPORTED, "tome_plan_build: shape bm=%d n=%d does not fit the cluster kernel", bm, n);
  if (ws_bytes < plan_cluster_workspace(bm, n))
    return set_error(TOME_ERR_WORKSPACE, "tome_plan_build: workspace %zu < %zu bytes", ws_bytes, plan_cluster_workspace(bm, n));
  PcParams p;
  p.bm = bm; p.n = n; p.na = na_of(n); p.nb = nb_of(n); p.cm = cm; p.cls = plan->class_token; p.distill = plan->distill_token;
  p.heads = heads; p.r = plan->r; p.stride_h = stride_h; p.v = v;
  p.CS = g.CS; p.RA = g.RA; p.RB = g.RB; p.BN = g.BN; p.stages = g.stages;
  p.rows_total = bm * p.nb;
  const float eps = 5e-5f + 2e-7f * (float)cm;        // error bound of h.h + h.m + m.h on unit vectors (match_sm100.cu)
  p.window = 2.0f * eps;
  const size_t rows = (size_t)bm * p.nb;
  p.hmB = (__nv_bfloat16*)ws;
  p.mhatB = (float*)((char*)ws + ((2 * rows * PC_K * sizeof(__nv_bfloat16) + 255) & ~(size_t)255));
  p.node_max = const_cast<float*>(plan->node_max); p.node_idx = const_cast<int*>(plan->node_idx);
  p.src_idx = plan->src_idx; p.unm_idx = plan->unm_idx; p.dst_idx = plan->dst_idx; p.a_map = plan->a_map;
  p.b_off = plan->b_off; p.b_src = plan->b_src; p.b_head = plan->b_head;
  p.trace = getenv("TOME_PC_TRACE") ? (long long*)strtoull(getenv("TOME_PC_TRACE"), nullptr, 0) : nullptr;
  alignas(64) CUtensorMap map_b;
  int rc = make_bf16_map(&map_b, p.hmB, 2LL * rows, PC_K, PC_K, g.BN, "tome_plan_build");
  if (rc) return rc;
  if (dtype == TOME_F32) return launch_pc_t<float>(map_b, metric, p, g, st);
  return launch_pc_t<__nv_bfloat16>(map_b, metric, p, g, st);
}

void plan_cluster_describe(int bm, int n, long long out[5]) {
  PcGeom g;
  if (!pc_geometry(bm, n, g)) { out[0] = out[1] = out[2] = out[3] = out[4] = 0; return; }
  out[0] = g.CS; out[1] = g.RA; out[2] = g.RB; out[3] = g.BN; out[4] = (long long)g.smem;
}

}  // namespace tome
