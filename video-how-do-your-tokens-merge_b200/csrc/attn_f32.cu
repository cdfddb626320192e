// fp32 attention on the tensor cores, at fp32 accuracy: softmax(scale * q k^T + key bias) v for the fp32
// models -- the reference benchmark runs fp32 with TF32 off (slowfast/utils/model_benchmark.py:21-45) -- where torch's
// fused fp32 attention was 42 % of the patched VideoMAE step once the linear layers had moved to tcgen05
// (profiles/r02_videomae_fp32_launches_after.csv).  Same idea as linear_f32.cu: q, k, v arrive as exact three-way bf16
// splits (h + m + l == x, tome_split3 of the QKV GEMM's output) and every bf16 product is exact in the fp32 accumulator.
//     S = Q K^T   six plane products  (h.h, h.m, m.h, m.m, h.l, l.h : what is dropped is <= 3 * 2^-24 |q||k|)
//     P           fp32 softmax numerators, split exactly into three bf16 planes in registers
//     O = P V     six plane products, V consumed as TMA wrote it (rows = keys: MN-major B operand)
// Flash-style: one CTA per (128 queries, head, clip) walks the keys in blocks of 64 with a running maximum; the block's
// scores are read out of TMEM ONCE (64 values per row), the block's P V lands in its own TMEM columns and is folded
// into the row's fp32 accumulator in registers with the usual rescale, so no accumulator is ever rescaled in TMEM.
//   warp 0      TMA producer: K / V planes of each key block (3-stage ring, SWIZZLE_128B, separate K and V barriers);
//   warp 1      S issuer, warp 10 P V issuer: converged warps, one elected lane, one UTCHMMA per MMA;
//   warps 2-9   TWO threads per query row, 32 key columns of the block and 32 channels of O each: Q planes into TMEM once,
//               then per block scores -> row maximum (halves exchanged through shared memory) -> exp2 / sum -> three P
//               planes into TMEM (tcgen05.st) -> accumulate O.
// History of the kernel (profiles/r02_attn_f32_ncu.txt): 637 us at 8 x 12 x 1568 with one barrier pair per stage; 480 us with
// separate K / V barriers; 326 us once the MMA warp ran converged (as `if (lane == 0)` every tcgen05.mma sat in an ELECT /
// vote loop of ~70 cycles, twice the MMA's own 32-cycle slot) with the A operands in tensor memory; 279 us with the two
// issue streams in separate warps, where the tensor pipe is busy back to back in steady state (1536 cycles per block).
#include <cstdlib>

#include "tc_ptx.cuh"

namespace tome {

constexpr int AF_BM = 128, AF_BKV = 64, AF_D = 64;
constexpr int AF_SM = 256;                                                  // softmax threads: two per query row
constexpr uint32_t AF_KP = AF_BKV * 128u;           // one K / V plane of a key block: 64 rows x 128 bytes
constexpr uint32_t AF_STAGE = 6u * AF_KP;           // K h,m,l | V h,m,l

// One problem per (clip, frame): the S = F * P queries of a clip (tokens tok0 .. tok0 + S - 1 of its N) against the P keys of
// frame f, softmax per frame -- the space stage of Motionformer's trajectory attention (tome/patch/motionformer.py:105-115).
// Plain attention is the case F = 1, P = N, tok0 = 0: xs (B, S, 1, C) is then the (B, N, C) output.
struct AfParams {
  int B, N, heads, nblk, nobias_q;
  int F, P, S, tok0;
  float scale_log2e;
  const float* bias;                                // (B, F * P) log size per key, or NULL
  float* out;                                       // xs (B, S, F, heads * 64), or NULL
  __nv_bfloat16* out3;                              // (B * S * F, 3 * heads * 64) bf16 planes of the same values (the next GEMM's operand), or NULL
  float* diag;                                      // (B, S, heads * 64): xs[b, s, frame(s)] (the trajectory "diagonal"), or NULL
};

__device__ __forceinline__ float af_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void af_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t af_desc_mn(uint32_t smem_addr) {     // MN-major SWIZZLE_128B (see attn_frames.cu)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t af_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- A operands in tensor memory ----------------------------------------------------------------------------------------
// An SS-mode 128 x 64 x 16 MMA reads 6 KB of operands per 32-cycle slot (192 B / cycle against a 128 B / cycle shared-memory
// port).  Both A operands live in tensor memory instead -- Q for the whole CTA (three planes x 32 columns, written once by the softmax threads
// straight from global memory: no Q tile, no TMA for it) and P (three planes x 32 columns, written with tcgen05.st in the
// layout the softmax threads already hold: lane = row, column c = keys 2c, 2c + 1) -- so every MMA reads only its 2 KB B
// operand from shared memory and runs at the tensor pipe's floor.  The 96 KB of Q / P tiles that leaves pays for a third
// K / V stage.  TMEM columns: S0 0..63 | S1 64..127 | O 128..191 | P h,m,l 192..287 | Q h,m,l 288..383 (512 allocated).
constexpr int AT_NST = 3;
// -DTOME_ATTN_TRACE: clock64 stamps of one CTA's issue / softmax events per key block (tools/trace_attn_f32.py; the timeline
// that found the serialised issue streams is in profiles/r02_attn_f32_ncu.txt)
#ifdef TOME_ATTN_TRACE
__device__ long long g_af_trace[8 * 32];
#define AF_TR(ev, j) do { if (tr && (j) < 32) g_af_trace[(ev) * 32 + (j)] = clock64(); } while (0)
#else
#define AF_TR(ev, j) do { } while (0)
#endif
constexpr uint32_t AT_S1 = 64u, AT_O = 128u, AT_P = 192u, AT_Q = 288u;

// p -> exact bf16 planes of a pair of probabilities without the single-value F2F conversions (XU pipe, 16 / clk / SM: with 64 of
// them per thread and block that pipe was as busy as the tensor pipe): the packed conversion is an ALU instruction and a
// bf16 is the upper half of its fp32
__device__ __forceinline__ void af_split_pair(float p0, float p1, uint32_t& wh, uint32_t& wm, uint32_t& wl) {
  wh = af_pack(p0, p1);
  const float r0 = p0 - __uint_as_float(wh << 16), r1 = p1 - __uint_as_float(wh & 0xffff0000u);
  wm = af_pack(r0, r1);
  wl = af_pack(r0 - __uint_as_float(wm << 16), r1 - __uint_as_float(wm & 0xffff0000u));
}

constexpr int AT_THREADS = 352;     // TMA, S issuer, 2 x 4 softmax warps, P V issuer

template <bool HAS_BIAS, bool FRAMES>                   // FRAMES: more than one frame per clip (the loop carries frame boundaries)
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_f32_kernel(const __grid_constant__ CUtensorMap map_kv, const __nv_bfloat16* __restrict__ qkv3, const AfParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler, see the MMA warp
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C = p.heads * AF_D, C3 = 3 * C;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sm_kv = base, bars = sm_kv + (uint32_t)AT_NST * AF_STAGE;
  const uint32_t bar_q = bars, bar_p = bars + 8, bar_o = bars + 16, tmem_slot = bars + 24, bar_s = bars + 32, bar_sfree = bars + 48,
                 bar_kfull = bars + 64, bar_kempty = bars + 96, bar_vfull = bars + 128, bar_vempty = bars + 160, sm_xch = bars + 192;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_kv);
    mbar_init(bar_q, AF_SM); mbar_init(bar_p, AF_SM); mbar_init(bar_o, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_s + 8u * s, 1); mbar_init(bar_sfree + 8u * s, AF_SM); }
    for (int s = 0; s < AT_NST; ++s) {
      mbar_init(bar_kfull + 8u * s, 1); mbar_init(bar_kempty + 8u * s, 1);
      mbar_init(bar_vfull + 8u * s, 1); mbar_init(bar_vempty + 8u * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int row0 = b * p.N + p.tok0;                 // first query token of the clip == first key of frame 0
  const int nbf = p.nblk;                            // 64-key blocks per frame
  const int nb = FRAMES ? p.F * nbf : nbf;           // the CTA walks the blocks of all frames: Q, TMEM and the pipeline are set up once
  // the softmax threads' query loads go out before the TMEM allocation and the barrier, so that their latency (the planes of
  // a 173 MB tensor: mostly HBM) overlaps the CTA's set-up: 32 channels (64 bytes) per plane and thread; rows past the tensor
  // are zero
  uint4 qv[12];
  if (warp >= 2 && warp < 10) {
    const int s_q = qt * AF_BM + (warp & 3) * 32 + lane;
    const bool in = (long long)row0 + s_q < (long long)p.B * p.N;
    const uint4* src = reinterpret_cast<const uint4*>(qkv3 + ((long long)row0 + s_q) * (3LL * C3) + h * AF_D + 32 * ((warp - 2) >> 2));
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[4 * pl + i] = in ? __ldg(src + (size_t)pl * (C3 / 8) + i) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
#ifdef TOME_ATTN_TRACE
  const bool tr = blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < nb; ++j) {
        const uint32_t s = (uint32_t)(j % AT_NST), k = (uint32_t)(j / AT_NST);
        const uint32_t st = sm_kv + s * AF_STAGE;
        const int fj = FRAMES ? j / nbf : 0;
        const int krow = row0 + fj * p.P + (j - fj * nbf) * AF_BKV;         // block j - fj * nbf of frame fj
        if (k >= 1) mbar_wait_sleep(bar_kempty + 8u * s, (k - 1) & 1u, 32);
        mbar_expect_tx(bar_kfull + 8u * s, 3u * AF_KP);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tma_load_2d(st + pl * AF_KP, &map_kv, pl * C3 + C + h * AF_D, krow, bar_kfull + 8u * s);
        if (k >= 1) mbar_wait_sleep(bar_vempty + 8u * s, (k - 1) & 1u, 32);
        mbar_expect_tx(bar_vfull + 8u * s, 3u * AF_KP);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tma_load_2d(st + (3 + pl) * AF_KP, &map_kv, pl * C3 + 2 * C + h * AF_D, krow, bar_vfull + 8u * s);
      }
    }
  } else if (warp == 1) {
    // The MMA warp runs CONVERGED with warp-uniform operands and one elected lane issuing: written as `if (lane == 0)` the
    // compiler cannot prove the operands uniform and wraps every tcgen05.mma in an ELECT / vote loop with register ->
    // uniform-register moves -- ~70 cycles per MMA measured (profiles/r02_attn_f32_ncu.txt), twice the 32-cycle slot of a
    // 128 x 64 x 16 MMA, and the issue of P V sits on the block-to-block critical path.  This form is one UTCHMMA per MMA.
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AF_BKV >> 3) << 17) | ((uint32_t)(AF_BM >> 4) << 24);
    const uint64_t dk0 = make_sw128_desc(sm_kv);                            // stage 0, plane h; a plane / stage further is
    constexpr uint64_t PL = AF_KP >> 4, ST = AF_STAGE >> 4;                               // a constant in the address field
    auto issue_s = [&](int j) {
      const uint32_t s = (uint32_t)(j % AT_NST), k = (uint32_t)(j / AT_NST), sb = (uint32_t)(j & 1), kb = (uint32_t)(j >> 1);
      mbar_wait(bar_kfull + 8u * s, k & 1u);
      if (kb >= 1) mbar_wait(bar_sfree + 8u * sb, (kb - 1) & 1u);           // the softmax has pulled the previous S out of this buffer
      tc_fence_after();
      AF_TR(0, j);
      if (leader) {
        const uint32_t d = tb + sb * AT_S1;
        const uint64_t dk = dk0 + (uint64_t)s * ST;
        // (query plane, key plane): l.h, h.l, m.m, m.h, h.m, h.h -- smallest products first; A: 8 columns per k-step,
        // K: +32 bytes inside the swizzle row
#define TOME_AT_S(QP_, KP_, ACC_)                                                                                        \
        _Pragma("unroll") for (int ks = 0; ks < AF_D / 16; ++ks)                                                         \
          umma_bf16_ts(d, tb + AT_Q + 32u * (QP_) + 8u * ks, dk + (KP_) * PL + (uint64_t)(2 * ks), idesc_s, (ACC_) || ks ? 1u : 0u);
        TOME_AT_S(2, 0, 0) TOME_AT_S(0, 2, 1) TOME_AT_S(1, 1, 1) TOME_AT_S(1, 0, 1) TOME_AT_S(0, 1, 1) TOME_AT_S(0, 0, 1)
#undef TOME_AT_S
        umma_commit(bar_s + 8u * sb);
        umma_commit(bar_kempty + 8u * s);
      }
      __syncwarp();
      AF_TR(1, j);
    };
    mbar_wait(bar_q, 0);
    for (int j = 0; j < nb; ++j) issue_s(j);                                // up to two blocks ahead of the softmax (S is double-buffered)
  } else if (warp == 10) {
    // P V from a warp of its own: in one warp the two issue streams serialise -- an MMA enters the pipe only as fast as
    // the pipe drains (~800 cycles per batch of 24) and every mbarrier poll between the batches is a ~100-cycle round
    // trip, 2250 cycles per block against 1536 of tensor time (the timeline in profiles/r02_attn_f32_ncu.txt)
    const bool leader = elect_one_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AF_D >> 3) << 17) | ((uint32_t)(AF_BM >> 4) << 24);
    const uint64_t dv0 = af_desc_mn(sm_kv + 3u * AF_KP);
    constexpr uint64_t PL = AF_KP >> 4, ST = AF_STAGE >> 4;
    for (int j = 0; j < nb; ++j) {
      const uint32_t s = (uint32_t)(j % AT_NST);
      mbar_wait(bar_vfull + 8u * s, (uint32_t)((j / AT_NST) & 1));
      mbar_wait(bar_p, (uint32_t)(j & 1));
      tc_fence_after();
      AF_TR(2, j);
      if (leader) {
        const uint32_t d = tb + AT_O;
        const uint64_t dv = dv0 + (uint64_t)s * ST;
        // (P plane, V plane), smallest products first; P: 8 columns per 16 keys; V (MN-major): 16 keys = 2048 bytes
#define TOME_AT_O(PP_, VP_, ACC_)                                                                                        \
        _Pragma("unroll") for (int ks = 0; ks < AF_BKV / 16; ++ks)                                                       \
          umma_bf16_ts(d, tb + AT_P + 32u * (PP_) + 8u * ks, dv + (VP_) * PL + (uint64_t)(128 * ks), idesc_o, (ACC_) || ks ? 1u : 0u);
        TOME_AT_O(2, 0, 0) TOME_AT_O(0, 2, 1) TOME_AT_O(1, 1, 1) TOME_AT_O(1, 0, 1) TOME_AT_O(0, 1, 1) TOME_AT_O(0, 0, 1)
#undef TOME_AT_O
        umma_commit(bar_o);
        umma_commit(bar_vempty + 8u * s);
      }
      __syncwarp();
      AF_TR(3, j);
    }
  } else if (warp >= 2) {
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;                          // which 32 key columns of a block / 32 channels of O
    const int row = q4 * 32 + lane;
    const int s_idx = qt * AF_BM + row;                        // query token within the clip
    const bool live = s_idx < p.S;
    const bool biased = p.bias != nullptr && s_idx >= p.nobias_q;
    const uint32_t tlane = (uint32_t)(q4 * 32) << 16;
    const float* bias_b = p.bias ? p.bias + (long long)b * p.S : nullptr;
    const bool bias_vec = ((p.P | p.S) & 3) == 0;
    float* xch = reinterpret_cast<float*>(gen + (sm_xch - base));          // [2 blocks][2 halves][128 rows]
    float* lx = xch + 4 * AF_BM;                                           // [2 frames][2 halves][128 rows]: a finished frame's row sums
    const float LOG2E = 1.4426950408889634f;
    {   // this row's query planes (loaded above) into tensor memory
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 v = qv[4 * pl + i];
          w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        tmem_st16(tmem_base + tlane + AT_Q + 32u * pl + 16u * half, w);
      }
      tmem_st_wait();
      tc_fence_before();
      af_arrive(bar_q);
    }
    float m = -INFINITY, l = 0.f;
    float oacc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) oacc[e] = 0.f;
    auto fold_o = [&]() {                                      // O_acc += the finished P V block
      float v[32];
      tmem_ld32(tmem_base + tlane + AT_O + 32u * half, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) oacc[e] += v[e];
    };
    // the finished frame fd: O / l -> xs (B, S, F, C), its planes, the trajectory diagonal
    auto write_frame = [&](int fd, float lsum) {
      if (!live) return;
      const float inv = 1.0f / lsum;
      float o[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = oacc[e] * inv;
      const long long orow = ((long long)b * p.S + s_idx) * p.F + fd;      // row of xs
      if (p.out) {
        float4* dst = reinterpret_cast<float4*>(p.out + orow * C + h * AF_D + 32 * half);
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
      }
      if (p.diag && s_idx / p.P == fd) {                                   // the query's own frame
        float4* dst = reinterpret_cast<float4*>(p.diag + ((long long)b * p.S + s_idx) * C + h * AF_D + 32 * half);
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
      }
      if (p.out3) {
        __nv_bfloat16* d3 = p.out3 + orow * 3 * C + h * AF_D + 32 * half;
#pragma unroll
        for (int e = 0; e < 4; ++e) store_planes8(d3 + 8 * e, C, reinterpret_cast<const float(&)[8]>(o[8 * e]));
      }
    };
    for (int j = 0; j < nb; ++j) {
      const uint32_t sb = (uint32_t)(j & 1), k = (uint32_t)(j >> 1);
      const int f = FRAMES ? j / nbf : 0, jf = j - f * nbf;    // frame, block within the frame
      const bool frame_start = FRAMES && jf == 0 && j > 0;     // (warp-uniform) the previous block closed frame f - 1
      if (frame_start) {                                       // park that frame's row sum for the other half, start afresh:
        lx[(((f - 1) & 1) * 2 + half) * AF_BM + row] = l;      // alpha below is 0 (m = -inf), which also clears the accumulator
        m = -INFINITY;
      }
      const float* brow = HAS_BIAS ? bias_b + (long long)f * p.P : nullptr;      // this frame's keys
      mbar_wait(bar_s + 8u * sb, k & 1u);
      tc_fence_after();
      if (warp == 2) AF_TR(4, j);
      float t[32];
      tmem_ld32(tmem_base + tlane + sb * AT_S1 + 32u * half, t);
      tc_fence_before();
      af_arrive(bar_sfree + 8u * sb);
      const int key0 = jf * AF_BKV + 32 * half;
      if (HAS_BIAS) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (biased && bias_vec && key0 + e + 3 < p.P) b4 = __ldg(reinterpret_cast<const float4*>(brow + key0 + e));
          else if (biased && key0 + e + 3 < p.P) {             // rows of the bias are not 16-byte aligned (n % 4 != 0)
            b4.x = __ldg(brow + key0 + e); b4.y = __ldg(brow + key0 + e + 1); b4.z = __ldg(brow + key0 + e + 2); b4.w = __ldg(brow + key0 + e + 3);
          } else if (biased) {
            b4.x = key0 + e < p.P ? __ldg(brow + key0 + e) : 0.f;
            b4.y = key0 + e + 1 < p.P ? __ldg(brow + key0 + e + 1) : 0.f;
            b4.z = key0 + e + 2 < p.P ? __ldg(brow + key0 + e + 2) : 0.f;
          }
          t[e] = fmaf(t[e], p.scale_log2e, b4.x * LOG2E);
          t[e + 1] = fmaf(t[e + 1], p.scale_log2e, b4.y * LOG2E);
          t[e + 2] = fmaf(t[e + 2], p.scale_log2e, b4.z * LOG2E);
          t[e + 3] = fmaf(t[e + 3], p.scale_log2e, b4.w * LOG2E);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) t[e] *= p.scale_log2e;    // (the reference's q * scale, in log2 units)
      }
      if (key0 + 32 > p.P) {                                   // last block: keys beyond the frame / clip
#pragma unroll
        for (int e = 0; e < 32; ++e) if (key0 + e >= p.P) t[e] = -INFINITY;
      }
      float bmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; e += 4) bmax = fmaxf(bmax, fmaxf(fmaxf(t[e], t[e + 1]), fmaxf(t[e + 2], t[e + 3])));
      xch[(sb * 2 + half) * AF_BM + row] = bmax;               // the row's other half
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bmax = fmaxf(bmax, xch[(sb * 2 + (half ^ 1)) * AF_BM + row]);
      const float m_new = fmaxf(m, bmax);                      // finite: the first half of every block holds a real key
      const float alpha = af_ex2(m - m_new);                   // 0 on the first block (m = -inf)
      float l0 = 0.f, l1 = 0.f;
      uint32_t wh[16], wm[16], wl[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = af_ex2(t[2 * i] - m_new), p1 = af_ex2(t[2 * i + 1] - m_new);
        l0 += p0; l1 += p1;
        af_split_pair(p0, p1, wh[i], wm[i], wl[i]);
      }
      if (warp == 2) AF_TR(5, j);
      if (j > 0) {                                             // P V of the previous block: also frees the P columns
        mbar_wait(bar_o, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (warp == 2) AF_TR(6, j);
        fold_o();
      }
      if (frame_start)                                         // the other half's sum became visible at this block's barrier
        write_frame(f - 1, l + lx[(((f - 1) & 1) * 2 + (half ^ 1)) * AF_BM + row]);
#pragma unroll
      for (int e = 0; e < 32; ++e) oacc[e] *= alpha;
      l *= alpha;
      m = m_new;
      // the three P planes: this thread's 32 keys are columns 16 * half .. + 15 of each plane (two keys per column)
      tmem_st16(tmem_base + tlane + AT_P + 16u * half, wh);
      tmem_st16(tmem_base + tlane + AT_P + 32u + 16u * half, wm);
      tmem_st16(tmem_base + tlane + AT_P + 64u + 16u * half, wl);
      l += l0 + l1;
      tmem_st_wait();
      tc_fence_before();                                       // the O columns were read above: P V (j) may overwrite them
      af_arrive(bar_p);
      if (warp == 2) AF_TR(7, j);
    }
    mbar_wait(bar_o, (uint32_t)((nb - 1) & 1));
    tc_fence_after();
    fold_o();
    asm volatile("bar.sync 1, 256;" ::: "memory");             // the row sum of both halves (same running maximum)
    xch[half * AF_BM + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xch[(half ^ 1) * AF_BM + row];
    write_frame(FRAMES ? p.F - 1 : 0, l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// qkv3 (B * N, 9 * heads * 64) bf16 planes; per clip the queries are tokens tok0 .. tok0 + F * P - 1 and frame f's keys tokens
// tok0 + f * P .. + P - 1 (plain attention: F = 1, P = N, tok0 = 0)
int launch_frames_attention_f32(const void* qkv3, int B, int N, int heads, int F, int P, int tok0, float scale, const float* bias,
                                int nobias_q, void* out, void* out3, void* diag, cudaStream_t st) {
  if (!out && !out3) return set_error(TOME_ERR_ARG, "tome_attention_f32: no output");
  if (F < 1 || P < 1 || tok0 < 0 || N != tok0 + F * P)
    return set_error(TOME_ERR_ARG, "tome_frames_attention_f32: N=%d != lead + F*P (lead=%d F=%d P=%d)", N, tok0, F, P);
  if (((uintptr_t)qkv3 & 15) || ((uintptr_t)out & 15) || ((uintptr_t)out3 & 15) || ((uintptr_t)diag & 15) || (bias && ((uintptr_t)bias & 15)))
    return set_error(TOME_ERR_ALIGN, "tome_attention_f32: buffers must be 16-byte aligned");
  AfParams p;
  p.B = B; p.N = N; p.heads = heads; p.nblk = (P + AF_BKV - 1) / AF_BKV; p.nobias_q = nobias_q;
  p.F = F; p.P = P; p.S = F * P; p.tok0 = tok0;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias; p.out = (float*)out; p.out3 = (__nv_bfloat16*)out3; p.diag = (float*)diag;
  const long long rows = (long long)B * N, cols = 9LL * heads * AF_D;
  alignas(64) CUtensorMap map_kv;
  int rc = make_bf16_map(&map_kv, qkv3, rows, cols, cols, AF_BKV, "tome_attention_f32");
  if (rc) return rc;
  if (B > 65535) return set_error(TOME_ERR_UNSUPPORTED, "tome_attention_f32: batch %d > 65535", B);
  dim3 grid((p.S + AF_BM - 1) / AF_BM, heads, B);
  const size_t smem = 1024 + AT_NST * AF_STAGE + 192 + 8 * AF_BM * sizeof(float);
  static PerDeviceOnce once;
  if (once.first_time()) {
    TOME_CUDA(cudaFuncSetAttribute(attn_f32_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOME_CUDA(cudaFuncSetAttribute(attn_f32_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOME_CUDA(cudaFuncSetAttribute(attn_f32_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TOME_CUDA(cudaFuncSetAttribute(attn_f32_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const __nv_bfloat16* q3 = (const __nv_bfloat16*)qkv3;
  if (F > 1) {
    if (bias) attn_f32_kernel<true, true><<<grid, AT_THREADS, smem, st>>>(map_kv, q3, p);
    else attn_f32_kernel<false, true><<<grid, AT_THREADS, smem, st>>>(map_kv, q3, p);
  } else {
    if (bias) attn_f32_kernel<true, false><<<grid, AT_THREADS, smem, st>>>(map_kv, q3, p);
    else attn_f32_kernel<false, false><<<grid, AT_THREADS, smem, st>>>(map_kv, q3, p);
  }
  TOME_LAUNCH_CHECK("attn_f32_kernel");
  return TOME_OK;
}

int launch_attention_f32(const void* qkv3, int B, int N, int heads, float scale, const float* bias, int nobias_q, void* out,
                         void* out3, cudaStream_t st) {
  return launch_frames_attention_f32(qkv3, B, N, heads, 1, N, 0, scale, bias, nobias_q, out, out3, nullptr, st);
}

}  // namespace tome

#ifdef TOME_ATTN_TRACE
extern "C" TOME_API int tome_debug_attn_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, tome::g_af_trace, sizeof(long long) * 8 * 32);
}
#endif
