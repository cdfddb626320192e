/*
 * tome_b200.h -- C ABI of the B200 (sm_100a) token-merging hot path.
 *
 * Drop-in boundary for the per-block ToMe step of sjpollard/video-how-do-your-tokens-merge:
 *   tome/merge.py:17-102   bipartite_soft_matching   -> tome_match + tome_select
 *   tome/merge.py:75-85    merge() closure           -> tome_merge
 *   tome/merge.py:87-100   unmerge() closure         -> tome_unmerge
 *   tome/merge.py:215-271  bipartite_soft_matching_drop   -> tome_merge(mode = TOME_MODE_DROP)
 *   tome/merge.py:274-352  bipartite_soft_matching_hybrid -> tome_merge(hybrid_threshold)
 *   tome/merge.py:355-369  merge_wavg                -> tome_merge(mode = TOME_MODE_WAVG)
 *   tome/merge.py:372-384  merge_source              -> tome_merge_source
 *   tome/merge.py:54-57    random_* modes (torch.rand scores) -> tome_rowmax
 *
 * Plain C: raw device pointers, sizes and strides; no torch / C++ types.  The caller owns
 * every buffer (inputs, outputs, workspace) and keeps it alive until the stream work is
 * done.  Every call only enqueues work on `stream` (a cudaStream_t passed as void*): no
 * host synchronisation, no allocation, CUDA-graph capturable.  Returns TOME_OK (0) or a
 * negative tome_status; tome_last_error() gives a thread-local message.  There is no CPU
 * fallback: on a device that is not sm_100 every compute entry point fails.
 */
#ifndef TOME_B200_H
#define TOME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOME_ABI_VERSION 26

#if defined(__GNUC__)
#define TOME_API __attribute__((visibility("default")))
#else
#define TOME_API
#endif

typedef enum tome_status {
  TOME_OK = 0,
  TOME_ERR_ARG = -1,          /* bad shape / null pointer / inconsistent plan            */
  TOME_ERR_DTYPE = -2,        /* unsupported element type                                */
  TOME_ERR_ALIGN = -3,        /* pointer or stride alignment the kernels cannot take     */
  TOME_ERR_WORKSPACE = -4,    /* workspace too small (see tome_*_workspace_bytes)        */
  TOME_ERR_CUDA = -5,         /* CUDA runtime / driver error (message has the string)    */
  TOME_ERR_ARCH = -6,         /* device is not compute capability 10.x                   */
  TOME_ERR_UNSUPPORTED = -7   /* valid request this build cannot serve                   */
} tome_status;

/* Largest matching batch (clips, or clips * frames) one call takes: it rides on gridDim.y / gridDim.z. */
#define TOME_MAX_BATCH 65535

typedef enum tome_dtype { TOME_F32 = 0, TOME_BF16 = 1, TOME_U8 = 2 /* tome_patchify input only */ } tome_dtype;

/* tome_match algorithm.  Both produce the SAME canonical node_max/node_idx bits. */
typedef enum tome_match_algo {
  TOME_MATCH_AUTO = 0,
  TOME_MATCH_EXACT_SIMT = 1,  /* fp64 CUDA-core tiles; any shape                          */
  TOME_MATCH_TCGEN05 = 2      /* bf16 h.h+h.m+m.h tcgen05/TMEM pass + exact fp64 refine    */
} tome_match_algo;

/* tome_merge reduction (merge.py:75 `mode`, plus the fused / drop forms). */
typedef enum tome_merge_mode {
  TOME_MODE_WAVG = 0,  /* merge_wavg: sum(x*size)/sum(size), also writes size / log(size) */
  TOME_MODE_SUM = 1,   /* scatter_reduce 'sum'                                            */
  TOME_MODE_MEAN = 2,  /* scatter_reduce 'mean', include_self=True                        */
  TOME_MODE_AMAX = 3,  /* scatter_reduce 'max' / 'amax' (NaN propagating)                 */
  TOME_MODE_DROP = 4   /* bipartite_soft_matching_drop: src tokens discarded              */
} tome_merge_mode;

/*
 * A matching plan: what the reference's merge()/unmerge() closures capture
 * (merge.py:75: unm_idx, src_idx, dst_idx, r) plus two derived maps the kernels use.
 * na = ceil(n/2) "A" (even) tokens, nb = floor(n/2) "B" (odd) tokens.  All int32, device.
 */
typedef struct tome_plan {
  int32_t bm;             /* matching batch (clips, or clips*frames)                      */
  int32_t n;              /* tokens per batch element before the merge                    */
  int32_t r;              /* EFFECTIVE r = min(r, (n - protected)/2) > 0  (merge.py:44)   */
  int32_t class_token;    /* merge.py:59-60, 71-73                                        */
  int32_t distill_token;  /* merge.py:61-62, 82-83                                        */
  const float* node_max;  /* (bm, na)   best score of each A token          (merge.py:64) */
  const int32_t* node_idx;/* (bm, na)   B index attaining it, lowest on ties              */
  int32_t* src_idx;       /* (bm, r)    A tokens merged away, best first    (merge.py:68) */
  int32_t* unm_idx;       /* (bm, na-r) A tokens kept (rank order; ascending if cls)      */
  int32_t* dst_idx;       /* (bm, r)    B token each src merges into        (merge.py:69) */
  int32_t* a_map;         /* (bm, na)   >=0: position in unm_idx; <0: -(dst+1)            */
  int32_t* b_off;         /* (bm, nb+1) CSR offsets into b_src                            */
  int32_t* b_src;         /* (bm, r)    A tokens grouped by dst, ascending k inside       */
  int32_t* b_head;        /* (bm, nb, 4) per B token {count, source 0, source 1, CSR begin}: one
                                        16-byte lookup names up to two merged A tokens      */
} tome_plan;

/* Tensor addressing: element (b, t, c) of a (bm, tokens, c) tensor lives at
 *   base + (b / inner) * stride_bo + (b % inner) * stride_bi + t * stride_n + c
 * (strides in ELEMENTS, channel stride 1).  inner = 1 is a plain batched tensor; inner = T
 * expresses TimeSformer/Motionformer's 'b (p t) m -> (b t) p m' view
 * (tome/patch/timesformer.py:89-90, motionformer.py:150-151) without a copy. */
typedef struct tome_view {
  int64_t stride_bo;
  int64_t stride_bi;
  int64_t stride_n;
  int32_t inner;
} tome_view;

TOME_API int tome_abi_version(void);
TOME_API const char* tome_last_error(void);
/* Number of kernels this library has enqueued since load (bench.py's gpu_launches). */
TOME_API unsigned long long tome_launch_count(void);

/* TOME_OK when `device` is compute capability 10.x and the kernels can launch. */
TOME_API int tome_device_check(int device);

/* --- kernel 1: match (merge.py:49-64) ------------------------------------------------
 * metric: (bm, n, cm) of `dtype`, addressed by `view`.  Writes node_max (bm, na) fp32 and
 * node_idx (bm, na) int32.  Canonical arithmetic (DESIGN.md "score definition"):
 *   norm = fp32(sqrt(sum fp64(x)^2)); mhat = fp32(x / norm);
 *   score = fp32(sum fp64(mhat_a) * fp64(mhat_b)); class token -> row 0 = -inf;
 *   distill token -> column 0 = -inf; argmax = lowest column on ties.
 * The (bm, na, nb) score matrix is never written to memory. */
TOME_API size_t tome_match_workspace_bytes(int32_t bm, int32_t n, int32_t cm, int32_t algo);
TOME_API int tome_match(const void* metric, int32_t dtype, int32_t bm, int32_t n, int32_t cm,
               const tome_view* view, int32_t class_token, int32_t distill_token, int32_t algo,
               float* node_max, int32_t* node_idx, void* workspace, size_t workspace_bytes,
               void* stream);

/* tome_match on metric = mean over heads of K, without materialising the mean
 * (tome/patch/videomae.py:72-73, timesformer.py:83, motionformer.py:143-144, vivit.py:123-124):
 * keys element (b, h, t, k) lives at  base + view(b, t) + h * stride_h + k.  The mean is rounded to
 * `dtype` before normalisation, as the reference's k.mean(1) tensor is.  Workspace as tome_match
 * with TOME_MATCH_TCGEN05. */
TOME_API int tome_match_heads(const void* keys, int32_t dtype, int32_t bm, int32_t heads, int32_t n, int32_t cm,
                     const tome_view* view, int64_t stride_h, int32_t class_token, int32_t distill_token,
                     float* node_max, int32_t* node_idx, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Row max/argmax of a MATERIALISED (bm, na, nb) fp32 score tensor, same masking and tie
 * rule.  Serves the random_merge / random_drop modes, whose scores are torch.rand
 * (merge.py:54-57, 235-238). */
TOME_API int tome_rowmax(const float* scores, int32_t bm, int32_t na, int32_t nb, int32_t class_token,
                int32_t distill_token, float* node_max, int32_t* node_idx, void* stream);

/* --- kernel 2: select (merge.py:65-73) -----------------------------------------------
 * Stable descending rank of node_max (ties: lower A index first; NaN above +inf), then
 * src = first r, unm = rest (ascending when class_token), dst = node_idx[src]; also fills
 * a_map / b_off / b_src / b_head.  plan->r must already be the effective r (> 0). */
TOME_API size_t tome_select_workspace_bytes(int32_t bm, int32_t n);
TOME_API int tome_select(const tome_plan* plan, void* workspace, size_t workspace_bytes, void* stream);

/* --- kernels 1 + 2 in one call (merge.py:49-73: everything bipartite_soft_matching computes) ---------
 * Matching on `metric` (heads == 1: (bm, n, cm) through `view`; heads > 1: the head-mean of K as in
 * tome_match_heads) followed by the selection: one ABI call, one workspace.  With algo == TOME_MATCH_AUTO,
 * cm <= 64 (cm % 8 == 0), 16-byte aligned rows and at most 2048 A tokens this is ONE kernel launch: a thread-block
 * cluster per batch element (csrc/plan_cluster.cu); other shapes, or an explicit algo, run the multi-launch chain.
 * Both give the same bits.  bm, n, r, class/distill flags
 * and every output buffer come from `plan`; plan->node_max / node_idx are OUTPUTS of this call. */
TOME_API size_t tome_plan_build_workspace_bytes(int32_t bm, int32_t n, int32_t cm);
TOME_API int tome_plan_build(const void* metric, int32_t dtype, int32_t heads, int64_t stride_h, const tome_view* view,
                    int32_t cm, int32_t algo, const tome_plan* plan, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Diagnostics for tests: geometry of the one-launch cluster form of tome_plan_build for a shape -- out5 = {CTAs per
 * cluster (0: the shape takes the multi-launch chain), A rows per CTA, B rows per CTA, tile width, shared memory bytes}. */
TOME_API void tome_plan_cluster_describe(int32_t bm, int32_t n, int64_t* out5);

/* Diagnostics for tests: geometry of the tensor-core pass for a shape -- out5 = {column tiles, BN, byte
 * offset of tile_max in the workspace, byte offset of tile_cnt, 1 if the exact refine is fused}. */
TOME_API void tome_match_tc_describe(int32_t bm, int32_t n, int32_t cm, int64_t* out5);

/* --- kernel 3: merge (merge.py:75-85, 260-269, 316-334, 355-369) ----------------------
 * x: (bm, n, c) -> out: (bm, n - r, c), same dtype.  Output token order is the
 * reference's: kept A tokens in unm_idx order, then ALL B tokens in original order
 * (distill: [unm0, dst0, unm1.., dst1..]).  Per B token the reduction order is the
 * reference CPU order (itself, then its sources by ascending k), fp32, no FMA
 * contraction, so fp32 results are bit-identical to the CPU reference.
 *   size_in : (bm, n) fp32 token sizes or NULL (= ones)          [WAVG only]
 *   size_out, logsize_out : (bm, n - r) fp32 or NULL             [WAVG, DROP: ones / zeros]
 *   hybrid_threshold : NaN = off; else merge.py:326 -- a B token hit by any edge with
 *                      node_max < threshold is zeroed before its sources are added. */
TOME_API int tome_merge(const tome_plan* plan, const void* x, int32_t dtype, int32_t c,
               const tome_view* x_view, const float* size_in, int32_t mode,
               float hybrid_threshold, void* out, const tome_view* out_view, float* size_out,
               float* logsize_out, void* stream);

/* tome_merge plus the LayerNorm the patched block applies next (norm2, tome/patch/videomae.py:21-22,
 * timesformer.py:56, motionformer.py:29, vivit.py:39): also writes normed_out = LayerNorm(out) * w + b
 * (same dtype; fp32 statistics over the rounded row), saving the separate LayerNorm pass over x'.
 * ln_weight / ln_bias: (c) in x's dtype (bias may be NULL). */
TOME_API int tome_merge_norm(const tome_plan* plan, const void* x, int32_t dtype, int32_t c,
                    const tome_view* x_view, const float* size_in, int32_t mode, float hybrid_threshold,
                    void* out, const tome_view* out_view, float* size_out, float* logsize_out,
                    const void* ln_weight, const void* ln_bias, float ln_eps, void* normed_out,
                    const tome_view* normed_view, void* stream);

/* tome_merge_norm whose input rows are the sum of two tensors: every row read is round(x + residual)
 * (`residual` laid out like x), i.e. the block's `x = x + attn(...)` (tome/patch/videomae.py:19-20) is taken
 * inside the merge instead of in a pass of its own; results are bit-identical to merging the
 * materialised sum.  residual == NULL: as tome_merge_norm.  ln_weight / normed_out may both be NULL
 * (no fused LayerNorm).  16-byte aligned rows only. */
TOME_API int tome_merge_add_norm(const tome_plan* plan, const void* x, const void* residual, int32_t dtype, int32_t c,
                        const tome_view* x_view, const float* size_in, int32_t mode, float hybrid_threshold,
                        void* out, const tome_view* out_view, float* size_out, float* logsize_out,
                        const void* ln_weight, const void* ln_bias, float ln_eps, void* normed_out,
                        const tome_view* normed_view, void* stream);

/* tome_merge_add_norm with the residual addressed through its OWN view.  TimeSformer / Motionformer hold the
 * residual stream as 'b (p t) m' behind a class token (x_view.inner = T) while the spatial attention's output is
 * '(b t) (1 + p) m' (tome/patch/timesformer.py:44-48, motionformer.py:24-27): the reference rearranges and cats it
 * back before adding; here the add happens inside the merge, each tensor read in place.  residual_view == NULL:
 * laid out like x. */
TOME_API int tome_merge_add_norm_rv(const tome_plan* plan, const void* x, const void* residual, const tome_view* residual_view,
                           int32_t dtype, int32_t c, const tome_view* x_view, const float* size_in, int32_t mode,
                           float hybrid_threshold, void* out, const tome_view* out_view, float* size_out, float* logsize_out,
                           const void* ln_weight, const void* ln_bias, float ln_eps, void* normed_out,
                           const tome_view* normed_view, void* stream);

/* merge_source (merge.py:372-384): source (bm, n, n0) fp32 0/1 adjacency, 'max' reduce.
 * source == NULL means the implicit identity (n0 == n), generated on the fly. */
TOME_API int tome_merge_source(const tome_plan* plan, const float* source, int32_t n0,
                      float hybrid_threshold, float* out, void* stream);

/* Caller-side fusion: sum_out = a + b, normed_out = LayerNorm(sum_out) * w + bias for
 * contiguous (rows, c) tensors -- the residual add that closes a patched block plus the LayerNorm that
 * opens the next (tome/patch/videomae.py:17-22), one pass instead of two kernels. */
TOME_API int tome_add_layernorm(const void* a, const void* b, int32_t dtype, int64_t rows, int32_t c,
                       const void* ln_weight, const void* ln_bias, float ln_eps, void* sum_out,
                       void* normed_out, void* stream);
/* Same with `b` broadcast over the batch: b has b_rows rows and row i of `a` takes b[i % b_rows] -- the
 * position-embedding add in front of the first block (videomae builder:276-278). */
TOME_API int tome_add_rows_layernorm(const void* a, const void* b, int64_t b_rows, int32_t dtype, int64_t rows, int32_t c,
                            const void* ln_weight, const void* ln_bias, float ln_eps, void* sum_out,
                            void* normed_out, void* stream);

/* The same add + LayerNorm through (b, p, t) row views, for the divided space-time blocks (SURVEY.md 8f-f2;
 * tome/patch/timesformer.py:38-56, slowfast/models/timesformer.py:115-153): row (b, p, t), b < nb, p < np, t < nt, of
 * each tensor lives at base + b*strides[0] + p*strides[1] + t*strides[2] (elements), so the hops between the
 * '(b p) t' (temporal attention), '(b t) (1 + p)' (spatial attention) and 'b (1 + p t)' (residual stream) row orders
 * are addressing instead of rearrange / cat copies.  sum_out = a + b (b == NULL: a) and / or
 * normed_out = LayerNorm(sum) * w + bias; either output may be NULL. */
TOME_API int tome_rows_add_layernorm(const void* a, const int64_t* a_strides, const void* b, const int64_t* b_strides,
                            int32_t dtype, int32_t nb, int32_t np, int32_t nt, int32_t c, const void* ln_weight,
                            const void* ln_bias, float ln_eps, void* sum_out, const int64_t* sum_strides,
                            void* normed_out, const int64_t* normed_strides, void* stream);

/* The class-token rows of the divided space-time blocks (tome/patch/timesformer.py:41-48, 56), one small launch instead of
 * ~10 on `batch` rows:  sum = a[b] (+ mean over t < mean_t of mean_src[b, t]) (+ add[b]), each step rounded to `dtype` like the
 * separate ops; sum_out[b] = sum (optional); normed_out[b, rep] = LayerNorm(sum) * w + bias for rep < reps (the per-frame
 * copies in front of the spatial attention; optional).  Strides in elements. */
TOME_API int tome_cls_rows(const void* a, int64_t a_stride_b, const void* add, int64_t add_stride_b, const void* mean_src,
                  int64_t mean_stride_b, int64_t mean_stride_t, int32_t mean_t, int32_t dtype, int32_t batch, int32_t c,
                  void* sum_out, int64_t sum_stride_b, const void* ln_weight, const void* ln_bias, float ln_eps,
                  void* normed_out, int64_t normed_stride_b, int64_t normed_stride_rep, int32_t reps, void* stream);

/* Caller-side piece of proportional attention (tome/patch/videomae.py:62-63,
 * vivit.py:103-104, timesformer.py:72-74: attn + log(size) of the key token).  q and k heads carry spare
 * channels d, d+1 (host-padded); this writes k[b, t, h, d..d+1] = two-term split of log_size / scale
 * (0 for the `lead` leading class tokens) so that  scale * (q . k)  over the padded head equals
 * scale * (q . k) + log(size_key)  and the attention kernel needs no mask.  When q != NULL its channels
 * d, d+1 are set to 1 (0 for the leading tokens, whose logits take no bias); with q == NULL they must
 * already hold 1.  log_size: (b, n - lead) fp32; element strides in units of `dtype`. */
TOME_API int tome_attn_key_bias(const float* log_size, int32_t b, int32_t n, int32_t lead, int32_t heads, int32_t d,
                       float scale, int32_t dtype, void* k, int64_t k_stride_b, int64_t k_stride_n,
                       int64_t k_stride_h, void* q, int64_t q_stride_b, int64_t q_stride_n, int64_t q_stride_h,
                       void* stream);

/* Caller-side attention over SHORT sequences: softmax(scale * q k^T) v for `seqs` independent
 * sequences of n_tok <= 32 tokens and `heads` heads of dimension d = 64 -- TimeSformer's temporal attention on
 * '(b p) t' rows (slowfast/models/timesformer.py:118-123), 18 816 problems of 8 x 8 per call at the bench shape,
 * which tile-based flash kernels run at a few percent of HBM speed.  q / k / v element (s, t, h, c) lives at
 * base + s*seq_stride + t*tok_stride + h*d + c (the QKV GEMM's output is read in place); out (seqs, n_tok, heads, d)
 * contiguous, same dtype (fp32 / bf16); fp32 scores, softmax and accumulation. */
TOME_API int tome_attn_short(const void* q, const void* k, const void* v, int32_t dtype, int64_t seqs, int32_t n_tok,
                    int32_t heads, int32_t d, int64_t seq_stride, int64_t tok_stride, float scale, void* out, void* stream);

/* Caller-side attention of Motionformer's trajectory attention (tome/patch/motionformer.py:105-115,
 * slowfast/models/motionformer_vit_helper.py:196-243), bf16, head dimension 64.
 * tome_frames_attention -- the space stage with the proportional-attention key bias: every one of the S = frames *
 *   keys_per_frame patch queries attends to the keys of EACH frame separately,
 *     xs[b, s, f, h*64 + c] = sum_p softmax_p(scale * q[b,h,s] . k[b,h,f,p] + key_bias[b, f*P + p]) * v[b,h,f,p][c],
 *   on tcgen05 / TMEM / TMA.  qkv: the QKV GEMM's output (b, n, 3 * heads * 64), n = lead + S tokens (`lead` leading
 *   tokens that are neither queries nor keys -- Motionformer's class token -- then the tokens in '(f p)' order), channel
 *   order (3, heads, 64), read in place; key_bias (b, S) fp32 or NULL; xs (b, S, frames, heads*64);
 *   x_diag (b, S, heads*64) = xs[b, s, frame of s] or NULL.  keys_per_frame <= 256.  With frames == 1 and lead == 0
 *   this is plain attention over up to 256 tokens with a key bias -- TimeSformer's spatial attention
 *   (tome/patch/timesformer.py:70-79), whose class QUERY takes no bias: unbiased_queries = 1.
 * tome_traj_temporal -- the temporal stage: out[r, h] = sum_f softmax_f(scale * q2[r,h] . k2[r,f,h]) * vals[r,f,h] for
 *   rows r = (b, s); q2 / out (rows, heads*64), k2 / vals (rows, frames, heads*64), frames <= 32; bf16 or fp32 tensors. */
TOME_API int tome_frames_attention(const void* qkv, int32_t dtype, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t frames,
                          int32_t keys_per_frame, int32_t lead, int32_t unbiased_queries, float scale, const float* key_bias,
                          void* xs, void* x_diag, void* stream);
TOME_API int tome_traj_temporal(const void* q2, const void* k2, const void* vals, int32_t dtype, int64_t rows, int32_t frames,
                       int32_t heads, int32_t d, float scale, void* out, void* stream);

/* Caller-side data format: the tubelet embedding of the four models is a Conv3d whose
 * kernel equals its stride (slowfast/models/videomae_video_model_builder.py:138-160), i.e. a GEMM over
 * non-overlapping tubelets.  x (b, c, t, h, w) contiguous -> out (b, (t/tt)(h/ph)(w/pw), c*tt*ph*pw), token
 * order (t', h', w'), feature order (c, tt, ph, pw) = the flattened conv weight's; in_dtype -> out_dtype
 * conversion (fp32 clips to a bf16 model) happens in the same pass.  in_dtype TOME_U8: decoder-style uint8
 * frames, converted as value / 255 (what `x.float() / 255` gives), so a clip crosses PCIe at one byte per
 * sample.  pw % 8 == 0. */
TOME_API int tome_patchify(const void* x, int32_t in_dtype, int32_t b, int32_t c, int32_t t, int32_t h, int32_t w,
                  int32_t tt, int32_t ph, int32_t pw, void* out, int32_t out_dtype, void* stream);

/* Caller-side fusion: the first half of the block's MLP right after the merge
 * (videomae builder:40-56: fc1 -> nn.GELU -> fc2), out = GELU_erf(x @ W^T + bias) from ONE tcgen05 GEMM whose
 * epilogue applies bias and activation, instead of a library GEMM plus an elementwise pass over the (m, n)
 * tensor.  bf16 only: x (m, k) with rows `x_row_stride` elements apart, W (n, k) and out (m, n) contiguous,
 * bias (n) or NULL; fp32 accumulation; the pre-activation is rounded to bf16 before the GELU, as the two
 * separate ops would.  gelu == 0: bias only; 1: erf GELU (nn.GELU); 2: HuggingFace "gelu_fast"
 * (0.5 x (1 + tanh(0.7978845608 x (1 + 0.044715 x^2))), ViViT's hidden_act).  n % 256 == 0, k % 8 == 0. */
TOME_API int tome_linear_gelu(const void* x, const void* w, const void* bias, int32_t m, int32_t n, int32_t k,
                     int64_t x_row_stride, int32_t gelu, void* out, void* stream);

/* Caller-side fp32 linear layers on tcgen05 at fp32 accuracy (the reference benchmark runs fp32 with
 * TF32 off, slowfast/utils/model_benchmark.py:21-45).  An fp32 value is exactly h + m + l with three bf16 terms and a
 * bf16 x bf16 product is exact in fp32, so accumulating the nine plane products in fp32 (terms = 9) is an fp32 GEMM
 * without TF32 truncation; terms = 8 leaves out l.l (<= 2^-32 of |a||b| per product, 2^-8 of the accumulator's own
 * rounding unit: the host mirror's default), terms = 6 also m.l and l.m (<= 2^-23 relative per product).
 * tome_split3: x (rows, k) fp32, rows `row_stride` elements apart -> out (rows, 3k) bf16 planes [h | m | l]; k % 4 == 0.
 * tome_linear_f32: out (m, n) fp32 = act(x @ W^T + bias) from the split planes x3 (m, 3k), w3 (n, 3k); bias (n) fp32 or
 *   NULL; gelu 0 / 1 (erf GELU, exact erf) / 2 (HuggingFace "gelu_fast", tanhf).  n % 256 == 0, k % 32 == 0.  out_planes (m, 3n) bf16: the result ALSO (or, with
 *   out == NULL, ONLY) as split planes -- the operand of the next tome_linear_f32 / tome_attention_f32 without an fp32
 *   round trip (fc1 -> fc2, qkv -> attention). */
TOME_API int tome_split3(const void* x, int64_t rows, int32_t k, int64_t row_stride, void* out, void* stream);
/* The inverse on a column slice: out (rows, ncols) fp32 = h + m + l (exact) of columns col0 .. col0 + ncols - 1 of a planes
 * tensor x3 (rows, 3n); multiples of 8.  (The K third of a QKV result held as planes, for the matching metric.) */
TOME_API int tome_planes_sum(const void* x3, int64_t rows, int32_t n, int32_t col0, int32_t ncols, void* out, void* stream);
TOME_API int tome_linear_f32(const void* x3, const void* w3, const void* bias, int32_t m, int32_t n, int32_t k, int32_t gelu,
                    int32_t terms, void* out, void* out_planes, void* stream);

/* Caller-side fp32 attention on tcgen05 at fp32 accuracy (tome/patch/videomae.py:58-68, vivit.py:98-117
 * in the fp32 models): out (b, n, heads*64) fp32 = softmax(scale * q k^T + key_bias) v per head, flash-style (running
 * maximum over 64-key blocks), from qkv3 = tome_split3 of the QKV GEMM's output viewed (b*n, 3*heads*64), i.e.
 * (b*n, 9*heads*64) bf16 planes [h: q k v | m: q k v | l: q k v].  Six plane products for q k^T and six for P V with
 * P split exactly in registers.  key_bias (b, n) fp32 or NULL; the first `unbiased_queries` queries
 * take no bias (TimeSformer's class token).  out_planes (b*n, 3*heads*64) bf16: the context also / only as split planes
 * (the output projection's operand). */
TOME_API int tome_attention_f32(const void* qkv3, int32_t b, int32_t n, int32_t heads, int32_t d, float scale,
                       const float* key_bias, int32_t unbiased_queries, void* out, void* out_planes, void* stream);

/* One query token against the whole sequence (Motionformer's class token, slowfast vit_helper.py:181-189 `cls_out`): out (b,
 * heads*64) = softmax(scale * q[b, query_token] . k[b, :]) v[b, :] per head from the QKV GEMM's contiguous (b, n, 3*heads*64)
 * output, bf16 or fp32 (fp32 arithmetic either way). */
TOME_API int tome_cls_attention(const void* qkv, int32_t dtype, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t query_token,
                       float scale, void* out, void* stream);

/* tome_frames_attention in fp32 accuracy (the fp32 Motionformer: the reference benchmark's arithmetic): the same per-frame
 * attention on the exact-split kernel of tome_attention_f32 (which is its frames == 1, lead == 0 case), from qkv3 =
 * tome_split3 of the QKV GEMM's output viewed (b*n, 3*heads*64).  Any keys_per_frame; xs (b, S, frames, heads*64) fp32 and /
 * or xs_planes (b*S*frames, 3*heads*64) bf16 planes (the operand of the K projection that follows); x_diag (b, S, heads*64)
 * fp32 or NULL. */
TOME_API int tome_frames_attention_f32(const void* qkv3, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t frames,
                              int32_t keys_per_frame, int32_t lead, float scale, const float* key_bias, void* xs,
                              void* xs_planes, void* x_diag, void* stream);

/* Caller-side bf16 attention with the proportional-attention key bias for sequences of any length (SURVEY.md 8f-f1;
 * tome/patch/videomae.py:58-68, vivit.py:98-117: `attn + size.log()`): out (b, n, heads*64) bf16 = softmax(scale * q k^T +
 * key_bias) v per head, flash-style (running maximum over 128-key blocks, fp32 softmax and accumulation), q / k / v read in
 * place from the QKV GEMM's contiguous (b, n, 3*heads*64) bf16 output.  key_bias (b, n) fp32 or NULL; the first
 * `unbiased_queries` queries take no bias. */
TOME_API int tome_attention_bf16(const void* qkv, int32_t b, int32_t n, int32_t heads, int32_t d, float scale,
                        const float* key_bias, int32_t unbiased_queries, void* out, void* stream);

/* unmerge (merge.py:87-100): x (bm, n - r, c) -> out (bm, n, c); contiguous tensors. */
TOME_API int tome_unmerge(const tome_plan* plan, const void* x, int32_t dtype, int32_t c, void* out,
                 void* stream);

/* --- matching between two arbitrary token sets --------------------------------------------------------
 * The upstream-ToMe variants the reference keeps: kth_bipartite_soft_matching (tome/merge.py:105-158: every
 * k-th token is a destination, the rest are sources) and random_bipartite_soft_matching (merge.py:161-212: r
 * random sources).  tok_of_row (int32, batch element b at tok_of_row + b * tok_stride_b; stride 0 = shared)
 * lists token indices: rows [0, ra) the A / source set, rows [ra, ra + nb) the B / destination set.
 * tome_match_sets: canonical scores (as tome_match) of every A row against every B row; node_max / node_idx
 *   (bm, ra): best B ROW per A row, lowest on ties (merge.py:135 / 190 `scores.max(dim=-1)`).
 * tome_group_reduce: out (bm, nb, c) contiguous = every B row reduced with the A rows assigned to it by
 *   dst_idx (bm, ra), include_self, in the reference CPU order (itself, then ascending k), fp32 arithmetic
 *   -- `dst.scatter_reduce(-2, dst_idx, src, reduce=mode)`, merge.py:141 / 196.  mode: SUM, MEAN or AMAX.
 * tome_gather_rows: out[b, t] = x[b, map[b, t]] (zeros where map < 0) -- the unmerge of both variants
 *   (merge.py:145-156, 200-210) once the caller has composed the row map. */
TOME_API size_t tome_match_sets_workspace_bytes(int32_t bm, int32_t ra, int32_t nb, int32_t cm);
TOME_API int tome_match_sets(const void* metric, int32_t dtype, int32_t bm, int32_t n, int32_t cm, const tome_view* view,
                    const int32_t* tok_of_row, int64_t tok_stride_b, int32_t ra, int32_t nb, float* node_max,
                    int32_t* node_idx, void* workspace, size_t workspace_bytes, void* stream);
TOME_API int tome_group_reduce(const void* x, int32_t dtype, int32_t bm, int32_t n, int32_t c, const tome_view* x_view,
                      const int32_t* tok_of_row, int64_t tok_stride_b, int32_t ra, int32_t nb, const int32_t* dst_idx,
                      int32_t mode, void* out, void* stream);
TOME_API int tome_gather_rows(const void* x, int32_t dtype, int32_t bm, int32_t n_in, int32_t c, const int32_t* map,
                     int32_t n_out, void* out, void* stream);

/* --- `source` consumers and random modes ------------------------------------------
 * Compact form of the reference's dense source matrix (tome/merge.py:372-384; tome/vis.py:55,102,146 read it as
 * `source.argmax(dim=1)`): group (bm, n0) int32, group[b, t] = index of the merged token that holds original
 * token t, -1 once the token is gone (drop modes, merge.py:260-269; destinations zeroed by a hybrid threshold,
 * merge.py:326).  dense[b, s, t] == 1 exactly when group[b, t] == s.
 * tome_source_compose: one block's update.  group_in == NULL is the implicit identity (n0 == plan->n).  drop != 0:
 *   merged-away A tokens are discarded instead of joining their destination.  hybrid_threshold: NaN = off.
 * tome_source_dense: the fp32 (bm, n_tokens, n0) matrix the reference API exposes, expanded on demand. */
TOME_API int tome_source_compose(const tome_plan* plan, const int32_t* group_in, int32_t n0, int32_t drop,
                        float hybrid_threshold, int32_t* group_out, void* stream);
TOME_API int tome_source_dense(const int32_t* group, int32_t bm, int32_t n_tokens, int32_t n0, float* out, void* stream);

/* random_merge / random_drop scores (merge.py:54-57, 235-238: torch.rand of shape (bm, na, nb), then max) from a
 * counter-based Philox4x32-10 stream, fused with the masked row max / first argmax; the score tensor is never
 * written unless scores_out (bm, na, nb) is given (tests).  key = seed; counter = (column / 4, A row, clip0 + b,
 * call); score = (word >> 8) * 2^-24.  An edge's score therefore depends only on (seed, call, clip, row, column):
 * not on the batch composition or on how clips are sharded over GPUs.  philox_state: DEVICE memory,
 * {uint64 seed, uint64 call}; advance != 0 bumps `call` on the stream after the draw, so a captured CUDA graph
 * draws fresh scores at every replay. */
TOME_API int tome_random_rowmax(void* philox_state, int64_t clip0, int32_t bm, int32_t na, int32_t nb, int32_t class_token,
                       int32_t distill_token, float* node_max, int32_t* node_idx, float* scores_out, int32_t advance,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOME_B200_H */
