// extern "C" entry points declared in include/tome_b200.h: argument validation, error
// strings, dispatch.  No torch types, no allocation, no host synchronisation.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace tome {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

static int g_dev_ok[64];   // 0 unknown, 1 ok, -1 bad
static int g_sms[64];

int ensure_device_ok() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(TOME_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= 64) return set_error(TOME_ERR_ARG, "device index %d out of range", dev);
  if (g_dev_ok[dev] == 0) {
    int major = 0, sms = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return set_error(TOME_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    g_sms[dev] = sms;
    g_dev_ok[dev] = (major == 10) ? 1 : -1;
  }
  if (g_dev_ok[dev] < 0)
    return set_error(TOME_ERR_ARCH, "device %d is not compute capability 10.x (this library is sm_100a only, no fallback)", dev);
  return TOME_OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64 && g_sms[dev] > 0) ? g_sms[dev] : 148;
}

// implemented in the kernel translation units
size_t match_exact_workspace(int bm, int n, int cm);
int launch_match_exact(const void*, int, int, int, int, const View&, int, int, float*, int*, void*, size_t, cudaStream_t);
int launch_rowmax(const float*, int, int, int, int, int, float*, int*, cudaStream_t);
size_t match_tc_workspace(int bm, int n, int cm);
bool match_tc_supported(int dtype, int bm, int n, int cm, const View& v, const void* metric);
int launch_match_tc(const void*, int, int, int, int, const View&, int, int, float*, int*, void*, size_t, cudaStream_t, int, long long,
                    unsigned long long**, unsigned long long**);
int launch_select_one(const tome_plan*, const unsigned long long*, unsigned long long*, cudaStream_t);
bool match_tc_fused_refine(int bm, int n, int cm);
bool select_two_launches();
void match_tc_describe(int, int, int, long long[5]);
size_t select_workspace(int bm, int n);
int launch_select(const tome_plan*, void*, size_t, cudaStream_t);
int launch_merge(const tome_plan*, const void*, int, int, const View&, const float*, int, float, void*, const View&,
                 float*, float*, cudaStream_t, const void*, const void*, float, void*, const View*, const void*, const View*);
int launch_rows_add_layernorm(const void*, const long long*, const void*, const long long*, int, int, int, int, int, const void*, const void*,
                              float, void*, const long long*, void*, const long long*, cudaStream_t);
int launch_merge_source(const tome_plan*, const float*, int, float, float*, cudaStream_t);
int launch_unmerge(const tome_plan*, const void*, int, int, void*, cudaStream_t);
int launch_add_layernorm(const void*, const void*, long long, int, long long, int, const void*, const void*, float, void*, void*, cudaStream_t);
int launch_linear_gelu(const void*, const void*, const void*, int, int, int, long long, int, void*, cudaStream_t);
int launch_patchify(const void*, int, int, int, int, int, int, int, int, int, void*, int, cudaStream_t);
int launch_key_bias(const float*, int, int, int, int, int, float, int, void*, long long, long long, long long, void*, long long,
                    long long, long long, cudaStream_t);

size_t plan_cluster_workspace(int bm, int n);
bool plan_cluster_supported(int dtype, int bm, int n, int cm, int heads, const View& v, long long stride_h, const void* metric);
int launch_plan_cluster(const void*, int, int, long long, const View&, int, const tome_plan*, void*, size_t, cudaStream_t);
void plan_cluster_describe(int, int, long long[5]);
size_t match_sets_workspace(int bm, int rows, int ra, int cm);
int launch_match_sets(const void*, int, int, int, const View&, const int*, long long, int, int, float*, int*, void*, size_t, cudaStream_t);
int launch_group_reduce(const void*, int, int, int, const View&, const int*, long long, int, int, const int*, int, void*, cudaStream_t);
int launch_gather_rows(const void*, int, int, int, int, const int*, int, void*, cudaStream_t);
int launch_attn_short(const void*, const void*, const void*, int, long long, int, int, int, long long, long long, float, void*, cudaStream_t);
int launch_cls_attention(const void*, int, int, int, int, int, float, void*, cudaStream_t);
int launch_frames_attention(const void*, int, int, int, int, int, float, const float*, void*, void*, int, int, cudaStream_t);
int launch_traj_temporal(const void*, const void*, const void*, int, long long, int, int, float, void*, cudaStream_t);
int launch_frames_attention_f32(const void*, int, int, int, int, int, int, float, const float*, int, void*, void*, void*, cudaStream_t);
int launch_split3(const void*, long long, int, long long, void*, cudaStream_t);
int launch_planes_sum(const void*, long long, int, int, int, void*, cudaStream_t);
int launch_linear_f32(const void*, const void*, const void*, int, int, int, int, int, void*, void*, cudaStream_t);
int launch_attention_f32(const void*, int, int, int, float, const float*, int, void*, void*, cudaStream_t);
int launch_attention_bf16(const void*, int, int, int, float, const float*, int, void*, cudaStream_t);
struct ClsArgs {
  const void *a, *add, *mean_src;
  long long a_sb, add_sb, m_sb, m_st;
  int mean_t;
  void *sum_out, *normed_out;
  long long s_sb, n_sb, n_sr;
  int reps, c;
  const void *ln_w, *ln_b;
  float eps;
};
int launch_cls_rows(const ClsArgs&, int, int, cudaStream_t);
int launch_source_compose(const tome_plan*, const int*, int, int, float, int*, cudaStream_t);
int launch_source_dense(const int*, int, int, int, float*, cudaStream_t);
int launch_random_rowmax(void*, long long, int, int, int, int, int, float*, int*, float*, int, cudaStream_t);

static int check_plan(const tome_plan* p, const char* who) {
  if (!p) return set_error(TOME_ERR_ARG, "%s: plan is NULL", who);
  if (p->bm <= 0 || p->n < 2) return set_error(TOME_ERR_ARG, "%s: bad plan shape bm=%d n=%d", who, p->bm, p->n);
  if (p->bm > TOME_MAX_BATCH)       // the matching batch rides on gridDim.y / gridDim.z of several kernels
    return set_error(TOME_ERR_UNSUPPORTED, "%s: matching batch %d > %d; split the batch", who, p->bm, TOME_MAX_BATCH);
  const int prot = (p->class_token ? 1 : 0) + (p->distill_token ? 1 : 0);
  if (p->r <= 0 || p->r > (p->n - prot) / 2)
    return set_error(TOME_ERR_ARG, "%s: plan->r=%d is not an effective r for n=%d (max %d)", who, p->r, p->n, (p->n - prot) / 2);
  const bool unm_ok = p->unm_idx || (p->n + 1) / 2 == p->r;   // nothing is kept when r == na
  if (!p->node_max || !p->node_idx || !p->src_idx || !unm_ok || !p->dst_idx || !p->a_map || !p->b_off || !p->b_src || !p->b_head)
    return set_error(TOME_ERR_ARG, "%s: plan has NULL buffers", who);
  return TOME_OK;
}

}  // namespace tome

using namespace tome;

extern "C" {

int tome_abi_version(void) { return TOME_ABI_VERSION; }
const char* tome_last_error(void) { return g_err; }
unsigned long long tome_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int tome_device_check(int device) {
  int cur = 0;
  TOME_CUDA(cudaGetDevice(&cur));
  if (device != cur) TOME_CUDA(cudaSetDevice(device));
  int rc = ensure_device_ok();
  if (device != cur) cudaSetDevice(cur);
  return rc;
}

size_t tome_match_workspace_bytes(int32_t bm, int32_t n, int32_t cm, int32_t algo) {
  if (bm <= 0 || n <= 0 || cm <= 0) return 0;
  size_t e = match_exact_workspace(bm, n, cm), t = match_tc_workspace(bm, n, cm);
  if (algo == TOME_MATCH_EXACT_SIMT) return e;
  if (algo == TOME_MATCH_TCGEN05) return t;
  return e > t ? e : t;
}

int tome_match(const void* metric, int32_t dtype, int32_t bm, int32_t n, int32_t cm, const tome_view* view,
               int32_t class_token, int32_t distill_token, int32_t algo, float* node_max, int32_t* node_idx,
               void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(metric && node_max && node_idx && workspace, "tome_match: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && n >= 2 && cm > 0, "tome_match: bad shape bm=%d n=%d cm=%d", bm, n, cm);
  if (bm > TOME_MAX_BATCH) return set_error(TOME_ERR_UNSUPPORTED, "tome_match: matching batch %d > %d; split the batch", bm, TOME_MAX_BATCH);
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_match: unsupported dtype %d", dtype);
  TOME_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "tome_match: workspace must be 256-byte aligned");
  const View v = make_view(view, n, cm);
  cudaStream_t st = (cudaStream_t)stream;
  if (algo == TOME_MATCH_AUTO)
    algo = match_tc_supported(dtype, bm, n, cm, v, metric) ? TOME_MATCH_TCGEN05 : TOME_MATCH_EXACT_SIMT;
  if (algo == TOME_MATCH_TCGEN05) {
    if (!match_tc_supported(dtype, bm, n, cm, v, metric))
      return set_error(TOME_ERR_UNSUPPORTED, "tome_match: tcgen05 path needs an fp32/bf16 metric with cm %% 8 == 0, cm <= 4096 and even strides (got cm=%d)", cm);
    return launch_match_tc(metric, dtype, bm, n, cm, v, class_token, distill_token, node_max, node_idx, workspace,
                           workspace_bytes, st, 1, 0, nullptr, nullptr);
  }
  if (algo != TOME_MATCH_EXACT_SIMT) return set_error(TOME_ERR_ARG, "tome_match: unknown algo %d", algo);
  return launch_match_exact(metric, dtype, bm, n, cm, v, class_token, distill_token, node_max, node_idx, workspace,
                            workspace_bytes, st);
}

int tome_match_heads(const void* keys, int32_t dtype, int32_t bm, int32_t heads, int32_t n, int32_t cm,
                     const tome_view* view, int64_t stride_h, int32_t class_token, int32_t distill_token,
                     float* node_max, int32_t* node_idx, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(keys && node_max && node_idx && workspace && view, "tome_match_heads: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && n >= 2 && cm > 0 && heads >= 1, "tome_match_heads: bad shape bm=%d heads=%d n=%d cm=%d", bm, heads, n, cm);
  if (bm > TOME_MAX_BATCH) return set_error(TOME_ERR_UNSUPPORTED, "tome_match_heads: matching batch %d > %d; split the batch", bm, TOME_MAX_BATCH);
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_match_heads: unsupported dtype %d", dtype);
  TOME_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "tome_match_heads: workspace must be 256-byte aligned");
  const View v = make_view(view, n, cm);
  if (!match_tc_supported(dtype, bm, n, cm, v, keys))
    return set_error(TOME_ERR_UNSUPPORTED, "tome_match_heads: needs cm %% 8 == 0 and even strides (got cm=%d); average the heads and call tome_match", cm);
  return launch_match_tc(keys, dtype, bm, n, cm, v, class_token, distill_token, node_max, node_idx, workspace,
                         workspace_bytes, (cudaStream_t)stream, heads, stride_h, nullptr, nullptr);
}

static size_t align256_(size_t x) { return (x + 255) & ~(size_t)255; }

size_t tome_plan_build_workspace_bytes(int32_t bm, int32_t n, int32_t cm) {
  if (bm <= 0 || n <= 0 || cm <= 0) return 0;
  const size_t chain = align256_(tome_match_workspace_bytes(bm, n, cm, TOME_MATCH_AUTO)) + align256_(select_workspace(bm, n));
  const size_t cluster = align256_(plan_cluster_workspace(bm, n));
  return chain > cluster ? chain : cluster;
}

void tome_plan_cluster_describe(int32_t bm, int32_t n, int64_t* out5) {
  long long o[5];
  plan_cluster_describe(bm, n, o);
  for (int i = 0; i < 5; ++i) out5[i] = o[i];
}

void tome_match_tc_describe(int32_t bm, int32_t n, int32_t cm, int64_t* out5) {
  long long o[5];
  match_tc_describe(bm, n, cm, o);
  for (int i = 0; i < 5; ++i) out5[i] = o[i];
}

int tome_plan_build(const void* metric, int32_t dtype, int32_t heads, int64_t stride_h, const tome_view* view, int32_t cm,
                    int32_t algo, const tome_plan* plan, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_plan_build");
  if (rc) return rc;
  const int bm = plan->bm, n = plan->n;
  TOME_CHECK_ARG(metric && workspace && cm > 0 && heads >= 1 && n >= 2, "tome_plan_build: NULL pointer or bad shape (heads=%d n=%d cm=%d)", heads, n, cm);
  TOME_CHECK_ARG(heads == 1 || view, "tome_plan_build: a head-mean metric needs an explicit view");
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_plan_build: unsupported dtype %d", dtype);
  TOME_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "tome_plan_build: workspace must be 256-byte aligned");
  if (workspace_bytes < tome_plan_build_workspace_bytes(bm, n, cm))
    return set_error(TOME_ERR_WORKSPACE, "tome_plan_build: workspace %zu < %zu bytes", workspace_bytes, tome_plan_build_workspace_bytes(bm, n, cm));
  const size_t match_bytes = align256_(tome_match_workspace_bytes(bm, n, cm, TOME_MATCH_AUTO));
  void* select_ws = (char*)workspace + match_bytes;
  const size_t select_bytes = workspace_bytes - match_bytes;
  const View v = make_view(view, n, cm);
  cudaStream_t st = (cudaStream_t)stream;
  float* node_max = const_cast<float*>(plan->node_max);
  int32_t* node_idx = const_cast<int32_t*>(plan->node_idx);
  const bool tc_ok = match_tc_supported(dtype, bm, n, cm, v, metric);
  // one cluster launch for kernels 1 + 2 (plan_cluster.cu) whenever the shape fits; TOME_MATCH_TCGEN05 names the
  // multi-launch tensor-core chain explicitly (kept for wide metrics and as a cross-check)
  if (algo == TOME_MATCH_AUTO && plan_cluster_supported(dtype, bm, n, cm, heads, v, stride_h, metric))
    return launch_plan_cluster(metric, dtype, heads, stride_h, v, cm, plan, workspace, workspace_bytes, st);
  if (algo == TOME_MATCH_AUTO) algo = tc_ok ? TOME_MATCH_TCGEN05 : TOME_MATCH_EXACT_SIMT;
  if (algo == TOME_MATCH_TCGEN05) {
    if (!tc_ok)
      return set_error(TOME_ERR_UNSUPPORTED, "tome_plan_build: tcgen05 path needs cm %% 8 == 0, cm <= 4096 and even strides (got cm=%d)", cm);
    // three launches: normalise -> tensor-core match (packed keys only) -> one-launch select, which decodes the packed keys
    // itself and whose threshold flags the normalisation kernel zeroed; TOME_SELECT_TWO=1 keeps the rank + finish pair
    if (!select_two_launches() && match_tc_fused_refine(bm, n, cm)) {
      unsigned long long *packed = nullptr, *flags = nullptr;
      rc = launch_match_tc(metric, dtype, bm, n, cm, v, plan->class_token, plan->distill_token, nullptr, nullptr, workspace,
                           match_bytes, st, heads, stride_h, &packed, &flags);
      if (rc) return rc;
      return launch_select_one(plan, packed, flags, st);
    }
    rc = launch_match_tc(metric, dtype, bm, n, cm, v, plan->class_token, plan->distill_token, node_max, node_idx, workspace,
                         match_bytes, st, heads, stride_h, nullptr, nullptr);
    if (rc) return rc;
    return launch_select(plan, select_ws, select_bytes, st);
  }
  if (algo != TOME_MATCH_EXACT_SIMT) return set_error(TOME_ERR_ARG, "tome_plan_build: unknown algo %d", algo);
  if (heads != 1) return set_error(TOME_ERR_UNSUPPORTED, "tome_plan_build: the exact SIMT path takes a materialised metric (heads == 1)");
  rc = launch_match_exact(metric, dtype, bm, n, cm, v, plan->class_token, plan->distill_token, node_max, node_idx, workspace,
                          match_bytes, st);
  if (rc) return rc;
  return launch_select(plan, select_ws, select_bytes, st);
}

int tome_rowmax(const float* scores, int32_t bm, int32_t na, int32_t nb, int32_t class_token, int32_t distill_token,
                float* node_max, int32_t* node_idx, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(scores && node_max && node_idx, "tome_rowmax: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && na > 0 && nb > 0, "tome_rowmax: bad shape bm=%d na=%d nb=%d", bm, na, nb);
  if (bm > TOME_MAX_BATCH) return set_error(TOME_ERR_UNSUPPORTED, "tome_rowmax: matching batch %d > %d; split the batch", bm, TOME_MAX_BATCH);
  return launch_rowmax(scores, bm, na, nb, class_token, distill_token, node_max, node_idx, (cudaStream_t)stream);
}

size_t tome_select_workspace_bytes(int32_t bm, int32_t n) { return (bm > 0 && n > 0) ? select_workspace(bm, n) : 0; }

int tome_select(const tome_plan* plan, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_select");
  if (rc) return rc;
  TOME_CHECK_ARG(workspace, "tome_select: workspace is NULL");
  return launch_select(plan, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tome_merge(const tome_plan* plan, const void* x, int32_t dtype, int32_t c, const tome_view* x_view,
               const float* size_in, int32_t mode, float hybrid_threshold, void* out, const tome_view* out_view,
               float* size_out, float* logsize_out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_merge");
  if (rc) return rc;
  TOME_CHECK_ARG(x && out && c > 0, "tome_merge: NULL tensor or c=%d", c);
  TOME_CHECK_ARG(mode >= TOME_MODE_WAVG && mode <= TOME_MODE_DROP, "tome_merge: unknown mode %d", mode);
  TOME_CHECK_ARG(mode == TOME_MODE_WAVG || size_in == nullptr, "tome_merge: size_in is only meaningful for TOME_MODE_WAVG");
  TOME_CHECK_ARG(x != out, "tome_merge: in-place merge is not supported");
  const View xv = make_view(x_view, plan->n, c), ov = make_view(out_view, plan->n - plan->r, c);
  return launch_merge(plan, x, dtype, c, xv, size_in, mode, hybrid_threshold, out, ov, size_out, logsize_out,
                      (cudaStream_t)stream, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, nullptr);
}

int tome_merge_norm(const tome_plan* plan, const void* x, int32_t dtype, int32_t c, const tome_view* x_view,
                    const float* size_in, int32_t mode, float hybrid_threshold, void* out, const tome_view* out_view,
                    float* size_out, float* logsize_out, const void* ln_weight, const void* ln_bias, float ln_eps,
                    void* normed_out, const tome_view* normed_view, void* stream) {
  return tome_merge_add_norm(plan, x, nullptr, dtype, c, x_view, size_in, mode, hybrid_threshold, out, out_view, size_out,
                             logsize_out, ln_weight, ln_bias, ln_eps, normed_out, normed_view, stream);
}

int tome_merge_add_norm(const tome_plan* plan, const void* x, const void* residual, int32_t dtype, int32_t c,
                        const tome_view* x_view, const float* size_in, int32_t mode, float hybrid_threshold, void* out,
                        const tome_view* out_view, float* size_out, float* logsize_out, const void* ln_weight,
                        const void* ln_bias, float ln_eps, void* normed_out, const tome_view* normed_view, void* stream) {
  return tome_merge_add_norm_rv(plan, x, residual, nullptr, dtype, c, x_view, size_in, mode, hybrid_threshold, out, out_view, size_out,
                                logsize_out, ln_weight, ln_bias, ln_eps, normed_out, normed_view, stream);
}

int tome_merge_add_norm_rv(const tome_plan* plan, const void* x, const void* residual, const tome_view* residual_view, int32_t dtype,
                           int32_t c, const tome_view* x_view, const float* size_in, int32_t mode, float hybrid_threshold, void* out,
                           const tome_view* out_view, float* size_out, float* logsize_out, const void* ln_weight,
                           const void* ln_bias, float ln_eps, void* normed_out, const tome_view* normed_view, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_merge_norm");
  if (rc) return rc;
  TOME_CHECK_ARG(x && out && c > 0, "tome_merge_norm: NULL tensor or c=%d", c);
  TOME_CHECK_ARG((ln_weight != nullptr) == (normed_out != nullptr), "tome_merge_norm: LayerNorm weight and normed_out go together");
  TOME_CHECK_ARG(ln_weight || residual, "tome_merge_norm: nothing to fuse (use tome_merge)");
  TOME_CHECK_ARG(residual != out && residual != normed_out, "tome_merge_norm: buffers must not alias");
  TOME_CHECK_ARG(mode >= TOME_MODE_WAVG && mode <= TOME_MODE_DROP, "tome_merge_norm: unknown mode %d", mode);
  TOME_CHECK_ARG(mode == TOME_MODE_WAVG || size_in == nullptr, "tome_merge_norm: size_in is only meaningful for TOME_MODE_WAVG");
  TOME_CHECK_ARG(x != out && x != normed_out && out != normed_out, "tome_merge_norm: buffers must not alias");
  const View xv = make_view(x_view, plan->n, c), ov = make_view(out_view, plan->n - plan->r, c);
  const View nv = make_view(normed_view, plan->n - plan->r, c);
  const View rv = residual_view ? make_view(residual_view, plan->n, c) : xv;
  return launch_merge(plan, x, dtype, c, xv, size_in, mode, hybrid_threshold, out, ov, size_out, logsize_out,
                      (cudaStream_t)stream, ln_weight, ln_bias, ln_eps, normed_out, &nv, residual, &rv);
}

int tome_rows_add_layernorm(const void* a, const int64_t* a_strides, const void* b, const int64_t* b_strides, int32_t dtype, int32_t nb,
                            int32_t np, int32_t nt, int32_t c, const void* ln_weight, const void* ln_bias, float ln_eps, void* sum_out,
                            const int64_t* sum_strides, void* normed_out, const int64_t* normed_strides, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(a && a_strides && nb > 0 && np > 0 && nt > 0 && c > 0, "tome_rows_add_layernorm: NULL input or empty shape");
  TOME_CHECK_ARG(sum_out || normed_out, "tome_rows_add_layernorm: nothing to write");
  TOME_CHECK_ARG((!b || b_strides) && (!sum_out || sum_strides) && (!normed_out || (normed_strides && ln_weight)),
                 "tome_rows_add_layernorm: a tensor without its strides (or LayerNorm without a weight)");
  static_assert(sizeof(long long) == sizeof(int64_t), "stride type");
  return launch_rows_add_layernorm(a, (const long long*)a_strides, b, (const long long*)b_strides, dtype, nb, np, nt, c, ln_weight, ln_bias,
                                   ln_eps, sum_out, (const long long*)sum_strides, normed_out, (const long long*)normed_strides,
                                   (cudaStream_t)stream);
}

int tome_merge_source(const tome_plan* plan, const float* source, int32_t n0, float hybrid_threshold, float* out,
                      void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_merge_source");
  if (rc) return rc;
  TOME_CHECK_ARG(out, "tome_merge_source: out is NULL");
  TOME_CHECK_ARG(source || n0 == plan->n, "tome_merge_source: implicit identity needs n0 == n (n0=%d n=%d)", n0, plan->n);
  TOME_CHECK_ARG(n0 > 0, "tome_merge_source: n0=%d", n0);
  return launch_merge_source(plan, source, n0, hybrid_threshold, out, (cudaStream_t)stream);
}

int tome_add_layernorm(const void* a, const void* b, int32_t dtype, int64_t rows, int32_t c, const void* ln_weight,
                       const void* ln_bias, float ln_eps, void* sum_out, void* normed_out, void* stream) {
  return tome_add_rows_layernorm(a, b, rows, dtype, rows, c, ln_weight, ln_bias, ln_eps, sum_out, normed_out, stream);
}

int tome_add_rows_layernorm(const void* a, const void* b, int64_t b_rows, int32_t dtype, int64_t rows, int32_t c,
                            const void* ln_weight, const void* ln_bias, float ln_eps, void* sum_out, void* normed_out,
                            void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(a && b && ln_weight && sum_out && normed_out && rows > 0 && c > 0 && b_rows > 0 && b_rows <= rows,
                 "tome_add_layernorm: NULL pointer or empty shape");
  return launch_add_layernorm(a, b, b_rows, dtype, rows, c, ln_weight, ln_bias, ln_eps, sum_out, normed_out, (cudaStream_t)stream);
}

int tome_linear_gelu(const void* x, const void* w, const void* bias, int32_t m, int32_t n, int32_t k, int64_t x_row_stride,
                     int32_t gelu, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x && w && out && m > 0 && n > 0 && k > 0 && x_row_stride >= k, "tome_linear_gelu: NULL pointer or bad shape");
  TOME_CHECK_ARG(x != out, "tome_linear_gelu: in-place is not supported");
  return launch_linear_gelu(x, w, bias, m, n, k, x_row_stride, gelu, out, (cudaStream_t)stream);
}

int tome_patchify(const void* x, int32_t in_dtype, int32_t b, int32_t c, int32_t t, int32_t h, int32_t w, int32_t tt, int32_t ph,
                  int32_t pw, void* out, int32_t out_dtype, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x && out && b > 0 && c > 0 && t > 0 && h > 0 && w > 0 && tt > 0 && ph > 0 && pw > 0, "tome_patchify: NULL pointer or empty shape");
  return launch_patchify(x, in_dtype, b, c, t, h, w, tt, ph, pw, out, out_dtype, (cudaStream_t)stream);
}

int tome_attn_key_bias(const float* log_size, int32_t b, int32_t n, int32_t lead, int32_t heads, int32_t d, float scale,
                       int32_t dtype, void* k, int64_t k_sb, int64_t k_sn, int64_t k_sh, void* q, int64_t q_sb, int64_t q_sn,
                       int64_t q_sh, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(log_size && k && b > 0 && heads > 0 && d > 0 && lead >= 0 && n > lead && scale > 0.f,
                 "tome_attn_key_bias: NULL pointer or bad shape (b=%d n=%d lead=%d heads=%d d=%d scale=%g)", b, n, lead, heads, d, (double)scale);
  return launch_key_bias(log_size, b, n, lead, heads, d, 1.0f / scale, dtype, k, k_sb, k_sn, k_sh, q, q_sb, q_sn, q_sh,
                         (cudaStream_t)stream);
}

int tome_unmerge(const tome_plan* plan, const void* x, int32_t dtype, int32_t c, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_unmerge");
  if (rc) return rc;
  TOME_CHECK_ARG(x && out && c > 0, "tome_unmerge: NULL tensor or c=%d", c);
  return launch_unmerge(plan, x, dtype, c, out, (cudaStream_t)stream);
}

size_t tome_match_sets_workspace_bytes(int32_t bm, int32_t ra, int32_t nb, int32_t cm) {
  if (bm <= 0 || ra <= 0 || nb <= 0 || cm <= 0) return 0;
  return match_sets_workspace(bm, ra + nb, ra, cm);
}

int tome_match_sets(const void* metric, int32_t dtype, int32_t bm, int32_t n, int32_t cm, const tome_view* view,
                    const int32_t* tok_of_row, int64_t tok_stride_b, int32_t ra, int32_t nb, float* node_max,
                    int32_t* node_idx, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(metric && tok_of_row && node_max && node_idx && workspace, "tome_match_sets: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && n > 0 && cm > 0 && ra > 0 && nb > 0 && ra + nb <= n, "tome_match_sets: bad shape bm=%d n=%d cm=%d ra=%d nb=%d", bm, n, cm, ra, nb);
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_match_sets: unsupported dtype %d", dtype);
  TOME_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "tome_match_sets: workspace must be 256-byte aligned");
  return launch_match_sets(metric, dtype, bm, cm, make_view(view, n, cm), tok_of_row, tok_stride_b, ra, nb, node_max, node_idx,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

int tome_group_reduce(const void* x, int32_t dtype, int32_t bm, int32_t n, int32_t c, const tome_view* x_view,
                      const int32_t* tok_of_row, int64_t tok_stride_b, int32_t ra, int32_t nb, const int32_t* dst_idx,
                      int32_t mode, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x && tok_of_row && dst_idx && out, "tome_group_reduce: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && c > 0 && ra > 0 && nb > 0 && ra + nb <= n, "tome_group_reduce: bad shape bm=%d n=%d c=%d ra=%d nb=%d", bm, n, c, ra, nb);
  TOME_CHECK_ARG(mode == TOME_MODE_SUM || mode == TOME_MODE_MEAN || mode == TOME_MODE_AMAX, "tome_group_reduce: mode %d is not sum / mean / amax", mode);
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_group_reduce: unsupported dtype %d", dtype);
  return launch_group_reduce(x, dtype, bm, c, make_view(x_view, n, c), tok_of_row, tok_stride_b, ra, nb, dst_idx, mode, out,
                             (cudaStream_t)stream);
}

int tome_gather_rows(const void* x, int32_t dtype, int32_t bm, int32_t n_in, int32_t c, const int32_t* map, int32_t n_out,
                     void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x && map && out && bm > 0 && n_in > 0 && n_out > 0 && c > 0, "tome_gather_rows: bad argument");
  if (dtype != TOME_F32 && dtype != TOME_BF16) return set_error(TOME_ERR_DTYPE, "tome_gather_rows: unsupported dtype %d", dtype);
  return launch_gather_rows(x, dtype, bm, n_in, c, map, n_out, out, (cudaStream_t)stream);
}

int tome_source_compose(const tome_plan* plan, const int32_t* group_in, int32_t n0, int32_t drop, float hybrid_threshold,
                        int32_t* group_out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  rc = check_plan(plan, "tome_source_compose");
  if (rc) return rc;
  TOME_CHECK_ARG(group_out && n0 > 0, "tome_source_compose: NULL output or n0=%d", n0);
  TOME_CHECK_ARG(group_in || n0 == plan->n, "tome_source_compose: implicit identity needs n0 == n (n0=%d n=%d)", n0, plan->n);
  return launch_source_compose(plan, group_in, n0, drop, hybrid_threshold, group_out, (cudaStream_t)stream);
}

int tome_source_dense(const int32_t* group, int32_t bm, int32_t n_tokens, int32_t n0, float* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(group && out && bm > 0 && n_tokens > 0 && n0 > 0, "tome_source_dense: NULL pointer or empty shape");
  if (bm > TOME_MAX_BATCH) return set_error(TOME_ERR_UNSUPPORTED, "tome_source_dense: matching batch %d > %d; split the batch", bm, TOME_MAX_BATCH);
  return launch_source_dense(group, bm, n_tokens, n0, out, (cudaStream_t)stream);
}

int tome_random_rowmax(void* philox_state, int64_t clip0, int32_t bm, int32_t na, int32_t nb, int32_t class_token,
                       int32_t distill_token, float* node_max, int32_t* node_idx, float* scores_out, int32_t advance, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(philox_state && node_max && node_idx, "tome_random_rowmax: NULL pointer argument");
  TOME_CHECK_ARG(bm > 0 && na > 0 && nb > 0 && clip0 >= 0, "tome_random_rowmax: bad shape bm=%d na=%d nb=%d", bm, na, nb);
  TOME_CHECK_ARG(((uintptr_t)philox_state & 7) == 0, "tome_random_rowmax: philox_state must be 8-byte aligned");
  if (bm > TOME_MAX_BATCH) return set_error(TOME_ERR_UNSUPPORTED, "tome_random_rowmax: matching batch %d > %d; split the batch", bm, TOME_MAX_BATCH);
  return launch_random_rowmax(philox_state, clip0, bm, na, nb, class_token, distill_token, node_max, node_idx, scores_out, advance,
                              (cudaStream_t)stream);
}

int tome_attn_short(const void* q, const void* k, const void* v, int32_t dtype, int64_t seqs, int32_t n_tok, int32_t heads, int32_t d,
                    int64_t seq_stride, int64_t tok_stride, float scale, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(q && k && v && out && seqs > 0 && n_tok > 0 && heads > 0 && d > 0, "tome_attn_short: NULL pointer or empty shape");
  TOME_CHECK_ARG(seqs * heads * n_tok < (1LL << 38), "tome_attn_short: too many rows");
  return launch_attn_short(q, k, v, dtype, seqs, n_tok, heads, d, seq_stride, tok_stride, scale, out, (cudaStream_t)stream);
}

int tome_frames_attention(const void* qkv, int32_t dtype, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t frames,
                          int32_t keys_per_frame, int32_t lead, int32_t unbiased_queries, float scale, const float* key_bias, void* xs,
                          void* x_diag, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(qkv && xs && b > 0 && n > 1 && heads > 0 && frames > 0 && keys_per_frame > 0, "tome_frames_attention: NULL pointer or empty shape");
  if (dtype != TOME_BF16 || d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_frames_attention: bf16 with head dimension 64 only (dtype %d, d %d)", dtype, d);
  TOME_CHECK_ARG(lead >= 0 && unbiased_queries >= 0, "tome_frames_attention: negative lead");
  return launch_frames_attention(qkv, b, n, heads, frames, keys_per_frame, scale, key_bias, xs, x_diag, lead, unbiased_queries,
                                 (cudaStream_t)stream);
}

int tome_traj_temporal(const void* q2, const void* k2, const void* vals, int32_t dtype, int64_t rows, int32_t frames, int32_t heads,
                       int32_t d, float scale, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(q2 && k2 && vals && out && rows > 0 && frames > 0 && heads > 0, "tome_traj_temporal: NULL pointer or empty shape");
  if ((dtype != TOME_BF16 && dtype != TOME_F32) || d != 64)
    return set_error(TOME_ERR_UNSUPPORTED, "tome_traj_temporal: bf16 / fp32 with head dimension 64 only (dtype %d, d %d)", dtype, d);
  return launch_traj_temporal(q2, k2, vals, dtype, rows, frames, heads, scale, out, (cudaStream_t)stream);
}

int tome_split3(const void* x, int64_t rows, int32_t k, int64_t row_stride, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x && out && rows > 0 && k > 0 && row_stride >= k, "tome_split3: NULL pointer or bad shape");
  return launch_split3(x, rows, k, row_stride, out, (cudaStream_t)stream);
}

int tome_planes_sum(const void* x3, int64_t rows, int32_t n, int32_t col0, int32_t ncols, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x3 && out && rows > 0 && n > 0, "tome_planes_sum: NULL pointer or empty shape");
  return launch_planes_sum(x3, rows, n, col0, ncols, out, (cudaStream_t)stream);
}

int tome_linear_f32(const void* x3, const void* w3, const void* bias, int32_t m, int32_t n, int32_t k, int32_t gelu, int32_t terms,
                    void* out, void* out_planes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(x3 && w3 && (out || out_planes) && m > 0 && n > 0 && k > 0, "tome_linear_f32: NULL pointer or bad shape");
  TOME_CHECK_ARG(gelu >= 0 && gelu <= 2, "tome_linear_f32: gelu must be 0, 1 or 2");
  return launch_linear_f32(x3, w3, bias, m, n, k, gelu, terms, out, out_planes, (cudaStream_t)stream);
}

int tome_attention_f32(const void* qkv3, int32_t b, int32_t n, int32_t heads, int32_t d, float scale, const float* key_bias,
                       int32_t unbiased_queries, void* out, void* out_planes, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(qkv3 && (out || out_planes) && b > 0 && n > 0 && heads > 0 && unbiased_queries >= 0, "tome_attention_f32: NULL pointer or empty shape");
  if (d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_attention_f32: head dimension %d (64 only)", d);
  return launch_attention_f32(qkv3, b, n, heads, scale, key_bias, unbiased_queries, out, out_planes, (cudaStream_t)stream);
}

int tome_cls_attention(const void* qkv, int32_t dtype, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t query_token, float scale,
                       void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(qkv && out && b > 0 && n > 0 && heads > 0, "tome_cls_attention: NULL pointer or empty shape");
  if (d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_cls_attention: head dimension %d (64 only)", d);
  if ((long long)b * heads > 0x7fffffffLL) return set_error(TOME_ERR_UNSUPPORTED, "tome_cls_attention: too many (clip, head) pairs");
  return launch_cls_attention(qkv, dtype, b, n, heads, query_token, scale, out, (cudaStream_t)stream);
}

int tome_frames_attention_f32(const void* qkv3, int32_t b, int32_t n, int32_t heads, int32_t d, int32_t frames, int32_t keys_per_frame,
                              int32_t lead, float scale, const float* key_bias, void* xs, void* xs_planes, void* x_diag, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(qkv3 && (xs || xs_planes) && b > 0 && n > 0 && heads > 0 && frames > 0 && keys_per_frame > 0 && lead >= 0,
                 "tome_frames_attention_f32: NULL pointer or empty shape");
  if (d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_frames_attention_f32: head dimension %d (64 only)", d);
  return launch_frames_attention_f32(qkv3, b, n, heads, frames, keys_per_frame, lead, scale, key_bias, 0, xs, xs_planes, x_diag,
                                     (cudaStream_t)stream);
}

int tome_attention_bf16(const void* qkv, int32_t b, int32_t n, int32_t heads, int32_t d, float scale, const float* key_bias,
                        int32_t unbiased_queries, void* out, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(qkv && out && b > 0 && n > 0 && heads > 0 && unbiased_queries >= 0, "tome_attention_bf16: NULL pointer or empty shape");
  if (d != 64) return set_error(TOME_ERR_UNSUPPORTED, "tome_attention_bf16: head dimension %d (64 only)", d);
  return launch_attention_bf16(qkv, b, n, heads, scale, key_bias, unbiased_queries, out, (cudaStream_t)stream);
}

int tome_cls_rows(const void* a, int64_t a_stride_b, const void* add, int64_t add_stride_b, const void* mean_src, int64_t mean_stride_b,
                  int64_t mean_stride_t, int32_t mean_t, int32_t dtype, int32_t batch, int32_t c, void* sum_out, int64_t sum_stride_b,
                  const void* ln_weight, const void* ln_bias, float ln_eps, void* normed_out, int64_t normed_stride_b,
                  int64_t normed_stride_rep, int32_t reps, void* stream) {
  int rc = ensure_device_ok();
  if (rc) return rc;
  TOME_CHECK_ARG(a && batch > 0 && c > 0 && (sum_out || normed_out), "tome_cls_rows: NULL input or nothing to write");
  TOME_CHECK_ARG(!normed_out || (ln_weight && reps > 0), "tome_cls_rows: LayerNorm output without a weight / replica count");
  TOME_CHECK_ARG(!mean_src || mean_t > 0, "tome_cls_rows: mean source without a frame count");
  const int e = dtype == TOME_F32 ? 4 : 8;
  const uintptr_t al = (uintptr_t)a | (uintptr_t)add | (uintptr_t)mean_src | (uintptr_t)sum_out | (uintptr_t)normed_out | (uintptr_t)ln_weight | (uintptr_t)ln_bias;
  if ((al & 15) || a_stride_b % e || add_stride_b % e || mean_stride_b % e || mean_stride_t % e || sum_stride_b % e || normed_stride_b % e || normed_stride_rep % e)
    return set_error(TOME_ERR_ALIGN, "tome_cls_rows: 16-byte aligned rows required");
  ClsArgs g;
  g.a = a; g.add = add; g.mean_src = mean_src; g.a_sb = a_stride_b; g.add_sb = add_stride_b; g.m_sb = mean_stride_b; g.m_st = mean_stride_t;
  g.mean_t = mean_src ? mean_t : 0; g.sum_out = sum_out; g.normed_out = normed_out; g.s_sb = sum_stride_b; g.n_sb = normed_stride_b;
  g.n_sr = normed_stride_rep; g.reps = reps; g.c = c; g.ln_w = ln_weight; g.ln_b = ln_bias; g.eps = ln_eps;
  return launch_cls_rows(g, dtype, batch, (cudaStream_t)stream);
}

}  // extern "C"
