// C ABI (include/svx.h): handle management, error reporting, and the scorer that strings the scoring kernels
// together (cohort GEMM on tcgen05 → exact top-k statistics → trial gather).
#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "kernels.cuh"
#include <cstdlib>

#include "model.h"
#include "umma.cuh"

namespace svx {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace svx

using namespace svx;

struct svx_extractor { Model* model; };

struct svx_scorer {
  int device = 0;
  long long launches = 0;
  // workspaces
  __nv_bfloat16* d_a = nullptr; size_t a_bytes = 0;     // split test block [block_rows, 3d]
  __nv_bfloat16* d_b = nullptr; size_t b_bytes = 0;     // split cohort [c_pad, 3d]
  float* d_s = nullptr; size_t s_bytes = 0;             // score block [block_rows, c_pad]
  // fused path (asnorm_fused.cu): split operands [rows_pad, 2*dp] and the rows handed back to the unfused path
  __nv_bfloat16* d_a2 = nullptr; size_t a2_bytes = 0;
  __nv_bfloat16* d_b2 = nullptr; size_t b2_bytes = 0;
  int* d_flag = nullptr; size_t flag_bytes = 0;          // [0] = count, [1..] = row indices
  float* d_fx = nullptr; size_t fx_bytes = 0;            // gathered flagged rows, then their results
  float* d_fr = nullptr; size_t fr_bytes = 0;
  int sms = 148;
  int use_fused = 1;                                     // option "fused" (svx_scorer_set_option)
  long long fused_rows = 0, fallback_rows = 0;           // of the last call
};

#define API_CUDA(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                 \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

static int require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_last_error(std::string("no CUDA device available: ") + cudaGetErrorString(e) + " — libsvx has no CPU fallback");
    return 1;
  }
  if (device < 0 || device >= n) { set_last_error("invalid device index"); return 1; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { set_last_error("cudaGetDeviceProperties failed"); return 1; }
  if (p.major != 10) {
    set_last_error("libsvx is built for sm_100a (B200); device is sm_" + std::to_string(p.major) + std::to_string(p.minor));
    return 1;
  }
  return 0;
}

// Every entry point that launches work runs on the device of its handle (or, without a handle, the device the caller's
// stream belongs to = the caller's current device) and leaves the calling thread's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (device >= 0 && device != prev) cudaSetDevice(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename P>
static int grow_buf(P** p, size_t* have, size_t need) {
  if (need <= *have) return 0;
  cudaDeviceSynchronize();
  cudaFree(*p);
  *p = nullptr;
  if (cudaMalloc(p, need) != cudaSuccess) { set_last_error("cudaMalloc failed in scorer"); return 1; }
  *have = need;
  return 0;
}

extern "C" {

int svx_version(void) { return 200; }

int svx_build_flags(void) {
  int f = 0;
#ifdef SVX_DEBUG_SWITCHES
  f |= 1;
#endif
#ifdef SVX_ENABLE_PAIR
  f |= 2;
#endif
  return f;
}

const char* svx_last_error(void) { return g_last_error.c_str(); }

int svx_extractor_create(const svx_model_config* cfg, int device, int precision, svx_extractor** out) {
  if (!cfg || !out) { set_last_error("null argument"); return 1; }
  *out = nullptr;
  if (precision != SVX_PRECISION_FP16 && precision != SVX_PRECISION_BF16) { set_last_error("unknown precision"); return 1; }
  if (require_device(device)) return 1;
  Model* m = new (std::nothrow) Model(*cfg, device, precision);
  if (!m) { set_last_error("out of host memory"); return 1; }
  if (m->build()) { delete m; return 1; }
  svx_extractor* h = new svx_extractor{m};
  *out = h;
  return 0;
}

int svx_extractor_destroy(svx_extractor* h) {
  if (!h) return 0;
  DeviceGuard guard(h->model->device());
  delete h->model;
  delete h;
  return 0;
}

int svx_extractor_num_tensors(svx_extractor* h) { return h ? static_cast<int>(h->model->vars().size()) : -1; }

int svx_extractor_tensor_info(svx_extractor* h, int index, const char** name, int* ndim, int64_t shape[4]) {
  if (!h || index < 0 || index >= static_cast<int>(h->model->vars().size())) { set_last_error("bad tensor index"); return 1; }
  const VarSpec& v = h->model->vars()[index];
  if (name) *name = v.name.c_str();
  if (ndim) *ndim = static_cast<int>(v.shape.size());
  if (shape) for (size_t i = 0; i < v.shape.size() && i < 4; ++i) shape[i] = v.shape[i];
  return 0;
}

int svx_extractor_set_tensor(svx_extractor* h, const char* name, const float* data, int ndim, const int64_t* shape) {
  if (!h || !name || !data || !shape) { set_last_error("null argument"); return 1; }
  return h->model->set_tensor(name, data, ndim, shape);
}

int svx_extractor_finalize(svx_extractor* h) {
  if (!h) { set_last_error("null handle"); return 1; }
  DeviceGuard guard(h->model->device());
  return h->model->finalize();
}
int svx_extractor_embed_dim(svx_extractor* h) { return h ? h->model->embed_dim() : -1; }
int svx_extractor_set_option(svx_extractor* h, const char* key, int value) {
  if (!h || !key) { set_last_error("null argument"); return 1; }
  return h->model->set_option(key, value);
}

int svx_extractor_run_segments(svx_extractor* h, const float* feats_dev, const int32_t* frame_offsets_host, int n_segments,
                               float* out_dev, void* cuda_stream) {
  if (!h || !feats_dev || !frame_offsets_host || !out_dev) { set_last_error("null argument"); return 1; }
  DeviceGuard guard(h->model->device());
  return h->model->run_segments(feats_dev, frame_offsets_host, n_segments, out_dev, static_cast<cudaStream_t>(cuda_stream));
}

int svx_extractor_extract(svx_extractor* h, const float* feats, int feats_on_device, const int32_t* frame_offsets_host, int n_utts,
                          float* out, int out_on_device, void* cuda_stream) {
  if (!h || !feats || !frame_offsets_host || !out) { set_last_error("null argument"); return 1; }
  DeviceGuard guard(h->model->device());
  return h->model->extract(feats, feats_on_device, frame_offsets_host, n_utts, out, out_on_device,
                           static_cast<cudaStream_t>(cuda_stream));
}

int svx_extractor_set_dump_dir(svx_extractor* h, const char* dir) {
  if (!h) { set_last_error("null handle"); return 1; }
  h->model->set_dump_dir(dir);
  return 0;
}

long long svx_extractor_last_launches(svx_extractor* h) { return h ? h->model->last_launches() : -1; }

int svx_extractor_conv_time(svx_extractor* h, double* ms, double* flops) {
  if (!h || !ms || !flops) { set_last_error("null argument"); return 1; }
  return h->model->conv_time(ms, flops);
}

// ------------------------------------------------------------------------------------------------ scoring
int svx_scorer_create(int device, svx_scorer** out) {
  if (!out) { set_last_error("null argument"); return 1; }
  *out = nullptr;
  if (require_device(device)) return 1;
  DeviceGuard guard(device);
  API_CUDA(conv_umma_init());
  API_CUDA(asnorm_fused_init());
  svx_scorer* s = new svx_scorer();
  s->device = device;
  cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device);
  *out = s;
  return 0;
}

int svx_scorer_destroy(svx_scorer* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaFree(h->d_a); cudaFree(h->d_b); cudaFree(h->d_s);
  cudaFree(h->d_a2); cudaFree(h->d_b2); cudaFree(h->d_flag); cudaFree(h->d_fx); cudaFree(h->d_fr);
  delete h;
  return 0;
}

int svx_cmvn_sliding(const float* feats_dev, float* out_dev, const int32_t* frame_offsets_host, int n_utts, int feat_dim, int cmn_window,
                     int center, int min_window, void* cuda_stream) {
  if (!feats_dev || !out_dev || !frame_offsets_host) { set_last_error("null argument"); return 1; }
  if (n_utts <= 0) return 0;
  if (feat_dim <= 0 || feat_dim > 128 || cmn_window <= 0) { set_last_error("feat_dim must be in 1..128 and cmn_window positive"); return 1; }
  if (min_window <= 0 || min_window > cmn_window) { set_last_error("min_window must be in 1..cmn_window"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const long long total = frame_offsets_host[n_utts];
  for (int i = 0; i < n_utts; ++i)
    if (frame_offsets_host[i + 1] <= frame_offsets_host[i]) { set_last_error("empty utterance"); return 1; }
  // workspaces live for this call only: offsets, frame→utterance map, double running sums (one extra row per utterance)
  int32_t* d_off = nullptr; int32_t* d_utt = nullptr; double* d_csum = nullptr;
  API_CUDA(cudaMallocAsync(&d_off, static_cast<size_t>(n_utts + 1) * 4, st));
  API_CUDA(cudaMallocAsync(&d_utt, static_cast<size_t>(total) * 4, st));
  API_CUDA(cudaMallocAsync(&d_csum, static_cast<size_t>(total + n_utts) * feat_dim * 8, st));
  API_CUDA(cudaMemcpyAsync(d_off, frame_offsets_host, static_cast<size_t>(n_utts + 1) * 4, cudaMemcpyHostToDevice, st));
  cudaError_t e = launch_cmn_sliding(feats_dev, out_dev, d_off, n_utts, total, feat_dim, cmn_window, center, min_window, d_csum, d_utt, st);
  cudaFreeAsync(d_off, st); cudaFreeAsync(d_utt, st); cudaFreeAsync(d_csum, st);
  API_CUDA(e);
  API_CUDA(cudaStreamSynchronize(st));   // frame_offsets_host is the caller's buffer
  return 0;
}

int svx_decode_compressed(const uint8_t* blob_dev, int64_t blob_bytes, const int64_t* record_offsets_host, const int32_t* frame_offsets_host,
                          int n_records, int feat_dim, float* out_dev, void* cuda_stream) {
  if (!blob_dev || !record_offsets_host || !frame_offsets_host || !out_dev) { set_last_error("null argument"); return 1; }
  if (n_records <= 0) return 0;
  if (feat_dim <= 0) { set_last_error("feat_dim must be positive"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const long long total = frame_offsets_host[n_records];
  // what can be checked on the host: every record (header + per-column percentiles + rows x cols bytes) fits between its
  // offset and the next one; the headers themselves live on the device and are checked there (cm_check_kernel)
  for (int i = 0; i < n_records; ++i) {
    const long long rows = frame_offsets_host[i + 1] - frame_offsets_host[i];
    const long long need = 16 + 8LL * feat_dim + rows * feat_dim;
    const long long end = i + 1 < n_records ? record_offsets_host[i + 1] : blob_bytes;
    if (rows <= 0 || record_offsets_host[i] < 0 || record_offsets_host[i] + need > end) {
      set_last_error("compressed record " + std::to_string(i) + ": " + std::to_string(rows) + " rows x " + std::to_string(feat_dim) +
                     " columns need " + std::to_string(need) + " bytes, the record has " + std::to_string(end - record_offsets_host[i]));
      return 1;
    }
  }
  long long* d_rec = nullptr; int32_t* d_off = nullptr; int32_t* d_utt = nullptr; int32_t* d_bad = nullptr;
  API_CUDA(cudaMallocAsync(&d_bad, 4, st));
  API_CUDA(cudaMemsetAsync(d_bad, 0x7f, 4, st));
  static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
  API_CUDA(cudaMallocAsync(&d_rec, static_cast<size_t>(n_records) * 8, st));
  API_CUDA(cudaMallocAsync(&d_off, static_cast<size_t>(n_records + 1) * 4, st));
  API_CUDA(cudaMallocAsync(&d_utt, static_cast<size_t>(total > 0 ? total : 1) * 4, st));
  API_CUDA(cudaMemcpyAsync(d_rec, record_offsets_host, static_cast<size_t>(n_records) * 8, cudaMemcpyHostToDevice, st));
  API_CUDA(cudaMemcpyAsync(d_off, frame_offsets_host, static_cast<size_t>(n_records + 1) * 4, cudaMemcpyHostToDevice, st));
  cudaError_t e = launch_cm_decode(blob_dev, d_rec, d_off, n_records, total, feat_dim, out_dev, d_utt, blob_bytes, d_bad, st);
  int32_t bad = kCmNoBadRecord;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(d_rec, st); cudaFreeAsync(d_off, st); cudaFreeAsync(d_utt, st); cudaFreeAsync(d_bad, st);
  API_CUDA(e);
  API_CUDA(cudaStreamSynchronize(st));
  if (bad != kCmNoBadRecord) {
    set_last_error("compressed record " + std::to_string(bad) + ": its header (rows, cols) does not match frame_offsets / feat_dim " +
                   std::to_string(feat_dim) + " — nothing was decoded");
    return 1;
  }
  return 0;
}

int svx_l2norm_rows(const float* in_dev, float* out_dev, int64_t n, int d, void* cuda_stream) {
  if (!in_dev || !out_dev) { set_last_error("null argument"); return 1; }
  API_CUDA(launch_l2norm_rows(in_dev, out_dev, n, d, static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

int svx_group_means(const float* unit_rows_dev, int d, const int32_t* member_rows_dev, const int32_t* group_offsets_dev, float* out_dev,
                    int n_groups, void* cuda_stream) {
  if (!unit_rows_dev || !member_rows_dev || !group_offsets_dev || !out_dev) { set_last_error("null argument"); return 1; }
  if (d <= 0 || n_groups < 0) { set_last_error("bad group-mean shape"); return 1; }
  API_CUDA(launch_group_mean(unit_rows_dev, d, member_rows_dev, group_offsets_dev, out_dev, n_groups, static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

static int cohort_pass(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                       float* mean_dev, float* std_dev, float* vals_dev, void* cuda_stream);
static int cohort_pass_unfused(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                               float* mean_dev, float* std_dev, float* vals_dev, int vals_ld, cudaStream_t st);

int svx_scorer_set_option(svx_scorer* h, const char* key, int value) {
  if (!h || !key) { set_last_error("null argument"); return 1; }
  if (!strcmp(key, "fused")) { h->use_fused = value; return 0; }
  set_last_error(std::string("unknown scorer option: ") + key);
  return 1;
}

int svx_scorer_last_path(svx_scorer* h, long long* fused_rows, long long* fallback_rows) {
  if (!h) { set_last_error("null handle"); return 1; }
  if (fused_rows) *fused_rows = h->fused_rows;
  if (fallback_rows) *fallback_rows = h->fallback_rows;
  return 0;
}

int svx_asnorm_stats(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                     float* mean_dev, float* std_dev, void* cuda_stream) {
  if (!h || !test_dev || !cohort_dev || !mean_dev || !std_dev) { set_last_error("null argument"); return 1; }
  return cohort_pass(h, test_dev, n, cohort_dev, c, d, topk, mean_dev, std_dev, nullptr, cuda_stream);
}

int svx_cohort_topk_values(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                           float* vals_dev, void* cuda_stream) {
  if (!h || !test_dev || !cohort_dev || !vals_dev) { set_last_error("null argument"); return 1; }
  return cohort_pass(h, test_dev, n, cohort_dev, c, d, topk, nullptr, nullptr, vals_dev, cuda_stream);
}

int svx_topk_stats(const float* vals_dev, int ld, int64_t n, int m, int topk, float* mean_dev, float* std_dev, void* cuda_stream) {
  if (!vals_dev || !mean_dev || !std_dev) { set_last_error("null argument"); return 1; }
  if (m <= 0 || topk <= 0 || ld < m) { set_last_error("bad top-k merge shape"); return 1; }
  API_CUDA(launch_topk_stats(vals_dev, ld, n, m, topk, mean_dev, std_dev, nullptr, 0, static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

}  // extern "C"

// z with P(Z > z) = q for a standard normal Z (bisection on erfc): where the k-th largest of c roughly-normal scores sits.
static double normal_upper_quantile(double q) {
  double lo = -8.0, hi = 8.0;
  for (int i = 0; i < 80; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (0.5 * std::erfc(mid / std::sqrt(2.0)) > q) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

static int cohort_pass(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                       float* mean_dev, float* std_dev, float* vals_dev, void* cuda_stream) {
  if (c <= 0 || d <= 0 || topk <= 0) { set_last_error("cohort size, dimension and topk must be positive"); return 1; }
  if (d % 8 != 0) { set_last_error("embedding dimension must be a multiple of 8"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(h->device);
  h->launches = 0; h->fused_rows = 0; h->fallback_rows = 0;
  if (n <= 0) return 0;
  // The fused kernel takes the shapes it was built for: operand resident in shared memory (d <= 256), a threshold well inside the
  // cohort (k <= c/3) and a cohort large enough for the first 128 scores to be a sample (c >= 1024).  Everything else — and any
  // row the fused kernel hands back — goes through the unfused kernels.
  static const bool env_unfused = dbg_env("SVX_SCORE_UNFUSED") != nullptr;   // debug switch
  const bool fused_ok = h->use_fused && !env_unfused && d <= 256 && d % 4 == 0 && c >= 1024 && c <= 65535 && static_cast<long long>(topk) * 3 <= c && n <= (1LL << 30);
  if (!fused_ok) return cohort_pass_unfused(h, test_dev, n, cohort_dev, c, d, topk, mean_dev, std_dev, vals_dev, topk, st);

  const int dp = (d + 63) / 64 * 64;
  const long long n_pad = (n + 127) / 128 * 128;
  const int c_pad = (c + 127) / 128 * 128;
  if (grow_buf(&h->d_b2, &h->b2_bytes, static_cast<size_t>(c_pad) * 2 * dp * 2)) return 1;
  if (grow_buf(&h->d_flag, &h->flag_bytes, (static_cast<size_t>(n) + 1) * 4)) return 1;
  API_CUDA(launch_split2(cohort_dev, h->d_b2, c, c_pad, d, dp, st));
  API_CUDA(cudaMemsetAsync(h->d_flag, 0, 4, st));
  AsnormFusedParams fp;
  memset(&fp, 0, sizeof fp);
  fp.x = test_dev; fp.d = d;
  fp.n_rows = static_cast<int>(n); fp.c = c; fp.n_row_blocks = static_cast<int>(n_pad / 128); fp.n_tiles = c_pad / 128;
  fp.dp = dp; fp.kboxes = dp / 64; fp.topk = topk;
  const double zq = normal_upper_quantile(static_cast<double>(topk) / c);
  // statistics only: one-term pass 0, candidates = three bins around the threshold bin -> narrower bins over a narrower range
  if (vals_dev) { fp.z_lo = static_cast<float>(zq - 1.0); fp.z_hi = static_cast<float>(zq + 2.2); fp.nb = 96; }
  else { fp.z_lo = static_cast<float>(zq - 0.9); fp.z_hi = static_cast<float>(zq + 1.3); fp.nb = 112; }
  fp.cap = 48;                                // candidates per epilogue group; histogram counts are 16-bit (c <= 65535)
  fp.mean = mean_dev; fp.stdv = std_dev; fp.vals = vals_dev; fp.vals_ld = topk;
  fp.flag_count = h->d_flag; fp.flag_rows = h->d_flag + 1;
  fp.stages = 12;
  { static const char* kn = dbg_env("SVX_ASNORM_KNOCK"); fp.knock = kn ? atoi(kn) : 0; }
  while (fp.stages > 2 && asnorm_fused_smem_bytes(fp) > 227 * 1024) --fp.stages;
  if (asnorm_fused_smem_bytes(fp) > 227 * 1024) { set_last_error("fused cohort statistics: shared-memory plan does not fit"); return 1; }
  CUtensorMap mb;
  {
    const uint64_t bdims[2] = {static_cast<uint64_t>(2 * dp), static_cast<uint64_t>(c_pad)};
    const uint64_t str[1] = {static_cast<uint64_t>(2 * dp) * 2};
    const uint32_t box[2] = {64u, 128u};
    if (encode_tmap(&mb, 1, h->d_b2, 2, bdims, str, box, 128)) return 1;
  }
  API_CUDA(launch_asnorm_fused(fp, mb, h->sms, st));
  h->launches += 2;
  // rows the fused kernel could not finish (threshold bin outside its histogram, or overfull): usually none
  int n_flag = 0;
  API_CUDA(cudaMemcpyAsync(&n_flag, h->d_flag, 4, cudaMemcpyDeviceToHost, st));
  API_CUDA(cudaStreamSynchronize(st));
  h->fused_rows = n - n_flag; h->fallback_rows = n_flag;
  if (n_flag > 0) {
    const int w = vals_dev ? topk : 2;
    if (grow_buf(&h->d_fx, &h->fx_bytes, static_cast<size_t>(n_flag) * d * 4)) return 1;
    if (grow_buf(&h->d_fr, &h->fr_bytes, static_cast<size_t>(n_flag) * (w + 2) * 4)) return 1;
    API_CUDA(launch_gather_rows(test_dev, h->d_flag + 1, n_flag, d, h->d_fx, st));
    float* fm = h->d_fr; float* fs = h->d_fr + n_flag; float* fv = h->d_fr + 2 * static_cast<size_t>(n_flag);
    const long long l0 = h->launches;
    if (cohort_pass_unfused(h, h->d_fx, n_flag, cohort_dev, c, d, topk, mean_dev ? fm : nullptr, std_dev ? fs : nullptr,
                            vals_dev ? fv : nullptr, topk, st)) return 1;
    h->launches += l0;
    if (mean_dev) API_CUDA(launch_scatter_rows(fm, h->d_flag + 1, n_flag, 1, mean_dev, 1, st));
    if (std_dev) API_CUDA(launch_scatter_rows(fs, h->d_flag + 1, n_flag, 1, std_dev, 1, st));
    if (vals_dev) API_CUDA(launch_scatter_rows(fv, h->d_flag + 1, n_flag, topk, vals_dev, topk, st));
    h->launches += 4;
  }
  return 0;
}

static int cohort_pass_unfused(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                               float* mean_dev, float* std_dev, float* vals_dev, int vals_ld, cudaStream_t st) {
  (void)vals_ld;
  h->launches = 0;
  if (n <= 0) return 0;
  const int K = 3 * d;
  const int kbox = (K % 64 == 0) ? 64 : (K % 32 == 0 ? 32 : 16);
  const int nkc = (K + kbox - 1) / kbox;
  int n_tile = 256;
  while (n_tile > 16 && n_tile / 2 >= c) n_tile /= 2;
  const int c_pad = (c + n_tile - 1) / n_tile * n_tile;
  // rows per pass: a whole number of waves of (128 x n_tile) tiles over the SMs, near 4096 rows (the fp32 score block then
  // stays close to the L2 size); SVX_SCORE_BLOCK_ROWS overrides (debug)
  static const int env_rows = dbg_env("SVX_SCORE_BLOCK_ROWS") ? atoi(dbg_env("SVX_SCORE_BLOCK_ROWS")) : 0;
  int block_rows = 4096;
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int n_tiles_ = c_pad / n_tile;
    int best = 4096; double best_eff = 0.0;
    for (int mt = 24; mt <= 40; ++mt) {
      const long long tiles = static_cast<long long>(mt) * n_tiles_;
      const double eff = static_cast<double>(tiles) / (static_cast<double>((tiles + sms - 1) / sms) * sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = mt * 128; }
    }
    block_rows = best;
  }
  if (env_rows >= 128) block_rows = env_rows / 128 * 128;
  if (grow_buf(&h->d_a, &h->a_bytes, static_cast<size_t>(block_rows) * K * 2)) return 1;
  if (grow_buf(&h->d_b, &h->b_bytes, static_cast<size_t>(c_pad) * K * 2)) return 1;
  if (grow_buf(&h->d_s, &h->s_bytes, static_cast<size_t>(block_rows) * c_pad * 4)) return 1;
  API_CUDA(launch_split3(cohort_dev, h->d_b, c, c_pad, d, 1, st));
  ++h->launches;

  UmmaConvParams up;
  memset(&up, 0, sizeof up);
  up.out_W = 1; up.out_Wp = 1; up.w_box = 1; up.h_box = 128; up.w_tiles = 1;
  up.taps = 1; up.nkc = nkc; up.kbox = kbox; up.n_tile = n_tile; up.n_tiles = c_pad / n_tile;
  const int sw_bytes = kbox * 2;
  up.layout_type = sw_bytes == 128 ? 2u : sw_bytes == 64 ? 4u : 6u;
  up.sbo = 8u * sw_bytes;
  up.idesc = ptx::make_idesc_f16(1u, 128u, static_cast<uint32_t>(n_tile));
  up.a_stage_bytes = 128u * sw_bytes;
  up.b_stage_bytes = static_cast<uint32_t>((n_tile * sw_bytes + 1023) / 1024 * 1024);
  if (!conv_umma_finish_params(up)) { set_last_error("scoring GEMM tile does not fit"); return 1; }
  up.epi.n_valid = c_pad; up.epi.n_split = c_pad; up.epi.out_f32 = h->d_s; up.epi.ldf = c_pad;
  AMaps am; CUtensorMap bm; CUtensorMap auxm; OMaps om;
  memset(&auxm, 0, sizeof auxm);
  memset(&om, 0, sizeof om);
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(K), 1, static_cast<uint64_t>(block_rows)};
    const uint64_t str[2] = {static_cast<uint64_t>(K) * 2, static_cast<uint64_t>(K) * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(kbox), 1, 128};
    if (encode_tmap(&am.m[0], 1, h->d_a, 3, dims, str, box, sw_bytes)) return 1;
    for (int i = 1; i < 4; ++i) am.m[i] = am.m[0];
    const uint64_t bdims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(c_pad)};
    const uint64_t bstr[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t bbox[2] = {static_cast<uint32_t>(kbox), static_cast<uint32_t>(n_tile)};
    if (encode_tmap(&bm, 1, h->d_b, 2, bdims, bstr, bbox, sw_bytes)) return 1;
  }
  for (int64_t r0 = 0; r0 < n; r0 += block_rows) {
    const int rows = static_cast<int>(n - r0 < block_rows ? n - r0 : block_rows);
    const int rows_pad = (rows + 127) / 128 * 128;
    API_CUDA(launch_split3(test_dev + r0 * d, h->d_a, rows, rows_pad, d, 0, st));
    up.out_rows = rows;
    API_CUDA(launch_conv_umma(up, am, bm, auxm, om, 1, st));
    API_CUDA(launch_topk_stats(h->d_s, c_pad, rows, c, topk, mean_dev ? mean_dev + r0 : nullptr, std_dev ? std_dev + r0 : nullptr,
                               vals_dev ? vals_dev + r0 * topk : nullptr, topk, st));
    h->launches += 3;
  }
  return 0;
}

extern "C" {

int svx_trial_scores(const float* emb_dev, int d, const int32_t* idx1_dev, const int32_t* idx2_dev, int64_t n_trials,
                     const float* mean_dev, const float* std_dev, float* cos_dev, float* snorm_dev, void* cuda_stream) {
  if (!emb_dev || !idx1_dev || !idx2_dev || !cos_dev) { set_last_error("null argument"); return 1; }
  if (snorm_dev && (!mean_dev || !std_dev)) { set_last_error("snorm requested without cohort statistics"); return 1; }
  API_CUDA(launch_trial_scores(emb_dev, d, idx1_dev, idx2_dev, n_trials, mean_dev, std_dev, cos_dev, snorm_dev,
                               static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

long long svx_scorer_last_launches(svx_scorer* h) { return h ? h->launches : -1; }

int svx_eer_min_dcf(const float* scores_dev, const int32_t* labels_dev, int64_t n, double c_miss, double c_fa, double p_target,
                    double* out_host, void* cuda_stream) {
  if (!scores_dev || !labels_dev || !out_host) { set_last_error("null argument"); return 1; }
  if (n < 2 || n > 0x7fffffff) { set_last_error("EER needs between 2 and 2^31-1 trials"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  size_t bytes = 0;
  API_CUDA(eer_min_dcf(scores_dev, labels_dev, n, c_miss, c_fa, p_target, nullptr, nullptr, &bytes, st));
  void* ws = nullptr; double* d_out = nullptr;
  API_CUDA(cudaMallocAsync(&ws, bytes, st));
  API_CUDA(cudaMallocAsync(&d_out, 4 * sizeof(double), st));
  cudaError_t e = eer_min_dcf(scores_dev, labels_dev, n, c_miss, c_fa, p_target, d_out, ws, &bytes, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(ws, st); cudaFreeAsync(d_out, st);
  API_CUDA(e);
  API_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
