// EER / minDCF on the device (reference tensorflow/eer_minDCF.py:41-64, which builds the ROC with
// sklearn.metrics.roc_curve).  The same curve, point for point:
//   sort scores descending (CUB radix sort — library code, this is not a hot path) -> cumulative positives (CUB scan) -> one
//   point per distinct score (CUB select) -> drop_intermediate (a point stays if it is first, last or a corner: a second
//   difference of fps or tps is non-zero) -> origin prepended with threshold +inf -> fpr = fps / N, fnr = 1 - tps / P in fp64.
// EER = fpr at the FIRST minimum of |fnr - fpr| (np.nanargmin), minDCF = first minimum of c_miss*fnr*p_t + c_fa*fpr*(1-p_t),
// divided by min(c_miss*p_t, c_fa*(1-p_t)).  The reductions carry (value, index) pairs so that ties resolve like NumPy's.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace svx {

namespace {

struct Best { double v; long long i; };
__device__ __forceinline__ Best better(const Best& a, const Best& b) { return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

__global__ void distinct_flags_kernel(const float* __restrict__ s, long long n, uint8_t* __restrict__ flag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == n - 1 || s[i] != s[i + 1]) ? 1 : 0;
}

// One block: min over the kept ROC points (origin included) of both objectives.  idx[j] = position of distinct point j in the
// sorted arrays, tps_cum = inclusive positive counts over the sorted labels.
__global__ void __launch_bounds__(1024) roc_reduce_kernel(const float* __restrict__ s, const int* __restrict__ tps_cum, const long long* __restrict__ idx,
                                                          const int* __restrict__ m_ptr, long long n, double c_miss, double c_fa, double p_target,
                                                          double* __restrict__ out) {
  const long long m = *m_ptr;
  const double P = static_cast<double>(tps_cum[n - 1]);
  const double N = static_cast<double>(n) - P;
  __shared__ Best sh_e[1024], sh_d[1024];
  // origin: fpr = 0, fnr = 1
  Best be{fabs(1.0 - 0.0), -1}, bd{__dadd_rn(__dmul_rn(__dmul_rn(c_miss, 1.0), p_target), __dmul_rn(__dmul_rn(c_fa, 0.0), 1.0 - p_target)), -1};
  for (long long j = threadIdx.x; j < m; j += blockDim.x) {
    const long long ij = idx[j];
    const long long tp = tps_cum[ij], fp = 1 + ij - tp;
    bool keep = (j == 0 || j == m - 1 || m <= 2);
    if (!keep) {
      const long long ia = idx[j - 1], ib = idx[j + 1];
      const long long tpa = tps_cum[ia], tpb = tps_cum[ib];
      const long long fpa = 1 + ia - tpa, fpb = 1 + ib - tpb;
      keep = (fpb - 2 * fp + fpa) != 0 || (tpb - 2 * tp + tpa) != 0;
    }
    if (!keep) continue;
    const double fpr = static_cast<double>(fp) / N, fnr = 1.0 - static_cast<double>(tp) / P;
    be = better(be, Best{fabs(fnr - fpr), j});
    bd = better(bd, Best{__dadd_rn(__dmul_rn(__dmul_rn(c_miss, fnr), p_target), __dmul_rn(__dmul_rn(c_fa, fpr), 1.0 - p_target)), j});
  }
  sh_e[threadIdx.x] = be; sh_d[threadIdx.x] = bd;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh_e[threadIdx.x] = better(sh_e[threadIdx.x], sh_e[threadIdx.x + o]); sh_d[threadIdx.x] = better(sh_d[threadIdx.x], sh_d[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const Best e = sh_e[0], d = sh_d[0];
    double eer = 0.0, eer_thr = INFINITY;
    if (e.i >= 0) { const long long ij = idx[e.i]; eer = static_cast<double>(1 + ij - tps_cum[ij]) / N; eer_thr = static_cast<double>(s[ij]); }
    const double c_def = fmin(c_miss * p_target, c_fa * (1.0 - p_target));
    out[0] = eer; out[1] = eer_thr; out[2] = d.v / c_def; out[3] = d.i >= 0 ? static_cast<double>(s[idx[d.i]]) : INFINITY;
  }
}

}  // namespace

// workspace: returns bytes needed when ws == nullptr
cudaError_t eer_min_dcf(const float* scores, const int32_t* labels, long long n, double c_miss, double c_fa, double p_target, double* out_dev,
                        void* ws, size_t* ws_bytes, cudaStream_t st) {
  const size_t a = 256;
  auto up = [&](size_t v) { return (v + a - 1) / a * a; };
  size_t sort_b = 0, scan_b = 0, sel_b = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_b, scores, static_cast<float*>(nullptr), labels, static_cast<int*>(nullptr), static_cast<int>(n), 0, 32, st);
  cub::DeviceScan::InclusiveSum(nullptr, scan_b, static_cast<int*>(nullptr), static_cast<int*>(nullptr), static_cast<int>(n), st);
  cub::CountingInputIterator<long long> counting(0);
  cub::DeviceSelect::Flagged(nullptr, sel_b, counting, static_cast<uint8_t*>(nullptr), static_cast<long long*>(nullptr), static_cast<int*>(nullptr),
                             static_cast<int>(n), st);
  const size_t tmp_b = up(std::max(sort_b, std::max(scan_b, sel_b)));
  const size_t need = tmp_b + up(n * 4) * 3 + up(n) + up(n * 8) + a;
  if (!ws) { *ws_bytes = need; return cudaSuccess; }
  if (*ws_bytes < need) return cudaErrorInvalidValue;
  uint8_t* p = static_cast<uint8_t*>(ws);
  void* tmp = p; p += tmp_b;
  float* s_sorted = reinterpret_cast<float*>(p); p += up(n * 4);
  int* l_sorted = reinterpret_cast<int*>(p); p += up(n * 4);
  int* tps_cum = reinterpret_cast<int*>(p); p += up(n * 4);
  uint8_t* flag = p; p += up(n);
  long long* idx = reinterpret_cast<long long*>(p); p += up(n * 8);
  int* m_ptr = reinterpret_cast<int*>(p);
  size_t b = sort_b;
  cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(tmp, b, scores, s_sorted, labels, l_sorted, static_cast<int>(n), 0, 32, st);
  if (e != cudaSuccess) return e;
  b = scan_b;
  e = cub::DeviceScan::InclusiveSum(tmp, b, l_sorted, tps_cum, static_cast<int>(n), st);
  if (e != cudaSuccess) return e;
  distinct_flags_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(s_sorted, n, flag);
  b = sel_b;
  e = cub::DeviceSelect::Flagged(tmp, b, counting, flag, idx, m_ptr, static_cast<int>(n), st);
  if (e != cudaSuccess) return e;
  roc_reduce_kernel<<<1, 1024, 0, st>>>(s_sorted, tps_cum, idx, m_ptr, n, c_miss, c_fa, p_target, out_dev);
  return cudaGetLastError();
}

}  // namespace svx
