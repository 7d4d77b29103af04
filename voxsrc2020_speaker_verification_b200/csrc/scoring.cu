// Scoring kernels (reference tensorflow/snorm.py): L2 normalisation, split-bf16 operand preparation for the
// cohort GEMM, exact per-row top-k mean/std, trial gather + adaptive symmetric normalisation, cohort means.
#include <cstdlib>

#include "kernels.cuh"

namespace svx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// x / ||x||_2 per row (snorm.py:23-25,32).  One warp per row.
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* in, float* out, long long n, int d) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = in + row * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) { const float v = x[i]; ss += v * v; }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  float* o = out + row * d;
  for (int i = lane; i < d; i += 32) o[i] = x[i] / nrm;
}

cudaError_t launch_l2norm_rows(const float* in, float* out, long long n, int d, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long blocks = (n * 32 + 255) / 256;
  l2norm_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, out, n, d);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fp32 → split bf16 with the three partial products laid out along K so that ONE bf16 GEMM of depth 3d gives
//   hi(x)·hi(c) + lo(x)·hi(c) + hi(x)·lo(c)   (~2^-16 relative error per product, fp32 accumulation in TMEM):
//   test side   row = [hi | lo | hi],  cohort side row = [hi | hi | lo].  Rows n..n_pad-1 are zero.
__global__ void __launch_bounds__(256) split3_kernel(const float* in, __nv_bfloat16* out, long long n, long long n_pad, int d,
                                                     int cohort_side) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_pad * d) return;
  const long long row = idx / d;
  const int j = static_cast<int>(idx - row * d);
  float v = row < n ? in[idx] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* o = out + row * (3LL * d);
  o[j] = hi;
  o[d + j] = cohort_side ? hi : lo;
  o[2 * d + j] = cohort_side ? lo : hi;
}

cudaError_t launch_split3(const float* in, __nv_bfloat16* out, long long n, long long n_pad, int d, int cohort_side, cudaStream_t st) {
  const long long total = n_pad * d;
  if (total <= 0) return cudaSuccess;
  split3_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, n, n_pad, d, cohort_side);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Per-row top-k statistics (snorm.py:100-109): mean and population std of the k largest of c scores, exact for
// ties (only the VALUES of the top-k multiset matter).  One CTA per row:
//   1. row → shared memory, block min/max
//   2. 2048 linear buckets over [min,max] (monotone → exact partition), shared-memory histogram, suffix scan →
//      bucket B* that holds the k-th largest, r = how many of B*'s members are needed
//   3. 4-pass 8-bit radix select restricted to B* → exact threshold key t and number of copies of t needed
//   4. two reduction passes: mean, then sum of squared deviations
constexpr int kTopkThreads = 256;
constexpr int kBuckets = 2048;

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < kTopkThreads / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) { t = warp_sum(t); if (l == 0) red[0] = t; }
  __syncthreads();
  return red[0];
}

// Optionally also emits the selected values (unordered, padded with -1e30 up to topk) so that per-shard
// candidates can be merged across cohort shards: top-k of the union of per-shard top-k lists is the global top-k.
constexpr float kPadValue = -1e30f;

__global__ void __launch_bounds__(kTopkThreads) topk_stats_kernel(const float* scores, int ld, int c, int topk, float* mean_out,
                                                                  float* std_out, float* vals_out, int vals_ld) {
  extern __shared__ float vals[];                    // [c]
  __shared__ int hist[kBuckets];
  __shared__ float red[32];
  __shared__ int ired[kTopkThreads / 32];
  __shared__ int sh_bstar, sh_r, sh_above;
  __shared__ uint32_t sh_prefix;
  __shared__ int sh_krem;
  __shared__ int sh_emit;
  const long long row = blockIdx.x;
  const float* src = scores + row * ld;
  const int tid = threadIdx.x;
  const int k = min(topk, c);

  float vmin = INFINITY, vmax = -INFINITY;
  for (int i = tid; i < c; i += kTopkThreads) {
    const float v = src[i];
    vals[i] = v;
    vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  for (int i = tid; i < kBuckets; i += kTopkThreads) hist[i] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = vmin; red[8 + (tid >> 5)] = vmax; }
  __syncthreads();
  vmin = red[0]; vmax = red[8];
#pragma unroll
  for (int w = 1; w < kTopkThreads / 32; ++w) { vmin = fminf(vmin, red[w]); vmax = fmaxf(vmax, red[8 + w]); }
  const float bscale = vmax > vmin ? (static_cast<float>(kBuckets) - 0.5f) / (vmax - vmin) : 0.f;
  auto bucket = [&](float v) -> int { return min(kBuckets - 1, static_cast<int>((v - vmin) * bscale)); };

  for (int i = tid; i < c; i += kTopkThreads) atomicAdd(&hist[bucket(vals[i])], 1);
  __syncthreads();

  // suffix scan over buckets, highest first: thread t owns buckets [hi-7, hi], hi = kBuckets-1-8t
  {
    const int hi = kBuckets - 1 - 8 * tid;
    int part = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) part += hist[hi - j];
    int inc = part;                                  // inclusive scan over threads (ascending tid = descending buckets)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) ired[tid >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (tid >> 5); ++w) woff += ired[w];
    inc += woff;
    const int exc = inc - part;
    if (exc < k && inc >= k) {                        // exactly one thread
      int cum = exc;
      for (int j = 0; j < 8; ++j) {
        const int h = hist[hi - j];
        if (cum + h >= k) { sh_bstar = hi - j; sh_above = cum; sh_r = k - cum; break; }
        cum += h;
      }
    }
    __syncthreads();
  }
  const int bstar = sh_bstar, r = sh_r;

  // radix select of the r-th largest key among members of bucket bstar
  if (tid == 0) { sh_prefix = 0u; sh_krem = r; }
  uint32_t mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    __syncthreads();
    for (int i = tid; i < 256; i += kTopkThreads) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = sh_prefix;
    for (int i = tid; i < c; i += kTopkThreads) {
      const float v = vals[i];
      if (bucket(v) == bstar) {
        const uint32_t key = f2key(v);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
      }
    }
    __syncthreads();
    if (tid < 32) {                                  // warp 0: lane l owns digits [hi-7, hi], hi = 255-8l
      const int hi = 255 - 8 * tid;
      int part = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) part += hist[hi - j];
      int inc = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
      const int exc = inc - part;
      const int krem = sh_krem;
      if (exc < krem && inc >= krem) {
        int cum = exc;
        for (int j = 0; j < 8; ++j) {
          const int h = hist[hi - j];
          if (cum + h >= krem) {
            sh_prefix = prefix | (static_cast<uint32_t>(hi - j) << shift);
            sh_krem = krem - cum;
            break;
          }
          cum += h;
        }
      }
    }
    mask |= 0xFFu << shift;
  }
  __syncthreads();
  const uint32_t tkey = sh_prefix;
  const float tval = key2f(tkey);
  const float ties = static_cast<float>(sh_krem);

  if (vals_out) {
    if (tid == 0) sh_emit = 0;
    __syncthreads();
    float* vo = vals_out + row * vals_ld;
    for (int i = tid; i < c; i += kTopkThreads) {
      const float v = vals[i];
      const int b = bucket(v);
      if (b > bstar || (b == bstar && f2key(v) > tkey)) vo[atomicAdd(&sh_emit, 1)] = v;
    }
    __syncthreads();
    const int base = sh_emit;                       // = k - ties
    for (int i = base + tid; i < topk; i += kTopkThreads) vo[i] = (i < k) ? tval : kPadValue;
  }
  float s = 0.f;
  for (int i = tid; i < c; i += kTopkThreads) {
    const float v = vals[i];
    const int b = bucket(v);
    if (b > bstar || (b == bstar && f2key(v) > tkey)) s += v;
  }
  const float total = block_sum(s, red) + ties * tval;
  const float mean = total / static_cast<float>(k);
  float q = 0.f;
  for (int i = tid; i < c; i += kTopkThreads) {
    const float v = vals[i];
    const int b = bucket(v);
    if (b > bstar || (b == bstar && f2key(v) > tkey)) { const float d = v - mean; q += d * d; }
  }
  const float ss = block_sum(q, red) + ties * (tval - mean) * (tval - mean);
  if (tid == 0 && mean_out) {
    mean_out[row] = mean;
    std_out[row] = sqrtf(ss / static_cast<float>(k));
  }
}

// Register-resident form of the same algorithm for c <= 24*256 (the VoxCeleb2 cohort has 5994 speakers): every thread keeps
// its 24 scores and their bucket ids in registers, so the row is read once (coalesced) and the select / reduce passes touch
// no memory.  The generic kernel above spends ~3000 instructions per thread re-reading shared memory and recomputing buckets;
// this one ~800 (profiles/r01_score_launches.txt).  Same exact-selection logic: 2048 monotone buckets, 8-bit radix select
// inside the bucket of the k-th largest, ties counted.
// kVPT values per thread: instantiated for 4 / 8 / 12 / 24 so that a cohort shard (or a merged candidate list) pays for its own size
template <int kVPT>
__global__ void __launch_bounds__(kTopkThreads, 3) topk_stats_reg_kernel(const float* __restrict__ scores, int ld, int c, int topk,
                                                                      float* mean_out, float* std_out, float* vals_out, int vals_ld) {
  __shared__ int hist[kBuckets];
  __shared__ float red[32];
  __shared__ int ired[kTopkThreads / 32];
  __shared__ int sh_bstar, sh_r;
  __shared__ uint32_t sh_prefix;
  __shared__ int sh_krem;
  __shared__ int sh_emit;
  const long long row = blockIdx.x;
  const float* src = scores + row * ld;
  const int tid = threadIdx.x;
  const int k = min(topk, c);

  float v[kVPT];
  float vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
  for (int j = 0; j < kVPT; ++j) {
    const int i = tid + j * kTopkThreads;
    v[j] = i < c ? src[i] : -INFINITY;                 // -inf: never selected, never the minimum of real values' range below
    if (i < c) { vmin = fminf(vmin, v[j]); vmax = fmaxf(vmax, v[j]); }
  }
  for (int i = tid; i < kBuckets; i += kTopkThreads) hist[i] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = vmin; red[8 + (tid >> 5)] = vmax; }
  __syncthreads();
  vmin = red[0]; vmax = red[8];
#pragma unroll
  for (int w = 1; w < kTopkThreads / 32; ++w) { vmin = fminf(vmin, red[w]); vmax = fmaxf(vmax, red[8 + w]); }
  const float bscale = vmax > vmin ? (static_cast<float>(kBuckets) - 0.5f) / (vmax - vmin) : 0.f;

  uint32_t bk[kVPT / 2];                               // bucket ids, two per register; 0xffff = padding element
#pragma unroll
  for (int j = 0; j < kVPT; ++j) {
    const int i = tid + j * kTopkThreads;
    uint32_t b = 0xffffu;
    if (i < c) {
      b = static_cast<uint32_t>(min(kBuckets - 1, static_cast<int>((v[j] - vmin) * bscale)));
      atomicAdd(&hist[b], 1);
    }
    if (j & 1) bk[j >> 1] |= b << 16; else bk[j >> 1] = b;
  }
  __syncthreads();
  {   // suffix scan over buckets, highest first: thread t owns buckets [hi-7, hi], hi = kBuckets-1-8t
    const int hi = kBuckets - 1 - 8 * tid;
    int part = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) part += hist[hi - j];
    int inc = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) ired[tid >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (tid >> 5); ++w) woff += ired[w];
    inc += woff;
    const int exc = inc - part;
    if (exc < k && inc >= k) {
      int cum = exc;
      for (int j = 0; j < 8; ++j) {
        const int h = hist[hi - j];
        if (cum + h >= k) { sh_bstar = hi - j; sh_r = k - cum; break; }
        cum += h;
      }
    }
    __syncthreads();
  }
  const uint32_t bstar = static_cast<uint32_t>(sh_bstar);
  // which of this thread's values sit in the bucket of the k-th largest (usually none)
  uint32_t in_star = 0;
#pragma unroll
  for (int j = 0; j < kVPT; ++j) {
    const uint32_t b = (j & 1) ? (bk[j >> 1] >> 16) : (bk[j >> 1] & 0xffffu);
    if (b == bstar) in_star |= 1u << j;
  }
  if (tid == 0) { sh_prefix = 0u; sh_krem = sh_r; }
  uint32_t mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    __syncthreads();
    hist[tid] = 0;                                     // 256 digits, 256 threads
    __syncthreads();
    const uint32_t prefix = sh_prefix;
    if (in_star) {
#pragma unroll
      for (int j = 0; j < kVPT; ++j)
        if ((in_star >> j) & 1u) {
          const uint32_t key = f2key(v[j]);
          if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
        }
    }
    __syncthreads();
    if (tid < 32) {
      const int hi = 255 - 8 * tid;
      int part = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) part += hist[hi - j];
      int inc = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
      const int exc = inc - part;
      const int krem = sh_krem;
      if (exc < krem && inc >= krem) {
        int cum = exc;
        for (int j = 0; j < 8; ++j) {
          const int h = hist[hi - j];
          if (cum + h >= krem) {
            sh_prefix = prefix | (static_cast<uint32_t>(hi - j) << shift);
            sh_krem = krem - cum;
            break;
          }
          cum += h;
        }
      }
    }
    mask |= 0xFFu << shift;
  }
  __syncthreads();
  const uint32_t tkey = sh_prefix;
  const float tval = key2f(tkey);
  const float ties = static_cast<float>(sh_krem);
  // selected = strictly above the threshold key (buckets are monotone in the value, so this equals the bucket test of the
  // generic kernel); the threshold value itself contributes `ties` copies
  if (vals_out) {
    if (tid == 0) sh_emit = 0;
    __syncthreads();
    float* vo = vals_out + row * vals_ld;
#pragma unroll
    for (int j = 0; j < kVPT; ++j)
      if (tid + j * kTopkThreads < c && f2key(v[j]) > tkey) vo[atomicAdd(&sh_emit, 1)] = v[j];
    __syncthreads();
    const int base = sh_emit;
    for (int i = base + tid; i < topk; i += kTopkThreads) vo[i] = (i < k) ? tval : kPadValue;
  }
  float ssum = 0.f;
#pragma unroll
  for (int j = 0; j < kVPT; ++j)
    if (tid + j * kTopkThreads < c && f2key(v[j]) > tkey) ssum += v[j];
  const float total = block_sum(ssum, red) + ties * tval;
  const float mean = total / static_cast<float>(k);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kVPT; ++j)
    if (tid + j * kTopkThreads < c && f2key(v[j]) > tkey) { const float d = v[j] - mean; q += d * d; }
  const float ss = block_sum(q, red) + ties * (tval - mean) * (tval - mean);
  if (tid == 0 && mean_out) {
    mean_out[row] = mean;
    std_out[row] = sqrtf(ss / static_cast<float>(k));
  }
}

cudaError_t launch_topk_stats(const float* scores, int ld, long long n_rows, int c, int topk, float* mean, float* stdv,
                              float* vals_out, int vals_ld, cudaStream_t st) {
  if (n_rows <= 0) return cudaSuccess;
  static const bool no_reg = dbg_env("SVX_TOPK_GENERIC") != nullptr;   // debug switch
  if (c <= 24 * kTopkThreads && !no_reg) {
    const unsigned g = static_cast<unsigned>(n_rows);
    if (c <= 4 * kTopkThreads) topk_stats_reg_kernel<4><<<g, kTopkThreads, 0, st>>>(scores, ld, c, topk, mean, stdv, vals_out, vals_ld);
    else if (c <= 8 * kTopkThreads) topk_stats_reg_kernel<8><<<g, kTopkThreads, 0, st>>>(scores, ld, c, topk, mean, stdv, vals_out, vals_ld);
    else if (c <= 12 * kTopkThreads) topk_stats_reg_kernel<12><<<g, kTopkThreads, 0, st>>>(scores, ld, c, topk, mean, stdv, vals_out, vals_ld);
    else topk_stats_reg_kernel<24><<<g, kTopkThreads, 0, st>>>(scores, ld, c, topk, mean, stdv, vals_out, vals_ld);
    return cudaGetLastError();
  }
  const size_t smem = static_cast<size_t>(c) * sizeof(float);
  if (smem > 48 * 1024) {   // the attribute is per device and the call is cheap: set it on whatever device is current
    cudaError_t e = cudaFuncSetAttribute(topk_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  topk_stats_kernel<<<static_cast<unsigned>(n_rows), kTopkThreads, smem, st>>>(scores, ld, c, topk, mean, stdv, vals_out, vals_ld);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Trial gather (snorm.py:113-131): cos = <x[i1], x[i2]> in fp32, asnorm = 0.5*((cos-m1)/s1 + (cos-m2)/s2).
// One warp per trial, 128-bit loads; the embedding table (≤150 MB) mostly lives in L2.
__global__ void __launch_bounds__(256) trial_scores_kernel(const float* __restrict__ emb, int d, const int32_t* __restrict__ idx1,
                                                           const int32_t* __restrict__ idx2, long long t,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv,
                                                           float* cos_out, float* snorm_out) {
  const long long tr = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tr >= t) return;
  const int a = idx1[tr], b = idx2[tr];
  const float* xa = emb + static_cast<size_t>(a) * d;
  const float* xb = emb + static_cast<size_t>(b) * d;
  float acc = 0.f;
  if ((d & 3) == 0) {
    const float4* pa = reinterpret_cast<const float4*>(xa);
    const float4* pb = reinterpret_cast<const float4*>(xb);
    for (int i = lane; i < (d >> 2); i += 32) {
      const float4 u = pa[i], v = pb[i];
      acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
    }
  } else {
    for (int i = lane; i < d; i += 32) acc += xa[i] * xb[i];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    cos_out[tr] = acc;
    if (snorm_out) snorm_out[tr] = 0.5f * ((acc - mean[a]) / stdv[a] + (acc - mean[b]) / stdv[b]);
  }
}

cudaError_t launch_trial_scores(const float* emb, int d, const int32_t* idx1, const int32_t* idx2, long long t, const float* mean,
                                const float* stdv, float* cos_out, float* snorm_out, cudaStream_t st) {
  if (t <= 0) return cudaSuccess;
  const long long blocks = (t * 32 + 255) / 256;
  trial_scores_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(emb, d, idx1, idx2, t, mean, stdv, cos_out, snorm_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Cohort / enrolment models (snorm.py:45-67): mean of the unit vectors of each group, not re-normalised.
// Groups arrive as a CSR list (member_rows[group_off[g] .. group_off[g+1]) = rows of group g, in the order the reference's
// dict iteration appends them).  One thread per (group, column) adds its members in that order and divides by the count:
// the same sequential fp32 sum + true divide np.mean(matrix, axis=0) performs, so the result is run-to-run deterministic
// and bit-identical to the reference (no atomics).  Consecutive threads read consecutive columns of a row: coalesced.
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ unit_rows, int d, const int32_t* __restrict__ member_rows,
                                                         const int32_t* __restrict__ group_off, float* out, int n_groups) {
  const int g = blockIdx.x;
  if (g >= n_groups) return;
  const int lo = group_off[g], hi = group_off[g + 1];
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float acc = 0.f;
    for (int i = lo; i < hi; ++i) acc = __fadd_rn(acc, unit_rows[static_cast<size_t>(member_rows[i]) * d + col]);
    out[static_cast<size_t>(g) * d + col] = hi > lo ? __fdiv_rn(acc, static_cast<float>(hi - lo)) : 0.f;
  }
}

cudaError_t launch_group_mean(const float* unit_rows, int d, const int32_t* member_rows, const int32_t* group_off, float* out, int n_groups,
                              cudaStream_t st) {
  if (n_groups <= 0) return cudaSuccess;
  group_mean_kernel<<<static_cast<unsigned>(n_groups), d >= 256 ? 256 : (d > 32 ? (d + 31) / 32 * 32 : 32), 0, st>>>(unit_rows, d, member_rows,
                                                                                                              group_off, out, n_groups);
  return cudaGetLastError();
}

}  // namespace svx
