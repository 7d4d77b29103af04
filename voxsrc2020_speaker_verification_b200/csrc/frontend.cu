// Feature front-end on the device: what the reference gets from the Kaldi pipe
// `apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300` in front of the network (tf_extract.py:63).
//
// Sliding-window mean normalisation [ext: Kaldi SlidingWindowCmn]: frame t is centred on the mean of a window of up to
// `cmn_window` frames around t, shifted to stay inside [0, T); sums in double precision like Kaldi.  The window rule (the
// oracle restates it frame by frame, oracle/cmn_oracle.py):
//   center:     [t - W/2, t - W/2 + W);   otherwise: [t - W, t + 1)
//   start < 0 → both ends move right by -start
//   !center and end > t → end = max(t + 1, min_window)        (the first frames look ahead until min_window frames are seen)
//   end > T → both ends move left by end - T, start clamped at 0  Two kernels over the
// packed [total_frames, F] matrix: per-utterance, per-bin running sums (one thread per bin walks the frames; consecutive
// threads read consecutive bins, so every step is one coalesced row), then one thread per element subtracts the window mean.
#include "kernels.cuh"

namespace svx {

__global__ void __launch_bounds__(128) cmn_prefix_kernel(const float* __restrict__ feats, const int32_t* __restrict__ frame_off, double* csum,
                                                         int F) {
  const int u = blockIdx.x;
  const int f = threadIdx.x;
  if (f >= F) return;
  const int t0 = frame_off[u], t1 = frame_off[u + 1];
  // csum has one extra row per utterance: rows [t0 + u, t1 + u + 1)
  double* c = csum + (static_cast<size_t>(t0) + u) * F + f;
  const float* x = feats + static_cast<size_t>(t0) * F + f;
  double acc = 0.0;
  c[0] = 0.0;
  for (int t = 0; t < t1 - t0; ++t) {
    acc += static_cast<double>(x[static_cast<size_t>(t) * F]);
    c[static_cast<size_t>(t + 1) * F] = acc;
  }
}

__global__ void __launch_bounds__(256) cmn_apply_kernel(const float* __restrict__ feats, const int32_t* __restrict__ frame_off,
                                                        const int32_t* __restrict__ utt_of_frame, const double* __restrict__ csum, float* out,
                                                        long long total, int F, int window, int center, int min_window) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total * F) return;
  const long long row = idx / F;
  const int f = static_cast<int>(idx - row * F);
  const int u = utt_of_frame[row];
  const int t0 = frame_off[u], T = frame_off[u + 1] - t0;
  const int t = static_cast<int>(row - t0);
  int ws, we;
  if (center) { ws = t - window / 2; we = ws + window; } else { ws = t - window; we = t + 1; }
  if (ws < 0) { we -= ws; ws = 0; }
  if (!center && we > t) we = max(t + 1, min_window);
  if (we > T) { ws -= we - T; we = T; if (ws < 0) ws = 0; }
  const double* c = csum + (static_cast<size_t>(t0) + u) * F + f;
  const double mean = (c[static_cast<size_t>(we) * F] - c[static_cast<size_t>(ws) * F]) / static_cast<double>(we - ws);
  out[idx] = static_cast<float>(static_cast<double>(feats[idx]) - mean);
}

__global__ void utt_of_frame_kernel(int32_t* utt_of_frame, long long total, const int32_t* frame_off, int n_utts) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = n_utts - 1, ans = 0;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (frame_off[mid] <= i) { ans = mid; lo = mid + 1; } else { hi = mid - 1; }
  }
  utt_of_frame[i] = ans;
}

cudaError_t launch_cmn_sliding(const float* feats, float* out, const int32_t* frame_off_dev, int n_utts, long long total_frames, int F,
                               int window, int center, int min_window, double* csum_ws, int32_t* utt_ws, cudaStream_t st) {
  if (n_utts <= 0 || total_frames <= 0) return cudaSuccess;
  if (F > 128) return cudaErrorInvalidValue;
  utt_of_frame_kernel<<<static_cast<unsigned>((total_frames + 255) / 256), 256, 0, st>>>(utt_ws, total_frames, frame_off_dev, n_utts);
  cmn_prefix_kernel<<<n_utts, 128, 0, st>>>(feats, frame_off_dev, csum_ws, F);
  const long long n = total_frames * F;
  cmn_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(feats, frame_off_dev, utt_ws, csum_ws, out, total_frames, F, window,
                                                                          center, min_window);
  return cudaGetLastError();
}

}  // namespace svx

// ---------------------------------------------------------------------------------------------------------
// Kaldi CompressedMatrix ('CM ') decode on the device: the format `copy-feats --compress=true` writes for the prepared FBANK
// (reference prepare_data.sh:68-69) and kaldi_io._read_compressed_mat reads (kaldi_io.py:471-504).  Per record: global header
// {min f32, range f32, rows i32, cols i32}, per column 4 x uint16 percentiles (0, 25, 75, 100), then uint8 data column-major.
// The arithmetic follows kaldi_io.py operation by operation in fp32 without contraction, so the result is bit-identical.
namespace svx {

__global__ void __launch_bounds__(256) cm_decode_kernel(const uint8_t* __restrict__ blob, const long long* __restrict__ rec_off,
                                                        const int32_t* __restrict__ frame_off, const int32_t* __restrict__ utt_of_frame,
                                                        float* out, long long total, int cols, const int32_t* __restrict__ bad) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total * cols) return;
  const long long row = idx / cols;
  const int c = static_cast<int>(idx - row * cols);
  if (*bad != kCmNoBadRecord) return;                                // a record failed validation: nothing is decoded, the call reports it
  const int u = utt_of_frame[row];
  const int r = static_cast<int>(row - frame_off[u]);
  const int rows = frame_off[u + 1] - frame_off[u];
  const uint8_t* rec = blob + rec_off[u];
  float gmin, grange;
  memcpy(&gmin, rec, 4); memcpy(&grange, rec + 4, 4);
  const uint8_t* ch = rec + 16 + static_cast<size_t>(c) * 8;
  float p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned short q = static_cast<unsigned short>(ch[2 * i] | (ch[2 * i + 1] << 8));
    p[i] = __fadd_rn(__fmul_rn(__fmul_rn(static_cast<float>(q), grange), 1.52590218966964e-05f), gmin);     // kaldi_io.py:480-483
  }
  const uint8_t b = rec[16 + static_cast<size_t>(cols) * 8 + static_cast<size_t>(c) * rows + r];
  const float d = static_cast<float>(b);
  float v;
  if (b <= 64) v = __fadd_rn(p[0], __fmul_rn(__fdiv_rn(__fsub_rn(p[1], p[0]), 64.f), d));                        // kaldi_io.py:497-503
  else if (b > 192) v = __fadd_rn(p[2], __fmul_rn(__fdiv_rn(__fsub_rn(p[3], p[2]), 63.f), __fsub_rn(d, 192.f)));
  else v = __fadd_rn(p[1], __fmul_rn(__fdiv_rn(__fsub_rn(p[2], p[1]), 128.f), __fsub_rn(d, 64.f)));
  out[idx] = v;
}

// Every record's own header must agree with what the caller says about it: rows = its slice of frame_off, cols = the model's
// feature dimension, and the record must end inside the blob.  *bad starts at kCmNoBadRecord and receives the
// smallest offending record index.
__global__ void cm_check_kernel(const uint8_t* __restrict__ blob, const long long* __restrict__ rec_off, const int32_t* __restrict__ frame_off,
                                int n, int cols, long long blob_bytes, int32_t* bad) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const long long off = rec_off[u];
  const int rows = frame_off[u + 1] - frame_off[u];
  bool ok = off >= 0 && rows > 0 && off + 16 <= blob_bytes;
  if (ok) {
    int32_t hr, hc;
    memcpy(&hr, blob + off + 8, 4); memcpy(&hc, blob + off + 12, 4);
    ok = hr == rows && hc == cols && off + 16 + 8LL * cols + static_cast<long long>(rows) * cols <= blob_bytes;
  }
  if (!ok) atomicMin(bad, u);
}

cudaError_t launch_cm_decode(const uint8_t* blob, const long long* rec_off_dev, const int32_t* frame_off_dev, int n_utts, long long total_frames,
                             int cols, float* out, int32_t* utt_ws, long long blob_bytes, int32_t* bad_record, cudaStream_t st) {
  if (n_utts <= 0 || total_frames <= 0) return cudaSuccess;
  cm_check_kernel<<<(n_utts + 127) / 128, 128, 0, st>>>(blob, rec_off_dev, frame_off_dev, n_utts, cols, blob_bytes, bad_record);
  utt_of_frame_kernel<<<static_cast<unsigned>((total_frames + 255) / 256), 256, 0, st>>>(utt_ws, total_frames, frame_off_dev, n_utts);
  const long long n = total_frames * cols;
  cm_decode_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(blob, rec_off_dev, frame_off_dev, utt_ws, out, total_frames, cols, bad_record);
  return cudaGetLastError();
}

}  // namespace svx
