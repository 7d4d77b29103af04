// Bandwidth-bound kernels around the convolutions: input packing, Cin=1 stem conv, DPN pre-activation,
// Res2Net stride-2 average pool, statistics pooling, the embedding FC and the chunk combine.
// All use 128-bit accesses along the channel axis (NHWC, channels fastest).
#include "kernels.cuh"

namespace svx {

// ---------------------------------------------------------------------------------------------------------
// Row map of one stage in ONE launch, one warp per tall-image row: seg_of_row[row] = the segment the row belongs to (-1 = zero
// padding row; binary search over the segment offsets, the same in every lane), and pix_valid[p] for the flat pixel sequence
// p = row*Wp + col: 1 when the row belongs to a segment and col < W (0 on gap rows and on the zero column).
__global__ void __launch_bounds__(256) fill_row_map_kernel(int32_t* seg_of_row, uint8_t* pix_valid, int rows, const int32_t* seg_row_off,
                                                           const int32_t* seg_h, int n_seg, int W, int Wp) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  int lo = 0, hi = n_seg - 1, ans = -1;     // last segment whose offset <= row
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (seg_row_off[mid] <= row) { ans = mid; lo = mid + 1; } else { hi = mid - 1; }
  }
  int v = -1;
  if (ans >= 0 && row < seg_row_off[ans] + seg_h[ans]) v = ans;
  if (lane == 0) seg_of_row[row] = v;
  uint8_t* pr = pix_valid + static_cast<size_t>(row) * Wp;
  for (int col = lane; col < Wp; col += 32) pr[col] = (col < W && v >= 0) ? 1 : 0;
}

cudaError_t launch_fill_row_map(int32_t* seg_of_row, uint8_t* pix_valid, int rows, const int32_t* seg_row_off, const int32_t* seg_h, int n_seg,
                                int W, int Wp, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  fill_row_map_kernel<<<(rows + 7) / 8, 256, 0, st>>>(seg_of_row, pix_valid, rows, seg_row_off, seg_h, n_seg, W, Wp);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// TDNN input: fp32 [frames, F] per segment → tall image [rows, 1, Cpad] 16-bit (reference tf_extract.py:32 with
// expand_dim=2: the feature axis is the channel axis).  Padding rows and channels F..Cpad-1 are zero.
template <typename T>
__global__ void pack_input_kernel(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off,
                                  const int32_t* seg_of_row, T* out, int rows, int F, int Cpad) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(rows) * Cpad;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cpad);
  const int row = static_cast<int>(idx / Cpad);
  const int seg = seg_of_row[row];
  float v = 0.f;
  if (seg >= 0 && c < F) v = feats[(static_cast<size_t>(seg_frame_off[seg]) + (row - seg_row_off[seg])) * F + c];
  out[idx] = TypeOps<T>::from_f(v);
}

cudaError_t launch_pack_input(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off,
                              const int32_t* seg_of_row, void* out, int rows, int F, int Cpad, int is_bf16, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * Cpad;
  if (total <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16)
    pack_input_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(feats, seg_frame_off, seg_row_off, seg_of_row,
                                                              static_cast<__nv_bfloat16*>(out), rows, F, Cpad);
  else
    pack_input_kernel<__half><<<blocks, 256, 0, st>>>(feats, seg_frame_off, seg_row_off, seg_of_row,
                                                       static_cast<__half*>(out), rows, F, Cpad);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Stem: 3x3 conv, Cin = 1, stride 1, zero padding (1,1) in time and feature, then BN and ReLU
// (res2net_model.py:192-203, dpn_model.py:32-37).  Reads the fp32 features directly ([N,T,F,1] with expand_dim=3),
// writes the stage-0 tall image [rows, F, Cpad].
constexpr int kStemPix = 8;   // feature columns per thread

template <typename T>
__global__ void __launch_bounds__(256, 2) stem_conv_kernel(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off,
                                                        const int32_t* seg_h, const int32_t* seg_of_row, const float* w9,
                                                        const float* scale, const float* shift, T* out, int rows, int F,
                                                        int Wp, int C, int Cpad, int pitch) {
  extern __shared__ float sw[];   // [9][Cpad] weights, then scale[Cpad], shift[Cpad]
  for (int i = threadIdx.x; i < 9 * Cpad; i += blockDim.x) {
    const int c = i % Cpad;
    sw[i] = c < C ? w9[(i / Cpad) * C + c] : 0.f;
  }
  for (int i = threadIdx.x; i < Cpad; i += blockDim.x) {
    sw[9 * Cpad + i] = i < C ? scale[i] : 0.f;
    sw[10 * Cpad + i] = i < C ? shift[i] : 0.f;
  }
  __syncthreads();
  const int groups = Cpad >> 3;
  // One thread = kStemPix (8) consecutive feature columns x 8 output channels, taps outermost: the 64 accumulators and the 3 x 10 input window
  // live in registers, the 8 weights of a tap are fetched from shared memory (one 2 x 128-bit broadcast read per 64 FMAs).  The
  // first form of this kernel kept all 72 weights of the thread in registers (141 registers: 8 warps per SM, every one of them
  // waiting on its own input loads half of the time: 199 us for a 215 MB write).  32-bit index arithmetic (the launcher
  // guarantees the count fits).
  const int nch = (F + kStemPix - 1) / kStemPix;
  const unsigned total = static_cast<unsigned>(rows) * nch * groups;
  // grid-stride: a block stages the weights once and then works through many chunks (one chunk per block made the staging round
  // trip — two dependent global loads and a barrier — longer than the block's arithmetic)
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
  const int g = static_cast<int>(idx % groups);
  const unsigned chunk = idx / groups;
  const int f0 = static_cast<int>(chunk % nch) * kStemPix;
  const int row = static_cast<int>(chunk / nch);
  const int seg = seg_of_row[row];
  T* orow = out + (static_cast<size_t>(row) * Wp + f0) * pitch + g * 8;
  if (seg < 0) {
#pragma unroll
    for (int p = 0; p < kStemPix; ++p)
      if (f0 + p < F) *reinterpret_cast<uint4*>(orow + static_cast<size_t>(p) * pitch) = make_uint4(0, 0, 0, 0);
    continue;
  }
  const int t = row - seg_row_off[seg];
  const int T_ = seg_h[seg];
  const float* base = feats + static_cast<size_t>(seg_frame_off[seg]) * F;
  float x[3][kStemPix + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int tt = t + r - 1;
    const bool rok = tt >= 0 && tt < T_;
#pragma unroll
    for (int q = 0; q < kStemPix + 2; ++q) {
      const int ff = f0 + q - 1;
      x[r][q] = (rok && ff >= 0 && ff < F) ? __ldg(base + static_cast<size_t>(tt) * F + ff) : 0.f;
    }
  }
  float v[kStemPix][8];
#pragma unroll
  for (int p = 0; p < kStemPix; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[p][j] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s_ = 0; s_ < 3; ++s_) {          // same accumulation order per output as before: taps in (r, s) order
      const float4 a = *reinterpret_cast<const float4*>(sw + (r * 3 + s_) * Cpad + g * 8), b4 = *reinterpret_cast<const float4*>(sw + (r * 3 + s_) * Cpad + g * 8 + 4);
      const float w[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int p = 0; p < kStemPix; ++p) {
        const float xv = x[r][p + s_];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[p][j] = fmaf(xv, w[j], v[p][j]);
      }
    }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = sw[9 * Cpad + g * 8 + j]; sh[j] = sw[10 * Cpad + g * 8 + j]; }
#pragma unroll
  for (int p = 0; p < kStemPix; ++p) {
    if (f0 + p >= F) break;
    uint4 o;
    o.x = TypeOps<T>::pack2(fmaxf(fmaf(v[p][0], sc[0], sh[0]), 0.f), fmaxf(fmaf(v[p][1], sc[1], sh[1]), 0.f));
    o.y = TypeOps<T>::pack2(fmaxf(fmaf(v[p][2], sc[2], sh[2]), 0.f), fmaxf(fmaf(v[p][3], sc[3], sh[3]), 0.f));
    o.z = TypeOps<T>::pack2(fmaxf(fmaf(v[p][4], sc[4], sh[4]), 0.f), fmaxf(fmaf(v[p][5], sc[5], sh[5]), 0.f));
    o.w = TypeOps<T>::pack2(fmaxf(fmaf(v[p][6], sc[6], sh[6]), 0.f), fmaxf(fmaf(v[p][7], sc[7], sh[7]), 0.f));
    *reinterpret_cast<uint4*>(orow + static_cast<size_t>(p) * pitch) = o;
  }
  }
}

cudaError_t launch_stem_conv(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off, const int32_t* seg_h,
                             const int32_t* seg_of_row, const float* w9, const float* scale, const float* shift, void* out,
                             int rows, int F, int Wp, int C, int Cpad, int pitch, int is_bf16, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * ((F + kStemPix - 1) / kStemPix) * (Cpad / 8);
  if (total <= 0) return cudaSuccess;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;   // stage-0 capacity is 2^17 rows: far below
  unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned cap = static_cast<unsigned>(sms) * 8u;       // four rounds of the two resident blocks per SM
    if (blocks > cap) blocks = cap;
  }
  const size_t smem = 11 * Cpad * sizeof(float);
  if (is_bf16)
    stem_conv_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(feats, seg_frame_off, seg_row_off, seg_h, seg_of_row, w9, scale,
                                                                shift, static_cast<__nv_bfloat16*>(out), rows, F, Wp, C, Cpad, pitch);
  else
    stem_conv_kernel<__half><<<blocks, 256, smem, st>>>(feats, seg_frame_off, seg_row_off, seg_h, seg_of_row, w9, scale, shift,
                                                         static_cast<__half*>(out), rows, F, Wp, C, Cpad, pitch);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// y = relu(x*scale + shift) over a channel slice, optionally sub-sampled by 2 in time and feature
// (DPN pre-activation BN→ReLU, dpn_model.py:41-42, for the block input that several convs normalise differently;
// the stride-2 1x1 projection, dpn_model.py:75, samples even positions of each segment [ext: TF SAME, k=1]).
template <typename T>
__global__ void __launch_bounds__(256) bn_relu_kernel(const T* in, int in_C, int in_coff, int in_Wp, const float* scale,
                                                      const float* shift, T* out, int out_C, int out_rows, int out_W, int out_Wp, int C,
                                                      int stride, const int32_t* out_seg_of_row, const int32_t* out_seg_row_off,
                                                      const int32_t* in_seg_row_off) {
  const int groups = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(out_rows) * out_W * groups;
  if (idx >= total) return;
  const int g = static_cast<int>(idx % groups);
  const long long pix = idx / groups;
  const int col = static_cast<int>(pix % out_W);
  const int row = static_cast<int>(pix / out_W);
  const int seg = out_seg_of_row[row];
  uint4 o = make_uint4(0, 0, 0, 0);
  if (seg >= 0) {
    const int in_row = stride == 1 ? row : in_seg_row_off[seg] + (row - out_seg_row_off[seg]) * stride;
    const int in_col = col * stride;
    const uint4 x = *reinterpret_cast<const uint4*>(in + (static_cast<size_t>(in_row) * in_Wp + in_col) * in_C + in_coff + g * 8);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
    uint32_t os[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = TypeOps<T>::unpack2(xs[j]);
      const int c = g * 8 + j * 2;
      f.x = fmaxf(f.x * scale[c] + shift[c], 0.f);
      f.y = fmaxf(f.y * scale[c + 1] + shift[c + 1], 0.f);
      os[j] = TypeOps<T>::pack2(f.x, f.y);
    }
    o = make_uint4(os[0], os[1], os[2], os[3]);
  }
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(row) * out_Wp + col) * out_C + g * 8) = o;
}

cudaError_t launch_bn_relu(const void* in, int in_C, int in_coff, int in_Wp, const float* scale, const float* shift, void* out,
                           int out_C, int out_rows, int out_W, int out_Wp, int C, int stride, const int32_t* out_seg_of_row,
                           const int32_t* out_seg_row_off, const int32_t* in_seg_row_off, int is_bf16, cudaStream_t st) {
  const long long total = static_cast<long long>(out_rows) * out_W * (C / 8);
  if (total <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16)
    bn_relu_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), in_C, in_coff, in_Wp, scale, shift,
                                                           static_cast<__nv_bfloat16*>(out), out_C, out_rows, out_W, out_Wp, C, stride,
                                                           out_seg_of_row, out_seg_row_off, in_seg_row_off);
  else
    bn_relu_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(in), in_C, in_coff, in_Wp, scale, shift,
                                                    static_cast<__half*>(out), out_C, out_rows, out_W, out_Wp, C, stride, out_seg_of_row,
                                                    out_seg_row_off, in_seg_row_off);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Res2Net stride-2 last split: avg_pool 3x3 / 2, VALID over the (1,1) zero-padded tensor → divisor always 9
// (res2net_model.py:76-77).  Uniform row map in_row = 2*out_row + r - 1 (layout guarantees it per segment).
// One thread = two horizontally adjacent output pixels x 8 channels: 3 x 5 input vectors instead of 2 x 9, 32-bit index arithmetic
// (the first form spent more instructions on four 64-bit divisions per thread than on the pooling).
template <typename T>
__global__ void __launch_bounds__(256) avgpool3x3s2_kernel(const T* in, int in_C, int in_coff, int in_rows, int in_W, int in_Wp, T* out,
                                                           int out_C, int out_coff, int out_rows, int out_W, int out_Wp, int C,
                                                           const int32_t* out_seg_of_row) {
  const unsigned groups = static_cast<unsigned>(C) >> 3;
  const unsigned pairs = (static_cast<unsigned>(out_W) + 1u) >> 1;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = static_cast<unsigned>(out_rows) * pairs * groups;
  if (idx >= total) return;
  const unsigned g = idx % groups;
  const unsigned pp = idx / groups;
  const int col = static_cast<int>(pp % pairs) * 2;
  const int row = static_cast<int>(pp / pairs);
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
  if (out_seg_of_row[row] >= 0) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ir = 2 * row + r - 1;
      if (ir < 0 || ir >= in_rows) continue;
      const T* irow = in + static_cast<size_t>(ir) * in_Wp * in_C + in_coff + g * 8;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int ic = 2 * col + s - 1;
        if (ic < 0 || ic >= in_W) continue;
        const uint4 x = *reinterpret_cast<const uint4*>(irow + static_cast<size_t>(ic) * in_C);
        const float2 a = TypeOps<T>::unpack2(x.x), b = TypeOps<T>::unpack2(x.y), c = TypeOps<T>::unpack2(x.z),
                     d = TypeOps<T>::unpack2(x.w);
        if (s <= 2) { acc0[0] += a.x; acc0[1] += a.y; acc0[2] += b.x; acc0[3] += b.y; acc0[4] += c.x; acc0[5] += c.y; acc0[6] += d.x; acc0[7] += d.y; }
        if (s >= 2) { acc1[0] += a.x; acc1[1] += a.y; acc1[2] += b.x; acc1[3] += b.y; acc1[4] += c.x; acc1[5] += c.y; acc1[6] += d.x; acc1[7] += d.y; }
      }
    }
  }
  const float k = 1.f / 9.f;
  T* o0 = out + (static_cast<size_t>(row) * out_Wp + col) * out_C + out_coff + g * 8;
  uint4 o;
  o.x = TypeOps<T>::pack2(acc0[0] * k, acc0[1] * k); o.y = TypeOps<T>::pack2(acc0[2] * k, acc0[3] * k);
  o.z = TypeOps<T>::pack2(acc0[4] * k, acc0[5] * k); o.w = TypeOps<T>::pack2(acc0[6] * k, acc0[7] * k);
  *reinterpret_cast<uint4*>(o0) = o;
  if (col + 1 < out_W) {
    o.x = TypeOps<T>::pack2(acc1[0] * k, acc1[1] * k); o.y = TypeOps<T>::pack2(acc1[2] * k, acc1[3] * k);
    o.z = TypeOps<T>::pack2(acc1[4] * k, acc1[5] * k); o.w = TypeOps<T>::pack2(acc1[6] * k, acc1[7] * k);
    *reinterpret_cast<uint4*>(o0 + out_C) = o;
  }
}

cudaError_t launch_avgpool3x3s2(const void* in, int in_C, int in_coff, int in_rows, int in_W, int in_Wp, void* out, int out_C, int out_coff,
                                int out_rows, int out_W, int out_Wp, int C, const int32_t* out_seg_of_row, int is_bf16, cudaStream_t st) {
  const long long total = static_cast<long long>(out_rows) * ((out_W + 1) / 2) * (C / 8);
  if (total <= 0) return cudaSuccess;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16)
    avgpool3x3s2_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), in_C, in_coff, in_rows, in_W, in_Wp,
                                                                static_cast<__nv_bfloat16*>(out), out_C, out_coff, out_rows,
                                                                out_W, out_Wp, C, out_seg_of_row);
  else
    avgpool3x3s2_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(in), in_C, in_coff, in_rows, in_W, in_Wp,
                                                         static_cast<__half*>(out), out_C, out_coff, out_rows, out_W, out_Wp, C,
                                                         out_seg_of_row);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Even pixels of a tensor into (a channel slice of) a tensor of the next stage: the input of a 1x1 stride-2 projection shortcut
// (res2net_model.py:85-87 with the explicit padding of models.py:121-134: none for a 1x1 kernel, so output (R, c) reads input
// (2 R, 2 c)), laid behind the concat slices of the block so that conv3 reads [y | x_even] as one K axis (ConvDesc::fold_*).
// 16-byte vectors, zero on gap rows and on the zero column.  Type-agnostic (16-bit elements).
__global__ void __launch_bounds__(256) subsample2_kernel(const uint16_t* in, int in_C, int in_coff, int in_Wp, uint16_t* out, int out_C,
                                                         int out_coff, int out_rows, int out_W, int out_Wp, int C,
                                                         const int32_t* out_seg_of_row) {
  const unsigned groups = static_cast<unsigned>(C) >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = static_cast<unsigned>(out_rows) * out_Wp * groups;
  if (idx >= total) return;
  const unsigned g = idx % groups;
  const unsigned pix = idx / groups;
  const int col = static_cast<int>(pix % out_Wp);
  const int row = static_cast<int>(pix / out_Wp);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (col < out_W && out_seg_of_row[row] >= 0)
    v = *reinterpret_cast<const uint4*>(in + (static_cast<size_t>(2 * row) * in_Wp + 2 * col) * in_C + in_coff + g * 8);
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(row) * out_Wp + col) * out_C + out_coff + g * 8) = v;
}

cudaError_t launch_subsample2(const void* in, int in_C, int in_coff, int in_Wp, void* out, int out_C, int out_coff, int out_rows, int out_W,
                              int out_Wp, int C, const int32_t* out_seg_of_row, cudaStream_t st) {
  const long long total = static_cast<long long>(out_rows) * out_Wp * (C / 8);
  if (total <= 0) return cudaSuccess;
  if (total >= (1LL << 31)) return cudaErrorInvalidValue;
  subsample2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(static_cast<const uint16_t*>(in), in_C, in_coff, in_Wp,
                                                                               static_cast<uint16_t*>(out), out_C, out_coff, out_rows, out_W,
                                                                               out_Wp, C, out_seg_of_row);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Statistics pooling (models.py:262-269): per segment, per (w, c): mean over time and sqrt(population var + eps).
// Two passes over the segment's rows (they sit in L2), fp32.  Optional fused pre-activation relu(x*scale+shift)
// (DPN concat_bn_relu, dpn_model.py:24-29).  Output fp32 [n_seg, W*2C], index w*2C + {c | C + c}.
// One thread = 2 adjacent channels of one (segment, w); lanes run along channels → coalesced 4-byte loads.
template <typename T>
__global__ void __launch_bounds__(256) stats_pool_kernel(const T* in, int C_tot, int C, int W, int Wp, const int32_t* seg_row_off,
                                                         const int32_t* seg_h, const float* scale, const float* shift, float* out,
                                                         float eps) {
  const int seg = blockIdx.y;
  const int half_c = C >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * half_c) return;
  const int w = idx / half_c;
  const int c = (idx - w * half_c) * 2;
  const int r0 = seg_row_off[seg];
  const int H = seg_h[seg];
  float s0 = 1.f, s1 = 1.f, b0 = 0.f, b1 = 0.f;
  const bool act = scale != nullptr;
  if (act) { s0 = scale[c]; s1 = scale[c + 1]; b0 = shift[c]; b1 = shift[c + 1]; }
  const T* base = in + (static_cast<size_t>(r0) * Wp + w) * C_tot + c;
  const size_t rstride = static_cast<size_t>(Wp) * C_tot;
  float m0 = 0.f, m1 = 0.f;
#pragma unroll 8
  for (int h = 0; h < H; ++h) {
    float2 f = TypeOps<T>::unpack2(*reinterpret_cast<const uint32_t*>(base + h * rstride));
    if (act) { f.x = fmaxf(f.x * s0 + b0, 0.f); f.y = fmaxf(f.y * s1 + b1, 0.f); }
    m0 += f.x; m1 += f.y;
  }
  const float inv = 1.f / static_cast<float>(H);
  m0 *= inv; m1 *= inv;
  float v0 = 0.f, v1 = 0.f;
#pragma unroll 8
  for (int h = 0; h < H; ++h) {
    float2 f = TypeOps<T>::unpack2(*reinterpret_cast<const uint32_t*>(base + h * rstride));
    if (act) { f.x = fmaxf(f.x * s0 + b0, 0.f); f.y = fmaxf(f.y * s1 + b1, 0.f); }
    const float d0 = f.x - m0, d1 = f.y - m1;
    v0 += d0 * d0; v1 += d1 * d1;
  }
  float* o = out + static_cast<size_t>(seg) * W * 2 * C + static_cast<size_t>(w) * 2 * C;
  *reinterpret_cast<float2*>(o + c) = make_float2(m0, m1);
  *reinterpret_cast<float2*>(o + C + c) = make_float2(sqrtf(v0 * inv + eps), sqrtf(v1 * inv + eps));
}

cudaError_t launch_stats_pool(const void* in, int C_tot, int C, int W, int Wp, const int32_t* seg_row_off, const int32_t* seg_h, int n_seg,
                              const float* scale, const float* shift, float* out, float eps, int is_bf16, cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  dim3 grid((W * (C / 2) + 255) / 256, n_seg);
  if (is_bf16)
    stats_pool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), C_tot, C, W, Wp, seg_row_off, seg_h,
                                                            scale, shift, out, eps);
  else
    stats_pool_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(in), C_tot, C, W, Wp, seg_row_off, seg_h, scale, shift,
                                                     out, eps);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Attentive statistics pooling (models.py:273-303).  The first 1x1 conv acts on concat(x, tiled mean, tiled std); its
// mean/std part is constant over time, so it is a per-(segment, w) bias:  bias[seg, w, a] = sum_k pooled[seg, w, k] * Wms[k, a]
// (pooled = plain statistics [mean | std], Wms = kernel rows C..3C).  One thread per output, fp32.
__global__ void __launch_bounds__(128) att_bias_kernel(const float* __restrict__ pooled, const float* __restrict__ Wms, float* bias,
                                                       int n_sw, int C2, int A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int sw = blockIdx.y;
  if (a >= A || sw >= n_sw) return;
  const float* p = pooled + static_cast<size_t>(sw) * C2;
  float acc = 0.f;
  for (int k = 0; k < C2; ++k) acc += p[k] * Wms[static_cast<size_t>(k) * A + a];
  bias[static_cast<size_t>(sw) * A + a] = acc;
}

cudaError_t launch_att_bias(const float* pooled, const float* Wms, float* bias, int n_seg, int W, int C2, int A, cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  dim3 grid((A + 127) / 128, n_seg * W);
  att_bias_kernel<<<grid, 128, 0, st>>>(pooled, Wms, bias, n_seg * W, C2, A);
  return cudaGetLastError();
}

// t = tanh(t + bias[seg(row), col]) in place on the [pixels, A] pre-activations of the first attention conv; 0 outside segments.
template <typename T>
__global__ void __launch_bounds__(256) att_tanh_kernel(T* t, int A, long long n_pix, int Wp, int W, const int32_t* seg_of_row,
                                                       const float* __restrict__ bias) {
  const int groups = A >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_pix * groups) return;
  const int g = static_cast<int>(idx % groups);
  const long long pix = idx / groups;
  const int row = static_cast<int>(pix / Wp), col = static_cast<int>(pix - static_cast<long long>(row) * Wp);
  const int seg = seg_of_row[row];
  uint4* ptr = reinterpret_cast<uint4*>(t + pix * A + g * 8);
  uint4 o = make_uint4(0, 0, 0, 0);
  if (seg >= 0 && col < W) {
    const uint4 x = *ptr;
    const float* b = bias + (static_cast<size_t>(seg) * W + col) * A + g * 8;
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
    uint32_t os[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = TypeOps<T>::unpack2(xs[j]);
      os[j] = TypeOps<T>::pack2(tanhf(f.x + b[2 * j]), tanhf(f.y + b[2 * j + 1]));
    }
    o = make_uint4(os[0], os[1], os[2], os[3]);
  }
  *ptr = o;
}

cudaError_t launch_att_tanh(void* t, int A, long long n_pix, int Wp, int W, const int32_t* seg_of_row, const float* bias, int is_bf16,
                            cudaStream_t st) {
  const long long total = n_pix * (A / 8);
  if (total <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (is_bf16) att_tanh_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<__nv_bfloat16*>(t), A, n_pix, Wp, W, seg_of_row, bias);
  else att_tanh_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<__half*>(t), A, n_pix, Wp, W, seg_of_row, bias);
  return cudaGetLastError();
}

// Softmax over time of the attention logits per (segment, w, c), then weighted mean and sqrt(weighted second moment - mean^2 + eps)
// (models.py:298-303).  Same thread mapping and output layout as stats_pool_kernel (index w*2C + {c | C + c}).
template <typename T>
__global__ void __launch_bounds__(256) att_pool_kernel(const T* x, const T* logits, int C, int W, int Wp, const int32_t* seg_row_off,
                                                       const int32_t* seg_h, float* out, float eps) {
  const int seg = blockIdx.y;
  const int half_c = C >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * half_c) return;
  const int w = idx / half_c;
  const int c = (idx - w * half_c) * 2;
  const int r0 = seg_row_off[seg];
  const int H = seg_h[seg];
  const size_t base = (static_cast<size_t>(r0) * Wp + w) * C + c;
  const size_t rstride = static_cast<size_t>(Wp) * C;
  float m0 = -INFINITY, m1 = -INFINITY;
  for (int h = 0; h < H; ++h) {
    const float2 l = TypeOps<T>::unpack2(*reinterpret_cast<const uint32_t*>(logits + base + h * rstride));
    m0 = fmaxf(m0, l.x); m1 = fmaxf(m1, l.y);
  }
  float s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int h = 0; h < H; ++h) {
    const float2 l = TypeOps<T>::unpack2(*reinterpret_cast<const uint32_t*>(logits + base + h * rstride));
    const float2 v = TypeOps<T>::unpack2(*reinterpret_cast<const uint32_t*>(x + base + h * rstride));
    const float e0 = __expf(l.x - m0), e1 = __expf(l.y - m1);
    s0 += e0; s1 += e1;
    a0 += v.x * e0; a1 += v.y * e1;
    q0 += v.x * v.x * e0; q1 += v.y * v.y * e1;
  }
  const float wm0 = a0 / s0, wm1 = a1 / s1;
  float* o = out + static_cast<size_t>(seg) * W * 2 * C + static_cast<size_t>(w) * 2 * C;
  *reinterpret_cast<float2*>(o + c) = make_float2(wm0, wm1);
  *reinterpret_cast<float2*>(o + C + c) = make_float2(sqrtf(q0 / s0 - wm0 * wm0 + eps), sqrtf(q1 / s1 - wm1 * wm1 + eps));
}

cudaError_t launch_att_pool(const void* x, const void* logits, int C, int W, int Wp, const int32_t* seg_row_off, const int32_t* seg_h,
                            int n_seg, float* out, float eps, int is_bf16, cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  dim3 grid((W * (C / 2) + 255) / 256, n_seg);
  if (is_bf16)
    att_pool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(logits), C, W, Wp,
                                                          seg_row_off, seg_h, out, eps);
  else
    att_pool_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<const __half*>(logits), C, W, Wp, seg_row_off,
                                                   seg_h, out, eps);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Embedding FC with both 2-D batch norms folded in (BN → dense → BN, res2net_model.py:240-242):
//   out[n, e] = bias[e] + sum_d pooled[n, d] * Wf[d, e]       Wf = diag(s1) W diag(s2), fp32.
// Split-K SGEMM on the CUDA cores, fp32 throughout: one wave of CTAs (row tiles x K splits ~ the SM count), CTA tile 128 segments x
// 256 outputs, thread tile 8 x 16 (128 accumulators: 6 x 128-bit shared-memory reads per 128 FMAs), K in slabs of 16 through two
// shared-memory buffers with the next slab's global loads in flight; every split writes a partial-sum slab and a second kernel adds
// the slabs in a fixed order (bit-reproducible, no atomics).  The first form (one output column x 32 segments per thread, 1 280
// small blocks) ran at 23 TFLOP/s; this one is bound by the FMA pipe.
constexpr int kFcTileM = 128, kFcTileN = 256, kFcSlab = 16;

__global__ void __launch_bounds__(256, 1) fc_kernel(const float* __restrict__ pooled, const float* __restrict__ Wf, float* partial, int n,
                                                    int D, int E, int slabs_per_split) {
  __shared__ __align__(16) float As[2][kFcSlab][kFcTileM];     // [k][segment]
  __shared__ __align__(16) float Bs[2][kFcSlab][kFcTileN];     // [k][output]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                      // outputs tx*4 + {0,64,128,192} + 0..3, segments ty*8 + 0..7
  const int n0 = blockIdx.y * kFcTileM;
  const int e0 = blockIdx.x * kFcTileN;
  const int k_begin = blockIdx.z * slabs_per_split * kFcSlab;
  const int k_end = min(D, k_begin + slabs_per_split * kFcSlab);
  const bool vec = (D & 3) == 0 && (E & 3) == 0;
  // global -> registers of one slab: A 128 x 16 (two float4 per thread: row tid & 127, k quads (tid >> 7) and (tid >> 7) + 2),
  // B 16 x 256 (four float4 per thread: k = tid >> 6 (+4, +8, +12), outputs (tid & 63) * 4)
  float4 ra[2], rb[4];
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = tid & 127, kq = (tid >> 7) + 2 * i, k = k0 + kq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < n) {
        const float* src = pooled + static_cast<size_t>(n0 + r) * D + k;
        if (vec && k + 4 <= k_end) v = __ldg(reinterpret_cast<const float4*>(src));
        else { if (k < k_end) v.x = src[0]; if (k + 1 < k_end) v.y = src[1]; if (k + 2 < k_end) v.z = src[2]; if (k + 3 < k_end) v.w = src[3]; }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = (tid >> 6) + 4 * i, c = (tid & 63) * 4, k = k0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end) {
        const float* src = Wf + static_cast<size_t>(k) * E + e0 + c;
        if (vec && e0 + c + 4 <= E) v = __ldg(reinterpret_cast<const float4*>(src));
        else { if (e0 + c < E) v.x = src[0]; if (e0 + c + 1 < E) v.y = src[1]; if (e0 + c + 2 < E) v.z = src[2]; if (e0 + c + 3 < E) v.w = src[3]; }
      }
      rb[i] = v;
    }
  };
  auto store_slab = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = tid & 127, kq = (tid >> 7) + 2 * i;
      As[buf][kq * 4 + 0][r] = ra[i].x; As[buf][kq * 4 + 1][r] = ra[i].y; As[buf][kq * 4 + 2][r] = ra[i].z; As[buf][kq * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = (tid >> 6) + 4 * i, c = (tid & 63) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][c]) = rb[i];
    }
  };
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  int buf = 0;
  if (k_begin < k_end) { load_slab(k_begin); store_slab(0); }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += kFcSlab) {
    const bool more = k0 + kFcSlab < k_end;
    if (more) load_slab(k0 + kFcSlab);
#pragma unroll
    for (int kk = 0; kk < kFcSlab; ++kk) {                     // K ascending inside the slab, slabs ascending: a fixed summation order
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][c * 64 + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_slab(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  float* po = partial + static_cast<size_t>(blockIdx.z) * n * E;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = n0 + ty * 8 + i;
    if (r >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int e = e0 + c * 64 + tx * 4;
      float* dst = po + static_cast<size_t>(r) * E + e;
      if (vec && e + 4 <= E) *reinterpret_cast<float4*>(dst) = make_float4(acc[i][c * 4], acc[i][c * 4 + 1], acc[i][c * 4 + 2], acc[i][c * 4 + 3]);
      else
        for (int q = 0; q < 4; ++q) if (e + q < E) dst[q] = acc[i][c * 4 + q];
    }
  }
}

__global__ void fc_reduce_kernel(const float* partial, const float* bias, float* out, int n, int E, int splits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * E) return;
  float acc = bias[i % E];
  for (int z = 0; z < splits; ++z) acc += partial[static_cast<size_t>(z) * n * E + i];
  out[i] = acc;
}

// K splits: one wave of CTAs over the SMs (row tiles x output tiles x splits <= SM count), at least 4 slabs per split
static void fc_plan(int n, int D, int E, int* splits, int* slabs_per_split) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ((n + kFcTileM - 1) / kFcTileM) * ((E + kFcTileN - 1) / kFcTileN);
  const int slabs = (D + kFcSlab - 1) / kFcSlab;
  int want = tiles > 0 ? sms / tiles : 1;
  if (want < 1) want = 1;
  int per = (slabs + want - 1) / want;
  if (per < 4) per = 4;
  *slabs_per_split = per;
  *splits = (slabs + per - 1) / per;
}

int fc_splits(int D, int n, int E) {
  int splits = 1, per = 1;
  fc_plan(n, D, E, &splits, &per);
  return splits;
}

cudaError_t launch_fc(const float* pooled, const float* Wf, const float* bias, float* partial, float* out, int n, int D, int E,
                      cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int splits = 1, per = 1;
  fc_plan(n, D, E, &splits, &per);
  dim3 grid((E + kFcTileN - 1) / kFcTileN, (n + kFcTileM - 1) / kFcTileM, splits);
  fc_kernel<<<grid, 256, 0, st>>>(pooled, Wf, partial, n, D, E, per);
  fc_reduce_kernel<<<(n * E + 255) / 256, 256, 0, st>>>(partial, bias, out, n, E, splits);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Chunk combine (tf_extract.py:104-111): utterance embedding = sum_i y_i * len_i / sum_i len_i over its chunks.
__global__ void chunk_combine_kernel(const float* seg_emb, const int32_t* utt_seg_off, const int32_t* seg_len, float* out, int n_utt,
                                     int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_utt * E) return;
  const int u = i / E, e = i % E;
  const int s0 = utt_seg_off[u], s1 = utt_seg_off[u + 1];
  float acc = 0.f, tot = 0.f;
  for (int s = s0; s < s1; ++s) {
    const float l = static_cast<float>(seg_len[s]);
    acc += seg_emb[static_cast<size_t>(s) * E + e] * l;
    tot += l;
  }
  out[i] = acc / tot;
}

cudaError_t launch_chunk_combine(const float* seg_emb, const int32_t* utt_seg_off, const int32_t* seg_len, float* out, int n_utt,
                                 int E, cudaStream_t st) {
  if (n_utt <= 0) return cudaSuccess;
  chunk_combine_kernel<<<(n_utt * E + 255) / 256, 256, 0, st>>>(seg_emb, utt_seg_off, seg_len, out, n_utt, E);
  return cudaGetLastError();
}

}  // namespace svx
