// Plain CUDA-core direct convolution over the tall-image layout: any stride, dilation, padding and groups.
// It runs the layers the tcgen05 kernel does not cover (stride-2 and grouped convs, tiny channel counts) and is
// the on-device cross-check for the tensor-core path (SVX_FORCE_SIMPLE=1).  fp32 accumulation.
#include "conv.cuh"

namespace svx {

template <typename T>
__global__ void __launch_bounds__(256) conv_simple_kernel(const SimpleConvParams p) {
  const int n_valid = p.epi.n_valid;
  const long long total = static_cast<long long>(p.out_rows) * p.out_W * n_valid;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % n_valid);
  const long long pix = idx / n_valid;
  const int col = static_cast<int>(pix % p.out_W);
  const int row = static_cast<int>(pix / p.out_W);
  const bool valid = p.epi.seg_of_row ? (p.epi.seg_of_row[row] >= 0) : true;
  float acc = 0.f;
  if (valid) {
    const int g = c / p.cout_g;
    const int cin0 = g * p.cin_g;                                   // first input channel of this group
    const int abase = p.grp_ntile > 0 ? (c / p.grp_ntile) * p.grp_cstep : 0;
    const T* w = static_cast<const T*>(p.wgt) + static_cast<size_t>(c) * (p.kh * p.kw * p.kpad);
    const T* in = static_cast<const T*>(p.in);
    for (int r = 0; r < p.kh; ++r) {
      const int ir = row * p.sh + r * p.dh - p.ph;
      if (ir < 0 || ir >= p.in_rows) continue;
      for (int s = 0; s < p.kw; ++s) {
        const int ic = col * p.sw + s * p.dw - p.pw;
        if (ic < 0 || ic >= p.in_W) continue;
        const T* x = in + (static_cast<size_t>(ir) * p.in_Wp + ic) * p.in_C + p.in_coff + cin0;
        const T* wk = w + (r * p.kw + s) * p.kpad + (cin0 - abase);
        for (int ci = 0; ci < p.cin_g; ++ci) acc += TypeOps<T>::to_f(x[ci]) * TypeOps<T>::to_f(wk[ci]);
      }
    }
  }
  epilogue_scalar<T>(p.epi, acc, row, col, p.out_Wp, c, valid);
}

cudaError_t launch_conv_simple(const SimpleConvParams& p, int is_bf16, cudaStream_t stream) {
  const long long total = static_cast<long long>(p.out_rows) * p.out_W * p.epi.n_valid;
  if (total <= 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  if (is_bf16)
    conv_simple_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(p);
  else
    conv_simple_kernel<__half><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace svx
