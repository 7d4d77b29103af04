// Microbenchmark: tcgen05.mma issue/execution rate, M=128, K=16 (kind::f16), N and smem layout swept.
//   variant 0: constant descriptors, 8x unrolled           (pure execution rate)
//   variant 1: descriptors recomputed per MMA from loop counters (what conv_flat's issue loop does)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../voxsrc2020_speaker_verification_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace svx::ptx;

template <int MT, int KS>
__device__ __forceinline__ void issue_tap(uint64_t adesc0, uint64_t bdesc0, uint32_t d_tmem, uint32_t n_tile, uint32_t idesc, uint32_t first) {
  constexpr uint32_t kSub16 = 64u * KS;     // sub-tiles 32 rows apart here so that 4 of them + shifts stay inside the operand area
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < MT; ++j) {
#pragma unroll
      for (int k = 0; k < KS; ++k)
        umma_f16(d_tmem + j * n_tile, adesc0 + (j * kSub16 + k * 2), bdesc0 + k * 2, idesc, k == 0 ? first : 1u);
    }
  }
}

template <int VARIANT>
__global__ void __launch_bounds__(128, 1) k(int n, uint32_t row_bytes, int iters, int shift_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    const uint32_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
    const uint64_t base = make_kmajor_desc(0, 8 * row_bytes, layout);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = make_idesc_f16(0, 128, n);
    const int ksteps = row_bytes / 32;
    long long t0 = clock64();
    if (VARIANT == 0) {
      const uint64_t ad = base | (((a0 + shift_rows * row_bytes) >> 4) & 0x3FFF), bd = base | ((b0 >> 4) & 0x3FFF);
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (elect_one()) umma_f16(tb, ad, bd, idesc, 1);
      }
    } else if (VARIANT == 2) {
      int it = 0;
      while (it < iters) {
        for (int tap = 0; tap < 9; ++tap, it += 8) {
          const uint32_t a_tap = a0 + (shift_rows + tap * 7) * row_bytes;
          issue_tap<4, 2>(base + (a_tap >> 4), base + (b0 >> 4), tb, n <= 64 ? n : 64, idesc, 1u);
        }
      }
    } else {
      int it = 0;
      while (it < iters) {
        for (int tap = 0; tap < 9 && it < iters; ++tap) {
          const uint32_t a_tap = a0 + (shift_rows + tap) * row_bytes;
          for (int j = 0; j < 4; ++j) {
            const uint32_t a_j = a_tap + j * 128u * row_bytes / 4;   // stays inside the 64 KB operand area
            const uint32_t d_j = tb + (j * n) % 256;
            for (int kk = 0; kk < ksteps; ++kk, ++it) {
              const uint64_t ad = base | (((a_j + kk * 32) >> 4) & 0x3FFF), bd = base | (((b0 + kk * 32) >> 4) & 0x3FFF);
              if (elect_one()) umma_f16(d_j, ad, bd, idesc, 1);
            }
          }
        }
      }
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4608;
  for (int grid : {148})
    for (uint32_t rb : {128u, 64u})
      for (int n : {32, 64})
        for (int variant : {0, 2})
          for (int shift : {0, 82}) {
            if (variant == 0) k<0><<<grid, 128, 200 * 1024>>>(n, rb, iters, shift, d);
            else k<2><<<grid, 128, 200 * 1024>>>(n, rb, iters, shift, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2] = {0, 0};
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("grid %3d rowbytes %3u N %3d variant %d shift %d: issue %.1f cyc/MMA, done %.1f cyc/MMA (floor %d) %s\n", grid, rb, n, variant, shift,
                   double(h[0]) / iters, double(h[1]) / iters, 128 * n / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  return 0;
}
