// Probe of tcgen05.mma.cta_group::2 semantics (M = 256 over a CTA pair), preparing the 2-CTA conv kernel of the next round:
// which halves of A / B each CTA supplies and where D lands.  One cluster of 2 CTAs; operands are written with plain stores in
// the SWIZZLE_128B K-major layout; the leader issues the MMAs; both CTAs read their own TMEM and the host checks hypotheses.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../voxsrc2020_speaker_verification_b200/csrc umma_2cta.cu -o umma_2cta
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace svx::ptx;

constexpr int K = 64;       // one SWIZZLE_128B box of fp16

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sw_store(uint8_t* tile, int r, int k, float v) {   // element (r, k) of a [rows][64] fp16 K-major SW128 tile
  const uint32_t off = r * 128 + ((((k >> 3) ^ (r & 7)) << 4)) + (k & 7) * 2;
  *reinterpret_cast<__half*>(tile + off) = __float2half(v);
}

// A value patterns: A[row, k], B[n, k] with row in 0..255 (global), n in 0..N-1 (global)
__host__ __device__ inline float a_val(int row, int k) { return float((row * 3 + k * 5) % 7 - 3); }
__host__ __device__ inline float b_val(int n, int k) { return float((n * 2 + k * 3) % 5 - 2); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int N, int b_split, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = smem;                 // [128][64] fp16 = 16 KB
  uint8_t* b_tile = smem + 16384;         // [<=256][64] fp16
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * K; i += 128) sw_store(a_tile, i / K, i % K, a_val(rank * 128 + i / K, i % K));
  const int b_rows = b_split ? N / 2 : N;
  for (int i = threadIdx.x; i < b_rows * K; i += 128) sw_store(b_tile, i / K, i % K, b_val((b_split ? rank * (N / 2) : 0) + i / K, i % K));
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  if (rank == 0 && warp == 0) {
    const uint64_t base = make_kmajor_desc(0, 1024, 2);
    const uint32_t idesc = make_idesc_f16(0, 256, N);
    const uint64_t ad = base + (smem_u32(a_tile) >> 4), bd = base + (smem_u32(b_tile) >> 4);
    if (elect_one()) {
      for (int k = 0; k < K / 16; ++k) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tb), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // every warp reads its 32 lanes x N columns
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(static_cast<size_t>(rank) * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {64, 128, 256})
    for (int b_split : {1, 0}) {
      float* d; cudaMalloc(&d, 256 * N * 4); cudaMemset(d, 0xff, 256 * N * 4);
      probe<<<2, 128, 64 * 1024>>>(N, b_split, d);
      cudaError_t e = cudaDeviceSynchronize();
      float* h = (float*)malloc(256 * N * 4);
      cudaMemcpy(h, d, 256 * N * 4, cudaMemcpyDeviceToHost);
      // hypothesis: D[row, n] = sum_k A[row, k] * B[n, k] with A rows 0-127 from CTA 0, 128-255 from CTA 1
      int bad = 0, bad_first = -1;
      for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
          float want = 0;
          for (int k = 0; k < K; ++k) want += a_val(r, k) * b_val(n, k);
          if (h[r * N + n] != want) { if (!bad) bad_first = r * N + n; ++bad; }
        }
      printf("N %3d b_split %d: %s  mismatches %d/%d (first at row %d col %d: got %g)\n", N, b_split,
             e == cudaSuccess ? "ok" : cudaGetErrorString(e), bad, 256 * N, bad_first >= 0 ? bad_first / N : -1, bad_first >= 0 ? bad_first % N : -1,
             bad_first >= 0 ? h[bad_first] : 0.f);
      if (e != cudaSuccess) return 1;
      cudaFree(d); free(h);
    }
  return 0;
}
