// Microbenchmark: what slows tcgen05.mma down inside conv_flat?  Warp 0 issues the kernel's narrow-3x3 MMA pattern
// (M=128, N=32, K=16, 64-byte rows, 4 accumulators, shifted descriptors) while other warps of the same CTA generate one kind of
// traffic each:  bit0 tcgen05.ld of other TMEM columns (8 warps) | bit1 ld/st.shared (8 warps) | bit2 TMA loads (64-B rows) |
// bit3 TMA stores (64-B rows).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../voxsrc2020_speaker_verification_b200/csrc mma_contend.cu -o mma_contend
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace svx::ptx;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MT, int KS, int SUBROWS>
__device__ __forceinline__ void issue_tap(uint64_t adesc0, uint64_t bdesc0, uint32_t d_tmem, uint32_t n_tile, uint32_t idesc) {
  constexpr uint32_t kSub16 = SUBROWS * 2u * KS;      // (SUBROWS rows x 32*KS bytes) >> 4
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < MT; ++j) {
#pragma unroll
      for (int k = 0; k < KS; ++k) umma_f16(d_tmem + j * n_tile, adesc0 + (j * kSub16 + k * 2), bdesc0 + k * 2, idesc, 1u);
    }
  }
}

constexpr int kStages = 4;
constexpr int kBoxBytes = 128 * 64;

__constant__ int c_shift[9] = {0, 1, 2, 81, 82, 83, 162, 163, 164};

__global__ void __launch_bounds__(384, 1) k(const __grid_constant__ CUtensorMap map, int n, int iters, int mode, int ld_cols, int variant, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, full[kStages], empty[kStages];
  __shared__ uint32_t slot;
  __shared__ volatile int done, prod_done;
  __shared__ volatile unsigned long long cnt[4];
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (variant & 16) {      // random fp16 values in (-1, 1) instead of zeros
      uint32_t h = (i * 2654435761u) ^ (i >> 7);
      h *= 0x9E3779B1u;
      v = (h & 0x83ff83ffu) | 0x38003800u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    done = 0; prod_done = 0; cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  uint8_t* ring = smem + 96 * 1024;          // TMA load ring (4 x 8 KB)
  uint8_t* stbuf = smem + 128 * 1024;        // TMA store source (8 KB)
  uint8_t* lsbuf = smem + 136 * 1024;        // ld/st.shared area (32 KB)
  if (warp == 0) {
    const uint32_t row_bytes = 64;
    const uint64_t base = make_kmajor_desc(0, 8 * row_bytes, 4u);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);   // 9 x 2 KB weight items
    const uint32_t idesc = make_idesc_f16(0, 128, n);
    long long t0 = clock64();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    int it = 0;
    while (it < iters) {
      for (int tap = 0; tap < 9; ++tap, it += 8) {
        // variant bit0: sub-tiles 128 rows apart (as in the kernel) instead of 32; bit1: the kernel's tap shifts; bit2: fence per tap;
        // bit3: a different weight item per tap; bit4: random operand data instead of zeros
        const uint32_t a_tap = a0 + ((variant & 2) ? c_shift[tap] : 82 + tap * 7) * row_bytes;
        const uint32_t b_tap = b0 + ((variant & 8) ? tap * 2048 : 0);
        if (variant & 4) tc_fence_after();
        if (variant & 1) issue_tap<4, 2, 128>(base + (a_tap >> 4), base + (b_tap >> 4), tb, n, idesc);
        else issue_tap<4, 2, 32>(base + (a_tap >> 4), base + (b_tap >> 4), tb, n, idesc);
      }
    }
    if (elect_one()) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    done = 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = static_cast<long long>(g1 - g0);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t2 - t0;
  } else if (warp == 1) {
    if ((mode & 4) && lane == 0) {
      unsigned long long c = 0;
      for (int i = 0; !done; ++i, ++c) {
        const int s = i % kStages; const uint32_t ph = (i / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], kBoxBytes);
        tma_load_2d(ring + s * kBoxBytes, &map, &full[s], 0, ((blockIdx.x * 131 + i) % 512) * 128);
      }
      cnt[2] = c;
      __threadfence_block();
      prod_done = 1;
    }
  } else if (warp == 2) {
    if ((mode & 4) && lane == 0) {
      for (int i = 0;; ++i) {                 // drains everything the producer issued before leaving
        const int s = i % kStages; const uint32_t ph = (i / kStages) & 1;
        bool got;
        while (!(got = mbar_try_wait(&full[s], ph))) if (prod_done && static_cast<unsigned long long>(i) >= cnt[2]) break;
        if (!got) break;
        mbar_arrive(&empty[s]);
      }
    }
  } else if (warp == 3) {
    if ((mode & 8) && lane == 0) {
      unsigned long long c = 0;
      for (int i = 0; !done; ++i, ++c) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&map),
                     "r"(smem_u32(stbuf)), "r"(0), "r"(((blockIdx.x * 131 + i) % 512) * 128) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      cnt[3] = c;
    }
  } else {
    const int q4 = warp & 3;
    unsigned long long c = 0;
    uint32_t acc = 0;
    if (mode & 1) {
      uint32_t v[16];
      while (!done) {
        for (int col = 0; col < ld_cols; col += 16) {
          tmem_ld16(tb + (static_cast<uint32_t>(q4 * 32) << 16) + 256 + col, v);
          tmem_ld_wait();
          acc += v[0] ^ v[15];
        }
        ++c;
      }
      if (lane == 0) atomicAdd(const_cast<unsigned long long*>(&cnt[0]), c);
    }
    if (mode & 2) {
      uint4* p = reinterpret_cast<uint4*>(lsbuf) + (warp - 4) * 256;
      while (!done) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint4 x = p[u * 32 + (lane ^ u)];
          x.x += acc;
          p[u * 32 + lane] = x;
          acc += x.y;
        }
        ++c;
      }
      if (lane == 0) atomicAdd(const_cast<unsigned long long*>(&cnt[1]), c);
    }
    if (acc == 0x1234567u) out[3] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[4] = cnt[0]; out[5] = cnt[1]; out[6] = cnt[2]; out[7] = cnt[3]; }
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const int P = 512 * 128 + 1024, C = 32;
  void* g; cudaMalloc(&g, (size_t)P * C * 2); cudaMemset(g, 0, (size_t)P * C * 2);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)P}; cuuint64_t str[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 72 * 256;
  const char* names[16] = {"alone", "tmem.ld", "ld/st.shared", "tmem.ld + ld/st.shared", "TMA load", "", "", "", "TMA store", "", "", "",
                           "TMA load+store", "", "", "all"};
  for (int n : {32, 64})
    for (int variant : {0, 15, 16, 31})
    for (int mode : {0, 1, 2, 3, 4, 8, 12, 15})
      for (int ld_cols : {32, 128}) {
        if (!(mode & 1) && ld_cols != 32) continue;
        if (variant != 31 && mode != 0) continue;
        cudaMemset(d, 0, 64);
        k<<<148, 384, 200 * 1024>>>(m, n, iters, mode, ld_cols, variant, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8] = {0};
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        const double cyc = double(h[0]);
        printf("N %2d variant %2d  %-24s ld_cols %3d: %.1f cyc/MMA %.1f ns/MMA | per 1000 cyc: tmem.ld sweeps %.1f, lds/sts rounds %.1f, TMA load boxes %.2f, TMA store boxes %.2f %s\n",
               n, variant, names[mode], ld_cols, cyc / iters, double(h[1]) / iters, h[4] * 1e3 / cyc, h[5] * 1e3 / cyc, h[6] * 1e3 / cyc, h[7] * 1e3 / cyc,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
