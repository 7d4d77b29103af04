// Micro-benchmark: per-SM rate of TMA tile loads / stores as a function of the box row width, against a
// cooperative LDG.128 → STS gather.  Data is L2-resident (small tensor, reused), 148 persistent CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_rate tma_rate.cu && ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../voxsrc2020_speaker_verification_b200/csrc/umma.cuh"
using namespace svx::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled enc;

static CUtensorMap make_map(void* base, int C, int W, int rows, int box_c, int box_w, int box_h, int promo) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)rows};
  cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h};
  cuuint32_t es[3] = {1, 1, 1};
  int sw = box_c * 2;
  CUtensorMapSwizzle s = sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s,
                   promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

constexpr int MAXST = 32;

// One producer thread issues `iters` box loads per CTA through an 8-stage ring; one consumer thread frees the slots.
__global__ void __launch_bounds__(64) tma_load_kernel(const __grid_constant__ CUtensorMap map, int iters, int box_bytes, int W_tiles,
                                                      int row_tiles, int w_box, int h_box, int STAGES, int seq) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)smem; uint64_t* empty = full + MAXST;
  uint8_t* tiles = smem + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      int s = it % STAGES; uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], box_bytes);
      int t = seq ? (int)(((long long)it * gridDim.x + blockIdx.x) % ((long long)W_tiles * row_tiles)) : (blockIdx.x * 131 + it) % (W_tiles * row_tiles);
      tma_load_3d(tiles + s * box_bytes, &map, &full[s], 0, (t % W_tiles) * w_box, (t / W_tiles) * h_box);
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      int s = it % STAGES; uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
}

// TMA store: one thread stores a box `iters` times (bulk groups, at most 4 in flight).
__global__ void __launch_bounds__(64) tma_store_kernel(const __grid_constant__ CUtensorMap map, int iters, int W_tiles, int row_tiles,
                                                       int w_box, int h_box) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = i;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      int t = (blockIdx.x * 131 + it) % (W_tiles * row_tiles);
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&map),
                   "r"(smem_u32(smem)), "r"(0), "r"((t % W_tiles) * w_box), "r"((t / W_tiles) * h_box) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// 128 threads gather a {row_bytes x 128 rows} tile with LDG.128 and store it to (unswizzled) smem, `iters` times.
__global__ void __launch_bounds__(128) ldg_gather_kernel(const uint8_t* base, int iters, int row_bytes, int pix_pitch, int W, int W_tiles,
                                                         int row_tiles, int w_box, int h_box, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint4* sm = (uint4*)raw;
  const int units = row_bytes / 16;                 // 16-byte units per row
  const int total = 128 * units;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    int t = (blockIdx.x * 131 + it) % (W_tiles * row_tiles);
    int w0 = (t % W_tiles) * w_box, r0 = (t / W_tiles) * h_box;
    for (int i = threadIdx.x; i < total; i += 128) {
      int m = i / units, u = i % units;
      int h = m / w_box, w = m % w_box;
      const uint4* p = (const uint4*)(base + ((size_t)(r0 + h) * W + (w0 + w)) * pix_pitch + u * 16);
      sm[(it & 1) * total + i] = __ldg(p);
    }
    __syncthreads();
    acc += ((uint32_t*)raw)[threadIdx.x];
  }
  if (acc == 0x12345678) sink[0] = acc;
}

// coalesced stores: 128 threads write a 128-row x row_bytes tile from registers, 4 rows x 128 B per warp instruction
__global__ void __launch_bounds__(128) stg_kernel(uint8_t* base, int iters, int row_bytes, int pix_pitch, int W, int W_tiles, int row_tiles,
                                                  int w_box, int h_box, int coalesced) {
  const int units = row_bytes / 16;
  const int total = 128 * units;
  uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  for (int it = 0; it < iters; ++it) {
    int t = (blockIdx.x * 131 + it) % (W_tiles * row_tiles);
    int w0 = (t % W_tiles) * w_box, r0 = (t / W_tiles) * h_box;
    if (coalesced) {
      for (int i = threadIdx.x; i < total; i += 128) {
        int m = i / units, u = i % units;
        int h = m / w_box, w = m % w_box;
        *(uint4*)(base + ((size_t)(r0 + h) * W + (w0 + w)) * pix_pitch + u * 16) = v;
      }
    } else {                                         // thread = row, one 16-byte unit per instruction (the v0 epilogue pattern)
      int m = threadIdx.x, h = m / w_box, w = m % w_box;
      uint8_t* p = base + ((size_t)(r0 + h) * W + (w0 + w)) * pix_pitch;
      for (int u = 0; u < units; ++u) *(uint4*)(p + u * 16) = v;
    }
  }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  enc = (PFN_encodeTiled)fn;
  const int W = 80, w_box = 16, h_box = 8;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(tma_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct Cfg { int C; int box_c; const char* name; };
  Cfg cfgs[] = {{96, 32, "64B rows of 96-ch tensor"}, {128, 64, "128B rows of 128-ch tensor"}, {64, 64, "128B rows dense 64-ch"}};
  for (int dram = 0; dram < 2; ++dram) {
    const int rows = dram ? 262144 : 1024;      // 262144 x 80 px x C x 2 B = 4-5 GB (DRAM) vs 16-20 MB (L2)
    for (auto& c : cfgs) {
      void* d; size_t bytes = (size_t)rows * W * c.C * 2;
      CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes));
      const int W_tiles = W / w_box, row_tiles = rows / h_box;
      CUtensorMap m = make_map(d, c.C, W, rows, c.box_c, w_box, h_box, 0);
      int box_bytes = c.box_c * 2 * 128;
      for (int st : {2, 4, 8, 12, 16, 24}) {
        if ((size_t)st * box_bytes > 200 * 1024) continue;
        size_t smem = 2048 + (size_t)st * box_bytes;
        const int iters = 2000;
        tma_load_kernel<<<148, 64, smem>>>(m, 200, box_bytes, W_tiles, row_tiles, w_box, h_box, st, dram);
        CK(cudaEventRecord(e0));
        tma_load_kernel<<<148, 64, smem>>>(m, iters, box_bytes, W_tiles, row_tiles, w_box, h_box, st, dram);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s TMA load %-28s stages=%2d: %.3f us/box -> %.2f ns/row, %.1f GB/s/SM, %.2f TB/s chip, implied latency %.2f us\n",
               dram ? "DRAM" : "L2  ", c.name, st, ms * 1e3 / iters, ms * 1e6 / iters / 128, box_bytes / (ms * 1e-3 / iters) / 1e9,
               148.0 * box_bytes / (ms * 1e-3 / iters) / 1e12, st * ms * 1e3 / iters);
      }
      cudaFree(d);
    }
  }
  return 0;
}
