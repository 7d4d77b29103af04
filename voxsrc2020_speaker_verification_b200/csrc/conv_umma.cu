// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM, operands by TMA).
//
//   D[128 pixels, n_tile] = sum over taps, k-boxes of  A_tap[128 pixels, kbox] * B[n_tile, kbox]^T
//
// A is never materialised (no im2col): for every filter tap the producer issues one 3-D TMA box load
// {kbox channels, w_box columns, h_box rows} from the NHWC tall image, displaced by the tap's (dh, dw).  The box
// lands in shared memory as 128 rows of one swizzle span each, which is exactly the canonical K-major UMMA
// operand layout; out-of-image coordinates (left/right feature-axis padding, first/last rows, channels past
// the end of a channel slice) are zero-filled by the TMA unit.  Stride-2 convs read 4 parity-phase views of
// the input through separate tensor maps.  B is the [Cout, taps*kpad] weight matrix, one 2-D box per k-step.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (TMEM → registers → fused scale/shift/ReLU/residual/concat routing → 16-byte stores).
// Shared memory is sized so that two CTAs share an SM: one CTA's epilogue overlaps the other's main loop.
#include "conv.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

constexpr int kUmmaThreads = 192;
constexpr int kMaxStages = 8;

template <typename T>
__device__ __forceinline__ void epilogue8(const Epilogue& e, float (&v)[8], int c, size_t pix, int row, bool valid) {
  // c: first of 8 consecutive output channels (multiple of 8), all < n_valid
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (e.pre_relu) v[j] = fmaxf(v[j], 0.f);
  if (e.scale) {
    const float4 s0 = *reinterpret_cast<const float4*>(e.scale + c);
    const float4 s1 = *reinterpret_cast<const float4*>(e.scale + c + 4);
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
    v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
  }
  if (e.shift) {
    const float4 s0 = *reinterpret_cast<const float4*>(e.shift + c);
    const float4 s1 = *reinterpret_cast<const float4*>(e.shift + c + 4);
    v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
    v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
  }
  if (e.out_f32) {
    float* o = e.out_f32 + static_cast<size_t>(row) * e.ldf + c;
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  if (c < e.n_split) {
    if (e.res) {
      const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const T*>(e.res) + pix * e.res_C + e.res_coff + c);
      const float2 a = TypeOps<T>::unpack2(r.x), b = TypeOps<T>::unpack2(r.y), cc = TypeOps<T>::unpack2(r.z),
                   d = TypeOps<T>::unpack2(r.w);
      v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += cc.x; v[5] += cc.y; v[6] += d.x; v[7] += d.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (e.post_relu) v[j] = fmaxf(v[j], 0.f);
      if (!valid) v[j] = 0.f;
    }
    uint4 o;
    o.x = TypeOps<T>::pack2(v[0], v[1]); o.y = TypeOps<T>::pack2(v[2], v[3]);
    o.z = TypeOps<T>::pack2(v[4], v[5]); o.w = TypeOps<T>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(static_cast<T*>(e.out) + pix * e.out_C + e.out_coff + c) = o;
    if (e.out2) {
      if (valid) {
        const uint4 r = *reinterpret_cast<const uint4*>(static_cast<const T*>(e.add2) + pix * e.add2_C + e.add2_coff + c);
        const float2 a = TypeOps<T>::unpack2(r.x), b = TypeOps<T>::unpack2(r.y), cc = TypeOps<T>::unpack2(r.z),
                     d = TypeOps<T>::unpack2(r.w);
        v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += cc.x; v[5] += cc.y; v[6] += d.x; v[7] += d.y;
      }
      o.x = TypeOps<T>::pack2(v[0], v[1]); o.y = TypeOps<T>::pack2(v[2], v[3]);
      o.z = TypeOps<T>::pack2(v[4], v[5]); o.w = TypeOps<T>::pack2(v[6], v[7]);
      *reinterpret_cast<uint4*>(static_cast<T*>(e.out2) + pix * e.out2_C + e.out2_coff + c) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (e.post_relu) v[j] = fmaxf(v[j], 0.f);
      if (!valid) v[j] = 0.f;
    }
    uint4 o;
    o.x = TypeOps<T>::pack2(v[0], v[1]); o.y = TypeOps<T>::pack2(v[2], v[3]);
    o.z = TypeOps<T>::pack2(v[4], v[5]); o.w = TypeOps<T>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(static_cast<T*>(e.outb) + pix * e.outb_C + e.outb_coff + (c - e.n_split)) = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(kUmmaThreads)
conv_umma_kernel(const __grid_constant__ UmmaConvParams p, const __grid_constant__ AMaps amaps,
                 const __grid_constant__ CUtensorMap bmap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint8_t* tiles = smem + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&amaps.m[i]);
    prefetch_tmap(&bmap);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mt = blockIdx.x;
  const int row0 = (mt / p.w_tiles) * p.h_box;
  const int w0 = (mt % p.w_tiles) * p.w_box;
  const int n_blk = blockIdx.y;
  const int n0 = n_blk * p.n_tile;
  const int total_it = p.taps * p.nkc;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = static_cast<uint32_t>(128 + p.n_tile) * p.kbox * 2u;
      const int a_c0 = n_blk * p.a_c_step;
      int it = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const CUtensorMap* am = &amaps.m[p.tap_map[tap]];
        const int wc = w0 + p.tap_dw[tap];
        const int rc = row0 + p.tap_dh[tap];
        for (int kc = 0; kc < p.nkc; ++kc, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], tx_bytes);
          uint8_t* a_dst = tiles + static_cast<size_t>(s) * stage_bytes;
          tma_load_3d(a_dst, am, &full_bar[s], a_c0 + kc * p.kbox, wc, rc);
          tma_load_2d(a_dst + p.a_stage_bytes, &bmap, &full_bar[s], (tap * p.nkc + kc) * p.kbox, n0);
        }
      }
    }
  } else if (warp == 1) {
    const int ksteps = p.kbox >> 4;   // UMMA K = 16 elements = 32 bytes
    for (int it = 0; it < total_it; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(tiles + static_cast<size_t>(s) * stage_bytes);
        const uint32_t b_addr = a_addr + p.a_stage_bytes;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = make_kmajor_desc(a_addr + k * 32, p.sbo, p.layout_type);
          const uint64_t bdesc = make_kmajor_desc(b_addr + k * 32, p.sbo, p.layout_type);
          umma_f16(tmem_base, adesc, bdesc, p.idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                       // frees the smem slot when these MMAs retire
        if (it == total_it - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // accumulator row = pixel of the tile
    const int h = m / p.w_box;
    const int w = m - h * p.w_box;
    const int row = row0 + h;
    const int col = w0 + w;
    const bool in_range = (row < p.out_rows) && (col < p.out_W);
    bool valid = in_range;
    if (in_range && p.epi.seg_of_row) valid = p.epi.seg_of_row[row] >= 0;
    const size_t pix = static_cast<size_t>(row) * p.out_W + col;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + c0, r);
      tmem_ld_wait();
      if (in_range) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = n0 + c0 + g * 8;
          if (c < p.epi.n_valid) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
            epilogue8<T>(p.epi, v, c, pix, row, valid);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

size_t conv_umma_smem_bytes(const UmmaConvParams& p) {
  return 2048 + static_cast<size_t>(p.stages) * (p.a_stage_bytes + p.b_stage_bytes);
}

cudaError_t conv_umma_init() {
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_conv_umma(const UmmaConvParams& p, const AMaps& amaps, const CUtensorMap& bmap, int n_tiles,
                             int is_bf16, cudaStream_t stream) {
  const int row_tiles = (p.out_rows + p.h_box - 1) / p.h_box;
  if (row_tiles <= 0) return cudaSuccess;
  dim3 grid(static_cast<unsigned>(row_tiles * p.w_tiles), static_cast<unsigned>(n_tiles), 1);
  const size_t smem = conv_umma_smem_bytes(p);
  if (is_bf16)
    conv_umma_kernel<__nv_bfloat16><<<grid, kUmmaThreads, smem, stream>>>(p, amaps, bmap);
  else
    conv_umma_kernel<__half><<<grid, kUmmaThreads, smem, stream>>>(p, amaps, bmap);
  return cudaGetLastError();
}

}  // namespace svx
