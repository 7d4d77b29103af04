// Fused hierarchical 3x3 chain of a stride-1 Res2Net bottleneck (res2net_model.py:26-78, strides == 1, split == 4):
//
//   y_0 = relu(bn_0(conv3x3(x_0)))            s_1 = x_1 + y_0
//   y_1 = relu(bn_1(conv3x3(s_1)))            s_2 = x_2 + y_1
//   y_2 = relu(bn_2(conv3x3(s_2)))
//
// in ONE persistent kernel: the running sums s_i never leave the SM.  Unfused, every conv of the chain is a launch that reads
// its input and the next split and writes its output AND the running sum (640 B per pixel for the three); fused, the kernel
// reads the three planar splits once and writes the three concat slices once (384 B per pixel).
//
// The image is the flat pixel sequence of conv_flat.cu (p = row*Wp + col, one zero column per row), so filter tap (dh, dw) is
// the constant shift dh*Wp + dw, and a CTA STREAMS through a contiguous band of 128-pixel tiles.  The three convs run as a
// software pipeline kD = 4 tiles apart (the halo Wp + 1 is less than a tile, so one tile would do for the data; the other
// three are SLACK): conv k+1 may start tile U once the epilogue of conv k has finished tile U+1, which it was handed kD - 1
// iterations earlier — with a skew of 2 every iteration waited for the full MMA -> commit -> epilogue -> arrive round trip
// (2.4 us against 1.1 us of MMAs).  Each conv's epilogue warpgroup turns accumulators (TMEM, four buffers per conv) into y
// (written straight to the concat with 256-bit stores) and into the next conv's operand s = x + y, x read by the thread that owns
// the pixel one iteration ahead.  Operands live in three shared-memory rings of 128-pixel tiles in the canonical K-major
// SWIZZLE_64B layout (64-byte pixel rows = 32 padded channels); a tap is a shared-memory descriptor displaced by its shift.
// A ring of L tiles carries one extra slot that mirrors slot 0, so that a 128-row operand window that starts anywhere in the
// ring is contiguous (a UMMA operand cannot wrap).
//
//   warp 0      producer: weights (resident, 54 KB), then one tile of x_0 per iteration (TMA, 2 iterations ahead)
//   warp 1      TMEM allocator + tcgen05.mma issuer: 3 convs x 9 taps x 2 K-steps of M=128, N=32, K=16 per iteration
//   warps 4-15  three epilogue warpgroups, one per conv
//
// A band starts two tiles (conv 0) / one tile (conv 1) early to rebuild the halo of its first tile; those outputs are not
// stored.  Arithmetic, rounding points and MMA order are those of the unfused path (conv_flat.cu with aux mode 2), so the
// results are bit-identical to it.
#include <cstdio>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

namespace {

constexpr int kChainThreads = 512;
constexpr int kD = 4;                               // pipeline skew between consecutive convs, in tiles
constexpr int kL0 = 5, kL1 = kD + 2, kL2 = kD + 2;   // ring depths in tiles (each ring has one more slot: the mirror of slot 0)
constexpr uint32_t kTileBytes = 128u * 64u;        // 128 pixels x 32 channels x 2 bytes
constexpr uint32_t kTapBytes = 2048u;              // one tap's weights: [32 n][32 k] 16-bit, SWIZZLE_64B
constexpr uint32_t kConvWBytes = 9u * kTapBytes;
constexpr uint32_t kHeaderBytes = 2048u;
constexpr uint32_t kRing0 = 0u;
constexpr uint32_t kRing1 = kRing0 + (kL0 + 1) * kTileBytes;
constexpr uint32_t kRing2 = kRing1 + (kL1 + 1) * kTileBytes;
constexpr uint32_t kRingBytes = kRing2 + (kL2 + 1) * kTileBytes;

struct ChainSmem {
  uint64_t w_bar;
  uint64_t x_full[8];          // ring 0, slot: tile of split 0 landed (tx bytes)
  uint64_t mma_done[3][4];     // conv k, its tile counter & 3: the MMAs have completed (accumulators ready, operand tiles read)
  uint64_t epi_done[3][4];     // conv k, its tile counter & 3: accumulators drained, s tile of the next ring written (128 arrivals)
  uint32_t tmem_slot;
  int32_t tap16[12];           // tap shift in 16-byte units of a 64-byte pixel row
  alignas(16) float scale[3][32];   // read as float4
  alignas(16) float shift[3][32];
};
static_assert(sizeof(ChainSmem) <= kHeaderBytes, "header");

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// Bounded wait; on a timeout the first starved thread of each warp records (code, iteration) before the trap.
__device__ __forceinline__ void wait_chain(uint64_t* bar, uint32_t parity, unsigned long long* dbg, int code, int it) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  bool reported = false;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) != 0) continue;
    const unsigned long long now = global_ns();
    if (t0 == 0) { t0 = now; continue; }
    if (!reported && now - t0 > 1000000000ull && dbg && (threadIdx.x & 31) == 0) {
      reported = true;
      dbg[(blockIdx.x & 3) * 16 + (threadIdx.x >> 5)] = (static_cast<unsigned long long>(static_cast<unsigned>(it)) << 32) |
                                                         (static_cast<unsigned long long>(blockIdx.x) << 16) | static_cast<unsigned long long>(code & 0xff);
      __threadfence_system();
    }
    if (now - t0 > kWatchdogNs) __trap();
  }
}

__device__ __forceinline__ int pmod(int a, int m) { const int r = a % m; return r < 0 ? r + m : r; }

// All MMAs of one conv on one tile: 9 taps x 2 K-steps, operand window = the tile's ring slot displaced by the tap shift.
template <int L>
__device__ __forceinline__ void issue_conv(uint32_t ring_lo, int slot, const int32_t* tap16, uint32_t w_lo, uint32_t hi, uint32_t d_tmem, uint32_t idesc) {
  const int base16 = slot * static_cast<int>(kTileBytes >> 4);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    int u = base16 + tap16[tap];
    if (u < 0) u += L * static_cast<int>(kTileBytes >> 4);          // window starts before slot 0: it is the tail of slot L-1 + the mirror
    const uint32_t a_lo = ring_lo + static_cast<uint32_t>(u);
    const uint32_t b_lo = w_lo + static_cast<uint32_t>(tap) * (kTapBytes >> 4);
    if (elect_one()) {
      umma_lo(d_tmem, a_lo, b_lo, hi, idesc, tap == 0 ? 0u : 1u);
      umma_lo(d_tmem, a_lo + 2u, b_lo + 2u, hi, idesc, 1u);
    }
  }
}

}  // namespace

template <typename T>
__global__ void __launch_bounds__(kChainThreads, 1)
res2_chain_kernel(const __grid_constant__ ChainParams p, const __grid_constant__ ChainMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  ChainSmem& S = *reinterpret_cast<ChainSmem*>(smem);
  uint8_t* w_smem = smem + kHeaderBytes;
  uint8_t* ring_smem = w_smem + 3 * kConvWBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // band of tiles of this CTA.  Conv 0 covers tiles tA .. t1+1 (n0 of them), conv 1 tiles t0-1 .. t1 (n1), conv 2 the band itself (n2);
  // conv k starts f_k iterations into the loop, and every barrier / ring / TMEM index below is the conv's OWN tile counter.
  const long long n_tiles = (p.P + 127) / 128;
  const long long band = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = static_cast<long long>(blockIdx.x) * band;
  const long long t1 = t0 + band < n_tiles ? t0 + band : n_tiles;
  const int nb = t1 > t0 ? static_cast<int>(t1 - t0) : 0;
  const int n0 = nb + 4, n1 = nb + 2, n2 = nb;
  constexpr int f1 = kD + 1, f2 = 2 * kD + 2;
  const int n_it = nb > 0 ? f2 + n2 : 0;
  const long long tA = t0 - 2;            // conv 0's first tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.x[0]);
    for (int k = 0; k < 3; ++k) prefetch_tmap(&maps.w[k]);
    mbar_init(&S.w_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&S.x_full[i], 1);
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 4; ++i) { mbar_init(&S.mma_done[k][i], 1); mbar_init(&S.epi_done[k][i], 128); }
    for (int t = 0; t < 12; ++t) S.tap16[t] = t < 9 ? p.tap_shift[t] * 4 : 0;
    fence_barrier_init();
    mbar_expect_tx(&S.w_bar, 3 * kConvWBytes);     // weights are static: fetched before the dependency wait
    for (int k = 0; k < 3; ++k)
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_smem + k * kConvWBytes + tap * kTapBytes, &maps.w[k], &S.w_bar, tap * 32, 0);
  }
  if (warp == 1) { tmem_alloc(&S.tmem_slot, 512); tmem_relinquish(); }
  if (threadIdx.x < 96) {
    const int k = threadIdx.x >> 5, c = threadIdx.x & 31;
    S.scale[k][c] = p.scale[k] ? p.scale[k][c] : 1.f;
    S.shift[k][c] = p.shift[k] ? p.shift[k][c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: x_0 tiles tA-1 .. t1+2, tile tA-1+r in slot r % kL0 (slot 0 also in the mirror)
    if (lane == 0 && nb > 0) {
      auto load = [&](int rel) {
        const int slot = rel % kL0;
        const bool mirror = slot == 0;
        uint64_t* bar = &S.x_full[slot];
        mbar_expect_tx(bar, mirror ? 2 * kTileBytes : kTileBytes);
        const int px = static_cast<int>((tA - 1 + rel) * 128);
        tma_load_2d(ring_smem + kRing0 + static_cast<uint32_t>(slot) * kTileBytes, &maps.x[0], bar, 0, px);
        if (mirror) tma_load_2d(ring_smem + kRing0 + static_cast<uint32_t>(kL0) * kTileBytes, &maps.x[0], bar, 0, px);
      };
      for (int r = 0; r < 4; ++r) load(r);
      for (int i = 0; i + 4 <= n0 + 1; ++i) {       // conv 0's tile i reads ring numbers i, i+1, i+2; number i+4 replaces i-1
        if (i >= 1) wait_chain(&S.mma_done[0][(i - 1) & 3], static_cast<uint32_t>((i - 1) >> 2) & 1u, p.dbg, 0x01, i);
        load(i + 4);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    if (nb > 0) {
      wait_chain(&S.w_bar, 0, p.dbg, 0x10, 0);
      const uint64_t desc_base = make_kmajor_desc(0, 512u, 4u);        // SWIZZLE_64B, 8-row groups 512 bytes apart
      const uint32_t hi = static_cast<uint32_t>(desc_base >> 32);
      const uint32_t lo0 = static_cast<uint32_t>(desc_base);
      const uint32_t ring_u32 = smem_u32(ring_smem);
      const uint32_t r0_lo = lo0 + ((ring_u32 + kRing0) >> 4), r1_lo = lo0 + ((ring_u32 + kRing1) >> 4), r2_lo = lo0 + ((ring_u32 + kRing2) >> 4);
      const uint32_t w_lo = lo0 + (smem_u32(w_smem) >> 4);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = p.idesc;
      for (int it = 0; it < n_it; ++it) {
        // ---- conv 0, its tile i: x_0 ring numbers i, i+1, i+2
        if (it < n0) {
          const int i = it;
          if (i == 0) {
            wait_chain(&S.x_full[0], 0, p.dbg, 0x11, i);
            wait_chain(&S.x_full[1], 0, p.dbg, 0x11, i);
          }
          wait_chain(&S.x_full[(i + 2) % kL0], static_cast<uint32_t>((i + 2) / kL0) & 1u, p.dbg, 0x12, i);
          if (i >= 4) wait_chain(&S.epi_done[0][i & 3], static_cast<uint32_t>((i - 4) >> 2) & 1u, p.dbg, 0x13, i);      // accumulator buffer drained
          tc_fence_after();
          issue_conv<kL0>(r0_lo, (i + 1) % kL0, S.tap16, w_lo, hi, tb + (0u * 4u + static_cast<uint32_t>(i & 3)) * 32u, idesc);
          if (elect_one()) umma_commit(&S.mma_done[0][i & 3]);
        }
        // ---- conv 1, its tile i (image tile tA+1+i): s_1 ring numbers i, i+1, i+2, the last one written by conv 0's epilogue of ITS tile i+2
        if (it >= f1 && it - f1 < n1) {
          const int i = it - f1;
          wait_chain(&S.epi_done[0][(i + 2) & 3], static_cast<uint32_t>((i + 2) >> 2) & 1u, p.dbg, 0x14, i);
          if (i >= 4) wait_chain(&S.epi_done[1][i & 3], static_cast<uint32_t>((i - 4) >> 2) & 1u, p.dbg, 0x15, i);
          tc_fence_after();
          issue_conv<kL1>(r1_lo, (i + 1) % kL1, S.tap16, w_lo + (kConvWBytes >> 4), hi, tb + (1u * 4u + static_cast<uint32_t>(i & 3)) * 32u, idesc);
          if (elect_one()) umma_commit(&S.mma_done[1][i & 3]);
        }
        // ---- conv 2, its tile i (image tile tA+2+i): s_2 ring numbers i, i+1, i+2
        if (it >= f2) {
          const int i = it - f2;
          wait_chain(&S.epi_done[1][(i + 2) & 3], static_cast<uint32_t>((i + 2) >> 2) & 1u, p.dbg, 0x16, i);
          if (i >= 4) wait_chain(&S.epi_done[2][i & 3], static_cast<uint32_t>((i - 4) >> 2) & 1u, p.dbg, 0x17, i);
          tc_fence_after();
          issue_conv<kL2>(r2_lo, (i + 1) % kL2, S.tap16, w_lo + 2u * (kConvWBytes >> 4), hi, tb + (2u * 4u + static_cast<uint32_t>(i & 3)) * 32u, idesc);
          if (elect_one()) umma_commit(&S.mma_done[2][i & 3]);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warpgroup of conv k
    const int k = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int Ln = k == 0 ? kL1 : kL2;                                   // ring that receives s (k < 2)
    const uint32_t ring_next = smem_u32(ring_smem) + (k == 0 ? kRing1 : kRing2);
    const uint32_t row_off = static_cast<uint32_t>(m) * 64u;
    const uint32_t row_xor = static_cast<uint32_t>((m >> 1) & 3) << 4;
    const uint32_t sc_base = smem_u32(&S.scale[k][0]), sh_base = smem_u32(&S.shift[k][0]);
    uint8_t* y_base = p.y + static_cast<size_t>(k) * 64u;
    const uint8_t* xn = k < 2 ? p.xs[k] : nullptr;                        // planar split x_{k+1}: 64 bytes per pixel
    const int nk = nb > 0 ? (k == 0 ? n0 : k == 1 ? n1 : n2) : 0;
    const long long tile0 = tA + k;                                       // image tile of this conv's tile 0
    uint4 xa[4], xb[4];
    auto load_x = [&](long long pp, uint4 (&d)[4]) {
      if (pp >= 0 && pp < p.P_cap) {
        ldg256(xn + static_cast<size_t>(pp) * 64u, d[0], d[1]);
        ldg256(xn + static_cast<size_t>(pp) * 64u + 32u, d[2], d[3]);
      } else d[0] = d[1] = d[2] = d[3] = make_uint4(0, 0, 0, 0);
    };
    if (k < 2 && nk > 0) load_x(tile0 * 128 + m, xa);
    for (int i = 0; i < nk; ++i) {
      const long long tile = tile0 + i;
      const long long pp = tile * 128 + m;
      const bool valid = pp >= 0 && pp < p.P && p.pix_valid[pp] != 0;
      const uint32_t vmask = valid ? 0xffffffffu : 0u;
      const bool store_y = tile >= t0 && tile < t1 && pp < p.P_cap;
      if (k < 2 && i + 1 < nk) load_x(pp + 128, xb);                      // next tile's x values: in flight under this tile's arithmetic
      wait_chain(&S.mma_done[k][i & 3], static_cast<uint32_t>(i >> 2) & 1u, p.dbg, 0x20 + k, i);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(k * 4 + (i & 3)) * 32u;
      uint32_t ra[16], rb[16];
      tmem_ld16(taddr, ra);
      tmem_ld16(taddr + 16, rb);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 s4 = lds_f4(sc_base + g * 16), b4 = lds_f4(sh_base + g * 16);
        const uint32_t* r = g < 4 ? ra + g * 4 : rb + (g - 4) * 4;
        v[g * 4 + 0] = fmaxf(fmaf(__uint_as_float(r[0]), s4.x, b4.x), 0.f);
        v[g * 4 + 1] = fmaxf(fmaf(__uint_as_float(r[1]), s4.y, b4.y), 0.f);
        v[g * 4 + 2] = fmaxf(fmaf(__uint_as_float(r[2]), s4.z, b4.z), 0.f);
        v[g * 4 + 3] = fmaxf(fmaf(__uint_as_float(r[3]), s4.w, b4.w), 0.f);
      }
      if (k < 2) {
        // s = x_{k+1} + y_k -> ring number i of the next ring (slot 0 also into the mirror), before y leaves: the next conv waits for it
        const int slot = i % Ln;
        const uint32_t base = ring_next + static_cast<uint32_t>(slot) * kTileBytes + row_off;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = base + ((static_cast<uint32_t>(c) << 4) ^ row_xor);
          const uint4 ax = xa[c];
          const float2 a0 = TypeOps<T>::unpack2(ax.x), a1 = TypeOps<T>::unpack2(ax.y), a2 = TypeOps<T>::unpack2(ax.z), a3 = TypeOps<T>::unpack2(ax.w);
          uint4 s;
          s.x = TypeOps<T>::pack2(v[c * 8 + 0] + a0.x, v[c * 8 + 1] + a0.y) & vmask; s.y = TypeOps<T>::pack2(v[c * 8 + 2] + a1.x, v[c * 8 + 3] + a1.y) & vmask;
          s.z = TypeOps<T>::pack2(v[c * 8 + 4] + a2.x, v[c * 8 + 5] + a2.y) & vmask; s.w = TypeOps<T>::pack2(v[c * 8 + 6] + a3.x, v[c * 8 + 7] + a3.y) & vmask;
          sts_u4(addr, s);
          if (slot == 0) sts_u4(addr + static_cast<uint32_t>(Ln) * kTileBytes, s);
        }
        fence_proxy_async();            // generic-proxy writes of this thread -> visible to the tensor core's operand reads
      }
      tc_fence_before();
      mbar_arrive(&S.epi_done[k][i & 3]);
      if (store_y) {
        uint4 o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          o[c].x = TypeOps<T>::pack2(v[c * 8 + 0], v[c * 8 + 1]) & vmask; o[c].y = TypeOps<T>::pack2(v[c * 8 + 2], v[c * 8 + 3]) & vmask;
          o[c].z = TypeOps<T>::pack2(v[c * 8 + 4], v[c * 8 + 5]) & vmask; o[c].w = TypeOps<T>::pack2(v[c * 8 + 6], v[c * 8 + 7]) & vmask;
        }
        uint8_t* dst = y_base + static_cast<size_t>(pp) * p.y_pitch;
        stg256(dst, o[0], o[1]);
        stg256(dst + 32, o[2], o[3]);
      }
      if (k < 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) xa[c] = xb[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

size_t res2_chain_smem_bytes() { return 1024 + kHeaderBytes + 3 * kConvWBytes + kRingBytes; }

static int g_chain_sms = 0;

cudaError_t res2_chain_init() {
  cudaError_t e = cudaFuncSetAttribute(res2_chain_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(res2_chain_smem_bytes()));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(res2_chain_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(res2_chain_smem_bytes()));
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_chain_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t launch_res2_chain(const ChainParams& p, const ChainMaps& maps, int is_bf16, cudaStream_t st) {
  if (p.P <= 0) return cudaSuccess;
  const long long n_tiles = (p.P + 127) / 128;
  const int sms = g_chain_sms > 0 ? g_chain_sms : 148;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(static_cast<unsigned>(n_tiles < sms ? n_tiles : sms), 1, 1);
  cfg.blockDim = dim3(kChainThreads, 1, 1);
  cfg.dynamicSmemBytes = res2_chain_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = is_bf16 ? cudaLaunchKernelEx(&cfg, res2_chain_kernel<__nv_bfloat16>, p, maps) : cudaLaunchKernelEx(&cfg, res2_chain_kernel<__half>, p, maps);
  if (le != cudaSuccess) return le;
  return cudaGetLastError();
}

}  // namespace svx
