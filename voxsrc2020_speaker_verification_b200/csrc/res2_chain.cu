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
// software pipeline two tiles apart (the halo Wp + 1 is less than a tile): in iteration i the MMA warp issues conv 0 on tile
// T, conv 1 on tile T-2 and conv 2 on tile T-4; each conv's epilogue warpgroup has the other two convs' MMAs (~1.7 us) to
// turn its accumulators into y (written straight to the concat with 256-bit stores) and into the next conv's operand
// (s = x + y, added IN PLACE over the x tile that TMA put in the next ring).  Operands live in three shared-memory rings of
// 128-pixel tiles in the canonical K-major SWIZZLE_64B layout (64-byte pixel rows = 32 padded channels); a tap is a
// shared-memory descriptor displaced by its shift.  A ring of L tiles carries one extra slot that mirrors slot 0, so that a
// 128-row operand window that starts anywhere in the ring is contiguous (a UMMA operand cannot wrap).
//
//   warp 0      producer: weights (resident, 54 KB), then per iteration one tile of x_0, x_1, x_2 (TMA, 2 iterations ahead)
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
constexpr int kL0 = 5, kL1 = 6, kL2 = 6;           // ring depths in tiles (each ring has one more slot: the mirror of slot 0)
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
  uint64_t x_full[3][8];       // ring k, slot: tile of split k landed (tx bytes)
  uint64_t mma_done[3][4];     // conv k, iteration & 3: its MMAs have completed (accumulators ready, operand tiles read)
  uint64_t epi_done[3][4];     // conv k, iteration & 3: accumulators drained, s tile of the next ring written (128 arrivals)
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

  // band of tiles of this CTA
  const long long n_tiles = (p.P + 127) / 128;
  const long long band = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = static_cast<long long>(blockIdx.x) * band;
  const long long t1 = t0 + band < n_tiles ? t0 + band : n_tiles;
  const int n_it = t1 > t0 ? static_cast<int>(t1 - t0) + 6 : 0;
  const long long tA = t0 - 2;            // conv 0's tile in iteration 0

  if (warp == 0 && lane == 0) {
    for (int k = 0; k < 3; ++k) { prefetch_tmap(&maps.x[k]); prefetch_tmap(&maps.w[k]); }
    mbar_init(&S.w_bar, 1);
    for (int k = 0; k < 3; ++k) {
      for (int i = 0; i < 8; ++i) mbar_init(&S.x_full[k][i], 1);
      for (int i = 0; i < 4; ++i) { mbar_init(&S.mma_done[k][i], 1); mbar_init(&S.epi_done[k][i], 128); }
    }
    for (int t = 0; t < 12; ++t) S.tap16[t] = t < 9 ? p.tap_shift[t] * 4 : 0;
    fence_barrier_init();
    mbar_expect_tx(&S.w_bar, 3 * kConvWBytes);     // weights are static: fetched before the dependency wait
    for (int k = 0; k < 3; ++k)
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_smem + k * kConvWBytes + tap * kTapBytes, &maps.w[k], &S.w_bar, tap * 32, 0);
  }
  if (warp == 1) { tmem_alloc(&S.tmem_slot, 256); tmem_relinquish(); }
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
    // ------------------------------------------------------------------ producer
    if (lane == 0 && n_it > 0) {
      // ring k receives the tiles of split k in order; tile number `rel` of a ring goes to slot rel % L (slot 0 also to the mirror)
      auto load = [&](int k, int rel, long long tile) {
        const int L = k == 0 ? kL0 : k == 1 ? kL1 : kL2;
        const uint32_t rb = k == 0 ? kRing0 : k == 1 ? kRing1 : kRing2;
        const int slot = rel % L;
        const bool mirror = k == 0 && slot == 0;       // rings 1 and 2 are mirrored by the epilogue that writes s
        uint64_t* bar = &S.x_full[k][slot];
        mbar_expect_tx(bar, mirror ? 2 * kTileBytes : kTileBytes);
        const int px = static_cast<int>(tile * 128);
        tma_load_2d(ring_smem + rb + static_cast<uint32_t>(slot) * kTileBytes, &maps.x[k], bar, 0, px);
        if (mirror) tma_load_2d(ring_smem + rb + static_cast<uint32_t>(L) * kTileBytes, &maps.x[k], bar, 0, px);
      };
      // prologue: x_0 tiles tA-1 .. tA+2, x_1 tiles tA, tA+1, x_2 tiles tA-2, tA-1
      for (int r = 0; r < 4; ++r) load(0, r, tA - 1 + r);
      for (int r = 0; r < 2; ++r) load(1, r, tA + r);
      for (int r = 0; r < 2; ++r) load(2, r, tA - 2 + r);
      for (int it = 0; it < n_it; ++it) {
        if (it >= 1) {      // the slots refilled below were last read by the MMAs of iteration it-1
          const uint32_t par = static_cast<uint32_t>((it - 1) >> 2) & 1u;
          for (int k = 0; k < 3; ++k) wait_chain(&S.mma_done[k][(it - 1) & 3], par, p.dbg, 0x01 + k, it);
        }
        if (it + 4 <= n_it + 1) load(0, it + 4, tA + it + 3);
        if (it + 2 <= n_it - 1) { load(1, it + 2, tA + it + 2); load(2, it + 2, tA + it); }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    if (n_it > 0) {
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
        const uint32_t buf = static_cast<uint32_t>(it & 1);
        const uint32_t par_m2 = static_cast<uint32_t>((it - 2) >> 2) & 1u, par_m1 = static_cast<uint32_t>((it - 1) >> 2) & 1u;
        // ---- conv 0 on tile tA + it: x_0 tiles (ring numbers) it, it+1, it+2
        if (it == 0) {
          wait_chain(&S.x_full[0][0], 0, p.dbg, 0x11, it);
          wait_chain(&S.x_full[0][1], 0, p.dbg, 0x11, it);
        }
        wait_chain(&S.x_full[0][(it + 2) % kL0], static_cast<uint32_t>((it + 2) / kL0) & 1u, p.dbg, 0x12, it);
        if (it >= 2) wait_chain(&S.epi_done[0][(it - 2) & 3], par_m2, p.dbg, 0x13, it);      // accumulator buffer drained
        tc_fence_after();
        issue_conv<kL0>(r0_lo, (it + 1) % kL0, S.tap16, w_lo, hi, tb + (0u * 2u + buf) * 32u, idesc);
        if (elect_one()) umma_commit(&S.mma_done[0][it & 3]);
        // ---- conv 1 on tile tA + it - 2: s_1 tiles written by conv 0's epilogue up to iteration it-1
        if (it >= 1) wait_chain(&S.epi_done[0][(it - 1) & 3], par_m1, p.dbg, 0x14, it);
        if (it >= 2) wait_chain(&S.epi_done[1][(it - 2) & 3], par_m2, p.dbg, 0x15, it);
        tc_fence_after();
        issue_conv<kL1>(r1_lo, pmod(it - 2, kL1), S.tap16, w_lo + (kConvWBytes >> 4), hi, tb + (1u * 2u + buf) * 32u, idesc);
        if (elect_one()) umma_commit(&S.mma_done[1][it & 3]);
        // ---- conv 2 on tile tA + it - 4: s_2 tiles written by conv 1's epilogue up to iteration it-1
        if (it >= 1) wait_chain(&S.epi_done[1][(it - 1) & 3], par_m1, p.dbg, 0x16, it);
        if (it >= 2) wait_chain(&S.epi_done[2][(it - 2) & 3], par_m2, p.dbg, 0x17, it);
        tc_fence_after();
        issue_conv<kL2>(r2_lo, pmod(it - 2, kL2), S.tap16, w_lo + 2u * (kConvWBytes >> 4), hi, tb + (2u * 2u + buf) * 32u, idesc);
        if (elect_one()) umma_commit(&S.mma_done[2][it & 3]);
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
    for (int it = 0; it < n_it; ++it) {
      const long long tile = tA + it - 2 * k;
      const long long pp = tile * 128 + m;
      const bool valid = pp >= 0 && pp < p.P && p.pix_valid[pp] != 0;
      const uint32_t vmask = valid ? 0xffffffffu : 0u;
      const bool store_y = tile >= t0 && tile < t1 && pp < p.P_cap;
      wait_chain(&S.mma_done[k][it & 3], static_cast<uint32_t>(it >> 2) & 1u, p.dbg, 0x20 + k, it);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(k * 2 + (it & 1)) * 32u;
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
        // s = x_{k+1} + y_k in place over the x tile (ring number `it` of the next ring), slot 0 also into the mirror
        const int slot = it % Ln;
        wait_chain(&S.x_full[k + 1][slot], static_cast<uint32_t>(it / Ln) & 1u, p.dbg, 0x28 + k, it);
        const uint32_t base = ring_next + static_cast<uint32_t>(slot) * kTileBytes + row_off;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = base + ((static_cast<uint32_t>(c) << 4) ^ row_xor);
          const uint4 ax = lds_u4(addr);
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
      mbar_arrive(&S.epi_done[k][it & 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
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
