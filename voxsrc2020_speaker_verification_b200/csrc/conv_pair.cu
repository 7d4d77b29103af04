// 1x1 stride-1 convs of the deep stages as a CTA-PAIR GEMM (tcgen05.mma.cta_group::2): out[pixel, n] = epi(sum_k in[pixel, k] * w[n, k]).
//
// Why a second GEMM kernel: with one CTA per tile (conv_flat.cu) the deep 1x1 convs (K, N = 384 ... 1024) are bound by the bytes
// an SM can RECEIVE from L2 (~60 GB/s per SM measured, profiles/r02_ncu_summary.md): a 256-pixel x 128-channel tile takes
// (256 + 128) * K * 2 bytes for 2 * 256 * 128 * K FLOP = 85 FLOP per received byte, i.e. at most ~5 TFLOP/s per SM.  Multicast
// does not help (every CTA still receives the whole operand).  A CTA pair does: the two SMs of a TPC compute ONE M = 256 x
// N = 256 tile, each holding its own 128 pixels of A and only HALF of the weight rows (the tensor cores read the peer's half
// through the pair's operand path), so a CTA receives (128 + 128) * K * 2 bytes for 2 * 128 * 256 * K FLOP = 128 FLOP per byte.
//
//   cluster = 2 CTAs (one TPC).  Per pair tile: 256 pixels (CTA r: pixels [128 r, 128 r + 128)) x n_tile channels (CTA r
//   holds weight rows [r n_tile/2, (r+1) n_tile/2)), K streamed in 64-element boxes (SWIZZLE_128B) through a ring of stages.
//   warp 0      producer of its CTA: A box + weight half-box per stage; BOTH CTAs' loads complete on the LEADER's `full`
//               barrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped to CTA 0), which the leader
//               arms for the bytes of both
//   warp 1      (leader only) tcgen05.mma.cta_group::2 issuer; its commits are multicast to the `empty` / `tmem_full`
//               barriers of both CTAs
//   warps 4-11  two epilogue warpgroups per CTA, each on half of the tile's columns: accumulators (TMEM, double-buffered,
//               2 x 256 columns) -> BN scale/shift, residual, ReLU -> a 32 x 64 transposition through the warp's private
//               staging tile -> 256-bit global accesses, four lanes per 128-byte line, one 16-channel group (32 bytes) per
//               lane through a per-group routing table (dense output, planar splits, second destination); the peer's
//               epilogue warps hand the accumulator buffer back with a remote arrive on the leader's barrier
//
// The same kernel runs the 3x3 stride-1 convs whose half of the weights fits in shared memory (the 48- and 96-channel hierarchical
// convs of Res2Net stages 2-3): the image is the flat pixel sequence of conv_flat.cu, every CTA loads ONE haloed span
// (128 + 2 halo pixels) per K box and the nine taps are shared-memory descriptors displaced by the tap shift — the same
// displacement in both CTAs of the pair — against RESIDENT weights (9 taps x K boxes x n_tile/2 rows per CTA, loaded once).
// The flat kernel has to stream those weights through a 2-4 deep ring for every span (they do not fit beside its slots), which
// left it at 2.5x its MMA time.  Aux mode 2 writes the hierarchical second output s = x_next + y beside y.
//
// Stride-2 convs of the Res2Net down-sampling blocks (2-D tile mode): a CTA's tile is tile_h output rows; the input is seen through
// four PARITY-PHASE views of the high-resolution tensor (3-D tensor maps with strides of two pixels / two rows: in_row = 2 out_row +
// dh - 1 lands in phase (dh - 1) & 1 at row out_row + floor((dh - 1) / 2), the same for columns), one haloed box per phase and K
// box, laid out [box_h][tile_bw][64 channels] so that box rows are again a flat pixel sequence of pitch tile_bw and every tap is a
// descriptor displaced by (dr + 1) * tile_bw + dc + 1.  conv_umma.cu reloaded one box per TAP (nine per tile).
//
// All three forms share one main loop: a tile is a list of K BOXES (A loads) and every box carries a list of ITEMS (tap
// displacement, weight column, K steps) — 1x1: one item per box; 3x3: nine per box; stride-2 3x3: 1 / 2 / 2 / 4 per phase.
//
// Arithmetic and rounding points are those of conv_flat.cu (same K order, same epilogue expression).
#include <cstdio>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

namespace {

constexpr int kPairThreads = 384;
constexpr uint32_t kPairHeader = 1024u + 8192u + 1024u;   // barriers | scale[1024], shift[1024] | routing table

struct PairSmem {
  uint64_t full[8], empty[8];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t bres_bar;
  uint32_t tmem_slot;
  uint32_t item_off16[24];      // descriptor displacement of each item (its tap's row displacement inside the box * 128 bytes >> 4)
};
static_assert(sizeof(PairSmem) <= 1024, "barrier block");

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// Remote arrive WITHOUT cluster-scope release: what it orders (the tcgen05.ld reads of the accumulators) is covered by
// tcgen05.fence::before_thread_sync; a .release.cluster arrive made the lane wait for all of its earlier global stores to be
// performed (MEMBAR.ALL.GPU + ERRBAR, ~6 % of all stall samples) before the accumulator buffer went back to the MMA warp.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes go to a barrier that may live in the PEER CTA of the pair (cluster address).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma2_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Timing experiments (debug library only, SVX_PAIR_KNOCK): 1 no epilogue at all, 2 no global stores, 4 no residual loads.
#ifdef SVX_DEBUG_SWITCHES
#define PKNOCK(bit) ((p.knock & (bit)) != 0)
#else
#define PKNOCK(bit) false
#endif

// Bounded wait (wall time, see umma.cuh); on a timeout the first lane of the warp records (code, counter) before the trap.
__device__ __forceinline__ void wait_pair(uint64_t* bar, uint32_t parity, unsigned long long* dbg, int code, unsigned cnt) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  bool reported = false;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) != 0) continue;
    const unsigned long long now = global_ns();
    if (t0 == 0) { t0 = now; continue; }
    if (!reported && now - t0 > 1000000000ull && dbg && (threadIdx.x & 31) == 0) {
      reported = true;
      dbg[(blockIdx.x & 3) * 16 + (threadIdx.x >> 5)] = (static_cast<unsigned long long>(cnt) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 16) |
                                                         0x8000ull | static_cast<unsigned long long>(code & 0xff);
      __threadfence_system();
    }
    if (now - t0 > kWatchdogNs) __trap();
  }
}

// Two fp32 values -> packed 16-bit pair in ONE instruction (F2FP): round-to-nearest, optional ReLU, and for fp16 the clamp to
// +-65504 that TypeOps<__half>::pack2 spells as two FMNMX per value (the epilogue is bound by its instruction count: 8 warps per
// SM, ~12 instructions per output before this).  Results equal pack2(relu(...)) for every non-NaN input.
template <typename T, bool RELU> struct PackSat;
template <> struct PackSat<__half, true> {
  static __device__ __forceinline__ uint32_t f(float lo, float hi) { uint32_t r; asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
};
template <> struct PackSat<__half, false> {
  static __device__ __forceinline__ uint32_t f(float lo, float hi) { uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
};
template <> struct PackSat<__nv_bfloat16, true> {
  static __device__ __forceinline__ uint32_t f(float lo, float hi) { uint32_t r; asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
};
template <> struct PackSat<__nv_bfloat16, false> {
  static __device__ __forceinline__ uint32_t f(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
};

// 8 consecutive channels: BN scale/shift (shared memory), residual, ReLU (RELU: 0 none, 1 after scale/shift/residual, 2 before
// scale/shift), 16-bit packing.
template <typename T, int AUX, int RELU>
__device__ __forceinline__ uint4 pair_epi8(const uint32_t* r, uint32_t sc, uint32_t sh, const uint4& ax, bool res_here, uint32_t vmask) {
  const float4 s0 = lds_f4(sc), s1 = lds_f4(sc + 16), b0 = lds_f4(sh), b1 = lds_f4(sh + 16);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = __uint_as_float(r[j]);
    if (RELU == 2) v[j] = fmaxf(v[j], 0.f);
  }
  v[0] = fmaf(v[0], s0.x, b0.x); v[1] = fmaf(v[1], s0.y, b0.y); v[2] = fmaf(v[2], s0.z, b0.z); v[3] = fmaf(v[3], s0.w, b0.w);
  v[4] = fmaf(v[4], s1.x, b1.x); v[5] = fmaf(v[5], s1.y, b1.y); v[6] = fmaf(v[6], s1.z, b1.z); v[7] = fmaf(v[7], s1.w, b1.w);
  if (AUX == 1) {
    if (res_here) {
      const float2 a0 = TypeOps<T>::unpack2(ax.x), a1 = TypeOps<T>::unpack2(ax.y), a2 = TypeOps<T>::unpack2(ax.z), a3 = TypeOps<T>::unpack2(ax.w);
      v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y;
    }
  }
  uint4 o;
  o.x = PackSat<T, RELU == 1>::f(v[0], v[1]) & vmask; o.y = PackSat<T, RELU == 1>::f(v[2], v[3]) & vmask;
  o.z = PackSat<T, RELU == 1>::f(v[4], v[5]) & vmask; o.w = PackSat<T, RELU == 1>::f(v[6], v[7]) & vmask;
  return o;
}

// Aux mode 2 (hierarchical 3x3): y = relu(scale * acc + shift) and the second output s = y + x_next, y taken BEFORE its rounding
// to 16 bits (conv_flat.cu epi8, AUX >= 2).
template <typename T>
__device__ __forceinline__ void pair_epi8_two(const uint32_t* r, uint32_t sc, uint32_t sh, const uint4& ax, uint32_t vmask, uint4& o, uint4& o2) {
  const float4 s0 = lds_f4(sc), s1 = lds_f4(sc + 16), b0 = lds_f4(sh), b1 = lds_f4(sh + 16);
  float v[8];
  v[0] = fmaxf(fmaf(__uint_as_float(r[0]), s0.x, b0.x), 0.f); v[1] = fmaxf(fmaf(__uint_as_float(r[1]), s0.y, b0.y), 0.f);
  v[2] = fmaxf(fmaf(__uint_as_float(r[2]), s0.z, b0.z), 0.f); v[3] = fmaxf(fmaf(__uint_as_float(r[3]), s0.w, b0.w), 0.f);
  v[4] = fmaxf(fmaf(__uint_as_float(r[4]), s1.x, b1.x), 0.f); v[5] = fmaxf(fmaf(__uint_as_float(r[5]), s1.y, b1.y), 0.f);
  v[6] = fmaxf(fmaf(__uint_as_float(r[6]), s1.z, b1.z), 0.f); v[7] = fmaxf(fmaf(__uint_as_float(r[7]), s1.w, b1.w), 0.f);
  const float2 a0 = TypeOps<T>::unpack2(ax.x), a1 = TypeOps<T>::unpack2(ax.y), a2 = TypeOps<T>::unpack2(ax.z), a3 = TypeOps<T>::unpack2(ax.w);
  o.x = PackSat<T, false>::f(v[0], v[1]) & vmask; o.y = PackSat<T, false>::f(v[2], v[3]) & vmask;
  o.z = PackSat<T, false>::f(v[4], v[5]) & vmask; o.w = PackSat<T, false>::f(v[6], v[7]) & vmask;
  o2.x = PackSat<T, false>::f(v[0] + a0.x, v[1] + a0.y) & vmask; o2.y = PackSat<T, false>::f(v[2] + a1.x, v[3] + a1.y) & vmask;
  o2.z = PackSat<T, false>::f(v[4] + a2.x, v[5] + a2.y) & vmask; o2.w = PackSat<T, false>::f(v[6] + a3.x, v[7] + a3.y) & vmask;
}

}  // namespace

template <typename T, int AUX, int RELU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ PairConvParams p, const __grid_constant__ PairMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  PairSmem& S = *reinterpret_cast<PairSmem*>(smem);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);
  float* s_shift = s_scale + 1024;
  unsigned long long* s_dst = reinterpret_cast<unsigned long long*>(smem + 1024 + 8192);     // [64] destination of each 16-channel group (0: none)
  uint32_t* s_pitch = reinterpret_cast<uint32_t*>(smem + 1024 + 8192 + 512);                 // [64] its bytes per pixel
  uint8_t* bres_smem = smem + kPairHeader;                                   // resident weights (taps * nkb items), if any
  const int n_items = p.box_item0[p.n_boxes];
  const uint32_t bres_bytes = p.b_resident ? static_cast<uint32_t>(n_items) * p.b_item_bytes : 0u;
  uint8_t* stage_smem = bres_smem + bres_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = static_cast<int>(blockIdx.x >> 1);
  const int n_pairs = static_cast<int>(gridDim.x >> 1);
  const bool tile2d = p.tile_bw != 0;
  const int n_mb = tile2d ? (p.rows + 2 * p.tile_h - 1) / (2 * p.tile_h) : static_cast<int>((p.P + 255) / 256);     // pair tiles along the pixels
  const int n_total = n_mb * p.n_tiles;
  const int half_rows = p.n_tile >> 1;
  const int ng = p.n_tile >> 4;            // 16-channel groups per tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a); prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.a2[0]); prefetch_tmap(&maps.a2[1]); prefetch_tmap(&maps.a2[2]);
    for (int i = 0; i < 8; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], 16); }   // 8 epilogue warps of each CTA
    mbar_init(&S.bres_bar, 1);
    for (int j = 0; j < 24; ++j) S.item_off16[j] = j < n_items ? p.item_off16[j] : 0u;
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_gemm; i += kPairThreads) {
    s_scale[i] = (p.scale && i < p.n_valid) ? p.scale[i] : 1.f;
    s_shift[i] = (p.shift && i < p.n_valid) ? p.shift[i] : 0.f;
  }
  if (threadIdx.x < 64) {
    const int g = threadIdx.x;
    const int r = g < (p.n_gemm >> 4) ? p.route[g] : 0xff;
    s_dst[g] = r == 0xff ? 0ull : reinterpret_cast<unsigned long long>(p.dst_base[r]) + p.goff[g];
    s_pitch[g] = r == 0xff ? 0u : p.dst_pitch[r];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;
  if (p.b_resident && warp == 0 && lane == 0) {      // weights are static: fetched before the dependency wait; both halves complete on the leader's barrier
    const uint32_t bres_leader = mapa_u32(smem_u32(&S.bres_bar), 0);
    const uint32_t item_tx = static_cast<uint32_t>(half_rows) * 128u;
    if (rank == 0) mbar_expect_tx(&S.bres_bar, 2u * item_tx * static_cast<uint32_t>(n_items));
    for (int j = 0; j < n_items; ++j)
      tma_load_2d_pair(bres_smem + static_cast<size_t>(j) * p.b_item_bytes, &maps.b, bres_leader, p.item_wcol[j], static_cast<int>(rank) * half_rows);
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (both CTAs)
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(&S.full[0]), 0);
      const uint32_t tx = static_cast<uint32_t>(p.a_rows) * 128u + (p.b_resident ? 0u : static_cast<uint32_t>(half_rows) * 128u);   // a_rows = box rows in either mode
      uint32_t it = 0;
      for (int t = pair; t < n_total; t += n_pairs) {
        const int mb = t / p.n_tiles, nb = t - mb * p.n_tiles;
        const int ti = (p.reverse ? n_mb - 1 - mb : mb) * 2 + static_cast<int>(rank);         // this CTA's tile along the pixels
        const int nrow = nb * p.n_tile + static_cast<int>(rank) * half_rows;
        for (int b = 0; b < p.n_boxes; ++b, ++it) {
          const uint32_t s = it % static_cast<uint32_t>(p.stages);
          wait_pair(&S.empty[s], ((it / static_cast<uint32_t>(p.stages)) & 1u) ^ 1u, p.dbg, 0x01, it);
          if (rank == 0) mbar_expect_tx(&S.full[s], 2u * tx);          // the leader's barrier collects the bytes of both CTAs
          uint8_t* dst = stage_smem + static_cast<size_t>(s) * p.stage_bytes;
          // haloed span / box: rows before the image or past its end are zero-filled
          if (tile2d) tma_load_3d_pair(dst, p.box_map[b] == 0 ? &maps.a : &maps.a2[p.box_map[b] - 1], full_leader + s * 8u, p.box_c[b], p.a_col0, ti * p.tile_h + p.a_row0);
          else tma_load_2d_pair(dst, &maps.a, full_leader + s * 8u, p.box_c[b], ti * 128 - p.halo);
          if (!p.b_resident) tma_load_2d_pair(dst + p.a_bytes, &maps.b, full_leader + s * 8u, p.item_wcol[p.box_item0[b]], nrow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; converged warp, elected lane issues)
    if (rank == 0) {
      const uint64_t desc_base = make_kmajor_desc(0, 1024u, 2u);       // SWIZZLE_128B, 8-row groups 1024 bytes apart
      const uint32_t hi = static_cast<uint32_t>(desc_base >> 32);
      const uint32_t lo0 = static_cast<uint32_t>(desc_base) + (smem_u32(stage_smem) >> 4);
      const uint32_t stage16 = p.stage_bytes >> 4;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = p.idesc;
      const uint32_t bres_lo = static_cast<uint32_t>(desc_base) + (smem_u32(bres_smem) >> 4);
      const uint32_t a16 = p.a_bytes >> 4, item16 = p.b_item_bytes >> 4;
      if (p.b_resident) wait_pair(&S.bres_bar, 0, p.dbg, 0x10, 0);
      uint32_t it = 0;
      int lt = 0;
      for (int t = pair; t < n_total; t += n_pairs, ++lt) {
        const uint32_t buf = static_cast<uint32_t>(lt) & 1u;
        wait_pair(&S.tmem_empty[buf], ((static_cast<uint32_t>(lt) >> 1) & 1u) ^ 1u, p.dbg, 0x11, lt);
        tc_fence_after();
        const uint32_t d_tmem = tb + buf * 256u;
        int j = 0;
        for (int b = 0; b < p.n_boxes; ++b, ++it) {
          const uint32_t s = it % static_cast<uint32_t>(p.stages);
          const int j_end = p.box_item0[b + 1];
          uint32_t toff = S.item_off16[j];
          int ks = p.item_ks[j];                                         // K = 16 steps of this item that hold real channels
          wait_pair(&S.full[s], (it / static_cast<uint32_t>(p.stages)) & 1u, p.dbg, 0x12, it);
          tc_fence_after();
          const uint32_t a_lo = lo0 + s * stage16;
          uint32_t b_lo = p.b_resident ? bres_lo + static_cast<uint32_t>(j) * item16 : a_lo + a16;
#pragma unroll 1
          for (; j < j_end; ++j) {
            const uint32_t toff_next = S.item_off16[j + 1];              // next item's displacement loads while this one issues
            const int ks_next = p.item_ks[j + 1];
            const uint32_t at = a_lo + toff;
            const uint32_t first = j == 0 ? 0u : 1u;
            if (elect_one()) {
              umma2_lo(d_tmem, at, b_lo, hi, idesc, first);
              if (ks > 1) umma2_lo(d_tmem, at + 2u, b_lo + 2u, hi, idesc, 1u);
              if (ks > 2) umma2_lo(d_tmem, at + 4u, b_lo + 4u, hi, idesc, 1u);
              if (ks > 3) umma2_lo(d_tmem, at + 6u, b_lo + 6u, hi, idesc, 1u);
            }
            b_lo += item16;
            toff = toff_next;
            ks = ks_next;
          }
          if (elect_one()) {
            umma2_commit_mc(&S.empty[s], 3);                             // stage s is free in BOTH CTAs once these MMAs have read it
            if (b == p.n_boxes - 1) umma2_commit_mc(&S.tmem_full[buf], 3);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warpgroups (both CTAs)
    // A thread owns a pixel (= TMEM lane), but 32 lanes touching 32 different pixel rows cost 32 L1 wavefronts per 1 KB
    // (measured: 5-9 us per tile, the bound of the first version).  So every warp transposes its 32 pixels x 64 channels
    // through a private 4 KB staging tile ([pixel][128 bytes], 16-byte chunks XOR-swizzled by the row): residual values come in
    // and results go out with lanes mapped (pixel = 8 i + lane / 4, 32-byte chunk = lane % 4), i.e. 8 wavefronts per 1 KB.
    const int wg = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const uint32_t empty_leader = mapa_u32(smem_u32(&S.tmem_empty[0]), 0);
    const uint32_t sc_base = smem_u32(s_scale), sh_base = smem_u32(s_shift);
    const int n_res_grp = p.n_res >> 4;
    const uint32_t st_base = smem_u32(stage_smem + static_cast<size_t>(p.stages) * p.stage_bytes) + static_cast<uint32_t>(warp - 4) * 4096u;
    const uint32_t my_row = st_base + static_cast<uint32_t>(lane) * 128u;
    const uint32_t my_x = static_cast<uint32_t>(lane & 7);
    const int cpx = lane >> 2, c32 = lane & 3;
    const uint32_t co_off0 = static_cast<uint32_t>(cpx) * 128u + ((static_cast<uint32_t>(2 * c32) ^ static_cast<uint32_t>(cpx)) << 4);       // row 8 i + cpx: + i * 1024
    const uint32_t co_off1 = static_cast<uint32_t>(cpx) * 128u + ((static_cast<uint32_t>(2 * c32 + 1) ^ static_cast<uint32_t>(cpx)) << 4);
    const int g_first = wg ? (ng + 1) >> 1 : 0;          // this warpgroup's 16-channel groups of the tile: the first or the second half
    const int ngh = wg ? ng >> 1 : (ng + 1) >> 1;
    const int nchunk = (ngh + 3) >> 2;
    // Residual values of chunk c of tile t, in the coalesced mapping.  They are loaded ONE TILE AHEAD: as soon as chunk c of the
    // current tile has been copied into the staging tile, its registers take the same chunk of the pair's next tile, so the
    // DRAM latency of the residual never sits between two tiles of an epilogue warp.
    uint4 rx[2][8];
    auto load_res = [&](int t, int c) {
      const int mb = t / p.n_tiles, nb = t - mb * p.n_tiles;
      const long long px0 = static_cast<long long>(p.reverse ? n_mb - 1 - mb : mb) * 256 + static_cast<int>(rank) * 128 + q4 * 32;   // flat mode only (no aux tile in 2-D tile mode)
      const int gl = c * 4 + c32, G = nb * ng + g_first + gl;
      const bool here = t < n_total && gl < ngh && G < n_res_grp && !PKNOCK(4);
      const uint8_t* rsrc = p.res + static_cast<size_t>(G) * 32u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long px = px0 + 8 * i + cpx;
        if (here && px < p.P_cap) ldg256(rsrc + static_cast<size_t>(px) * p.res_pitch, rx[c][2 * i], rx[c][2 * i + 1]);
        else rx[c][2 * i] = rx[c][2 * i + 1] = make_uint4(0, 0, 0, 0);
      }
    };
    if (AUX != 0) { load_res(pair, 0); load_res(pair, 1); }
    int lt = 0;
    for (int t = pair; t < n_total; t += n_pairs, ++lt) {
      const int mb = t / p.n_tiles, nb = t - mb * p.n_tiles;
      const int ti = (p.reverse ? n_mb - 1 - mb : mb) * 2 + static_cast<int>(rank);           // this CTA's tile along the pixels
      // accumulator row m of the tile -> output pixel (flat mode: consecutive pixels; 2-D tile mode: tile_h output rows laid out with
      // the box pitch tile_bw, the columns past the output row pitch and the rows past tile_h are not output pixels)
      auto pix_of = [&](int m, long long& q) -> bool {
        if (!tile2d) { q = static_cast<long long>(ti) * 128 + m; return true; }
        const int r = m / p.tile_bw, c = m - r * p.tile_bw;
        q = static_cast<long long>(ti * p.tile_h + r) * p.out_wp + c;
        return r < p.tile_h && c < p.out_wp;
      };
      long long pp;
      const bool own = pix_of(q4 * 32 + lane, pp);
      const uint32_t vmask = (own && pp < p.P && p.pix_valid[pp]) ? 0xffffffffu : 0u;
      long long cq[4];                                     // output pixels of the four rows this lane stores in the coalesced mapping
      bool cok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { cok[i] = pix_of(q4 * 32 + 8 * i + cpx, cq[i]); cok[i] = cok[i] && cq[i] < p.P_cap; }
      const int G0 = nb * ng + g_first;                  // first global 16-channel group of this warpgroup
      const uint32_t buf = static_cast<uint32_t>(lt) & 1u;
      wait_pair(&S.tmem_full[buf], (static_cast<uint32_t>(lt) >> 1) & 1u, p.dbg, 0x41, lt);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * 256u + static_cast<uint32_t>(g_first) * 16u;
      if (PKNOCK(1)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(empty_leader + buf * 8u);
        continue;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nchunk) {
          const int gvalid = min(4, ngh - 4 * c);
          uint32_t acc[2][16];
          tmem_ld16(taddr + static_cast<uint32_t>(c * 64), acc[0]);
          if (AUX != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sts_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off0, rx[c][2 * i]);
              sts_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off1, rx[c][2 * i + 1]);
            }
            __syncwarp();
            load_res(t + n_pairs, c);
          }
          uint4 s2[8];                                       // aux mode 2: the second output of this chunk, stored after y has left the staging tile
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < gvalid) {
              tmem_ld_wait();
              if (g + 1 < gvalid) tmem_ld16(taddr + static_cast<uint32_t>(c * 64 + (g + 1) * 16), acc[(g + 1) & 1]);
              const int G = G0 + c * 4 + g;
              const uint32_t sc = sc_base + static_cast<uint32_t>(G) * 64u, sh = sh_base + static_cast<uint32_t>(G) * 64u;
              const bool res_here = AUX == 1 && G < n_res_grp;
              const uint32_t a0 = my_row + ((static_cast<uint32_t>(2 * g) ^ my_x) << 4), a1 = my_row + ((static_cast<uint32_t>(2 * g + 1) ^ my_x) << 4);
              uint4 ax0 = make_uint4(0, 0, 0, 0), ax1 = ax0;
              if (AUX != 0) { ax0 = lds_u4(a0); ax1 = lds_u4(a1); }
              uint4 o0, o1;
              if (AUX == 2) {
                pair_epi8_two<T>(acc[g & 1], sc, sh, ax0, vmask, o0, s2[2 * g]);
                pair_epi8_two<T>(acc[g & 1] + 8, sc + 32, sh + 32, ax1, vmask, o1, s2[2 * g + 1]);
              } else {
                o0 = pair_epi8<T, AUX, RELU>(acc[g & 1], sc, sh, ax0, res_here, vmask);
                o1 = pair_epi8<T, AUX, RELU>(acc[g & 1] + 8, sc + 32, sh + 32, ax1, res_here, vmask);
              }
              sts_u4(a0, o0);
              sts_u4(a1, o1);
            }
          }
          if (c == nchunk - 1) tc_fence_before();          // all accumulators of this tile have been read
          __syncwarp();
          if (c == nchunk - 1 && lane == 0) mbar_arrive_cluster(empty_leader + buf * 8u);     // the leader's MMA warp waits for the epilogue warps of BOTH CTAs
          if (c32 < gvalid) {
            const int G = G0 + c * 4 + c32;
            const unsigned long long d = s_dst[G];
            const unsigned long long pitch = s_pitch[G];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 v0 = lds_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off0);
              const uint4 v1 = lds_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off1);
              if (d != 0ull && cok[i] && !PKNOCK(2)) stg256(reinterpret_cast<void*>(d + static_cast<unsigned long long>(cq[i]) * pitch), v0, v1);
            }
          }
          __syncwarp();                                      // the staging tile is reused by the next chunk / tile
          if (AUX == 2) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g < gvalid) {
                sts_u4(my_row + ((static_cast<uint32_t>(2 * g) ^ my_x) << 4), s2[2 * g]);
                sts_u4(my_row + ((static_cast<uint32_t>(2 * g + 1) ^ my_x) << 4), s2[2 * g + 1]);
              }
            }
            __syncwarp();
            if (c32 < gvalid) {
              uint8_t* d2 = p.out2 + static_cast<size_t>(G0 + c * 4 + c32) * 32u;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 v0 = lds_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off0);
                const uint4 v1 = lds_u4(st_base + static_cast<uint32_t>(i) * 1024u + co_off1);
                if (cok[i] && !PKNOCK(2)) stg256(d2 + static_cast<size_t>(cq[i]) * p.out2_pitch, v0, v1);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the leader multicasts commits into the peer's barriers: nobody leaves early
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

size_t conv_pair_smem_bytes(const PairConvParams& p) {
  const size_t bres = p.b_resident ? static_cast<size_t>(p.box_item0[p.n_boxes]) * p.b_item_bytes : 0;
  return 1024 + kPairHeader + bres + static_cast<size_t>(p.stages) * p.stage_bytes + 8 * 4096;
}   // + one 4 KB staging tile per epilogue warp

static int g_pair_clusters = 0;

template <typename T, int AUX, int RELU>
static cudaError_t pair_attr() { return cudaFuncSetAttribute(conv_pair_kernel<T, AUX, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }
template <typename T>
static cudaError_t pair_attr_type() {
  cudaError_t e;
  if ((e = pair_attr<T, 0, 0>()) != cudaSuccess) return e;
  if ((e = pair_attr<T, 0, 1>()) != cudaSuccess) return e;
  if ((e = pair_attr<T, 0, 2>()) != cudaSuccess) return e;
  if ((e = pair_attr<T, 1, 0>()) != cudaSuccess) return e;
  if ((e = pair_attr<T, 1, 1>()) != cudaSuccess) return e;
  return pair_attr<T, 2, 1>();
}

cudaError_t conv_pair_init() {
  cudaError_t e;
  if ((e = pair_attr_type<__half>()) != cudaSuccess) return e;
  if ((e = pair_attr_type<__nv_bfloat16>()) != cudaSuccess) return e;
  // resident pairs: one CTA per SM (every plan takes most of the shared memory), pairs never straddle a TPC
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(2 * 128, 1, 1); cfg.blockDim = dim3(kPairThreads, 1, 1); cfg.dynamicSmemBytes = 200 * 1024;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_pair_kernel<__half, 0, 1>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    n = sms / 2;
  }
  g_pair_clusters = n;
  if (dbg_env("SVX_PLAN_LOG")) fprintf(stderr, "conv_pair: %d resident CTA pairs\n", n);
  return cudaSuccess;
}

cudaError_t launch_conv_pair(const PairConvParams& p, const PairMaps& maps, int is_bf16, cudaStream_t st) {
  if (p.P <= 0) return cudaSuccess;
  const long long n_mb = p.tile_bw ? (p.rows + 2 * p.tile_h - 1) / (2 * p.tile_h) : (p.P + 255) / 256;
  const long long n_total = n_mb * p.n_tiles;
  long long pairs = g_pair_clusters > 0 ? g_pair_clusters : 74;
  if (pairs > n_total) pairs = n_total;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs), 1, 1);
  cfg.blockDim = dim3(kPairThreads, 1, 1);
  cfg.dynamicSmemBytes = conv_pair_smem_bytes(p);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (p.pre_relu && (p.post_relu || p.aux_mode)) return cudaErrorInvalidValue;
  if (p.aux_mode == 2 && !p.post_relu) return cudaErrorInvalidValue;
  const int relu = p.pre_relu ? 2 : p.post_relu ? 1 : 0;
  cudaError_t le = cudaErrorInvalidValue;
#define SVX_PAIR(T)                                                                                     \
  do {                                                                                                  \
    if (p.aux_mode == 2) le = cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 2, 1>, p, maps);             \
    else if (p.aux_mode) le = relu == 1 ? cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 1, 1>, p, maps)       \
                                   : cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 1, 0>, p, maps);      \
    else le = relu == 1 ? cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 0, 1>, p, maps)                  \
            : relu == 2 ? cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 0, 2>, p, maps)                  \
                        : cudaLaunchKernelEx(&cfg, conv_pair_kernel<T, 0, 0>, p, maps);                 \
  } while (0)
  if (is_bf16) SVX_PAIR(__nv_bfloat16); else SVX_PAIR(__half);
#undef SVX_PAIR
  if (le != cudaSuccess) {
    fprintf(stderr, "conv_pair launch failed: grid %u smem %zu n_tile %d x%d boxes %d stages %d\n", cfg.gridDim.x, cfg.dynamicSmemBytes, p.n_tile, p.n_tiles, p.n_boxes, p.stages);
    return le;
  }
  return cudaGetLastError();
}

}  // namespace svx
