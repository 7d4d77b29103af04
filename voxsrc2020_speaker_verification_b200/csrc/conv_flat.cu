// Flat implicit-GEMM convolution for every stride-1 layer (1x1, 3x3, dilated kx1) on tcgen05 tensor cores.
//
// The NHWC image is a 1-D sequence of pixels p = row*Wp + col with one all-zero column per row (Wp = W+1) and
// all-zero gap rows between segments, so filter tap (dh, dw) is the constant shift s = dh*Wp + dw and
//
//   D[p, n] = sum_tap sum_k  X[p + s_tap, k] * Wt[n, tap, k]
//
// A span of mt*128 pixels plus a halo of max|s| pixels on each side is loaded ONCE per K-box by TMA (2-D boxes
// {kbox channels, <=256 pixels}, swizzled rows = the canonical K-major UMMA layout); each tap is then just a
// shared-memory descriptor whose start address is displaced by s rows.  (The swizzle is a function of the
// absolute shared-memory address, so a row-displaced descriptor with base_offset 0 addresses the tile exactly as
// TMA wrote it — profiles/r01_experiment_shifted_descriptor.txt.)  Compared with one 128-pixel box per tap this
// divides the TMA row requests of a 3x3 conv by ~7 (the rate that bounded the previous kernel).
//
// Persistent warp-specialised CTA, 384 threads, one per SM; spans dealt round-robin:
//   warp 0   producer: A spans (ring of a_stages) and weight boxes (resident in smem when small, else a ring)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer; 2 accumulator sets x mt sub-tiles in TMEM
//   warp 2   aux producer: residual / add2 tiles into the epilogue slots, several sub-tiles ahead
//   warp 3   store issuer: TMA stores of finished slots, releases slots when the stores have read them
//   warps 4-7 / 8-11   two epilogue warpgroups (even / odd spans): tcgen05.ld -> scale/shift, ReLU, residual,
//            second output (x_{i+1} + o_i) -> swizzled smem boxes (in place over the aux tile)
// The epilogue body is specialised at compile time on <aux mode, pre-ReLU, post-ReLU>; the previous generic
// body cost ~35 SASS instructions per output value and bounded every 1x1 conv.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

namespace {

constexpr int kFlatThreads = 384;
constexpr int kRing = 8;
constexpr uint32_t kHeaderBytes = 3072;   // barriers + progress words (1024) + scale/shift of one n-tile (2 x 256 floats)

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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 1-D bulk copies (no tensor map, no rows): a 128-pixel tile of a DENSE planar tensor is one contiguous run in global memory.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

// Timing experiments only (tools/build_knock.sh builds a second library with -DSVX_KNOCK): parts of the pipeline can be
// switched off at run time to see what the MMA phase costs without them.  The production build compiles these out.
#ifdef SVX_KNOCK
#define KNOCK(bit) ((p.knock & (bit)) != 0)
#else
#define KNOCK(bit) false
#endif

// Bounded wait that says which barrier starved before trapping: code = role*16 + barrier kind, a = ring index, b = counter.
// On a timeout the thread also copies the CTA's 16 role progress words (prog) so the host can see where every role stood.
__device__ __forceinline__ void wait_dbg(uint64_t* bar, uint32_t parity, unsigned long long* dbg, int code, int a, unsigned b,
                                         const volatile uint32_t* prog) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  bool reported = false;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) != 0) continue;
    const unsigned long long now = global_ns();       // wall-time watchdog (see umma.cuh): report after 2 s, trap after kWatchdogNs
    if (t0 == 0) { t0 = now; continue; }
    if (!reported && now - t0 > 2000000000ull && dbg) {      // first report, then keep waiting so that every starved role can report
      reported = true;
      if ((threadIdx.x & 31) == 0 || (threadIdx.x >> 5) >= 4) {
        dbg[(blockIdx.x & 3) * 16 + (threadIdx.x >> 5)] = (static_cast<unsigned long long>(b) << 32) | (static_cast<unsigned long long>(blockIdx.x) << 16) |
                                                        (static_cast<unsigned long long>(a & 0xff) << 8) | static_cast<unsigned long long>(code & 0xff);
        if ((threadIdx.x & 31) == 0)
          for (int i = 0; i < 16; ++i) dbg[64 + (blockIdx.x & 3) * 16 + i] = (static_cast<unsigned long long>(blockIdx.x) << 32) | prog[i];
        __threadfence_system();
      }
    }
    if (now - t0 > kWatchdogNs) __trap();
  }
}

struct Tracer {
  unsigned long long* buf; int n;
  __device__ __forceinline__ void init(unsigned long long* base, int role) { buf = (base && blockIdx.x == 0) ? base + role * kTraceEvents : nullptr; n = 0; }
  __device__ __forceinline__ void ev(int code) {
    if (buf && n < kTraceEvents) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      buf[n++] = (t << 8) | static_cast<unsigned long long>(code & 0xff);
    }
  }
};

// 8 consecutive channels of one pixel.  v: accumulators; sc/sh: shared addresses of scale/shift for these channels.
template <typename T, int AUX, bool PRE, bool POST>
__device__ __forceinline__ void epi8(const uint32_t* r, uint32_t sc, uint32_t sh, uint32_t ua, uint32_t ub, bool res_here, uint32_t vmask) {
  const float4 s0 = lds_f4(sc), s1 = lds_f4(sc + 16), b0 = lds_f4(sh), b1 = lds_f4(sh + 16);
  uint4 ax = make_uint4(0, 0, 0, 0);
  if (AUX == 1) { if (res_here) ax = lds_u4(ua); }
  if (AUX >= 2) ax = lds_u4(ub);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = __uint_as_float(r[j]);
    if (PRE) v[j] = fmaxf(v[j], 0.f);
  }
  v[0] = fmaf(v[0], s0.x, b0.x); v[1] = fmaf(v[1], s0.y, b0.y); v[2] = fmaf(v[2], s0.z, b0.z); v[3] = fmaf(v[3], s0.w, b0.w);
  v[4] = fmaf(v[4], s1.x, b1.x); v[5] = fmaf(v[5], s1.y, b1.y); v[6] = fmaf(v[6], s1.z, b1.z); v[7] = fmaf(v[7], s1.w, b1.w);
  float2 a0, a1, a2, a3;
  if (AUX != 0) { a0 = TypeOps<T>::unpack2(ax.x); a1 = TypeOps<T>::unpack2(ax.y); a2 = TypeOps<T>::unpack2(ax.z); a3 = TypeOps<T>::unpack2(ax.w); }
  if (AUX == 1) { v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y; }
  if (POST) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  uint4 o;
  o.x = TypeOps<T>::pack2(v[0], v[1]) & vmask; o.y = TypeOps<T>::pack2(v[2], v[3]) & vmask;
  o.z = TypeOps<T>::pack2(v[4], v[5]) & vmask; o.w = TypeOps<T>::pack2(v[6], v[7]) & vmask;
  sts_u4(ua, o);
  if (AUX >= 2) {   // second output = v + add2, in place over the add2 tile
    uint4 o2;
    o2.x = TypeOps<T>::pack2(v[0] + a0.x, v[1] + a0.y) & vmask; o2.y = TypeOps<T>::pack2(v[2] + a1.x, v[3] + a1.y) & vmask;
    o2.z = TypeOps<T>::pack2(v[4] + a2.x, v[5] + a2.y) & vmask; o2.w = TypeOps<T>::pack2(v[6] + a3.x, v[7] + a3.y) & vmask;
    sts_u4(ub, o2);
  }
}

// 256-bit global accesses (sm_100: LDG/STG.256): one full 32-byte sector per lane, half the LSU requests of two 128-bit ones.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// The same 8 channels with the aux values already in registers and the results written straight to global memory.
template <typename T, int AUX, bool PRE, bool POST>
__device__ __forceinline__ void epi8_direct(const uint32_t* r, uint32_t sc, uint32_t sh, const uint4& ax, uint4& o, uint4& o2, uint32_t vmask) {
  const float4 s0 = lds_f4(sc), s1 = lds_f4(sc + 16), b0 = lds_f4(sh), b1 = lds_f4(sh + 16);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = __uint_as_float(r[j]);
    if (PRE) v[j] = fmaxf(v[j], 0.f);
  }
  v[0] = fmaf(v[0], s0.x, b0.x); v[1] = fmaf(v[1], s0.y, b0.y); v[2] = fmaf(v[2], s0.z, b0.z); v[3] = fmaf(v[3], s0.w, b0.w);
  v[4] = fmaf(v[4], s1.x, b1.x); v[5] = fmaf(v[5], s1.y, b1.y); v[6] = fmaf(v[6], s1.z, b1.z); v[7] = fmaf(v[7], s1.w, b1.w);
  float2 a0, a1, a2, a3;
  if (AUX != 0) { a0 = TypeOps<T>::unpack2(ax.x); a1 = TypeOps<T>::unpack2(ax.y); a2 = TypeOps<T>::unpack2(ax.z); a3 = TypeOps<T>::unpack2(ax.w); }
  if (AUX == 1) { v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y; }
  if (POST) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  o.x = TypeOps<T>::pack2(v[0], v[1]) & vmask; o.y = TypeOps<T>::pack2(v[2], v[3]) & vmask;
  o.z = TypeOps<T>::pack2(v[4], v[5]) & vmask; o.w = TypeOps<T>::pack2(v[6], v[7]) & vmask;
  if (AUX >= 2) {
    o2.x = TypeOps<T>::pack2(v[0] + a0.x, v[1] + a0.y) & vmask; o2.y = TypeOps<T>::pack2(v[2] + a1.x, v[3] + a1.y) & vmask;
    o2.z = TypeOps<T>::pack2(v[4] + a2.x, v[5] + a2.y) & vmask; o2.w = TypeOps<T>::pack2(v[6] + a3.x, v[7] + a3.y) & vmask;
  }
}

// All MMAs of one filter tap: MT sub-tiles x KS K-steps, fully unrolled so that every descriptor is the tap's base
// (one value) plus a compile-time constant.  A generic loop costs ~150 cycles of issue per MMA (descriptor arithmetic in
// vector registers + 5 R2UR each, tools/microbench/mma_rate.cu) against 16-64 cycles of tensor-pipe time.
template <int MT, int KS>
__device__ __forceinline__ void issue_tap(uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t d_tmem, uint32_t n_tile, uint32_t idesc, uint32_t first,
                                          int knock = 0) {
  constexpr uint32_t kSub16 = 256u * KS;      // (128 rows x 32*KS bytes) >> 4
  if (elect_one()) {      // one elected region for the whole tap: the R2UR traffic of consecutive MMAs overlaps
#pragma unroll
    for (int j = 0; j < MT; ++j) {
#pragma unroll
      for (int k = 0; k < KS; ++k)
#ifdef SVX_KNOCK
        if (!(knock & 32))
#endif
        umma_lo(d_tmem + j * n_tile, a_lo + (j * kSub16 + k * 2), b_lo + k * 2, hi, idesc, k == 0 ? first : 1u);
    }
  }
}

struct Smem {
  uint64_t a_full[kRing], a_empty[kRing], b_full[kRing], b_empty[kRing];
  uint64_t slot_full[kRing], slot_empty[kRing], slot_ready[kRing];
  uint64_t tmem_full[4], tmem_empty[4];   // accumulator buffers: 2, or 4 when 4 * mt * n_tile columns fit (p.tmem_bufs)
  uint64_t bres_bar;
  uint32_t tmem_slot;
  uint32_t tapoff16[12];        // ((halo + tap_shift[tap]) * row_bytes) >> 4: descriptor displacement of each filter tap
  volatile uint32_t prog[16];   // debug: progress of each role (warp), dumped by wait_dbg on a timeout
};
static_assert(sizeof(Smem) <= 1024, "barrier block");

// MMA issuer role (warp 1), one instantiation per <sub-tiles per span, K-steps per box>.
// The whole warp runs the loop CONVERGED and computes the descriptors (uniform values); only the tcgen05 instructions
// themselves sit under the elected-lane branch.  Computing them inside an `if (lane == 0)` region made ptxas wrap every
// UTCHMMA in an ELECT / 5x R2UR.BROADCAST waterfall: ~200 ns per MMA whatever its shape.
//
// Multicast clusters (MC): an A stage is filled by all cn CTAs that share the span and a streamed weight item by all cm CTAs
// that share the n-tile, so a stage may be refilled only when every one of them has consumed it: the "empty" commits are
// multicast to the same barrier in all CTAs of mask_a / mask_b (their barriers count cn / cm arrivals).
template <int MT, int KS, bool MC>
__device__ __forceinline__ void mma_role(const FlatConvParams& p, Smem& S, uint32_t tb, uint32_t a_base, uint32_t b_base, int my_group, int n_units,
                                         int groups, uint16_t mask_a, uint16_t mask_b, int lane, Tracer& tr) {
  if (p.b_resident) wait_dbg(&S.bres_bar, 0, p.dbg, 0x10, 0, 0, S.prog);
  const uint64_t desc_base = make_kmajor_desc(0, p.sbo, p.layout_type);
  const uint32_t hi = static_cast<uint32_t>(desc_base >> 32);
  const uint32_t lo0 = static_cast<uint32_t>(desc_base);           // LBO field; the address field is added below (smem < 256 KB: no carry out of it)
  const uint32_t a_lo_base = lo0 + (a_base >> 4), b_lo_base = lo0 + (b_base >> 4);
  const uint32_t a_stage16 = p.a_stage_bytes >> 4, b_item16 = p.b_item_bytes >> 4;
  const uint32_t nt = static_cast<uint32_t>(p.n_tile);
  const uint32_t idesc = p.idesc;
  const int taps = p.taps, nkc = p.nkc;
  const bool b_res = p.b_resident != 0;
  uint32_t ia = 0, ib = 0;
  int ls = 0;
  for (int unit = my_group; unit < n_units; unit += groups, ++ls) {
    const int buf = ls & (p.tmem_bufs - 1);       // span ls -> buffer ls mod bufs; warpgroup ls & 1 reads it (buffers wg, wg + 2)
    tr.ev(1);
    if (lane == 0) S.prog[1] = ls;
    wait_dbg(&S.tmem_empty[buf], ((ls >> p.tmem_bufs_log2) & 1) ^ 1, p.dbg, 0x11, buf, ls, S.prog);
    tr.ev(2);
    const uint32_t d_tmem = tb + static_cast<uint32_t>(buf * MT) * nt;
    uint32_t first = 0u;
    for (int kc = 0; kc < nkc; ++kc, ++ia) {
      const int sa = ia % p.a_stages;
      const uint32_t a_par = (ia / p.a_stages) & 1;
      uint32_t toff = S.tapoff16[0];                       // fetched before the wait: off the critical path
      wait_dbg(&S.a_full[sa], a_par, p.dbg, 0x12, sa, ia, S.prog);
      if (kc == 0) tr.ev(3);
      tc_fence_after();
      const uint32_t a_lo = a_lo_base + static_cast<uint32_t>(sa) * a_stage16;
      if (b_res) {
        uint32_t b_lo = b_lo_base + static_cast<uint32_t>(kc * taps) * b_item16;
#pragma unroll 1
        for (int tap = 0; tap < taps; ++tap, b_lo += b_item16) {
          const uint32_t toff_next = S.tapoff16[tap + 1];   // next tap's displacement loads while this tap issues
          issue_tap<MT, KS>(a_lo + toff, b_lo, hi, d_tmem, nt, idesc, first, p.knock);
          first = 1u;
          toff = toff_next;
        }
      } else {
#pragma unroll 1
        for (int tap = 0; tap < taps; ++tap, ++ib) {
          const int sb = ib % p.b_stages;
          const uint32_t b_par = (ib / p.b_stages) & 1;
          const uint32_t toff_next = S.tapoff16[tap + 1];
          wait_dbg(&S.b_full[sb], b_par, p.dbg, 0x13, sb, ib, S.prog);
          tc_fence_after();
          issue_tap<MT, KS>(a_lo + toff, b_lo_base + static_cast<uint32_t>(sb) * b_item16, hi, d_tmem, nt, idesc, first, p.knock);
          first = 1u;
          if (elect_one()) { if constexpr (MC) umma_commit_mc(&S.b_empty[sb], mask_b); else umma_commit(&S.b_empty[sb]); }
          toff = toff_next;
        }
      }
      if (elect_one()) {
        if constexpr (MC) umma_commit_mc(&S.a_empty[sa], mask_a); else umma_commit(&S.a_empty[sa]);
        if (kc == nkc - 1) umma_commit(&S.tmem_full[buf]);
      }
    }
    tr.ev(4);
  }
}

}  // namespace

// MC (multicast clusters) is a template parameter so that the plain instantiation carries no cluster instruction at all and
// launches without a cluster attribute.  In an MC launch a cluster is cm x cn CTAs (rank = im * cn + in): the cn CTAs of a row
// work on the SAME span with different n-tiles, each loads 1/cn of every A stage and multicasts it to the row; the cm CTAs of a
// column work on consecutive spans with the SAME n-tile, each loads 1/cm of every streamed weight item and multicasts it to the
// column.  L2 reads per CTA drop to A/cn + B/cm — the L2 -> SM path (~6.5 TB/s for the whole chip, the same as HBM) is what
// bounded the deep stages, where A was re-read once per n-tile and the weights once per span.
template <typename T, int AUX, bool PRE, bool POST, bool MC>
__global__ void __launch_bounds__(kFlatThreads, 1)
conv_flat_kernel(const __grid_constant__ FlatConvParams p, const __grid_constant__ FlatMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  Smem& S = *reinterpret_cast<Smem*>(smem);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);
  float* s_shift = s_scale + 256;
  const int total_items = p.taps * p.nkc;
  const uint32_t b_bytes = p.b_resident ? static_cast<uint32_t>(total_items) * p.b_item_bytes : static_cast<uint32_t>(p.b_stages) * p.b_item_bytes;
  uint8_t* b_smem = smem + kHeaderBytes;
  uint8_t* slot_smem = b_smem + b_bytes;
  uint8_t* a_smem = slot_smem + 2 * static_cast<size_t>(p.slots) * p.slot_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int span_px = p.mt * 128;
  constexpr int n_parts = 1;
  const int n_spans = static_cast<int>((p.P + span_px - 1) / span_px);
  // scheduling unit = one cluster (one CTA without MC): it takes cm consecutive spans per step (CTA row im takes span cm*u + im)
  // and cn consecutive n-tiles (CTA column in); clusters are dealt over n_tiles/cn n-tile columns x `groups` span groups
  const int cn = MC ? p.cn : 1, cm = MC ? p.cm : 1;
  uint32_t rank = 0u;
  if constexpr (MC) rank = cluster_ctarank();
  const int in = static_cast<int>(rank) % cn, im = static_cast<int>(rank) / cn;
  const int cluster_id = static_cast<int>(blockIdx.x) / (cn * cm);
  const int n_clusters = static_cast<int>(gridDim.x) / (cn * cm);
  const int n_cols = p.n_tiles / cn;
  const int sstride = cm;
  const int n_units = (n_spans + sstride - 1) / sstride;
  const int n_spans_all = n_units * sstride;          // includes phantom spans past the end (fully masked, loads zero-filled)
  const int groups = n_clusters / n_cols;
  const int my_group = cluster_id / n_cols;
  const int n_blk = (cluster_id % n_cols) * cn + in;
  const int n0 = n_blk * p.n_tile;
  // CTAs that receive the A slices this CTA loads (its row) and the weight slices it loads (its column)
  const uint16_t mask_a = static_cast<uint16_t>(((1u << cn) - 1u) << (im * cn));
  uint16_t mask_b = 0;
  for (int j = 0; j < cm; ++j) mask_b |= static_cast<uint16_t>(1u << (j * cn + in));
  const uint32_t row_bytes = static_cast<uint32_t>(p.kbox) * 2u;
  const uint32_t box_bytes = 128u * static_cast<uint32_t>(p.box_ch) * 2u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a); prefetch_tmap(&maps.b);
    if (AUX) prefetch_tmap(&maps.aux);
    prefetch_tmap(&maps.o[0]);
    if (AUX >= 2) prefetch_tmap(&maps.o2);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&S.a_full[i], 1); mbar_init(&S.a_empty[i], cn);     // a stage is free when every CTA of the row has consumed it
      mbar_init(&S.b_full[i], 1); mbar_init(&S.b_empty[i], cm);     // a weight item when every CTA of the column has
      mbar_init(&S.slot_full[i], 1); mbar_init(&S.slot_empty[i], 1); mbar_init(&S.slot_ready[i], 128);
    }
    for (int b = 0; b < 4; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], 4); }
    mbar_init(&S.bres_bar, 1);
    for (int t = 0; t < 12; ++t) S.tapoff16[t] = t < p.taps ? (static_cast<uint32_t>(p.halo + p.tap_shift[t]) * row_bytes) >> 4 : 0u;
    fence_barrier_init();
    if (p.b_resident) {     // weights are static: fetch them before the dependency wait
      const uint32_t bb = static_cast<uint32_t>(p.b_rows) * row_bytes;
      mbar_expect_tx(&S.bres_bar, bb * total_items);
      for (int kc = 0; kc < p.nkc; ++kc)
        for (int tap = 0; tap < p.taps; ++tap)
          tma_load_2d(b_smem + static_cast<size_t>(kc * p.taps + tap) * p.b_item_bytes, &maps.b, &S.bres_bar, tap * p.kpad + kc * p.kbox, n0);
    }
  }
  if (warp == 1) {
    tmem_alloc(&S.tmem_slot, p.tmem_cols); tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.n_tile; i += kFlatThreads) {
    const int c = n0 + i;
    s_scale[i] = (p.scale && c < p.n_valid) ? p.scale[c] : 1.f;
    s_shift[i] = (p.shift && c < p.n_valid) ? p.shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();          // the peers' barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;
  // Programmatic dependent launch: the next conv's CTAs may be scheduled as soon as SMs free up and run their own prologue
  // (barriers, TMEM, resident weights, scale/shift — all static data) under this grid's tail; everything that touches
  // activations waits here for the previous grid to have completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: A spans + weights
    if (lane == 0) {
      const uint32_t b_box_bytes = static_cast<uint32_t>(p.b_rows) * row_bytes;
      const uint32_t a_tx = static_cast<uint32_t>(p.a_boxes) * p.a_box_rows * row_bytes;
      // MC: this CTA loads rows [in * a_slice_rows, +a_slice_rows) of every A box and rows [im * b_slice_rows, +b_slice_rows) of
      // every weight item, into the same offsets of every CTA of its row / column (the full barrier of each receiver, armed by
      // its own producer for the whole stage, collects the bytes of all slices)
      const uint32_t a_slice_off = static_cast<uint32_t>(in) * p.a_slice_rows * row_bytes;
      const uint32_t b_slice_off = static_cast<uint32_t>(im) * p.b_slice_rows * row_bytes;
      Tracer tr; tr.init(p.trace, 0);
      uint32_t ia = 0, ib = 0;
      for (int unit = my_group; unit < n_units; unit += groups) {
        const int span = unit * sstride + im;
        const int p0 = (p.reverse ? n_spans_all - 1 - span : span) * span_px;
        tr.ev(1);
        S.prog[0] = ia;
        for (int kc = 0; kc < p.nkc; ++kc, ++ia) {
          const int sa = ia % p.a_stages;
          wait_dbg(&S.a_empty[sa], ((ia / p.a_stages) & 1) ^ 1, p.dbg, 0x01, sa, ia, S.prog);
          mbar_expect_tx(&S.a_full[sa], a_tx);
          uint8_t* dst = a_smem + static_cast<size_t>(sa) * p.a_stage_bytes;
          for (int b = 0; b < p.a_boxes; ++b) {
            if constexpr (MC)
              tma_load_2d_mc(dst + static_cast<size_t>(b) * p.a_box_rows * row_bytes + a_slice_off, &maps.a, &S.a_full[sa],
                             kc * p.kbox + n_blk * p.a_c_step, p0 - p.halo + b * p.a_box_rows + in * p.a_slice_rows, mask_a);
            else
              tma_load_2d(dst + static_cast<size_t>(b) * p.a_box_rows * row_bytes, &maps.a, &S.a_full[sa], kc * p.kbox + n_blk * p.a_c_step,
                          p0 - p.halo + b * p.a_box_rows);
          }
          if (!p.b_resident) {
            for (int tap = 0; tap < p.taps; ++tap, ++ib) {
              const int sb = ib % p.b_stages;
              wait_dbg(&S.b_empty[sb], ((ib / p.b_stages) & 1) ^ 1, p.dbg, 0x02, sb, ib, S.prog);
              mbar_expect_tx(&S.b_full[sb], b_box_bytes);
              if constexpr (MC)
                tma_load_2d_mc(b_smem + static_cast<size_t>(sb) * p.b_item_bytes + b_slice_off, &maps.b, &S.b_full[sb], tap * p.kpad + kc * p.kbox,
                               n0 + im * p.b_slice_rows, mask_b);
              else
                tma_load_2d(b_smem + static_cast<size_t>(sb) * p.b_item_bytes, &maps.b, &S.b_full[sb], tap * p.kpad + kc * p.kbox, n0);
            }
          }
        }
        tr.ev(2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const int ksteps = p.kbox >> 4;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    Tracer tr; tr.init(lane == 0 ? p.trace : nullptr, 1);
    // The shape switch sits OUTSIDE the loops: inside the tap loop its jump-table load, the indexed constant loads of the tap
    // shifts and the 64-bit descriptor arithmetic formed one serial dependent chain of ~500 cycles per tap in this single warp
    // (knock-out measurement, profiles/r01_knockout_issue_loop.txt) — more than the MMAs of a narrow tap take to execute.
    switch ((p.mt << 4) | ksteps) {
      case 0x11: mma_role<1, 1, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x12: mma_role<1, 2, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x14: mma_role<1, 4, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x21: mma_role<2, 1, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x22: mma_role<2, 2, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x24: mma_role<2, 4, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x41: mma_role<4, 1, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      case 0x42: mma_role<4, 2, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
      default:   mma_role<4, 4, MC>(p, S, tb, smem_u32(a_smem), smem_u32(b_smem), my_group, n_units, groups, mask_a, mask_b, lane, tr); break;
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ aux producer (residual / add2 tiles)
    if (AUX != 0 && lane == 0 && !p.direct) {
      const int aux_hi = AUX == 1 ? p.n_res : p.n_valid;
      const uint32_t buf_off = AUX >= 2 ? static_cast<uint32_t>(p.boxes) * box_bytes : 0u;
      Tracer tr; tr.init(p.trace, 2);
      int ls = 0;
      for (int unit = my_group; unit < n_units; unit += groups, ++ls) {
        const int span = unit * sstride + im;
        const int p0 = (p.reverse ? n_spans_all - 1 - span : span) * span_px;
        for (int jh = 0; jh < p.mt * n_parts; ++jh) {
          const int j = jh / n_parts, cb = n0 + (jh % n_parts) * p.part_cols;   // sub-tile, first channel of this column part
          const uint32_t q = static_cast<uint32_t>(ls >> 1) * (p.mt * n_parts) + jh;   // slot counter of the warpgroup that owns this span
          const int slot = (ls & 1) * p.slots + q % p.slots;
          S.prog[2] = (ls << 8) | jh;
          int nb = 0;
          for (int b = 0; b < p.boxes; ++b)
            if (cb + b * p.box_ch < aux_hi) ++nb;
          tr.ev(1);
          wait_dbg(&S.slot_empty[slot], ((q / p.slots) & 1) ^ 1, p.dbg, 0x21, slot, q, S.prog);
          tr.ev(2);
          if (AUX == 3 && !KNOCK(4)) {       // add2 tile of a dense planar tensor: one contiguous run, no rows
            const uint32_t bytes = 128u * p.d_aux_pitch;
            mbar_expect_tx(&S.slot_full[slot], bytes);
            bulk_load_1d(slot_smem + static_cast<size_t>(slot) * p.slot_bytes + buf_off, p.d_aux + static_cast<size_t>(p0 + j * 128) * p.d_aux_pitch, bytes,
                         &S.slot_full[slot]);
          } else if (nb > 0 && !KNOCK(4)) {
            mbar_expect_tx(&S.slot_full[slot], static_cast<uint32_t>(nb) * box_bytes);
            uint8_t* dst = slot_smem + static_cast<size_t>(slot) * p.slot_bytes + buf_off;
            for (int b = 0; b < nb; ++b)
              tma_load_2d(dst + static_cast<size_t>(b) * box_bytes, &maps.aux, &S.slot_full[slot], cb + b * p.box_ch, p0 + j * 128);
          } else {
            mbar_arrive(&S.slot_full[slot]);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ store issuer
    if (lane == 0 && p.direct != 1) {
      Tracer tr; tr.init(p.trace, 3);
      const uint32_t slot_base = smem_u32(slot_smem);
      int prev_slot = -1;
      int ls = 0;
      for (int unit = my_group; unit < n_units; unit += groups, ++ls) {
        const int span = unit * sstride + im;
        const int p0 = (p.reverse ? n_spans_all - 1 - span : span) * span_px;
        for (int jh = 0; jh < p.mt * n_parts; ++jh) {
          const int j = jh / n_parts, cb = n0 + (jh % n_parts) * p.part_cols;
          const uint32_t q = static_cast<uint32_t>(ls >> 1) * (p.mt * n_parts) + jh;
          const int slot = (ls & 1) * p.slots + q % p.slots;
          S.prog[3] = (ls << 8) | jh;
          wait_dbg(&S.slot_ready[slot], (q / p.slots) & 1, p.dbg, 0x31, slot, q, S.prog);
          tr.ev(1);
          const uint32_t bufA = slot_base + static_cast<uint32_t>(slot) * p.slot_bytes;
          const uint32_t bufB = bufA + static_cast<uint32_t>(p.boxes) * box_bytes;
          const int px = p0 + j * 128;
          for (int b = 0; b < p.boxes; ++b) {
            const int cg = cb + b * p.box_ch;
            if (cg >= p.n_valid || KNOCK(2)) break;
            if (p.direct == 2) { tma_store_2d(&maps.o2, bufA + b * box_bytes, cg, px); continue; }   // hybrid: only out2 goes through the slot
            const int gb = cg / p.box_ch;           // global staging box (tiles and parts start on box boundaries)
            if (p.route_map[gb] != 0xff) tma_store_2d(&maps.o[p.route_map[gb]], bufA + b * box_bytes, p.route_c[gb], px);
            if (AUX == 2) tma_store_2d(&maps.o2, bufB + b * box_bytes, cg, px);
          }
          if (AUX == 3 && !KNOCK(2)) bulk_store_1d(p.d_out2 + static_cast<size_t>(px) * p.d_out2_pitch, bufB, 128u * p.d_out2_pitch);
          bulk_commit();
          if (p.slots == 1) {             // a single slot per warpgroup cannot stay held until the next commit
            bulk_wait_read<0>(); mbar_arrive(&S.slot_empty[slot]);
          } else {
            if (prev_slot >= 0) { bulk_wait_read<1>(); mbar_arrive(&S.slot_empty[prev_slot]); }
            prev_slot = slot;
          }
          tr.ev(2);
        }
      }
      if (prev_slot >= 0) { bulk_wait_read<0>(); mbar_arrive(&S.slot_empty[prev_slot]); }
      bulk_wait_all();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warpgroups
    const int wg = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    uint32_t row_off, row_xor; int bsh; uint32_t bmask;
    if (p.box_ch == 64) { row_off = static_cast<uint32_t>(m) * 128u; row_xor = static_cast<uint32_t>(m & 7) << 4; bsh = 6; bmask = 63u; }
    else { row_off = static_cast<uint32_t>(m) * 64u; row_xor = static_cast<uint32_t>((m >> 1) & 3) << 4; bsh = 5; bmask = 31u; }
    const uint32_t slot_base = smem_u32(slot_smem);
    const uint32_t sc_base = smem_u32(s_scale), sh_base = smem_u32(s_shift);
    const uint32_t bufB_off = static_cast<uint32_t>(p.boxes) * box_bytes;
    Tracer tr; tr.init((q4 == 0 && lane == 0) ? p.trace : nullptr, 4 + wg);
    int ls = 0;
    if (p.direct) {
      // ---- direct epilogue: this thread's pixel is one contiguous channel run in every tensor it touches
      const int ngrp1 = (min(p.n_tile, p.n_store - n0) + 7) >> 3;      // groups written to the primary output (pad groups: zeros)
      const int ngrp2 = (min(p.n_tile, p.n_store2 - n0) + 7) >> 3;     // groups of the aux tile / second output (padded planar tensors)
      const uint8_t* aux_base = p.d_aux + static_cast<size_t>(n0) * 2;
      uint8_t* out_base = p.d_out + static_cast<size_t>(n0) * 2;
      uint8_t* out2_base = p.d_out2 + static_cast<size_t>(n0) * 2;
      const bool aux32 = AUX != 0 && ((reinterpret_cast<uintptr_t>(aux_base) | p.d_aux_pitch) & 31u) == 0;     // 32-byte aligned pixel runs
      // 256-bit accesses only for runs that are whole sectors: a 48-byte run as 32 + 16 bytes measured slower than 3 x 16 (345 vs 322 us)
      const bool out32 = (ngrp1 & 1) == 0 && ((reinterpret_cast<uintptr_t>(out_base) | p.d_out_pitch) & 31u) == 0;
      const bool out2_32 = AUX >= 2 && (ngrp2 & 1) == 0 && ((reinterpret_cast<uintptr_t>(out2_base) | p.d_out2_pitch) & 31u) == 0;
      uint4 ax[8], nx[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) ax[g] = nx[g] = make_uint4(0, 0, 0, 0);
      auto load_aux = [&](const uint4* src, uint4 (&d)[8]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (2 * q + 1 < ngrp2 && aux32) ldg256(src + 2 * q, d[2 * q], d[2 * q + 1]);
          else {
            if (2 * q < ngrp2) d[2 * q] = __ldg(src + 2 * q);
            if (2 * q + 1 < ngrp2) d[2 * q + 1] = __ldg(src + 2 * q + 1);
          }
        }
      };
      for (int unit = my_group; unit < n_units; unit += groups, ++ls) {
        const int span = unit * sstride + im;
        if ((ls & 1) != wg) continue;
        const long long p0 = static_cast<long long>(p.reverse ? n_spans_all - 1 - span : span) * span_px;
        uint32_t vm[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long pp = p0 + j * 128 + m;
          vm[j] = (j < p.mt && pp < p.P && p.pix_valid[pp]) ? 0xffffffffu : 0u;
        }
        if (AUX != 0 && p0 + m < p.P_cap) {        // aux values of sub-tile 0: in flight while the MMAs of this span finish
          load_aux(reinterpret_cast<const uint4*>(aux_base + static_cast<size_t>(p0 + m) * p.d_aux_pitch), ax);
        }
        tr.ev(1);
        if (lane == 0) S.prog[warp] = 0x10000u | (static_cast<uint32_t>(ls) << 4);
        const int tb_i = ls & (p.tmem_bufs - 1);
        wait_dbg(&S.tmem_full[tb_i], (ls >> p.tmem_bufs_log2) & 1, p.dbg, 0x41, wg, ls, S.prog);
        tr.ev(2);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < p.mt; ++j) {
          const long long pp = p0 + j * 128 + m;
          const bool inb = pp < p.P_cap;
          if (AUX != 0 && j + 1 < p.mt && pp + 128 < p.P_cap) {   // next sub-tile's aux values load under this one's arithmetic
            load_aux(reinterpret_cast<const uint4*>(aux_base + static_cast<size_t>(pp + 128) * p.d_aux_pitch), nx);
          }
          const uint32_t vmask = j == 0 ? vm[0] : j == 1 ? vm[1] : j == 2 ? vm[2] : vm[3];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>((tb_i * p.mt + j) * p.n_tile);
          uint4* o1 = reinterpret_cast<uint4*>(out_base + static_cast<size_t>(pp) * p.d_out_pitch);
          uint4* o2 = reinterpret_cast<uint4*>(out2_base + static_cast<size_t>(pp) * p.d_out2_pitch);
          // hybrid (direct == 2): the second output is staged in a slot and leaves by TMA (scattered STG rows cost ~2.5x a TMA
          // row, profiles/r01_microbench_tma_rate.txt), the aux values and the primary output stay on the LSU
          const bool hybrid = AUX >= 2 && p.direct == 2;
          const uint32_t qs = static_cast<uint32_t>(ls >> 1) * p.mt + j;
          const int slot = wg * p.slots + qs % p.slots;
          const uint32_t slotA = slot_base + static_cast<uint32_t>(slot) * p.slot_bytes + row_off;
          if (hybrid) wait_dbg(&S.slot_empty[slot], ((qs / p.slots) & 1) ^ 1, p.dbg, 0x43, slot, qs, S.prog);
          uint32_t ra[16];
          tr.ev(3);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (2 * q < max(ngrp1, ngrp2)) {
              tmem_ld16(taddr + q * 16, ra);
              tmem_ld_wait();
              if (inb && !KNOCK(1)) {
                const bool two1 = 2 * q + 1 < ngrp1, one2 = 2 * q < ngrp2, two2 = 2 * q + 1 < ngrp2;   // groups of out1 / of aux + out2
                uint4 va, vb = make_uint4(0, 0, 0, 0), wa, wb = make_uint4(0, 0, 0, 0);
                epi8_direct<T, AUX, PRE, POST>(ra, sc_base + q * 64, sh_base + q * 64, ax[2 * q], va, wa, vmask);
                if (two1) epi8_direct<T, AUX, PRE, POST>(ra + 8, sc_base + q * 64 + 32, sh_base + q * 64 + 32, ax[2 * q + 1], vb, wb, vmask);
                if (KNOCK(64)) { if (va.x == 0x12345u) o1[0] = va; }     // timing experiment: arithmetic kept, stores dropped
                else if (two1) { if (out32) stg256(o1 + 2 * q, va, vb); else { o1[2 * q] = va; o1[2 * q + 1] = vb; } }
                else o1[2 * q] = va;
                if (AUX >= 2) {
                  if (hybrid) {
                    const uint32_t c0s = q * 16u, c1s = q * 16u + 8u;
                    if (one2) sts_u4(slotA + (c0s >> bsh) * box_bytes + ((((c0s & bmask) >> 3) << 4) ^ row_xor), wa);
                    if (two2) sts_u4(slotA + (c1s >> bsh) * box_bytes + ((((c1s & bmask) >> 3) << 4) ^ row_xor), wb);
                  } else if (two2) { if (out2_32) stg256(o2 + 2 * q, wa, wb); else { o2[2 * q] = wa; o2[2 * q + 1] = wb; } }
                  else if (one2) o2[2 * q] = wa;
                }
              }
            }
          }
          if (j == p.mt - 1) {              // all accumulators of this span have been read: hand TMEM back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tmem_empty[tb_i]);
          }
          if (hybrid) {
            fence_proxy_async();              // generic-proxy writes of this thread -> visible to the TMA store
            mbar_arrive(&S.slot_ready[slot]);
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) ax[g] = nx[g];
          tr.ev(4);
        }
      }
    } else
    for (int unit = my_group; unit < n_units; unit += groups, ++ls) {
        const int span = unit * sstride + im;
      if ((ls & 1) != wg) continue;
      const long long p0 = static_cast<long long>(p.reverse ? n_spans_all - 1 - span : span) * span_px;
      uint32_t vm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long pp = p0 + j * 128 + m;
        vm[j] = (j < p.mt && pp < p.P && p.pix_valid[pp]) ? 0xffffffffu : 0u;
      }
      tr.ev(1);
      if (lane == 0) S.prog[warp] = 0x10000u | (static_cast<uint32_t>(ls) << 4);
      const int tb_i = ls & (p.tmem_bufs - 1);
      wait_dbg(&S.tmem_full[tb_i], (ls >> p.tmem_bufs_log2) & 1, p.dbg, 0x41, wg, ls, S.prog);
      tr.ev(2);
      tc_fence_after();
#pragma unroll 1
      for (int jh = 0; jh < p.mt * n_parts; ++jh) {
        const int j = jh / n_parts, cpart = (jh % n_parts) * p.part_cols;   // sub-tile, first tile-local column of this part
        const uint32_t q = static_cast<uint32_t>(ls >> 1) * (p.mt * n_parts) + jh;   // this warpgroup's own slot counter: its ring is
        const int slot = wg * p.slots + q % p.slots;                                     // private, so a parity wait is never > 1 phase ahead
        const uint32_t u = q / p.slots;
        if (lane == 0) S.prog[warp] = 0x20000u | q;
        if (AUX != 0) wait_dbg(&S.slot_full[slot], u & 1, p.dbg, 0x42, slot, q, S.prog);
        else wait_dbg(&S.slot_empty[slot], (u & 1) ^ 1, p.dbg, 0x43, slot, q, S.prog);
        tr.ev(3);
        const uint32_t bufA = slot_base + static_cast<uint32_t>(slot) * p.slot_bytes + row_off;
        // AUX 3: the add2 / out2 tile is the linear image [128 pixels][pitch] of a dense planar tensor (48-byte rows: conflict-free)
        const uint32_t linB = slot_base + static_cast<uint32_t>(slot) * p.slot_bytes + bufB_off + static_cast<uint32_t>(m) * p.d_aux_pitch;
        const uint32_t vmask = j == 0 ? vm[0] : j == 1 ? vm[1] : j == 2 ? vm[2] : vm[3];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>((tb_i * p.mt + j) * p.n_tile);
        uint32_t ra[16], rb[16];
        const int c_end = cpart + p.part_cols;
        if (!KNOCK(8)) tmem_ld16(taddr + cpart, ra);
#pragma unroll 1
        for (int c0 = cpart; c0 < c_end; c0 += 32) {
          tmem_ld_wait();
          if (c0 + 16 < c_end && !KNOCK(8)) tmem_ld16(taddr + c0 + 16, rb);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int cl = c0 + g * 8, cg = n0 + cl;
            if (cg < p.n_valid && (cg & p.grp_mask) < p.grp_w && !KNOCK(1)) {
              const uint32_t cs = static_cast<uint32_t>(cl - cpart);   // column inside the slot
              const uint32_t ua = bufA + (cs >> bsh) * box_bytes + ((((cs & bmask) >> 3) << 4) ^ row_xor);
              epi8<T, AUX, PRE, POST>(ra + g * 8, sc_base + cl * 4, sh_base + cl * 4, ua, AUX == 3 ? linB + (cs >> 3) * 16u : ua + bufB_off, cg < p.n_res, vmask);
            }
          }
          if (c0 + 16 < c_end) {
            tmem_ld_wait();
            if (c0 + 32 < c_end && !KNOCK(8)) tmem_ld16(taddr + c0 + 32, ra);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int cl = c0 + 16 + g * 8, cg = n0 + cl;
              if (cg < p.n_valid && (cg & p.grp_mask) < p.grp_w && !KNOCK(1)) {
                const uint32_t cs = static_cast<uint32_t>(cl - cpart);   // column inside the slot
              const uint32_t ua = bufA + (cs >> bsh) * box_bytes + ((((cs & bmask) >> 3) << 4) ^ row_xor);
                epi8<T, AUX, PRE, POST>(rb + g * 8, sc_base + cl * 4, sh_base + cl * 4, ua, AUX == 3 ? linB + (cs >> 3) * 16u : ua + bufB_off, cg < p.n_res, vmask);
              }
            }
          }
        }
        if (jh == p.mt * n_parts - 1) {   // all accumulators of this span are in registers / smem: hand TMEM back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.tmem_empty[tb_i]);
        }
        if (lane == 0) S.prog[warp] = 0x30000u | q;
        fence_proxy_async();              // generic-proxy writes of this thread → visible to the TMA store
        mbar_arrive(&S.slot_ready[slot]);
        tr.ev(4);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();          // peers multicast into this CTA's shared memory and arrive on its barriers: nobody leaves early
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

size_t conv_flat_smem_bytes(const FlatConvParams& p) {
  const size_t b = p.b_resident ? static_cast<size_t>(p.taps) * p.nkc * p.b_item_bytes : static_cast<size_t>(p.b_stages) * p.b_item_bytes;
  return 1024 + kHeaderBytes + b + 2 * static_cast<size_t>(p.slots) * p.slot_bytes + static_cast<size_t>(p.a_stages) * p.a_stage_bytes;
}

static int g_flat_sms = 0;

template <typename T, int AUX, bool PRE, bool POST>
static cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel<T, AUX, PRE, POST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
#ifdef SVX_DEBUG_SWITCHES      // the multicast-cluster instantiation exists in the debug build only (measured slower)
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_flat_kernel<T, AUX, PRE, POST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
#endif
  return e;
}

template <typename T>
static cudaError_t init_type() {
  cudaError_t e;
  if ((e = set_attr<T, 0, false, false>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 0, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 0, true, false>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 1, false, false>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 1, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 2, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr<T, 3, false, true>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t conv_flat_init() {
  cudaError_t e = init_type<__half>();
  if (e != cudaSuccess) return e;
  e = init_type<__nv_bfloat16>();
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_flat_sms, cudaDevAttrMultiProcessorCount, dev);
}

// How many clusters of `cs` CTAs (one CTA per SM: every plan uses most of the shared memory) can be resident at once: clusters
// never straddle a GPC, so this is less than SMs / cs when the GPCs' SM counts are not multiples of cs.  Queried once per size.
template <typename T>
static int max_clusters(int cs, size_t smem) {
  static int cache[17] = {0};
  if (cs < 1 || cs > 16) return 0;
  if (cache[cs] > 0) return cache[cs];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(static_cast<unsigned>(cs * 64), 1, 1); cfg.blockDim = dim3(kFlatThreads, 1, 1); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cs); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
#ifdef SVX_DEBUG_SWITCHES
  const cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, conv_flat_kernel<T, 0, false, true, true>, &cfg);
#else
  const cudaError_t qe = cudaErrorNotSupported;
#endif
  if (qe != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = (g_flat_sms > 0 ? g_flat_sms : 148) / cs * 3 / 4;      // conservative guess
  }
  cache[cs] = n;
  return n;
}

int conv_flat_max_clusters(int cluster_size) { return max_clusters<__half>(cluster_size, 200 * 1024); }

template <typename T>
static cudaError_t launch_typed(const FlatConvParams& p, const FlatMaps& maps, dim3 grid, size_t smem, cudaStream_t st) {
  static const bool no_pdl = dbg_env("SVX_NO_PDL") != nullptr;   // debug switch
  const bool mc = p.cn * p.cm > 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = dim3(kFlatThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (mc) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(p.cn * p.cm); attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = cudaSuccess;
  static const bool launch_log = dbg_env("SVX_LAUNCH_LOG") != nullptr;   // debug switch
  if (launch_log)
    fprintf(stderr, "conv_flat launch: grid %u smem %zu cluster %dx%d n_tiles %d n_tile %d mt %d P %lld aux %d\n", grid.x, smem, p.cm, p.cn, p.n_tiles,
            p.n_tile, p.mt, p.P, p.aux_mode);
#ifdef SVX_DEBUG_SWITCHES
#define SVX_FLAT(AUX, PRE, POST)                                                                       \
  le = mc ? cudaLaunchKernelEx(&cfg, conv_flat_kernel<T, AUX, PRE, POST, true>, p, maps)              \
          : cudaLaunchKernelEx(&cfg, conv_flat_kernel<T, AUX, PRE, POST, false>, p, maps)
#else
  if (mc) return cudaErrorNotSupported;
#define SVX_FLAT(AUX, PRE, POST) le = cudaLaunchKernelEx(&cfg, conv_flat_kernel<T, AUX, PRE, POST, false>, p, maps)
#endif
  if (p.aux_mode == 0) {
    if (p.pre_relu && !p.post_relu) SVX_FLAT(0, true, false);
    else if (!p.pre_relu && p.post_relu) SVX_FLAT(0, false, true);
    else if (!p.pre_relu && !p.post_relu) SVX_FLAT(0, false, false);
    else return cudaErrorInvalidValue;
  } else if (p.aux_mode == 1) {
    if (p.pre_relu) return cudaErrorInvalidValue;
    if (p.post_relu) SVX_FLAT(1, false, true); else SVX_FLAT(1, false, false);
  } else {
    if (p.pre_relu || !p.post_relu) return cudaErrorInvalidValue;
    if (p.lin) SVX_FLAT(3, false, true); else SVX_FLAT(2, false, true);
  }
#undef SVX_FLAT
  if (le != cudaSuccess) {
    fprintf(stderr, "conv_flat launch failed: grid %u block %d smem %zu cluster %dx%d n_tiles %d n_tile %d mt %d\n", grid.x, kFlatThreads, smem, p.cm, p.cn,
            p.n_tiles, p.n_tile, p.mt);
    return le;
  }
  return cudaGetLastError();
}

cudaError_t launch_conv_flat(const FlatConvParams& p, const FlatMaps& maps, int is_bf16, cudaStream_t stream) {
  if (p.P <= 0) return cudaSuccess;
  const long long n_spans = (p.P + p.mt * 128 - 1) / (p.mt * 128);
  const int sms = g_flat_sms > 0 ? g_flat_sms : 148;
  const size_t smem = conv_flat_smem_bytes(p);
  const int cs = p.cn * p.cm;                                            // CTAs per scheduling unit (cluster)
  const int n_cols = p.n_tiles / p.cn;                                   // clusters per span group
  const long long n_units = (n_spans + p.cm - 1) / p.cm;
  const int unit_cap = cs > 1 ? (is_bf16 ? max_clusters<__nv_bfloat16>(cs, smem) : max_clusters<__half>(cs, smem)) : sms;
  long long groups = unit_cap / n_cols;
  if (groups < 1) groups = 1;
  if (groups > n_units) groups = n_units;
  dim3 grid(static_cast<unsigned>(groups * n_cols * cs), 1, 1);
  return is_bf16 ? launch_typed<__nv_bfloat16>(p, maps, grid, smem, stream) : launch_typed<__half>(p, maps, grid, smem, stream);
}

}  // namespace svx
