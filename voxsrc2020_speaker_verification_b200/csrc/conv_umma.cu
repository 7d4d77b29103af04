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
// Persistent, warp-specialised CTA (one per SM, 320 threads), tiles dealt round-robin:
//   warp 0      TMA producer: runs ahead across tile boundaries through a shared-memory ring of (A,B) stages and
//               prefetches the tile's residual / add2 operand ("aux", 64-channel SWIZZLE_128B boxes)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; two accumulator buffers in TMEM so the
//               MMAs of tile i+1 overlap the epilogue of tile i
//   warps 2-5   epilogue warpgroup 0 (even tiles), warps 6-9 epilogue warpgroup 1 (odd tiles):
//               tcgen05.ld → scale/shift/ReLU/residual/concat routing → 16-byte stores
#include <cstdlib>

#include "conv.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

constexpr int kUmmaThreads = 320;
constexpr int kMaxStages = 8;
constexpr int kAuxBoxBytes = 128 * 128;   // 128 rows x 64 channels x 2 B

template <typename T>
__device__ __forceinline__ void add8(float (&v)[8], const uint4& r) {
  const float2 a = TypeOps<T>::unpack2(r.x), b = TypeOps<T>::unpack2(r.y), c = TypeOps<T>::unpack2(r.z), d = TypeOps<T>::unpack2(r.w);
  v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
}

template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  o.x = TypeOps<T>::pack2(v[0], v[1]); o.y = TypeOps<T>::pack2(v[2], v[3]);
  o.z = TypeOps<T>::pack2(v[4], v[5]); o.w = TypeOps<T>::pack2(v[6], v[7]);
  return o;
}

// c: first of 8 consecutive output channels (multiple of 8, < n_valid); c_local = c - n0.
// Shared-memory accesses of the epilogue use explicit shared-space addresses: generic-pointer loads cost the
// epilogue ~3x (profiles/r01_trace_*).
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

// Shared-space address of the 16-byte unit that holds channels [c_local, c_local+8) of tile row m inside a buffer of
// 64-channel SWIZZLE_128B boxes (unit u of row m lives at u ^ (m & 7)) — the layout TMA loads produce and TMA
// stores consume.  row_off = m * 128, row_xor = (m & 7) << 4 are per-thread constants.
__device__ __forceinline__ uint32_t box_unit(uint32_t buf, uint32_t row_off, uint32_t row_xor, int c_local) {
  return buf + static_cast<uint32_t>(c_local >> 6) * kAuxBoxBytes + row_off + ((static_cast<uint32_t>(c_local & 56) << 1) ^ row_xor);
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Per-thread, per-tile constants of the epilogue, all in registers.
struct EpiCtx {
  uint32_t scale_addr, shift_addr;   // shared-space addresses of the staged per-channel vectors
  uint32_t aux, stage;               // shared-space addresses of this warpgroup's aux / staging buffers
  uint32_t row_off, row_xor;         // m * 128, (m & 7) << 4
  int aux_mode, n_split;
  bool staged, pre_relu, post_relu, f32;
};

// One group of 8 consecutive output channels of one pixel: BN scale/shift, ReLUs, residual, routing.
//   c: first channel (multiple of 8, < n_valid); c_local = c - n0.
// Staged mode: 16-bit results go to swizzled shared-memory boxes that one thread later hands to TMA stores (the
// residual tile is overwritten in place); secondary-destination channels and the fp32 form are stored directly.
template <typename T>
__device__ __forceinline__ void epilogue8(const Epilogue& e, const EpiCtx& x, float (&v)[8], int c, int c_local, size_t pix, int row,
                                          bool in_range, bool valid) {
  const float4 s0 = lds_f4(x.scale_addr + c * 4), s1 = lds_f4(x.scale_addr + c * 4 + 16);
  const float4 b0 = lds_f4(x.shift_addr + c * 4), b1 = lds_f4(x.shift_addr + c * 4 + 16);
  const bool prim = c < x.n_split;
  const uint32_t au = box_unit(x.aux, x.row_off, x.row_xor, c_local);
  uint4 ax = make_uint4(0, 0, 0, 0);
  if (x.aux_mode != 0 && prim) ax = lds_u4(au);
  if (x.pre_relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  v[0] = fmaf(v[0], s0.x, b0.x); v[1] = fmaf(v[1], s0.y, b0.y); v[2] = fmaf(v[2], s0.z, b0.z); v[3] = fmaf(v[3], s0.w, b0.w);
  v[4] = fmaf(v[4], s1.x, b1.x); v[5] = fmaf(v[5], s1.y, b1.y); v[6] = fmaf(v[6], s1.z, b1.z); v[7] = fmaf(v[7], s1.w, b1.w);
  if (x.f32) {
    if (!in_range) return;
    float* o = e.out_f32 + static_cast<size_t>(row) * e.ldf + c;
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  if (x.aux_mode == 1 && prim) add8<T>(v, ax);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (x.post_relu) v[j] = fmaxf(v[j], 0.f);
    if (!valid) v[j] = 0.f;
  }
  const uint4 o = pack8<T>(v);
  if (prim) {
    if (!x.staged) {                                            // direct 16-byte stores (debug / fallback form)
      if (!in_range) return;
      *reinterpret_cast<uint4*>(static_cast<T*>(e.out) + pix * e.out_C + e.out_coff + c) = o;
      if (x.aux_mode == 2) {
        if (valid) add8<T>(v, ax);
        *reinterpret_cast<uint4*>(static_cast<T*>(e.out2) + pix * e.out2_C + e.out2_coff + c) = pack8<T>(v);
      }
    } else if (x.aux_mode == 1) {
      sts_u4(au, o);                                            // residual tile overwritten in place, stored from there
    } else {
      sts_u4(box_unit(x.stage, x.row_off, x.row_xor, c_local), o);
      if (x.aux_mode == 2) {                                    // out2 = v + add2, in place on the add2 tile
        if (valid) add8<T>(v, ax);
        sts_u4(au, pack8<T>(v));
      }
    }
  } else {
    if (!in_range) return;
    *reinterpret_cast<uint4*>(static_cast<T*>(e.outb) + pix * e.outb_C + e.outb_coff + (c - e.n_split)) = o;
  }
}

// debug timeline: event (role, code) stamped with the global ns timer, CTA 0 only
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

template <typename T>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ UmmaConvParams p, const __grid_constant__ AMaps amaps,
                 const __grid_constant__ CUtensorMap bmap, const __grid_constant__ CUtensorMap auxmap,
                 const __grid_constant__ OMaps omaps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);          // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;                      // [kMaxStages]
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;                 // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                     // [2]
  uint64_t* aux_full_bar = tmem_empty_bar + 2;                      // [2]
  uint64_t* aux_empty_bar = aux_full_bar + 2;                       // [2]
  uint64_t* bres_bar = aux_empty_bar + 2;                           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);          // [n_pad]
  float* s_shift = s_scale + p.n_tiles * p.n_tile;                  // [n_pad]
  uint8_t* bres = smem + 1024 + p.ss_bytes;                         // resident weights: [taps*nkc][b_stage_bytes]
  uint8_t* aux_smem = bres + p.bres_bytes;                          // [2][aux_boxes][16 KB]
  uint8_t* stage_smem = aux_smem + 2 * p.aux_bytes;                 // [2][stage_boxes][16 KB]
  uint8_t* tiles = stage_smem + 2 * p.stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool b_res = p.bres_bytes != 0;
  const uint32_t ring_bytes = p.a_stage_bytes + (b_res ? 0u : p.b_stage_bytes);
  const int aux_mode = p.aux_mode;
  const bool staged = p.store_mode == 1;
  const int aux_arrivals = staged ? 1 : 4;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&amaps.m[i]);
    prefetch_tmap(&bmap);
    if (aux_mode) prefetch_tmap(&auxmap);
    if (staged) { prefetch_tmap(&omaps.m[0]); prefetch_tmap(&omaps.m[1]); }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);     // one arrive per epilogue warp of the warpgroup
      mbar_init(&aux_full_bar[b], 1);
      mbar_init(&aux_empty_bar[b], aux_arrivals);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.n_tile; i += kUmmaThreads) {
    s_scale[i] = (p.epi.scale && i < p.epi.n_valid) ? p.epi.scale[i] : 1.f;
    s_shift[i] = (p.epi.shift && i < p.epi.n_valid) ? p.epi.shift[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile schedule: CTA b owns n-tile (b % n_tiles) for its whole life (its weights can stay resident) and walks the
  // m-tiles (b / n_tiles) + j * (gridDim.x / n_tiles).
  const int row_tiles = (p.out_rows + p.h_box - 1) / p.h_box;
  const int m_tiles = row_tiles * p.w_tiles;
  const int groups = gridDim.x / p.n_tiles;
  const int my_group = blockIdx.x / p.n_tiles;
  const int n_blk = blockIdx.x % p.n_tiles;
  const int n0 = n_blk * p.n_tile;
  const int total_it = p.taps * p.nkc;
  const bool active = my_group < groups;     // CTAs past the last full group (never launched by the host) stay idle

  if (warp == 0) {
    if (lane == 0 && active) {
      const uint32_t b_box_bytes = static_cast<uint32_t>(p.n_tile) * p.kbox * 2u;
#ifdef SVX_ASHIFT
      const uint32_t tx_bytes = 256u * p.kbox * 2u + (b_res ? 0u : b_box_bytes);
#else
      const uint32_t tx_bytes = 128u * p.kbox * 2u + (b_res ? 0u : b_box_bytes);
#endif
      if (b_res) {
        mbar_expect_tx(bres_bar, b_box_bytes * total_it);
        for (int i = 0; i < total_it; ++i) tma_load_2d(bres + static_cast<size_t>(i) * p.b_stage_bytes, &bmap, bres_bar, i * p.kbox, n0);
      }
      const int a_c0 = n_blk * p.a_c_step;
      uint32_t it = 0;
      int local = 0;
      Tracer tr; tr.init(p.trace, 0);
      for (int mt = my_group; mt < m_tiles; mt += groups, ++local) {
        const int row0 = (mt / p.w_tiles) * p.h_box;
        const int w0 = (mt % p.w_tiles) * p.w_box;
        tr.ev(1);                                             // tile start
        if (aux_mode) {
          const int b = local & 1;
          const int boxes = min(p.aux_boxes, (p.aux_width - n0 + 63) >> 6);
          mbar_wait(&aux_empty_bar[b], ((local >> 1) & 1) ^ 1);
          tr.ev(2);                                           // aux buffer free
          if (boxes > 0) {
            mbar_expect_tx(&aux_full_bar[b], static_cast<uint32_t>(boxes) * kAuxBoxBytes);
            for (int j = 0; j < boxes; ++j)
              tma_load_3d(aux_smem + b * p.aux_bytes + j * kAuxBoxBytes, &auxmap, &aux_full_bar[b], n0 + j * 64, w0, row0);
          } else {
            mbar_arrive(&aux_full_bar[b]);
          }
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const CUtensorMap* am = &amaps.m[p.tap_map[tap]];
          const int wc = w0 + p.tap_dw[tap];
          const int rc = row0 + p.tap_dh[tap];
          for (int kc = 0; kc < p.nkc; ++kc, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], tx_bytes);
            uint8_t* a_dst = tiles + static_cast<size_t>(s) * ring_bytes;
#ifdef SVX_ASHIFT
            // experiment (W = 1 only): the stage holds rows [rc - SHIFT, rc - SHIFT + 256); the MMA reads a view that starts
            // SHIFT rows into it
            tma_load_3d(a_dst, am, &full_bar[s], a_c0 + kc * p.kbox, wc, rc - SVX_ASHIFT);
            tma_load_3d(a_dst + 128 * p.kbox * 2, am, &full_bar[s], a_c0 + kc * p.kbox, wc, rc - SVX_ASHIFT + 128);
#else
            tma_load_3d(a_dst, am, &full_bar[s], a_c0 + kc * p.kbox, wc, rc);
#endif
            if (!b_res) tma_load_2d(a_dst + p.a_stage_bytes, &bmap, &full_bar[s], (tap * p.nkc + kc) * p.kbox, n0);
          }
        }
        tr.ev(3);                                             // all loads of the tile issued
      }
    }
  } else if (warp == 1) {
    if (active) {
      const int ksteps = p.kbox >> 4;   // UMMA K = 16 elements = 32 bytes
      if (b_res) mbar_wait(bres_bar, 0);
      uint32_t it = 0;
      int local = 0;
      Tracer tr; tr.init(lane == 0 ? p.trace : nullptr, 1);
      for (int mt = my_group; mt < m_tiles; mt += groups, ++local) {
        const int b = local & 1;
        tr.ev(1);
        mbar_wait(&tmem_empty_bar[b], ((local >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tr.ev(2);                                                   // accumulator free
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * p.n_tile);
        for (int i = 0; i < total_it; ++i, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          {
            // descriptors are computed by the whole (converged) warp and only the tcgen05 instructions sit under the elected
            // lane, fully unrolled: inside an `if (lane == 0)` region ptxas wrapped every UTCHMMA in an ELECT / R2UR.BROADCAST
            // waterfall that cost ~200 ns per MMA (tools/microbench/mma_rate.cu)
            const uint32_t a_addr0 = smem_u32(tiles + static_cast<size_t>(s) * ring_bytes);
            const uint32_t b_addr = b_res ? smem_u32(bres + static_cast<size_t>(i) * p.b_stage_bytes) : a_addr0 + p.a_stage_bytes;
#ifdef SVX_ASHIFT
            const uint32_t a_addr = a_addr0 + SVX_ASHIFT * p.kbox * 2;
#else
            const uint32_t a_addr = a_addr0;
#endif
            const uint64_t dbase = make_kmajor_desc(0, p.sbo, p.layout_type);
            const uint64_t adesc0 = dbase + (a_addr >> 4), bdesc0 = dbase + (b_addr >> 4);
            const uint32_t first = i != 0 ? 1u : 0u;
            const bool last = i == total_it - 1;
            if (elect_one()) {
              if (ksteps == 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc0 + 2 * k, bdesc0 + 2 * k, p.idesc, k == 0 ? first : 1u);
              } else if (ksteps == 2) {
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f16(d_tmem, adesc0 + 2 * k, bdesc0 + 2 * k, p.idesc, k == 0 ? first : 1u);
              } else {
                umma_f16(d_tmem, adesc0, bdesc0, p.idesc, first);
              }
              umma_commit(&empty_bar[s]);                       // frees the smem slot when these MMAs retire
              if (last) umma_commit(&tmem_full_bar[b]);
            }
          }
          __syncwarp();
          if (i == 0) tr.ev(3);                                     // first operands landed
        }
        tr.ev(4);                                                   // all MMAs issued
      }
    }
  } else if (active) {
    const int wg = (warp - 2) >> 2;          // epilogue warpgroup: even / odd local tiles
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;             // accumulator row = pixel of the tile
    const int h = m / p.w_box;
    const int w = m - h * p.w_box;
    const bool leader = (warp - 2) == wg * 4 && lane == 0;
    EpiCtx x;
    x.scale_addr = smem_u32(s_scale); x.shift_addr = smem_u32(s_shift);
    x.aux = smem_u32(aux_smem + wg * p.aux_bytes); x.stage = smem_u32(stage_smem + wg * p.stage_bytes);
    x.row_off = static_cast<uint32_t>(m) * 128u; x.row_xor = static_cast<uint32_t>(m & 7) << 4;
    x.aux_mode = aux_mode; x.n_split = p.epi.n_split; x.staged = staged;
    x.pre_relu = p.epi.pre_relu != 0; x.post_relu = p.epi.post_relu != 0; x.f32 = p.epi.out_f32 != nullptr;
    const int n_valid = p.epi.n_valid;
    int local = wg;
    // the segment id of this thread's row is fetched one tile ahead: the load competes with saturated TMA traffic
    // and would otherwise sit on the critical path of every tile
    auto seg_fetch = [&](int mt_) -> int {
      if (mt_ >= m_tiles) return -1;
      const int row_ = (mt_ / p.w_tiles) * p.h_box + h;
      const int col_ = (mt_ % p.w_tiles) * p.w_box + w;
      if (row_ >= p.out_rows || col_ >= p.out_W) return -1;
      return p.epi.seg_of_row ? p.epi.seg_of_row[row_] : 0;
    };
    int seg_next = seg_fetch(my_group + wg * groups);
    Tracer tr; tr.init(leader ? p.trace : nullptr, 2 + wg);
    for (int mt = my_group + wg * groups; mt < m_tiles; mt += 2 * groups, local += 2) {
      const int row0 = (mt / p.w_tiles) * p.h_box;
      const int w0 = (mt % p.w_tiles) * p.w_box;
      const int row = row0 + h;
      const int col = w0 + w;
      const bool in_range = (row < p.out_rows) && (col < p.out_W);
      const int seg_cur = seg_next;
      seg_next = seg_fetch(mt + 2 * groups);
      const size_t pix = static_cast<size_t>(row) * p.out_Wp + col;
      const uint32_t ph = (local >> 1) & 1;
      if (staged && p.stage_bytes) {          // the previous tile's TMA store must have finished reading the staging boxes
        if (leader) bulk_wait_read0();
        named_bar(1 + wg, 128);
      }
      tr.ev(1);
      if (aux_mode) mbar_wait(&aux_full_bar[wg], ph);
      tr.ev(2);                                                     // aux tile landed
      mbar_wait(&tmem_full_bar[wg], ph);
      tr.ev(3);                                                     // accumulator ready
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(wg * p.n_tile);
      const bool valid = in_range && seg_cur >= 0;
      // two register buffers: the TMEM load of chunk i+1 is in flight while chunk i is converted and stored
      uint32_t ra[16], rb[16];
      tmem_ld16(taddr, ra);
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        tmem_ld_wait();
        if (c0 + 16 < p.n_tile) tmem_ld16(taddr + c0 + 16, rb);
        if (in_range || staged) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int c = n0 + c0 + g * 8;
            if (c < n_valid) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(ra[g * 8 + j]);
              epilogue8<T>(p.epi, x, v, c, c0 + g * 8, pix, row, in_range, valid);
            }
          }
        }
        if (c0 + 16 < p.n_tile) {
          tmem_ld_wait();
          if (c0 + 32 < p.n_tile) tmem_ld16(taddr + c0 + 32, ra);
          if (in_range || staged) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int c = n0 + c0 + 16 + g * 8;
              if (c < n_valid) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rb[g * 8 + j]);
                epilogue8<T>(p.epi, x, v, c, c0 + 16 + g * 8, pix, row, in_range, valid);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[wg]);
      tr.ev(4);                                                     // tile converted
      if (staged) {
        fence_proxy_async();                  // generic-proxy writes to the boxes → visible to the TMA (async proxy)
        named_bar(1 + wg, 128);
        if (leader) {
          // primary destination: 64-channel boxes, clipped by the tensor map at the slice width (n_split)
          const int prim_hi = min(p.epi.n_split, p.epi.n_valid);
          const uint32_t src = aux_mode == 1 ? x.aux : x.stage;
          for (int j = 0; n0 + j * 64 < min(prim_hi, n0 + p.n_tile); ++j)
            tma_store_3d(&omaps.m[0], src + j * kAuxBoxBytes, n0 + j * 64, w0, row0);
          if (aux_mode == 2)
            for (int j = 0; n0 + j * 64 < min(prim_hi, n0 + p.n_tile); ++j)
              tma_store_3d(&omaps.m[1], x.aux + j * kAuxBoxBytes, n0 + j * 64, w0, row0);
          bulk_commit();
          tr.ev(5);                                                 // stores issued
          if (aux_mode) {                     // the producer may refill this aux buffer once the stores have read it
            bulk_wait_read0();
            mbar_arrive(&aux_empty_bar[wg]);
            tr.ev(6);                                               // stores have read the boxes
          }
        }
      } else if (aux_mode && lane == 0) {
        mbar_arrive(&aux_empty_bar[wg]);
      }
    }
    if (staged && leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

static size_t ring_stage_bytes(const UmmaConvParams& p) { return p.a_stage_bytes + (p.bres_bytes ? 0u : p.b_stage_bytes); }

size_t conv_umma_smem_bytes(const UmmaConvParams& p) {
  return 2048 + p.ss_bytes + p.bres_bytes + 2 * static_cast<size_t>(p.aux_bytes) + 2 * static_cast<size_t>(p.stage_bytes) +
         static_cast<size_t>(p.stages) * ring_stage_bytes(p);
}

// Fills the shared-memory plan (resident weights, aux / staging boxes, ring depth) and tmem_cols from the tile shape;
// returns false if the tile cannot be scheduled.
bool conv_umma_finish_params(UmmaConvParams& p) {
#ifdef SVX_ASHIFT
  p.a_stage_bytes *= 2;      // experiment: 256-row stage so a shifted 128-row view stays inside loaded data
#endif
  const int total_it = p.taps * p.nkc;
  const int boxes = (p.n_tile + 63) / 64;
  p.aux_bytes = p.aux_mode ? static_cast<uint32_t>(p.aux_boxes) * kAuxBoxBytes : 0u;
  p.stage_bytes = (p.store_mode == 1 && p.aux_mode != 1) ? static_cast<uint32_t>(boxes) * kAuxBoxBytes : 0u;
  p.ss_bytes = static_cast<uint32_t>((2 * p.n_tiles * p.n_tile * 4 + 1023) / 1024 * 1024);
  const long long fixed = 2048 + p.ss_bytes + 2LL * p.aux_bytes + 2LL * p.stage_bytes;
  // weights stay resident in shared memory when they are small (every layer of the two high-resolution stages)
  const long long b_total = static_cast<long long>(total_it) * p.b_stage_bytes;
  p.bres_bytes = 0;
  static const bool no_bres = dbg_env("SVX_NO_BRES") != nullptr;   // debug switch
  if (!no_bres && b_total <= 64 * 1024 && 225 * 1024 - fixed - b_total >= 4LL * p.a_stage_bytes)
    p.bres_bytes = static_cast<uint32_t>(b_total);
  const long long budget = 225 * 1024 - fixed - p.bres_bytes;
  long long stages = budget / static_cast<long long>(ring_stage_bytes(p));
  if (stages < 2) return false;
  if (stages > kMaxStages) stages = kMaxStages;
  static const int env_stages = dbg_env("SVX_MAX_STAGES") ? atoi(dbg_env("SVX_MAX_STAGES")) : 0;   // debug switch
  if (env_stages >= 2 && stages > env_stages) stages = env_stages;
  p.stages = static_cast<int>(stages);
  uint32_t tc = 32;
  while (tc < 2u * p.n_tile) tc *= 2;
  if (tc > 512) return false;
  p.tmem_cols = tc;
  return true;
}

static int g_num_sms = 0;

cudaError_t conv_umma_init() {
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t launch_conv_umma(const UmmaConvParams& p, const AMaps& amaps, const CUtensorMap& bmap, const CUtensorMap& auxmap,
                             const OMaps& omaps, int is_bf16, cudaStream_t stream) {
  const int row_tiles = (p.out_rows + p.h_box - 1) / p.h_box;
  const long long m_tiles = static_cast<long long>(row_tiles) * p.w_tiles;
  if (m_tiles <= 0) return cudaSuccess;
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  long long groups = sms / p.n_tiles;                  // CTAs come in groups of n_tiles (one per n-tile of an m-tile)
  if (groups < 1) groups = 1;
  if (groups > m_tiles) groups = m_tiles;
  dim3 grid(static_cast<unsigned>(groups * p.n_tiles), 1, 1);
  const size_t smem = conv_umma_smem_bytes(p);
  if (is_bf16)
    conv_umma_kernel<__nv_bfloat16><<<grid, kUmmaThreads, smem, stream>>>(p, amaps, bmap, auxmap, omaps);
  else
    conv_umma_kernel<__half><<<grid, kUmmaThreads, smem, stream>>>(p, amaps, bmap, auxmap, omaps);
  return cudaGetLastError();
}

}  // namespace svx
