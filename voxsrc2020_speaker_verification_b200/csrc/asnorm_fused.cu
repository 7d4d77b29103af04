// Fused cohort statistics for AS-norm (reference tensorflow/snorm.py:83-110, get_cohort_mean_std): for every test row the mean and
// population std of its top-k dot products with the cohort, WITHOUT ever writing the score matrix.
//
//   S[r, j] = <x_r, c_j>  as one tcgen05 GEMM over split-bf16 operands, x = xh + xl, c = ch + cl (bf16 each):
//             xh.ch + xl.ch + xh.cl  (fp32 accumulation in TMEM, ~2^-16 relative error per product — the same three terms, in the
//             same order, as the unfused path, which is pinned to the reference's golden vectors)
//
// One persistent CTA per SM owns 128 test rows at a time.  Their split operand ([hi | lo], 2 x Dp bf16 per row) is written by the
// epilogue threads into TENSOR MEMORY (tcgen05.mma with A from TMEM), the cohort streams through a TMA ring in tiles of 128 cohort
// rows, TWICE:
//   pass 0  (statistics only: on the xh.ch term alone — half of the cohort bytes, a third of the MMAs; its scores are off by ~2^-8 of
//           a score's standard deviation, a small fraction of a bin)  epilogue thread r (one per test row and epilogue group,
//           reading its accumulator row from TMEM) bins every score into a PRIVATE 16-bit histogram in shared memory: plain load /
//           add / store in batches of four with duplicate folding — no atomics (ATOMS costs ~2 cycles per lane, a predicated one
//           becomes a branch region).  The bin range is centred on where the k-th largest is expected: mean + z * std of the row's
//           first 128 scores, z from the normal quantile of k/c.  After the pass the thread walks the bins of both groups from the
//           top and finds the bin b* that holds the k-th largest.
//   pass 1  the exact three-term GEMM.  Binning is monotone in the score, so "bin >= b" is "score >= t_b" for a threshold found by
//           bisection once per row: scores at or above t_{b*+2} are accumulated (sum and sum of squares in double), the scores of the
//           band [t_{b*-1}, t_{b*+2}) are collected in a short per-row list, and the thread finally picks the largest
//           k - |above| of them.  Exact whenever 0 <= k - |above| <= |band|, whatever the error of pass 0 — pass 1 verifies it.
//           (When the selected scores themselves are wanted — cohort-sharded layout — pass 0 runs on all three terms and the band
//           is the bin b* alone.)  Exact for ties: only the VALUES of the top-k multiset matter.
// A row whose threshold bin falls outside the histogram range, whose band overflows its list or fails the check above is reported
// through flag_rows and finished by the unfused kernels (api.cu) — correctness never depends on the score distribution, only the
// speed does.
//
// Roles (384 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer (two 128-column accumulator buffers in TMEM, so the MMAs of
// tile t+1 overlap the epilogue of tile t), warps 4-7 / 8-11 two epilogue warpgroups on alternating tiles (warp & 3 = TMEM lane
// quarter).  All histogram / list words of a row live at word index == row (mod 128): every epilogue thread touches only its own
// words, bank = lane, no synchronisation inside a pass.
//
// Bounds (profiles/r02_ncu_final.md): per tile of 128 cohort rows the CTA ingests 2 * Dp * 128 * 2 B (128 KB at D = 256;
// half of it in pass 0) at ~50 GB/s per SM when every SM pulls the same cohort from L2: 1.35 ms for the two passes of the
// BASELINE job, plus the part of the epilogues that two accumulator buffers cannot hide (0.65 ms).  A third accumulator buffer for
// pass 0 (over the lo half of the operand, written between the passes) was built and measured: no gain — the two epilogue groups'
// own throughput is the bound of pass 0, not the number of buffers (profiles/r02_experiments.md).
#include <cstdio>
#include <cstring>

#include "kernels.cuh"
#include "umma.cuh"

namespace svx {

using namespace ptx;

namespace {

constexpr int kThreads = 384;
constexpr int kMaxStages = 12;
constexpr uint32_t kBoxBytes = 128u * 128u;   // 128 rows x 64 bf16
constexpr float kPad = -1e30f;

struct Bars {
  uint64_t a_full, a_empty;
  uint64_t b_full[kMaxStages], b_empty[kMaxStages];
  uint64_t t_full[2], t_empty[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

constexpr int kListSlack = 4;                  // spare words behind a candidate list: its write pointer is clamped once per four scores

// bytes of the area that holds the two private histograms (pass 0) and then the two candidate lists (pass 1), a multiple of 512
__host__ __device__ inline uint32_t asnorm_hist_bytes(int nb) { return static_cast<uint32_t>((nb + 2) >> 1) * 512u; }   // stored bins 0 .. nb
__device__ __forceinline__ uint32_t hist_addr(uint32_t hist_row, int b) {
  return hist_row + static_cast<uint32_t>(b) * 256u - static_cast<uint32_t>(b & 1) * 254u;
}
__host__ __device__ inline uint32_t asnorm_main_bytes(int nb, int cap) {
  const uint32_t h = 2u * asnorm_hist_bytes(nb);
  const uint32_t l = 2u * static_cast<uint32_t>(cap + kListSlack) * 512u;
  return h > l ? h : l;
}

// Stored bin of a score, three instructions: u = sat((v - lo) / range) in [0, 1] (FFMA.SAT, NaN -> 0), then 2^23 + floor(u * K) by a
// round-toward-zero FMA onto 2^23 (K just below nb + 1, so u = 1 lands in bin nb), whose mantissa is the bin.  Bin 0 = below the
// range (and its lowest sliver), bin nb = its top sliver and everything above.  Monotone in v.
__device__ __forceinline__ int asnorm_bin(float v, float scale, float off, float kf) {
  return __float_as_int(__fmaf_rz(__saturatef(fmaf(v, scale, off)), kf, 8388608.f)) - 0x4B000000;
}

// smallest float whose stored bin is >= target (+inf when there is none): bisection over the order-preserving integer key of a float
__device__ __forceinline__ float asnorm_bin_threshold(int target, float scale, float off1, float top) {
  auto unkey = [](uint32_t kk) { return __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk); };
  uint32_t lo = ~0xff7fffffu;                  // key(-FLT_MAX)
  uint32_t hi = 0x7f800000u | 0x80000000u;     // key(+inf)
  if (asnorm_bin(unkey(hi), scale, off1, top) < target) return unkey(hi);
  if (asnorm_bin(unkey(lo), scale, off1, top) >= target) return unkey(lo);
  while (hi - lo > 1u) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (asnorm_bin(unkey(mid), scale, off1, top) >= target) hi = mid; else lo = mid;
  }
  return unkey(hi);
}

}  // namespace

template <bool VALS>
__global__ void __launch_bounds__(kThreads, 1)
asnorm_fused_kernel(const __grid_constant__ AsnormFusedParams p, const __grid_constant__ CUtensorMap map_b) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  Bars& B = *reinterpret_cast<Bars*>(smem);
  const int kb_n = p.kboxes;                                   // 64-element K boxes per operand half (hi or lo)
  uint8_t* b_smem = smem + 1024;
  uint8_t* aux = b_smem + static_cast<size_t>(p.stages) * kBoxBytes;   // per-row histogram (pass 0) / candidate list (pass 1), [word][128 rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = 2 * kb_n;                                   // ring items per cohort tile: hi[0], lo[0], hi[1], lo[1], ...
  // One-term pass 0 (statistics only): the histogram pass runs on xh.ch alone — half of the cohort bytes, a third of the MMAs.  Its
  // scores are off by ~2^-8 of a score's standard deviation, a small fraction of a bin, so the k-th largest EXACT score lies in
  // the bins b*-1 .. b*+1 around the approximate threshold bin; pass 1 (exact scores) sums what lies above that band, lists the
  // band, and the row is exact whenever 0 <= k - |above| <= |band| — which pass 1 itself verifies (otherwise the row is handed back).
  constexpr bool kOneTerm = !VALS;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_b);
    mbar_init(&B.a_full, 4); mbar_init(&B.a_empty, 1);          // a_full: the 4 warps of epilogue group 0 have written the operand
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&B.b_full[s], 1); mbar_init(&B.b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&B.t_full[b], 1); mbar_init(&B.t_empty[b], 4); }      // the 4 warps of the group that owns the buffer
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&B.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = B.tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;                                             // ring position and its phase, kept incrementally (no division per item)
      for (int rb = blockIdx.x; rb < p.n_row_blocks; rb += gridDim.x) {
        for (int pass = 0; pass < 2; ++pass) {
          const bool one = kOneTerm && pass == 0;                    // pass 0 of the one-term form reads the cohort's hi half only
          for (int t = 0; t < p.n_tiles; ++t)
            for (int j = 0; j < (one ? kb_n : items); ++j) {
              mbar_wait(&B.b_empty[s], ph ^ 1);
              mbar_expect_tx(&B.b_full[s], kBoxBytes);
              tma_load_2d(b_smem + static_cast<size_t>(s) * kBoxBytes, &map_b, &B.b_full[s], one ? j * 64 : (j & 1) * p.dp + (j >> 1) * 64, t * 128);
              if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // descriptors as {32-bit low word, common high word}, computed by the converged warp; only the tcgen05 instructions sit under
    // the elected lane, unrolled with compile-time offsets (64-bit descriptor arithmetic inside the elected region costs ~200 ns
    // per MMA: the first build of this kernel spent 13 us per tile on 48 MMAs)
    const uint64_t dbase = make_kmajor_desc(0, 1024u, 2u);      // SWIZZLE_128B, 8-row groups 1024 B apart
    const uint32_t hi = static_cast<uint32_t>(dbase >> 32);
    const uint32_t lo0 = static_cast<uint32_t>(dbase);
    const uint32_t idesc = make_idesc_f16(1u, 128u, 128u);      // bf16 operands, fp32 accumulate, M = N = 128
    const uint32_t b_lo0 = lo0 + (smem_u32(b_smem) >> 4);
    const uint32_t a_t0 = tmem_base + 256u;                      // the test operand: lane = row, hi in columns [256, 256 + dp/2), lo right after
    constexpr uint32_t kBox16 = kBoxBytes >> 4;
    uint32_t tt = 0;
    int lb = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int rb = blockIdx.x; rb < p.n_row_blocks; rb += gridDim.x, ++lb) {
      mbar_wait(&B.a_full, lb & 1);
      tc_fence_after();
      for (int pass = 0; pass < 2; ++pass)
        for (int t = 0; t < p.n_tiles; ++t, ++tt) {
          const int buf = tt & 1;
          mbar_wait(&B.t_empty[buf], ((tt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * 128);
#pragma unroll 1
          for (int kb = 0; kb < kb_n; ++kb) {
            const uint32_t ah = a_t0 + static_cast<uint32_t>(kb) * 32u, al = a_t0 + static_cast<uint32_t>(kb_n + kb) * 32u;   // 64 elements = 32 columns
            {                                                     // cohort hi box: xh.ch then xl.ch
              mbar_wait(&B.b_full[s], ph);
              tc_fence_after();
              const uint32_t bd = b_lo0 + static_cast<uint32_t>(s) * kBox16;
              const uint32_t first = kb != 0 ? 1u : 0u;
              const bool one = kOneTerm && pass == 0;
              if (elect_one()) {
                if (!(p.knock & 4)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts_lo(d_tmem, ah + 8 * k, bd + 2 * k, hi, idesc, k == 0 ? first : 1u);
                if (!one) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts_lo(d_tmem, al + 8 * k, bd + 2 * k, hi, idesc, 1u);
                }
                }
                umma_commit(&B.b_empty[s]);
                if (one && kb == kb_n - 1) umma_commit(&B.t_full[buf]);
              }
              __syncwarp();
              if (++s == p.stages) { s = 0; ph ^= 1; }
              if (one) continue;
            }
            {                                                     // cohort lo box: xh.cl
              mbar_wait(&B.b_full[s], ph);
              tc_fence_after();
              const uint32_t bd = b_lo0 + static_cast<uint32_t>(s) * kBox16;
              if (elect_one()) {
                if (!(p.knock & 4)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts_lo(d_tmem, ah + 8 * k, bd + 2 * k, hi, idesc, 1u);
                }
                umma_commit(&B.b_empty[s]);
                if (kb == kb_n - 1) umma_commit(&B.t_full[buf]);
              }
              __syncwarp();
              if (++s == p.stages) { s = 0; ph ^= 1; }
            }
          }
        }
      if (elect_one()) umma_commit(&B.a_empty);                  // every MMA that reads this row block's operand has retired
      __syncwarp();
    }
  } else if (warp >= 4) {
    // Two epilogue warpgroups: group g takes the tiles whose accumulator buffer is g (even / odd tiles), so every test row has two
    // threads, each with a private histogram, private partial sums and a private candidate list; group 0 computes the bin range
    // from the row block's first tile and finishes the row.  Three named barriers per row block.
    const int g = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int nb = p.nb, cap = p.cap;                            // cap: candidates per GROUP
    const uint32_t aux_u32 = smem_u32(aux);
    const uint32_t aux_row = aux_u32 + static_cast<uint32_t>(row) * 4u;           // list word w of this row: aux_row + w * 512
    // one group's histogram: 16-bit counts, two bins per 32-bit word, word index == row (mod 128) -> bank = lane, no conflicts:
    // bin b of row r at (b >> 1) * 512 + r * 4 + (b & 1) * 2  ==  b * 256 + r * 4 - (b & 1) * 254
    const uint32_t hist_bytes = asnorm_hist_bytes(nb);
    const uint32_t hist_row0 = aux_row, hist_row1 = hist_row0 + hist_bytes;
    const uint32_t hist_row = g ? hist_row1 : hist_row0;                         // this group's PRIVATE histogram (pass 0)
    const uint32_t list_row = aux_row + static_cast<uint32_t>(g * (cap + kListSlack)) * 512u;   // this group's candidate list (pass 1)
    const uint32_t list_end = list_row + static_cast<uint32_t>(cap) * 512u;
    const uint32_t part_row = aux_row + asnorm_main_bytes(nb, cap);              // 8 words: range (pass 0), then group 1's partial results
    const uint32_t trash_row = part_row + 8u * 512u + static_cast<uint32_t>(g) * 512u;   // where the unconditional stores of unselected scores go
    const float top = __int_as_float(__float_as_int(static_cast<float>(nb + 1)) - 1);   // the float just below nb + 1
    const int k = p.topk;
    uint32_t tt = 0;
    int lbe = 0;
    for (int rb = blockIdx.x; rb < p.n_row_blocks; rb += gridDim.x) {
      const long long grow = static_cast<long long>(rb) * 128 + row;
      float off1 = 0.f, scale = 0.f;
      int bstar = -2, need = 0, above = 0;
      bool bad = false;
      double dsum = 0.0, dsq = 0.0;
      int pos = 0;
      uint32_t lptr = list_row;                                   // next free word of this group's candidate list
      float thr_gt = __int_as_float(0x7f800000), thr_ge = __int_as_float(0x7f800000);
      float* vo = (VALS && p.vals && grow < p.n_rows) ? p.vals + grow * p.vals_ld : nullptr;
      if (g == 0) {
        // this row's operand into tensor memory: x = hi + lo in bf16, two elements per 32-bit column (what tcgen05.mma reads as a
        // K-major A row).  The MMAs of the previous row block must have retired first.
        mbar_wait(&B.a_empty, (lbe & 1) ^ 1);
        tc_fence_after();
        const uint32_t a_t = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + 256u;
        const float* xr = p.x + grow * p.d;
        const bool live = grow < p.n_rows;
        for (int c0 = 0; c0 < p.dp; c0 += 32) {
          uint32_t wh[16], wl[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c0 + 4 * q < p.d) f = __ldg(reinterpret_cast<const float4*>(xr + c0 + 4 * q));
            const __nv_bfloat16 h0 = __float2bfloat16_rn(f.x), h1 = __float2bfloat16_rn(f.y), h2 = __float2bfloat16_rn(f.z), h3 = __float2bfloat16_rn(f.w);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(f.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f.y - __bfloat162float(h1));
            const __nv_bfloat16 l2 = __float2bfloat16_rn(f.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(f.w - __bfloat162float(h3));
            wh[2 * q] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
            wh[2 * q + 1] = static_cast<uint32_t>(__bfloat16_as_ushort(h2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h3)) << 16);
            wl[2 * q] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
            wl[2 * q + 1] = static_cast<uint32_t>(__bfloat16_as_ushort(l2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l3)) << 16);
          }
          tmem_st16(a_t + static_cast<uint32_t>(c0 >> 1), wh);
          tmem_st16(a_t + static_cast<uint32_t>((p.dp + c0) >> 1), wl);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.a_full);
      }
      ++lbe;
      // Both passes walk the tile through eight TMEM loads of 16 columns, load c+1 in flight while the scores of load c are
      // processed; the accumulator buffer goes back to the MMA warp as soon as the last load has landed.  The loop bodies are kept
      // SMALL on purpose (two 16-score blocks, not unrolled further): an earlier build with the whole tile unrolled for both
      // passes was 150 KB of code and ran 20 % slower than the same arithmetic in 40 KB — four roles in different code regions
      // share one instruction cache.
      const int npad = p.n_tiles * 128 - p.c;                      // zero cohort rows behind the last tile: each scores exactly 0
#define SVX_TILE_CHUNKS(PROCESS)                                                              \
      {                                                                                       \
        uint32_t ra[16], rq[16];                                                              \
        tmem_ld16(taddr, ra);                                                                 \
        _Pragma("unroll 1")                                                                   \
        for (uint32_t c0 = 0; c0 < 128u; c0 += 32u) {                                         \
          tmem_ld_wait_dep16(ra);                                                             \
          tmem_ld16(taddr + c0 + 16u, rq);                                                    \
          PROCESS(ra);                                                                        \
          tmem_ld_wait_dep16(rq);                                                             \
          if (c0 + 32u < 128u) tmem_ld16(taddr + c0 + 32u, ra);                               \
          else { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&B.t_empty[buf]); } \
          PROCESS(rq);                                                                        \
        }                                                                                     \
      }
      // ------------------------------------------------------------------ pass 0: histogram
      for (int t = 0; t < p.n_tiles; ++t, ++tt) {
        const int buf = tt & 1;
        if (t == 1 && g == 1) {                                   // group 1's first tile: the range is published, the histograms zeroed
          named_bar_sync(1, 256);
          off1 = __uint_as_float(lds_u32(part_row)); scale = __uint_as_float(lds_u32(part_row + 512u));
        }
        if (buf != g) continue;
        mbar_wait(&B.t_full[buf], (tt >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(buf * 128);
        if (t == 0) {
          // bin range from the statistics of the first tile's scores (a 128-score sample of this row; c >= 1024: a full tile)
          uint32_t r[16];
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 16) {
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float v = __uint_as_float(r[i]); s1 += v; s2 = fmaf(v, v, s2); }
          }
          const float mu = s1 * (1.f / 128.f);
          const float sd = sqrtf(fmaxf(s2 * (1.f / 128.f) - mu * mu, 1e-30f));
          const float lo = mu + p.z_lo * sd;
          scale = 1.f / ((p.z_hi - p.z_lo) * sd);                 // (v - lo) / range: [0, 1] over the bin range
          off1 = -lo * scale;
          for (uint32_t w = 0; w < 2u * hist_bytes; w += 512u) sts_u32(hist_row0 + w, 0u);
          sts_u32(part_row, __float_as_uint(off1)); sts_u32(part_row + 512u, __float_as_uint(scale));
          named_bar_sync(1, 256);
        }
        if (p.knock & 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&B.t_empty[buf]); continue; }
        // No shared-memory atomics (ATOMS costs ~2 cycles per LANE, and a predicated one becomes a branch region): every group has
        // a PRIVATE histogram, so a count is a plain 16-bit load / add / store of a word only this thread touches.  Four scores at
        // a time: their loads are in flight together, and a score whose bin an earlier score of the same four also hits carries
        // that score's increment (the later store wins).  Stored bin 0 collects the scores below the range and is never read.
        const bool skip0 = (p.knock & (2 | 8)) != 0;
        auto process0 = [&](uint32_t (&r)[16]) {
          if (skip0) return;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            int b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = asnorm_bin(__uint_as_float(r[i + j]), scale, off1, top);
            const uint32_t a0 = hist_addr(hist_row, b[0]), a1 = hist_addr(hist_row, b[1]), a2 = hist_addr(hist_row, b[2]), a3 = hist_addr(hist_row, b[3]);
            const uint32_t i1 = 1u + (b[1] == b[0] ? 1u : 0u);
            const uint32_t i2 = 1u + (b[2] == b[0] ? 1u : 0u) + (b[2] == b[1] ? 1u : 0u);
            const uint32_t i3 = 1u + (b[3] == b[0] ? 1u : 0u) + (b[3] == b[1] ? 1u : 0u) + (b[3] == b[2] ? 1u : 0u);
            const uint32_t h0 = lds_u16(a0), h1 = lds_u16(a1), h2 = lds_u16(a2), h3 = lds_u16(a3);
            sts_u16(a0, h0 + 1u); sts_u16(a1, h1 + i1); sts_u16(a2, h2 + i2); sts_u16(a3, h3 + i3);
          }
        };
        SVX_TILE_CHUNKS(process0)
      }
      {
        __threadfence_block();
        named_bar_sync(1, 256);                                   // both groups' histograms are complete
        // walk the bins from the top: b* holds the k-th largest score (both threads of a row find the same one); the zero
        // padding's scores are taken out of the bin of 0.0
        const int bin0 = asnorm_bin(0.f, scale, off1, top);
        int cum = 0;
        bstar = -1;
        for (int b = nb; b >= 1; --b) {
          int h = static_cast<int>(lds_u16(hist_addr(hist_row0, b)) + lds_u16(hist_addr(hist_row1, b)));
          if (b == bin0) h -= npad;
          if (cum + h >= k) { bstar = b; need = k - cum; above = cum; bad = h > 2 * cap - 2; break; }
          cum += h;
        }
        if (bstar < 0) bad = true;                                // fewer than k scores inside the bin range
        // the binning is monotone in the score, so "bin >= b" is "score >= the smallest float whose bin is b": two compares per
        // score in pass 1 instead of the binning arithmetic
        if (kOneTerm) {
          bad = bstar < 2 || bstar + 2 > nb;                      // the band needs a real edge on both sides
          if (!bad) {
            thr_ge = asnorm_bin_threshold(bstar - 1, scale, off1, top);
            thr_gt = asnorm_bin_threshold(bstar + 2, scale, off1, top);
          }
        } else if (!bad) {
          thr_ge = asnorm_bin_threshold(bstar, scale, off1, top);
          thr_gt = asnorm_bin_threshold(bstar + 1, scale, off1, top);
        }                                                         // bad: both stay +inf, pass 1 selects nothing for this row
        // the zero padding scores 0.0 in pass 1 too: above the band it only inflates the count (0 adds nothing to the sums);
        // inside the band it would sit in the list — such a row (its k-th largest score is ~0) is handed back
        // (with the selected scores written out the padding must not be selected at all)
        if (npad > 0 && !bad && thr_ge <= 0.f && (VALS || !(thr_gt <= 0.f))) { bad = true; thr_ge = thr_gt = __int_as_float(0x7f800000); }
        named_bar_sync(1, 256);                                   // everybody has read the histograms: their words become the lists
      }
      // ------------------------------------------------------------------ pass 1: sums above the threshold, candidates around it
      for (int t = 0; t < p.n_tiles; ++t, ++tt) {
        const int buf = tt & 1;
        if (buf != g) continue;
        mbar_wait(&B.t_full[buf], (tt >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(buf * 128);
        if (p.knock & 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&B.t_empty[buf]); continue; }
        const bool skip1 = (p.knock & (2 | 16)) != 0;
        // fp32 partial sums over ONE tile (only the scores above the threshold are non-zero: a handful per tile), the running sums in
        // double once per tile: per 16 scores the two F2F.F64 + DADD pairs were 12 % of the epilogue warps' stall samples (FP64 pipe)
        float s16 = 0.f, q16 = 0.f;
        auto process1 = [&](uint32_t (&r)[16]) {
          if (skip1) return;
#pragma unroll
          for (int i4 = 0; i4 < 16; i4 += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t rv = r[i4 + j];
              const float v = __uint_as_float(rv);
              const bool gt = v >= thr_gt;                        // above the threshold bin(s): in the top-k for certain
              const bool ge = v >= thr_ge;                        // in or above them
              const float vs = gt ? v : 0.f;
              s16 += vs;
              q16 = fmaf(vs, vs, q16);
              // (cohort-sharded layout only) the selected scores themselves: group 0 fills the row from the front, group 1 from
              // position above-1 backwards — together they write exactly `above` scores
              if constexpr (VALS) { if (vo != nullptr && gt && pos < above) vo[g == 0 ? pos : above - 1 - pos] = v; }
              pos += gt ? 1 : 0;
              const bool eq = ge && !gt;
              sts_u32(eq ? lptr : trash_row, rv);                 // unconditional store (a predicated one becomes a branch region)
              lptr += eq ? 512u : 0u;
            }
            lptr = min(lptr, list_end);                           // an overfull list stays inside its kListSlack spare words; the row is handed back
          }
        };
        SVX_TILE_CHUNKS(process1)
        dsum += static_cast<double>(s16);
        dsq += static_cast<double>(q16);
      }
#undef SVX_TILE_CHUNKS
      if (thr_gt <= 0.f) pos -= (g == ((p.n_tiles - 1 + p.n_tiles) & 1)) ? npad : 0;     // the padding counted by the group that took the last tile
      const int cnt = static_cast<int>((lptr - list_row) >> 9);
      if (g == 1) {                                               // hand this group's part to group 0
        sts_u32(part_row + 2 * 512u, static_cast<uint32_t>(cnt)); sts_u32(part_row + 3 * 512u, static_cast<uint32_t>(pos));
        const unsigned long long a = __double_as_longlong(dsum), b = __double_as_longlong(dsq);
        sts_u32(part_row + 4 * 512u, static_cast<uint32_t>(a)); sts_u32(part_row + 5 * 512u, static_cast<uint32_t>(a >> 32));
        sts_u32(part_row + 6 * 512u, static_cast<uint32_t>(b)); sts_u32(part_row + 7 * 512u, static_cast<uint32_t>(b >> 32));
      }
      named_bar_sync(1, 256);
      if (g == 0 && !p.knock) {
        const int cnt1 = static_cast<int>(lds_u32(part_row + 2 * 512u)), pos1 = static_cast<int>(lds_u32(part_row + 3 * 512u));
        dsum += __longlong_as_double(static_cast<long long>(lds_u32(part_row + 4 * 512u)) | (static_cast<long long>(lds_u32(part_row + 5 * 512u)) << 32));
        dsq += __longlong_as_double(static_cast<long long>(lds_u32(part_row + 6 * 512u)) | (static_cast<long long>(lds_u32(part_row + 7 * 512u)) << 32));
        if (kOneTerm) { above = pos + pos1; need = k - above; }     // counted on the exact scores
        if (!bad && (cnt >= cap || cnt1 >= cap || pos + pos1 != above || need < 0 || cnt + cnt1 < need)) bad = true;
        if (grow < p.n_rows) {
          if (bad) {
            const int slot = atomicAdd(p.flag_count, 1);
            p.flag_rows[slot] = static_cast<int>(grow);
          } else {
            // the `need` largest members of the threshold bin: group 1's candidates are appended to group 0's, then a partial
            // selection sort over this row's list words
            const uint32_t list1 = aux_row + static_cast<uint32_t>(cap + kListSlack) * 512u;
            for (int j = 0; j < cnt1; ++j) sts_u32(aux_row + static_cast<uint32_t>(cnt + j) * 512u, lds_u32(list1 + static_cast<uint32_t>(j) * 512u));
            const int total = cnt + cnt1;
            int n_sel = need;
            double sgn = 1.0;
            if (!vo && 2 * need > total) {
              // more than half of the list is wanted: add all of it, then take the total - need LARGEST OF THE NEGATED values out again
              for (int j = 0; j < total; ++j) {
                const uint32_t a = aux_row + static_cast<uint32_t>(j) * 512u;
                const float v = __uint_as_float(lds_u32(a));
                dsum += static_cast<double>(v);
                dsq = fma(static_cast<double>(v), static_cast<double>(v), dsq);
                sts_u32(a, __float_as_uint(-v));
              }
              n_sel = total - need;
              sgn = -1.0;
            }
            for (int i = 0; i < n_sel; ++i) {
              int best = i;
              float bv = __uint_as_float(lds_u32(aux_row + static_cast<uint32_t>(i) * 512u));
              const float first_v = bv;
              for (int j = i + 1; j < total; ++j) {
                const float v = __uint_as_float(lds_u32(aux_row + static_cast<uint32_t>(j) * 512u));
                if (v > bv) { bv = v; best = j; }
              }
              if (best != i) sts_u32(aux_row + static_cast<uint32_t>(best) * 512u, __float_as_uint(first_v));
              dsum += static_cast<double>(bv);                    // (negated list: bv = -v removes v)
              dsq = fma(sgn * static_cast<double>(bv), static_cast<double>(bv), dsq);
              if (vo) vo[above + i] = bv;
            }
            if (vo)
              for (int i = k; i < p.vals_ld; ++i) vo[i] = kPad;
            const double mean = dsum / static_cast<double>(k);
            const double var = dsq / static_cast<double>(k) - mean * mean;
            if (p.mean) p.mean[grow] = static_cast<float>(mean);
            if (p.stdv) p.stdv[grow] = static_cast<float>(sqrt(var > 0.0 ? var : 0.0));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// fp32 -> split bf16 rows [hi(dp) | lo(dp)], zero padded to dp columns and n_pad rows.
__global__ void __launch_bounds__(256) split2_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n, long long n_pad,
                                                     int d, int dp) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_pad * dp) return;
  const long long row = idx / dp;
  const int j = static_cast<int>(idx - row * dp);
  const float v = (row < n && j < d) ? in[row * d + j] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  out[row * (2LL * dp) + j] = hi;
  out[row * (2LL * dp) + dp + j] = lo;
}

cudaError_t launch_split2(const float* in, __nv_bfloat16* out, long long n, long long n_pad, int d, int dp, cudaStream_t st) {
  const long long total = n_pad * dp;
  if (total <= 0) return cudaSuccess;
  split2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, n, n_pad, d, dp);
  return cudaGetLastError();
}

size_t asnorm_fused_smem_bytes(const AsnormFusedParams& p) {
  // the histogram / list area, then the 8 partial-result words and two trash words of every row
  const size_t aux_words = asnorm_main_bytes(p.nb, p.cap) / 512 + 10;
  return 1024 + 1024 + static_cast<size_t>(p.stages) * kBoxBytes + aux_words * 512;
}

cudaError_t asnorm_fused_init() {
  cudaError_t e = cudaFuncSetAttribute(asnorm_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(asnorm_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_asnorm_fused(const AsnormFusedParams& p, const CUtensorMap& map_b, int sms, cudaStream_t st) {
  if (p.n_row_blocks <= 0) return cudaSuccess;
  const int grid = p.n_row_blocks < sms ? p.n_row_blocks : sms;
  if (p.vals) asnorm_fused_kernel<true><<<grid, kThreads, asnorm_fused_smem_bytes(p), st>>>(p, map_b);
  else asnorm_fused_kernel<false><<<grid, kThreads, asnorm_fused_smem_bytes(p), st>>>(p, map_b);
  return cudaGetLastError();
}

// gather / scatter of the rows the fused kernel handed back
__global__ void gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ rows, int n, int d, float* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx - static_cast<long long>(i) * d);
  out[idx] = in[static_cast<long long>(rows[i]) * d + j];
}
__global__ void scatter_rows_kernel(const float* __restrict__ in, const int* __restrict__ rows, int n, int d, float* __restrict__ out, int out_ld) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx - static_cast<long long>(i) * d);
  out[static_cast<long long>(rows[i]) * out_ld + j] = in[idx];
}
cudaError_t launch_gather_rows(const float* in, const int* rows, int n, int d, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gather_rows_kernel<<<static_cast<unsigned>((static_cast<long long>(n) * d + 255) / 256), 256, 0, st>>>(in, rows, n, d, out);
  return cudaGetLastError();
}
cudaError_t launch_scatter_rows(const float* in, const int* rows, int n, int d, float* out, int out_ld, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  scatter_rows_kernel<<<static_cast<unsigned>((static_cast<long long>(n) * d + 255) / 256), 256, 0, st>>>(in, rows, n, d, out, out_ld);
  return cudaGetLastError();
}

}  // namespace svx
