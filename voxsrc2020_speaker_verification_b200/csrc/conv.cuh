// Shared parameter blocks for the convolution kernels (tcgen05 implicit-GEMM and the plain CUDA-core form).
//
// Activation layout ("tall image"): all segments (≤1000-frame chunks of utterances) of one call are stacked
// along the time axis into one [rows, W, C] NHWC image per resolution stage, with at least one all-zero row
// between consecutive segments.  The zero rows are the convolution's time padding, so one uniform tiling of
// the tall image is exact for every segment, whatever its length.  seg_of_row[row] >= 0 marks rows that
// belong to a segment; every epilogue writes zeros to the other rows it covers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

namespace svx {

// Debug / tuning switches (SVX_* environment variables) exist only in the debug build (`build.py --debug` → libsvx_dbg.so,
// -DSVX_DEBUG_SWITCHES); in the production library every one of them folds to "unset" at compile time.
#ifdef SVX_DEBUG_SWITCHES
inline const char* dbg_env(const char* name) { return getenv(name); }
#else
inline const char* dbg_env(const char*) { return nullptr; }
#endif

constexpr int kMaxTaps = 9;
constexpr int kTraceEvents = 4096;

// Where a conv's output goes and what is fused on the way (BN as scale/shift, ReLU, residual, concat).
struct Epilogue {
  const float* scale;   // [n_total] or nullptr (=1)
  const float* shift;   // [n_total] or nullptr (=0)
  int pre_relu;         // relu(acc) before scale/shift   (TDNN: conv → ReLU → BN, tdnn_model.py:25-29)
  int post_relu;        // relu after scale/shift/residual (Res2Net: conv → BN → ReLU)
  int n_valid;          // real output channels (the MMA N may be padded up)
  // primary destination: channels [0, n_split) → out[(row*W + col)*out_C + out_coff + c]
  void* out; int out_C; int out_coff;
  // optional residual added before post_relu, same channel range
  const void* res; int res_C; int res_coff;
  // channels [n_split, n_valid) → outb[... outb_coff + (c - n_split)], no residual
  int n_split;
  void* outb; int outb_C; int outb_coff;
  // optional second output for channels [0, n_split): out2 = v + add2   (Res2Net x_{i+1} + o_i, res2net_model.py:65-66)
  void* out2; int out2_C; int out2_coff;
  const void* add2; int add2_C; int add2_coff;
  // planar splits (Res2Net conv1 → hierarchical 3x3 inputs): channel c = s*split_wp + j (j < split_w) goes to
  // split_ptr[s][pix*split_C[s] + split_coff[s] + j]; j >= split_w are padding channels.  n_splits = 0: unused.
  int n_splits, split_wp, split_w;
  void* split_ptr[8]; int split_C[8]; int split_coff[8];
  // fp32 row-major output instead of 16-bit NHWC (scoring GEMM): out_f32[row*ldf + c]
  float* out_f32; int ldf;
  const int32_t* seg_of_row;   // nullptr → every row valid
};

// Geometry for the plain CUDA-core kernel (any stride / dilation / groups).
struct SimpleConvParams {
  const void* in; int in_C; int in_coff; int in_rows; int in_W; int in_Wp;   // in_Wp: pixels per row in memory (W + zero column)
  const void* wgt;      // [n_pad][taps*kpad] K-major, same buffer the UMMA path uses
  int kpad;             // padded input channels per tap
  int cin_g;            // input channels per group
  int cout_g;           // output channels per group
  int grp_ntile;        // grouped weights: K positions of row n are relative to input channel
  int grp_cstep;        //   (n / grp_ntile) * grp_cstep   (0/0 for dense: relative to channel 0)
  int kh, kw, sh, sw, dh, dw, ph, pw;
  int out_rows, out_W, out_Wp;
  Epilogue epi;
};

// Geometry for the tcgen05 kernel.  One CTA = one 128-pixel (h_box x w_box) x n_tile output tile.
struct UmmaConvParams {
  int out_rows, out_W;
  int out_Wp;                  // row pitch of the output image in pixels (W + zero column)
  int w_box, h_box, w_tiles;
  int taps;
  int8_t tap_map[kMaxTaps];    // which of the (up to 4) A tensor maps the tap reads (stride-2 parity views)
  int8_t tap_dh[kMaxTaps];     // row / col displacement of the tap in that map's coordinates
  int8_t tap_dw[kMaxTaps];
  int nkc;                     // k-boxes per tap
  int kbox;                    // elements per k-box: 64 / 32 / 16 (= swizzle span / 2)
  int a_c_step;                // grouped conv: input-channel offset per n-tile (0 for dense)
  int n_tile;                  // UMMA N (multiple of 16, ≤ 256)
  int n_tiles;                 // number of n-tiles (Cout padded / n_tile)
  int stages;
  int aux_mode;                // 0 none, 1 residual added before post-ReLU, 2 second output = v + add2
  int aux_boxes;               // 64-channel TMA boxes per tile (≤ ceil(n_tile/64))
  int aux_width;               // channels of the aux slice (boxes past it are skipped)
  uint32_t aux_bytes;          // bytes of one aux buffer
  uint32_t ss_bytes;           // shared-memory bytes of the staged scale/shift vectors
  int store_mode;              // 0: direct 16-byte stores (fp32 output form); 1: swizzled smem boxes + TMA stores
  uint32_t stage_bytes;        // bytes of one output staging buffer (0 when the residual tile is reused in place)
  uint32_t bres_bytes;         // bytes of the resident weight region (0 → weights stream through the ring)
  unsigned long long* trace;   // debug: per-role event timestamps of CTA 0 (nullptr = off); [4 roles][kTraceEvents]
  uint32_t idesc, sbo, layout_type;
  uint32_t a_stage_bytes, b_stage_bytes, tmem_cols;
  Epilogue epi;
};

struct AMaps { CUtensorMap m[4]; };
struct OMaps { CUtensorMap m[2]; };   // TMA-store maps: [0] primary destination slice, [1] out2 slice

template <typename T> struct TypeOps;
template <> struct TypeOps<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) {
    return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
};
template <> struct TypeOps<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
};

// Scalar epilogue used by the CUDA-core kernel (and as the semantic definition of the fused one).
template <typename T>
__device__ __forceinline__ void epilogue_scalar(const Epilogue& e, float acc, int row, int col, int Wp, int c, bool valid) {
  if (c >= e.n_valid) return;
  float v = acc;
  if (e.pre_relu) v = fmaxf(v, 0.f);
  if (e.scale) v *= e.scale[c];
  if (e.shift) v += e.shift[c];
  const size_t pix = static_cast<size_t>(row) * Wp + col;
  if (e.n_splits > 0) {
    const int s_ = c / e.split_wp, j_ = c - s_ * e.split_wp;
    if (j_ >= e.split_w) return;
    if (e.post_relu) v = fmaxf(v, 0.f);
    if (!valid) v = 0.f;
    static_cast<T*>(e.split_ptr[s_])[pix * e.split_C[s_] + e.split_coff[s_] + j_] = TypeOps<T>::from_f(v);
    return;
  }
  if (e.out_f32) {
    e.out_f32[static_cast<size_t>(row) * e.ldf + c] = valid ? v : 0.f;
    return;
  }
  if (c < e.n_split) {
    if (e.res) v += TypeOps<T>::to_f(static_cast<const T*>(e.res)[pix * e.res_C + e.res_coff + c]);
    if (e.post_relu) v = fmaxf(v, 0.f);
    if (!valid) v = 0.f;
    static_cast<T*>(e.out)[pix * e.out_C + e.out_coff + c] = TypeOps<T>::from_f(v);
    if (e.out2) {
      float s = valid ? v + TypeOps<T>::to_f(static_cast<const T*>(e.add2)[pix * e.add2_C + e.add2_coff + c]) : 0.f;
      static_cast<T*>(e.out2)[pix * e.out2_C + e.out2_coff + c] = TypeOps<T>::from_f(s);
    }
  } else {
    if (e.post_relu) v = fmaxf(v, 0.f);
    if (!valid) v = 0.f;
    static_cast<T*>(e.outb)[pix * e.outb_C + e.outb_coff + (c - e.n_split)] = TypeOps<T>::from_f(v);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Flat implicit GEMM (conv_flat.cu): stride-1 convs over the image seen as a 1-D pixel sequence p = row*Wp + col
// (Wp = W + one zero column), so that filter tap (dh, dw) is the constant shift dh*Wp + dw.  One haloed span of
// pixels is loaded once per K-box; the taps are shifted shared-memory descriptors into it.
struct FlatConvParams {
  long long P;                 // pixels to cover in this launch (rows_used * Wp)
  int mt;                      // 128-pixel sub-tiles per span (share one haloed A load)
  int halo;                    // max |tap shift|
  int a_box_rows, a_boxes;     // A stage = a_boxes TMA boxes of a_box_rows pixel rows (<= 256, multiple of 8)
  int taps;
  int tap_shift[kMaxTaps];
  int nkc, kbox, kpad;         // K boxes per tap, elements per box (64/32/16), nkc*kbox
  int a_c_step;                // grouped conv (block-diagonal weights per n-tile): input-channel offset per n-tile; 0 for dense
  int n_tile, n_tiles;
  int a_stages, b_stages;
  uint32_t a_stage_bytes, b_item_bytes;
  int b_resident;
  uint32_t idesc, sbo, layout_type, tmem_cols;
  int tmem_bufs, tmem_bufs_log2;   // accumulator buffers in TMEM (2 or 4): with 4 an epilogue warpgroup may lag a whole span behind the MMAs
  int cn, cm;                  // multicast cluster cm x cn (1 x 1: none): cn CTAs share a span (A slices multicast), cm CTAs share an n-tile (weight slices multicast)
  int a_slice_rows;            // rows of every A box this CTA loads (a_box_rows / cn)
  int b_slice_rows;            // rows of every streamed weight item this CTA loads (n_tile / cm)
  int b_rows;                  // weight rows per item (= n_tile)
  // epilogue
  const float* scale; const float* shift;
  int n_valid;                 // real output channels
  int n_res;                   // residual (aux mode 1) applies to channels < n_res
  int grp_mask, grp_w;         // padded planar splits: channel groups with (c & grp_mask) >= grp_w are padding and are skipped
  const uint8_t* pix_valid;    // [P_cap] 1 = pixel belongs to a segment and is not the zero column
  int aux_mode;                // 0 none, 1 residual added before post-ReLU (in place), 2 second output = v + add2
  int box_ch;                  // channels per staging box: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int boxes;                   // staging boxes per buffer = ceil(part_cols / box_ch)
  int n_parts, part_cols;      // a slot holds part_cols = n_tile / n_parts columns: 256-wide tiles go through the slots in two halves
  int slots;                   // epilogue slot ring depth PER WARPGROUP (each of the two epilogue warpgroups owns its ring)
  uint32_t slot_bytes;
  uint8_t route_map[32];       // per global staging box (n0/box_ch + b): which output map it is stored through (0xff: none)
  int32_t route_c[32];         //   and at which channel coordinate of that map
  int pre_relu, post_relu;
  int reverse;                 // walk the spans from the last pixel to the first: consecutive layers alternate, so a consumer starts on
                               // the pixels its producer wrote last (still in the 126 MB L2)
  unsigned long long* trace;
  // direct epilogue (narrow tiles, n_tile <= 64, one destination): the epilogue threads read the aux tile and write the outputs
  // with their own 16-byte global accesses (a thread owns a pixel = one contiguous channel run), no slots, no TMA for them.
  // With 64-byte rows the TMA engine's ~2.5 ns per box ROW made the aux loads + stores the bound of every narrow 3x3 conv.
  int direct;
  int n_store2;                // same for the aux tile and the second output (padded planar tensors)
  int n_store;                 // direct epilogue: channels written to the primary output (>= n_valid: zeros in the pad of a padded concat slice)
  int lin;                     // aux mode 2 with add2 / out2 as DENSE planar tensors: their 128-pixel tiles move as 1-D bulk copies (kernel AUX = 3)
  long long P_cap;             // pixels allocated (stores/loads of the direct epilogue are clipped here, like the tensor maps clip)
  uint8_t* d_out; uint32_t d_out_pitch;        // primary output slice (first channel of the slice), bytes per pixel
  uint8_t* d_out2; uint32_t d_out2_pitch;      // second output (aux mode 2)
  const uint8_t* d_aux; uint32_t d_aux_pitch;  // residual (aux mode 1) / add2 (aux mode 2)
  int knock;                   // debug timing experiments only (SVX_FLAT_KNOCK): 1 skip epilogue math, 2 skip TMA stores, 4 skip aux loads, 8 skip tcgen05.ld, 32 skip the MMAs, 64 skip the direct epilogue's primary stores
  unsigned long long* dbg;     // host-mapped words: which barrier wait timed out (written before the trap)
};
struct FlatMaps { CUtensorMap a, b, aux, o2, o[8]; };   // o2: second output of aux mode 2
cudaError_t conv_flat_init();
int conv_flat_max_clusters(int cluster_size);   // resident clusters of that many CTAs (one CTA per SM)
size_t conv_flat_smem_bytes(const FlatConvParams& p);
cudaError_t launch_conv_flat(const FlatConvParams& p, const FlatMaps& maps, int is_bf16, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------------------
// Fused hierarchical 3x3 chain of a stride-1 Res2Net block with 4 splits of <= 32 channels (res2_chain.cu): three dependent
// 3x3 convs, the running sums x_{i+1} + y_i kept in shared memory.  Same flat pixel sequence as FlatConvParams.
struct ChainParams {
  long long P;                 // pixels to cover (rows_used * Wp)
  long long P_cap;             // pixels allocated
  int tap_shift[kMaxTaps];     // dh*Wp + dw of the 9 taps
  uint32_t idesc;              // M = 128, N = 32
  const float* scale[3]; const float* shift[3];   // folded BN of the three convs, 32 entries each (pad channels: 0)
  const uint8_t* pix_valid;
  uint8_t* y; uint32_t y_pitch;                   // concat: slice k of a pixel at y + pixel*y_pitch + k*64 bytes
  unsigned long long* dbg;
};
struct ChainMaps { CUtensorMap x[3], w[3]; };   // planar splits x_0..x_2 ({32 ch, 128 px} boxes) and the three weight matrices ({32 k, 32 n} per tap)
cudaError_t res2_chain_init();
size_t res2_chain_smem_bytes();
cudaError_t launch_res2_chain(const ChainParams& p, const ChainMaps& maps, int is_bf16, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------------------
// CTA-pair GEMM for the deep 1x1 stride-1 convs (conv_pair.cu): one M = 256 x n_tile tile per pair of CTAs (cta_group::2),
// every CTA loads its own 128 pixels and half of the weight rows.  Output leaves in 16-channel groups (32 bytes) through a
// routing table, so dense outputs, planar splits and a second destination are the same code.
struct PairConvParams {
  long long P, P_cap;          // pixels to cover / allocated
  // A tile's main loop is a list of K BOXES (one A load each) and every box carries ITEMS (tap displacement, weight column, K steps):
  //   1x1: one box per 64 input channels, one item each;  3x3 stride 1: the same boxes, nine items each (shifts of the flat pixel
  //   sequence);  stride 2 (2-D tile mode): one box per parity phase and 64 channels, 1 / 2 / 2 / 4 items per phase for a 3x3.
  int n_boxes;                 // <= 16
  uint8_t box_map[24];         // A tensor map of the box: 0 = PairMaps::a, 1..3 = PairMaps::a2[box_map - 1] (parity phases)
  int16_t box_c[24];           // channel coordinate of the box
  uint8_t box_item0[25];       // items of box b: [box_item0[b], box_item0[b + 1])
  uint16_t item_off16[24];     // (row displacement of the item's tap inside the box) * 128 bytes >> 4
  uint16_t item_wcol[24];      // weight column of the item (coordinate 0 of the B tensor map)
  uint8_t item_ks[24];         // K = 16 steps of the item that hold real channels (1..4)
  int halo;                    // flat modes: the A box starts halo pixels before the tile
  int a_rows; uint32_t a_bytes;         // rows of one A box (flat: 128 + 2 halo; 2-D: tile_bw * box_h) and its bytes rounded up to 1 KB
  int b_resident; uint32_t b_item_bytes;   // weights resident in shared memory (one item of n_tile/2 rows each) instead of streamed with A (one item per box)
  // 2-D tile mode (tile_bw != 0): a CTA tile is tile_h output rows; box = [box_h][tile_bw][64 ch] starting at (a_col0, first output row + a_row0)
  int tile_bw, tile_h, box_h, a_col0, a_row0;
  int out_wp;                  // pixels per output row in memory
  int rows;                    // output rows to cover (2-D tile mode; set per launch)
  int n_tile, n_tiles;         // 128, 192 or 256 output channels per pair tile
  int n_gemm;                  // GEMM N (n_tile * n_tiles, <= 1024)
  int stages; uint32_t stage_bytes;
  uint32_t idesc;              // M = 256, N = n_tile
  const float* scale; const float* shift; int n_valid;
  const uint8_t* pix_valid;
  int aux_mode;                // 0 none, 1 residual added before post-ReLU, 2 second output out2 = y + res (hierarchical 3x3, res = next split)
  int n_res;                   // residual applies to GEMM columns < n_res (multiple of 16)
  const uint8_t* res; uint32_t res_pitch;     // residual / add2 slice: column c of pixel p at res + p*res_pitch + 2c
  uint8_t* out2; uint32_t out2_pitch;         // aux mode 2: second output slice, same addressing
  int pre_relu, post_relu;
  int reverse;                 // walk the pixel blocks from the last to the first (see FlatConvParams::reverse)
  uint8_t route[64];           // per 16-column group: index into dst_base / dst_pitch, 0xff = not stored
  uint16_t goff[64];           //   and the byte offset of the group inside a pixel of that destination
  uint8_t* dst_base[10]; uint32_t dst_pitch[10];
  int knock;                   // debug timing experiments only (SVX_PAIR_KNOCK)
  unsigned long long* dbg;
};
struct PairMaps { CUtensorMap a, b, a2[3]; };   // a: {64 ch, a_rows px} boxes of the input (2-D tile mode: parity phase 0, 3-D), b: {64 k, n_tile/2 rows} boxes of the weights, a2: parity phases 1..3; SWIZZLE_128B
cudaError_t conv_pair_init();
size_t conv_pair_smem_bytes(const PairConvParams& p);
cudaError_t launch_conv_pair(const PairConvParams& p, const PairMaps& maps, int is_bf16, cudaStream_t stream);

cudaError_t launch_conv_simple(const SimpleConvParams& p, int is_bf16, cudaStream_t stream);
cudaError_t launch_conv_umma(const UmmaConvParams& p, const AMaps& amaps, const CUtensorMap& bmap, const CUtensorMap& auxmap,
                             const OMaps& omaps, int is_bf16, cudaStream_t stream);
bool conv_umma_finish_params(UmmaConvParams& p);   // stages / tmem_cols / aux_bytes from the tile shape
size_t conv_umma_smem_bytes(const UmmaConvParams& p);
cudaError_t conv_umma_init();   // sets max dynamic smem attribute

}  // namespace svx
