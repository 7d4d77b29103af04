// Extractor model: layer plans that restate the reference graph builders as fused GPU ops.
//   TDNN    reference tensorflow/models/tdnn_model.py:24-30,128-161
//   Res2Net reference tensorflow/models/res2net_model.py:26-136,185-262
//   DPN     reference tensorflow/models/dpn_model.py:24-171
//   ops     reference tensorflow/models/models.py:62-67 (BN), :107-203 (conv, padding), :262-269 (stats pool), :306-309 (dense)
#include "model.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kernels.cuh"
#include "umma.cuh"

namespace svx {

#define SVX_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int ceil_half(int a) { return (a + 1) / 2; }

constexpr float kBnEps4d = 1.001e-5f;   // TF fused batch norm clamps epsilon for 4-D inputs [ext]
constexpr float kBnEps2d = 1e-5f;       // models.py:20
constexpr float kPoolEps = 1e-5f;       // models.py:262
constexpr int kMaxChunk = 1000;         // tf_extract.py:96
constexpr int kMinFrames = 25;          // tf_extract.py:101-102

// ------------------------------------------------------------------------------------------------ TMA encode
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int encode_tmap(CUtensorMap* m, int is_bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes, int l2_promo_bytes) {
  PFN_encodeTiled fn = get_encode();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t gdim[5]; cuuint64_t gstr[5]; cuuint32_t bdim[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  l2_promo_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : l2_promo_bytes == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                                                                    : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u stride0 %llu", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
             box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    set_last_error(buf);
    return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ construction
Model::Model(const svx_model_config& cfg, int device, int precision) : cfg_(cfg), device_(device), is_bf16_(precision == SVX_PRECISION_BF16) {
  const char* fs = dbg_env("SVX_FORCE_SIMPLE");
  if (fs && fs[0] == '1') force_simple_ = 1;
  if (const char* dd = dbg_env("SVX_DUMP_DIR")) dump_dir_ = dd;
}

Model::~Model() {
  cudaSetDevice(device_);
  for (void* p : owned_) cudaFree(p);
  for (void* p : act_bufs_) cudaFree(p);
  for (auto p : d_seg_row_off_) cudaFree(p);
  for (auto p : d_seg_h_) cudaFree(p);
  for (auto p : d_seg_of_row_) cudaFree(p);
  for (auto p : d_pix_valid_) cudaFree(p);
  cudaFree(d_seg_frame_off_); cudaFree(d_seg_len_); cudaFree(d_utt_seg_off_);
  cudaFree(d_att_bias_);
  cudaFree(d_pooled_); cudaFree(d_fc_partial_); cudaFree(d_seg_emb_); cudaFree(d_feats_); cudaFree(d_out_);
  if (h_stage_) cudaFreeHost(h_stage_);
  for (auto e : events_) cudaEventDestroy(e);
}

int Model::new_tensor(int stage, int C) {
  ActTensor t; t.stage = stage; t.C = C;
  tensors_.push_back(t);
  return static_cast<int>(tensors_.size()) - 1;
}

std::string Model::next_name(std::map<std::string, int>& counts, const std::string& prefix, const std::string& base) {
  int k = counts[base]++;
  return prefix + (k == 0 ? base : base + "_" + std::to_string(k));
}

void Model::add_var(const std::string& name, std::vector<int64_t> shape) { vars_.push_back({name, std::move(shape)}); }

int Model::build() {
  if (cfg_.feat_dim <= 0 || cfg_.embed_dim <= 0) { set_last_error("bad feat_dim/embed_dim"); return 1; }
  switch (cfg_.family) {
    case SVX_FAMILY_TDNN: build_tdnn(); break;
    case SVX_FAMILY_RES2NET: build_res2net(); break;
    case SVX_FAMILY_DPN: build_dpn(); break;
    default: set_last_error("unknown model family"); return 1;
  }
  return 0;
}

void Model::build_tdnn() {
  std::map<std::string, int> root;
  n_stages_ = 1; gap_ = 3; stage_W_ = {1}; stage_Wp_ = {1};
  const int F = cfg_.feat_dim;
  int cur = new_tensor(0, round_up(F, 8));
  { Op op; op.kind = OP_PACK_INPUT; op.out = {cur, 0}; ops_.push_back(op); }
  int cin = F;
  for (int l = 0; l < cfg_.tdnn_layers; ++l) {
    const int f = cfg_.tdnn_filters[l], k = cfg_.tdnn_kernels[l], d = cfg_.tdnn_dilations[l];
    Op op; op.kind = OP_CONV;
    ConvDesc& c = op.conv;
    c.kernel_name = next_name(root, "", "conv2d") + "/kernel";            // tdnn_model.py:25-27
    add_var(c.kernel_name, {k, 1, cin, f});
    c.bn_name = next_name(root, "", "batch_normalization");               // tdnn_model.py:29
    add_var(c.bn_name + "/moving_mean", {f}); add_var(c.bn_name + "/moving_variance", {f});
    c.in = {cur, 0}; c.cin = cin; c.kh = k; c.kw = 1; c.dil = d; c.ph = d * (k - 1) / 2; c.pw = 0; c.cout = f;
    c.pre_relu = 1;                                                       // conv → ReLU → BN (tdnn_model.py:25-29)
    cur = new_tensor(0, f);
    c.out = {cur, 0};
    ops_.push_back(op);
    cin = f;
  }
  pool_tensor_ = cur; pool_C_ = cin; flat_dim_ = 2 * cin;
  tail_bn1_ = next_name(root, "", "batch_normalization");
  add_var(tail_bn1_ + "/moving_mean", {flat_dim_}); add_var(tail_bn1_ + "/moving_variance", {flat_dim_});
  add_var("dense/kernel", {flat_dim_, cfg_.embed_dim});
  tail_bn2_ = next_name(root, "", "batch_normalization");
  add_var(tail_bn2_ + "/moving_mean", {cfg_.embed_dim}); add_var(tail_bn2_ + "/moving_variance", {cfg_.embed_dim});
}

void Model::build_res2net() {
  std::map<std::string, int> root;
  const int F = cfg_.feat_dim, S = cfg_.split;
  // stage s = resolution after the s-th stride-2 layer
  n_stages_ = 1; gap_ = 1; stage_W_ = {F};
  for (int li = 0; li < 4; ++li)
    if (cfg_.block_strides[li] == 2) { stage_W_.push_back(ceil_half(stage_W_.back())); ++n_stages_; }
  for (int w : stage_W_) stage_Wp_.push_back(w + 1);
  int stage = 0;
  // stem (res2net_model.py:192-203)
  const int c0 = cfg_.num_filters[0];
  int cur = new_tensor(0, round_up(c0, 8));
  {
    Op op; op.kind = OP_STEM; op.out = {cur, 0}; op.C = c0;
    op.conv.kernel_name = next_name(root, "", "conv2d") + "/kernel";
    add_var(op.conv.kernel_name, {3, 3, 1, c0});
    op.bn_name = next_name(root, "", "batch_normalization");
    add_var(op.bn_name + "/moving_mean", {c0}); add_var(op.bn_name + "/moving_variance", {c0});
    ops_.push_back(op);
  }
  int cin = c0;
  for (int li = 0; li < 4; ++li) {
    const int filt = cfg_.num_filters[li], w = cfg_.width[li], cout = filt * 4, mid = S * w;
    const int st = cfg_.block_strides[li];
    const int in_stage = stage;
    if (st == 2) ++stage;
    const int xa = new_tensor(stage, cout), xb = new_tensor(stage, cout), sc = new_tensor(stage, cout);
    // stride-1 blocks keep the splits of the 1x1 output (x_i) and the running sums (x_{i+1} + o_i) PLANAR, one dense
    // [pixels, w] tensor each: a 3x3 then reads whole DRAM lines instead of a w-channel slice of every S*w-channel row
    // (measured 3.7-4.8x read amplification on the interleaved layout).  y (the concat, conv3's input) stays interleaved.
    std::vector<int> ms(S, -1), zs(S, -1);
    static const bool no_ypad0 = dbg_env("SVX_NO_YPAD") != nullptr;   // debug switch (see wp below)
    static const bool no_ppad = dbg_env("SVX_NO_PPAD") != nullptr;    // debug switch: planar tensors unpadded
    const int wpp = (!no_ypad0 && !no_ppad && w % 16 != 0) ? round_up(w, 16) : w;   // planar split tensors: padded like the concat slices, pad WRITTEN
    for (int i = 0; i + 1 < S; ++i) ms[i] = new_tensor(stage, wpp);
    for (int i = 1; i + 1 < S; ++i) zs[i] = new_tensor(stage, wpp);
    // The concat y keeps one slice per split.  When a slice (w channels) is not a whole number of 32-byte sectors, the slices are
    // padded to wp channels and the direct epilogue WRITES the pad (zeros): a 48-byte slice of a 192-byte pixel row is otherwise
    // stored as partial sectors that L2 read-fills from DRAM, which bounded the narrow 3x3 convs (profiles/r01_knockout_direct_stores.txt).
    // conv3 then reads S*wp channels; the weight rows of the pad positions are zero.
    static const bool no_ypad = dbg_env("SVX_NO_YPAD") != nullptr;   // debug switch
    const int wp = (!no_ypad && w % 16 != 0) ? round_up(w, 16) : w;
    // First stage, stride 1: the projection shortcut of the first block is folded into its conv3 as extra K (ConvDesc::fold_*).
    // The block input — the stem's output — then lives in the channels behind the concat slices of y (y is 64 channels wider),
    // written there by the stem and read from there by conv1: no shortcut launch, no shortcut tensor, no residual read.
    static const bool no_fold = dbg_env("SVX_NO_FOLD") != nullptr;   // debug switch
    const bool fold = li == 0 && st == 1 && !no_fold && cin % 8 == 0 && cin <= 64 && ops_.size() == 1 && ops_[0].kind == OP_STEM;
    const int y_all = new_tensor(stage, S * wp);
    // the first block's own concat, block input behind the slices (the other blocks keep the dense pitch).  64 spare channel slots
    // (pitch 384 B for 320 B of data): the dense form (pitch 320 B, SVX_FOLD_PAD=32) reads 0.3 GB less in conv3 but measured
    // 0.06-0.13 ms slower per step in a three-way A/B on one box (rows that straddle 128-byte lines)
    static const int fold_pad = dbg_env("SVX_FOLD_PAD") ? atoi(dbg_env("SVX_FOLD_PAD")) : 64;   // debug switch
    const int y_first = fold ? new_tensor(stage, S * wp + std::max(fold_pad, round_up(cin, 16))) : y_all;
    if (fold) { ops_[0].out = {y_first, S * wp}; cur = y_first; }
    // Stride-2 stages: the same fold with the block input's EVEN pixels (what a 1x1 stride-2 conv reads) copied behind the concat
    // slices by a small kernel (OP_SUBSAMPLE) — a quarter of the input instead of a shortcut launch that writes a full-width
    // tensor which conv3 then reads back as its residual.
    static const bool no_fold2 = dbg_env("SVX_NO_FOLD2") != nullptr;   // debug switch
    const bool fold2 = st == 2 && !no_fold && !no_fold2 && cin % 64 == 0 && (S * wp) % 64 == 0 && wp == w;
    const int y_first2 = fold2 ? new_tensor(stage, S * wp + cin) : -1;
    const int mp = st == 2 ? new_tensor(in_stage, mid) : -1;
    int out_t = xa;
    for (int b = 0; b < cfg_.block_sizes[li]; ++b) {
      const bool first = b == 0;
      const int bstride = first ? st : 1;
      int shortcut = cur;
      const bool folded = first && fold, folded2 = first && fold2;
      const int y = folded ? y_first : folded2 ? y_first2 : y_all;
      const int in_off = folded ? S * wp : 0;                    // where the block input sits in its tensor
      std::string fold_kernel, fold_bn;
      if (first) {   // projection shortcut: 1x1 conv stride s + BN (res2net_model.py:85-87,119-127)
        Op op; op.kind = OP_CONV; ConvDesc& c = op.conv;
        c.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(c.kernel_name, {1, 1, cin, cout});
        c.bn_name = next_name(root, "", "batch_normalization");
        add_var(c.bn_name + "/moving_mean", {cout}); add_var(c.bn_name + "/moving_variance", {cout});
        c.in = {cur, 0}; c.cin = cin; c.stride = bstride; c.cout = cout; c.out = {sc, 0};
        if (folded) { fold_kernel = c.kernel_name; fold_bn = c.bn_name; shortcut = -1; }
        else if (folded2) {
          fold_kernel = c.kernel_name; fold_bn = c.bn_name; shortcut = -1;
          Op sub; sub.kind = OP_SUBSAMPLE; sub.in = {cur, 0}; sub.out = {y, S * wp}; sub.C = cin;
          ops_.push_back(sub);
        } else { ops_.push_back(op); shortcut = sc; }
      }
      {   // conv1 1x1 + BN + ReLU (res2net_model.py:89-91)
        Op op; op.kind = OP_CONV; ConvDesc& c = op.conv;
        c.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(c.kernel_name, {1, 1, cin, mid});
        c.bn_name = next_name(root, "", "batch_normalization");
        add_var(c.bn_name + "/moving_mean", {mid}); add_var(c.bn_name + "/moving_variance", {mid});
        c.in = {cur, in_off}; c.cin = cin; c.cout = mid; c.post_relu = 1;
        if (bstride == 1) {   // x_0..x_{S-2} to their planar tensors, the last split passes straight into the concat (:74-75)
          c.out = {y, 0}; c.split_w = w; c.split_store = wpp;
          for (int i = 0; i + 1 < S; ++i) c.split_out.push_back({ms[i], 0});
          c.split_out.push_back({y, (S - 1) * wp});
        } else { c.out = {mp, 0}; }
        ops_.push_back(op);
      }
      {   // hierarchical 3x3 (res2net_model.py:26-78)
        std::map<std::string, int> inner;
        const std::string hscope = next_name(root, "", "conv2d");
        add_var(hscope + "/kernel", {3, 3, w, w * (S - 1)});
        for (int i = 0; i < S - 1; ++i) {
          Op op; op.kind = OP_CONV; ConvDesc& c = op.conv;
          c.kernel_name = hscope + "/kernel"; c.kernel_out_off = i * w;
          c.bn_name = next_name(inner, hscope + "/", "batch_normalization");
          add_var(c.bn_name + "/moving_mean", {w}); add_var(c.bn_name + "/moving_variance", {w});
          c.kh = c.kw = 3; c.stride = bstride; c.ph = c.pw = 1; c.cin = w; c.cout = w; c.post_relu = 1;
          c.out = {y, i * wp};
          if (bstride == 1) {
            c.out_store = wp; c.out2_store = wpp;
            c.in = {i == 0 ? ms[0] : zs[i], 0};
            if (i < S - 2) { c.out2 = {zs[i + 1], 0}; c.add2 = {ms[i + 1], 0}; }       // x_{i+1} + o_i (:65-66)
            if (S == 4 && wpp == 32 && wp == 32) c.chain_pos = i;                      // candidates for the fused chain (res2_chain.cu)
          } else {
            c.in = {mp, i * w};                                                        // no cross-split add when strided
          }
          ops_.push_back(op);
        }
        if (bstride == 2) {   // last split: avg_pool 3x3/2 over the padded tensor (:76-77)
          Op op; op.kind = OP_AVGPOOL; op.in = {mp, (S - 1) * w}; op.out = {y, (S - 1) * wp}; op.C = w;
          ops_.push_back(op);
        }
      }
      {   // conv3 1x1 + BN + shortcut + ReLU (res2net_model.py:98-101)
        Op op; op.kind = OP_CONV; ConvDesc& c = op.conv;
        c.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(c.kernel_name, {1, 1, mid, cout});
        c.bn_name = next_name(root, "", "batch_normalization");
        add_var(c.bn_name + "/moving_mean", {cout}); add_var(c.bn_name + "/moving_variance", {cout});
        c.in = {y, 0}; c.cin = S * wp; c.cout = cout; c.post_relu = 1; c.out = {out_t, 0};
        if (folded || folded2) { c.fold_kernel_name = fold_kernel; c.fold_bn_name = fold_bn; c.fold_cin = cin; c.fold_off = S * wp; c.cin = S * wp + cin; }
        else c.res = {shortcut, 0};
        if (wp != w) { c.in_gw = w; c.in_gwp = wp; }
        ops_.push_back(op);
      }
      cur = out_t; out_t = (out_t == xa) ? xb : xa; cin = cout;
    }
  }
  pool_tensor_ = cur; pool_C_ = cin; flat_dim_ = stage_W_.back() * 2 * cin;
  if (cfg_.att_pool) {   // att_stats_pool scope with two bias-free 1x1 convs (models.py:273-303)
    const int A = cfg_.att_dim > 0 ? cfg_.att_dim : 128;
    const std::string ascope = next_name(root, "", "att_stats_pool");
    has_att_ = true;
    const int t1 = new_tensor(stage, A), t2 = new_tensor(stage, cin);
    att_a_.kernel_name = ascope + "/conv2d/kernel"; add_var(att_a_.kernel_name, {1, 1, 3 * cin, A});
    att_a_.in = {cur, 0}; att_a_.cin = cin; att_a_.cout = A; att_a_.out = {t1, 0};      // x part of concat(x, mean, std); the rest is a bias
    att_b_.kernel_name = ascope + "/conv2d_1/kernel"; add_var(att_b_.kernel_name, {1, 1, A, cin});
    att_b_.in = {t1, 0}; att_b_.cin = A; att_b_.cout = cin; att_b_.out = {t2, 0};
  }
  tail_bn1_ = next_name(root, "", "batch_normalization");
  add_var(tail_bn1_ + "/moving_mean", {flat_dim_}); add_var(tail_bn1_ + "/moving_variance", {flat_dim_});
  add_var("dense/kernel", {flat_dim_, cfg_.embed_dim});
  tail_bn2_ = next_name(root, "", "batch_normalization");
  add_var(tail_bn2_ + "/moving_mean", {cfg_.embed_dim}); add_var(tail_bn2_ + "/moving_variance", {cfg_.embed_dim});
}

void Model::build_dpn() {
  std::map<std::string, int> root;
  const int F = cfg_.feat_dim;
  n_stages_ = 4; gap_ = 1; stage_W_ = {F, ceil_half(F), ceil_half(ceil_half(F)), ceil_half(ceil_half(ceil_half(F)))};
  for (int w : stage_W_) stage_Wp_.push_back(w + 1);
  const int c0 = cfg_.init_features;
  int cur = new_tensor(0, round_up(c0, 8));   // raw block input of the next stage (S0 first, then X[s])
  {
    Op op; op.kind = OP_STEM; op.out = {cur, 0}; op.C = c0;                 // dpn_model.py:32-37
    op.conv.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(op.conv.kernel_name, {3, 3, 1, c0});
    op.bn_name = next_name(root, "", "batch_normalization");
    add_var(op.bn_name + "/moving_mean", {c0}); add_var(op.bn_name + "/moving_variance", {c0});
    ops_.push_back(op);
  }
  int c_in = c0;
  auto bn_vars = [&](const std::string& n, int C) { add_var(n + "/moving_mean", {C}); add_var(n + "/moving_variance", {C}); };
  for (int s = 0; s < 4; ++s) {
    const int bw = cfg_.bw << s, inc = cfg_.inc_sec[s], r = cfg_.k_r * bw / cfg_.bw, card = cfg_.cardinality;
    const int c_out = bw + (cfg_.k_sec[s] + 2) * inc;
    const int st = s == 0 ? 1 : 2;
    const int in_stage = s == 0 ? 0 : s - 1;
    const int X = new_tensor(s, c_out);
    const int XA = new_tensor(s, c_out), Y1 = new_tensor(s, r), Y2 = new_tensor(s, r);
    const int XP = new_tensor(s, round_up(c_in, 8)), XAP = new_tensor(in_stage, round_up(c_in, 8)), Y1P = new_tensor(in_stage, r);
    int c = c_in;
    for (int b = 0; b < cfg_.k_sec[s]; ++b) {
      const bool first = b == 0;
      const int bstride = first ? st : 1;
      int a_in, y1;
      if (first) {
        {   // projection: BN → ReLU → 1x1 conv stride s (dpn_model.py:75), split into res0 | dense0 (:76-79)
          Op e; e.kind = OP_BN_RELU; e.bn_name = next_name(root, "", "batch_normalization"); bn_vars(e.bn_name, c_in);
          e.in = {cur, 0}; e.out = {XP, 0}; e.C = c_in; e.stride = bstride; ops_.push_back(e);
          Op op; op.kind = OP_CONV; ConvDesc& cv = op.conv;
          cv.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(cv.kernel_name, {1, 1, c_in, bw + 2 * inc});
          cv.in = {XP, 0}; cv.cin = c_in; cv.cout = bw + 2 * inc; cv.out = {X, 0};
          ops_.push_back(op);
        }
        {   // conv_a input: BN → ReLU of the un-projected block input at its own resolution (dpn_model.py:49,81)
          Op e; e.kind = OP_BN_RELU; e.bn_name = next_name(root, "", "batch_normalization"); bn_vars(e.bn_name, c_in);
          e.in = {cur, 0}; e.out = {XAP, 0}; e.C = c_in; e.stride = 1; ops_.push_back(e);
        }
        a_in = XAP; y1 = Y1P;
      } else {
        Op e; e.kind = OP_BN_RELU; e.bn_name = next_name(root, "", "batch_normalization"); bn_vars(e.bn_name, c);
        e.in = {X, 0}; e.out = {XA, 0}; e.C = c; e.stride = 1; ops_.push_back(e);
        a_in = XA; y1 = Y1;
      }
      const int ca = first ? c_in : c;
      // conv_a 1x1 → r, fused with the next BN → ReLU (dpn_model.py:49-50)
      const std::string ka = next_name(root, "", "conv2d") + "/kernel"; add_var(ka, {1, 1, ca, r});
      const std::string bnb = next_name(root, "", "batch_normalization"); bn_vars(bnb, r);
      {
        Op op; op.kind = OP_CONV; ConvDesc& cv = op.conv;
        cv.kernel_name = ka; cv.bn_name = bnb; cv.in = {a_in, 0}; cv.cin = ca; cv.cout = r; cv.post_relu = 1; cv.out = {y1, 0};
        ops_.push_back(op);
      }
      // conv_b 3x3 grouped, stride s, TF SAME, fused with the next BN → ReLU (dpn_model.py:50,53)
      const std::string kb = next_name(root, "", "conv2d") + "/kernel"; add_var(kb, {3, 3, r / card, r});
      const std::string bnc = next_name(root, "", "batch_normalization"); bn_vars(bnc, r);
      {
        Op op; op.kind = OP_CONV; ConvDesc& cv = op.conv;
        cv.kernel_name = kb; cv.bn_name = bnc; cv.in = {y1, 0}; cv.cin = r; cv.cout = r; cv.groups = card;
        cv.kh = cv.kw = 3; cv.stride = bstride; cv.post_relu = 1; cv.out = {Y2, 0};
        cv.ph = 1;   // rows: the segment layout makes in_row = stride*out_row + r - 1 hold for even and odd lengths
        cv.pw = bstride == 1 ? 1 : ((stage_W_[in_stage] % 2 == 0) ? 0 : 1);   // TF SAME, stride 2: pad_left = 0 for even extents [ext]
        ops_.push_back(op);
      }
      // conv_c 1x1 → bw+inc, raw: first bw channels add into the residual path, the rest append to the dense path (:83-87)
      {
        Op op; op.kind = OP_CONV; ConvDesc& cv = op.conv;
        cv.kernel_name = next_name(root, "", "conv2d") + "/kernel"; add_var(cv.kernel_name, {1, 1, r, bw + inc});
        const int c_now = first ? bw + 2 * inc : c;
        cv.in = {Y2, 0}; cv.cin = r; cv.cout = bw + inc; cv.n_split = bw;
        cv.out = {X, 0}; cv.res = {X, 0}; cv.outb = {X, c_now};
        ops_.push_back(op);
        c = c_now + inc;
      }
    }
    cur = X; c_in = c_out;
  }
  pool_tensor_ = cur; pool_C_ = c_in; flat_dim_ = stage_W_.back() * 2 * c_in;
  pool_bn_ = next_name(root, "", "batch_normalization"); bn_vars(pool_bn_, c_in);   // dpn_model.py:24-29
  tail_bn1_ = next_name(root, "", "batch_normalization"); bn_vars(tail_bn1_, flat_dim_);
  add_var("dense/kernel", {flat_dim_, cfg_.embed_dim});
  tail_bn2_ = next_name(root, "", "batch_normalization"); bn_vars(tail_bn2_, cfg_.embed_dim);
}

// ------------------------------------------------------------------------------------------------ weights
int Model::set_tensor(const char* name, const float* data, int ndim, const int64_t* shape) {
  for (const auto& v : vars_) {
    if (v.name != name) continue;
    if (static_cast<int>(v.shape.size()) != ndim) { set_last_error(std::string("rank mismatch for ") + name); return 1; }
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
      if (v.shape[i] != shape[i]) {
        set_last_error(std::string("shape mismatch for ") + name + ": expected dim " + std::to_string(i) + " = " +
                       std::to_string(v.shape[i]) + ", got " + std::to_string(shape[i]));
        return 1;
      }
      n *= static_cast<size_t>(shape[i]);
    }
    HostTensor& t = host_[name];
    t.shape.assign(shape, shape + ndim);
    t.data.assign(data, data + n);
    t.set = true;
    return 0;
  }
  set_last_error(std::string("unknown tensor name: ") + name);
  return 1;
}

int Model::fold_bn(const std::string& bn, int C, bool four_d, std::vector<float>& scale, std::vector<float>& shift) {
  const HostTensor& m = host_[bn + "/moving_mean"];
  const HostTensor& v = host_[bn + "/moving_variance"];
  const float eps = four_d ? kBnEps4d : kBnEps2d;
  scale.resize(C); shift.resize(C);
  for (int i = 0; i < C; ++i) {
    const float s = 1.0f / std::sqrt(v.data[i] + eps);
    scale[i] = s;
    shift[i] = -m.data[i] * s;
  }
  return 0;
}

template <typename T> static T host_cvt(float v);
template <> __half host_cvt<__half>(float v) { return __float2half_rn(std::min(std::max(v, -65504.f), 65504.f)); }
template <> __nv_bfloat16 host_cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// TMA issues one request per box row at a fixed ~1.75 ns whatever its width (profiles/r01_microbench_tma_rate.txt),
// so the widest swizzle span that the slice fills at all is the cheapest; padded K only costs idle MMA cycles.
static int pick_kbox(int cin) {
  if (dbg_env("SVX_KBOX_OLD")) {   // debug switch: exact-fit boxes
    if (cin % 64 == 0) return 64;
    if (cin % 32 == 0) return 32;
  }
  if (cin <= 16) return 16;
  if (cin <= 32) return 32;
  return 64;
}

// Output tiles are staged as 64-channel boxes for TMA stores, and a store box cannot be clipped at a tile boundary
// (only at the end of the destination slice), so a conv is either one n-tile (≤ 128 channels) or several n-tiles whose
// width is a multiple of 64.  Returns 0 when no such tiling exists (the caller then keeps direct stores).
static int pick_ntile(int cout, int max_multi) {
  const int n = round_up(cout, 16);
  if (n <= 128) return n;
  for (int t = max_multi; t >= 64; t -= 64)
    if (cout % t == 0) return t;
  return 0;
}

int Model::upload_conv_weights(ConvDesc& c) {
  const HostTensor& k = host_[c.kernel_name];   // [kh,kw,cin_g,cout_total]
  const int taps = c.kh * c.kw;
  const int cin_main = c.fold_cin > 0 ? c.fold_off : c.cin;               // channels of the conv's own input (a folded shortcut's follow)
  const int cin_tf = c.in_gw > 0 ? (cin_main / c.in_gwp) * c.in_gw : cin_main;   // input channels of the TF variable (padded concat: fewer than cin)
  const int cin_g = cin_tf / c.groups, cout_g = c.cout / c.groups;
  const int cout_total = static_cast<int>(k.shape[3]);
  const bool grouped = c.groups > 1;
  int grp_ntile = 0, grp_cstep = 0;
  if (grouped) {
    if (cin_g != cout_g || 32 % cin_g != 0) { set_last_error("unsupported grouped conv shape"); return 1; }
    grp_ntile = 32; grp_cstep = 32;
    c.kbox = 32; c.nkc = 1; c.kpad = 32; c.n_tile = 32;
    c.no_staged = true;   // 32-channel n-tiles: the 2-D kernel's 64-channel store boxes would spill into the neighbouring tile
  } else {
    c.kbox = pick_kbox(c.cin);
    c.nkc = (c.cin + c.kbox - 1) / c.kbox;
    c.kpad = c.nkc * c.kbox;
    // outputs are staged in 64-channel shared-memory boxes for TMA stores: ≤ 2 boxes per tile, 1 when an add2 tile
    // and an out2 tile are live as well
    c.n_tile = pick_ntile(c.cout, c.add2.id >= 0 ? 64 : 128);
    if (c.n_tile == 0) { c.n_tile = 128; c.no_staged = true; }
  }
  c.n_gemm = c.cout;
  if (c.split_w > 0) {   // box size and padded group width of the planar splits (see plan_flat)
    const int sw = c.split_w;
    c.split_box = (sw % 64 == 0) ? 64 : (sw % 32 == 0 || sw < 32) ? 32 : 64;
    c.split_wp = round_up(sw, c.split_box);
    c.n_gemm = static_cast<int>(c.split_out.size()) * c.split_wp;
    c.n_tile = std::min(256, c.n_gemm);
  }
  c.n_pad = round_up(c.n_gemm, c.n_tile);
  c.n_tiles = c.n_pad / c.n_tile;
  auto gemm_row = [&](int n) { return c.split_w > 0 ? (n / c.split_w) * c.split_wp + n % c.split_w : n; };
  const size_t K = static_cast<size_t>(taps) * c.kpad;
  std::vector<float> w(static_cast<size_t>(c.n_pad) * K, 0.f);
  for (int n = 0; n < c.cout; ++n) {
    const int g = n / cout_g;
    const int abase = grouped ? (n / grp_ntile) * grp_cstep : 0;
    for (int t = 0; t < taps; ++t)
      for (int ci = 0; ci < cin_g; ++ci) {
        int j = g * cin_g + ci - abase;           // position inside the k-box row
        if (c.in_gw > 0) j = (j / c.in_gw) * c.in_gwp + j % c.in_gw;
        w[static_cast<size_t>(gemm_row(n)) * K + static_cast<size_t>(t) * c.kpad + j] =
            k.data[(static_cast<size_t>(t) * cin_g + ci) * cout_total + c.kernel_out_off + n];
      }
  }
  std::vector<float> fold_shift;
  if (c.fold_cin > 0) {   // the folded shortcut: rows fold_off .. fold_off + fold_cin, scaled by s_shortcut / s_conv per output channel
    if (taps != 1 || grouped || c.split_w > 0 || c.bn_name.empty()) { set_last_error("folded shortcut on an unsupported conv"); return 1; }
    const HostTensor& kf = host_[c.fold_kernel_name];   // [1,1,fold_cin,cout]
    std::vector<float> s3, b3, ss, bs;
    fold_bn(c.bn_name, c.cout, true, s3, b3);
    fold_bn(c.fold_bn_name, c.cout, true, ss, bs);
    for (int n = 0; n < c.cout; ++n)
      for (int ci = 0; ci < c.fold_cin; ++ci)
        w[static_cast<size_t>(n) * K + c.fold_off + ci] = kf.data[static_cast<size_t>(ci) * c.cout + n] * (ss[n] / s3[n]);
    fold_shift = bs;
  }
  void* d = nullptr;
  SVX_CUDA(cudaMalloc(&d, w.size() * 2));
  owned_.push_back(d);
  if (is_bf16_) {
    std::vector<__nv_bfloat16> h(w.size());
    for (size_t i = 0; i < w.size(); ++i) h[i] = host_cvt<__nv_bfloat16>(w[i]);
    SVX_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  } else {
    std::vector<__half> h(w.size());
    for (size_t i = 0; i < w.size(); ++i) h[i] = host_cvt<__half>(w[i]);
    SVX_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  c.d_wgt = d;
  c.sp.grp_ntile = grp_ntile; c.sp.grp_cstep = grp_cstep;
  if (!c.bn_name.empty()) {
    std::vector<float> sc, sh;
    fold_bn(c.bn_name, c.cout, true, sc, sh);
    for (size_t n = 0; n < fold_shift.size(); ++n) sh[n] += fold_shift[n];
    if (c.split_w > 0) {
      std::vector<float> sc2(c.n_pad, 0.f), sh2(c.n_pad, 0.f);
      for (int n = 0; n < c.cout; ++n) { sc2[gemm_row(n)] = sc[n]; sh2[gemm_row(n)] = sh[n]; }
      sc.swap(sc2); sh.swap(sh2);
    }
    sc.resize(c.n_pad, 0.f); sh.resize(c.n_pad, 0.f);
    float* ds = nullptr; float* dh = nullptr;
    SVX_CUDA(cudaMalloc(&ds, c.n_pad * 4)); owned_.push_back(ds);
    SVX_CUDA(cudaMalloc(&dh, c.n_pad * 4)); owned_.push_back(dh);
    SVX_CUDA(cudaMemcpy(ds, sc.data(), c.n_pad * 4, cudaMemcpyHostToDevice));
    SVX_CUDA(cudaMemcpy(dh, sh.data(), c.n_pad * 4, cudaMemcpyHostToDevice));
    c.d_scale = ds; c.d_shift = dh;
  }
  return 0;
}

static int upload_floats(std::vector<void*>& owned, const std::vector<float>& v, float** out) {
  float* d = nullptr;
  if (cudaMalloc(&d, std::max<size_t>(v.size(), 1) * 4) != cudaSuccess) { set_last_error("cudaMalloc failed"); return 1; }
  owned.push_back(d);
  if (cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_last_error("cudaMemcpy failed"); return 1; }
  *out = d;
  return 0;
}

int Model::finalize() {
  SVX_CUDA(cudaSetDevice(device_));
  for (const auto& v : vars_)
    if (!host_.count(v.name) || !host_[v.name].set) { set_last_error("tensor not set: " + v.name); return 1; }
  SVX_CUDA(conv_umma_init());
  SVX_CUDA(conv_flat_init());
  SVX_CUDA(res2_chain_init());
  SVX_CUDA(conv_pair_init());
  for (Op& op : ops_) {
    if (op.kind == OP_CONV) {
      if (upload_conv_weights(op.conv)) return 1;
    } else if (op.kind == OP_STEM) {
      const HostTensor& k = host_[op.conv.kernel_name];   // [3,3,1,C] → [9][C]
      if (upload_floats(owned_, k.data, &op.d_w9)) return 1;
      std::vector<float> sc, sh;
      fold_bn(op.bn_name, op.C, true, sc, sh);
      if (upload_floats(owned_, sc, &op.d_scale) || upload_floats(owned_, sh, &op.d_shift)) return 1;
    } else if (op.kind == OP_BN_RELU) {
      std::vector<float> sc, sh;
      fold_bn(op.bn_name, op.C, true, sc, sh);
      const int cp = round_up(op.C, 8);
      sc.resize(cp, 0.f); sh.resize(cp, 0.f);
      if (upload_floats(owned_, sc, &op.d_scale) || upload_floats(owned_, sh, &op.d_shift)) return 1;
    }
  }
  if (has_att_) {
    if (upload_conv_weights(att_a_) || upload_conv_weights(att_b_)) return 1;
    const HostTensor& k = host_[att_a_.kernel_name];   // [1,1,3C,A]: rows C..3C multiply the tiled [mean | std]
    const int C = pool_C_, A = att_a_.cout;
    std::vector<float> wms(k.data.begin() + static_cast<size_t>(C) * A, k.data.begin() + static_cast<size_t>(3 * C) * A);
    if (upload_floats(owned_, wms, &d_att_wms_)) return 1;
  }
  if (!pool_bn_.empty()) {
    std::vector<float> sc, sh;
    fold_bn(pool_bn_, pool_C_, true, sc, sh);
    if (upload_floats(owned_, sc, &d_pool_scale_) || upload_floats(owned_, sh, &d_pool_shift_)) return 1;
  }
  {   // BN → dense → BN folded into one fp32 affine map (models.py:306-309 between two 2-D batch norms)
    std::vector<float> s1, b1, s2, b2;
    fold_bn(tail_bn1_, flat_dim_, false, s1, b1);
    fold_bn(tail_bn2_, cfg_.embed_dim, false, s2, b2);
    const HostTensor& W = host_["dense/kernel"];
    const int D = flat_dim_, E = cfg_.embed_dim;
    std::vector<float> Wf(static_cast<size_t>(D) * E);
    std::vector<double> bias(E, 0.0);
    for (int d = 0; d < D; ++d)
      for (int e = 0; e < E; ++e) {
        const float w = W.data[static_cast<size_t>(d) * E + e];
        Wf[static_cast<size_t>(d) * E + e] = s1[d] * w * s2[e];
        bias[e] += static_cast<double>(b1[d]) * w;
      }
    std::vector<float> bf(E);
    for (int e = 0; e < E; ++e) bf[e] = static_cast<float>(bias[e] * s2[e] + b2[e]);
    if (upload_floats(owned_, Wf, &d_Wf_) || upload_floats(owned_, bf, &d_bias_)) return 1;
  }
  host_.clear();
  finalized_ = true;
  return 0;
}

int Model::set_option(const char* key, int value) {
  if (!strcmp(key, "force_simple")) { force_simple_ = value; return 0; }
  if (!strcmp(key, "time_convs")) { time_convs_ = value; return 0; }
  if (!strcmp(key, "no_flat")) { force_no_flat_ = value; return 0; }
  if (!strcmp(key, "no_chain")) { no_chain_ = value; return 0; }
  if (!strcmp(key, "no_pair")) { no_pair_ = value; return 0; }
  if (!strcmp(key, "no_pair_s2")) { no_pair_s2_ = value; return 0; }
  set_last_error(std::string("unknown option: ") + key);
  return 1;
}

// ------------------------------------------------------------------------------------------------ workspace
int Model::plan_conv(ConvDesc& c) {
  const ActTensor& tin = tensors_[c.in.id];
  const ActTensor& tout = tensors_[c.out.id];
  const int in_W = stage_W_[tin.stage], out_W = stage_W_[tout.stage];
  const int in_Wp = stage_Wp_[tin.stage], out_Wp = stage_Wp_[tout.stage];
  const int in_rows = rows_cap_[tin.stage];
  auto tptr = [&](const TensorRef& r) -> void* { return r.id >= 0 ? tensors_[r.id].ptr : nullptr; };
  auto tC = [&](const TensorRef& r) -> int { return r.id >= 0 ? tensors_[r.id].C : 0; };
  Epilogue e;
  memset(&e, 0, sizeof e);
  e.scale = c.d_scale; e.shift = c.d_shift; e.pre_relu = c.pre_relu; e.post_relu = c.post_relu; e.n_valid = c.n_gemm;
  if (c.split_w > 0) {
    e.n_splits = static_cast<int>(c.split_out.size()); e.split_wp = c.split_wp; e.split_w = c.split_w;
    for (int i = 0; i < e.n_splits; ++i) {
      e.split_ptr[i] = tensors_[c.split_out[i].id].ptr; e.split_C[i] = tensors_[c.split_out[i].id].C; e.split_coff[i] = c.split_out[i].coff;
    }
  }
  e.out = tptr(c.out); e.out_C = tC(c.out); e.out_coff = c.out.coff;
  e.res = tptr(c.res); e.res_C = tC(c.res); e.res_coff = c.res.coff;
  e.n_split = c.n_split < 0 ? c.cout : c.n_split;
  e.outb = tptr(c.outb); e.outb_C = tC(c.outb); e.outb_coff = c.outb.coff;
  e.out2 = tptr(c.out2); e.out2_C = tC(c.out2); e.out2_coff = c.out2.coff;
  e.add2 = tptr(c.add2); e.add2_C = tC(c.add2); e.add2_coff = c.add2.coff;
  e.seg_of_row = d_seg_of_row_[tout.stage];
  // ---- CUDA-core form
  SimpleConvParams& sp = c.sp;
  sp.in = tin.ptr; sp.in_C = tin.C; sp.in_coff = c.in.coff; sp.in_rows = in_rows; sp.in_W = in_W; sp.in_Wp = in_Wp;
  sp.wgt = c.d_wgt; sp.kpad = c.kpad; sp.cin_g = c.cin / c.groups; sp.cout_g = c.n_gemm / c.groups;
  sp.kh = c.kh; sp.kw = c.kw; sp.sh = c.stride; sp.sw = c.stride; sp.dh = c.dil; sp.dw = 1; sp.ph = c.ph; sp.pw = c.pw;
  sp.out_rows = 0; sp.out_W = out_W; sp.out_Wp = out_Wp; sp.epi = e;
  c.use_flat = false;
  if (plan_flat(c)) return 1;
  c.use_pair = false;
  if ((c.use_flat || c.stride == 2) && plan_pair(c)) return 1;
  // ---- tcgen05 form
  c.use_umma = false;
  const int taps = c.kh * c.kw;
  auto mult8 = [](int v) { return v % 8 == 0; };
  const bool grouped_ok = c.groups > 1 && c.n_tile == 32 && c.kbox == 32 && c.nkc == 1 && c.cin == c.cout && c.cin % 32 == 0;
  bool ok = c.split_w == 0 && (c.groups == 1 || grouped_ok) && taps <= kMaxTaps && (c.stride == 1 || (c.stride == 2 && c.dil == 1)) && mult8(tin.C) && mult8(c.in.coff) &&
            mult8(c.cout) && mult8(e.out_C) && mult8(e.out_coff) && mult8(e.n_split) &&
            (!e.res || (mult8(e.res_C) && mult8(e.res_coff))) && (!e.outb || (mult8(e.outb_C) && mult8(e.outb_coff))) &&
            (!e.out2 || (mult8(e.out2_C) && mult8(e.out2_coff) && mult8(e.add2_C) && mult8(e.add2_coff)));
  if (!ok) return 0;
  UmmaConvParams& up = c.up;
  memset(&up, 0, sizeof up);
  const size_t esz = 2;
  int w_box = 1;
  while (w_box < 128 && out_W % (w_box * 2) == 0) w_box *= 2;
  up.out_rows = 0; up.out_W = out_W; up.out_Wp = out_Wp; up.w_box = w_box; up.h_box = 128 / w_box; up.w_tiles = out_W / w_box;
  up.taps = taps; up.nkc = c.nkc; up.kbox = c.kbox; up.a_c_step = c.groups > 1 ? 32 : 0; up.n_tile = c.n_tile; up.n_tiles = c.n_tiles;
  up.aux_mode = e.res ? 1 : (e.out2 ? 2 : 0);
  up.aux_boxes = up.aux_mode ? (c.n_tile + 63) / 64 : 0;
  up.aux_width = up.aux_mode == 1 ? e.n_split : (up.aux_mode == 2 ? c.cout : 0);
  {   // debug switch: SVX_STAGED_MASK bit 1 plain, 2 split, 4 residual, 8 add2, 16 stride-2 (default all)
    const char* mk = dbg_env("SVX_STAGED_MASK");
    const int mask = mk ? atoi(mk) : 63;
    int kind = up.aux_mode == 1 ? 4 : up.aux_mode == 2 ? 8 : (e.n_split < c.cout ? 2 : (c.cout % 64 == 0 ? 32 : 1));
    if (c.stride == 2 && !(mask & 16)) kind = 0;
    up.store_mode = ((mask & kind) && !c.no_staged) ? 1 : 0;
  }
  const int sw_bytes = c.kbox * 2;
  up.layout_type = sw_bytes == 128 ? 2u : sw_bytes == 64 ? 4u : 6u;
  up.sbo = 8u * sw_bytes;
  up.idesc = ptx::make_idesc_f16(is_bf16_ ? 1u : 0u, 128u, static_cast<uint32_t>(c.n_tile));
  up.a_stage_bytes = 128u * sw_bytes;
  up.b_stage_bytes = static_cast<uint32_t>(round_up(c.n_tile * sw_bytes, 1024));
  if (!conv_umma_finish_params(up)) return 0;
  up.epi = e;
  uint8_t* base = static_cast<uint8_t*>(tin.ptr);
  for (int t = 0; t < taps; ++t) {
    const int r = t / c.kw, s = t % c.kw;
    if (c.stride == 1) {
      up.tap_map[t] = 0; up.tap_dh[t] = static_cast<int8_t>(r * c.dil - c.ph); up.tap_dw[t] = static_cast<int8_t>(s - c.pw);
    } else {
      const int a = r - c.ph, b = s - c.pw;
      const int p = ((a % 2) + 2) % 2, q = ((b % 2) + 2) % 2;
      up.tap_map[t] = static_cast<int8_t>(p * 2 + q);
      up.tap_dh[t] = static_cast<int8_t>((a - p) / 2); up.tap_dw[t] = static_cast<int8_t>((b - q) / 2);
    }
  }
  const uint32_t box[3] = {static_cast<uint32_t>(c.kbox), static_cast<uint32_t>(w_box), static_cast<uint32_t>(128 / w_box)};
  memset(&c.auxmap, 0, sizeof c.auxmap);
  memset(&c.omaps, 0, sizeof c.omaps);
  {   // TMA-store maps over the destination slices (64-channel boxes, clipped at the slice width)
    const uint32_t obox[3] = {64u, static_cast<uint32_t>(w_box), static_cast<uint32_t>(128 / w_box)};
    const uint64_t dims[3] = {static_cast<uint64_t>(e.n_split), static_cast<uint64_t>(out_W), static_cast<uint64_t>(rows_cap_[tout.stage])};
    const uint64_t str[2] = {static_cast<uint64_t>(tout.C) * esz, static_cast<uint64_t>(out_Wp) * tout.C * esz};
    if (encode_tmap(&c.omaps.m[0], is_bf16_, static_cast<uint8_t*>(tout.ptr) + static_cast<size_t>(c.out.coff) * esz, 3, dims, str, obox, 128))
      return 1;
    c.omaps.m[1] = c.omaps.m[0];
    if (c.out2.id >= 0) {
      const ActTensor& t2 = tensors_[c.out2.id];
      const uint64_t d2[3] = {static_cast<uint64_t>(c.cout), static_cast<uint64_t>(out_W), static_cast<uint64_t>(rows_cap_[t2.stage])};
      const uint64_t s2[2] = {static_cast<uint64_t>(t2.C) * esz, static_cast<uint64_t>(out_Wp) * t2.C * esz};
      if (encode_tmap(&c.omaps.m[1], is_bf16_, static_cast<uint8_t*>(t2.ptr) + static_cast<size_t>(c.out2.coff) * esz, 3, d2, s2, obox, 128))
        return 1;
    }
  }
  if (up.aux_mode) {   // residual / add2 tile: 64-channel SWIZZLE_128B boxes over the output-resolution tensor slice
    const TensorRef& ar = up.aux_mode == 1 ? c.res : c.add2;
    const ActTensor& ta = tensors_[ar.id];
    const uint64_t dims[3] = {static_cast<uint64_t>(up.aux_width), static_cast<uint64_t>(out_W), static_cast<uint64_t>(rows_cap_[ta.stage])};
    const uint64_t str[2] = {static_cast<uint64_t>(ta.C) * esz, static_cast<uint64_t>(out_Wp) * ta.C * esz};
    const uint32_t abox[3] = {64u, static_cast<uint32_t>(w_box), static_cast<uint32_t>(128 / w_box)};
    if (encode_tmap(&c.auxmap, is_bf16_, static_cast<uint8_t*>(ta.ptr) + static_cast<size_t>(ar.coff) * esz, 3, dims, str, abox, 128)) return 1;
  }
  if (c.stride == 1) {
    const uint64_t dims[3] = {static_cast<uint64_t>(c.cin), static_cast<uint64_t>(in_W), static_cast<uint64_t>(in_rows)};
    const uint64_t str[2] = {static_cast<uint64_t>(tin.C) * esz, static_cast<uint64_t>(in_Wp) * tin.C * esz};
    if (encode_tmap(&c.amaps.m[0], is_bf16_, base + static_cast<size_t>(c.in.coff) * esz, 3, dims, str, box, sw_bytes)) return 1;
    for (int i = 1; i < 4; ++i) c.amaps.m[i] = c.amaps.m[0];
  } else {
    for (int p = 0; p < 2; ++p)
      for (int q = 0; q < 2; ++q) {
        const int wv = (in_W - q + 1) / 2, rv = (in_rows - p + 1) / 2;
        if (wv <= 0 || rv <= 0) { c.amaps.m[p * 2 + q] = c.amaps.m[0]; continue; }
        const uint64_t dims[3] = {static_cast<uint64_t>(c.cin), static_cast<uint64_t>(wv), static_cast<uint64_t>(rv)};
        const uint64_t str[2] = {2ull * tin.C * esz, 2ull * in_Wp * tin.C * esz};
        uint8_t* b = base + (static_cast<size_t>(p) * in_Wp + q) * tin.C * esz + static_cast<size_t>(c.in.coff) * esz;
        if (encode_tmap(&c.amaps.m[p * 2 + q], is_bf16_, b, 3, dims, str, box, sw_bytes)) return 1;
      }
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(taps) * c.kpad, static_cast<uint64_t>(c.n_pad)};
    const uint64_t str[1] = {static_cast<uint64_t>(taps) * c.kpad * esz};
    const uint32_t bbox[2] = {static_cast<uint32_t>(c.kbox), static_cast<uint32_t>(c.n_tile)};
    if (encode_tmap(&c.bmap, is_bf16_, c.d_wgt, 2, dims, str, bbox, sw_bytes)) return 1;
  }
  c.use_umma = true;
  return 0;
}

// Flat plan (conv_flat.cu) for stride-1, ungrouped convs: tap shifts in the pixel sequence, tile shape and the
// shared-memory budget (A span ring, weights resident or ringed, epilogue slots).  Leaves use_flat = false when the
// layer does not qualify; the caller then plans the 2-D tiled kernel.
int Model::plan_flat(ConvDesc& c) {
  static const bool disabled = dbg_env("SVX_NO_FLAT") != nullptr;   // debug switch
  if (disabled) return 0;
  const ActTensor& tin = tensors_[c.in.id];
  const ActTensor& tout = tensors_[c.out.id];
  if (c.stride != 1 || tin.stage != tout.stage) return 0;
  // grouped convs (DPN, 32 groups): n-tiles of 32 output channels with block-diagonal weights over their own 32 input channels
  // (upload_conv_weights lays the weights out that way); each n-tile's CTA offsets its A loads by 32 channels
  const bool grouped = c.groups > 1;
  if (grouped && (c.n_tile != 32 || c.kbox != 32 || c.nkc != 1 || c.cin != c.cout || c.cin % 32 != 0)) return 0;
  const int taps = c.kh * c.kw;
  if (taps > kMaxTaps) return 0;
  const int Wp = stage_Wp_[tout.stage], W = stage_W_[tout.stage];
  if (c.kw > 1 && (Wp <= W || c.pw > 1 || c.kw - 1 - c.pw > 1)) return 0;   // one zero column covers |dw| <= 1 only
  auto mult8 = [](int v) { return v % 8 == 0; };
  auto tC = [&](const TensorRef& r) -> int { return r.id >= 0 ? tensors_[r.id].C : 0; };
  const int n_split = c.n_split < 0 ? c.cout : c.n_split;
  if (!(mult8(tin.C) && mult8(c.in.coff) && mult8(c.cout) && mult8(tout.C) && mult8(c.out.coff) && mult8(n_split))) return 0;
  if (c.res.id >= 0 && !(mult8(tC(c.res)) && mult8(c.res.coff))) return 0;
  if (c.outb.id >= 0 && !(mult8(tC(c.outb)) && mult8(c.outb.coff))) return 0;
  if (c.out2.id >= 0 && !(mult8(tC(c.out2)) && mult8(c.out2.coff) && mult8(tC(c.add2)) && mult8(c.add2.coff))) return 0;
  if (c.res.id >= 0 && c.out2.id >= 0) return 0;
  if (c.pre_relu && (c.post_relu || c.res.id >= 0 || c.out2.id >= 0)) return 0;
  if (c.out2.id >= 0 && (!c.post_relu || n_split != c.cout)) return 0;
  const bool split = c.split_w > 0;
  if (split && (c.res.id >= 0 || c.out2.id >= 0 || c.outb.id >= 0 || c.split_out.size() > 8)) return 0;
  const int N = c.n_gemm;
  const int aux_mode = c.res.id >= 0 ? 1 : (c.out2.id >= 0 ? 2 : 0);

  FlatConvParams fp;
  memset(&fp, 0, sizeof fp);
  fp.taps = taps; fp.nkc = c.nkc; fp.kbox = c.kbox; fp.kpad = c.kpad;
  fp.a_c_step = grouped ? 32 : 0;
  int halo = 0;
  for (int t = 0; t < taps; ++t) {
    const int r = t / c.kw, s = t % c.kw;
    fp.tap_shift[t] = (r * c.dil - c.ph) * Wp + (s - c.pw);
    halo = std::max(halo, std::abs(fp.tap_shift[t]));
  }
  fp.halo = halo;
  const uint32_t row_bytes = static_cast<uint32_t>(c.kbox) * 2u;
  fp.layout_type = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  fp.sbo = 8u * row_bytes;
  fp.scale = c.d_scale; fp.shift = c.d_shift; fp.n_valid = N;
  fp.pix_valid = d_pix_valid_[tout.stage];
  fp.aux_mode = aux_mode; fp.pre_relu = c.pre_relu; fp.post_relu = c.post_relu;
  fp.n_res = aux_mode == 1 ? n_split : 0;
  fp.grp_mask = 0; fp.grp_w = 1;
  const int split_store = (split && c.split_store > c.split_w && c.split_store <= c.split_wp) ? c.split_store : c.split_w;   // pad channels stored as zeros
  if (split && (c.split_wp & (c.split_wp - 1)) == 0 && split_store < c.split_wp) { fp.grp_mask = c.split_wp - 1; fp.grp_w = split_store; }

  // Search over tile shapes with a small cost model (cycles per 128 output pixels, all n-tiles): tensor pipe (bounded by
  // operand reads from shared memory when N is small), TMA row requests (~5.6 cycles per box row per SM with every SM
  // pulling from L2, profiles/r01_microbench_tma_rate.txt), load latency over the bytes the rings keep in flight, and
  // the epilogue's issue slots.  The cheapest shape that fits the 227 KB of shared memory wins.
  const long long budget = 227 * 1024 - 1024 - 3072;
  const int n16 = round_up(N, 16);
  std::vector<int> cands;
  if (n16 <= 256) cands.push_back(n16);
  if (split) {
    for (int k = 4; k >= 1; --k)
      if (k * c.split_wp < n16 && k * c.split_wp <= 256 && N % (k * c.split_wp) == 0) cands.push_back(k * c.split_wp);
  } else {
    for (int t : {256, 192, 128, 64, 32})
      if (t < n16 && N % t == 0) cands.push_back(t);
  }
  static const int env_maxmt = dbg_env("SVX_FLAT_MAXMT") ? atoi(dbg_env("SVX_FLAT_MAXMT")) : 4;          // debug switches
  static const bool plan_log = dbg_env("SVX_PLAN_LOG") != nullptr;
  // Multicast clusters cm x cn (conv_flat.cu): cn CTAs share a span (A slices multicast across the n-tiles), cm CTAs share an
  // n-tile (streamed weight slices multicast across spans).  The cost model below charges the L2 row requests once per cluster;
  // SVX_MC_CN / SVX_MC_CM force a shape.
  // MEASURED (profiles/r02_multicast_clusters.txt): no shape is faster than single CTAs on any layer — A multicast x2 / x4 costs
  // +6 % / +13 % on the deep 1x1 convs, weight multicast the same — so the bytes a CTA RECEIVES (~45-50 GB/s per SM with all
  // SMs pulling), not the L2 reads, bound those layers; clusters are therefore opt-in (SVX_MC=1, debug build only).
  static const bool no_mc = dbg_env("SVX_MC") == nullptr;
  static const int force_cn = dbg_env("SVX_MC_CN") ? atoi(dbg_env("SVX_MC_CN")) : 0;
  static const int force_cm = dbg_env("SVX_MC_CM") ? atoi(dbg_env("SVX_MC_CM")) : 0;
  static const double mc_penalty = dbg_env("SVX_MC_PENALTY") ? atof(dbg_env("SVX_MC_PENALTY")) : 1.05;   // lockstep of the cluster's CTAs
  const int ksteps = c.kbox / 16;
  // direct epilogue (global accesses from the epilogue threads instead of slots + TMA) for narrow single-destination tiles
  static const bool no_direct = dbg_env("SVX_NO_DIRECT") != nullptr;   // debug switch
  const bool direct_ok = !no_direct && !split && n_split == c.cout && c.outb.id < 0;
  // aux mode 2 over dense planar tensors: the add2 / out2 tiles are contiguous runs -> 1-D bulk copies instead of 128 rows each
  // (measured: no gain — 14 354 vs 14 458 emb/s with it on the stage-3 3x3 convs, and none on stages 1-2 against 2-D TMA tiles —
  // so it is opt-in: SVX_LIN=1)
  static const bool use_lin = dbg_env("SVX_LIN") != nullptr;   // debug switch
  bool lin_ok = false;
  if (use_lin && aux_mode == 2 && !split) {
    const ActTensor& ta = tensors_[c.add2.id];
    const ActTensor& t2 = tensors_[c.out2.id];
    lin_ok = ta.C == c.cout && t2.C == c.cout && c.add2.coff == 0 && c.out2.coff == 0 && (c.cout * 2) % 16 == 0;
  }
  bool found = false;
  double best = 1e30;
  // debug build: SVX_FORCE_PLAN="taps,cin,cout,n_tile,mt,bres[;...]" pins the tile shape of one layer class (-1 = free)
  int f_ntile = -1, f_mt = -1, f_bres = -1;
  if (const char* fpl = dbg_env("SVX_FORCE_PLAN")) {
    const char* q = fpl;
    while (q && *q) {
      int a[6];
      if (sscanf(q, "%d,%d,%d,%d,%d,%d", &a[0], &a[1], &a[2], &a[3], &a[4], &a[5]) == 6 && a[0] == taps && a[1] == c.cin && a[2] == c.cout) {
        f_ntile = a[3]; f_mt = a[4]; f_bres = a[5];
      }
      q = strchr(q, ';');
      if (q) ++q;
    }
  }
  for (int n_tile : cands) {
    if (f_ntile > 0 && n_tile != f_ntile) continue;
    const int n_tiles = (N + n_tile - 1) / n_tile;
    for (int box_ch : {64, 32}) {
      if (split && box_ch != c.split_box) continue;
      if (!split && box_ch == 64 && n_tile <= 32) continue;                              // narrow tiles: 32-channel boxes halve the slot size
      if (n_tiles > 1 && n_tile % box_ch != 0) continue;
      if (!split && box_ch == 32 && n_tile > 32 && n_tile % 64 == 0) continue;   // 64-channel boxes are never worse there
      // routing of channels >= n_split
      if (!split && n_split < c.cout && n_split % box_ch != 0) continue;       // a staging box has exactly one destination
      if ((N + box_ch - 1) / box_ch > 32) continue;                              // routing table size
      if (grouped && (n_tile != 32 || box_ch != 32)) continue;
      const int n_parts = 1;
      const int part_cols = n_tile;
      const int boxes = (part_cols + box_ch - 1) / box_ch;
      const uint32_t box_bytes = 128u * box_ch * 2u;
      const bool direct = direct_ok && n_tile <= 64 && (n_tiles == 1 || grouped);   // several n-tiles would re-read A per 64 channels
      static const bool no_hybrid = dbg_env("SVX_NO_HYBRID") != nullptr;   // debug switch
      // out2 through a slot + TMA, aux and out1 on the LSU — where the pixel runs are not 32-byte multiples (24 channels); with
      // sector-aligned runs (48 channels) the 256-bit stores of the fully direct form are faster (118 vs 134 us in stage 2)
      const bool hybrid = direct && aux_mode == 2 && !no_hybrid && (c.cout * 2) % 32 != 0;
      const bool lin = lin_ok && !direct && n_tiles == 1;
      const uint32_t slot_bytes = hybrid ? boxes * box_bytes : direct ? 0u : lin ? boxes * box_bytes + static_cast<uint32_t>(round_up(128 * c.cout * 2, 1024))
                                                    : boxes * box_bytes * (aux_mode == 2 ? 2u : 1u);
      const int b_rows_cta = n_tile;
      const uint32_t b_item = static_cast<uint32_t>(round_up(b_rows_cta * static_cast<int>(row_bytes), 1024));
      const int items = taps * c.nkc;
      const long long b_total = static_cast<long long>(items) * b_item;
      for (int mt : {4, 2, 1}) {
        if (mt > env_maxmt) continue;
        if (mt * n_tile > 256) continue;
        for (int cn : {1, 2, 4, 8})
        for (int cm : {1, 2, 4}) {
        if (cn * cm > 8) continue;
        if (cn * cm > 1 && (no_mc || grouped || direct)) continue;
        if (force_cn > 0 && n_tiles % force_cn == 0 && !grouped && !direct && cn != force_cn) continue;
        if (force_cm > 0 && !grouped && !direct && cm != force_cm) continue;
        if (n_tiles % cn != 0) continue;
        if (n_tile % (8 * cm) != 0) continue;                       // weight slices are whole 8-row swizzle atoms
        const int a_rows_min = mt * 128 + 2 * halo;
        const int a_boxes = (a_rows_min + 255) / 256;
        const int a_box_rows = round_up((a_rows_min + a_boxes - 1) / a_boxes, 8 * cn);   // ... and so are the A slices
        if (a_box_rows > 256) continue;
        const uint32_t a_stage = static_cast<uint32_t>(round_up(a_boxes * a_box_rows * static_cast<int>(row_bytes), 1024));
        // clusters never straddle a GPC: fewer CTAs fit than there are SMs (conv_flat_max_clusters), which the cost pays for
        const int cs = cn * cm;
        const double sm_frac = cs == 1 ? 1.0 : std::min(1.0, static_cast<double>(conv_flat_max_clusters(cs)) * cs / 148.0);
        for (int b_res : {1, 0}) {
          if (cm > 1 && b_res) continue;                            // resident weights are loaded once: nothing to share
          static const long long bres_max = dbg_env("SVX_BRES_MAX") ? atoll(dbg_env("SVX_BRES_MAX")) : 64 * 1024;   // tuning knobs; resident weights above 64 KB starve the A ring (measured +1.2 % against 96 KB)
          static const double lat_cyc = dbg_env("SVX_LAT_CYC") ? atof(dbg_env("SVX_LAT_CYC")) : 3000.0;
          static const double slot_scale = dbg_env("SVX_SLOT_SCALE") ? atof(dbg_env("SVX_SLOT_SCALE")) : 1.0;
          if (b_res && b_total > bres_max && f_bres != 1) continue;
          if (f_bres >= 0 && b_res != f_bres) continue;
          if (f_mt > 0 && mt != f_mt) continue;
          if (!b_res && items < 2) continue;
          static const int slots0 = dbg_env("SVX_SLOTS0") ? atoi(dbg_env("SVX_SLOTS0")) : 2;
          int a_stages = 2, b_stages = b_res ? 0 : 2, slots = slots0;               // slots: per warpgroup (2: convert j+1 while j is stored)
          long long left = budget - (b_res ? b_total : 2LL * b_item) - 2LL * a_stage - 2LL * slots * slot_bytes;
          if (left < 0) { slots = 1; left += 2LL * slot_bytes; }
          if (left < 0) continue;
          {   // a 1x1 with resident weights gains more from a 2nd slot per warpgroup than from mt > 1 — when mt = 1 gets that slot
            const uint32_t a1 = static_cast<uint32_t>(round_up(128 * static_cast<int>(row_bytes), 1024));
            const bool mt1_two_slots = budget - b_total - 2LL * a1 - 4LL * slot_bytes >= 0;
            if (slots == 1 && b_res && taps == 1 && mt > 1 && mt1_two_slots) continue;
          }
          // ---- cost per 128 output pixels (cycles)
          const double halo_ovh = 1.0 + 2.0 * halo / (mt * 128.0);
          const double a_rows_cta = static_cast<double>(n_tiles) * c.nkc * 128.0 * halo_ovh;       // rows every CTA receives
          const double b_rows_rcv = b_res ? 0.0 : static_cast<double>(n_tiles) * items * b_rows_cta / mt;
          const double a_rows = a_rows_cta / cn, b_rows = b_rows_rcv / cm;                           // rows requested from L2
          const int aux_boxes = (aux_mode && !direct && !lin) ? boxes : 0;
          const double aux_rows = static_cast<double>(n_tiles) * aux_boxes * 128.0;
          const double st_rows = hybrid ? static_cast<double>(n_tiles) * boxes * 128.0 : direct ? 0.0 : static_cast<double>(n_tiles) * boxes * 128.0 * ((aux_mode == 2 && !lin) ? 2.0 : 1.0);
          const double t_req = (a_rows + b_rows + aux_rows) * 5.6;
          const double t_st = st_rows * 4.6;
          const double load_bytes = (a_rows_cta + b_rows_rcv) * row_bytes;
          const double mma_cyc = std::max(n_tile / 2.0, (4096.0 + n_tile * 32.0) / 128.0) + 6.0;
          const double t_mma = static_cast<double>(n_tiles) * items * ksteps * mma_cyc;
          const double t_epi = static_cast<double>(n_tiles) * (n_tile / 16.0) * 150.0;
          const double t_fixed = std::max(std::max(t_req, t_st), std::max(t_mma, t_epi));
          // latency terms shrink with ring depth: loads in flight against ~1.5 us, and slots held ~4 us with an aux tile
          // (prefetch → conversion → store read), ~1.5 us without (profiles/r01_trace_flat_*.txt); 2*slots are in flight
          auto t_lat = [&]() {
            return load_bytes * lat_cyc / static_cast<double>(static_cast<long long>(a_stages) * a_stage + static_cast<long long>(b_stages) * b_item);
          };
          auto t_slot = [&]() { return (direct && !hybrid) ? 0.0 : static_cast<double>(n_tiles) * n_parts * (aux_mode ? 7600.0 : 2900.0) * slot_scale / (2.0 * slots); };
          for (;;) {   // grow whichever ring currently bounds the tile, while it fits
            const double tl = t_lat(), ts = t_slot();
            if (std::max(tl, ts) <= t_fixed) break;
            bool grew = false;
            if (ts >= tl) {
              if (slots < 4 && left >= 2LL * slot_bytes) { ++slots; left -= 2LL * slot_bytes; grew = true; }
            }
            if (!grew) {
              const bool want_b = !b_res && b_stages < 8 && b_stages < items && static_cast<long long>(b_stages) * b_item <= static_cast<long long>(a_stages) * a_stage;
              if (want_b && left >= b_item) { ++b_stages; left -= b_item; grew = true; }
              else if (a_stages < 8 && left >= a_stage) { ++a_stages; left -= a_stage; grew = true; }
              else if (!b_res && b_stages < 8 && b_stages < items && left >= b_item) { ++b_stages; left -= b_item; grew = true; }
            }
            if (!grew && ts < tl && slots < 4 && left >= 2LL * slot_bytes && ts > t_fixed) { ++slots; left -= 2LL * slot_bytes; grew = true; }
            if (!grew) break;
          }
          // the model is optimistic about overlap: when shared memory is left over, take a second slot per warpgroup (convert
          // j+1 while j is being stored), a third A stage, and a third slot for aux tiles
          static const int a_min = dbg_env("SVX_A_MIN") ? atoi(dbg_env("SVX_A_MIN")) : 3;
          while (a_stages < a_min && a_min > 3 && left >= a_stage) { ++a_stages; left -= a_stage; }
          if (slots < 2 && left >= 2LL * slot_bytes) { ++slots; left -= 2LL * slot_bytes; }
          if (a_stages < 3 && left >= a_stage) { ++a_stages; left -= a_stage; }
          if (aux_mode && slots < 3 && left >= 2LL * slot_bytes) { ++slots; left -= 2LL * slot_bytes; }
          if (!b_res && b_stages < 4 && b_stages < items && left >= b_item) { ++b_stages; left -= b_item; }
          const double score = std::max(t_fixed, std::max(t_lat(), t_slot())) / sm_frac * (cs > 1 ? mc_penalty : 1.0) + 0.01 * n_tiles - 0.001 * mt +
                               (b_res ? 0.0 : 0.005);
          const double t_latv = t_lat(), t_slotv = t_slot();
          if (plan_log)
            fprintf(stderr, "  cand n_tile %d box %d mt %d cl %dx%d bres %d a_st %d b_st %d slots %d: req %.0f st %.0f lat %.0f mma %.0f epi %.0f slot %.0f -> %.0f\n", n_tile,
                    box_ch, mt, cm, cn, b_res, a_stages, b_stages, slots, t_req, t_st, t_latv, t_mma, t_epi, t_slotv, score);
          if (score < best) {
            best = score;
            fp.mt = mt; fp.a_box_rows = a_box_rows; fp.a_boxes = a_boxes; fp.n_tile = n_tile; fp.n_tiles = n_tiles;
            fp.a_stages = a_stages; fp.b_stages = b_stages; fp.a_stage_bytes = a_stage; fp.b_item_bytes = b_item;
            fp.b_resident = b_res; fp.box_ch = box_ch; fp.boxes = boxes; fp.slots = slots; fp.slot_bytes = slot_bytes;
            fp.n_parts = n_parts; fp.part_cols = part_cols;
            fp.cn = cn; fp.cm = cm; fp.a_slice_rows = a_box_rows / cn; fp.b_slice_rows = n_tile / cm;
            fp.direct = hybrid ? 2 : direct ? 1 : 0; fp.lin = lin ? 1 : 0;
            found = true;
          }
        }
        }
      }
    }
  }
  if (found && plan_log)
    fprintf(stderr, "plan %dx%d cin %d cout %d aux %d: n_tile %d x%d box %d mt %d cluster %dx%d bres %d a_st %d b_st %d slots %d direct %d lin %d score %.0f\n", c.kh,
            c.kw, c.cin, c.cout, aux_mode, fp.n_tile, fp.n_tiles, fp.box_ch, fp.mt, fp.cm, fp.cn, fp.b_resident, fp.a_stages, fp.b_stages, fp.slots, fp.direct,
            fp.lin, best);
  if (!found) return 0;
  // TMEM buffers: the epilogue of span s releases its accumulators only after its last sub-tile, so with 2 buffers the MMAs of
  // span s+2 wait for it; 4 buffers (when they fit in 512 columns) take that wait off the critical path
  static const bool no_tb4 = dbg_env("SVX_NO_TMEM4") != nullptr;   // debug switch
  const int bufs = (!no_tb4 && 4 * fp.mt * fp.n_tile <= 512) ? 4 : 2;
  fp.tmem_bufs = bufs; fp.tmem_bufs_log2 = bufs == 4 ? 2 : 1;
  uint32_t tc = 32;
  while (tc < static_cast<uint32_t>(bufs) * fp.mt * fp.n_tile) tc *= 2;
  if (tc > 512) return 0;
  fp.tmem_cols = tc;
  fp.b_rows = fp.n_tile;
  fp.idesc = ptx::make_idesc_f16(is_bf16_ ? 1u : 0u, 128u, static_cast<uint32_t>(fp.n_tile));

  // tensor maps over the flat pixel sequence
  const size_t esz = 2;
  const uint64_t P_cap = static_cast<uint64_t>(rows_cap_[tout.stage]) * Wp;
  const int sw_box = fp.box_ch * 2;   // 128 or 64
  FlatMaps& fm = c.fmaps;
  memset(&fm, 0, sizeof fm);
  {
    // a padded planar tensor (pad channels hold zeros, the weight rows of the pad are zero) is read whole: dense rows, full L2 promotion
    const int a_width = (c.in.coff == 0 && tin.C > c.cin && tin.C <= c.kpad && c.groups == 1) ? tin.C : c.cin;
    const uint64_t dims[2] = {static_cast<uint64_t>(a_width), P_cap};
    const uint64_t str[1] = {static_cast<uint64_t>(tin.C) * esz};
    const uint32_t box[2] = {static_cast<uint32_t>(c.kbox), static_cast<uint32_t>(fp.a_slice_rows)};      // one load per (box, cluster column)
    static const int env_promo = dbg_env("SVX_L2_PROMO") ? atoi(dbg_env("SVX_L2_PROMO")) : -1;   // debug switch (see promo_for below)
    const int pitch = tin.C * 2, off = c.in.coff * 2, wb = c.cin * 2;
    const int promo = env_promo >= 0 ? env_promo : (a_width == tin.C) ? 128 : (pitch % 128 == 0 && off % 128 == 0 && wb % 128 == 0) ? 128
                                                 : (pitch % 64 == 0 && off % 64 == 0 && wb % 64 == 0) ? 64 : 0;
    if (encode_tmap(&fm.a, is_bf16_, static_cast<uint8_t*>(tin.ptr) + static_cast<size_t>(c.in.coff) * esz, 2, dims, str, box, row_bytes, promo)) return 1;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(taps) * c.kpad, static_cast<uint64_t>(c.n_pad)};
    const uint64_t str[1] = {static_cast<uint64_t>(taps) * c.kpad * esz};
    const uint32_t box[2] = {static_cast<uint32_t>(c.kbox), static_cast<uint32_t>(fp.b_resident ? fp.b_rows : fp.b_slice_rows)};
    if (encode_tmap(&fm.b, is_bf16_, c.d_wgt, 2, dims, str, box, row_bytes)) return 1;
  }
  // L2 promotion widens every TMA request to the promotion size; on a narrow channel slice of a wider row that
  // multiplies the DRAM traffic (measured 3.7-4.8x on the 24-channel Res2Net splits), so promote only what the
  // slice geometry fills.
  auto promo_for = [&](const ActTensor& t, int coff, int width) -> int {
    static const int env_promo = dbg_env("SVX_L2_PROMO") ? atoi(dbg_env("SVX_L2_PROMO")) : -1;   // debug switch
    if (env_promo >= 0) return env_promo;
    const int pitch = t.C * 2, off = coff * 2, wb = width * 2;
    if (width == t.C) return 128;                                  // dense tensor: every fetched byte is used
    if (pitch % 128 == 0 && off % 128 == 0 && wb % 128 == 0) return 128;
    if (pitch % 64 == 0 && off % 64 == 0 && wb % 64 == 0) return 64;
    return 0;
  };
  auto slice_map = [&](CUtensorMap* m, const ActTensor& t, int coff, int width) -> int {
    const uint64_t dims[2] = {static_cast<uint64_t>(width), P_cap};
    const uint64_t str[1] = {static_cast<uint64_t>(t.C) * esz};
    const uint32_t box[2] = {static_cast<uint32_t>(fp.box_ch), 128u};
    return encode_tmap(m, is_bf16_, static_cast<uint8_t*>(t.ptr) + static_cast<size_t>(coff) * esz, 2, dims, str, box, sw_box,
                       promo_for(t, coff, width));
  };
  // destinations: one output map per planar split, or map 0 = primary slice and map 1 = the channels past n_split
  memset(fp.route_map, 0xff, sizeof fp.route_map);
  const int n_boxes_total = (N + fp.box_ch - 1) / fp.box_ch;
  if (split) {
    for (size_t i = 0; i < c.split_out.size(); ++i) {
      const ActTensor& ts = tensors_[c.split_out[i].id];
      if (ts.stage != tout.stage) return 0;
      if (slice_map(&fm.o[i], ts, c.split_out[i].coff, std::min(split_store, ts.C - c.split_out[i].coff))) return 1;
    }
    for (int gb = 0; gb < n_boxes_total; ++gb) {
      const int ch = gb * fp.box_ch, sidx = ch / c.split_wp, j0 = ch - sidx * c.split_wp;
      if (j0 < split_store) { fp.route_map[gb] = static_cast<uint8_t>(sidx); fp.route_c[gb] = j0; }
    }
  } else {
    if (slice_map(&fm.o[0], tout, c.out.coff, n_split)) return 1;
    if (n_split < c.cout) {
      const ActTensor& tb = tensors_[c.outb.id];
      if (tb.stage != tout.stage) return 0;
      if (slice_map(&fm.o[1], tb, c.outb.coff, c.cout - n_split)) return 1;
    }
    for (int gb = 0; gb < n_boxes_total; ++gb) {
      const int ch = gb * fp.box_ch;
      if (ch < n_split) { fp.route_map[gb] = 0; fp.route_c[gb] = ch; }
      else { fp.route_map[gb] = 1; fp.route_c[gb] = ch - n_split; }
    }
  }
  if (aux_mode == 1) {
    const ActTensor& tr = tensors_[c.res.id];
    if (tr.stage != tout.stage) return 0;
    if (slice_map(&fm.aux, tr, c.res.coff, n_split)) return 1;
  } else if (aux_mode == 2) {
    const ActTensor& ta = tensors_[c.add2.id];
    const ActTensor& t2 = tensors_[c.out2.id];
    if (ta.stage != tout.stage || t2.stage != tout.stage) return 0;
    if (slice_map(&fm.aux, ta, c.add2.coff, c.cout)) return 1;
    if (slice_map(&fm.o2, t2, c.out2.coff, (fp.direct == 2 && c.out2_store > c.cout && c.out2_store <= t2.C - c.out2.coff) ? c.out2_store : c.cout)) return 1;
  }
  fp.P_cap = static_cast<long long>(P_cap);
  fp.n_store2 = (fp.direct && c.out2_store > c.cout && fp.n_tiles == 1 && c.out2_store <= fp.n_tile) ? c.out2_store : N;
  fp.n_store = (fp.direct && c.out_store > c.cout && fp.n_tiles == 1 && c.out_store <= fp.n_tile) ? c.out_store : N;
  if (fp.direct || fp.lin) {
    fp.d_out = static_cast<uint8_t*>(tout.ptr) + static_cast<size_t>(c.out.coff) * esz; fp.d_out_pitch = static_cast<uint32_t>(tout.C * esz);
    if (aux_mode == 1) {
      const ActTensor& tr = tensors_[c.res.id];
      fp.d_aux = static_cast<const uint8_t*>(tr.ptr) + static_cast<size_t>(c.res.coff) * esz; fp.d_aux_pitch = static_cast<uint32_t>(tr.C * esz);
    } else if (aux_mode == 2) {
      const ActTensor& ta = tensors_[c.add2.id];
      const ActTensor& t2 = tensors_[c.out2.id];
      fp.d_aux = static_cast<const uint8_t*>(ta.ptr) + static_cast<size_t>(c.add2.coff) * esz; fp.d_aux_pitch = static_cast<uint32_t>(ta.C * esz);
      fp.d_out2 = static_cast<uint8_t*>(t2.ptr) + static_cast<size_t>(c.out2.coff) * esz; fp.d_out2_pitch = static_cast<uint32_t>(t2.C * esz);
    }
  }
  if (conv_flat_smem_bytes(fp) > 227 * 1024) return 0;
  c.fp = fp;
  c.use_flat = true;
  return 0;
}

static unsigned long long* flat_dbg_words();

// CTA-pair kernel (conv_pair.cu): 1x1 stride-1 convs with K >= 384 (a pair of SMs per M = 256 x N <= 256 tile halves the weight bytes
// every SM receives) and the 3x3 stride-1 convs whose half of the weights fits in shared memory beside the pixel ring (resident
// weights instead of a 2-4 deep weight ring per span).  Leaves use_pair = false when the shape does not qualify (the flat plan runs).
int Model::plan_pair(ConvDesc& c) {
  static const bool disabled = dbg_env("SVX_NO_PAIR") != nullptr;   // debug switch
  // 1x1: measured faster from K = 192 (Res2Net-50 stage 2: 192->256 + residual 348 -> 317 us, 256->192 275 -> 196 us); at K = 128 (stage 1,
  // HBM-bound at 4.1 M pixels) the flat kernel's TMA-store epilogue wins (495 vs 725 us)
  static const int min_k = dbg_env("SVX_PAIR_MIN_K") ? atoi(dbg_env("SVX_PAIR_MIN_K")) : 192;
  static const bool no_3x3 = dbg_env("SVX_NO_PAIR3") != nullptr;
  // 3x3: measured faster from 96 channels (97/93/86 -> 78/80/56 us per hierarchical chain of stage 3); at 48 channels the two
  // epilogue passes of aux mode 2 on three 16-channel groups cost more than the weight ring saves (120 -> 195 us)
  static const int min_c3 = dbg_env("SVX_PAIR3_MIN_C") ? atoi(dbg_env("SVX_PAIR3_MIN_C")) : 96;
  if (disabled) return 0;
  const ActTensor& tin = tensors_[c.in.id];
  const ActTensor& tout = tensors_[c.out.id];
  const int taps = c.kh * c.kw;
  static const bool no_s2 = dbg_env("SVX_NO_PAIR_S2") != nullptr;
  const bool s2 = c.stride == 2;          // Res2Net down-sampling convs: 2-D tile mode over four parity-phase views of the input
  if (c.groups != 1 || taps > kMaxTaps || c.dil != 1) return 0;
  if (!s2 && (c.stride != 1 || tin.stage != tout.stage)) return 0;
  if (c.kbox != 64 || c.kpad % 64 != 0) return 0;
  const int Wp = stage_Wp_[tout.stage], W = stage_W_[tout.stage];
  if (s2) {
    // explicit (k-1)/2 padding on both sides (models.py:121-134): in = 2 out + d - pad; the segment layout gives row offsets off[s] = 2 off[s+1]
    if (no_s2 || cfg_.family != SVX_FAMILY_RES2NET || tin.stage + 1 != tout.stage) return 0;
    if (!((c.kh == 1 && c.kw == 1 && c.ph == 0 && c.pw == 0) || (c.kh == 3 && c.kw == 3 && c.ph == 1 && c.pw == 1))) return 0;
    if (c.res.id >= 0 || c.out2.id >= 0 || c.split_w > 0 || c.pre_relu) return 0;
    if (Wp <= W || stage_Wp_[tin.stage] != 2 * W + 1) return 0;
  } else if (taps == 1) {
    // (K padded up to 192 from fewer channels stays on the flat kernel — except conv3 with the shortcut folded in, K = 160 on
    // 200 x 80 pixels: 490 us here, 576-672 us on every flat plan, profiles/r02_experiments.md)
    if (c.kpad < min_k || (c.cin < min_k && c.fold_cin == 0)) return 0;
  } else {
    if (no_3x3 || c.cin < min_c3) return 0;
    if (c.kw > 1 && (Wp <= W || c.pw > 1 || c.kw - 1 - c.pw > 1)) return 0;   // one zero column covers |dw| <= 1 only
  }
  const int aux_mode = c.res.id >= 0 ? 1 : (c.out2.id >= 0 ? 2 : 0);
  if (c.res.id >= 0 && c.out2.id >= 0) return 0;
  if (aux_mode == 2 && (c.add2.id < 0 || !c.post_relu || c.pre_relu)) return 0;
  const int N = c.n_gemm;
  if (N % 16 != 0 || N > 1024 || N < 32) return 0;
  int n_tile = N % 256 == 0 ? 256 : N % 192 == 0 ? 192 : N % 128 == 0 ? 128 : 0;
  if (n_tile == 0 && N <= 256) n_tile = N;            // one tile: any multiple of 16
  if (n_tile == 0 || c.n_pad < N || (n_tile / 2) % 8 != 0) return 0;
  const int n_split = c.n_split < 0 ? c.cout : c.n_split;
  const bool split = c.split_w > 0;
  if (split && aux_mode) return 0;
  PairConvParams pp;
  memset(&pp, 0, sizeof pp);
  memset(pp.route, 0xff, sizeof pp.route);
  const size_t esz = 2;
  int n_dst = 0;
  auto add_dst = [&](const TensorRef& r) -> int {
    const ActTensor& t = tensors_[r.id];
    if (t.stage != tout.stage || (t.C * esz) % 32 != 0 || n_dst >= 10) return -1;
    pp.dst_base[n_dst] = static_cast<uint8_t*>(t.ptr);
    pp.dst_pitch[n_dst] = static_cast<uint32_t>(t.C * esz);
    return n_dst++;
  };
  if (split) {
    if (c.split_wp % 16 != 0 || c.split_w % 16 != 0 || c.split_out.size() > 8) return 0;
    const int split_store = (c.split_store > c.split_w && c.split_store <= c.split_wp) ? c.split_store : c.split_w;
    if (split_store % 16 != 0) return 0;
    std::vector<int> di;
    for (const TensorRef& r : c.split_out) {
      if (r.coff % 16 != 0) return 0;
      const int d = add_dst(r);
      if (d < 0) return 0;
      di.push_back(d);
    }
    for (int g = 0; g < N / 16; ++g) {
      const int ch = g * 16, s = ch / c.split_wp, j = ch - s * c.split_wp;
      if (s >= static_cast<int>(di.size()) || j >= split_store) continue;
      pp.route[g] = static_cast<uint8_t>(di[s]);
      pp.goff[g] = static_cast<uint16_t>((c.split_out[s].coff + j) * esz);
    }
  } else {
    if (N != c.cout || n_split % 16 != 0 || c.out.coff % 16 != 0) return 0;
    const int d0 = add_dst(c.out);
    if (d0 < 0) return 0;
    int d1 = -1;
    if (n_split < c.cout) {
      if (c.outb.id < 0 || c.outb.coff % 16 != 0) return 0;
      d1 = add_dst(c.outb);
      if (d1 < 0) return 0;
    }
    for (int g = 0; g < N / 16; ++g) {
      const int ch = g * 16;
      if (ch < n_split) { pp.route[g] = static_cast<uint8_t>(d0); pp.goff[g] = static_cast<uint16_t>((c.out.coff + ch) * esz); }
      else { pp.route[g] = static_cast<uint8_t>(d1); pp.goff[g] = static_cast<uint16_t>((c.outb.coff + ch - n_split) * esz); }
    }
  }
  if (aux_mode == 1) {
    const ActTensor& tr = tensors_[c.res.id];
    if (tr.stage != tout.stage || (tr.C * esz) % 32 != 0 || c.res.coff % 16 != 0 || n_split % 16 != 0) return 0;
    pp.res = static_cast<const uint8_t*>(tr.ptr) + static_cast<size_t>(c.res.coff) * esz;
    pp.res_pitch = static_cast<uint32_t>(tr.C * esz);
    pp.n_res = n_split;
  } else if (aux_mode == 2) {
    const ActTensor& ta = tensors_[c.add2.id];
    const ActTensor& t2 = tensors_[c.out2.id];
    if (ta.stage != tout.stage || t2.stage != tout.stage || (ta.C * esz) % 32 != 0 || (t2.C * esz) % 32 != 0 || c.add2.coff % 16 != 0 || c.out2.coff % 16 != 0) return 0;
    if (n_split != c.cout || N != c.cout || N > 128 || ta.C - c.add2.coff < N || t2.C - c.out2.coff < N) return 0;
    pp.res = static_cast<const uint8_t*>(ta.ptr) + static_cast<size_t>(c.add2.coff) * esz;
    pp.res_pitch = static_cast<uint32_t>(ta.C * esz);
    pp.n_res = N;
    pp.out2 = static_cast<uint8_t*>(t2.ptr) + static_cast<size_t>(c.out2.coff) * esz;
    pp.out2_pitch = static_cast<uint32_t>(t2.C * esz);
  }
  pp.aux_mode = aux_mode;
  const int nkb = c.kpad / 64;
  const int ks_last = std::min(4, std::max(1, (c.cin - (nkb - 1) * 64 + 15) / 16));   // real channels of the last K box (the rest is zero-filled / zero weights)
  pp.n_tile = n_tile; pp.n_tiles = N / n_tile; pp.n_gemm = N; pp.n_valid = N;
  pp.b_item_bytes = static_cast<uint32_t>(round_up((n_tile / 2) * 128, 1024));
  pp.b_resident = taps > 1 ? 1 : 0;
  pp.out_wp = Wp;
  int n_items = 0;
  auto add_item = [&](int row_shift, int wcol, int ks) {
    pp.item_off16[n_items] = static_cast<uint16_t>(row_shift * 8); pp.item_wcol[n_items] = static_cast<uint16_t>(wcol); pp.item_ks[n_items] = static_cast<uint8_t>(ks);
    ++n_items;
  };
  if (!s2) {
    int halo = 0;
    int shift[kMaxTaps];
    for (int t = 0; t < taps; ++t) {
      const int r = t / c.kw, s = t % c.kw;
      shift[t] = (r * c.dil - c.ph) * Wp + (s - c.pw);
      halo = std::max(halo, std::abs(shift[t]));
    }
    pp.halo = halo;
    pp.a_rows = round_up(128 + 2 * halo, 8);
    if (nkb > 24 || nkb * taps > 24) return 0;
    pp.n_boxes = nkb;
    for (int kb = 0; kb < nkb; ++kb) {
      pp.box_map[kb] = 0; pp.box_c[kb] = static_cast<int16_t>(kb * 64); pp.box_item0[kb] = static_cast<uint8_t>(n_items);
      for (int t = 0; t < taps; ++t) add_item(halo + shift[t], t * c.kpad + kb * 64, kb == nkb - 1 ? ks_last : 4);
    }
    pp.box_item0[nkb] = static_cast<uint8_t>(n_items);
  } else {
    // 2-D tile mode: tile = tile_h output rows; phase (pr, pc) holds the input pixels (2 k + pr, 2 j + pc).  Tap (dh, dw) of output (R, c) reads
    // input (2 R + dh - ph, 2 c + dw - pw) = phase ((dh - ph) & 1, (dw - pw) & 1) at (R + dr, c + dc), dr / dc = floor((d - pad) / 2) in {-1, 0}.
    // 3x3: the box starts one row / one column early, tap displacement (dr + 1) * bw + (dc + 1); 1x1: phase (0, 0), no halo.
    const bool k3 = c.kh == 3;
    pp.tile_bw = k3 ? Wp + 1 : Wp;
    if (pp.tile_bw > 128) return 0;
    pp.tile_h = 128 / pp.tile_bw;                       // as many output rows as fit the 128 accumulator rows
    pp.box_h = k3 ? pp.tile_h + 1 : pp.tile_h;
    pp.a_col0 = k3 ? -1 : 0; pp.a_row0 = k3 ? -1 : 0;
    if (pp.box_h > 256 || pp.tile_bw * pp.box_h > 224) return 0;
    pp.a_rows = pp.tile_bw * pp.box_h;
    const int n_phase = k3 ? 4 : 1;
    if (n_phase * nkb > 16 || taps * nkb > 24) return 0;
    pp.n_boxes = 0;
    for (int ph = 0; ph < n_phase; ++ph)
      for (int kb = 0; kb < nkb; ++kb) {
        const int b = pp.n_boxes++;
        pp.box_map[b] = static_cast<uint8_t>(ph); pp.box_c[b] = static_cast<int16_t>(kb * 64); pp.box_item0[b] = static_cast<uint8_t>(n_items);
        for (int t = 0; t < taps; ++t) {
          const int dh = t / c.kw - c.ph, dw = t % c.kw - c.pw;                  // input offset relative to (2 R, 2 c)
          const int pr = dh & 1, pc = dw & 1, dr = (dh - pr) / 2, dc = (dw - pc) / 2;
          if (pr * 2 + pc != ph) continue;
          add_item(k3 ? (dr + 1) * pp.tile_bw + (dc + 1) : 0, t * c.kpad + kb * 64, kb == nkb - 1 ? ks_last : 4);
        }
      }
    pp.box_item0[pp.n_boxes] = static_cast<uint8_t>(n_items);
  }
  if (pp.a_rows > 256 && !s2) return 0;
  pp.a_bytes = static_cast<uint32_t>(round_up((pp.a_rows + (s2 ? 48 : 0)) * 128, 1024));   // 2-D tiles: the unused accumulator rows 126 / 127 read past the box
  if (pp.b_resident && (pp.n_tiles != 1 || (n_tile / 2) % 8 != 0)) return 0;
  const long long bres = pp.b_resident ? static_cast<long long>(n_items) * pp.b_item_bytes : 0;
  pp.stage_bytes = pp.a_bytes + (pp.b_resident ? 0u : pp.b_item_bytes);
  const long long room = 227 * 1024 - 1024 - (1024 + 8192 + 1024) - 8 * 4096 - bres;
  if (room < 3LL * pp.stage_bytes) return 0;
  pp.stages = static_cast<int>(room / pp.stage_bytes);
  if (pp.stages > 8) pp.stages = 8;
  pp.idesc = ptx::make_idesc_f16(is_bf16_ ? 1u : 0u, 256u, static_cast<uint32_t>(n_tile));
  pp.scale = c.d_scale; pp.shift = c.d_shift;
  pp.pix_valid = d_pix_valid_[tout.stage];
  pp.pre_relu = c.pre_relu; pp.post_relu = c.post_relu;
  const uint64_t P_cap = static_cast<uint64_t>(rows_cap_[tout.stage]) * Wp;
  pp.P_cap = static_cast<long long>(P_cap);
  PairMaps& pm = c.pmaps;
  memset(&pm, 0, sizeof pm);
  if (!s2) {
    const int a_width = (c.in.coff == 0 && tin.C > c.cin && tin.C <= c.kpad) ? tin.C : c.cin;
    const uint64_t dims[2] = {static_cast<uint64_t>(a_width), P_cap};
    const uint64_t str[1] = {static_cast<uint64_t>(tin.C) * esz};
    const uint32_t box[2] = {64u, static_cast<uint32_t>(pp.a_rows)};
    const int pitch = tin.C * 2, off = c.in.coff * 2, wb = c.cin * 2;
    const int promo = (a_width == tin.C || (pitch % 128 == 0 && off % 128 == 0 && wb % 128 == 0)) ? 128 : 64;
    if (encode_tmap(&pm.a, is_bf16_, static_cast<uint8_t*>(tin.ptr) + static_cast<size_t>(c.in.coff) * esz, 2, dims, str, box, 128, promo)) return 1;
  } else {
    // parity-phase views of the high-resolution input: pixel (k, j) of phase (pr, pc) = input pixel (2 k + pr, 2 j + pc)
    const int Wp_in = stage_Wp_[tin.stage], rows_in = rows_cap_[tin.stage];
    const int n_phase = c.kh == 3 ? 4 : 1;
    // channels past cin meet zero weight rows; they only have to exist and be finite: the rest of the tensor's row, at most the K pad
    const int a_width = std::min(tin.C - c.in.coff, c.kpad);
    for (int ph = 0; ph < n_phase; ++ph) {
      const int pr = ph >> 1, pc = ph & 1;
      const uint64_t dims[3] = {static_cast<uint64_t>(a_width), static_cast<uint64_t>((Wp_in - pc + 1) / 2), static_cast<uint64_t>((rows_in - pr + 1) / 2)};
      const uint64_t str[2] = {2ull * tin.C * esz, 2ull * Wp_in * tin.C * esz};
      const uint32_t box[3] = {64u, static_cast<uint32_t>(pp.tile_bw), static_cast<uint32_t>(pp.box_h)};
      uint8_t* base = static_cast<uint8_t*>(tin.ptr) + (static_cast<size_t>(pr) * Wp_in + pc) * tin.C * esz + static_cast<size_t>(c.in.coff) * esz;
      if (encode_tmap(ph == 0 ? &pm.a : &pm.a2[ph - 1], is_bf16_, base, 3, dims, str, box, 128, 64)) return 1;
    }
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(taps) * c.kpad, static_cast<uint64_t>(c.n_pad)};
    const uint64_t str[1] = {static_cast<uint64_t>(taps) * c.kpad * esz};
    const uint32_t box[2] = {64u, static_cast<uint32_t>(n_tile / 2)};
    if (encode_tmap(&pm.b, is_bf16_, c.d_wgt, 2, dims, str, box, 128)) return 1;
  }
  if (conv_pair_smem_bytes(pp) > 227 * 1024) return 0;
  c.pp = pp;
  c.use_pair = true;
  return 0;
}

int Model::ensure_capacity(int rows0) {
  if (!rows_cap_.empty() && rows0 <= rows_cap_[0]) return 0;
  SVX_CUDA(cudaDeviceSynchronize());
  for (void* p : act_bufs_) cudaFree(p);
  act_bufs_.clear();
  for (auto p : d_seg_of_row_) cudaFree(p);
  for (auto p : d_pix_valid_) cudaFree(p);
  d_seg_of_row_.assign(n_stages_, nullptr);
  d_pix_valid_.assign(n_stages_, nullptr);
  rows_cap_.assign(n_stages_, 0);
  int rows = round_up(std::max(rows0, 256), 256);
  for (int s = 0; s < n_stages_; ++s) {
    rows_cap_[s] = rows;
    SVX_CUDA(cudaMalloc(&d_seg_of_row_[s], static_cast<size_t>(rows) * 4));
    SVX_CUDA(cudaMalloc(&d_pix_valid_[s], static_cast<size_t>(rows) * stage_Wp_[s]));
    rows = rows / 2 + 8;
  }
  for (ActTensor& t : tensors_) {
    // + 1024 pixels of slack past the last row
    const size_t bytes = (static_cast<size_t>(rows_cap_[t.stage]) * stage_Wp_[t.stage] + 1024) * t.C * 2;
    SVX_CUDA(cudaMalloc(&t.ptr, bytes));
    SVX_CUDA(cudaMemset(t.ptr, 0, bytes));
    act_bufs_.push_back(t.ptr);
  }
  for (Op& op : ops_)
    if (op.kind == OP_CONV && plan_conv(op.conv)) return 1;
  for (size_t i = 0; i < ops_.size(); ++i)
    if (ops_[i].kind == OP_CONV && ops_[i].conv.chain_pos == 0 && plan_chain(i)) return 1;
  if (has_att_ && (plan_conv(att_a_) || plan_conv(att_b_))) return 1;
  return 0;
}

// Fused hierarchical chain (res2_chain.cu): ops i, i+1, i+2 are the three 3x3 convs of a stride-1 Res2Net block.  Leaves
// chain_ok = false when the shapes do not qualify; the convs then run one by one on the flat kernel.
int Model::plan_chain(size_t i0) {
  Op& op0 = ops_[i0];
  op0.chain_ok = false;
  static const bool disabled = dbg_env("SVX_NO_CHAIN") != nullptr;   // debug switch
  if (disabled || i0 + 2 >= ops_.size()) return 0;
  ConvDesc* cv[3];
  for (int k = 0; k < 3; ++k) {
    Op& o = ops_[i0 + k];
    if (o.kind != OP_CONV || o.conv.chain_pos != k) return 0;
    cv[k] = &o.conv;
  }
  const ActTensor& ty = tensors_[cv[0]->out.id];
  const int stage = ty.stage;
  const int Wp = stage_Wp_[stage], W = stage_W_[stage];
  if (Wp <= W || Wp + 1 >= 128) return 0;                      // one zero column; the halo must be shorter than a tile
  if ((ty.C * 2) % 32 != 0 || (cv[0]->out.coff * 2) % 32 != 0) return 0;
  for (int k = 0; k < 3; ++k) {
    const ConvDesc& c = *cv[k];
    if (c.kh != 3 || c.kw != 3 || c.stride != 1 || c.dil != 1 || c.groups != 1 || c.ph != 1 || c.pw != 1) return 0;
    if (c.kpad != 32 || c.n_pad != 32 || c.cin != c.cout || c.cout > 32 || !c.post_relu || c.pre_relu || c.res.id >= 0 || c.outb.id >= 0) return 0;
    if (c.out.id != cv[0]->out.id || c.out.coff != cv[0]->out.coff + 32 * k) return 0;
    if (!c.d_scale || !c.d_shift || !c.use_flat) return 0;
    const ActTensor& tin = tensors_[c.in.id];
    if (tin.C != 32 || c.in.coff != 0 || tin.stage != stage) return 0;
    if (k < 2) {
      // conv k writes x_{k+1} + y_k to the tensor conv k+1 reads, from the planar split add2
      if (c.out2.id < 0 || c.add2.id < 0 || c.out2.id != cv[k + 1]->in.id || c.out2.coff != 0 || c.add2.coff != 0) return 0;
      if (tensors_[c.add2.id].C != 32 || tensors_[c.add2.id].stage != stage) return 0;
    } else if (c.out2.id >= 0) return 0;
  }
  ChainParams cp;
  memset(&cp, 0, sizeof cp);
  for (int t = 0; t < 9; ++t) cp.tap_shift[t] = (t / 3 - 1) * Wp + (t % 3 - 1);
  cp.idesc = ptx::make_idesc_f16(is_bf16_ ? 1u : 0u, 128u, 32u);
  for (int k = 0; k < 3; ++k) { cp.scale[k] = cv[k]->d_scale; cp.shift[k] = cv[k]->d_shift; }
  cp.pix_valid = d_pix_valid_[stage];
  cp.y = static_cast<uint8_t*>(ty.ptr) + static_cast<size_t>(cv[0]->out.coff) * 2;
  cp.y_pitch = static_cast<uint32_t>(ty.C * 2);
  const uint64_t P_cap = static_cast<uint64_t>(rows_cap_[stage]) * Wp;
  cp.P_cap = static_cast<long long>(P_cap);
  ChainMaps& cm = op0.cmaps;
  memset(&cm, 0, sizeof cm);
  const int src[3] = {cv[0]->in.id, cv[0]->add2.id, cv[1]->add2.id};      // the planar splits x_0, x_1, x_2
  for (int k = 0; k < 3; ++k) {
    const uint64_t dims[2] = {32, P_cap};
    const uint64_t str[1] = {64};
    const uint32_t box[2] = {32, 128};
    if (encode_tmap(&cm.x[k], is_bf16_, tensors_[src[k]].ptr, 2, dims, str, box, 64, 128)) return 1;
    const uint64_t wd[2] = {9 * 32, 32};
    const uint64_t ws[1] = {9 * 32 * 2};
    const uint32_t wb[2] = {32, 32};
    if (encode_tmap(&cm.w[k], is_bf16_, cv[k]->d_wgt, 2, wd, ws, wb, 64)) return 1;
  }
  op0.cp = cp;
  op0.chain_ok = true;
  return 0;
}

int Model::launch_chain(size_t i0, cudaStream_t st) {
  Op& op0 = ops_[i0];
  const ConvDesc& c0 = op0.conv;
  const int stage = tensors_[c0.out.id].stage;
  if (time_convs_) {
    while (events_.size() < ev_used_ + 2) { cudaEvent_t e; SVX_CUDA(cudaEventCreate(&e)); events_.push_back(e); }
    SVX_CUDA(cudaEventRecord(events_[ev_used_], st));
  }
  op0.cp.P = static_cast<long long>(rows_used_[stage]) * stage_Wp_[stage];
  op0.cp.dbg = flat_dbg_words();
  if (tensor_dir_.size() != tensors_.size()) tensor_dir_.assign(tensors_.size(), 0);
  tensor_dir_[c0.out.id] = 0;                        // the chain walks the pixels forwards
  SVX_CUDA(launch_res2_chain(op0.cp, op0.cmaps, is_bf16_, st));
  if (time_convs_) {
    SVX_CUDA(cudaEventRecord(events_[ev_used_ + 1], st));
    ev_used_ += 2;
    double pix = 0.0;
    for (int h : seg_h_host_[stage]) pix += static_cast<double>(h) * stage_W_[stage];
    const double fl = 3.0 * 2.0 * pix * 9.0 * c0.cin * c0.cout;
    conv_flops_ += fl;
    if (conv_labels_.size() < ev_used_ / 2) {
      char lb[256];
      snprintf(lb, sizeof lb, "stage %d fused 3x3 chain x3 cin %4d cout %4d  %.1f GFLOP", stage, c0.cin, c0.cout, fl * 1e-9);
      conv_labels_.push_back(lb);
    }
  }
  ++launches_;
  return 0;
}

// Tall-image layout of one call: per stage, where each segment starts and how tall it is.
//   Res2Net: stride-2 layers pad (1,1) regardless of the extent (models.py:121-134) → in_row = 2*out_row + r - 1 holds
//            for every segment when row_off[s] = 2*row_off[s+1].
//   DPN:     TF SAME pads (0,1) on even and (1,1) on odd extents [ext] → row_off[s] = 2*row_off[s+1] - 1 + (H odd).
int Model::layout_segments(const std::vector<int>& seg_len) {
  const int n = static_cast<int>(seg_len.size());
  seg_h_host_.assign(n_stages_, std::vector<int>(n));
  seg_off_host_.assign(n_stages_, std::vector<int>(n));
  for (int i = 0; i < n; ++i) {
    int h = seg_len[i];
    for (int s = 0; s < n_stages_; ++s) { seg_h_host_[s][i] = h; h = ceil_half(h); }
  }
  const int last = n_stages_ - 1;
  int off = 1;
  for (int i = 0; i < n; ++i) { seg_off_host_[last][i] = off; off += seg_h_host_[last][i] + gap_; }
  for (int s = last - 1; s >= 0; --s)
    for (int i = 0; i < n; ++i) {
      if (cfg_.family == SVX_FAMILY_DPN)
        seg_off_host_[s][i] = 2 * seg_off_host_[s + 1][i] - 1 + (seg_h_host_[s][i] & 1);
      else
        seg_off_host_[s][i] = 2 * seg_off_host_[s + 1][i];
    }
  rows_used_.assign(n_stages_, 0);
  for (int s = 0; s < n_stages_; ++s) rows_used_[s] = n ? seg_off_host_[s][n - 1] + seg_h_host_[s][n - 1] + gap_ : 0;
  // every finer stage must still cover 2*rows of the coarser one (stride-2 reads) — guaranteed by the capacity rule
  if (ensure_capacity(rows_used_[0])) return 1;
  for (int s = 0; s < n_stages_; ++s)
    if (rows_used_[s] > rows_cap_[s]) { set_last_error("internal: stage capacity too small"); return 1; }
  return 0;
}

// Device time of the tensor-core conv launches of the last call (option "time_convs"): CUDA events bracket every
// launch on the launching stream; bench.py divides the algorithmic FLOPs of those launches by this.
int Model::conv_time(double* ms, double* flops) {
  double total = 0.0;
  static const bool print_each = dbg_env("SVX_CONV_TIMES") != nullptr;   // debug build: one line per conv launch
  for (size_t i = 0; i + 1 < ev_used_; i += 2) {
    float t = 0.f;
    SVX_CUDA(cudaEventSynchronize(events_[i + 1]));
    SVX_CUDA(cudaEventElapsedTime(&t, events_[i], events_[i + 1]));
    total += t;
    if (print_each && i / 2 < conv_labels_.size()) fprintf(stderr, "convtime %3zu %8.1f us  %s\n", i / 2, t * 1e3, conv_labels_[i / 2].c_str());
  }
  *ms = total; *flops = conv_flops_;
  return 0;
}

// Host-mapped diagnostic words of the flat kernel's bounded barrier waits (conv_flat.cu wait_dbg); they survive the trap.
static unsigned long long* g_flat_dbg_host = nullptr;
static unsigned long long* flat_dbg_words() {
  static unsigned long long* dev = nullptr;
  if (!g_flat_dbg_host) {
    if (cudaHostAlloc(&g_flat_dbg_host, 128 * 8, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(g_flat_dbg_host, 0, 128 * 8);
    if (cudaHostGetDevicePointer(&dev, g_flat_dbg_host, 0) != cudaSuccess) dev = nullptr;
  }
  return dev;
}

int Model::launch_conv(ConvDesc& c, cudaStream_t st) {
  const int out_stage = tensors_[c.out.id].stage;
  const int out_rows = rows_used_[out_stage];
  const bool flat = c.use_flat && !force_simple_ && !force_no_flat_;
  const bool pair = c.use_pair && !no_pair_ && !force_simple_ && (flat || (c.stride == 2 && !force_no_flat_ && !no_pair_s2_));
  const bool umma = !flat && !pair && c.use_umma && !force_simple_;
  if (pair || flat || umma) {
    static const char* trace_dir = dbg_env("SVX_TRACE_DIR");   // debug: per-launch event timeline of CTA 0
    static unsigned long long* d_trace = nullptr;
    static int trace_idx = 0;
    if (trace_dir) {
      if (!d_trace) cudaMalloc(&d_trace, 6 * kTraceEvents * 8);
      cudaMemsetAsync(d_trace, 0, 6 * kTraceEvents * 8, st);
    }
    if (time_convs_) {
      while (events_.size() < ev_used_ + 2) { cudaEvent_t e; SVX_CUDA(cudaEventCreate(&e)); events_.push_back(e); }
      SVX_CUDA(cudaEventRecord(events_[ev_used_], st));
    }
    if (pair) {
      c.pp.P = static_cast<long long>(out_rows) * stage_Wp_[out_stage];
      c.pp.rows = out_rows;
      c.pp.dbg = flat_dbg_words();
      { static const char* kn = dbg_env("SVX_PAIR_KNOCK"); c.pp.knock = kn ? atoi(kn) : 0; }
      if (tensor_dir_.size() != tensors_.size()) tensor_dir_.assign(tensors_.size(), 0);
      c.pp.reverse = 1 - tensor_dir_[c.in.id];
      auto mark = [&](const TensorRef& r) { if (r.id >= 0) tensor_dir_[r.id] = c.pp.reverse; };
      mark(c.out); mark(c.outb);
      for (const TensorRef& r : c.split_out) mark(r);
      SVX_CUDA(launch_conv_pair(c.pp, c.pmaps, is_bf16_, st));
    } else if (flat) {
      c.fp.P = static_cast<long long>(out_rows) * stage_Wp_[out_stage];
      c.fp.trace = trace_dir ? d_trace : nullptr;
      c.fp.dbg = flat_dbg_words();
      { static const char* kn = dbg_env("SVX_FLAT_KNOCK"); c.fp.knock = kn ? atoi(kn) : 0; }
      // walk the pixels in the opposite direction of the kernel that wrote the input: the consumer then starts on what is
      // still in L2 (+0.6 % measured on the headline step)
      static const bool no_rev = dbg_env("SVX_NO_REVERSE") != nullptr;   // debug switch
      if (tensor_dir_.size() != tensors_.size()) tensor_dir_.assign(tensors_.size(), 0);
      c.fp.reverse = no_rev ? 0 : 1 - tensor_dir_[c.in.id];
      auto mark = [&](const TensorRef& r) { if (r.id >= 0) tensor_dir_[r.id] = c.fp.reverse; };
      mark(c.out); mark(c.outb); mark(c.out2);
      for (const TensorRef& r : c.split_out) mark(r);
      SVX_CUDA(launch_conv_flat(c.fp, c.fmaps, is_bf16_, st));
    } else {
      if (tensor_dir_.size() != tensors_.size()) tensor_dir_.assign(tensors_.size(), 0);
      tensor_dir_[c.out.id] = 0;
      if (c.outb.id >= 0) tensor_dir_[c.outb.id] = 0;
      c.up.out_rows = out_rows;
      c.up.trace = trace_dir ? d_trace : nullptr;
      SVX_CUDA(launch_conv_umma(c.up, c.amaps, c.bmap, c.auxmap, c.omaps, is_bf16_, st));
    }
    if (trace_dir) {
      std::vector<unsigned long long> h(6 * kTraceEvents);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost);
      char path[512];
      snprintf(path, sizeof path, "%s/trace%04d_%s_k%dx%d_cin%d_cout%d_s%d_aux%d.bin", trace_dir, trace_idx++, flat ? "flat" : "umma", c.kh, c.kw,
               c.cin, c.cout, c.stride, flat ? c.fp.aux_mode : c.up.aux_mode);
      FILE* f = fopen(path, "wb");
      if (f) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
    if (time_convs_) {
      SVX_CUDA(cudaEventRecord(events_[ev_used_ + 1], st));
      ev_used_ += 2;
      // algorithmic FLOPs: 2 * valid output pixels * taps * cin * cout (no padding waste counted)
      double pix = 0.0;
      for (int h : seg_h_host_[out_stage]) pix += static_cast<double>(h) * stage_W_[out_stage];
      const int cin_own = c.fold_cin > 0 ? c.fold_off : c.cin;      // (a folded shortcut's channels are real input channels too)
      const double fl = 2.0 * pix * c.kh * c.kw * (((c.in_gw > 0 ? (cin_own / c.in_gwp) * c.in_gw : cin_own) + c.fold_cin) / c.groups) * c.cout;
      conv_flops_ += fl;
      if (conv_labels_.size() < ev_used_ / 2) {
        char lb[256];
        snprintf(lb, sizeof lb, "stage %d %dx%d s%d g%d cin %4d cout %4d aux %d %s n_tile %3d x%d mt %d bres %d a_st %d b_st %d slots %d direct %d  %.1f GFLOP", out_stage,
                 c.kh, c.kw, c.stride, c.groups, c.cin, c.cout, flat ? c.fp.aux_mode : c.up.aux_mode, pair ? "pair" : flat ? "flat" : "umma", pair ? c.pp.n_tile : flat ? c.fp.n_tile : c.up.n_tile,
                 pair ? c.pp.n_tiles : flat ? c.fp.n_tiles : c.up.n_tiles, flat ? c.fp.mt : 1, flat ? c.fp.b_resident : (c.up.bres_bytes != 0), flat ? c.fp.a_stages : c.up.stages,
                 flat ? c.fp.b_stages : 0, flat ? c.fp.slots : 0, flat ? c.fp.direct : 0, fl * 1e-9);
        conv_labels_.push_back(lb);
      }
    }
  } else {
    c.sp.out_rows = out_rows;
    SVX_CUDA(launch_conv_simple(c.sp, is_bf16_, st));
  }
  ++launches_;
  return 0;
}

// ------------------------------------------------------------------------------------------------ execution
int Model::ensure_seg_capacity(int n) {
  if (n <= seg_cap_) return 0;
  SVX_CUDA(cudaDeviceSynchronize());
  const int cap = round_up(n, 1024);
  for (auto p : d_seg_row_off_) cudaFree(p);
  for (auto p : d_seg_h_) cudaFree(p);
  cudaFree(d_seg_frame_off_); cudaFree(d_seg_len_); cudaFree(d_utt_seg_off_);
  d_seg_row_off_.assign(n_stages_, nullptr); d_seg_h_.assign(n_stages_, nullptr);
  for (int s = 0; s < n_stages_; ++s) {
    SVX_CUDA(cudaMalloc(&d_seg_row_off_[s], cap * 4));
    SVX_CUDA(cudaMalloc(&d_seg_h_[s], cap * 4));
  }
  SVX_CUDA(cudaMalloc(&d_seg_frame_off_, cap * 4));
  SVX_CUDA(cudaMalloc(&d_seg_len_, cap * 4));
  SVX_CUDA(cudaMalloc(&d_utt_seg_off_, (cap + 1) * 4));
  seg_cap_ = cap;
  return 0;
}

// Runs the network on segments given by (frame start, length) pairs → d_seg_emb_[n_seg, E].
static int grow(float** p, size_t* have, size_t need) {
  if (need <= *have) return 0;
  cudaFree(*p);
  *p = nullptr;
  if (cudaMalloc(p, need) != cudaSuccess) return 1;
  *have = need;
  return 0;
}

int Model::run_segments(const float* d_feats, const int32_t* h_frame_off, int n_seg, float* d_out, cudaStream_t st) {
  if (n_seg <= 0) { launches_ = 0; return 0; }
  std::vector<int> starts(n_seg), lens(n_seg);
  for (int i = 0; i < n_seg; ++i) { starts[i] = h_frame_off[i]; lens[i] = h_frame_off[i + 1] - h_frame_off[i]; }
  return run_segments_sl(d_feats, starts, lens, d_out, st);
}

int Model::run_segments_sl(const float* d_feats, const std::vector<int>& starts, const std::vector<int>& lens, float* d_out, cudaStream_t st) {
  if (!finalized_) { set_last_error("extractor not finalized"); return 1; }
  SVX_CUDA(cudaSetDevice(device_));
  launches_ = 0;
  if (!in_extract_) { ev_used_ = 0; conv_flops_ = 0.0; conv_labels_.clear(); }
  const int n_seg = static_cast<int>(starts.size());
  if (n_seg <= 0) return 0;
  for (int i = 0; i < n_seg; ++i)
    if (lens[i] <= 0) { set_last_error("empty segment"); return 1; }
  // sub-batches bounded by workspace rows
  const int max_rows0 = 1 << 17;
  int i0 = 0;
  while (i0 < n_seg) {
    int i1 = i0; long long rows = 8;
    while (i1 < n_seg && (i1 == i0 || rows + lens[i1] + 8 <= max_rows0)) { rows += lens[i1] + 8; ++i1; }
    const int nb = i1 - i0;
    std::vector<int> sl(lens.begin() + i0, lens.begin() + i1);
    if (layout_segments(sl)) return 1;
    if (ensure_seg_capacity(nb)) return 1;
    // stage the tables through pinned memory
    const size_t words = static_cast<size_t>(nb) * (2 * n_stages_ + 1);
    if (words * 4 > h_stage_bytes_) {
      SVX_CUDA(cudaStreamSynchronize(st));
      if (h_stage_) cudaFreeHost(h_stage_);
      h_stage_bytes_ = round_up(static_cast<int>(words * 4), 1 << 16);
      SVX_CUDA(cudaMallocHost(&h_stage_, h_stage_bytes_));
    } else {
      SVX_CUDA(cudaStreamSynchronize(st));   // previous call's async table copies must have drained
    }
    int32_t* hp = h_stage_;
    for (int s = 0; s < n_stages_; ++s) {
      for (int i = 0; i < nb; ++i) hp[i] = seg_off_host_[s][i];
      SVX_CUDA(cudaMemcpyAsync(d_seg_row_off_[s], hp, nb * 4, cudaMemcpyHostToDevice, st));
      hp += nb;
      for (int i = 0; i < nb; ++i) hp[i] = seg_h_host_[s][i];
      SVX_CUDA(cudaMemcpyAsync(d_seg_h_[s], hp, nb * 4, cudaMemcpyHostToDevice, st));
      hp += nb;
    }
    for (int i = 0; i < nb; ++i) hp[i] = starts[i0 + i];
    SVX_CUDA(cudaMemcpyAsync(d_seg_frame_off_, hp, nb * 4, cudaMemcpyHostToDevice, st));
    for (int s = 0; s < n_stages_; ++s) {
      SVX_CUDA(launch_fill_row_map(d_seg_of_row_[s], d_pix_valid_[s], rows_used_[s], d_seg_row_off_[s], d_seg_h_[s], nb, stage_W_[s], stage_Wp_[s], st));
      launches_ += 1;
    }
    if (grow(&d_pooled_, &pooled_bytes_, static_cast<size_t>(nb) * flat_dim_ * 4) ||
        grow(&d_fc_partial_, &fc_partial_bytes_, static_cast<size_t>(fc_splits(flat_dim_, nb, cfg_.embed_dim)) * nb * cfg_.embed_dim * 4)) {
      set_last_error("allocation failed"); return 1;
    }
    int op_index = 0, chain_skip = 0;
    const char* dump_dir = dump_dir_.empty() ? nullptr : dump_dir_.c_str();   // svx_extractor_set_dump_dir: raw dump of every op's destination tensor
    for (Op& op : ops_) {
      struct Dump {
        Model* m; Op& op; int idx; const char* dir; cudaStream_t st;
        ~Dump() {
          if (!dir) return;
          const TensorRef& r = op.kind == OP_CONV ? op.conv.out : op.out;
          if (r.id < 0) return;
          cudaStreamSynchronize(st);
          const ActTensor& t = m->tensors_[r.id];
          const size_t bytes = static_cast<size_t>(m->rows_used_[t.stage]) * m->stage_Wp_[t.stage] * t.C * 2;
          std::vector<char> h(bytes);
          cudaMemcpy(h.data(), t.ptr, bytes, cudaMemcpyDeviceToHost);
          char path[512];
          snprintf(path, sizeof path, "%s/op%03d_k%d_t%d_r%d_w%d_c%d_off%d_n%d.bin", dir, idx, (int)op.kind, r.id, m->rows_used_[t.stage],
                   m->stage_Wp_[t.stage], t.C, r.coff, op.kind == OP_CONV ? op.conv.cout : op.C);
          FILE* f = fopen(path, "wb");
          if (f) { fwrite(h.data(), 1, bytes, f); fclose(f); }
        }
      } dump{this, op, op_index++, dump_dir, st};
      switch (op.kind) {
        case OP_PACK_INPUT: {
          const ActTensor& t = tensors_[op.out.id];
          SVX_CUDA(launch_pack_input(d_feats, d_seg_frame_off_, d_seg_row_off_[0], d_seg_of_row_[0], t.ptr, rows_used_[0], cfg_.feat_dim,
                                     t.C, is_bf16_, st));
          ++launches_;
          break;
        }
        case OP_STEM: {
          const ActTensor& t = tensors_[op.out.id];
          SVX_CUDA(launch_stem_conv(d_feats, d_seg_frame_off_, d_seg_row_off_[0], d_seg_h_[0], d_seg_of_row_[0], op.d_w9, op.d_scale,
                                    op.d_shift, static_cast<uint8_t*>(t.ptr) + static_cast<size_t>(op.out.coff) * 2, rows_used_[0], cfg_.feat_dim,
                                    stage_Wp_[0], op.C, round_up(op.C, 8), t.C, is_bf16_, st));
          ++launches_;
          break;
        }
        case OP_CONV:
          if (op.conv.chain_pos > 0 && chain_skip > 0) { --chain_skip; break; }     // ran inside the fused chain launch
          if (op.conv.chain_pos == 0 && op.chain_ok && !no_chain_ && !force_simple_ && !force_no_flat_) {
            if (launch_chain(static_cast<size_t>(op_index - 1), st)) return 1;
            chain_skip = 2;
            break;
          }
          if (launch_conv(op.conv, st)) return 1;
          break;
        case OP_BN_RELU: {
          const ActTensor& ti = tensors_[op.in.id];
          const ActTensor& to = tensors_[op.out.id];
          SVX_CUDA(launch_bn_relu(ti.ptr, ti.C, op.in.coff, stage_Wp_[ti.stage], op.d_scale, op.d_shift, to.ptr, to.C,
                                  rows_used_[to.stage], stage_W_[to.stage], stage_Wp_[to.stage], round_up(op.C, 8), op.stride, d_seg_of_row_[to.stage],
                                  d_seg_row_off_[to.stage], d_seg_row_off_[ti.stage], is_bf16_, st));
          ++launches_;
          break;
        }
        case OP_AVGPOOL: {
          const ActTensor& ti = tensors_[op.in.id];
          const ActTensor& to = tensors_[op.out.id];
          SVX_CUDA(launch_avgpool3x3s2(ti.ptr, ti.C, op.in.coff, rows_cap_[ti.stage], stage_W_[ti.stage], stage_Wp_[ti.stage], to.ptr, to.C,
                                       op.out.coff, rows_used_[to.stage], stage_W_[to.stage], stage_Wp_[to.stage], op.C, d_seg_of_row_[to.stage], is_bf16_, st));
          ++launches_;
          break;
        }
        case OP_SUBSAMPLE: {
          const ActTensor& ti = tensors_[op.in.id];
          const ActTensor& to = tensors_[op.out.id];
          SVX_CUDA(launch_subsample2(ti.ptr, ti.C, op.in.coff, stage_Wp_[ti.stage], to.ptr, to.C, op.out.coff, rows_used_[to.stage],
                                     stage_W_[to.stage], stage_Wp_[to.stage], op.C, d_seg_of_row_[to.stage], st));
          ++launches_;
          break;
        }
        default: break;
      }
      static const bool sync_each = dbg_env("SVX_SYNC_EACH") != nullptr;   // debug: localise a failing launch
      if (sync_each) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          char buf[512];
          const ConvDesc& c = op.conv;
          snprintf(buf, sizeof buf, "op %d kind %d failed: %s (conv %dx%d cin %d cout %d stride %d flat %d mt %d n_tile %d n_tiles %d slots %d a_stages %d "
                   "b_res %d box_ch %d aux %d halo %d)", op_index - 1, (int)op.kind, cudaGetErrorString(e), c.kh, c.kw, c.cin, c.cout, c.stride,
                   (int)c.use_flat, c.fp.mt, c.fp.n_tile, c.fp.n_tiles, c.fp.slots, c.fp.a_stages, c.fp.b_resident, c.fp.box_ch, c.fp.aux_mode, c.fp.halo);
          std::string msg = buf;
          if (g_flat_dbg_host)
            for (int i = 0; i < 64; ++i)
              if (g_flat_dbg_host[i]) {
                snprintf(buf, sizeof buf, " [wait code 0x%02llx idx %llu cta %llu warp %d count %llu]", g_flat_dbg_host[i] & 0xff,
                         (g_flat_dbg_host[i] >> 8) & 0xff, (g_flat_dbg_host[i] >> 16) & 0xffff, i & 15, g_flat_dbg_host[i] >> 32);
                msg += buf;
              }
          if (g_flat_dbg_host)
            for (int g = 0; g < 4; ++g) {
              if (!g_flat_dbg_host[64 + g * 16]) continue;
              snprintf(buf, sizeof buf, " {cta %llu prog:", g_flat_dbg_host[64 + g * 16] >> 32);
              msg += buf;
              for (int i = 0; i < 12; ++i) { snprintf(buf, sizeof buf, " %x", (unsigned)(g_flat_dbg_host[64 + g * 16 + i] & 0xffffffffu)); msg += buf; }
              msg += "}";
            }
          set_last_error(msg);
          return 1;
        }
      }
    }
    const ActTensor& tp = tensors_[pool_tensor_];
    SVX_CUDA(launch_stats_pool(tp.ptr, tp.C, pool_C_, stage_W_[tp.stage], stage_Wp_[tp.stage], d_seg_row_off_[tp.stage], d_seg_h_[tp.stage], nb,
                               d_pool_scale_, d_pool_shift_, d_pooled_, kPoolEps, is_bf16_, st));
    if (has_att_) {   // the statistics above feed the attention; the weighted statistics replace them (models.py:280-303)
      const int A = att_a_.cout, Wl = stage_W_[tp.stage], Wpl = stage_Wp_[tp.stage];
      if (grow(&d_att_bias_, &att_bias_bytes_, static_cast<size_t>(nb) * Wl * A * 4)) { set_last_error("allocation failed"); return 1; }
      SVX_CUDA(launch_att_bias(d_pooled_, d_att_wms_, d_att_bias_, nb, Wl, 2 * pool_C_, A, st));
      if (launch_conv(att_a_, st)) return 1;
      const ActTensor& t1 = tensors_[att_a_.out.id];
      SVX_CUDA(launch_att_tanh(t1.ptr, A, static_cast<long long>(rows_used_[tp.stage]) * Wpl, Wpl, Wl, d_seg_of_row_[tp.stage], d_att_bias_,
                               is_bf16_, st));
      if (launch_conv(att_b_, st)) return 1;
      const ActTensor& t2 = tensors_[att_b_.out.id];
      SVX_CUDA(launch_att_pool(tp.ptr, t2.ptr, pool_C_, Wl, Wpl, d_seg_row_off_[tp.stage], d_seg_h_[tp.stage], nb, d_pooled_, kPoolEps,
                               is_bf16_, st));
      launches_ += 3;
    }
    SVX_CUDA(launch_fc(d_pooled_, d_Wf_, d_bias_, d_fc_partial_, d_out + static_cast<size_t>(i0) * cfg_.embed_dim, nb, flat_dim_,
                       cfg_.embed_dim, st));
    launches_ += 3;
    i0 = i1;
  }
  return 0;
}

int Model::extract(const float* feats, int feats_on_device, const int32_t* h_frame_off, int n_utts, float* out, int out_on_device,
                   cudaStream_t st) {
  if (!finalized_) { set_last_error("extractor not finalized"); return 1; }
  SVX_CUDA(cudaSetDevice(device_));
  if (n_utts <= 0) return 0;
  const int F = cfg_.feat_dim, E = cfg_.embed_dim;
  // chunk rule (tf_extract.py:101-110)
  std::vector<int32_t> starts, seg_len, utt_seg_off(n_utts + 1, 0);
  for (int u = 0; u < n_utts; ++u) {
    const int T = h_frame_off[u + 1] - h_frame_off[u];
    if (T < kMinFrames) {
      set_last_error("utterance " + std::to_string(u) + " has " + std::to_string(T) +
                     " frames; fewer than 25 is a division by zero in the reference (tf_extract.py:102,111)");
      return 2;
    }
    const int nch = 1 + (T - kMinFrames) / kMaxChunk;
    for (int i = 0; i < nch; ++i) {
      const int len = ((i + 1) * kMaxChunk <= T) ? kMaxChunk : T - i * kMaxChunk;
      starts.push_back(h_frame_off[u] + i * kMaxChunk);
      seg_len.push_back(len);
    }
    utt_seg_off[u + 1] = static_cast<int32_t>(starts.size());
  }
  const int n_seg = static_cast<int>(starts.size());
  const size_t total_frames = static_cast<size_t>(h_frame_off[n_utts]);
  const float* d_feats = feats;
  if (!feats_on_device) {
    if (grow(&d_feats_, &d_feats_bytes_, total_frames * F * 4)) { set_last_error("feature staging allocation failed"); return 1; }
    SVX_CUDA(cudaMemcpyAsync(d_feats_, feats, total_frames * F * 4, cudaMemcpyHostToDevice, st));
    d_feats = d_feats_;
  }
  if (grow(&d_seg_emb_, &seg_emb_bytes_, static_cast<size_t>(n_seg) * E * 4)) { set_last_error("allocation failed"); return 1; }
  float* d_out = out;
  if (!out_on_device) {
    if (grow(&d_out_, &d_out_bytes_, static_cast<size_t>(n_utts) * E * 4)) { set_last_error("allocation failed"); return 1; }
    d_out = d_out_;
  }
  // all chunks in one pass: segments are (first frame, length) pairs, so dropped tails leave no gaps to work around
  long long total_launches = 0;
  ev_used_ = 0; conv_flops_ = 0.0; conv_labels_.clear(); in_extract_ = true;
  struct Guard { bool& f; ~Guard() { f = false; } } guard{in_extract_};
  {
    std::vector<int> st_i(starts.begin(), starts.end()), ln_i(seg_len.begin(), seg_len.end());
    if (run_segments_sl(d_feats, st_i, ln_i, d_seg_emb_, st)) return 1;
    total_launches += launches_;
  }
  bool single = true;
  for (int u = 0; u < n_utts; ++u) if (utt_seg_off[u + 1] - utt_seg_off[u] != 1) { single = false; break; }
  if (single) {
    SVX_CUDA(cudaMemcpyAsync(d_out, d_seg_emb_, static_cast<size_t>(n_utts) * E * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    if (ensure_seg_capacity(std::max(n_seg, n_utts + 1))) return 1;
    SVX_CUDA(cudaStreamSynchronize(st));
    SVX_CUDA(cudaMemcpyAsync(d_seg_len_, seg_len.data(), n_seg * 4, cudaMemcpyHostToDevice, st));
    SVX_CUDA(cudaMemcpyAsync(d_utt_seg_off_, utt_seg_off.data(), (n_utts + 1) * 4, cudaMemcpyHostToDevice, st));
    SVX_CUDA(launch_chunk_combine(d_seg_emb_, d_utt_seg_off_, d_seg_len_, d_out, n_utts, E, st));
    SVX_CUDA(cudaStreamSynchronize(st));   // seg_len / utt_seg_off are stack-owned host vectors
    ++total_launches;
  }
  launches_ = total_launches;
  if (!out_on_device) {
    SVX_CUDA(cudaMemcpyAsync(out, d_out, static_cast<size_t>(n_utts) * E * 4, cudaMemcpyDeviceToHost, st));
    SVX_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

}  // namespace svx
