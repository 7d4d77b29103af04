// Extractor model: static layer plan for TDNN / Res2Net / DPN, weight folding, workspace + TMA maps, executor.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/svx.h"
#include "conv.cuh"

namespace svx {

void set_last_error(const std::string& msg);

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
};

struct VarSpec {
  std::string name;
  std::vector<int64_t> shape;
};

enum OpKind { OP_PACK_INPUT, OP_STEM, OP_CONV, OP_BN_RELU, OP_AVGPOOL, OP_POOL, OP_SUBSAMPLE };

struct TensorRef {
  int id = -1;      // activation tensor id
  int coff = 0;     // channel offset
};

// One convolution with its fused epilogue, described with tensor ids; pointers are resolved when the
// workspace is (re)allocated.
struct ConvDesc {
  TensorRef in; int cin = 0;
  int kh = 1, kw = 1, stride = 1, dil = 1, ph = 0, pw = 0, groups = 1, cout = 0;
  std::string kernel_name;            // TF variable holding the weights
  int kernel_out_off = 0;             // first output column of that variable (hierarchical conv split, res2net_model.py:54)
  // input read from a concat whose slices are padded (in_gw real channels in every in_gwp): cin is the PADDED count, the weight
  // rows of the pad positions are zero and the pad channels of the tensor hold zeros
  int in_gw = 0, in_gwp = 0;
  // a second bias-free 1x1 conv + BN folded in as extra K (the projection shortcut of a stride-1 block, res2net_model.py:85-87 /
  // :98-101: relu(bn3(W3 y) + bn_s(Ws x)) = relu(s3 * ([W3 ; Ws * s_s / s3] [y ; x]) + b3 + b_s)): its fold_cin input channels sit
  // in the input tensor from channel fold_off on, its weight rows are scaled per output channel by s_s / s3 in fp32 before the
  // one rounding to 16 bits
  std::string fold_kernel_name, fold_bn_name; int fold_cin = 0, fold_off = 0;
  int split_store = 0;                // planar-split conv: channels stored per split (> split_w: destination tensors padded, pad written as zeros)
  int out2_store = 0;                 // same for the second output / aux tile of the direct epilogue (padded planar tensors)
  int out_store = 0;                  // channels the direct epilogue writes (>= cout: the pad of a padded concat slice is written as zeros)
  std::string bn_name;                // BN applied to the conv output ("" = none)
  int pre_relu = 0, post_relu = 0;
  TensorRef out; int n_split = -1;    // -1 → all channels to `out`
  TensorRef outb;
  TensorRef res;
  TensorRef out2, add2;
  // resolved
  void* d_wgt = nullptr; int kpad = 0, kbox = 0, nkc = 0, n_pad = 0, n_tile = 0, n_tiles = 0;
  float* d_scale = nullptr; float* d_shift = nullptr;
  bool use_umma = false; bool no_staged = false;
  UmmaConvParams up; AMaps amaps; CUtensorMap bmap; CUtensorMap auxmap; OMaps omaps;
  bool use_flat = false;
  // planar splits: output channel group s (split_w wide) goes to split_out[s]; the GEMM's N axis is laid out as
  // n_splits groups padded to split_wp channels (weights, scale and shift are permuted accordingly at upload)
  std::vector<TensorRef> split_out; int split_w = 0, split_wp = 0, split_box = 0;
  int n_gemm = 0;                     // GEMM N: cout, or n_splits * split_wp
  FlatConvParams fp; FlatMaps fmaps;
  bool use_pair = false; PairConvParams pp; PairMaps pmaps;   // CTA-pair GEMM (conv_pair.cu) instead of the flat kernel
  SimpleConvParams sp;
  int chain_pos = -1;                 // position in the hierarchical 3x3 chain of a stride-1 Res2Net block (res2_chain.cu), -1: none
};

struct Op {
  OpKind kind;
  ConvDesc conv;                       // OP_CONV
  // OP_STEM: kernel_name / bn_name / out in conv
  // OP_BN_RELU / OP_AVGPOOL / OP_SUBSAMPLE (even pixels of `in` -> `out` of the next stage):
  TensorRef in, out; int C = 0; int stride = 1; std::string bn_name;
  float* d_scale = nullptr; float* d_shift = nullptr; float* d_w9 = nullptr;
  // OP_CONV with conv.chain_pos == 0: this conv and the next two run as ONE fused launch when chain_ok
  bool chain_ok = false; ChainParams cp; ChainMaps cmaps;
};

struct ActTensor {
  int stage = 0;
  int C = 0;
  void* ptr = nullptr;
};

class Model {
 public:
  Model(const svx_model_config& cfg, int device, int precision);
  ~Model();
  int build();                                             // create op list + variable specs
  const std::vector<VarSpec>& vars() const { return vars_; }
  int set_tensor(const char* name, const float* data, int ndim, const int64_t* shape);
  int finalize();
  int embed_dim() const { return cfg_.embed_dim; }
  int device() const { return device_; }
  int set_option(const char* key, int value);
  void set_dump_dir(const char* dir) { dump_dir_ = dir ? dir : ""; }
  // segments: contiguous [frame_off[i], frame_off[i+1]) rows of feats (device fp32 [total_frames, F])
  int run_segments(const float* d_feats, const int32_t* h_frame_off, int n_seg, float* d_out, cudaStream_t st);
  // the same for segments given as (first frame, length) pairs, which need not be contiguous (chunk rule with dropped tails)
  int run_segments_sl(const float* d_feats, const std::vector<int>& starts, const std::vector<int>& lens, float* d_out, cudaStream_t st);
  // utterances with the chunk rule; host or device feats/out
  int extract(const float* feats, int feats_on_device, const int32_t* h_frame_off, int n_utts, float* out, int out_on_device,
              cudaStream_t st);
  long long last_launches() const { return launches_; }
  int conv_time(double* ms, double* flops);

 private:
  int new_tensor(int stage, int C);
  std::string next_name(std::map<std::string, int>& counts, const std::string& prefix, const std::string& base);
  void add_var(const std::string& name, std::vector<int64_t> shape);
  void build_tdnn();
  void build_res2net();
  void build_dpn();
  int ensure_capacity(int rows0);
  int plan_conv(ConvDesc& c);
  int plan_flat(ConvDesc& c);
  int plan_pair(ConvDesc& c);
  int plan_chain(size_t op_index);
  int launch_chain(size_t op_index, cudaStream_t st);
  int fold_bn(const std::string& bn, int C, bool four_d, std::vector<float>& scale, std::vector<float>& shift);
  int upload_conv_weights(ConvDesc& c);
  int launch_conv(ConvDesc& c, cudaStream_t st);
  int layout_segments(const std::vector<int>& seg_len);
  int ensure_seg_capacity(int n);

  svx_model_config cfg_;
  int device_ = 0;
  int is_bf16_ = 0;
  int force_simple_ = 0;
  std::string dump_dir_;              // non-empty: every op's destination tensor is written there after it ran (parity tests)
  bool finalized_ = false;
  std::vector<VarSpec> vars_;
  std::map<std::string, HostTensor> host_;
  std::vector<Op> ops_;
  std::vector<ActTensor> tensors_;
  int n_stages_ = 1;
  int gap_ = 1;
  std::vector<int> stage_W_;
  std::vector<int> stage_Wp_;        // pixels per row in memory: W + 1 zero column for the 2-D networks, W for the TDNN
  std::vector<uint8_t*> d_pix_valid_;
  int force_no_flat_ = 0;
  int no_pair_s2_ = 0;                // option "no_pair_s2": stride-2 convs on conv_umma.cu instead of the pair kernel's 2-D tile mode
  int no_pair_ = 0;                   // option "no_pair": deep 1x1 convs on the flat kernel instead of the CTA-pair GEMM
  int no_chain_ = 0;                  // option "no_chain": run the hierarchical 3x3 convs as separate launches
  std::vector<int> tensor_dir_;       // per activation tensor: 1 if its last writer walked the pixels backwards
  std::vector<int> rows_cap_, rows_used_;
  int seg_cap_ = 0;
  // per stage device tables
  std::vector<int32_t*> d_seg_row_off_, d_seg_h_, d_seg_of_row_;
  int32_t* d_seg_frame_off_ = nullptr;
  int32_t* d_seg_len_ = nullptr;
  int32_t* d_utt_seg_off_ = nullptr;
  int32_t* h_stage_ = nullptr;       // pinned staging for the tables
  size_t h_stage_bytes_ = 0;
  // attentive statistics pooling (models.py:273-303): two 1x1 convs on the tensor-core path + three small kernels
  bool has_att_ = false;
  ConvDesc att_a_, att_b_;
  float* d_att_wms_ = nullptr; float* d_att_bias_ = nullptr; size_t att_bias_bytes_ = 0;
  // tail
  int pool_tensor_ = -1, pool_C_ = 0, flat_dim_ = 0;
  std::string pool_bn_;
  float* d_pool_scale_ = nullptr; float* d_pool_shift_ = nullptr;
  size_t pooled_bytes_ = 0, seg_emb_bytes_ = 0, fc_partial_bytes_ = 0;
  float* d_fc_partial_ = nullptr;
  float* d_pooled_ = nullptr; float* d_Wf_ = nullptr; float* d_bias_ = nullptr; float* d_seg_emb_ = nullptr;
  std::string tail_bn1_, tail_bn2_;
  // io staging
  float* d_feats_ = nullptr; size_t d_feats_bytes_ = 0;
  float* d_out_ = nullptr; size_t d_out_bytes_ = 0;
  std::vector<void*> owned_;         // weight / scale buffers
  std::vector<void*> act_bufs_;
  long long launches_ = 0;
  int time_convs_ = 0; bool in_extract_ = false;
  std::vector<cudaEvent_t> events_; size_t ev_used_ = 0; double conv_flops_ = 0.0;
  std::vector<std::string> conv_labels_;   // one per timed conv launch (shape and plan), printed by the debug build
  std::vector<std::vector<int>> seg_h_host_, seg_off_host_;
};

// TMA descriptor encode through the driver entry point (no link-time libcuda dependency).
int encode_tmap(CUtensorMap* m, int elem_bytes_is2, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes, int l2_promo_bytes = 128);

}  // namespace svx
