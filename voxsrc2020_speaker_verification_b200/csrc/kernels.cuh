// Launchers for the non-GEMM kernels (elementwise.cu, scoring.cu).
#pragma once
#include "conv.cuh"

namespace svx {

cudaError_t launch_fill_row_map(int32_t* seg_of_row, uint8_t* pix_valid, int rows, const int32_t* seg_row_off, const int32_t* seg_h, int n_seg,
                                int W, int Wp, cudaStream_t st);
cudaError_t launch_pack_input(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off,
                              const int32_t* seg_of_row, void* out, int rows, int F, int Cpad, int is_bf16, cudaStream_t st);
cudaError_t launch_stem_conv(const float* feats, const int32_t* seg_frame_off, const int32_t* seg_row_off, const int32_t* seg_h,
                             const int32_t* seg_of_row, const float* w9, const float* scale, const float* shift, void* out,
                             int rows, int F, int Wp, int C, int Cpad, int pitch, int is_bf16, cudaStream_t st);   // pitch: channels per pixel of the destination tensor (>= Cpad)
cudaError_t launch_bn_relu(const void* in, int in_C, int in_coff, int in_Wp, const float* scale, const float* shift, void* out,
                           int out_C, int out_rows, int out_W, int out_Wp, int C, int stride, const int32_t* out_seg_of_row,
                           const int32_t* out_seg_row_off, const int32_t* in_seg_row_off, int is_bf16, cudaStream_t st);
// out[R, c, out_coff + ch] = in[2 R, 2 c, in_coff + ch] on the pixels of a segment, 0 elsewhere (the input of a stride-2 1x1 conv)
cudaError_t launch_subsample2(const void* in, int in_C, int in_coff, int in_Wp, void* out, int out_C, int out_coff, int out_rows, int out_W,
                              int out_Wp, int C, const int32_t* out_seg_of_row, cudaStream_t st);
cudaError_t launch_avgpool3x3s2(const void* in, int in_C, int in_coff, int in_rows, int in_W, int in_Wp, void* out, int out_C, int out_coff,
                                int out_rows, int out_W, int out_Wp, int C, const int32_t* out_seg_of_row, int is_bf16, cudaStream_t st);
cudaError_t launch_stats_pool(const void* in, int C_tot, int C, int W, int Wp, const int32_t* seg_row_off, const int32_t* seg_h, int n_seg,
                              const float* scale, const float* shift, float* out, float eps, int is_bf16, cudaStream_t st);
// attentive statistics pooling (models.py:273-303)
cudaError_t launch_att_bias(const float* pooled, const float* Wms, float* bias, int n_seg, int W, int C2, int A, cudaStream_t st);
cudaError_t launch_att_tanh(void* t, int A, long long n_pix, int Wp, int W, const int32_t* seg_of_row, const float* bias, int is_bf16,
                            cudaStream_t st);
cudaError_t launch_att_pool(const void* x, const void* logits, int C, int W, int Wp, const int32_t* seg_row_off, const int32_t* seg_h,
                            int n_seg, float* out, float eps, int is_bf16, cudaStream_t st);
int fc_splits(int D, int n, int E);   // partial-sum slabs launch_fc writes for this shape
cudaError_t launch_fc(const float* pooled, const float* Wf, const float* bias, float* partial, float* out, int n, int D, int E,
                      cudaStream_t st);
cudaError_t launch_chunk_combine(const float* seg_emb, const int32_t* utt_seg_off, const int32_t* seg_len, float* out, int n_utt,
                                 int E, cudaStream_t st);

// front-end
constexpr int32_t kCmNoBadRecord = 0x7f7f7f7f;   // cudaMemset(0x7f) pattern: no record failed validation
cudaError_t launch_cmn_sliding(const float* feats, float* out, const int32_t* frame_off_dev, int n_utts, long long total_frames, int F,
                               int window, int center, int min_window, double* csum_ws, int32_t* utt_ws, cudaStream_t st);

cudaError_t launch_cm_decode(const uint8_t* blob, const long long* rec_off_dev, const int32_t* frame_off_dev, int n_utts, long long total_frames,
                             int cols, float* out, int32_t* utt_ws, long long blob_bytes, int32_t* bad_record, cudaStream_t st);

// scoring
cudaError_t launch_l2norm_rows(const float* in, float* out, long long n, int d, cudaStream_t st);
cudaError_t launch_split3(const float* in, __nv_bfloat16* out, long long n, long long n_pad, int d, int cohort_side, cudaStream_t st);
cudaError_t launch_topk_stats(const float* scores, int ld, long long n_rows, int c, int topk, float* mean, float* stdv,
                              float* vals_out, int vals_ld, cudaStream_t st);
cudaError_t launch_trial_scores(const float* emb, int d, const int32_t* idx1, const int32_t* idx2, long long t, const float* mean,
                                const float* stdv, float* cos_out, float* snorm_out, cudaStream_t st);
// fused cohort statistics (asnorm_fused.cu)
struct AsnormFusedParams {
  const float* x; int d;       // test rows [n_rows, d] fp32 (split into bf16 hi + lo by the kernel itself, into tensor memory)
  int n_rows, c;               // valid test rows / cohort rows
  int n_row_blocks, n_tiles;   // 128-row blocks of the test matrix, 128-row tiles of the cohort (both zero padded)
  int dp, kboxes;              // padded embedding dimension (multiple of 64, <= 256) and dp / 64
  int topk;                    // k (< c)
  float z_lo, z_hi;            // histogram range: sample mean + [z_lo, z_hi) sample standard deviations
  int nb, cap;                 // bins (even) and candidate list capacity per row and epilogue group
  int stages;                  // cohort ring depth
  int knock;                   // timing experiments (debug build, SVX_ASNORM_KNOCK): 1 skip the epilogue's TMEM reads, 2 skip its arithmetic, 4 skip the MMAs
  float* mean; float* stdv;    // [n_rows] (may be null when vals is wanted only)
  float* vals; int vals_ld;    // optional: the selected scores of every row, unordered, padded with -1e30 to vals_ld
  int* flag_count; int* flag_rows;   // rows handed back to the unfused path (threshold bin outside the range / overfull)
};
cudaError_t asnorm_fused_init();
size_t asnorm_fused_smem_bytes(const AsnormFusedParams& p);
cudaError_t launch_asnorm_fused(const AsnormFusedParams& p, const CUtensorMap& map_b, int sms, cudaStream_t st);
cudaError_t launch_split2(const float* in, __nv_bfloat16* out, long long n, long long n_pad, int d, int dp, cudaStream_t st);
cudaError_t launch_gather_rows(const float* in, const int* rows, int n, int d, float* out, cudaStream_t st);
cudaError_t launch_scatter_rows(const float* in, const int* rows, int n, int d, float* out, int out_ld, cudaStream_t st);
// EER / minDCF (metrics.cu): ws == nullptr -> *ws_bytes = workspace size; out_dev = {eer, eer_threshold, min_dcf, min_dcf_threshold}
cudaError_t eer_min_dcf(const float* scores, const int32_t* labels, long long n, double c_miss, double c_fa, double p_target, double* out_dev,
                        void* ws, size_t* ws_bytes, cudaStream_t st);
cudaError_t launch_group_mean(const float* unit_rows, int d, const int32_t* member_rows, const int32_t* group_off, float* out, int n_groups,
                              cudaStream_t st);

}  // namespace svx
