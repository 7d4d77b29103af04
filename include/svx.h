/* libsvx — C ABI of the B200 speaker-verification inference library.
 *
 * The reference (xx205/voxsrc2020_speaker_verification) is pure Python over TensorFlow 1.x and NumPy and has no
 * FFI of its own; each entry point below states the reference interface it stands in for, so that a maintainer
 * can bind it with ctypes.  Conventions: every function returns 0 on success and a
 * non-zero status otherwise, svx_last_error() then describes the failure (thread-local); no exceptions cross
 * the ABI; the caller owns all data buffers and the CUDA stream, the library owns handles and workspaces;
 * one handle per GPU, a handle is not thread-safe, different handles are independent.
 */
#ifndef SVX_H_
#define SVX_H_

#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define SVX_FAMILY_TDNN 0
#define SVX_FAMILY_RES2NET 1
#define SVX_FAMILY_DPN 2

#define SVX_PRECISION_FP16 0 /* fp16 operands, fp32 accumulate (default: meets cos >= 0.9999) */
#define SVX_PRECISION_BF16 1 /* bf16 operands, fp32 accumulate */

/* Architecture of one extractor: the fields of the reference's module-level model instances
 * (tensorflow/models/tdnn_model.py:158-161, res2net_model.py:246-280, dpn_model.py:171). */
typedef struct svx_model_config {
  int32_t family;   /* SVX_FAMILY_* */
  int32_t feat_dim; /* FBANK bins: 40 or 80 (global_config.sh:18) */
  int32_t embed_dim;
  /* TDNN */
  int32_t tdnn_layers;
  int32_t tdnn_filters[8];
  int32_t tdnn_kernels[8];
  int32_t tdnn_dilations[8];
  /* Res2Net */
  int32_t num_filters[4];
  int32_t width[4];
  int32_t split;
  int32_t block_sizes[4];
  int32_t block_strides[4];
  /* DPN */
  int32_t init_features;
  int32_t bw;
  int32_t k_r;
  int32_t cardinality;
  int32_t k_sec[4];
  int32_t inc_sec[4];
  /* attentive statistics pooling (tensorflow/models/models.py:273-303; res2net_model.py:264-280): 0 = plain stats_pool */
  int32_t att_pool;
  int32_t att_dim; /* 128 in the reference */
} svx_model_config;

typedef struct svx_extractor svx_extractor;

int svx_version(void);
/* bit 0: debug build (SVX_* environment switches compiled in); bit 1: the cta_group::2 instantiation of the flat conv kernel. */
int svx_build_flags(void);
const char* svx_last_error(void);

/* ---- extractor: replaces the frozen graph `model/inputs:0 → model/outputs:0` that tf_extract.py:75-82 imports
 * and runs with sess.run (tf_extract.py:108). */
int svx_extractor_create(const svx_model_config* cfg, int device, int precision, svx_extractor** out);
int svx_extractor_destroy(svx_extractor* h);
/* The frozen-graph variables the model expects, in TF creation order (names as produced by freeze_graph,
 * export_inference_model.sh:40-44): conv2d{,_k}/kernel, batch_normalization{,_k}/moving_{mean,variance}, dense/kernel. */
int svx_extractor_num_tensors(svx_extractor* h);
int svx_extractor_tensor_info(svx_extractor* h, int index, const char** name, int* ndim, int64_t shape[4]);
/* Host fp32 data in TF layout (kernels HWIO). Shape is validated. */
int svx_extractor_set_tensor(svx_extractor* h, const char* name, const float* data, int ndim, const int64_t* shape);
/* Folds batch norms into per-channel scale/shift, converts and re-lays-out weights, uploads. */
int svx_extractor_finalize(svx_extractor* h);
int svx_extractor_embed_dim(svx_extractor* h);
/* "force_simple" = 1 routes every conv through the CUDA-core kernel (debug cross-check). */
int svx_extractor_set_option(svx_extractor* h, const char* key, int value);
/* One graph evaluation per segment: segment i = rows [frame_offsets[i], frame_offsets[i+1]) of feats
 * ([total_frames, feat_dim] fp32, device pointer); out = device fp32 [n_segments, embed_dim].
 * Same as sess.run(outputs, {inputs: x}) for each segment on its own (batch 1, tf_extract.py:84,108). */
int svx_extractor_run_segments(svx_extractor* h, const float* feats_dev, const int32_t* frame_offsets_host, int n_segments,
                               float* out_dev, void* cuda_stream);
/* Whole utterances with the reference chunk rule (tf_extract.py:96-111): <=1000-frame chunks, tail < 25 frames
 * dropped, length-weighted mean of chunk embeddings.  feats/out may be host (pinned or pageable) or device.
 * Fails for utterances shorter than 25 frames (the reference divides by zero there). */
int svx_extractor_extract(svx_extractor* h, const float* feats, int feats_on_device, const int32_t* frame_offsets_host,
                          int n_utts, float* out, int out_on_device, void* cuda_stream);
/* Parity tooling: with a directory set, every op of the next runs writes its destination tensor there after it ran (one raw
 * 16-bit NHWC file per op, name = op index, kind, tensor id, rows, row pitch, channels, channel offset, width); NULL / "" = off.
 * tests/test_gpu_blocks.py feeds these to the oracle block by block. */
int svx_extractor_set_dump_dir(svx_extractor* h, const char* dir);
/* Number of kernels launched by the last run/extract call. */
long long svx_extractor_last_launches(svx_extractor* h);
/* With option "time_convs" = 1: summed device time (CUDA events on the launching stream) of the tensor-core conv
 * launches of the last call, and their algorithmic FLOPs (2*MACs over valid pixels, no padding waste). */
int svx_extractor_conv_time(svx_extractor* h, double* ms, double* flops);

/* ---- front-end: replaces the Kaldi pipe `apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:... ark:- |`
 * that tf_extract.py:63 reads its features through.  feats/out: device fp32 [total_frames, feat_dim] (may alias),
 * frame_offsets_host: int32 [n_utts + 1].  Window mean in double precision, window shifted to stay inside the utterance. */
/* center = 0 follows Kaldi's rule for causal windows: [t - cmn_window, t + 1), and while fewer than min_window frames have been
 * seen the window ends at min_window (Kaldi's --min-cmn-window, default 100), so the first frames look ahead. */
int svx_cmvn_sliding(const float* feats_dev, float* out_dev, const int32_t* frame_offsets_host, int n_utts, int feat_dim,
                     int cmn_window, int center, int min_window, void* cuda_stream);

/* Kaldi CompressedMatrix ('CM ') records → fp32 [total_frames, feat_dim] on the device, bit-identical to
 * kaldi_io._read_compressed_mat (kaldi_io.py:471-504).  blob_dev: the records' payloads back to back, each starting at its
 * global header {min f32, range f32, rows i32, cols i32} (i.e. right after the "CM " token); record_offsets_host[i] = byte
 * offset of record i in the blob; frame_offsets_host: int32 [n + 1] cumulative row counts. */
/* Validated before anything is decoded: every record's own header must say rows = its frame_offsets slice and cols = feat_dim,
 * and the record must end inside blob_bytes (a 40-dim ark fed to an 80-dim model fails here instead of decoding garbage). */
int svx_decode_compressed(const uint8_t* blob_dev, int64_t blob_bytes, const int64_t* record_offsets_host,
                          const int32_t* frame_offsets_host, int n_records, int feat_dim, float* out_dev, void* cuda_stream);

/* ---- scoring: replaces the NumPy body of tensorflow/snorm.py.  All pointers are device pointers. */
typedef struct svx_scorer svx_scorer;
int svx_scorer_create(int device, svx_scorer** out);
int svx_scorer_destroy(svx_scorer* h);
/* snorm.l2norm (snorm.py:23-25) applied per row, as read_xvector does on load (snorm.py:32). */
int svx_l2norm_rows(const float* in_dev, float* out_dev, int64_t n, int d, void* cuda_stream);
/* read_speaker_xvector (snorm.py:45-67): out[g] = mean of the unit rows member_rows[group_offsets[g] .. group_offsets[g+1])
 * (CSR; members in the order the reference appends them).  Sequential fp32 sum in that order, then a true divide by the
 * count — what np.mean(matrix, axis=0) does (snorm.py:63) — so the result is deterministic and bit-identical to it. */
int svx_group_means(const float* unit_rows_dev, int d, const int32_t* member_rows_dev, const int32_t* group_offsets_dev,
                    float* out_dev, int n_groups, void* cuda_stream);
/* get_cohort_mean_std (snorm.py:83-110): for every test row, mean and population std of its topk largest dot
 * products with the cohort rows (whole cohort when topk > c). test [n,d] unit rows, cohort [c,d]. */
int svx_asnorm_stats(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                     float* mean_dev, float* std_dev, void* cuda_stream);
/* Cohort row-sharding (one shard per GPU): the topk largest dot products of every test row with THIS shard's
 * cohort rows, unordered, vals [n, topk] (padded with -1e30 when the shard has fewer than topk rows) ... */
int svx_cohort_topk_values(svx_scorer* h, const float* test_dev, int64_t n, const float* cohort_dev, int c, int d, int topk,
                           float* vals_dev, void* cuda_stream);
/* ... and, after an all-gather of the shards' candidate lists into vals [n, m] (row pitch ld), the same mean /
 * population std of the topk largest as get_cohort_mean_std (snorm.py:104-106). */
int svx_topk_stats(const float* vals_dev, int ld, int64_t n, int m, int topk, float* mean_dev, float* std_dev, void* cuda_stream);
/* get_cosine_score + get_asnorm1_score (snorm.py:113-131) for index-pair trials; snorm_dev may be NULL
 * (cosine only, mean/std unused). */
int svx_trial_scores(const float* emb_dev, int d, const int32_t* idx1_dev, const int32_t* idx2_dev, int64_t n_trials,
                     const float* mean_dev, const float* std_dev, float* cos_dev, float* snorm_dev, void* cuda_stream);
long long svx_scorer_last_launches(svx_scorer* h);
/* compute_eer_and_min_dcf (eer_minDCF.py:41-64): scores / labels (1 = target) of n trials on the device -> out_host[4] =
 * {eer, eer_threshold, min_dcf, min_dcf_threshold}, on the ROC sklearn.metrics.roc_curve builds (descending distinct thresholds,
 * intermediate collinear points dropped, origin with threshold +inf), ties broken like np.nanargmin / the reference's `<` scan. */
int svx_eer_min_dcf(const float* scores_dev, const int32_t* labels_dev, int64_t n, double c_miss, double c_fa, double p_target,
                    double* out_host, void* cuda_stream);
/* "fused" = 0 routes svx_asnorm_stats / svx_cohort_topk_values through the unfused kernels (GEMM → score block → per-row select);
 * the default (1) is the fused kernel, which never writes the scores (shapes it does not take fall back on their own). */
int svx_scorer_set_option(svx_scorer* h, const char* key, int value);
/* How many rows of the last statistics call the fused kernel finished itself and how many it handed to the unfused kernels. */
int svx_scorer_last_path(svx_scorer* h, long long* fused_rows, long long* fallback_rows);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* SVX_H_ */
