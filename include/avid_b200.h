/*
 * avid_b200.h — C ABI of the B200-native AVID / AVID-CMA training hot path.
 *
 * The reference (facebookresearch/AVID-CMA) has no FFI layer: its hot path is a
 * chain of ATen calls issued from Python (criterions/{nce,avid,avid_cma}.py,
 * models/{video,audio,network_blocks,av_wrapper}.py).  Every entry point below
 * replaces one such chain; the comment on each cites the reference lines it
 * stands in for.  The Python modules in avid_cma_b200/{criterions,models} bind
 * these through ctypes and are the only callers.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - sizes are explicit, tensors are dense row-major in the layout stated;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing
 *     allocates (scratch comes in through `workspace`);
 *   - return value: 0 = ok, otherwise an AVID_E* code; avid_last_error() holds
 *     a human-readable message for the calling thread.
 *   - activations of the encoders are channels-last: [N, T, H, W, C] fp32
 *     (audio uses T = 1).  Convolution filters are [taps, Cin, Cout] fp32 where
 *     taps = kT*kH*kW in (kt, kh, kw) order ("tap-major"); avid_filter_to_tapmajor
 *     converts from / to the PyTorch [Cout, Cin, kT, kH, kW] parameter layout.
 */
#ifndef AVID_B200_H
#define AVID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVID_OK            0
#define AVID_EINVAL        1   /* bad argument (shape, null pointer, unsupported size) */
#define AVID_EWORKSPACE    2   /* workspace too small */
#define AVID_ECUDA         3   /* CUDA runtime error, see avid_last_error() */
#define AVID_EUNSUPPORTED  4

#define AVID_ABI_VERSION   1
#define AVID_MAX_KEYS      8
#define AVID_EMB_DIM       128 /* embedding width the criterion kernels are specialised for */

int         avid_version(void);
const char* avid_last_error(void);
/* number of kernels launched by this library since load / since the last reset */
uint64_t    avid_launch_count(void);
void        avid_reset_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Criterion: memory bank + NCE                                               */
/* ------------------------------------------------------------------------- */

/* One score key of AVIDSimilarityMemoryBank.forward (criterions/avid.py:69-75) or
 * AVIDSimilarityPositiveExpansion.forward (criterions/avid_cma.py:169-188).
 *   ctx      0 = video embedding is the context, 1 = audio embedding
 *   bank     0 = targets come from view1_mem (video bank), 1 = view2_mem (audio)
 *   pos_mode 0 = positive is the instance's own row y[b] (P = 1)
 *            1 = positives are positive_set[y[b], :] (P = pos_k)
 *   num_neg  this key uses the first num_neg of the shared negatives
 *   weight   contribution of the key's loss to the total (coeff / 2)            */
typedef struct avid_nce_key {
    int32_t ctx;
    int32_t bank;
    int32_t pos_mode;
    int32_t num_neg;
    float   weight;
} avid_nce_key_t;

typedef struct avid_nce_args {
    /* (B, 128) un-normalised embeddings, as returned by the towers */
    const float*   emb_video;
    const float*   emb_audio;
    const int64_t* y;              /* (B) instance indices */
    const float*   bank_video;     /* rows [row_begin, row_end) of view1_mem, (row_end-row_begin, 128) */
    const float*   bank_audio;     /* same rows of view2_mem */
    int64_t        num_rows;       /* N: logical bank size */
    int64_t        row_begin;      /* first row held by this process (0 when not sharded) */
    int64_t        row_end;        /* one past the last row held (N when not sharded) */
    int32_t        batch;          /* B: instances scored by this call */
    int32_t        mean_batch;     /* divisor of the batch mean (nce.py:57); 0 = batch.  A sharded caller
                                      scoring all W*B gathered instances passes the per-rank B here. */
    int32_t        num_neg;        /* K: shared negatives per instance */
    /* Negatives: either given (B, K) int64 (the reference's host-drawn indices,
     * alias_method.py:64-71 + avid.py:84-85 already applied) or NULL: drawn in the
     * kernel from Philox4x32-10 keyed by (seed, offset), uniform over the same support. */
    const int64_t* neg_idx;
    uint64_t       seed;
    uint64_t       offset;
    const int32_t* positive_set;   /* (N, pos_k) int32 sorted ascending per row, or NULL */
    int32_t        pos_k;
    int32_t        num_keys;
    avid_nce_key_t keys[AVID_MAX_KEYS];
    const float*   avg_exp_score;  /* device scalar Z (criterions/nce.py:21-36); must be > 0 */
    float          temperature;    /* 0.07 in the reference (avid.py:32) */
    /* outputs */
    float*         loss_keys;      /* (num_keys) per-key NCE loss (mean over the batch) */
    float*         loss_total;     /* (1) sum_k weight_k * loss_k */
    float*         grad_video;     /* (B, 128) dL_total / d emb_video (through F.normalize) */
    float*         grad_audio;     /* (B, 128) */
    float*         scores;         /* optional (num_keys, B, 1 + pos_k + K): s for [self | positives | negatives]; unused slots untouched */
    int64_t*       neg_idx_out;    /* optional (B, K): the negatives that were used */
    /* Sharded mode (row_begin/row_end a strict sub-range): the kernel only scores rows it
     * holds and leaves partial sums in grad_hat_* / loss_part for the caller to reduce;
     * loss_keys / loss_total / grad_* are then produced by avid_nce_finalize.  */
    float*         grad_hat_video; /* (B, 128) partial dL/d normalised video embedding; may be NULL when not sharded */
    float*         grad_hat_audio; /* (B, 128) */
    float*         loss_part;      /* (num_keys, B) per-instance loss terms (before the batch mean) */
    /* Optional int32 the kernel sets to 1 when some y[b] is outside [0, num_rows) (the reference raises IndexError on such an
     * index, avid.py:57-58).  The offending instance then contributes no positive; nothing is read or written out of bounds.
     * May live in pinned host memory (the caller polls it without a device sync).  NULL: not reported. */
    int32_t*       bad_index;
    /* Sharded mode, packed records (ONE all-gather in, ONE reduce-scatter out).  group_batch > 0: query b of
     * the `batch` gathered queries is row (b % group_batch) of rank-record (b / group_batch); emb_video / emb_audio / y / neg_idx
     * point at the fields of record 0 and record g starts in_group_stride BYTES further; grad_hat_* / loss_part (then laid out
     * (num_keys, group_batch) per record) likewise with out_group_stride.  0: dense (B, ...) arrays. */
    int32_t        group_batch;
    int64_t        in_group_stride;
    int64_t        out_group_stride;
} avid_nce_args_t;

/* Scratch of the criterion kernels.  Its first 4 * (batch + 1) bytes are ticket counters (splits done per query, queries
 * done) that MUST BE ZERO before the first call with a given workspace; every call leaves them zero again, so a
 * workspace zeroed once at allocation can be reused for the life of the criterion. */
size_t avid_nce_workspace_bytes(int32_t batch, int32_t num_neg, int32_t pos_k, int32_t num_keys);

/* Fused replacement for: F.normalize (avid.py:52-53), positive/negative gathers
 * (avid.py:57-62 / avid_cma.py:158-165,196-209), bmm/T scores (avid.py:65-75),
 * NCECriterion.forward (nce.py:38-58) for every key, the coefficient mix
 * (avid.py:216-233 / avid_cma.py:338-359) AND the backward of all of those
 * w.r.t. the embeddings, in one pass over the gathered rows and ONE kernel launch:
 * the last CTA of a query reduces it, the last query forms the batch means.       */
int avid_nce_forward_backward(const avid_nce_args_t* args_host, void* workspace, size_t workspace_bytes, void* stream);

/* Second half of the sharded protocol: given (all-reduced) grad_hat_* and loss_part,
 * produce loss_keys, loss_total, grad_video, grad_audio.                         */
int avid_nce_finalize(const avid_nce_args_t* args_host, void* workspace, size_t workspace_bytes, void* stream);

/* NCECriterion.compute_partition_function (nce.py:21-36) for the first batch:
 * mean over (b, k < keys[key].num_neg) of exp(score) for one key -> out_mean (1).
 * In sharded mode writes the partial SUM over held rows instead (caller divides). */
int avid_nce_partition_mean(const avid_nce_args_t* args_host, int32_t key, float* out_mean, void* workspace, size_t workspace_bytes, void* stream);

/* AVIDSimilarityMemoryBank.update_memory (avid.py:103-129) after the all-gather:
 * for i < n: if y[i] in [row_begin,row_end): m = bank[y[i]]; m = mom*m + (1-mom)*normalize(emb[i]);
 * bank[y[i]] = normalize(m), for both banks.  Duplicate y: the LAST occurrence in y wins (one complete update per row,
 * like index_copy_; never a torn row).  group_batch > 0: the n instances come as packed per-rank records of the one
 * all-gather of the step (instance i = row i % group_batch of record i / group_batch, records group_stride BYTES apart,
 * emb_video / emb_audio / y point at the fields of record 0); 0: dense arrays.  */
int avid_bank_update(float* bank_video, float* bank_audio, int64_t row_begin, int64_t row_end,
                     const float* emb_video, const float* emb_audio, const int64_t* y, int32_t n, int32_t group_batch, int64_t group_stride,
                     float momentum_video, float momentum_audio, void* stream);

/* init_memory (avid.py:88-96) without the rank-0 broadcast (avid.py:99-101): rows [row_begin, row_begin + rows) of bank
 * `which` (0 video / 1 audio) <- L2-normalised N(0,1) rows.  Row r depends only on (seed, which, r) (Philox4x32-10 +
 * Box-Muller), so every rank of a sharded run fills its own rows and replicated ranks fill identical banks. */
int avid_bank_init(float* bank, int64_t row_begin, int64_t rows, uint64_t seed, int32_t which, void* stream);

/* In-place row-wise x / max(||x||_2, 1e-12) on (rows, 128): init_memory (avid.py:92,95). */
int avid_rows_l2_normalize(float* x, int64_t rows, void* stream);

/* sample_negatives on device (avid.py:82-86; avid_cma.py:200-207): the same Philox stream
 * avid_nce_forward_backward uses when neg_idx == NULL, exposed for tests and statistics. */
int avid_sample_negatives(const int64_t* y, int32_t batch, int32_t num_neg, int64_t num_rows,
                          const int32_t* positive_set, int32_t pos_k,
                          uint64_t seed, uint64_t offset, int64_t* neg_idx_out, void* stream);

/* CMASampler.sample_instance (avid_cma.py:42-73), all queries of [q_begin, q_end):
 * sim = combine(V V_q^T, A A_q^T) over all N candidate rows, top-(pos_k+1) by
 * similarity, first hit dropped, remaining pos_k indices sorted ascending.
 *   mode 0 consensus(min) 1 union(max) 2 video 3 audio
 * cand_* are the candidate rows [cand_begin, cand_begin+num_cand) of the banks, q_* the
 * query rows; a sharded caller passes each shard in turn with the same top_val/top_idx
 * scratch (running top lists, (q_end-q_begin, 64) each) and calls avid_cma_topk_finish. */
size_t avid_cma_topk_workspace_bytes(int64_t num_queries);
int avid_cma_topk_begin(int64_t num_queries, void* workspace, size_t workspace_bytes, void* stream);
int avid_cma_topk_scan(const float* q_video, const float* q_audio, int64_t num_queries,
                       const float* cand_video, const float* cand_audio, int64_t cand_begin, int64_t num_cand,
                       int32_t mode, int32_t pos_k, void* workspace, size_t workspace_bytes, void* stream);
int avid_cma_topk_finish(int64_t num_queries, int32_t pos_k, int32_t* positive_set_out /* (num_queries, pos_k) */,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The same search (avid_cma.py:52-70: mm, mm, min/max, topk) with the N x N similarity work on the tensor cores
 * (csrc/cma_tc.cu): avid_cma_topk_scan_tc keeps an APPROXIMATE top-64 per query from fp16 inputs (tcgen05, fp32
 * accumulation; *_h are fp16 copies of the rows made by avid_cma_to_half), avid_cma_topk_rescore replaces the scores of
 * the listed candidates of the current shard by exact fp32 dot products, and -- after the last shard --
 * avid_cma_topk_certify proves per query that the exact top-(pos_k+1) lies inside the list (every outside candidate
 * has exact similarity <= list minimum + eps; eps = 1e-3 bounds the fp16 rounding of a dot product of unit rows) and
 * moves it to the front for avid_cma_topk_finish.  Queries without a certificate are counted in *fail_count and
 * listed in fail_list (num_queries entries); the caller re-mines them with avid_cma_topk_scan.  Same workspace,
 * same begin / finish calls, same shard protocol as the fp32 path. */
int avid_cma_to_half(const float* x, void* out_f16, int64_t n, void* stream);
int avid_cma_topk_scan_tc(const void* q_video_h, const void* q_audio_h, int64_t num_queries,
                          const void* cand_video_h, const void* cand_audio_h, int64_t cand_begin, int64_t num_cand,
                          int32_t mode, void* workspace, size_t workspace_bytes, void* stream);
int avid_cma_topk_rescore(const float* q_video, const float* q_audio, int64_t num_queries,
                          const float* cand_video, const float* cand_audio, int64_t cand_begin, int64_t num_cand,
                          int32_t mode, void* workspace, size_t workspace_bytes, void* stream);
int avid_cma_topk_certify(int64_t num_queries, int32_t pos_k, float eps, void* workspace, size_t workspace_bytes,
                          int32_t* fail_count, int32_t* fail_list, void* stream);

/* ------------------------------------------------------------------------- */
/* Encoders: (2+1)D / 2D convolution stacks, train-mode BatchNorm, pools, heads */
/* ------------------------------------------------------------------------- */

typedef struct avid_conv_shape {
    int32_t n, ti, hi, wi, ci;      /* input  [n, ti, hi, wi, ci] */
    int32_t to, ho, wo, co;         /* output [n, to, ho, wo, co] */
    int32_t kt, kh, kw;             /* filter taps */
    int32_t st, sh, sw;             /* strides  */
    int32_t pt, ph, pw;             /* zero padding */
} avid_conv_shape_t;

/* math mode of the contraction */
#define AVID_MATH_FP32    0   /* CUDA-core fp32 FMA (exact-parity mode)             */
#define AVID_MATH_BF16X3  1   /* tcgen05 bf16 tensor cores, 3-term split (~fp32)    */
#define AVID_MATH_BF16    2   /* tcgen05 bf16 tensor cores, single pass             */

/* nn.Conv3d / nn.Conv2d forward without bias (network_blocks.py:18,20,35-49; video.py:20;
 * audio.py:22): out = conv(in, filt) (+ addend when non-NULL: the residual sum of
 * network_blocks.py:58-59).  filt is tap-major [taps, ci, co].                   */
int avid_conv_forward(const avid_conv_shape_t* s_host, const float* in, const float* filt, const float* addend,
                      float* out, int32_t math, void* stream);
/* gradient w.r.t. the input: din = conv_transpose(dout, filt) (+ addend); filt_t is the transposed
 * tap-major filter [taps, co, ci] written by avid_filter_to_tapmajor */
int avid_conv_dgrad(const avid_conv_shape_t* s_host, const float* dout, const float* filt_t, const float* addend,
                    float* din, int32_t math, void* stream);
/* gradient w.r.t. the filter, tap-major [taps, ci, co]; dfilt must be zeroed by the caller
 * (split over pixels, accumulated with atomics) */
int avid_conv_wgrad(const avid_conv_shape_t* s_host, const float* in, const float* dout, float* dfilt,
                    int32_t math, void* stream);

/* ---- tensor-core (tcgen05) convolutions ------------------------------------------------------------
 * Operands are bf16 planes of the channels-last tensors: x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
 * (avid_split_bf16).  With both planes the kernel accumulates hi*hi + hi*lo + lo*hi in fp32 ("bf16x3",
 * AVID_MATH_BF16X3, ~16 significand bits per operand); with lo == NULL it is a plain bf16 product
 * (AVID_MATH_BF16).  Activations are fetched by TMA im2col-mode loads, filters by tiled TMA loads.
 * The forward kernels take an optional `bn_stats` (2, co) double buffer, zeroed by the caller: the epilogue adds the
 * per-channel sum and sum of squares of the stored output (after the addend), i.e. what avid_bn_stats would compute in a
 * separate pass over the tensor, so that avid_bn_finalize can follow directly.
 * Channel counts must be multiples of 64 (every layer of both towers except the two stems).
 *   forward: in planes [n,ti,hi,wi,ci], filter planes K-major [taps][co][ci] (the w_tap_t layout)
 *   dgrad  : dout planes [n,to,ho,wo,co], filter planes [taps][ci][co] (the w_tap layout); a strided gradient runs as one
 *            stride-1 correlation per stride-parity class of the input pixels (st*sh*sw classes, ONE launch)            */
int avid_split_bf16(const float* x, void* hi, void* lo /* may be NULL */, int64_t n, void* stream);
int avid_conv_forward_tc(const avid_conv_shape_t* s_host, const void* in_hi, const void* in_lo, const void* filt_hi, const void* filt_lo,
                         const float* addend, float* out, double* bn_stats, void* stream);
/* Optional fusion for avid_conv_dgrad_tc: `din` (+ addend) is the gradient at the ReLU output of the PREVIOUS conv-BN-ReLU
 * layer; with this struct the epilogue also accumulates that layer's BatchNorm-backward sums
 * (sums[0][c] += sum g, sums[1][c] += sum g * xhat, g = din * relu'(bn(z)), xhat = (z - mean) * invstd) so that
 * avid_bn_relu_backward_apply can follow without the separate avid_bn_relu_backward_reduce pass over din and z. */
typedef struct avid_bn_backward_fuse {
    const float* z;        /* previous layer's conv output, same shape as din */
    const float* mean;     /* (ci) saved batch mean / inverse std / affine parameters of that layer's BatchNorm */
    const float* invstd;
    const float* gamma;
    const float* beta;
    double*      sums;     /* (2, ci) doubles, zeroed by the caller */
} avid_bn_backward_fuse_t;
int avid_conv_dgrad_tc(const avid_conv_shape_t* s_host, const void* dout_hi, const void* dout_lo, const void* filt_hi, const void* filt_lo,
                       const float* addend, float* din, const avid_bn_backward_fuse_t* fuse_host /* may be NULL */, void* stream);
/* avid_conv_dgrad_tc with a SUBSAMPLED addend: `addend` is [n, ceil(ti/at), ceil(hi/ah), ceil(wi/aw), ci] and is added to the
 * input pixels (t, h, w) with t % at == 0, h % ah == 0, w % aw == 0 only (addend_stride = {at, ah, aw}; NULL or all ones: the dense
 * addend of avid_conv_dgrad_tc).  That is the input gradient of the strided 1x1x1 residual convolution of a stage entry
 * (network_blocks.py:46-49, :58 of the reference: x_res = res_conv(x)), which is zero at every other pixel: it is computed as a
 * stride-1 1x1x1 input gradient over the OUTPUT grid, never zero-filled to the input grid and never read back from there. */
int avid_conv_dgrad_tc_sub(const avid_conv_shape_t* s_host, const void* dout_hi, const void* dout_lo, const void* filt_hi, const void* filt_lo,
                           const float* addend, const int32_t* addend_stride_host /* {at, ah, aw}, may be NULL */, float* din,
                           const avid_bn_backward_fuse_t* fuse_host /* may be NULL */, void* stream);
/* Host-only query: the launch plan of avid_conv_forward_tc (dgrad == 0) / avid_conv_dgrad_tc (dgrad != 0) for a geometry
 * (network_blocks.py:35-49: the strided stage entries are the interesting cases).  out[0] = stride-parity classes that have work (all
 * run in ONE launch), out[1] = 128-pixel tiles of the largest class, out[2] = filter taps covered by the classes (each tap belongs
 * to exactly one class: kt*kh*kw), out[3] = destination pixels covered (every pixel a tap reaches, exactly once), out[4] = 1 when the
 * multiply-high divisions of the tile -> pixel decode reproduce integer division on a sample of indices of every class. */
int avid_conv_tc_plan(const avid_conv_shape_t* s_host, int32_t dgrad, int64_t* out_host /* [5] */);
/* Host-only query: which kernel avid_conv_forward_tc (dgrad == 0) / avid_conv_dgrad_tc (dgrad != 0) launches for this geometry:
 * 1 = conv_pair_kernel (64 -> 64 channel 1x3x3 stride-1 layers: CTA pairs, tcgen05 cta_group::2, halo strip, resident filter;
 * network_blocks.py:35-37 / :14-16 at 64 channels), 0 = conv_tc_kernel (im2col TMA, every other layer).  bench.py labels its
 * per-launch roofline records with it. */
int avid_conv_tc_uses_cta_pairs(const avid_conv_shape_t* s_host, int32_t dgrad);
/*   wgrad  : in planes [n,ti,hi,wi,ci], dout planes [n,to,ho,wo,co] -> dfilt fp32 tap-major [taps][ci][co], zeroed by the
 *            caller (split over pixels, accumulated with fp32 vector atomics); any stride                               */
int avid_conv_wgrad_tc(const avid_conv_shape_t* s_host, const void* in_hi, const void* in_lo, const void* dout_hi, const void* dout_lo,
                       float* dfilt, void* stream);

/* ---- tensor-core stems: the 7x7 / 3x7x7 stride-2 first convolutions with Cin <= 4 (video.py:20, audio.py:22) ------------
 * The input is packed once per step as bf16 planes [n][t][h][wp][4] (channels zero-padded to 4, rows zero-padded in w to
 * wp = 2*wo + 8 with pixel w stored at w + pad_left, pad_left = the convolution's pw) so that a tiled tensor map with a
 * 16-byte pixel-pair stride hands the kernel A[pixel][(kw, c)] rows without an im2col buffer (csrc/stem_tc.cu).
 * The shape's `ci` is the REAL channel count; sh == sw == 2, kw <= 8, kh <= 8, kt <= 4, co == 64.
 *   avid_stem_pack        x [n][c][t][h][w] fp32 (the reference's NCDHW / NCHW input) -> hi / lo planes
 *   avid_stem_filter_pack PyTorch filter [co][ci][kt][kh][kw] fp32 -> planes [co][kt*kh][32 = 8 kw x 4 c] (forward operand)
 *   avid_stem_forward_tc  out [n,to,ho,wo,64] fp32
 *   avid_stem_wgrad_tc    dfilt fp32 tap-major [taps][4][64] (the avid_conv_wgrad layout with ci_pad = 4), zeroed by the caller */
int avid_stem_pack(const float* x, void* hi, void* lo /* may be NULL */, int32_t n, int32_t c, int32_t t, int32_t h, int32_t w,
                   int32_t wp, int32_t pad_left, void* stream);
int avid_stem_filter_pack(const float* w_oihw, void* hi, void* lo /* may be NULL */, int32_t co, int32_t ci, int32_t kt, int32_t kh, int32_t kw,
                          void* stream);
int avid_stem_forward_tc(const avid_conv_shape_t* s_host, const void* x_hi, const void* x_lo, int32_t wp, const void* filt_hi, const void* filt_lo,
                         float* out, double* bn_stats, void* stream);
int avid_stem_wgrad_tc(const avid_conv_shape_t* s_host, const void* x_hi, const void* x_lo, int32_t wp, const void* dout_hi, const void* dout_lo,
                       float* dfilt, void* stream);

/* PyTorch parameter layout [co, ci, taps] -> tap-major [taps, ci_pad, co] (channels ci..ci_pad-1 zero) and,
 * when w_tap_t != NULL, its transpose [taps, co, ci_pad] (the filter operand of avid_conv_dgrad). */
int avid_filter_to_tapmajor(const float* w_oihw, float* w_tap, float* w_tap_t, int32_t co, int32_t ci, int32_t taps, int32_t ci_pad, void* stream);
/* tap-major [taps, ci_pad, co] (a filter gradient) -> PyTorch layout [co, ci, taps] */
int avid_filter_from_tapmajor(const float* w_tap, float* w_oihw, int32_t co, int32_t ci, int32_t taps, int32_t ci_pad, void* stream);

/* PyTorch filter [co, ci, taps] fp32 -> the bf16 (hi, lo) planes of both tensor-core operand layouts in one pass:
 * forward planes [taps][co][ci], dgrad planes [taps][ci][co] (lo planes NULL for single-pass bf16) */
int avid_filter_to_planes(const float* w_oihw, void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* dgrad_lo, int32_t co, int32_t ci, int32_t taps,
                          void* stream);
/* The same conversions for many filters in one launch each (arrays of `count` pointers / shapes on the HOST): a tower converts
 * all its tensor-core filters at the start of a step and all its filter gradients at the end of the backward pass. */
int avid_filter_to_planes_multi(const float* const* w_oihw, void* const* fwd_hi, void* const* fwd_lo, void* const* dgrad_hi,
                                void* const* dgrad_lo, const int32_t* co, const int32_t* ci, const int32_t* taps, int32_t count,
                                void* stream);
int avid_filter_from_tapmajor_multi(const float* const* w_tap, float* const* w_oihw, const int32_t* co, const int32_t* ci,
                                    const int32_t* taps, const int32_t* ci_pad, int32_t count, void* stream);

/* [n, c, thw] -> [n, thw, c_pad] (c_pad > c only for c <= 4: the 3-channel clip / 1-channel spectrogram)
 * and back [n, thw, c] -> [n, c, thw] */
int avid_nchw_to_nhwc(const float* in, float* out, int32_t n, int32_t c, int64_t thw, int32_t c_pad, void* stream);
int avid_nhwc_to_nchw(const float* in, float* out, int32_t n, int32_t c, int64_t thw, void* stream);

/* Train-mode BatchNorm{2d,3d} statistics (PyTorch defaults eps=1e-5, momentum=0.1):
 * per-channel sum / sum-of-squares of x (rows, c) in fp64 (stats, (2, c) doubles, zeroed
 * by the caller), then finalize -> scale = gamma*invstd, shift = beta - mean*scale,
 * saved mean / invstd, running stats EMA with the unbiased variance.              */
int avid_bn_stats(const float* x, int64_t rows, int32_t c, double* stats, void* stream);
int avid_bn_finalize(const double* stats, int64_t rows, int32_t c, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var,
                     float* mean, float* invstd, float* scale, float* shift, void* stream);
/* y = relu(x*scale + shift) */
int avid_bn_relu_forward(const float* x, const float* scale, const float* shift, float* y,
                         int64_t rows, int32_t c, void* stream);
/* same, writing any of: fp32 y, bf16 planes y_hi / y_lo (y = hi + lo, the operands of the tcgen05 convolutions) */
int avid_bn_relu_forward_ex(const float* x, const float* scale, const float* shift, float* y, void* y_hi, void* y_lo,
                            int64_t rows, int32_t c, void* stream);
/* backward of y = relu(bn(x)): sums (2, c) doubles zeroed by caller: sum(g), sum(g*xhat), g = dy*(y>0) */
int avid_bn_relu_backward_reduce(const float* x, const float* dy, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, int64_t rows, int32_t c,
                                 double* sums, void* stream);
/* dx = gamma*invstd*(g - sum_g/rows - xhat*sum_gx/rows); dgamma = sum_gx; dbeta = sum_g */
int avid_bn_relu_backward_apply(const float* x, const float* dy, const float* mean, const float* invstd,
                                const float* gamma, const float* beta, const double* sums,
                                int64_t rows, int32_t c, float* dx, float* dgamma, float* dbeta, void* stream);

int avid_bn_relu_backward_apply_ex(const float* x, const float* dy, const float* mean, const float* invstd,
                                   const float* gamma, const float* beta, const double* sums,
                                   int64_t rows, int32_t c, float* dx, void* dx_hi, void* dx_lo, float* dgamma, float* dbeta, void* stream);

/* Fused BatchNorm -> ReLU -> MaxPool3d((1,3,3),(1,2,2),(0,1,1)) of the video stem (video.py:21-23): the ReLU output
 * [nt, h, w, c] is never materialised.  forward: pooled = maxpool(relu(z*scale + shift)) as fp32 and / or bf16 planes, plus
 * the winning window position per pooled element; backward: the gradient at the ReLU output is gathered from the pooled
 * gradient dyp [nt, ho, wo, c] through that argmax inside the BatchNorm backward apply kernel; the backward reduce needs only
 * the pooled tensors (the gradient is non-zero at window maxima only, where y = pooled and xhat = (pooled - beta) / gamma). */
int avid_bn_relu_maxpool_forward(const float* z, const float* scale, const float* shift, float* pooled, void* pooled_hi, void* pooled_lo,
                                 uint8_t* argmax, int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo, void* stream);
int avid_bn_relu_maxpool_backward_reduce(const float* z, const float* pooled, const uint8_t* argmax, const float* dyp, const float* mean,
                                         const float* invstd, const float* gamma, const float* beta, int32_t nt, int32_t h, int32_t w, int32_t c,
                                         int32_t ho, int32_t wo, double* sums, void* stream);
int avid_bn_relu_maxpool_backward_apply(const float* z, const uint8_t* argmax, const float* dyp, const float* mean, const float* invstd,
                                        const float* gamma, const float* beta, const double* sums, int32_t nt, int32_t h, int32_t w, int32_t c,
                                        int32_t ho, int32_t wo, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta, void* stream);

/* nn.MaxPool3d((1,3,3), stride (1,2,2), padding (0,1,1)) on [n*t, h, w, c] (video.py:23).  argmax (optional, one byte per
 * output element) records the winning window position dh*3+dw -- the first maximum in scan order, like ATen. */
int avid_maxpool_1x3x3_forward(const float* x, float* y, uint8_t* argmax, int32_t nt, int32_t h, int32_t w, int32_t c,
                               int32_t ho, int32_t wo, void* stream);
/* routes dy to the recorded argmax of each window; gather form: every dx element is written exactly once, x is not read */
int avid_maxpool_1x3x3_backward(const uint8_t* argmax, const float* dy, float* dx,
                                int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo, void* stream);

/* nn.AdaptiveMaxPool{2d,3d}(1) (video.py:41, audio.py:31): x [n, thw, c] -> y [n, c], argmax [n, c] */
int avid_global_maxpool_forward(const float* x, float* y, int32_t* argmax, int32_t n, int64_t thw, int32_t c, void* stream);
/* dx zeroed by caller */
int avid_global_maxpool_backward(const float* dy, const int32_t* argmax, float* dx, int32_t n, int64_t thw, int32_t c, void* stream);

/* Head (av_wrapper.py:17-33): y = x W^T + b (optionally ReLU), W in PyTorch [out, in] layout */
int avid_linear_forward(const float* x, const float* w, const float* b, float* y,
                        int32_t rows, int32_t in_f, int32_t out_f, int32_t relu, void* stream);
/* dy is overwritten with dy*(y>0) when relu != 0; dx may be NULL; dw/db are overwritten */
int avid_linear_backward(const float* x, const float* w, const float* y, float* dy,
                         float* dx, float* dw, float* db,
                         int32_t rows, int32_t in_f, int32_t out_f, int32_t relu, void* stream);

/* elementwise helpers used by the towers */
int avid_add_inplace(float* a, const float* b, int64_t n, void* stream);   /* a += b */

/* torch.optim.Adam.step (utils/main_utils.py:250-256: lr, betas, weight_decay as L2 added to the gradient)
 * on flat fp32 buffers; `step` counts from 1; the gradient is multiplied by grad_scale first
 * (1/world_size after a sum all-reduce). */
int avid_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* the same update for `count` tensors given as HOST arrays of device pointers / element counts (32 tensors per launch) */
int avid_adam_step_multi(float* const* params_host, const float* const* grads_host, float* const* exp_avgs_host, float* const* exp_avg_sqs_host,
                         const int64_t* sizes_host, int32_t count, int64_t step, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, void* stream);
/* cudaMemsetAsync(p, 0, bytes): one call zeroes the per-step arena of accumulators (filter gradients, BatchNorm sums) */
int avid_zero_bytes(void* p, size_t bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Input side: LogSpectrogram.__call__ of datasets/preprocessing.py:158-186 for a batch of mono clips on
 * the device: |stft(n_fft, hop)|^2 (hann window, centred frames, reflect padding), bin 0 + pair means -> n_fft/4 + 1 bins,
 * first num_frames frames, power_to_db (amin 1e-10, ref 1, top_db per clip; top_db < 0: no floor), optional per-bin
 * (x - mean) / (std + 1e-5).  wave (batch, num_samples) fp32 -> out (batch, 1, num_frames, n_fft/4 + 1) fp32. */
size_t avid_log_spectrogram_workspace_bytes(int32_t batch);
int avid_log_spectrogram(const float* wave, int32_t batch, int32_t num_samples, int32_t n_fft, int32_t hop, int32_t num_frames,
                         float top_db, const float* mean, const float* stdv, float* out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Video half of the input side: VideoPrep_MSC_CJ.__call__ with augment=True (datasets/preprocessing.py:15-57) for ONE clip of
 * uint8 RGB frames on the device, with the random decisions already drawn by the host in the reference's order
 * (RandomResizedCrop.get_params video_transforms.py:330-371, RandomHorizontalFlip :86, ColorJitter.get_params + shuffle :413-463):
 * crop -> Pillow's BILINEAR Image.resize -> optional left-right flip -> the colour ops in the given order on uint8 images
 * (torchvision adjust_brightness / adjust_saturation / adjust_hue / adjust_contrast = ImageEnhance blends and an HSV round trip)
 * -> ClipToTensor (/255, volume_transforms.py:14-70) -> Normalize ((x - mean) / std, tensor_transforms.py:13-38).  Bit-identical to
 * Pillow 12: fixed-point resampling weights (Resample.c), float32 blends (Blend.c), Convert.c luma / HSV.
 * frames (T, height, width, 3) uint8 -> out (3, T, out_h, out_w) float32. */
#define AVID_VIDEO_OP_BRIGHTNESS 0
#define AVID_VIDEO_OP_SATURATION 1
#define AVID_VIDEO_OP_HUE        2
#define AVID_VIDEO_OP_CONTRAST   3
typedef struct avid_video_prep {
    int32_t frames, height, width;                    /* the clip */
    int32_t crop_top, crop_left, crop_h, crop_w;      /* (i, j, h, w) of RandomResizedCrop.get_params */
    int32_t out_h, out_w;                             /* `crop` of VideoPrep_MSC_CJ */
    int32_t flip;                                     /* random.random() < 0.5 */
    int32_t num_ops;                                  /* colour ops in application order (0..4, at most one contrast) */
    int32_t op_kind[4];                               /* AVID_VIDEO_OP_* */
    float   op_factor[4];                             /* brightness / saturation / contrast factor, hue_factor in [-0.5, 0.5] */
    int32_t hue_shift;                                /* np.uint8(hue_factor * 255) of F.adjust_hue, computed by the host in double */
    int32_t normalize;                                /* 0: stop after /255 */
    float   mean[3], std[3];
} avid_video_prep_t;
size_t avid_video_prep_workspace_bytes(const avid_video_prep_t* p_host);      /* 0 on invalid parameters */
int avid_video_prep(const uint8_t* frames, const avid_video_prep_t* p_host, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* the same for `count` clips (one loader batch) given as HOST arrays of device pointers / parameter structs: 4 launches per 16 clips */
size_t avid_video_prep_batch_workspace_bytes(const avid_video_prep_t* params_host, int32_t count);
int avid_video_prep_batch(const uint8_t* const* frames_host, const avid_video_prep_t* params_host, int32_t count, float* const* out_host,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- Sharded optimizer over NVLink peer memory -------------------------------------------------------------------------
 * Replaces DistributedDataParallel's gradient all-reduce (utils/main_utils.py:105-117) followed by torch.optim.Adam.step on every
 * rank (main_utils.py:250-256) for runs with W > 1 ranks on one node: rank r owns elements [r * S, (r + 1) * S) of the flat
 * parameter vector.  The buffers are symmetric allocations (same layout on every rank, peer-mapped); ptr[r] is rank r's buffer as
 * seen from this process.  The caller orders the ranks (gradients complete before avid_adam_shard_step, shards updated before
 * avid_pull_shards) with a barrier each. */
#define AVID_MAX_PEERS 16
typedef struct avid_peer_ptrs {
    const void* ptr[AVID_MAX_PEERS];
} avid_peer_ptrs_t;

/* Reduce-scatter + Adam in one pass over the shard [begin, begin + count) (both multiples of 4):
 *   g = grad_scale * sum_{r < world} grads->ptr[r][begin + i]   (grad_scale = 1 / world is DDP's average; P2P loads, rank order)
 *   torch.optim.Adam update (L2 weight decay, bias correction at `step`) of param_flat[begin + i], exp_avg[i], exp_avg_sq[i];
 * exp_avg / exp_avg_sq hold `count` elements (only the owner keeps the moments of its shard).  The hyper-parameters are doubles and
 * are rounded to fp32 the way torch.optim.Adam's kernels see them (beta and 1 - beta separately). */
int avid_adam_shard_step(float* param_flat, const avid_peer_ptrs_t* grads, int32_t world, float* exp_avg, float* exp_avg_sq, int64_t begin,
                         int64_t count, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                         void* stream);

/* All-gather of the updated parameters by P2P loads: param_flat[r * shard, (r + 1) * shard) <- params->ptr[r][same range] for
 * every r != rank (shard a multiple of 4; the flat buffers hold world * shard elements). */
int avid_pull_shards(float* param_flat, const avid_peer_ptrs_t* params, int32_t world, int32_t rank, int64_t shard, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVID_B200_H */
