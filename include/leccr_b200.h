/*
 * leccr_b200 -- C ABI of the B200-native dense cross-modal similarity stage of LECCR.
 *
 * The reference (LiJiaBei-7/leccr) has no FFI layer: its boundary for this path is four Python
 * callables.  The Python shims in leccr_b200/ keep those signatures and call
 * the entry points below through ctypes; any other host can bind the same symbols.  Citations are
 * relative to /root/reference/LECCR/.
 *
 * Conventions
 *  - plain C, no C++ types, no exceptions; every function returns LECCR_OK (0) or a negative code;
 *    leccr_strerror() names it, leccr_last_cuda_error() returns the captured CUDA error string.
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *    (including workspaces); nothing is retained after the call returns; no device allocation and no
 *    host synchronisation happen inside (calls are stream-ordered on `stream`, a cudaStream_t).
 *  - matrices are row-major; `ld` counts elements; 16-bit operands must be 16-byte aligned with
 *    ld % 8 == 0 (TMA requirement).
 *  - the library targets sm_100a only and refuses other devices (LECCR_ERR_ARCH). There is no CPU
 *    path.
 */
#ifndef LECCR_B200_H_
#define LECCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LECCR_OK 0
#define LECCR_ERR_ARG (-1)        /* bad shape / null pointer / unsupported option            */
#define LECCR_ERR_ALIGN (-2)      /* pointer or leading dimension violates the TMA alignment   */
#define LECCR_ERR_ARCH (-3)       /* current device is not compute capability 10.x             */
#define LECCR_ERR_CUDA (-4)       /* a CUDA runtime call failed (see leccr_last_cuda_error)    */
#define LECCR_ERR_WORKSPACE (-5)  /* workspace too small                                       */
#define LECCR_ERR_DRIVER (-6)     /* cuTensorMapEncodeTiled unavailable or failed              */
#define LECCR_ERR_NCCL (-7)       /* libnccl.so.2 unavailable or an NCCL call failed           */

/* element types of caller tensors */
#define LECCR_F32 0
#define LECCR_F16 1
#define LECCR_BF16 2
/* tensor-core operand formats (tcgen05 kind::f16) */
#define LECCR_FMT_F16 0
#define LECCR_FMT_BF16 1
/* operand layouts produced by leccr_prep */
#define LECCR_LAYOUT_HI 0    /* [hi]            K = D                                   */
#define LECCR_LAYOUT_X3_ROWS 1 /* [hi | lo | hi]  K = 3D  (rows role of the split product) */
#define LECCR_LAYOUT_X3_COLS 2 /* [hi | hi | lo]  K = 3D  (cols role)                      */
/* double_sim fusion modes */
#define LECCR_FUSE_NORM 1 /* video_Retrieval_caption_double_sim.py:178 */
#define LECCR_FUSE_RAW 2  /* image_Retrieval_caption.py:244-246        */

#define LECCR_STAT_WORDS 4 /* floats per tensor written by leccr_prep: max|hi row|, max|residual row|, max|x|, bad flag */
#define LECCR_TOPK_KP 16   /* list length behind the top-k / Recall decision (k <= 16) */
#define LECCR_RANK_CAP 10  /* ranks >= this are reported as lower bounds (Recall@1/5/10 only needs < 10) */

typedef void* leccr_stream_t; /* cudaStream_t */

const char* leccr_strerror(int code);
const char* leccr_last_cuda_error(void);
int leccr_abi_version(void);
/* LECCR_OK when the current CUDA device can run the kernels (cc 10.x). */
int leccr_check_device(void);
/* Measurement aid (bench.py roofline figure): while enabled, every tensor-core launch is bracketed with
 * CUDA events on its own stream (no synchronisation: it can stay on inside a timed loop; not under CUDA-graph
 * capture); read waits for the recorded launches and returns their accumulated device time and count. */
void leccr_profile_enable(int on);
int leccr_profile_read(double* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------
 * Operand preparation.  Replaces nothing in the reference by itself; it is the prologue of every
 * similarity GEMM (the reference multiplies fp32 tensors directly, models/xvlm.py:273,
 * image_Retrieval_caption.py:151).  `normalize` fuses F.normalize(x, dim=-1)
 * (models/xvlm.py:245-256).  src_dtype F32 casts; F16/BF16 sources are used in place by the GEMMs
 * and only need leccr_stats16.
 *   dst16     : [n][ld_dst] 16-bit, ld_dst >= D (layout HI) or 3D (X3), ld_dst % 8 == 0
 *   rn_hi/lo  : [n] row norms of the 16-bit operand / of its rounding residual (may be NULL)
 *   stats     : LECCR_STAT_WORDS floats, zero-initialised by the caller, max-combined
 * ------------------------------------------------------------------------------------------ */
int leccr_prep(const float* src, int64_t n, int D, int64_t ld_src, int normalize, int fmt, int layout,
               void* dst16, int64_t ld_dst, float* rn_hi, float* rn_lo, float* stats,
               leccr_stream_t stream);
/* leccr_prep for the two embedding sets of an evaluation in ONE launch (same D, format, layout, normalize). */
int leccr_prep_pair(const float* src0, int64_t n0, int64_t ld_src0, void* dst16_0, int64_t ld_dst0, float* rn_hi0,
                    float* rn_lo0, float* stats0, const float* src1, int64_t n1, int64_t ld_src1, void* dst16_1,
                    int64_t ld_dst1, float* rn_hi1, float* rn_lo1, float* stats1, int D, int normalize, int fmt,
                    int layout, leccr_stream_t stream);
/* The cast prologue fused with its all-gather (replaces AllGather.forward, models/xvlm.py:53-59, for the
 * contrastive operands): casts this rank's n rows and stores them into EVERY rank's gathered operand
 * buffer at [dst_row0 + i][dst_col0 ...] through peer pointers (NVLink).  dst_ptrs_dev: device array of
 * `world` base pointers of symmetric buffers (own included), leading dimension ld_dst elements.
 * The caller issues a cross-rank barrier before anyone reads the gathered buffers. */
int leccr_prep_push(const float* src, int64_t n, int D, int64_t ld_src, int normalize, int fmt,
                    void* const* dst_ptrs_dev, int world, int64_t dst_row0, int64_t dst_col0, int64_t ld_dst,
                    leccr_stream_t stream);
/* Same exchange for a block of 8-byte words (the idx labels, models/xvlm.py:285). */
int leccr_push_words(const void* src, int64_t n_words, void* const* dst_ptrs_dev, int world, int64_t dst_word0,
                     leccr_stream_t stream);
int leccr_stats16(const void* src16, int fmt, int64_t n, int D, int64_t ld_src, float* rn_hi,
                  float* rn_lo, float* stats, leccr_stream_t stream);
/* [n][D] -> [D][ld_dst] (ld_dst >= n, columns n..ld_dst zero-filled). */
int leccr_transpose16(const void* src16, int64_t n, int D, int64_t ld_src, void* dst16, int64_t ld_dst,
                      leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_sim_f32: S = scale * rows16 . cols16^T   (fp32 out, materialised)
 * Replaces: score_matrix_i2t = image_embeds @ text_embeds.t()   image_Retrieval_caption.py:151
 *           c_sim = caption_embeds.reshape(-1,d) @ text_embeds.T video_Retrieval_caption_double_sim.py:174
 * With X3 operands (K = 3D) the product is fp32-accurate; with HI operands it carries the 16-bit
 * rounding of the inputs.  scale_dev (optional device scalar) is multiplied into scale.
 * ------------------------------------------------------------------------------------------ */
int leccr_sim_f32(const void* rows16, int64_t ld_rows, const void* cols16, int64_t ld_cols, int64_t n_rows,
                  int64_t n_cols, int K, int fmt, float* S, int64_t ld_S, float scale,
                  const float* scale_dev, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_sim_topk: fused similarity + streaming per-row top-k (+ exact Recall ranks) for one or two
 * orientations in ONE tensor-core launch; the n_rows x n_cols matrix never reaches HBM.
 * Replaces: image_Retrieval_caption.py:151-163 (matmul + D2H) and :261-295 (np.argsort ranking).
 * ------------------------------------------------------------------------------------------ */
typedef struct leccr_topk_problem {
  /* 16-bit tensor-core operands (from leccr_prep, or the caller's own fp16/bf16 tensors) */
  const void* rows16;
  const void* cols16;
  int64_t ld_rows16, ld_cols16;
  int64_t n_rows, n_cols;
  /* outputs */
  float* topk_val; /* [n_rows][k] approximate scores, descending (ties: lower column first) */
  int32_t* topk_idx;
  /* exact Recall support (optional: gt_off == NULL disables everything below) */
  const int32_t* gt_off; /* CSR offsets [n_rows + 1] of each row's ground-truth columns */
  const int32_t* gt_ids;
  const void* rows_x; /* original (un-rounded) operands used for exact re-scoring */
  const void* cols_x;
  int64_t ld_rows_x, ld_cols_x;
  int x_dtype;              /* LECCR_F32 / F16 / BF16 */
  const float* rn_hi;       /* [n_rows] from leccr_prep / leccr_stats16 of the rows operand */
  const float* rn_lo;
  const float* col_stats;   /* LECCR_STAT_WORDS of the cols operand */
  int32_t* rank;            /* [n_rows]: min over GT of #{j : s_j > s_gt}; exact when < LECCR_RANK_CAP */
  int32_t* recall_counts;   /* [3] += #{rank < 1, 5, 10}  (zeroed by the caller) */
  float* gt_score;          /* [nnz] exact fp32 score of each ground-truth pair (may be NULL) */
} leccr_topk_problem;

size_t leccr_sim_topk_workspace(const leccr_topk_problem* probs, int n_prob, int tiles_per_chunk);
int leccr_sim_topk(const leccr_topk_problem* probs, int n_prob, int D, int fmt, int k,
                   int tiles_per_chunk /* 0 = auto */, void* workspace, size_t workspace_bytes,
                   leccr_stream_t stream);

/* leccr_sim_rank: Recall evaluation WITHOUT candidate lists -- exactly what the reference's itm_eval returns
 * (image_Retrieval_caption.py:261-317: the position of each row's best ground-truth column, Recall@1/5/10; no
 * top-k list leaves that function).  Same problem description as leccr_sim_topk (gt_off .. gt_score required,
 * topk_val / topk_idx ignored).  The best ground-truth score of every row is computed exactly first; the
 * tensor-core pass only COUNTS the scores that are definitely greater (16-bit-operand score > t + eps) and writes
 * the rare (row, column) pairs inside the band t +- eps, which are re-scored exactly afterwards; a row whose pairs
 * overflow takes the exact fallback.  rank[row] = number of columns scoring above the row's best ground truth:
 * exact when < LECCR_RANK_CAP, otherwise a lower bound >= LECCR_RANK_CAP (the same contract as leccr_sim_topk;
 * LECCR_RANK_CAP for rows without ground truth); recall_counts += #{rank < 1, 5, 10}. */
size_t leccr_sim_rank_workspace(const leccr_topk_problem* probs, int n_prob);
int leccr_sim_rank(const leccr_topk_problem* probs, int n_prob, int D, int fmt, void* workspace, size_t workspace_bytes,
                   leccr_stream_t stream);

/* Streamed evaluation: the columns of a problem arrive in windows (e.g. while the rest of the gallery is
 * still crossing PCIe).  Each call may run the tensor-core phase over ONE window (cols16 / n_cols describe
 * the window, col_begin its global position; the candidates go to list slots [sub_begin, sub_begin +
 * sub_count) of the problem's persistent workspace) and / or the finalize phase over everything collected
 * (then cols_x must address ALL columns and n_cols_total gives their number).  Per-row thresholds carry over
 * from window to window.  A problem whose rows are complete in one call uses all three phases at once;
 * the problems of one call share one tensor-core launch.  sub_total <= 8. */
#define LECCR_TOPK_INIT 1     /* first call of a problem: reset its workspace */
#define LECCR_TOPK_GEMM 2     /* similarity + candidate lists for the columns given */
#define LECCR_TOPK_FINALIZE 4 /* merge all slots: top-k, exact ranks, Recall counts */
#define LECCR_TOPK_LONG 8     /* set on EVERY call of a problem whose windows are long (> 8192 columns per slot,
                                 e.g. the gallery windows of a 1M-row search): filter-epilogue list shape */
typedef struct leccr_topk_stream {
  int32_t phases;
  int32_t sub_begin, sub_count, sub_total;
  int64_t col_begin;
  int64_t n_cols_total; /* finalize without a tensor-core phase: total number of columns (else 0) */
  void* workspace;      /* leccr_sim_topk_stream_workspace(n_rows, sub_total) bytes, kept from INIT to FINALIZE */
  size_t workspace_bytes;
} leccr_topk_stream;
size_t leccr_sim_topk_stream_workspace(int64_t n_rows, int sub_total);
int leccr_sim_topk_stream(const leccr_topk_problem* probs, const leccr_topk_stream* streams, int n_prob, int D,
                          int fmt, int k, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_infonce_fwd / leccr_infonce_bwd: symmetric InfoNCE over all-gathered embeddings.
 * Replaces: XVLMBase.get_contrastive_loss  models/xvlm.py:260-292 (forward: :273-292; backward:
 * autograd of the same lines followed by AllGather.backward :62-67, which keeps the local rows).
 *   a16, b16     : [n][D] 16-bit all-gathered image / text operands (leccr_prep, layout HI)
 *   idx          : [n] int64 labels (idx_all, :285) or NULL for the arange labels of :277
 *   temp         : device scalar self.temp
 *   out          : [6] loss, dloss/dtemp, loss_i2t, loss_t2i, dloss_i2t/dtemp, dloss_t2i/dtemp
 *   lse2 / rcnt  : [2][n] per-row log2-domain log-sum-exp and 1/|positives| (saved for backward)
 * Backward (local rows [row_begin, row_begin + row_count) only):
 *   aT16, bT16   : ignored (may be null): the gradient products dA = G' B, dB = G'^T A read the gathered rows
 *                  as MN-major tensor-core operands (the contraction index is the row index of a16 / b16, no
 *                  transposed copy).  Only with LECCR_BWD_MN=0 (development comparison) they are the [D][ldT]
 *                  transposed operands (leccr_transpose16)
 *   grad_out     : device scalar dL/dloss
 *   dA, dB       : [row_count][D] fp32 (overwritten)
 * ------------------------------------------------------------------------------------------ */
size_t leccr_infonce_fwd_workspace(int64_t n, int tiles_per_chunk);
int leccr_infonce_fwd(const void* a16, const void* b16, int64_t ld16, const int64_t* idx, int64_t n, int D,
                      int fmt, const float* temp, float* out, float* lse2, float* rcnt,
                      int tiles_per_chunk, void* workspace, size_t workspace_bytes,
                      leccr_stream_t stream);
size_t leccr_infonce_bwd_workspace(int64_t n, int64_t row_count, int D);
int leccr_infonce_bwd(const void* a16, const void* b16, int64_t ld16, const void* aT16, const void* bT16,
                      int64_t ldT, const int64_t* idx, int64_t n, int D, int fmt, const float* temp,
                      const float* lse2, const float* rcnt, int64_t row_begin, int64_t row_count,
                      const float* grad_out, float* dA, float* dB, void* workspace,
                      size_t workspace_bytes, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Ranking of a materialised fp32 score matrix S [R][C] (the drop-in itm_eval).
 * Replaces: itm_eval  image_Retrieval_caption.py:261-295 == video_..._double_sim.py:194-230.
 *   rank_rows: rank[r] = min_{g in gt(r)} #{c : S[r][c] > S[r][g]}          (i2t, :265-278)
 *   rank_cols: rank[c] = min_{g in gt(c)} #{r : S[r][c] > S[g][c]}          (t2i, :288-290, read
 *              from the same matrix: the reference's t2i matrix is the transpose view, :152)
 *   recall_counts: counts[0..2] += #{rank < 1, 5, 10}                        (:281-283, :293-295)
 * scratch for rank_cols: int32[nnz], zeroed by the caller.
 * ------------------------------------------------------------------------------------------ */
int leccr_rank_rows(const float* S, int64_t ld, int64_t R, int64_t C, const int32_t* gt_off,
                    const int32_t* gt_ids, int32_t* rank, leccr_stream_t stream);
int leccr_rank_cols(const float* S, int64_t ld, int64_t R, int64_t C, const int32_t* gt_off,
                    const int32_t* gt_ids, int32_t* scratch_nnz, int32_t* rank, leccr_stream_t stream);
int leccr_recall_counts(const int32_t* rank, int64_t n, int32_t* counts, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_double_sim_fuse: in place S <- w1 * f(S) + w2 * f(max_n Cn)
 * Replaces: norm_score + fusion  video_Retrieval_caption_double_sim.py:87-91,175-178 (mode NORM,
 * f = (x - max x) / (max x - min x) over the whole matrix) and image_Retrieval_caption.py:239-246
 * (mode RAW, f = identity).
 *   S  : [numel] fp32 text-video similarity;  Cn : [n_cap][numel] caption similarities
 *   Cmax : [numel] scratch receiving max_n Cn;  mm : 4 x uint32 scratch
 * ------------------------------------------------------------------------------------------ */
int leccr_double_sim_fuse(float* S, const float* Cn, int n_cap, int64_t numel, float* Cmax, uint32_t* mm,
                          float w1, float w2, int mode, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_double_sim_topk: the double_sim evaluation fused into the tensor-core epilogue -- no N x M matrix.
 * Replaces: video_Retrieval_caption_double_sim.py:170-179 (S = V T^T, C = max_n caption_n T^T,
 * alpha * norm_score(S) + (1 - alpha) * norm_score(C), both directions; norm_score :87-91) followed by itm_eval
 * (:194-247), and the raw variant image_Retrieval_caption.py:239-246 (mode RAW).
 *   vc16 : [G * n_vid][K] 16-bit, the videos and their caption queries INTERLEAVED: row G i = video i, rows
 *          G i + 1 .. G i + n_cap = caption_1..n (i), rows above repeat the last caption.  G in {2, 4, 8} > n_cap.
 *          Split-precision layout LECCR_LAYOUT_X3_COLS (K = 3 D) for fp32-faithful scores, or HI (K = D).
 *   t16  : [n_txt][K] texts, layout LECCR_LAYOUT_X3_ROWS (or HI).  Both contiguous (ld = K).
 *   w1, w2 : alpha, 1 - alpha.   txt_gt: [n_txt] ground-truth video of every text (txt2img), or NULL: no ranks.
 *   vid_gt_off / vid_gt_ids : CSR of img2txt; it must be the inverse map of txt_gt (every dataset of the
 *          reference builds them together, dataset/retrieval_dataset_video.py:201-219).
 *   topk_*_vc : [G * n_vid][k] per-VC-row lists (row G i = video i's top-k texts) or NULL (orientation A skipped);
 *   topk_*_txt: [n_txt][k] per-text top-k videos or NULL.   rank_vid [n_vid], rank_txt [n_txt]: number of scores
 *   strictly above the row's best ground-truth score;  recall_counts [6]: #{rank < 1, 5, 10} i2t then t2i.
 * Pass 1 (tensor cores, nothing stored): min / max of S and max_n C_n, scores at the ground truth.  Pass 2
 * (recompute): fuse with the reference's fp32 operation order, count, top-k lists.
 * ------------------------------------------------------------------------------------------ */
size_t leccr_double_sim_topk_workspace(int64_t n_vid, int64_t n_txt, int G);
int leccr_double_sim_topk(const void* vc16, const void* t16, int64_t n_vid, int64_t n_txt, int K, int fmt, int G,
                          int n_cap, float w1, float w2, int mode, const int32_t* txt_gt, const int32_t* vid_gt_off,
                          const int32_t* vid_gt_ids, int k, float* topk_val_vc, int32_t* topk_idx_vc,
                          float* topk_val_txt, int32_t* topk_idx_txt, int32_t* rank_vid, int32_t* rank_txt,
                          int32_t* recall_counts, void* workspace, size_t workspace_bytes, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_normalize_fwd / leccr_normalize_bwd: row-wise L2 normalisation as a training op.
 * Replaces: the F.normalize(..., dim=-1) of XVLMBase.get_features, models/xvlm.py:245-256 and
 * models/xvlm_video.py:264-277 (the projection in front of it stays torch's Linear).
 *   x   : [n][ld_x] fp32 projected features;  y: [n][ld_y] fp32 = x / max(||x||, 1e-12)
 *   inv : [n] out, 1 / max(||x||, 1e-12), NEGATED for rows whose norm was clamped (saved for the backward)
 *   y16 : optional [n][ld_y16] 16-bit copy of y in `fmt` (the similarity stage's operand, same pass)
 *   backward: dx = inv * (g - y (y . g)); clamped rows: dx = g * inv (torch's clamp_min semantics)
 * ------------------------------------------------------------------------------------------ */
int leccr_normalize_fwd(const float* x, int64_t n, int D, int64_t ld_x, float* y, int64_t ld_y, float* inv,
                        void* y16, int64_t ld_y16, int fmt, leccr_stream_t stream);
int leccr_normalize_bwd(const float* y, int64_t ld_y, const float* inv, const float* g, int64_t ld_g, int64_t n, int D,
                        float* dx, int64_t ld_dx, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Cross-rank exchange over peer memory (one node, NVLink / NVSwitch).  The pointer
 * tables are device arrays of `world` device pointers into peer-mapped (symmetric) buffers, own rank
 * included; the Python shim obtains them from torch.distributed._symmetric_memory.
 *
 * leccr_peer_barrier: stream-ordered barrier across the ranks.  flag_ptrs_dev[p] -> rank p's block of
 *   64 uint32 words (zero-initialised once: `world` flags, word 48 an epoch counter); `epoch` must grow by one
 *   per barrier on every rank, or be 0 on every call: the kernel then takes the next epoch from the counter
 *   word itself, the launch has no per-call argument and can be captured in a CUDA graph.
 *   The wait is bounded by wall time: LECCR_PEER_TIMEOUT_S seconds (default 600, 0 = for ever), then the
 *   kernel reports the missing peer and traps.
 *   Orders all earlier peer stores of this stream before all later work of the peers' streams.
 * leccr_topk_merge_peers: the exchange step of the row-partitioned gallery (replaces nothing in the
 *   reference, which ranks on one CPU; north_star layout).  Rank p published its per-query local top-k
 *   ([Q][k_in] fp32 descending / int32 local columns, ties by lower column) in its peer-visible buffer;
 *   this call pulls the lists of queries [q_begin, q_begin + q_count) from all ranks and merges them in
 *   ONE kernel: out [q_count][k_out], global column = local + col_offset_host[p].
 * ------------------------------------------------------------------------------------------ */
int leccr_peer_barrier(uint32_t* const* flag_ptrs_dev, int world, int rank, uint32_t epoch, leccr_stream_t stream);
/* leccr_memcpy_peer_async: stream-ordered copy between device pointers, either of which may be a peer-mapped
 * address (copy engines over NVLink, no SMs): the gallery windows a rank uploaded are pushed to the ranks
 * that share its gallery part while the tensor cores rank the windows that have arrived (the reference moves
 * every feature through the host: image_Retrieval_caption.py:163). */
int leccr_memcpy_peer_async(void* dst, const void* src, size_t bytes, leccr_stream_t stream);
int leccr_topk_merge_peers(const float* const* val_ptrs_dev, const int32_t* const* idx_ptrs_dev, int world, int k_in,
                           int64_t q_begin, int64_t q_count, const int64_t* col_offset_host, int k_out,
                           float* out_val, int32_t* out_idx, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_itc_forward / leccr_itc_backward: the whole get_contrastive_loss step in two calls
 * (models/xvlm.py:260-292 forward incl. its three AllGather calls :271-272,285; backward = autograd of the
 * same lines + AllGather.backward :62-67).  They issue exactly the launches of leccr_prep(_push),
 * leccr_peer_barrier, leccr_infonce_fwd / leccr_transpose16, leccr_infonce_bwd from C++ so the host cost of a
 * training step is two library calls.
 *   image_feat, text_feat : this rank's [B][D] fp32 rows;  idx: [B] int64 or NULL
 *   world == 1 : rows_ptrs_dev .. local_idx are ignored
 *   world == 1 : rows_ptrs_dev .. local_slot_bytes are ignored; 3 launches: cast, tensor-core pass, finalize
 *   world  > 1 : rows_ptrs_dev[p] -> rank p's peer-mapped slot: [n][2D] 16-bit rows followed (at idx_all -
 *                both16 bytes) by [n] int64 labels -- the private buffer's own layout; idx_ptrs_dev[p] -> the
 *                labels part; flag_ptrs_dev / epoch as in leccr_peer_barrier; local_slot = this rank's slot,
 *                local_slot_bytes the bytes to copy into both16 (rows, + labels when idx != NULL).
 *                The caller alternates two slots.  5 launches: push, barrier, copy, tensor-core pass, finalize.
 *   stat_ptrs_dev != NULL (world > 1): STRIP forward.  Every row's log-sum-exp, positives and E_softmax[z]
 *                depend on that row alone (:279-290 are row-wise), so the rank runs the tensor-core pass for ITS
 *                B rows of both orientations only (1 / world of the work) and the per-row statistics are
 *                exchanged: stat_ptrs_dev[p] -> rank p's peer-mapped statistics slot, float lse2[2][n] |
 *                float rcnt[2][n] | double partial[world][4]; local_stat_slot = this rank's.  A second barrier
 *                (epoch + 1, or 0 again in counter mode: the caller advances its epoch by TWO per call) orders the
 *                exchange.  7 launches:
 *                push, barrier, copy, tensor-core pass, finalize + push, barrier, reduce.
 *   both16 : out, private [n][2D] 16-bit gathered operands [image | text] (saved for the backward)
 *   idx_all: out, [n] int64 (when idx != NULL);  out/lse2/rcnt as in leccr_infonce_fwd
 *   backward: dA, dB [row_count][D] fp32, dtemp scalar = grad_out * out[1] (may be NULL);
 *             one_directional: gradient of out[2] = loss_i2t alone (caption_vision_loss,
 *             models/model_retrieval_caption.py:141), dtemp = grad_out * out[4]
 * ------------------------------------------------------------------------------------------ */
size_t leccr_itc_fwd_workspace(int64_t n, int tiles_per_chunk);
int leccr_itc_forward(const float* image_feat, int64_t ld_img, const float* text_feat, int64_t ld_txt,
                      const int64_t* idx, int64_t B, int D, int fmt, int rank, int world,
                      void* const* rows_ptrs_dev, void* const* idx_ptrs_dev, uint32_t* const* flag_ptrs_dev,
                      uint32_t epoch, const void* local_slot, size_t local_slot_bytes, void* const* stat_ptrs_dev,
                      const void* local_stat_slot, void* both16,
                      int64_t* idx_all, const float* temp, float* out, float* lse2, float* rcnt,
                      void* workspace, size_t workspace_bytes, leccr_stream_t stream);
size_t leccr_itc_bwd_workspace(int64_t n, int64_t row_count, int D);
int leccr_itc_backward(const void* both16, const int64_t* idx_all, int64_t n, int D, int fmt, const float* temp,
                       const float* lse2, const float* rcnt, const float* out, int64_t row_begin, int64_t row_count,
                       const float* grad_out, float* dA, float* dB, float* dtemp, int one_directional, void* workspace,
                       size_t workspace_bytes, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_caploss_fwd / leccr_caploss_bwd: caption contrastive loss of the LOCAL batch (no gather).
 * Replaces: RetrievalModel.get_caption_contrastive_loss  models/model_retrieval_caption.py:145-152
 * (== models/video_model_retrieval_caption.py:171-178): sim = caption.reshape(n*B, d) @ text.T,
 * logits = max_n sim / temp, arange labels, symmetric cross entropy; backward = autograd of the same lines
 * (the max routes each gradient to its arg-max caption query, first index on ties).
 *   cap16 : [n_cap * B][K] 16-bit caption queries, row a * B + i = query a of sample i
 *   txt16 : [B][K] 16-bit text features.  Forward: K = 3D split-precision operands (leccr_prep layouts
 *           X3_ROWS / X3_COLS) so the arg max is the fp32 one; backward: the plain [.][D] halves of the same
 *           buffers (ld = 3D).
 *   out   : [2] loss, d loss / d temp;  L: [B][B] fp32 max_n sim;  amax: [B][B] arg max;  stats: [4][B]
 *           (row / column log-sum-exp in log2 units, row / column E_softmax[z]) -- saved for the backward
 *   dcap  : [n_cap * B][D], dtxt: [B][D] fp32 (overwritten), dtemp scalar = grad_out * out[1] (may be NULL)
 * ------------------------------------------------------------------------------------------ */
size_t leccr_caploss_fwd_workspace(int n_cap, int64_t B);
int leccr_caploss_fwd(const void* cap16, int64_t ld_cap, const void* txt16, int64_t ld_txt, int n_cap, int64_t B, int K,
                      int fmt, const float* temp, float* out, float* L, uint8_t* amax, float* stats, void* workspace,
                      size_t workspace_bytes, leccr_stream_t stream);
size_t leccr_caploss_bwd_workspace(int n_cap, int64_t B, int D);
int leccr_caploss_bwd(const float* L, const uint8_t* amax, const float* stats, const void* cap16, int64_t ld_cap,
                      const void* txt16, int64_t ld_txt, int n_cap, int64_t B, int D, int fmt, const float* temp,
                      const float* out, const float* grad_out, float* dcap, float* dtxt, float* dtemp, void* workspace,
                      size_t workspace_bytes, leccr_stream_t stream);

/* leccr_topk_dense: top-k (k <= 16) of every row -- or, with by_columns, every column -- of a MATERIALISED
 * fp32 score matrix S [R][ld]; used for the double_sim fusion (video_Retrieval_caption_double_sim.py:178-179),
 * whose matrix is materialised.  out: [R or C][k], score descending, ties by lower index, -1 / -inf padding. */
int leccr_topk_dense(const float* S, int64_t ld, int64_t R, int64_t C, int by_columns, int k, float* out_val,
                     int32_t* out_idx, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * NCCL collectives behind the C ABI, for hosts without torch.distributed and for groups that span nodes
 * (inside one NVLink node the peer-memory entries above are used instead).  Replaces the transport of
 * AllGather.forward, models/xvlm.py:53-59 (dist.all_gather + torch.cat -> one ncclAllGather into the
 * gathered buffer); the top-k lists of a row-partitioned gallery are exchanged with the same call and merged
 * by leccr_topk_merge_peers over a table of pointers into the gathered buffer.
 *   leccr_comm_unique_id : rank 0 creates the 128-byte id and ships it to the others (any side channel)
 *   leccr_comm_init      : collective; one communicator per process / GPU (current device)
 *   leccr_allgather      : recv[r * bytes_per_rank ...] = rank r's send, stream-ordered
 * libnccl.so.2 is resolved at run time (LECCR_ERR_NCCL when it is missing).
 * ------------------------------------------------------------------------------------------ */
typedef struct leccr_nccl_id { char internal[128]; } leccr_nccl_id;
int leccr_comm_unique_id(leccr_nccl_id* id_host);
int leccr_comm_init(const leccr_nccl_id* id_host, int rank, int world, void** comm);
int leccr_comm_destroy(void* comm);
int leccr_allgather(void* comm, const void* send, void* recv, size_t bytes_per_rank, leccr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * leccr_dstl_fwd / leccr_dstl_bwd: the distillation loss on ALL-GATHERED tensors (N = world * B rows).
 * Replaces: RetrievalModel.dstl_loss  models/model_retrieval_caption.py:99-116 after its four allgathers
 * (:95-98; the gathers stay with the host): logits_tv = text_t image^T, labels = softmax_rows(alpha *
 * norm(text_s image^T) + (1 - alpha) * norm(max_n caption_n text_s^T)) (norm_score :87-90), loss =
 * kl_div(log_softmax(logits_tv), labels, batchmean).  Backward = autograd of the same lines followed by
 * AllGather.backward (models/xvlm.py:62-67): the local rows of d image and d text_t (labels are detached).
 *   16-bit operands with K = 3D split precision (leccr_prep): tt16 / ts16_rows / cap16 in layout X3_ROWS,
 *   img16 / ts16_cols in layout X3_COLS; cap16 is [n_cap * N][K], row a * N + i = query a of sample i
 *   out: [1] loss;  Fm, TV: [N][N] fp32 label logits / text_t-image logits;  lse: [2][N]  (saved for backward)
 *   backward: img16 / tt16 = the plain [N][D] halves of the same buffers (equal ld), dimg / dtt [row_count][D]
 * ------------------------------------------------------------------------------------------ */
size_t leccr_dstl_fwd_workspace(int n_cap, int64_t N);
int leccr_dstl_fwd(const void* tt16, int64_t ld_tt, const void* ts16_rows, int64_t ld_tsr, const void* ts16_cols,
                   int64_t ld_tsc, const void* img16, int64_t ld_img, const void* cap16, int64_t ld_cap, int n_cap,
                   int64_t N, int K, int fmt, float alpha, float* out, float* Fm, float* TV, float* lse, void* workspace,
                   size_t workspace_bytes, leccr_stream_t stream);
size_t leccr_dstl_bwd_workspace(int64_t N, int64_t row_count, int D);
int leccr_dstl_bwd(const float* Fm, const float* TV, const float* lse, const void* img16, int64_t ld_img, const void* tt16,
                   int64_t ld_tt, int64_t N, int D, int fmt, int64_t row_begin, int64_t row_count, const float* grad_out,
                   float* dimg, float* dtt, void* workspace, size_t workspace_bytes, leccr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LECCR_B200_H_ */
