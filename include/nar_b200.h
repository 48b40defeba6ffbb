/*
 * nar_b200.h - C ABI of the B200-native NAR (CHAMELEON next-article recommendation)
 * training hot path.  libnar_b200.so exports exactly these symbols.
 *
 * The reference (gabrielspmoreira/chameleon_recsys @ 2e50af5) has NO native / FFI layer:
 * its boundary is the TF-Estimator Python contract (nar_module/nar/nar_trainer_gcom.py:234-332
 * model_fn, nar_module/nar/datasets.py:166-179 input_fn) and every "kernel" is a TensorFlow
 * 1.12 library op.  Each entry point below therefore cites the reference op group it
 * replaces (file:line of the TF call site) instead of a pre-existing FFI declaration.
 *
 * Conventions: extern "C"; plain pointers and sizes; every pointer is a DEVICE pointer
 * unless it says "host"; the caller owns every buffer; `stream` is a cudaStream_t passed
 * as void*; functions return 0 on success, a negative nar_status, or a positive
 * cudaError_t; nothing throws; no hidden global state beyond the opaque nar_ctx.
 * There is no CPU fallback anywhere: without a CUDA device every call fails.
 */
#ifndef NAR_B200_H
#define NAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAR_ABI_VERSION 2

typedef enum {
  NAR_OK = 0,
  NAR_ERR_INVALID = -1,        /* bad argument (null pointer, misaligned ld, size limit) */
  NAR_ERR_UNSUPPORTED = -2,    /* valid request the library does not implement */
  NAR_ERR_NO_DEVICE = -3,      /* no sm_100 device / driver entry point missing */
  NAR_ERR_WORKSPACE = -4       /* workspace too small */
} nar_status;

typedef struct nar_ctx nar_ctx;

/* ---- context ------------------------------------------------------------------------- */
int  nar_abi_version(void);
/* sizeof of an ABI struct as this library was compiled: 0 nar_feature_plan, 1 nar_model_cfg, 2 nar_step_io,
 * 3 nar_row_layout, 4 nar_gemm_epilogue, 5 nar_segment (bindings check their mirrors against it); -1 otherwise */
int  nar_abi_struct_size(int which);
const char* nar_status_string(int status);
int  nar_ctx_create(int device, nar_ctx** out);
int  nar_ctx_destroy(nar_ctx* ctx);

/* ---- feature-row gather  (replaces tf.nn.embedding_lookup nar_model.py:948 (ACR, frozen),
 *      :918 (trainable item embedding), tf.gather :929/:1067/:1095/:1138, tf.one_hot :734,
 *      small embedding_lookup :741, recency/novelty normalisation :996-1193, the concat
 *      :332/:992 and scale_center_features :905).  One output row per (position, item)
 *      pair; written straight into the GEMM A operand.                                     */
typedef enum {
  NAR_SEG_CTX_OHE = 0, NAR_SEG_CTX_EMBED = 1, NAR_SEG_CTX_NUM = 2, NAR_SEG_CTX_ZERO = 3,
  NAR_SEG_META_OHE = 4, NAR_SEG_META_EMBED = 5, NAR_SEG_META_NUM = 6,
  NAR_SEG_ACR = 7, NAR_SEG_ITEM_EMB = 8, NAR_SEG_RECENCY = 9, NAR_SEG_NOVELTY = 10
} nar_seg_kind;

typedef struct {
  int32_t kind;        /* nar_seg_kind */
  int32_t col;         /* first column in the output row */
  int32_t width;       /* columns written */
  int32_t card;        /* categorical cardinality (OHE / EMBED) */
  int32_t src;         /* index into ctx_int / ctx_float / meta pointer arrays */
  int32_t ld;          /* leading dimension of `table` (floats) */
  const float* table;  /* embedding table (EMBED, ACR, ITEM_EMB) */
  float* grad;         /* gradient table for trainable embeddings (backward only) */
} nar_segment;

#define NAR_MAX_SEGMENTS 24
#define NAR_MAX_SRC 16
#define NAR_MAX_COLS 1024

typedef struct {
  int32_t n_segments;
  int32_t row_ld;                       /* floats per output row (Fp) */
  nar_segment seg[NAR_MAX_SEGMENTS];
  const int64_t* ctx_int[NAR_MAX_SRC];  /* [B*T] int64 context ids per feature */
  const float*   ctx_float[NAR_MAX_SRC];/* [B*T] float context values per feature */
  const int64_t* meta[NAR_MAX_SRC];     /* [V] int64 metadata per feature */
  const int64_t* created_at_ts;         /* [V] int64 ms */
  const float*   pop_norm;              /* [V] articles_recent_pop_norm */
  const float*   gamma;                 /* [Fp] scale  (nar_model.py:891) */
  const float*   beta;                  /* [Fp] centre (nar_model.py:895) */
  const float*   stats;                 /* [3][8] normalisation stats (input / positive / negative rows), see nar_feature_stats */
  float log_base_recency;               /* elapsed_days_smooth_log_base */
  float log_base_novelty;               /* popularity_smooth_log_base */
  /* column map: col_seg[c] = index into seg[] of the segment that owns output column c, 255 = padding.
   * narrow_begin/end: the column ranges NOT covered by the wide (ACR / item-embedding) segments. */
  int32_t n_narrow;
  int32_t narrow_begin[4];
  int32_t narrow_end[4];
  uint8_t col_seg[NAR_MAX_COLS];
} nar_feature_plan;

/* Which rows a row list holds.  Rows [0, n_input) are clicked items (reference timestamp = event_timestamp[row_pos],
 * normalisation statistics group 0, nar_model.py:328); the others use max_ts (:343, :356) and are either
 *   n_cand > 0 : groups of n_cand rows per position, the positive first (group 1) then its negatives (group 2), or
 *   n_cand == 0: n_positive positive rows (group 1) followed by negative rows (group 2) - the base rows of the
 *                per-unique-id CAR layer 1 (nar_build_base_rows).
 * Rows >= n_full carry ITEM features only: their context columns (internal column >= ctx_col0) are written as 0 and
 * skipped by the backward pass.  n_full >= n_rows: every row is a full row.                                         */
typedef struct {
  int64_t n_rows, n_input, n_cand, n_positive, n_full, ctx_col0;
} nar_row_layout;

/* rows: row_pos[r] = flat index b*T+t of the position that owns row r (context features,
 * reference timestamp), row_item[r] = article id.                                           */
int nar_gather_features(nar_ctx* ctx, const nar_feature_plan* plan /*host*/,
                        const int32_t* row_pos, const int64_t* row_item, const nar_row_layout* rows /*host*/,
                        const int64_t* event_timestamp /*[B*T]*/, const int64_t* max_ts /*[1]*/,
                        float* out /*[n_rows,row_ld]*/, void* stream);

/* backward of the same: d_gamma += sum_r dX*raw, d_beta += sum_r dX, trainable embedding
 * grads += dX*gamma scattered by id (IndexedSlices scatter-add, nar_model.py:918/:741).     */
int nar_gather_features_bwd(nar_ctx* ctx, const nar_feature_plan* plan /*host*/,
                            const int32_t* row_pos, const int64_t* row_item, const nar_row_layout* rows /*host*/,
                            const int64_t* event_timestamp, const int64_t* max_ts,
                            const float* d_out /*[n_rows,row_ld]*/, float* d_gamma, float* d_beta, void* stream);

/* row lists of one step.  The L valid positions (pos_idx[l] = b*T+t, session-major) produce
 * n_rows = L + L*(1+K) rows: input rows [0,L) = clicked items (nar_model.py:328), then for each
 * position its candidates contiguously: positive label_next_item (:343) followed by its K
 * negatives (:356).                                                                        */
int nar_build_rows(const int32_t* pos_idx, int64_t L, const int64_t* item_clicked, const int64_t* label_next_item,
                   const int64_t* negatives /*[B*T,K]*/, int64_t K, int32_t* row_pos, int64_t* row_item, void* stream);

/* base rows of the per-unique-id CAR layer 1 (csrc/car.cu): n_base = 2L + U rows = the L clicked items, the L positives,
 * then one ITEM-ONLY row per entry of the step's unique-negative table (unique_items / n_unique as returned by
 * nar_sample_negatives_uidx; U = table capacity K*20 plus one trailing slot for the padding negative, id 0).  Also
 * writes the inverse map Mt [U, ld_mt] uint16: Mt[u][l] = k+1 when position l drew unique entry u as its k-th
 * negative (neg_uidx [B*T, K]), 0 otherwise - what nar_car_segsum walks to sum gradients in a fixed order.         */
int nar_build_base_rows(const int32_t* pos_idx, int64_t L, const int64_t* item_clicked, const int64_t* label_next_item,
                        const int64_t* unique_items, const int32_t* n_unique /*[1] device*/, int64_t U,
                        const int32_t* neg_uidx, int64_t K, int32_t* base_pos /*[2L+U]*/, int64_t* base_item /*[2L+U]*/,
                        uint16_t* Mt, int64_t ld_mt, void* stream);

/* normalisation statistics of recency / novelty over the first n_norm nonzero buffer entries
 * (nar_model.py:1062-1089, :1150-1193, :1011-1039).  stats[g][8], g = 0 input / 1 positive /
 * 2 negative rows: {rec_mean, rec_std, rec_zmin, rec_zmax, nov_mean, nov_std, nov_zmin, nov_zmax};
 * all three groups are equal unless the buffer is empty (first batch), where each group
 * uses its own non-padded rows (the tf.cond at nar_model.py:1082 / :1179): pass the rows.    */
int nar_feature_stats(nar_ctx* ctx, const int64_t* buffer, int64_t buf_len, int64_t n_norm,
                      const int64_t* created_at_ts, const float* pop_norm, const int64_t* max_ts,
                      float log_base_recency, float log_base_novelty,
                      const int32_t* row_pos, const int64_t* row_item, int64_t n_rows, int64_t n_input,
                      int64_t n_cand, const int64_t* event_timestamp,
                      float* stats /*[24]*/, void* stream);

/* plain row gather / scatter-add used by the parity tests and the roofline micro-benchmark
 * (tf.nn.embedding_lookup nar_model.py:948 and its IndexedSlices gradient :918).           */
int nar_gather_rows_f32(const float* table, int64_t n_table_rows, int64_t ld, int width,
                        const int64_t* ids, int64_t n, float* out, int64_t ld_out, void* stream);
int nar_scatter_add_rows_f32(float* table, int64_t n_table_rows, int64_t ld, int width,
                             const int64_t* ids, int64_t n, const float* src, int64_t ld_src, void* stream);

/* ---- dense contraction (replaces every tf.layers.Dense nar_model.py:375-473 and the
 *      UGRNN input projection :1317 -> Eigen/MKL or cuBLAS sgemm in the reference).
 *      D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ), TMA-fed tcgen05.mma kind::tf32, fp32
 *      accumulation in TMEM.  a_kmajor: A(m,k) = A[m*lda + k] else A[k*lda + m];
 *      b_kmajor: B(n,k) = B[n*ldb + k] else B[k*ldb + n].                                  */
typedef enum { NAR_ACT_NONE = 0, NAR_ACT_LEAKY_RELU = 1, NAR_ACT_TANH = 2 } nar_act;

typedef struct {
  const float* bias;      /* [N] added before the activation, or NULL */
  int32_t act;            /* nar_act applied to acc+bias */
  int32_t dact;           /* nar_act whose DERIVATIVE (evaluated from the forward OUTPUT aux) multiplies the result */
  const float* aux;       /* [M,N] forward output of the layer being differentiated (dact != NONE) */
  int64_t ld_aux;
  int32_t accumulate;     /* 1: D += result with atomics (required when split_k > 1) */
  int32_t split_k;        /* >=1 */
  int32_t precision;      /* 1 = TF32, 3 = 3xTF32 (error-compensated, ~fp32 accuracy) */
  const float* b_lo;      /* precision 3 only, optional: x - tf32_trunc(x) of operand B, same shape / ld as B (see
                             nar_tf32_lo; the weights' lo plane is maintained by nar_adam_tf).  NULL: split B in-kernel */
  const void* b_bf16;     /* precision 4 (bf16x3: bf16 hi + lo pieces on the kind::f16 path, fp32 accumulate; A must be K-major
                             fp32): operand B as the pre-split transposed plane written by nar_pack_bf16x3 - [N, ld_bf16]
                             bf16, row n = per block of 32 k the 32 hi values then the 32 lo values; B / ldb are ignored */
  int64_t ld_bf16;        /* elements per row of b_bf16 (>= ceil(K/32)*64, multiple of 8) */
} nar_gemm_epilogue;

/* bf16x3 weight planes for n matrices in one launch: W[i] [K[i], N[i]] fp32 (row stride ldw[i], i.e. stored [in, out]) ->
 * out[i] [N[i], ld_out[i]] bf16 as nar_gemm_epilogue.b_bf16 describes (zero padded to whole 32-k blocks).
 * descs_dev: caller-owned device scratch of >= 32*32 bytes the call keeps its table in.                               */
int nar_pack_bf16x3(const float* const* W /*host array*/, void* const* out /*host array*/, const int32_t* K, const int32_t* N,
                    const int32_t* ldw, const int32_t* ld_out, int n, void* descs_dev, void* stream);

int nar_gemm_tf32(nar_ctx* ctx, int64_t M, int64_t N, int64_t K,
                  const float* A, int64_t lda, int a_kmajor,
                  const float* B, int64_t ldb, int b_kmajor,
                  float* D, int64_t ldd, const nar_gemm_epilogue* epi /*host*/, void* stream);

/* ---- session RNN (replaces tf.contrib.rnn.UGRNNCell in MultiRNNCell / dynamic_rnn,
 *      nar_model.py:1308-1342).  Rows are the valid positions only, grouped by session:
 *      session b owns rows [sess_off[b], sess_off[b+1]).  gx = x*Wx + b for all rows
 *      (nar_gemm_tf32), gate cols [0,Hp), candidate cols [Hp,2Hp).                          */
int nar_ugrnn_fwd(nar_ctx* ctx, const float* gx /*[L,2Hp]*/, const float* Wh /*[Hp,2Hp]*/,
                  const int32_t* sess_off /*[B+1]*/, int64_t B, int64_t Hp,
                  float* h_out /*[L,Hp]*/, float* gate /*[L,Hp]*/, float* cand /*[L,Hp]*/, void* stream);
int nar_ugrnn_bwd(nar_ctx* ctx, const float* d_hout /*[L,Hp]*/, const float* h_out, const float* gate,
                  const float* cand, const float* WhT /*[2Hp,Hp]*/, const int32_t* sess_off, int64_t B,
                  int64_t Hp, float* d_gx /*[L,2Hp]*/, float* h_prev /*[L,Hp]*/, void* stream);

/* GRU recurrence (tf.nn.rnn_cell.GRUCell; rnn_cell='gru'): gx [L,3Hp] = x*Wxg + bg | x*Wxc + bc (r | u | c pre-activations
 * of the input), Whg [Hp,2Hp], Whc [Hp,Hp]:  [r,u] = sigmoid(gx_ru + h*Whg) ; c = tanh(gx_c + (r*h)*Whc) ;
 * h' = u*h + (1-u)*c.  Outputs per row: state h_out, gates r / u, candidate c, rh = r * (state entering the step).      */
int nar_gru_fwd(nar_ctx* ctx, const float* gx, const float* Whg, const float* Whc, const int32_t* sess_off, int64_t B,
                int64_t Hp, float* h_out, float* r_out, float* u_out, float* c_out, float* rh_out, void* stream);
/* d_gx [L,3Hp] = dL/d(pre-activations); h_prev [L,Hp] = state entering the step (dWhg = h_prev^T d_gx[:, :2Hp],
 * dWhc = rh^T d_gx[:, 2Hp:]); WhgT [2Hp,Hp], WhcT [Hp,Hp] are the transposed recurrent blocks.                        */
int nar_gru_bwd(nar_ctx* ctx, const float* d_hout, const float* h_out, const float* r_out, const float* u_out,
                const float* c_out, const float* WhgT, const float* WhcT, const int32_t* sess_off, int64_t B, int64_t Hp,
                float* d_gx, float* h_prev, void* stream);

/* ---- negative sampler (replaces nar_model.py:1220-1304: tf.random_shuffle x(2+clicks),
 *      tf.unique, unsorted_segment_min, tf.setdiff1d inside nested tf.map_fn).  RNG spec:
 *      oracle/sampler_ref.py.  all_items_global [Bg,T1] builds the pool; negatives are
 *      produced for local sessions [sess0, sess0+B).  out [B,T1-1,K] int64, zero padded.   */
int nar_sample_negatives_workspace(int64_t Bg, int64_t T1, int64_t buf_len, int64_t K, int64_t* bytes /*host*/);
int nar_sample_negatives(nar_ctx* ctx, const int64_t* all_items_global, int64_t Bg, int64_t T1,
                         int64_t sess0, int64_t B, const int64_t* buffer, int64_t buf_len,
                         int64_t K, int64_t n_from_buffer, uint64_t seed, uint32_t step,
                         int64_t* out, void* workspace, int64_t workspace_bytes, void* stream);

/* the same, additionally returning for every negative its index in the pool's sorted unique-item table
 * (out_uidx [B,T1-1,K] int32; K*20 = the padding slot for an id-0 negative) and device pointers to that table /
 * its length inside `workspace` (valid until the next call with the same workspace).                              */
int nar_sample_negatives_uidx(nar_ctx* ctx, const int64_t* all_items_global, int64_t Bg, int64_t T1,
                              int64_t sess0, int64_t B, const int64_t* buffer, int64_t buf_len,
                              int64_t K, int64_t n_from_buffer, uint64_t seed, uint32_t step,
                              int64_t* out, int32_t* out_uidx /*or NULL*/, const int64_t** unique_items /*host, out*/,
                              const int32_t** n_unique /*host, out*/, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- per-unique-id CAR layer 1 (nar_model.py:343-405; see csrc/car.cu): pre-activation of candidate (l, j) =
 *      j == 0 ? PP[l] : PC[l] + PI[neg_uidx[pos_idx[l], j-1]]; H1c [L*(1+K), C] = act(pre).                        */
int nar_car_combine(const float* PP /*[L,C]*/, const float* PC /*[L,C]*/, const float* PI /*[U,C]*/,
                    const int32_t* pos_idx, const int32_t* neg_uidx, int64_t L, int64_t K, int64_t C, int act,
                    float* H1c, void* stream);
/* backward: dPP[l] = dH1c[l,0]; dPC[l] = sum_k dH1c[l,1+k]; dPI[u] = sum of the rows that drew u (fixed order).   */
int nar_car_segsum(const float* dH1c, int64_t L, int64_t K, int64_t C, int64_t U, const uint16_t* Mt, int64_t ld_mt,
                   const int32_t* pos_idx, const int32_t* neg_uidx, float* dPP, float* dPC, float* dPI, void* stream);

/* ---- scorer + loss (replaces tf.multiply + matching_dense_layer_1..4 :478-500, softmax
 *      :515, log :660, masked mean :664).                                                  */
/* prod[r,:] = cand[r,:] * pred[r / n_cand,:]   (tf.multiply nar_model.py:478,:493)          */
int nar_mul_pred(const float* cand, const float* pred, int64_t n_pos, int64_t n_cand, int64_t C,
                 float* prod, void* stream);
/* d_cand = d_prod*pred*act'(cand) ; d_pred[l] = sum_j d_prod[l,j]*cand[l,j].  cand_act (nar_act): the
 * activation that produced cand (the CAR tanh) is differentiated in the same pass; NAR_ACT_NONE = plain product rule */
int nar_mul_pred_bwd(const float* d_prod, const float* cand, const float* pred, int64_t n_pos, int64_t n_cand,
                     int64_t C, int cand_act, float* d_cand, float* d_pred, void* stream);
/* last Dense(32->1) + /temperature + log-softmax over the 1+K candidates + masked mean CE,
 * forward and backward in one pass.  z3 [n_pos*n_cand, ld_z] ; logits [n_pos,n_cand] ;
 * loss_sum += sum_l -(logp[l,0]) * inv_count ; d_z3 = d(loss)/d(z3) (before leaky');
 * d_m4[k] += ..., d_c4 += ...                                                              */
/* optional novelty regulariser (nar_model.py:517, :531-544, :673-683): total_loss -= factor * mean over the valid
 * positions of sum_k q_k * nov_k, q = softmax over the NEGATIVES only of the scaled scores, nov_k =
 * -log_base(pop_norm[id_k]); its value is added to loss_nov[0] and its gradient to the negatives' score gradients */
typedef struct {
  float factor;               /* novelty_reg_factor; <= 0: off */
  float log_base;             /* popularity_smooth_log_base */
  const float* pop_norm;      /* [num_items] articles_recent_pop_norm */
  const int64_t* cand_ids;    /* [n_pos*n_cand] candidate ids, positive first */
  float* loss_nov;            /* [1] */
} nar_novelty_reg;

int nar_score_softmax_ce(const float* z3, int64_t ld_z, int64_t width, const float* m4, int64_t ld_m4,
                         const float* c4, int64_t n_pos, int64_t n_cand, float inv_temperature,
                         float inv_count, float* logits, float* loss_sum, float* d_z3,
                         float* d_m4, float* d_c4, const nar_novelty_reg* nov /*host, or NULL*/, void* stream);
/* cosine mode (north_star wording; nar_model.py:437 commented l2-normalise): logits =
 * <l2n(pred), l2n(cand)>/temperature fused with the same softmax-CE; writes d_cand, d_pred. */
int nar_cosine_softmax_ce(const float* cand, const float* pred, int64_t n_pos, int64_t n_cand, int64_t C,
                          float inv_temperature, float inv_count, float* logits, float* loss_sum,
                          float* d_cand, float* d_pred, const nar_novelty_reg* nov /*host, or NULL*/, void* stream);

/* ---- evaluation ranking (ModeKeys.EVAL: rank_items_by_predicted_prob nar_model.py:777-795 = tf.nn.top_k over all
 *      1+K candidates, sparse_recall_at_top_k :835-840, define_mrr_metric :862-885).  logits [n_pos,n_cand] as written
 *      by the two *_softmax_ce kernels (already / temperature); cand_ids [n_pos*n_cand]: per position the positive
 *      followed by its negatives.  pred_ids / pred_probs [n_pos,n_cand] (either may be NULL): candidates sorted by
 *      softmax probability, descending, ties to the lower candidate index (top_k order).  metrics[0] += number of
 *      positions whose positive is in the top_n, metrics[1] += sum of 1/rank for those, metrics[2] += n_pos.       */
int nar_rank_candidates(const float* logits, const int64_t* cand_ids, int64_t n_pos, int64_t n_cand, int32_t top_n,
                        int64_t* pred_ids, float* pred_probs, double* metrics /*[3] float64*/, void* stream);

/* ---- host state (CPU, no CUDA): ClickedItemsState.update_items_state (clicked_items_state.py:187-250) in one pass.
 *      buffer [cap,2] int64 {item, timestamp} newest first, zero padded (in/out); batch_items / batch_ts: the step's
 *      non-padded clicks in batch order (nar_model.py:1635-1646); hours_ms = recent_clicks_buffer_hours * 3.6e6;
 *      scratch [cap,2]; recent_pop [V] (out), pop_norm [V] float64 (out) = max(pop / (sum(pop) + 1), min_norm_pop);
 *      articles_pop [V] (in/out, += bincount(batch)).                                                           */
int nar_host_state_update(int64_t* buffer, int64_t cap, const int64_t* batch_items, const int64_t* batch_ts,
                          int64_t n_batch, int64_t hours_ms, int64_t* scratch, int64_t* recent_pop,
                          double* pop_norm, int64_t* articles_pop, int64_t num_items, double min_norm_pop);
/* the same, straight from the padded batch (ItemsStateUpdaterHook.after_run nar_model.py:1635-1646): item_clicked /
 * event_ts [B,T], label_last [B]; batch_scratch [2*B*(T+1)] int64                                                */
int nar_host_state_update_batch(int64_t* buffer, int64_t cap, const int64_t* item_clicked, const int64_t* event_ts,
                                const int64_t* label_last, int64_t B, int64_t T, int64_t hours_ms,
                                int64_t* batch_scratch, int64_t* scratch, int64_t* recent_pop, double* pop_norm,
                                int64_t* articles_pop, int64_t num_items, double min_norm_pop);

/* ---- device-resident ClickedItemsState (same update as above, in HBM; STAGED: not yet on the default training loop).
 *      old_items / old_ts [cap] -> new_items / new_ts [cap] (distinct buffers: ping-pong); all_items [Bg,T+1] =
 *      [item_clicked | label_last_item], event_ts [Bg,T]; recent_pop [V] int64 scratch / output; pop_norm [V] float32
 *      (what the graph reads), pop_norm64 [V] float64 or NULL; articles_pop [V] in/out; err[0] = 1 on an id outside
 *      [0, V).  A batch without clicks leaves new_* untouched: the caller keeps using old_*.                        */
int nar_state_update(const int64_t* old_items, const int64_t* old_ts, int64_t cap, const int64_t* all_items,
                     const int64_t* event_ts, int64_t Bg, int64_t T, int64_t hours_ms, int64_t* new_items,
                     int64_t* new_ts, int64_t* recent_pop, float* pop_norm, double* pop_norm64, int64_t* articles_pop,
                     int64_t num_items, double min_norm_pop, int* err, void* stream);

/* ---- dropout (replaces tf.layers.dropout nar_model.py:338-340 / :351-353 / :367-369 / :417-419 and
 *      DropoutWrapper(output_keep_prob) :1330-1333).  dst[r,c] = src[r,c] * keep(r,c) / keep_prob; the masks come from
 *      the counter-based generator specified in oracle/dropout_ref.py (Philox4x32-10 keyed by seed, counter = column
 *      block, row key, tensor id, step) so that forward, backward and the oracle draw the same bits.
 *      tensor_id > 0: every row belongs to that tensor, row key = row_pos[r] (flat position b*T+t).
 *      tensor_id == 0: feature rows in the n_cand > 0 layout of nar_row_layout: rows [0,n_input) tensor 1 (key =
 *      position), then per position the positive (tensor 2, key = position) and K negatives (tensor 3, key =
 *      position*K + k).  In place (dst == src) is allowed.                                                        */
int nar_dropout_rows(const float* src, float* dst, int64_t rows, int64_t cols, int64_t ld, const int32_t* row_pos,
                     int64_t n_input, int64_t n_cand, int64_t K, int tensor_id, float keep_prob, uint64_t seed,
                     uint32_t step, void* stream);

/* ---- small helpers ------------------------------------------------------------------ */
/* out[c] += sum_r x[r,c]   (bias gradients)                                                */
int nar_colsum_add(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream);
/* y = x * act'(aux) elementwise                                                            */
int nar_act_bwd(const float* dy, const float* y, int64_t n, int act, float* dx, void* stream);
/* out[0] += scale * sum(x^2) / 2  (l2_regularizer, nar_model.py:655)                       */
int nar_l2_loss_add(const float* x, int64_t n, float scale, float* out, void* stream);
int nar_transpose_f32(const float* src, int64_t rows, int64_t cols, int64_t ld_src, float* dst, int64_t ld_dst, void* stream);

/* ---- optimiser (replaces tf.train.AdamOptimizer(lr,.9,.999,1e-8) nar_model.py:708-722;
 *      TF form: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t*m/(sqrt(v)+eps); the gradient of
 *      elements [0,reg_end) gets + reg_l2*w (l2_regularizer); grad is scaled by grad_scale
 *      first (1/world after a sum-allreduce is NOT needed: losses are already global means) */
int nar_adam_tf(float* params, const float* grads, float* m, float* v, int64_t n, int64_t reg_end,
                float reg_l2, float lr, float beta1, float beta2, float eps, int64_t step,
                float* params_lo /* optional [n]: receives w - tf32_trunc(w) of the updated weights */, void* stream);
/* lo[i] = x[i] - tf32_trunc(x[i])  (the second operand plane of the 3xTF32 GEMM)          */
int nar_tf32_lo(const float* x, int64_t n, float* lo, void* stream);

/* =====================================================================================================
 * The whole step behind ONE call (replaces the single session.run(train_op) of nar_trainer_gcom.py:515-517 /
 * MonitoredTrainingSession: sampler -> features -> CAR -> RNN -> FC -> scorer -> loss -> backward -> Adam,
 * nar_model.py:102-728).  The engine sequences every kernel of the step from C: the caller stages the batch in
 * HBM, fills a nar_step_io and makes one call per phase; no per-kernel host round trip is left.
 *   nar_engine_prepare  the weight-independent front of a step (negatives, row lists, normalisation statistics,
 *                       base rows): may run one step ahead on a side stream
 *   nar_engine_step     forward (+ backward when io->train): gradients complete in cfg.grads on return-stream order
 *   nar_engine_apply    TF-Adam over the flat parameter buffer (after the data-parallel gradient exchange, if any)
 * Internally the step forks weight / bias gradients onto an engine-owned auxiliary stream (events, no host sync).
 * ===================================================================================================== */
#define NAR_MAX_LAYERS 4

typedef struct {
  /* dimensions */
  int64_t num_items, C /*CAR_embedding_size*/, Hp /*rnn_units padded to 4*/, Fp /*feature row width*/, ctx_col0;
  int32_t layers, rnn_cell /*0 = UGRNNCell (nar_model.py:1318), 1 = GRUCell (:1315)*/, ranking /*0 = MLP scorer (:444-500), 1 = cosine*/;
  int32_t fwd_precision, bwd_precision;       /* nar_gemm_epilogue.precision of the forward (3 or 4) / backward (1 or 3) GEMMs */
  int32_t dedup;                              /* 1: per-unique-id CAR layer 1 (csrc/car.cu); 0: every candidate row materialised */
  int32_t use_aux_stream;                     /* 1: weight / bias gradients (and the forward session branch) on the auxiliary stream */
  float keep_prob;                            /* dropout_keep_prob (training steps only; < 1 needs dedup == 0) */
  float novelty_reg_factor;                   /* nar_model.py:673-683; 0 = off */
  uint64_t dropout_seed;
  /* hyper-parameters */
  int64_t K /*negatives per click*/, n_from_buffer, buf_len, n_norm;
  float inv_temperature, reg_l2, lr, beta1, beta2, eps;
  uint64_t sampler_seed;
  int32_t world, rank;                        /* data parallel: rank 0 adds the regulariser term to the loss */
  /* flat parameter buffers: params / params_lo / grads / adam_m / adam_v share offsets (floats) */
  float *params, *params_lo, *grads, *adam_m, *adam_v;
  int64_t n_params, reg_end;
  int64_t off_W1, off_b1, off_W2, off_b2, off_W3, off_b3, off_W4, off_b4, off_gamma, off_beta;
  int64_t off_M[4], off_c[4], ld_M[4];        /* matching_dense_layer_1..4 kernels / biases, leading dimensions */
  int64_t off_Wx[NAR_MAX_LAYERS], off_Wh[NAR_MAX_LAYERS], off_rb[NAR_MAX_LAYERS];     /* UGRNN: [in|H, 2Hp] (gate | candidate); GRU: gates (r | u) */
  int64_t off_Wxc[NAR_MAX_LAYERS], off_Whc[NAR_MAX_LAYERS], off_bc[NAR_MAX_LAYERS];  /* GRU only: candidate blocks [in|H, Hp] */
  /* feature plan: static part (segments, tables, metadata, created_at_ts, gamma / beta, column map) */
  nar_feature_plan plan;
} nar_model_cfg;

typedef struct {
  int64_t B /*local sessions*/, Bg /*global sessions*/, T, sess0 /*first local session*/, L /*local valid positions*/,
          L_global /*sum(mask) over the global batch: the loss normaliser*/;
  int64_t L_cap;                  /* positions the workspace was sized for (>= L) */
  int64_t global_step;            /* optimiser steps applied so far; this step is number global_step + 1 */
  uint32_t sampler_step;          /* counter of the negative sampler (training: global_step + 1) */
  int32_t train;                  /* 1: forward + backward, 0: forward + loss only */
  /* staged inputs (device) */
  const int64_t *all_items /*[Bg,T+1] item_clicked | label_last_item*/, *event_ts /*[Bg,T]*/, *item_clicked /*[Bg,T]*/,
                *label_next /*[Bg,T]*/, *buffer /*[buf_len]*/, *max_ts /*[1]*/;
  const float* pop_norm;          /* [num_items] */
  const int64_t* ctx_int[NAR_MAX_SRC];
  const float* ctx_float[NAR_MAX_SRC];
  const int32_t *pos_idx /*[L] flat b*T+t of the valid positions, session-major*/, *sess_off /*[B+1]*/;
  /* buffers */
  void* prep_ws;  int64_t prep_ws_bytes;      /* results of nar_engine_prepare (one slot per step in flight) */
  void* ws;       int64_t ws_bytes;           /* activations / activation gradients of the step */
  float* loss;                                /* [4] device: {cross-entropy (mean over L_global), l2 regulariser, -, -} */
} nar_step_io;

typedef struct nar_engine nar_engine;

int nar_engine_create(nar_ctx* ctx, const nar_model_cfg* cfg /*host*/, nar_engine** out);
int nar_engine_destroy(nar_engine* eng);
/* cfg fields that may change between steps without re-creating the engine (lr, precisions, stream use, world / rank) */
int nar_engine_update_cfg(nar_engine* eng, const nar_model_cfg* cfg /*host*/);
/* bytes of nar_step_io.prep_ws / .ws for a global batch of Bg x T, B local sessions, room for L_cap valid positions */
int nar_engine_workspace_bytes(const nar_engine* eng, int64_t Bg, int64_t B, int64_t T, int64_t L_cap, int32_t train,
                               int64_t* prep_bytes /*host*/, int64_t* ws_bytes /*host*/);
int nar_engine_prepare(nar_engine* eng, const nar_step_io* io /*host*/, void* stream);
int nar_engine_step(nar_engine* eng, const nar_step_io* io /*host*/, void* stream);
int nar_engine_apply(nar_engine* eng, const nar_step_io* io /*host*/, void* stream);
/* after the caller wrote the weights itself (initialisation, checkpoint restore): rebuild what the engine derives from
 * them (the bf16x3 planes of the forward weights, fwd_precision 4)                                                      */
int nar_engine_refresh(nar_engine* eng, void* stream);
/* device address / shape of a named intermediate of the LAST nar_engine_prepare / nar_engine_step with this io
 * (parity tests, evaluation ranking): "neg", "neg_uidx", "row_pos", "row_item", "stats", "X", "H1", "E", "HO<i>", "F1",
 * "PR", "logits", "base_pos", "base_item", ...  Returns NAR_ERR_INVALID for an unknown name.                       */
int nar_engine_buffer(const nar_engine* eng, const nar_step_io* io /*host*/, const char* name, void** ptr /*host*/,
                      int64_t* rows /*host*/, int64_t* ld /*host*/);
/* kernels launched by this engine so far */
int64_t nar_engine_launch_count(const nar_engine* eng);

#ifdef __cplusplus
}
#endif
#endif /* NAR_B200_H */
