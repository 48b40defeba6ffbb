// The NAR training / evaluation step sequenced in C: one call per phase instead of ~50 host round trips.
//
// Replaces the single session.run(train_op) of the reference trainer (nar_trainer_gcom.py:515-517) over the graph of
// nar_model.py:102-728: sampler (:265-276) -> features (:314-370) -> CAR (:374-405) -> RNN (:408, :1308-1342) ->
// FC1/FC2 (:410-438) -> scorer (:444-517) -> loss (:639-704) -> gradients -> TF-Adam (:706-722).  Only the valid
// positions (mask == 1) are materialised: padded positions never reach the loss (:660-664).
//
// Row layouts (L = valid local positions, n_cand = 1+K, R = L + L*n_cand):
//   H1 / E [R, C]  rows [0,L) = clicked items, then per position its positive followed by its K negatives
//   full mode      X [R, Fp] feature rows in the same order (every candidate row gathered and multiplied by W1)
//   dedup mode     XB [2L+U, Fp] base rows: L clicked, L positives, U unique-negative ITEM rows (csrc/car.cu):
//                  PP = XB[L:2L] W1 + b1 ; PI = XB[2L:, item cols] W1[item rows] ; PC = XB[:L, ctx cols] W1[ctx rows] + b1
//                  H1[L + l*n_cand + j] = leaky(j == 0 ? PP[l] : PC[l] + PI[u(l, j-1)])
//                  backward: DB = [dH1(in) | dPP | dPI | dPC] by fixed-order segment sums, then ONE wgrad / dgrad
//                  over the 2L+U base rows (+ the context block), instead of two GEMMs over all R rows.
// Streams: the caller's stream carries the critical path (forward, dgrad chain); everything only Adam needs (weight /
// bias gradients) and the forward session branch run on an engine-owned auxiliary stream behind events.
#include "common.cuh"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" int nar_sample_negatives_uidx(nar_ctx*, const int64_t*, int64_t, int64_t, int64_t, int64_t, const int64_t*, int64_t,
                                         int64_t, int64_t, uint64_t, uint32_t, int64_t*, int32_t*, const int64_t**,
                                         const int32_t**, void*, int64_t, void*);

namespace {

constexpr int N_EVENTS = 64;

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Carver {
  char* base; int64_t off;
  explicit Carver(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T> T* take(int64_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + (n > 0 ? n : 1) * (int64_t)sizeof(T), 256);
    return p;
  }
};

struct PrepBufs {
  void* sampler_ws; int64_t sampler_bytes;
  int64_t* neg; int32_t* neg_uidx; float* stats; int32_t* row_pos; int64_t* row_item;
  int32_t* base_pos; int64_t* base_item; uint16_t* Mt; int64_t ld_mt, U;
};

struct StepBufs {
  float *X, *dX, *H1, *E, *dE, *dH1, *F1, *PR, *logits, *PD, *dPD, *Z1, *Z2, *Z3, *dZ1, *dZ2, *dZ3, *dPR, *dF1, *dHO;
  float *GX[NAR_MAX_LAYERS], *HO[NAR_MAX_LAYERS], *GT[NAR_MAX_LAYERS], *CD[NAR_MAX_LAYERS], *dGX[NAR_MAX_LAYERS],
        *HPV[NAR_MAX_LAYERS], *dHOb[NAR_MAX_LAYERS], *HOd[NAR_MAX_LAYERS],   // HOd: RNN outputs after DropoutWrapper
        *UO[NAR_MAX_LAYERS], *RH[NAR_MAX_LAYERS];                          // GRU: update gate, r * previous state
  float *PP, *PI, *PC, *DB;        // dedup: layer-1 pre-activations and their gradients
};

}  // namespace

// bf16x3 planes of the forward weights (fwd_precision 4; see gemm_tcgen05.cu MODE 4): one entry per (weight block, K) a
// forward GEMM uses; refreshed by one nar_pack_bf16x3 launch after every optimiser step / weight load
constexpr int MAX_PLANES = 32;
struct PlaneSet {
  int n;
  int64_t off_W[MAX_PLANES]; int32_t K[MAX_PLANES], N[MAX_PLANES], ldw[MAX_PLANES], ld_out[MAX_PLANES];
  int64_t dst[MAX_PLANES];                 // element offset of the plane in `buf`
  const float* W[MAX_PLANES]; void* out[MAX_PLANES];
  uint16_t* buf; void* descs;
};

struct nar_engine {
  nar_ctx* ctx;
  nar_model_cfg cfg;
  PlaneSet planes;
  cudaStream_t aux, aux2;
  cudaEvent_t ev[N_EVENTS];
  int ev_i;
  float* WhT[NAR_MAX_LAYERS];
  float* WhcT[NAR_MAX_LAYERS];     // GRU: transposed candidate recurrent block
  int64_t launches;
};

namespace {

int64_t prep_carve(const nar_engine* e, int64_t Bg, int64_t B, int64_t T, int64_t L_cap, void* base, PrepBufs* pb) {
  const nar_model_cfg& c = e->cfg;
  const int64_t K = c.K, n_cand = K + 1, Rcap = L_cap * (n_cand + 1);
  Carver cv(base);
  int64_t sb = 0;
  nar_sample_negatives_workspace(Bg, T + 1, c.buf_len, K, &sb);
  pb->sampler_bytes = sb;
  pb->sampler_ws = cv.take<char>(sb);
  pb->neg = cv.take<int64_t>(Bg * T * K);
  pb->neg_uidx = cv.take<int32_t>(c.dedup ? Bg * T * K : 1);
  pb->stats = cv.take<float>(24);
  pb->row_pos = cv.take<int32_t>(Rcap);
  pb->row_item = cv.take<int64_t>(Rcap);
  pb->U = K * 20 + 1;
  pb->ld_mt = align_up(L_cap > 0 ? L_cap : 1, 8);
  pb->base_pos = cv.take<int32_t>(c.dedup ? 2 * L_cap + pb->U : 1);
  pb->base_item = cv.take<int64_t>(c.dedup ? 2 * L_cap + pb->U : 1);
  pb->Mt = cv.take<uint16_t>(c.dedup ? pb->U * pb->ld_mt : 1);
  (void)B;
  return cv.off;
}

int64_t step_carve(const nar_engine* e, int64_t L_cap, int train, void* base, StepBufs* sb) {
  const nar_model_cfg& c = e->cfg;
  const int64_t K = c.K, n_cand = K + 1, Rc = L_cap * n_cand, R = L_cap + Rc, C = c.C, Hp = c.Hp, Fp = c.Fp;
  const int64_t U = K * 20 + 1, NB = 2 * L_cap + U;
  memset(sb, 0, sizeof(*sb));
  Carver cv(base);
  if (c.dedup) { sb->X = cv.take<float>(NB * Fp); } else { sb->X = cv.take<float>(R * Fp); }
  sb->H1 = cv.take<float>(R * C);
  sb->E = cv.take<float>(R * C);
  const int64_t gw = c.rnn_cell == 1 ? 3 : 2;           // gate blocks per unit: UGRNN (gate | candidate), GRU (r | u | candidate)
  for (int i = 0; i < c.layers; ++i) {
    sb->GX[i] = cv.take<float>(L_cap * gw * Hp); sb->HO[i] = cv.take<float>(L_cap * Hp);
    sb->GT[i] = cv.take<float>(L_cap * Hp); sb->CD[i] = cv.take<float>(L_cap * Hp);
    if (c.rnn_cell == 1) { sb->UO[i] = cv.take<float>(L_cap * Hp); sb->RH[i] = cv.take<float>(L_cap * Hp); }
  }
  sb->F1 = cv.take<float>(L_cap * 512);
  sb->PR = cv.take<float>(L_cap * C);
  sb->logits = cv.take<float>(L_cap * n_cand);
  if (c.dedup) { sb->PP = cv.take<float>(L_cap * C); sb->PI = cv.take<float>(U * C); sb->PC = cv.take<float>(L_cap * C); }
  if (c.ranking == 0) {
    sb->PD = cv.take<float>(Rc * C); sb->Z1 = cv.take<float>(Rc * 128); sb->Z2 = cv.take<float>(Rc * 64); sb->Z3 = cv.take<float>(Rc * 32);
  }
  if (train) {
    sb->dE = cv.take<float>(R * C);
    sb->dPR = cv.take<float>(L_cap * C); sb->dF1 = cv.take<float>(L_cap * 512); sb->dHO = cv.take<float>(L_cap * Hp);
    for (int i = 0; i < c.layers; ++i) {
      sb->dGX[i] = cv.take<float>(L_cap * gw * Hp); sb->HPV[i] = cv.take<float>(L_cap * Hp); sb->dHOb[i] = cv.take<float>(L_cap * Hp);
    }
    if (c.ranking == 0) {
      sb->dZ3 = cv.take<float>(Rc * 32); sb->dZ2 = cv.take<float>(Rc * 64); sb->dZ1 = cv.take<float>(Rc * 128); sb->dPD = cv.take<float>(Rc * C);
    }
    if (c.keep_prob < 1.f)
      for (int i = 0; i < c.layers; ++i) sb->HOd[i] = cv.take<float>(L_cap * Hp);
    if (c.dedup) {
      sb->dH1 = cv.take<float>(Rc * C);                    // candidate rows only; the input rows' dH1 is DB[0:L]
      sb->DB = cv.take<float>((3 * L_cap + U) * C);
      sb->dX = cv.take<float>(NB * Fp);
    } else {
      sb->dH1 = cv.take<float>(R * C);
      sb->dX = cv.take<float>(R * Fp);
    }
  }
  return cv.off;
}

// ---------------------------------------------------------------------------------------------- one step
struct Seq {
  nar_engine* e; const nar_step_io* io; cudaStream_t main, aux;
  bool use_aux, aux_dirty = false;
  int rc = 0;
  const nar_model_cfg& c;
  Seq(nar_engine* e_, const nar_step_io* io_, cudaStream_t s)
      : e(e_), io(io_), main(s), aux(e_->aux), use_aux(e_->cfg.use_aux_stream != 0), c(e_->cfg) {}

  cudaEvent_t next_event() { cudaEvent_t v = e->ev[e->ev_i]; e->ev_i = (e->ev_i + 1) % N_EVENTS; return v; }
  // stream that deferred work (weight / bias gradients, forward session branch) runs on, after everything queued on main so far
  cudaStream_t fork() { return fork(main); }
  cudaStream_t fork(cudaStream_t from) {           // deferred work of a chain that itself runs on `from`
    if (!use_aux) return main;
    cudaEvent_t v = next_event();
    if (cudaEventRecord(v, from) != cudaSuccess || cudaStreamWaitEvent(aux, v, 0) != cudaSuccess) rc = rc ? rc : (int)cudaGetLastError();
    aux_dirty = true;
    return aux;
  }
  // second auxiliary stream: an independent CHAIN (the session backward) next to the main stream's
  bool two_chains = false;
  cudaStream_t fork2() {
    if (!use_aux || !two_chains) return main;
    cudaEvent_t v = next_event();
    if (cudaEventRecord(v, main) != cudaSuccess || cudaStreamWaitEvent(e->aux2, v, 0) != cudaSuccess) rc = rc ? rc : (int)cudaGetLastError();
    return e->aux2;
  }
  void join2() {
    if (!use_aux || !two_chains) return;
    cudaEvent_t v = next_event();
    if (cudaEventRecord(v, e->aux2) != cudaSuccess || cudaStreamWaitEvent(main, v, 0) != cudaSuccess) rc = rc ? rc : (int)cudaGetLastError();
  }
  void join() {
    if (!use_aux || !aux_dirty) return;
    cudaEvent_t v = next_event();
    if (cudaEventRecord(v, aux) != cudaSuccess || cudaStreamWaitEvent(main, v, 0) != cudaSuccess) rc = rc ? rc : (int)cudaGetLastError();
    aux_dirty = false;
  }
  void chk(int r) { if (r && !rc) rc = r; ++e->launches; }

  const float* W(int64_t off) const { return c.params + off; }
  float* G(int64_t off) const { return c.grads + off; }

  // Y[M,N] = act(X[M,Kd] * W[Kd,N] + b)      (W stored [in,out]: MN-major B operand, or its bf16x3 plane)
  void fwd(const float* X, int64_t ldx, int64_t off_W, int64_t ldw, int64_t off_b, float* Y, int64_t ldy, int64_t M, int64_t N,
           int64_t Kd, int act, cudaStream_t st) {
    nar_gemm_epilogue ep; memset(&ep, 0, sizeof(ep));
    ep.bias = off_b >= 0 ? W(off_b) : nullptr; ep.act = act; ep.split_k = 1; ep.precision = c.fwd_precision;
    ep.b_lo = c.fwd_precision == 3 ? c.params_lo + off_W : nullptr;
    if (c.fwd_precision == 4) {
      const PlaneSet& ps = e->planes;
      int i = 0;
      for (; i < ps.n; ++i) if (ps.off_W[i] == off_W && ps.K[i] == Kd && ps.N[i] == N) break;
      if (i == ps.n) { ep.precision = 3; ep.b_lo = c.params_lo + off_W; }      // no plane for this block: 3xTF32
      else { ep.b_bf16 = ps.buf + ps.dst[i]; ep.ld_bf16 = ps.ld_out[i]; }
    }
    chk(nar_gemm_tf32(e->ctx, M, N, Kd, X, ldx, 1, W(off_W), ldw, 0, Y, ldy, &ep, st));
  }
  // dX[M,n_in] (+)= dY[M,n_out] * W^T, optionally times act'(aux)
  void dgrad(const float* dY, int64_t lddy, int64_t off_W, int64_t ldw, float* dX, int64_t lddx, int64_t M, int64_t n_in,
             int64_t n_out, int dact, const float* auxp, int64_t ld_aux, int accumulate, cudaStream_t st) {
    nar_gemm_epilogue ep; memset(&ep, 0, sizeof(ep));
    ep.dact = dact; ep.aux = auxp; ep.ld_aux = ld_aux; ep.accumulate = accumulate; ep.split_k = accumulate ? 0 : 1;
    ep.precision = c.bwd_precision;
    chk(nar_gemm_tf32(e->ctx, M, n_in, n_out, dY, lddy, 1, W(off_W), ldw, 1, dX, lddx, &ep, st));
  }
  // dW[n_in,n_out] += X[rows,n_in]^T * dY[rows,n_out]   (split-K, red.add into the gradient buffer)
  void wgrad(const float* X, int64_t ldx, const float* dY, int64_t lddy, int64_t off_W, int64_t ldw, int64_t n_in, int64_t n_out,
             int64_t rows, cudaStream_t st) {
    nar_gemm_epilogue ep; memset(&ep, 0, sizeof(ep));
    ep.accumulate = 1; ep.split_k = 0; ep.precision = c.bwd_precision;
    chk(nar_gemm_tf32(e->ctx, n_in, n_out, rows, X, ldx, 0, dY, lddy, 0, G(off_W), ldw, &ep, st));
  }
  void bgrad(const float* dY, int64_t ld, int64_t rows, int64_t cols, int64_t off_b, cudaStream_t st) {
    chk(nar_colsum_add(dY, rows, cols, ld, G(off_b), st));
  }
};

int run_step(nar_engine* e, const nar_step_io* io, cudaStream_t main) {
  const nar_model_cfg& c = e->cfg;
  const int64_t B = io->B, T = io->T, L = io->L, K = c.K, n_cand = K + 1, Rc = L * n_cand, R = L + Rc;
  const int64_t C = c.C, Hp = c.Hp, Fp = c.Fp, c0 = c.ctx_col0;
  const int train = io->train;
  if (L > io->L_cap || L < 0 || B <= 0) return NAR_ERR_INVALID;
  PrepBufs pb; StepBufs sb;
  if (prep_carve(e, io->Bg, B, T, io->L_cap, io->prep_ws, &pb) > io->prep_ws_bytes) return NAR_ERR_WORKSPACE;
  if (step_carve(e, io->L_cap, train, io->ws, &sb) > io->ws_bytes) return NAR_ERR_WORKSPACE;
  NAR_CHECK_CUDA(cudaMemsetAsync(io->loss, 0, 4 * sizeof(float), main));
  if (train) NAR_CHECK_CUDA(cudaMemsetAsync(c.grads, 0, (size_t)c.n_params * sizeof(float), main));
  if (L == 0) return NAR_OK;
  // dropout (training steps only): masks are per candidate row, so every row must be materialised
  const bool drop = train && c.keep_prob < 1.f;
  if (drop && c.dedup) return NAR_ERR_INVALID;
  const uint32_t dstep = (uint32_t)(io->global_step + 1);
  Seq s(e, io, main);
  // NAR_DEBUG_DROP_ONLY=<tensor id> (diagnostics, mirrored by the oracle): dropout at that site only
  int drop_only = -1;
  if (drop) { const char* v = getenv("NAR_DEBUG_DROP_ONLY"); if (v) drop_only = atoi(v); }
  auto dropout = [&](const float* src, float* dst, int64_t rows, int64_t cols, const int32_t* rpos, int tid, cudaStream_t st) {
    if (drop_only >= 0 && drop_only != tid) {
      if (src != dst) cudaMemcpyAsync(dst, src, (size_t)rows * cols * sizeof(float), cudaMemcpyDeviceToDevice, st);
      return;
    }
    s.chk(nar_dropout_rows(src, dst, rows, cols, cols, rpos, L, n_cand, K, tid, c.keep_prob, c.dropout_seed, dstep, st));
  };
  const float inv_count = 1.0f / (float)(io->L_global > 0 ? io->L_global : 1);
  const int64_t U = pb.U, NB = 2 * L + U;

  // ---- feature plan of this step: static part + the staged inputs
  nar_feature_plan plan = c.plan;
  for (int i = 0; i < NAR_MAX_SRC; ++i) { plan.ctx_int[i] = io->ctx_int[i]; plan.ctx_float[i] = io->ctx_float[i]; }
  plan.pop_norm = io->pop_norm;
  plan.stats = pb.stats;
  nar_row_layout rl;
  if (c.dedup) { rl.n_rows = NB; rl.n_input = L; rl.n_cand = 0; rl.n_positive = L; rl.n_full = 2 * L; rl.ctx_col0 = c0; }
  else { rl.n_rows = R; rl.n_input = L; rl.n_cand = n_cand; rl.n_positive = 0; rl.n_full = R; rl.ctx_col0 = c0; }
  const int32_t* g_pos = c.dedup ? pb.base_pos : pb.row_pos;
  const int64_t* g_item = c.dedup ? pb.base_item : pb.row_item;
  s.chk(nar_gather_features(e->ctx, &plan, g_pos, g_item, &rl, io->event_ts, io->max_ts, sb.X, main));
  if (drop) dropout(sb.X, sb.X, R, Fp, pb.row_pos, 0, main);          // nar_model.py:338-340, :351-353, :367-369

  // ---- session branch: RNN (nar_model.py:408, :1308-1342) + FC1 / FC2 (:410-438) on the L clicked rows
  auto session_branch = [&](cudaStream_t st) {
    const float* rnn_in = sb.E; int64_t n_in = C;
    for (int i = 0; i < c.layers; ++i) {
      if (c.rnn_cell == 1) {
        // GRUCell: gx = (x Wxg + bg | x Wxc + bc), then the recurrence (csrc/gru.cu)
        s.fwd(rnn_in, i == 0 ? C : Hp, c.off_Wx[i], 2 * Hp, c.off_rb[i], sb.GX[i], 3 * Hp, L, 2 * Hp, n_in, NAR_ACT_NONE, st);
        s.fwd(rnn_in, i == 0 ? C : Hp, c.off_Wxc[i], Hp, c.off_bc[i], sb.GX[i] + 2 * Hp, 3 * Hp, L, Hp, n_in, NAR_ACT_NONE, st);
        s.chk(nar_gru_fwd(e->ctx, sb.GX[i], s.W(c.off_Wh[i]), s.W(c.off_Whc[i]), io->sess_off, B, Hp, sb.HO[i], sb.GT[i], sb.UO[i],
                          sb.CD[i], sb.RH[i], st));
      } else {
      s.fwd(rnn_in, i == 0 ? C : Hp, c.off_Wx[i], 2 * Hp, c.off_rb[i], sb.GX[i], 2 * Hp, L, 2 * Hp, n_in, NAR_ACT_NONE, st);
      s.chk(nar_ugrnn_fwd(e->ctx, sb.GX[i], s.W(c.off_Wh[i]), io->sess_off, B, Hp, sb.HO[i], sb.GT[i], sb.CD[i], st));
      }
      // DropoutWrapper(output_keep_prob) (nar_model.py:1330-1333): the cell's OUTPUT is dropped, its state is not
      if (drop) dropout(sb.HO[i], sb.HOd[i], L, Hp, io->pos_idx, 8 + i, st);
      rnn_in = drop ? sb.HOd[i] : sb.HO[i]; n_in = Hp;
    }
    s.fwd(rnn_in, Hp, c.off_W3, 512, c.off_b3, sb.F1, 512, L, 512, Hp, NAR_ACT_LEAKY_RELU, st);
    if (drop) dropout(sb.F1, sb.F1, L, 512, io->pos_idx, 4, st);                                  // nar_model.py:417-419
    s.fwd(sb.F1, 512, c.off_W4, C, c.off_b4, sb.PR, C, L, C, 512, NAR_ACT_TANH, st);
  };

  // ---- CAR (nar_model.py:374-405): the clicked rows first, so that the session branch can run under the candidates
  // (moving these two GEMMs into the session branch as well was measured neutral and is not used: DESIGN.md section 6)
  float* H1c = sb.H1 + L * C; float* Ec = sb.E + L * C;
  s.fwd(sb.X, Fp, c.off_W1, C, c.off_b1, sb.H1, C, L, C, Fp, NAR_ACT_LEAKY_RELU, main);
  s.fwd(sb.H1, C, c.off_W2, C, c.off_b2, sb.E, C, L, C, C, NAR_ACT_TANH, main);
  { cudaStream_t st = s.fork(); session_branch(st); }
  if (c.dedup) {
    s.fwd(sb.X + L * Fp, Fp, c.off_W1, C, c.off_b1, sb.PP, C, L, C, Fp, NAR_ACT_NONE, main);                       // positives: full rows
    s.fwd(sb.X + 2 * L * Fp, Fp, c.off_W1, C, -1, sb.PI, C, U, C, c0, NAR_ACT_NONE, main);                          // item half, once per unique id
    s.fwd(sb.X + c0, Fp, c.off_W1 + c0 * C, C, c.off_b1, sb.PC, C, L, C, Fp - c0, NAR_ACT_NONE, main);             // context half, once per position
    s.chk(nar_car_combine(sb.PP, sb.PC, sb.PI, io->pos_idx, pb.neg_uidx, L, K, C, NAR_ACT_LEAKY_RELU, H1c, main));
  } else {
    s.fwd(sb.X + L * Fp, Fp, c.off_W1, C, c.off_b1, H1c, C, Rc, C, Fp, NAR_ACT_LEAKY_RELU, main);
  }
  s.fwd(H1c, C, c.off_W2, C, c.off_b2, Ec, C, Rc, C, C, NAR_ACT_TANH, main);
  s.join();

  // ---- scorer + loss (nar_model.py:444-517, :639-667), optional novelty regulariser (:673-683)
  nar_novelty_reg nov; memset(&nov, 0, sizeof(nov));
  nov.factor = c.novelty_reg_factor; nov.log_base = c.plan.log_base_novelty; nov.pop_norm = io->pop_norm;
  nov.cand_ids = pb.row_item + L; nov.loss_nov = io->loss + 2;
  const nar_novelty_reg* novp = c.novelty_reg_factor > 0.f ? &nov : nullptr;
  if (c.ranking == 0) {
    s.chk(nar_mul_pred(Ec, sb.PR, L, n_cand, C, sb.PD, main));
    s.fwd(sb.PD, C, c.off_M[0], c.ld_M[0], c.off_c[0], sb.Z1, 128, Rc, 128, C, NAR_ACT_LEAKY_RELU, main);
    s.fwd(sb.Z1, 128, c.off_M[1], c.ld_M[1], c.off_c[1], sb.Z2, 64, Rc, 64, 128, NAR_ACT_LEAKY_RELU, main);
    s.fwd(sb.Z2, 64, c.off_M[2], c.ld_M[2], c.off_c[2], sb.Z3, 32, Rc, 32, 64, NAR_ACT_LEAKY_RELU, main);
    s.chk(nar_score_softmax_ce(sb.Z3, 32, 32, s.W(c.off_M[3]), c.ld_M[3], s.W(c.off_c[3]), L, n_cand, c.inv_temperature, inv_count,
                               sb.logits, io->loss, train ? sb.dZ3 : nullptr, train ? s.G(c.off_M[3]) : nullptr,
                               train ? s.G(c.off_c[3]) : nullptr, novp, main));
  } else {
    s.chk(nar_cosine_softmax_ce(Ec, sb.PR, L, n_cand, C, c.inv_temperature, inv_count, sb.logits, io->loss,
                                train ? sb.dE + L * C : nullptr, train ? sb.dPR : nullptr, novp, main));
  }
  // every rank holds the same weights: the regulariser is added once (rank 0) so that a sum over ranks is exact
  // (off the critical path: it only feeds the reported loss)
  if (c.reg_l2 > 0.f && c.rank == 0) { cudaStream_t st = s.fork(); s.chk(nar_l2_loss_add(c.params, c.reg_end, c.reg_l2, io->loss + 1, st)); }
  if (!train || s.rc) { s.join(); return s.rc; }

  // =============================================================================================== backward
  float* dEc = sb.dE + L * C;
  if (c.ranking == 0) {
    { cudaStream_t st = s.fork(); s.wgrad(sb.Z2, 64, sb.dZ3, 32, c.off_M[2], c.ld_M[2], 64, 32, Rc, st); s.bgrad(sb.dZ3, 32, Rc, 32, c.off_c[2], st); }
    s.dgrad(sb.dZ3, 32, c.off_M[2], c.ld_M[2], sb.dZ2, 64, Rc, 64, 32, NAR_ACT_LEAKY_RELU, sb.Z2, 64, 0, main);
    { cudaStream_t st = s.fork(); s.wgrad(sb.Z1, 128, sb.dZ2, 64, c.off_M[1], c.ld_M[1], 128, 64, Rc, st); s.bgrad(sb.dZ2, 64, Rc, 64, c.off_c[1], st); }
    s.dgrad(sb.dZ2, 64, c.off_M[1], c.ld_M[1], sb.dZ1, 128, Rc, 128, 64, NAR_ACT_LEAKY_RELU, sb.Z1, 128, 0, main);
    { cudaStream_t st = s.fork(); s.wgrad(sb.PD, C, sb.dZ1, 128, c.off_M[0], c.ld_M[0], C, 128, Rc, st); s.bgrad(sb.dZ1, 128, Rc, 128, c.off_c[0], st); }
    s.dgrad(sb.dZ1, 128, c.off_M[0], c.ld_M[0], sb.dPD, C, Rc, C, 128, NAR_ACT_NONE, nullptr, 0, 0, main);
    // candidate rows: through the product and the CAR tanh in one pass; d(pred) reduced over the candidates
    s.chk(nar_mul_pred_bwd(sb.dPD, Ec, sb.PR, L, n_cand, C, NAR_ACT_TANH, dEc, sb.dPR, main));
  } else {
    s.chk(nar_act_bwd(dEc, Ec, Rc * C, NAR_ACT_TANH, dEc, main));
  }
  // ---- two independent chains from here on:
  //   S  session branch: FC2 / FC1 (nar_model.py:410-426) -> BPTT -> d(E) of the L clicked rows - a dozen small kernels
  //   C  candidates: CAR layer-2 backward over the L*(1+K) candidate rows - the big GEMMs (+ the segment sums)
  // S runs on a second auxiliary stream next to C; they meet before the clicked rows' CAR backward.
  auto session_backward = [&](cudaStream_t ss) {
    s.chk(nar_act_bwd(sb.dPR, sb.PR, L * C, NAR_ACT_TANH, sb.dPR, ss));
    { cudaStream_t st = s.two_chains ? ss : s.fork(ss); s.wgrad(sb.F1, 512, sb.dPR, C, c.off_W4, C, 512, C, L, st); s.bgrad(sb.dPR, C, L, C, c.off_b4, st); }
    s.dgrad(sb.dPR, C, c.off_W4, C, sb.dF1, 512, L, 512, C, NAR_ACT_LEAKY_RELU, sb.F1, 512, 0, ss);
    if (drop) dropout(sb.dF1, sb.dF1, L, 512, io->pos_idx, 4, ss);     // F1 holds the dropped activations: re-apply the mask to the gradient
    const float* rnn_out = drop ? sb.HOd[c.layers - 1] : sb.HO[c.layers - 1];
    { cudaStream_t st = s.two_chains ? ss : s.fork(ss); s.wgrad(rnn_out, Hp, sb.dF1, 512, c.off_W3, 512, Hp, 512, L, st); s.bgrad(sb.dF1, 512, L, 512, c.off_b3, st); }
    s.dgrad(sb.dF1, 512, c.off_W3, 512, sb.dHO, Hp, L, Hp, 512, NAR_ACT_NONE, nullptr, 0, 0, ss);
    float* dho = sb.dHO;
    for (int i = c.layers - 1; i >= 0; --i) {
      if (drop) dropout(dho, dho, L, Hp, io->pos_idx, 8 + i, ss);       // gradient of the dropped cell output
      const float* x_in = i == 0 ? sb.E : (drop ? sb.HOd[i - 1] : sb.HO[i - 1]);
      const int64_t n_in = i == 0 ? C : Hp;
      if (c.rnn_cell == 1) {
        const int64_t W3 = 3 * Hp;
        s.chk(nar_transpose_f32(s.W(c.off_Wh[i]), Hp, 2 * Hp, 2 * Hp, e->WhT[i], Hp, ss));
        s.chk(nar_transpose_f32(s.W(c.off_Whc[i]), Hp, Hp, Hp, e->WhcT[i], Hp, ss));
        s.chk(nar_gru_bwd(e->ctx, dho, sb.HO[i], sb.GT[i], sb.UO[i], sb.CD[i], e->WhT[i], e->WhcT[i], io->sess_off, B, Hp, sb.dGX[i],
                          sb.HPV[i], ss));
        const float* dg = sb.dGX[i]; const float* dc = sb.dGX[i] + 2 * Hp;
        {
          cudaStream_t st = s.two_chains ? ss : s.fork(ss);
          s.wgrad(x_in, n_in, dg, W3, c.off_Wx[i], 2 * Hp, n_in, 2 * Hp, L, st);
          s.wgrad(x_in, n_in, dc, W3, c.off_Wxc[i], Hp, n_in, Hp, L, st);
          s.wgrad(sb.HPV[i], Hp, dg, W3, c.off_Wh[i], 2 * Hp, Hp, 2 * Hp, L, st);
          s.wgrad(sb.RH[i], Hp, dc, W3, c.off_Whc[i], Hp, Hp, Hp, L, st);
          s.bgrad(dg, W3, L, 2 * Hp, c.off_rb[i], st);
          s.bgrad(dc, W3, L, Hp, c.off_bc[i], st);
        }
        // d(input) = d_gx[:, :2Hp] Wxg^T + d_gx[:, 2Hp:] Wxc^T (two GEMMs into one buffer), then through the CAR tanh for layer 0
        float* dxin = i == 0 ? sb.dE : sb.dHOb[i];
        const int64_t ldx = i == 0 ? C : Hp;
        s.dgrad(dg, W3, c.off_Wx[i], 2 * Hp, dxin, ldx, L, n_in, 2 * Hp, NAR_ACT_NONE, nullptr, 0, 0, ss);
        s.dgrad(dc, W3, c.off_Wxc[i], Hp, dxin, ldx, L, n_in, Hp, NAR_ACT_NONE, nullptr, 0, 1, ss);
        if (i == 0) s.chk(nar_act_bwd(sb.dE, sb.E, L * C, NAR_ACT_TANH, sb.dE, ss));
        else dho = sb.dHOb[i];
        continue;
      }
      s.chk(nar_transpose_f32(s.W(c.off_Wh[i]), Hp, 2 * Hp, 2 * Hp, e->WhT[i], Hp, ss));
      s.chk(nar_ugrnn_bwd(e->ctx, dho, sb.HO[i], sb.GT[i], sb.CD[i], e->WhT[i], io->sess_off, B, Hp, sb.dGX[i], sb.HPV[i], ss));
      {
        cudaStream_t st = s.two_chains ? ss : s.fork(ss);
        s.wgrad(x_in, n_in, sb.dGX[i], 2 * Hp, c.off_Wx[i], 2 * Hp, n_in, 2 * Hp, L, st);
        s.wgrad(sb.HPV[i], Hp, sb.dGX[i], 2 * Hp, c.off_Wh[i], 2 * Hp, Hp, 2 * Hp, L, st);
        s.bgrad(sb.dGX[i], 2 * Hp, L, 2 * Hp, c.off_rb[i], st);
      }
      if (i == 0) {
        s.dgrad(sb.dGX[0], 2 * Hp, c.off_Wx[0], 2 * Hp, sb.dE, C, L, C, 2 * Hp, NAR_ACT_TANH, sb.E, C, 0, ss);   // clicked rows of dE (pre-tanh)
      } else {
        s.dgrad(sb.dGX[i], 2 * Hp, c.off_Wx[i], 2 * Hp, sb.dHOb[i], Hp, L, Hp, 2 * Hp, NAR_ACT_NONE, nullptr, 0, 0, ss);
        dho = sb.dHOb[i];
      }
    }
  };
  // (NAR_BWD_CHAINS=1.  With two chains S keeps its own weight gradients on its stream: sharing the ONE auxiliary stream
  // in enqueue order made the big layer-2 wgrad of C wait behind the last small wgrad of S: 1.22 -> 1.27 ms per G1 step.)
  { const char* v = getenv("NAR_BWD_CHAINS"); s.two_chains = s.use_aux && v && atoi(v) != 0; }
  { cudaStream_t ss = s.fork2(); session_backward(ss); }
  // ---- C: CAR layer 2 of the candidate rows (shared weights: the clicked rows follow once S has produced their dE)
  { cudaStream_t st = s.fork(); s.wgrad(H1c, C, dEc, C, c.off_W2, C, C, C, Rc, st); s.bgrad(dEc, C, Rc, C, c.off_b2, st); }
  float* dH1c = c.dedup ? sb.dH1 : sb.dH1 + L * C;
  s.dgrad(dEc, C, c.off_W2, C, dH1c, C, Rc, C, C, NAR_ACT_LEAKY_RELU, H1c, C, 0, main);
  float* DBin = sb.DB; float* DBpp = sb.DB + L * C; float* DBpi = sb.DB + 2 * L * C; float* DBpc = sb.DB + NB * C;
  if (c.dedup) s.chk(nar_car_segsum(sb.dH1, L, K, C, U, pb.Mt, pb.ld_mt, io->pos_idx, pb.neg_uidx, DBpp, DBpc, DBpi, main));
  s.join2();                                       // dE[0:L] is final
  { cudaStream_t st = s.fork(); s.wgrad(sb.H1, C, sb.dE, C, c.off_W2, C, C, C, L, st); s.bgrad(sb.dE, C, L, C, c.off_b2, st); }
  if (c.dedup) {
    s.dgrad(sb.dE, C, c.off_W2, C, DBin, C, L, C, C, NAR_ACT_LEAKY_RELU, sb.H1, C, 0, main);
    {
      cudaStream_t st = s.fork();
      s.wgrad(sb.X, Fp, sb.DB, C, c.off_W1, C, Fp, C, NB, st);                                    // clicked + positive + unique item rows
      s.wgrad(sb.X + c0, Fp, DBpc, C, c.off_W1 + c0 * C, C, Fp - c0, C, L, st);                    // context block of the negatives
      s.bgrad(sb.DB, C, NB, C, c.off_b1, st);
    }
    s.dgrad(sb.DB, C, c.off_W1, C, sb.dX, Fp, NB, Fp, C, NAR_ACT_NONE, nullptr, 0, 0, main);
    // the negatives' context gradient lands on the clicked row of the same position (identical raw context features)
    s.dgrad(DBpc, C, c.off_W1 + c0 * C, C, sb.dX + c0, Fp, L, Fp - c0, C, NAR_ACT_NONE, nullptr, 0, 1, main);
  } else {
    s.dgrad(sb.dE, C, c.off_W2, C, sb.dH1, C, L, C, C, NAR_ACT_LEAKY_RELU, sb.H1, C, 0, main);
    { cudaStream_t st = s.fork(); s.wgrad(sb.X, Fp, sb.dH1, C, c.off_W1, C, Fp, C, R, st); s.bgrad(sb.dH1, C, R, C, c.off_b1, st); }
    s.dgrad(sb.dH1, C, c.off_W1, C, sb.dX, Fp, R, Fp, C, NAR_ACT_NONE, nullptr, 0, 0, main);
    if (drop) dropout(sb.dX, sb.dX, R, Fp, pb.row_pos, 0, main);
  }
  s.chk(nar_gather_features_bwd(e->ctx, &plan, g_pos, g_item, &rl, io->event_ts, io->max_ts, sb.dX, s.G(c.off_gamma),
                                s.G(c.off_beta), main));
  s.join();
  return s.rc;
}

}  // namespace

void planes_build(nar_engine* e) {
  const nar_model_cfg& c = e->cfg;
  PlaneSet& ps = e->planes;
  ps.n = 0;
  int64_t total = 0;
  auto add = [&](int64_t off, int64_t K, int64_t N, int64_t ldw) {
    if (ps.n >= MAX_PLANES) return;
    const int i = ps.n++;
    ps.off_W[i] = off; ps.K[i] = (int32_t)K; ps.N[i] = (int32_t)N; ps.ldw[i] = (int32_t)ldw;
    ps.ld_out[i] = (int32_t)((K + 31) / 32 * 64);
    ps.dst[i] = total;
    total += align_up((int64_t)N * ps.ld_out[i], 128);
  };
  const int64_t C = c.C, Hp = c.Hp, Fp = c.Fp, c0 = c.ctx_col0;
  add(c.off_W1, Fp, C, C); add(c.off_W1, c0, C, C); add(c.off_W1 + c0 * C, Fp - c0, C, C);
  add(c.off_W2, C, C, C); add(c.off_W3, Hp, 512, 512); add(c.off_W4, 512, C, C);
  add(c.off_M[0], C, 128, c.ld_M[0]); add(c.off_M[1], 128, 64, c.ld_M[1]); add(c.off_M[2], 64, 32, c.ld_M[2]);
  for (int i = 0; i < c.layers; ++i) {
    const int64_t n_in = i == 0 ? C : Hp;
    add(c.off_Wx[i], n_in, 2 * Hp, 2 * Hp);
    if (c.rnn_cell == 1) add(c.off_Wxc[i], n_in, Hp, Hp);
  }
  if (!ps.buf) {
    cudaMalloc(&ps.buf, (size_t)total * sizeof(uint16_t));
    cudaMemset(ps.buf, 0, (size_t)total * sizeof(uint16_t));
    cudaMalloc(&ps.descs, MAX_PLANES * 32);
  }
  for (int i = 0; i < ps.n; ++i) { ps.W[i] = c.params + ps.off_W[i]; ps.out[i] = ps.buf + ps.dst[i]; }
}

int planes_refresh(nar_engine* e, cudaStream_t st) {
  PlaneSet& ps = e->planes;
  if (!ps.buf || ps.n == 0) return NAR_ERR_INVALID;
  ++e->launches;
  return nar_pack_bf16x3(ps.W, ps.out, ps.K, ps.N, ps.ldw, ps.ld_out, ps.n, ps.descs, st);
}

// ================================================================================================ C ABI
extern "C" int nar_engine_create(nar_ctx* ctx, const nar_model_cfg* cfg, nar_engine** out) {
  if (!ctx || !cfg || !out) return NAR_ERR_INVALID;
  *out = nullptr;
  if (cfg->layers < 1 || cfg->layers > NAR_MAX_LAYERS || cfg->rnn_cell < 0 || cfg->rnn_cell > 1 || cfg->ranking < 0 || cfg->ranking > 1)
    return NAR_ERR_UNSUPPORTED;
  if ((cfg->C & 3) || (cfg->Hp & 3) || (cfg->Fp & 3) || (cfg->ctx_col0 & 3) || cfg->ctx_col0 <= 0 || cfg->ctx_col0 >= cfg->Fp)
    return NAR_ERR_INVALID;
  if (!cfg->params || !cfg->grads || !cfg->adam_m || !cfg->adam_v || !cfg->params_lo) return NAR_ERR_INVALID;
  nar_engine* e = new nar_engine();
  memset(e, 0, sizeof(*e));
  e->ctx = ctx; e->cfg = *cfg;
  NAR_CHECK_CUDA(cudaSetDevice(ctx->device));
  if (cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->aux2, cudaStreamNonBlocking) != cudaSuccess) { delete e; return NAR_ERR_NO_DEVICE; }
  for (int i = 0; i < N_EVENTS; ++i)
    if (cudaEventCreateWithFlags(&e->ev[i], cudaEventDisableTiming) != cudaSuccess) { delete e; return NAR_ERR_NO_DEVICE; }
  for (int i = 0; i < cfg->layers; ++i) {
    if (cudaMalloc(&e->WhT[i], (size_t)2 * cfg->Hp * cfg->Hp * sizeof(float)) != cudaSuccess) { delete e; return NAR_ERR_NO_DEVICE; }
    if (cfg->rnn_cell == 1 && cudaMalloc(&e->WhcT[i], (size_t)cfg->Hp * cfg->Hp * sizeof(float)) != cudaSuccess) { delete e; return NAR_ERR_NO_DEVICE; }
  }
  planes_build(e);
  if (!e->planes.buf || !e->planes.descs) { delete e; return NAR_ERR_NO_DEVICE; }
  *out = e;
  return NAR_OK;
}

extern "C" int nar_engine_refresh(nar_engine* e, void* stream) {
  if (!e) return NAR_ERR_INVALID;
  return planes_refresh(e, as_stream(stream));
}

extern "C" int nar_engine_destroy(nar_engine* e) {
  if (!e) return NAR_OK;
  cudaStreamSynchronize(e->aux);
  if (e->aux2) { cudaStreamSynchronize(e->aux2); cudaStreamDestroy(e->aux2); }
  for (int i = 0; i < NAR_MAX_LAYERS; ++i) { if (e->WhT[i]) cudaFree(e->WhT[i]); if (e->WhcT[i]) cudaFree(e->WhcT[i]); }
  for (int i = 0; i < N_EVENTS; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->aux) cudaStreamDestroy(e->aux);
  if (e->planes.buf) cudaFree(e->planes.buf);
  if (e->planes.descs) cudaFree(e->planes.descs);
  delete e;
  return NAR_OK;
}

extern "C" int nar_engine_update_cfg(nar_engine* e, const nar_model_cfg* cfg) {
  if (!e || !cfg) return NAR_ERR_INVALID;
  if (cfg->layers != e->cfg.layers || cfg->Hp != e->cfg.Hp || cfg->C != e->cfg.C || cfg->Fp != e->cfg.Fp ||
      cfg->rnn_cell != e->cfg.rnn_cell) return NAR_ERR_INVALID;        // structural changes need a new engine
  e->cfg = *cfg;
  planes_build(e);                      // the parameter buffer may have changed (share_params)
  return NAR_OK;
}

extern "C" int nar_engine_workspace_bytes(const nar_engine* e, int64_t Bg, int64_t B, int64_t T, int64_t L_cap, int32_t train,
                                          int64_t* prep_bytes, int64_t* ws_bytes) {
  if (!e || Bg <= 0 || B <= 0 || T <= 0 || L_cap < 0) return NAR_ERR_INVALID;
  PrepBufs pb; StepBufs sb;
  if (prep_bytes) *prep_bytes = prep_carve(e, Bg, B, T, L_cap, nullptr, &pb);
  if (ws_bytes) *ws_bytes = step_carve(e, L_cap, train, nullptr, &sb);
  return NAR_OK;
}

extern "C" int nar_engine_prepare(nar_engine* e, const nar_step_io* io, void* stream) {
  if (!e || !io || !io->prep_ws || !io->all_items || !io->buffer) return NAR_ERR_INVALID;
  const nar_model_cfg& c = e->cfg;
  const int64_t B = io->B, Bg = io->Bg, T = io->T, L = io->L, K = c.K, n_cand = K + 1, R = L + L * n_cand;
  if (L > io->L_cap || io->sess0 < 0 || io->sess0 + B > Bg) return NAR_ERR_INVALID;
  PrepBufs pb;
  if (prep_carve(e, Bg, B, T, io->L_cap, io->prep_ws, &pb) > io->prep_ws_bytes) return NAR_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int64_t* neg_local = pb.neg + io->sess0 * T * K;
  int32_t* uidx_local = c.dedup ? pb.neg_uidx + io->sess0 * T * K : nullptr;
  const int64_t* uitems = nullptr; const int32_t* n_unique = nullptr;
  int rc = nar_sample_negatives_uidx(e->ctx, io->all_items, Bg, T + 1, io->sess0, B, io->buffer, c.buf_len, K, c.n_from_buffer,
                                     c.sampler_seed, io->sampler_step, neg_local, uidx_local, &uitems, &n_unique, pb.sampler_ws,
                                     pb.sampler_bytes, st);
  e->launches += 2;
  if (rc) return rc;
  if (L <= 0) return NAR_OK;
  rc = nar_build_rows(io->pos_idx, L, io->item_clicked, io->label_next, pb.neg, K, pb.row_pos, pb.row_item, st);
  ++e->launches;
  if (rc) return rc;
  rc = nar_feature_stats(e->ctx, io->buffer, c.buf_len, c.n_norm, c.plan.created_at_ts, io->pop_norm, io->max_ts,
                         c.plan.log_base_recency, c.plan.log_base_novelty, pb.row_pos, pb.row_item, R, L, n_cand, io->event_ts,
                         pb.stats, st);
  ++e->launches;
  if (rc) return rc;
  if (c.dedup) {
    rc = nar_build_base_rows(io->pos_idx, L, io->item_clicked, io->label_next, uitems, n_unique, pb.U, pb.neg_uidx, K, pb.base_pos,
                             pb.base_item, pb.Mt, pb.ld_mt, st);
    e->launches += 2;
  }
  return rc;
}

extern "C" int nar_engine_step(nar_engine* e, const nar_step_io* io, void* stream) {
  if (!e || !io || !io->prep_ws || !io->ws || !io->loss) return NAR_ERR_INVALID;
  return run_step(e, io, as_stream(stream));
}

extern "C" int nar_engine_apply(nar_engine* e, const nar_step_io* io, void* stream) {
  if (!e || !io) return NAR_ERR_INVALID;
  const nar_model_cfg& c = e->cfg;
  ++e->launches;
  int rc = nar_adam_tf(c.params, c.grads, c.adam_m, c.adam_v, c.n_params, c.reg_end, c.reg_l2, c.lr, c.beta1, c.beta2, c.eps,
                       io->global_step + 1, c.params_lo, stream);
  if (rc == NAR_OK && c.fwd_precision == 4) rc = planes_refresh(e, as_stream(stream));
  return rc;
}

extern "C" int64_t nar_engine_launch_count(const nar_engine* e) { return e ? e->launches : 0; }

extern "C" int nar_engine_buffer(const nar_engine* e, const nar_step_io* io, const char* name, void** ptr, int64_t* rows, int64_t* ld) {
  if (!e || !io || !name || !ptr) return NAR_ERR_INVALID;
  const nar_model_cfg& c = e->cfg;
  const int64_t L = io->L, K = c.K, n_cand = K + 1, Rc = L * n_cand, R = L + Rc;
  PrepBufs pb; StepBufs sb;
  prep_carve(e, io->Bg, io->B, io->T, io->L_cap, io->prep_ws, &pb);
  step_carve(e, io->L_cap, io->train, io->ws, &sb);
  const int64_t NB = 2 * L + pb.U;
  struct Ent { const char* n; void* p; int64_t r, l; };
  const Ent tab[] = {
      {"neg", pb.neg, io->Bg * io->T, K}, {"neg_uidx", pb.neg_uidx, io->Bg * io->T, K}, {"stats", pb.stats, 1, 24},
      {"row_pos", pb.row_pos, R, 1}, {"row_item", pb.row_item, R, 1}, {"base_pos", pb.base_pos, NB, 1},
      {"base_item", pb.base_item, NB, 1}, {"Mt", pb.Mt, pb.U, pb.ld_mt},
      {"X", sb.X, c.dedup ? NB : R, c.Fp}, {"dX", sb.dX, c.dedup ? NB : R, c.Fp}, {"H1", sb.H1, R, c.C}, {"E", sb.E, R, c.C},
      {"dE", sb.dE, R, c.C}, {"dH1", sb.dH1, c.dedup ? Rc : R, c.C}, {"F1", sb.F1, L, 512}, {"PR", sb.PR, L, c.C},
      {"logits", sb.logits, L, n_cand}, {"PD", sb.PD, Rc, c.C}, {"Z1", sb.Z1, Rc, 128}, {"Z2", sb.Z2, Rc, 64}, {"Z3", sb.Z3, Rc, 32}, {"PP", sb.PP, L, c.C},
      {"PI", sb.PI, pb.U, c.C}, {"PC", sb.PC, L, c.C}, {"DB", sb.DB, 3 * L + pb.U, c.C},
      {"HO0", sb.HO[0], L, c.Hp}, {"HO1", sb.HO[1], L, c.Hp}, {"HO2", sb.HO[2], L, c.Hp}, {"HO3", sb.HO[3], L, c.Hp},
      {"HOd0", sb.HOd[0], L, c.Hp}, {"HOd1", sb.HOd[1], L, c.Hp}, {"HOd2", sb.HOd[2], L, c.Hp}, {"HOd3", sb.HOd[3], L, c.Hp},
      {"GX0", sb.GX[0], L, 2 * c.Hp}, {"dGX0", sb.dGX[0], L, 2 * c.Hp}};
  for (const Ent& t : tab)
    if (strcmp(t.n, name) == 0) {
      *ptr = t.p;
      if (rows) *rows = t.r;
      if (ld) *ld = t.l;
      return t.p ? NAR_OK : NAR_ERR_INVALID;
    }
  return NAR_ERR_INVALID;
}
