// UGRNN recurrence of the session RNN (tf.contrib.rnn.UGRNNCell inside dynamic_rnn,
// nar_model.py:1308-1342).  The input projection x*Wx + b of ALL time steps is one tcgen05 GEMM
// (nar_gemm_tf32); what is left is the sequential part, independent per session:
//     act = gx[t] + h * Wh ;  g = sigmoid(act_g + 1) ; c = tanh(act_c) ; h' = g*h + (1-g)*c
// Rows are the valid positions only (session b owns rows [sess_off[b], sess_off[b+1])), so
// "zero output / state pass-through past sequence_length" needs no work at all.
// One CTA owns SB sessions and walks their time steps; thread j owns gate column j and candidate
// column Hp+j, Wh streams from L2 (512 KB at H=256; coalesced rows).
#include "common.cuh"

namespace nar {
namespace rnn {

constexpr int SB = 4;             // sessions per CTA
constexpr int THREADS = 256;
constexpr int MAX_HP = 1024;      // h carried in shared memory: SB * Hp floats

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(THREADS)
ugrnn_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ Wh, const int32_t* __restrict__ sess_off,
                 int64_t B, int Hp, float* __restrict__ h_out, float* __restrict__ gate, float* __restrict__ cand) {
  extern __shared__ float sh[];                 // h[SB][Hp]
  float* h = sh;
  const int64_t b0 = (int64_t)blockIdx.x * SB;
  int off[SB], len[SB];
  int maxlen = 0;
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    const int64_t b = b0 + s;
    off[s] = b < B ? sess_off[b] : 0;
    len[s] = b < B ? sess_off[b + 1] - sess_off[b] : 0;
    maxlen = max(maxlen, len[s]);
  }
  for (int i = threadIdx.x; i < SB * Hp; i += THREADS) h[i] = 0.f;
  __syncthreads();
  const int W2 = 2 * Hp;
  for (int t = 0; t < maxlen; ++t) {
    for (int j = threadIdx.x; j < Hp; j += THREADS) {
      float ag[SB], ac[SB];
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (t < len[s]) {
          const float* g = gx + (int64_t)(off[s] + t) * W2;
          ag[s] = g[j]; ac[s] = g[Hp + j];
        } else { ag[s] = 0.f; ac[s] = 0.f; }
      }
      if (t > 0) {      // h == 0 at t == 0
        for (int k = 0; k < Hp; ++k) {
          const float wg = __ldg(Wh + (int64_t)k * W2 + j), wc = __ldg(Wh + (int64_t)k * W2 + Hp + j);
#pragma unroll
          for (int s = 0; s < SB; ++s) { const float hv = h[s * Hp + k]; ag[s] = fmaf(hv, wg, ag[s]); ac[s] = fmaf(hv, wc, ac[s]); }
        }
      }
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (t < len[s]) {
          const float g = sigmoidf(ag[s] + 1.0f), c = tanhf(ac[s]);
          const float hn = g * h[s * Hp + j] + (1.0f - g) * c;
          const int64_t row = (int64_t)(off[s] + t) * Hp + j;
          h_out[row] = hn; gate[row] = g; cand[row] = c;
        }
      }
    }
    __syncthreads();
    // second phase: publish h' (read back from h_out: each thread wrote its own columns)
    for (int j = threadIdx.x; j < Hp; j += THREADS) {
#pragma unroll
      for (int s = 0; s < SB; ++s)
        if (t < len[s]) h[s * Hp + j] = h_out[(int64_t)(off[s] + t) * Hp + j];
    }
    __syncthreads();
  }
}

// backward through time.  d_gx = dL/d(act) (feeds the Wx / bias / Wh wgrads and the dgrad GEMM);
// h_prev[row] = state entering the step (for dWh = h_prev^T * d_gx).
__global__ void __launch_bounds__(THREADS)
ugrnn_bwd_kernel(const float* __restrict__ d_hout, const float* __restrict__ h_out, const float* __restrict__ gate,
                 const float* __restrict__ cand, const float* __restrict__ WhT, const int32_t* __restrict__ sess_off,
                 int64_t B, int Hp, float* __restrict__ d_gx, float* __restrict__ h_prev) {
  extern __shared__ float sh[];
  float* dact = sh;                  // [SB][2Hp]
  float* dh = sh + SB * 2 * Hp;      // [SB][Hp] carried gradient wrt the state leaving step t-1
  float* gsave = dh + SB * Hp;       // [SB][Hp] dh_total * g of the current step
  const int64_t b0 = (int64_t)blockIdx.x * SB;
  int off[SB], len[SB];
  int maxlen = 0;
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    const int64_t b = b0 + s;
    off[s] = b < B ? sess_off[b] : 0;
    len[s] = b < B ? sess_off[b + 1] - sess_off[b] : 0;
    maxlen = max(maxlen, len[s]);
  }
  for (int i = threadIdx.x; i < SB * Hp; i += THREADS) dh[i] = 0.f;
  __syncthreads();
  const int W2 = 2 * Hp;
  for (int t = maxlen - 1; t >= 0; --t) {
    for (int j = threadIdx.x; j < Hp; j += THREADS) {
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        float dg_act = 0.f, dc_act = 0.f, keep = 0.f;
        if (t < len[s]) {
          const int64_t row = (int64_t)(off[s] + t) * Hp + j;
          const float dht = d_hout[row] + dh[s * Hp + j];
          const float hp = t > 0 ? h_out[row - Hp] : 0.f;
          const float g = gate[row], c = cand[row];
          dg_act = dht * (hp - c) * g * (1.0f - g);
          dc_act = dht * (1.0f - g) * (1.0f - c * c);
          keep = dht * g;
          d_gx[(int64_t)(off[s] + t) * W2 + j] = dg_act;
          d_gx[(int64_t)(off[s] + t) * W2 + Hp + j] = dc_act;
          h_prev[row] = hp;
        }
        dact[s * W2 + j] = dg_act;
        dact[s * W2 + Hp + j] = dc_act;
        gsave[s * Hp + j] = keep;
      }
    }
    __syncthreads();
    if (t > 0) {
      for (int k = threadIdx.x; k < Hp; k += THREADS) {
        float acc[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) acc[s] = gsave[s * Hp + k];
        for (int j = 0; j < W2; ++j) {
          const float w = __ldg(WhT + (int64_t)j * Hp + k);
#pragma unroll
          for (int s = 0; s < SB; ++s) acc[s] = fmaf(dact[s * W2 + j], w, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) dh[s * Hp + k] = (t < len[s]) ? acc[s] : dh[s * Hp + k];
      }
    }
    __syncthreads();
  }
}

}  // namespace rnn
}  // namespace nar

extern "C" int nar_ugrnn_fwd(nar_ctx* ctx, const float* gx, const float* Wh, const int32_t* sess_off, int64_t B, int64_t Hp,
                             float* h_out, float* gate, float* cand, void* stream) {
  using namespace nar::rnn;
  if (!ctx || !gx || !Wh || !sess_off || !h_out || !gate || !cand) return NAR_ERR_INVALID;
  if (Hp <= 0 || Hp > MAX_HP) return NAR_ERR_UNSUPPORTED;
  if (B <= 0) return NAR_OK;
  const size_t smem = (size_t)SB * Hp * sizeof(float);
  ugrnn_fwd_kernel<<<(unsigned)((B + SB - 1) / SB), THREADS, smem, as_stream(stream)>>>(gx, Wh, sess_off, B, (int)Hp, h_out, gate, cand);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_ugrnn_bwd(nar_ctx* ctx, const float* d_hout, const float* h_out, const float* gate, const float* cand,
                             const float* WhT, const int32_t* sess_off, int64_t B, int64_t Hp, float* d_gx, float* h_prev,
                             void* stream) {
  using namespace nar::rnn;
  if (!ctx || !d_hout || !h_out || !gate || !cand || !WhT || !sess_off || !d_gx || !h_prev) return NAR_ERR_INVALID;
  if (Hp <= 0 || Hp > MAX_HP) return NAR_ERR_UNSUPPORTED;
  if (B <= 0) return NAR_OK;
  const size_t smem = (size_t)SB * Hp * 4 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    NAR_CHECK_CUDA(cudaFuncSetAttribute(ugrnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB * MAX_HP * 4 * 4));
    attr_set = true;
  }
  ugrnn_bwd_kernel<<<(unsigned)((B + SB - 1) / SB), THREADS, smem, as_stream(stream)>>>(d_hout, h_out, gate, cand, WhT, sess_off, B, (int)Hp, d_gx, h_prev);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}
