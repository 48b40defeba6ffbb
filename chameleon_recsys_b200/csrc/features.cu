// Feature-row assembly for the NAR hot path: the embedding gather (HBM-bound kernel the
// north_star names), its backward (gamma/beta grads + IndexedSlices scatter-add), and the
// recency / novelty normalisation statistics.
//
// Reference: nar_model.py:921-994 get_item_features, :730-773 get_features, :887-907
// scale_center_features, :996-1039 normalisation, :1055-1089 recency, :1134-1193 novelty.
#include "common.cuh"

namespace nar {
namespace feat {

constexpr float MS_PER_DAY = 1000.0f * 60.0f * 60.0f * 24.0f;

// nar_model.py:1055-1060 (int64 -> float32 BEFORE the subtraction) + log_1p :28-34
__device__ __forceinline__ float recency_raw(int64_t ts_ref, int64_t created, float inv_log_base) {
  // explicit _rn intrinsics: no FMA contraction, so that the value a row gets is bit-identical to the value
  // the statistics kernel saw (with degenerate variance, 1 ulp of difference is amplified by 1/1e-12)
  const float days = fmaxf(__fdiv_rn(__fsub_rn(__ll2float_rn(ts_ref), __ll2float_rn(created)), MS_PER_DAY), 0.f);
  return __fmul_rn(logf(__fadd_rn(days, 1.0f)), inv_log_base);
}
__device__ __forceinline__ float novelty_raw(float pop_norm, float inv_log_base) {
  return -__fmul_rn(logf(pop_norm), inv_log_base);
}
// normalize_values + min_max_normalization (:1011-1039, :996-1009); st = {mean, std, zmin, zmax}
__device__ __forceinline__ float normalize(float x, const float* st) {
  const float z = __fdiv_rn(__fsub_rn(x, st[0]), st[1]);
  const float scaled = __fdiv_rn(__fadd_rn(__fsub_rn(z, st[2]), 1e-24f), fmaxf(__fsub_rn(st[3], st[2]), 2e-24f));
  return __fsub_rn(__fmul_rn(scaled, 2.0f), 1.0f);
}

__device__ __forceinline__ int row_group(int64_t r, int64_t n_input, int64_t n_cand) {
  if (r < n_input) return 0;
  if (n_cand <= 0) return 2;
  return ((r - n_input) % n_cand) == 0 ? 1 : 2;
}

// raw (un-scaled) value of column (c - seg.col) of segment `sg` for one row
__device__ __forceinline__ float seg_value(const nar_feature_plan& P, const nar_segment& sg, int j, int64_t pos,
                                           int64_t item, int64_t ts_ref, const float* st) {
  switch (sg.kind) {
    case NAR_SEG_CTX_OHE: { const int64_t id = P.ctx_int[sg.src][pos]; return id == j ? 1.f : 0.f; }
    case NAR_SEG_CTX_EMBED: {
      int64_t id = P.ctx_int[sg.src][pos]; id = id < 0 ? 0 : (id >= sg.card ? sg.card - 1 : id);
      return sg.table[id * sg.ld + j];
    }
    case NAR_SEG_CTX_NUM: return P.ctx_float[sg.src][pos];
    case NAR_SEG_CTX_ZERO: return 0.f;
    case NAR_SEG_META_OHE: { const int64_t id = P.meta[sg.src][item]; return id == j ? 1.f : 0.f; }
    case NAR_SEG_META_EMBED: {
      int64_t id = P.meta[sg.src][item]; id = id < 0 ? 0 : (id >= sg.card ? sg.card - 1 : id);
      return sg.table[id * sg.ld + j];
    }
    case NAR_SEG_META_NUM: return (float)P.meta[sg.src][item];
    case NAR_SEG_ACR:
    case NAR_SEG_ITEM_EMB: return sg.table[item * sg.ld + j];
    case NAR_SEG_RECENCY:
      return normalize(recency_raw(ts_ref, P.created_at_ts[item], 1.0f / logf(P.log_base_recency)), st);
    case NAR_SEG_NOVELTY:
      return normalize(novelty_raw(P.pop_norm[item], 1.0f / logf(P.log_base_novelty)), st + 4);
  }
  return 0.f;
}

// ------------------------------------------------------------------ forward gather
// One warp per output row.  Design for HBM speed:
//  * the per-CTA copy of the column map / segment table lives in shared memory (divergent lookups in the
//    constant bank would serialise 32x);
//  * phase A: every scalar the row needs (context ids / floats at its position, metadata of its item,
//    created_at, pop_norm) is fetched by ONE lane each, all independent -> one memory round trip, then
//    handed to the lanes that need it with warp shuffles (no dependent load chains per column);
//  * phase B: the wide table rows (ACR, item embedding) move as 128-bit loads / streaming 128-bit stores;
//  * phase C: narrow columns, one lane per column.
constexpr int GATHER_WARPS = 8;
constexpr int LANE_CTX_INT = 0, LANE_CTX_FLOAT = 12, LANE_META = 20, LANE_CREATED = 28, LANE_POP = 29;

struct SegS { int kind, col, card, src, ld; const float* table; };

__global__ void __launch_bounds__(GATHER_WARPS * 32)
gather_features_kernel(const __grid_constant__ nar_feature_plan P, int n_ctx_int, int n_ctx_float, int n_meta,
                       const int32_t* __restrict__ row_pos,
                       const int64_t* __restrict__ row_item, int64_t n_rows, int64_t n_input, int64_t n_cand,
                       const int64_t* __restrict__ event_ts, const int64_t* __restrict__ max_ts,
                       float* __restrict__ out) {
  __shared__ uint8_t s_colseg[NAR_MAX_COLS];
  __shared__ SegS s_seg[NAR_MAX_SEGMENTS];
  for (int i = threadIdx.x; i < P.row_ld; i += blockDim.x) s_colseg[i] = P.col_seg[i];
  if (threadIdx.x < P.n_segments) {
    const nar_segment& g = P.seg[threadIdx.x];
    s_seg[threadIdx.x] = SegS{g.kind, g.col, g.card, g.src, g.ld, g.table};
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int64_t pos = row_pos[r];
  const int64_t item = row_item[r];
  // ---- phase A: scalars, one lane each
  long long mine = 0;
  if (lane < LANE_CTX_FLOAT) { if (lane < n_ctx_int) mine = P.ctx_int[lane][pos]; }
  else if (lane < LANE_META) { if (lane - LANE_CTX_FLOAT < n_ctx_float) mine = (long long)__float_as_int(P.ctx_float[lane - LANE_CTX_FLOAT][pos]); }
  else if (lane < LANE_CREATED) { if (lane - LANE_META < n_meta) mine = P.meta[lane - LANE_META][item]; }
  else if (lane == LANE_CREATED) mine = P.created_at_ts[item];
  else if (lane == LANE_POP) mine = (long long)__float_as_int(P.pop_norm[item]);
  const int64_t ts_ref = (r < n_input) ? event_ts[pos] : max_ts[0];
  const float* st = P.stats + 8 * row_group(r, n_input, n_cand);
  float* orow = out + r * (int64_t)P.row_ld;
  // ---- phase B: wide table rows
  for (int s = 0; s < P.n_segments; ++s) {
    const SegS sg = s_seg[s];
    if (sg.kind != NAR_SEG_ACR && sg.kind != NAR_SEG_ITEM_EMB) continue;
    const int width = P.seg[s].width;
    const float* src = sg.table + item * (int64_t)sg.ld;
    const bool vec = ((sg.col & 3) == 0) && ((sg.ld & 3) == 0) && ((P.row_ld & 3) == 0);
    int j0 = 0;
    if (vec) {
      const int nv = width >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      const float4* g4 = reinterpret_cast<const float4*>(P.gamma + sg.col);
      const float4* b4 = reinterpret_cast<const float4*>(P.beta + sg.col);
      float4* o4 = reinterpret_cast<float4*>(orow + sg.col);
      for (int j = lane; j < nv; j += 32) {
        float4 v = __ldg(s4 + j);
        const float4 g = __ldg(g4 + j), b = __ldg(b4 + j);
        v.x = v.x * g.x + b.x; v.y = v.y * g.y + b.y; v.z = v.z * g.z + b.z; v.w = v.w * g.w + b.w;
        __stcs(o4 + j, v);                          // streaming store: the row is consumed once by the GEMM's TMA
      }
      j0 = nv << 2;
    }
    for (int j = j0 + lane; j < width; j += 32)
      orow[sg.col + j] = __ldg(src + j) * P.gamma[sg.col + j] + P.beta[sg.col + j];
  }
  // ---- phase C: narrow columns (one-hot, small embeddings, numerics, recency, novelty, padding)
  const float ilr = 1.0f / logf(P.log_base_recency), iln = 1.0f / logf(P.log_base_novelty);
  for (int q = 0; q < P.n_narrow; ++q) {
    const int begin = P.narrow_begin[q], end = P.narrow_end[q];
    for (int c0 = begin; c0 < end; c0 += 32) {             // warp-uniform trip count (shuffles inside)
      const int c = c0 + lane;
      const bool valid = c < end;
      const int si = valid ? s_colseg[c] : 255;
      SegS sg = s_seg[si == 255 ? 0 : si];
      int src_lane = lane;
      switch (sg.kind) {
        case NAR_SEG_CTX_OHE: case NAR_SEG_CTX_EMBED: src_lane = LANE_CTX_INT + sg.src; break;
        case NAR_SEG_CTX_NUM: src_lane = LANE_CTX_FLOAT + sg.src; break;
        case NAR_SEG_META_OHE: case NAR_SEG_META_EMBED: case NAR_SEG_META_NUM: src_lane = LANE_META + sg.src; break;
        case NAR_SEG_RECENCY: src_lane = LANE_CREATED; break;
        case NAR_SEG_NOVELTY: src_lane = LANE_POP; break;
        default: break;
      }
      const long long val = __shfl_sync(0xffffffffu, mine, src_lane);
      if (!valid) continue;
      float v = 0.f;
      if (si != 255) {
        const int j = c - sg.col;
        float raw = 0.f;
        switch (sg.kind) {
          case NAR_SEG_CTX_OHE: case NAR_SEG_META_OHE: raw = (val == (long long)j) ? 1.f : 0.f; break;
          case NAR_SEG_CTX_EMBED: case NAR_SEG_META_EMBED: {
            const long long id = val < 0 ? 0 : (val >= sg.card ? sg.card - 1 : val);
            raw = __ldg(sg.table + id * sg.ld + j);
          } break;
          case NAR_SEG_CTX_NUM: raw = __int_as_float((int)val); break;
          case NAR_SEG_META_NUM: raw = (float)val; break;
          case NAR_SEG_RECENCY: raw = normalize(recency_raw(ts_ref, (int64_t)val, ilr), st); break;
          case NAR_SEG_NOVELTY: raw = normalize(novelty_raw(__int_as_float((int)val), iln), st + 4); break;
          default: break;
        }
        v = raw * P.gamma[c] + P.beta[c];
      }
      orow[c] = v;
    }
  }
}

// ------------------------------------------------------------------ backward
// thread per column, CTA per block of rows: d_beta[c] += sum dX ; d_gamma[c] += sum dX*raw ;
// trainable embeddings: grad[id, j] += dX*gamma (scatter-add of the IndexedSlices gradient)
constexpr int BWD_ROWS = 32;
constexpr int BWD_THREADS = 256;

__global__ void __launch_bounds__(BWD_THREADS)
gather_features_bwd_kernel(const __grid_constant__ nar_feature_plan P, const int32_t* __restrict__ row_pos,
                           const int64_t* __restrict__ row_item, int64_t n_rows, int64_t n_input, int64_t n_cand,
                           const int64_t* __restrict__ event_ts, const int64_t* __restrict__ max_ts,
                           const float* __restrict__ d_out, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ int64_t s_pos[BWD_ROWS], s_item[BWD_ROWS], s_ts[BWD_ROWS];
  __shared__ int s_grp[BWD_ROWS];
  const int64_t r0 = (int64_t)blockIdx.x * BWD_ROWS;
  const int nr = (int)min((int64_t)BWD_ROWS, n_rows - r0);
  if (threadIdx.x < nr) {
    const int64_t r = r0 + threadIdx.x;
    const int64_t pos = row_pos[r];
    s_pos[threadIdx.x] = pos;
    s_item[threadIdx.x] = row_item[r];
    s_ts[threadIdx.x] = (r < n_input) ? event_ts[pos] : max_ts[0];
    s_grp[threadIdx.x] = row_group(r, n_input, n_cand);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < P.row_ld; c += BWD_THREADS) {
    const int si = P.col_seg[c];
    if (si == 255) continue;
    const nar_segment& sg = P.seg[si];
    const int j = c - sg.col;
    const float gam = P.gamma[c];
    float acc_b = 0.f, acc_g = 0.f;
    for (int i = 0; i < nr; ++i) {
      const float d = d_out[(r0 + i) * (int64_t)P.row_ld + c];
      const float raw = seg_value(P, sg, j, s_pos[i], s_item[i], s_ts[i], P.stats + 8 * s_grp[i]);
      acc_b += d;
      acc_g += d * raw;
      if (sg.grad != nullptr) {
        int64_t id;
        if (sg.kind == NAR_SEG_ITEM_EMB) id = s_item[i];
        else if (sg.kind == NAR_SEG_CTX_EMBED) id = P.ctx_int[sg.src][s_pos[i]];
        else id = P.meta[sg.src][s_item[i]];
        if (sg.kind != NAR_SEG_ITEM_EMB) id = id < 0 ? 0 : (id >= sg.card ? sg.card - 1 : id);
        atomicAdd(sg.grad + id * (int64_t)sg.ld + j, d * gam);
      }
    }
    atomicAdd(d_beta + c, acc_b);
    atomicAdd(d_gamma + c, acc_g);
  }
}

// ------------------------------------------------------------------ statistics
// single CTA.  Pass A: the first n_norm nonzero buffer entries (order kept).  If the buffer is
// empty: each row group uses its own non-padded rows (first batch only).
constexpr int STATS_THREADS = 1024;

struct Acc { float sum, mn, mx; float cnt; };

__device__ float block_reduce(float v, int op, float* sh) {   // op 0 sum, 1 min, 2 max
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = op == 0 ? warp_sum(v) : (op == 1 ? warp_min(v) : warp_max(v));
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : (op == 0 ? 0.f : (op == 1 ? INFINITY : -INFINITY));
    x = op == 0 ? warp_sum(x) : (op == 1 ? warp_min(x) : warp_max(x));
    if (lane == 0) sh[32] = x;
  }
  __syncthreads();
  return sh[32];
}

__global__ void __launch_bounds__(STATS_THREADS)
feature_stats_kernel(const int64_t* __restrict__ buffer, int64_t buf_len, int64_t n_norm,
                     const int64_t* __restrict__ created, const float* __restrict__ pop_norm,
                     const int64_t* __restrict__ max_ts_p, float log_base_rec, float log_base_nov,
                     const int32_t* __restrict__ row_pos, const int64_t* __restrict__ row_item, int64_t n_rows,
                     int64_t n_input, int64_t n_cand, const int64_t* __restrict__ event_ts, float* __restrict__ stats) {
  __shared__ float sh[40];
  __shared__ int s_cnt[STATS_THREADS];
  __shared__ int s_total;
  const int t = threadIdx.x;
  const float ilr = 1.0f / logf(log_base_rec), iln = 1.0f / logf(log_base_nov);
  const int64_t max_ts = max_ts_p[0];
  // contiguous chunk per thread so that ranks follow buffer order
  const int64_t chunk = (buf_len + STATS_THREADS - 1) / STATS_THREADS;
  const int64_t lo = min(buf_len, (int64_t)t * chunk), hi = min(buf_len, lo + chunk);
  int c = 0;
  for (int64_t i = lo; i < hi; ++i) c += buffer[i] != 0;
  s_cnt[t] = c;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < STATS_THREADS; ++i) { const int x = s_cnt[i]; s_cnt[i] = run; run += x; }
    s_total = run;
  }
  __syncthreads();
  const int total = s_total;
  if (total > 0) {
    const int64_t n_use = min((int64_t)total, n_norm);
    // pass 1: sums
    float sr = 0.f, sn = 0.f, mnr = INFINITY, mxr = -INFINITY, mnn = INFINITY, mxn = -INFINITY;
    int rank = s_cnt[t];
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t id = buffer[i];
      if (id == 0) continue;
      if (rank < n_use) {
        const float a = recency_raw(max_ts, created[id], ilr), b = novelty_raw(pop_norm[id], iln);
        sr += a; sn += b; mnr = fminf(mnr, a); mxr = fmaxf(mxr, a); mnn = fminf(mnn, b); mxn = fmaxf(mxn, b);
      }
      ++rank;
    }
    const float inv_n = 1.0f / (float)n_use;
    const float mean_r = block_reduce(sr, 0, sh) * inv_n;
    const float mean_n = block_reduce(sn, 0, sh) * inv_n;
    const float min_r = block_reduce(mnr, 1, sh), max_r = block_reduce(mxr, 2, sh);
    const float min_n = block_reduce(mnn, 1, sh), max_n = block_reduce(mxn, 2, sh);
    float vr = 0.f, vn = 0.f;
    rank = s_cnt[t];
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t id = buffer[i];
      if (id == 0) continue;
      if (rank < n_use) {
        const float a = recency_raw(max_ts, created[id], ilr) - mean_r, b = novelty_raw(pop_norm[id], iln) - mean_n;
        vr += a * a; vn += b * b;
      }
      ++rank;
    }
    const float var_r = block_reduce(vr, 0, sh) * inv_n, var_n = block_reduce(vn, 0, sh) * inv_n;
    if (t == 0) {
      const float sd_r = sqrtf(var_r + 1e-24f), sd_n = sqrtf(var_n + 1e-24f);
      for (int g = 0; g < 3; ++g) {
        float* s = stats + 8 * g;
        s[0] = mean_r; s[1] = sd_r; s[2] = (min_r - mean_r) / sd_r; s[3] = (max_r - mean_r) / sd_r;
        s[4] = mean_n; s[5] = sd_n; s[6] = (min_n - mean_n) / sd_n; s[7] = (max_n - mean_n) / sd_n;
      }
    }
    return;
  }
  // ---- empty buffer: statistics of each row group over its own non-padded rows
  for (int g = 0; g < 3; ++g) {
    float sr = 0.f, sn = 0.f, mnr = INFINITY, mxr = -INFINITY, mnn = INFINITY, mxn = -INFINITY, cnt = 0.f;
    for (int64_t r = t; r < n_rows; r += STATS_THREADS) {
      if (row_group(r, n_input, n_cand) != g) continue;
      const int64_t id = row_item[r];
      if (id == 0) continue;
      const int64_t ts = (r < n_input) ? event_ts[row_pos[r]] : max_ts;
      const float a = recency_raw(ts, created[id], ilr), b = novelty_raw(pop_norm[id], iln);
      sr += a; sn += b; mnr = fminf(mnr, a); mxr = fmaxf(mxr, a); mnn = fminf(mnn, b); mxn = fmaxf(mxn, b); cnt += 1.f;
    }
    const float n = block_reduce(cnt, 0, sh);
    const float inv_n = n > 0.f ? 1.0f / n : 0.f;
    const float mean_r = block_reduce(sr, 0, sh) * inv_n, mean_n = block_reduce(sn, 0, sh) * inv_n;
    const float min_r = block_reduce(mnr, 1, sh), max_r = block_reduce(mxr, 2, sh);
    const float min_n = block_reduce(mnn, 1, sh), max_n = block_reduce(mxn, 2, sh);
    float vr = 0.f, vn = 0.f;
    for (int64_t r = t; r < n_rows; r += STATS_THREADS) {
      if (row_group(r, n_input, n_cand) != g) continue;
      const int64_t id = row_item[r];
      if (id == 0) continue;
      const int64_t ts = (r < n_input) ? event_ts[row_pos[r]] : max_ts;
      const float a = recency_raw(ts, created[id], ilr) - mean_r, b = novelty_raw(pop_norm[id], iln) - mean_n;
      vr += a * a; vn += b * b;
    }
    const float var_r = block_reduce(vr, 0, sh) * inv_n, var_n = block_reduce(vn, 0, sh) * inv_n;
    if (t == 0) {
      const float sd_r = sqrtf(var_r + 1e-24f), sd_n = sqrtf(var_n + 1e-24f);
      float* s = stats + 8 * g;
      s[0] = mean_r; s[1] = sd_r; s[2] = (min_r - mean_r) / sd_r; s[3] = (max_r - mean_r) / sd_r;
      s[4] = mean_n; s[5] = sd_n; s[6] = (min_n - mean_n) / sd_n; s[7] = (max_n - mean_n) / sd_n;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ plain row gather / scatter-add
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, int64_t n_table_rows, int64_t ld, int width,
                   const int64_t* __restrict__ ids, int64_t n, float* __restrict__ out, int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  int64_t id = ids[r];
  id = id < 0 ? 0 : (id >= n_table_rows ? n_table_rows - 1 : id);
  const float* src = table + id * ld;
  float* dst = out + r * ld_out;
  const bool vec = ((ld & 3) == 0) && ((ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(table) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  int j0 = 0;
  if (vec) {
    const int nv = width >> 2;
    for (int j = lane; j < nv; j += 32) reinterpret_cast<float4*>(dst)[j] = __ldg(reinterpret_cast<const float4*>(src) + j);
    j0 = nv << 2;
  }
  for (int j = j0 + lane; j < width; j += 32) dst[j] = __ldg(src + j);
}

__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(float* __restrict__ table, int64_t n_table_rows, int64_t ld, int width,
                        const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ src, int64_t ld_src) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  int64_t id = ids[r];
  if (id < 0 || id >= n_table_rows) return;
  for (int j = lane; j < width; j += 32) atomicAdd(table + id * ld + j, src[r * ld_src + j]);
}

__global__ void __launch_bounds__(256)
build_rows_kernel(const int32_t* __restrict__ pos_idx, int64_t L, const int64_t* __restrict__ item_clicked,
                  const int64_t* __restrict__ label_next, const int64_t* __restrict__ negatives, int64_t K,
                  int32_t* __restrict__ row_pos, int64_t* __restrict__ row_item) {
  const int64_t n_cand = K + 1;
  const int64_t total = L * (n_cand + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = i / (n_cand + 1), j = i - l * (n_cand + 1);     // j = 0 input, 1 positive, 2.. negatives
    const int32_t pos = pos_idx[l];
    int64_t r, item;
    if (j == 0) { r = l; item = item_clicked[pos]; }
    else if (j == 1) { r = L + l * n_cand; item = label_next[pos]; }
    else { r = L + l * n_cand + (j - 1); item = negatives[(int64_t)pos * K + (j - 2)]; }
    row_pos[r] = pos;
    row_item[r] = item;
  }
}

}  // namespace feat
}  // namespace nar

extern "C" int nar_build_rows(const int32_t* pos_idx, int64_t L, const int64_t* item_clicked, const int64_t* label_next_item,
                              const int64_t* negatives, int64_t K, int32_t* row_pos, int64_t* row_item, void* stream) {
  if (!pos_idx || !item_clicked || !label_next_item || !negatives || !row_pos || !row_item || K < 0) return NAR_ERR_INVALID;
  if (L <= 0) return NAR_OK;
  const int64_t total = L * (K + 2);
  int64_t g = (total + 255) / 256; if (g > 148 * 8) g = 148 * 8;
  nar::feat::build_rows_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(pos_idx, L, item_clicked, label_next_item, negatives, K, row_pos, row_item);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_gather_features(nar_ctx* ctx, const nar_feature_plan* plan, const int32_t* row_pos,
                                   const int64_t* row_item, int64_t n_rows, int64_t n_input, int64_t n_cand,
                                   const int64_t* event_timestamp, const int64_t* max_ts, float* out, void* stream) {
  if (!ctx || !plan || !row_pos || !row_item || !out || !max_ts) return NAR_ERR_INVALID;
  if (plan->n_segments > NAR_MAX_SEGMENTS) return NAR_ERR_INVALID;
  if (n_rows <= 0) return NAR_OK;
  const unsigned grid = (unsigned)((n_rows + nar::feat::GATHER_WARPS - 1) / nar::feat::GATHER_WARPS);
  // lane budget of the scalar prefetch (phase A): <= 12 context ids, <= 8 context floats, <= 8 metadata arrays
  int n_ci = 0, n_cf = 0, n_me = 0;
  for (int i = 0; i < plan->n_segments; ++i) {
    const nar_segment& g = plan->seg[i];
    if (g.kind == NAR_SEG_CTX_OHE || g.kind == NAR_SEG_CTX_EMBED) n_ci = g.src + 1 > n_ci ? g.src + 1 : n_ci;
    if (g.kind == NAR_SEG_CTX_NUM) n_cf = g.src + 1 > n_cf ? g.src + 1 : n_cf;
    if (g.kind == NAR_SEG_META_OHE || g.kind == NAR_SEG_META_EMBED || g.kind == NAR_SEG_META_NUM) n_me = g.src + 1 > n_me ? g.src + 1 : n_me;
  }
  if (n_ci > 12 || n_cf > 8 || n_me > 8 || plan->row_ld > NAR_MAX_COLS) return NAR_ERR_UNSUPPORTED;
  nar::feat::gather_features_kernel<<<grid, nar::feat::GATHER_WARPS * 32, 0, as_stream(stream)>>>(
      *plan, n_ci, n_cf, n_me, row_pos, row_item, n_rows, n_input, n_cand, event_timestamp, max_ts, out);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_gather_features_bwd(nar_ctx* ctx, const nar_feature_plan* plan, const int32_t* row_pos,
                                       const int64_t* row_item, int64_t n_rows, int64_t n_input, int64_t n_cand,
                                       const int64_t* event_timestamp, const int64_t* max_ts, const float* d_out,
                                       float* d_gamma, float* d_beta, void* stream) {
  if (!ctx || !plan || !row_pos || !row_item || !d_out || !d_gamma || !d_beta) return NAR_ERR_INVALID;
  if (n_rows <= 0) return NAR_OK;
  const unsigned grid = (unsigned)((n_rows + nar::feat::BWD_ROWS - 1) / nar::feat::BWD_ROWS);
  nar::feat::gather_features_bwd_kernel<<<grid, nar::feat::BWD_THREADS, 0, as_stream(stream)>>>(
      *plan, row_pos, row_item, n_rows, n_input, n_cand, event_timestamp, max_ts, d_out, d_gamma, d_beta);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_feature_stats(nar_ctx* ctx, const int64_t* buffer, int64_t buf_len, int64_t n_norm,
                                 const int64_t* created_at_ts, const float* pop_norm, const int64_t* max_ts,
                                 float log_base_recency, float log_base_novelty, const int32_t* row_pos,
                                 const int64_t* row_item, int64_t n_rows, int64_t n_input, int64_t n_cand,
                                 const int64_t* event_timestamp, float* stats, void* stream) {
  if (!ctx || !buffer || !created_at_ts || !pop_norm || !max_ts || !stats) return NAR_ERR_INVALID;
  nar::feat::feature_stats_kernel<<<1, nar::feat::STATS_THREADS, 0, as_stream(stream)>>>(
      buffer, buf_len, n_norm, created_at_ts, pop_norm, max_ts, log_base_recency, log_base_novelty, row_pos,
      row_item, n_rows, n_input, n_cand, event_timestamp, stats);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_gather_rows_f32(const float* table, int64_t n_table_rows, int64_t ld, int width,
                                   const int64_t* ids, int64_t n, float* out, int64_t ld_out, void* stream) {
  if (!table || !ids || !out) return NAR_ERR_INVALID;
  if (n <= 0) return NAR_OK;
  nar::feat::gather_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(table, n_table_rows, ld, width, ids, n, out, ld_out);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}

extern "C" int nar_scatter_add_rows_f32(float* table, int64_t n_table_rows, int64_t ld, int width,
                                        const int64_t* ids, int64_t n, const float* src, int64_t ld_src, void* stream) {
  if (!table || !ids || !src) return NAR_ERR_INVALID;
  if (n <= 0) return NAR_OK;
  nar::feat::scatter_add_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(table, n_table_rows, ld, width, ids, n, src, ld_src);
  NAR_LAUNCH_CHECK();
  return NAR_OK;
}
