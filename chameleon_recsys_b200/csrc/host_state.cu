// Host-side (CPU) fast path of ClickedItemsState.update_items_state (reference clicked_items_state.py:187-250; numpy
// restatement and spec: chameleon_recsys_b200/clicked_items_state.py).  Plain C++, no CUDA: the numpy version costs
// ~0.8 ms per G1 step (boolean-mask copy of the 20000 x 2 buffer, bincount, V-long float64 division), which made the
// end-to-end loop host bound once the GPU step dropped to 1.6 ms.  One pass here: ~40 us.
#include <stdint.h>
#include <string.h>
#include "../../include/nar_b200.h"

extern "C" int nar_host_state_update(int64_t* buffer, int64_t cap, const int64_t* batch_items, const int64_t* batch_ts,
                                     int64_t n_batch, int64_t hours_ms, int64_t* scratch, int64_t* recent_pop,
                                     double* pop_norm, int64_t* articles_pop, int64_t num_items, double min_norm_pop) {
  if (!buffer || !scratch || !recent_pop || !pop_norm || !articles_pop || cap <= 0 || num_items <= 0 || n_batch < 0 ||
      (n_batch > 0 && (!batch_items || !batch_ts)))
    return NAR_ERR_INVALID;
  // _update_recently_clicked_items_buffer: batch reversed (newest click first), then the old entries whose timestamp
  // is >= min(batch ts) - hours (order kept, padding rows (ts 0) fall out unless the threshold is <= 0), clipped to
  // `cap` rows and zero padded
  int64_t min_ts = 0;
  for (int64_t i = 0; i < n_batch; ++i) min_ts = (i == 0 || batch_ts[i] < min_ts) ? batch_ts[i] : min_ts;
  const int64_t thr = min_ts - hours_ms;
  int64_t n = 0;
  for (int64_t i = n_batch - 1; i >= 0 && n < cap; --i, ++n) { scratch[2 * n] = batch_items[i]; scratch[2 * n + 1] = batch_ts[i]; }
  for (int64_t i = 0; i < cap && n < cap; ++i) {
    if (buffer[2 * i + 1] >= thr) { scratch[2 * n] = buffer[2 * i]; scratch[2 * n + 1] = buffer[2 * i + 1]; ++n; }
  }
  memcpy(buffer, scratch, (size_t)n * 2 * sizeof(int64_t));
  memset(buffer + 2 * n, 0, (size_t)(cap - n) * 2 * sizeof(int64_t));
  // _update_recent_pop_items / _update_recent_pop_norm: bincount of the nonzero ids, pop / (sum + 1) floored
  memset(recent_pop, 0, (size_t)num_items * sizeof(int64_t));
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t id = buffer[2 * i];
    if (id != 0) {
      if (id < 0 || id >= num_items) return NAR_ERR_INVALID;
      ++recent_pop[id]; ++total;
    }
  }
  const double denom = (double)(total + 1);
  for (int64_t v = 0; v < num_items; ++v) {
    const double x = (double)recent_pop[v] / denom;
    pop_norm[v] = x > min_norm_pop ? x : min_norm_pop;
  }
  // _update_pop_items
  for (int64_t i = 0; i < n_batch; ++i) {
    const int64_t id = batch_items[i];
    if (id < 0 || id >= num_items) return NAR_ERR_INVALID;
    ++articles_pop[id];
  }
  return NAR_OK;
}

// ItemsStateUpdaterHook.after_run, train-mode part (nar_model.py:1635-1646) + the update above in one call:
// [item_clicked | label_last_item] flattened row-major, padding (id 0) dropped, the last label carrying its session's
// maximum timestamp.  batch_scratch: [B*(T+1), 2] int64.
extern "C" int nar_host_state_update_batch(int64_t* buffer, int64_t cap, const int64_t* item_clicked,
                                           const int64_t* event_ts, const int64_t* label_last, int64_t B, int64_t T,
                                           int64_t hours_ms, int64_t* batch_scratch, int64_t* scratch, int64_t* recent_pop,
                                           double* pop_norm, int64_t* articles_pop, int64_t num_items, double min_norm_pop) {
  if (!item_clicked || !event_ts || !label_last || !batch_scratch || B < 0 || T < 0) return NAR_ERR_INVALID;
  int64_t* items = batch_scratch;
  int64_t* ts = batch_scratch + B * (T + 1);
  int64_t n = 0;
  for (int64_t b = 0; b < B; ++b) {
    int64_t mx = 0;
    for (int64_t t = 0; t < T; ++t) {
      const int64_t v = event_ts[b * T + t];
      mx = (t == 0 || v > mx) ? v : mx;
      if (item_clicked[b * T + t] != 0) { items[n] = item_clicked[b * T + t]; ts[n] = v; ++n; }
    }
    if (label_last[b] != 0) { items[n] = label_last[b]; ts[n] = mx; ++n; }
  }
  if (n == 0) return NAR_OK;                        // the hook skips empty batches
  return nar_host_state_update(buffer, cap, items, ts, n, hours_ms, scratch, recent_pop, pop_norm, articles_pop, num_items,
                               min_norm_pop);
}
