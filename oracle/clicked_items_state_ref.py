"""TEST INFRASTRUCTURE - numpy restatement of the reference's ``ClickedItemsState`` update
(nar_module/nar/clicked_items_state.py:187-250) and of the hook's batch flattening (nar_model.py:1635-1646).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product class (chameleon_recsys_b200/clicked_items_state.py) runs the single-pass C implementation in
libnar_b200 and has no numpy fallback; this file is the specification that C pass is bit-checked against
(tests/test_host_state.py), itself pinned to fixtures produced by the reference class
(tests/golden/make_state_golden.py).

Buffer semantics (clicked_items_state.py:206-228): ``[max_size, 2]`` int64 rows (article_id, click_timestamp_ms),
newest first; on update the batch is reversed and prepended, rows older than ``min(batch ts) - hours`` are dropped
first (that also drops zero padding rows), then the buffer is clipped / zero padded back to ``max_size``.
Popularity (:231-246): ``articles_recent_pop`` = bincount of nonzero buffer ids, ``articles_recent_pop_norm`` =
max(pop / (sum(pop)+1), 1/recent_clicks_for_normalization) in float64.
"""
from __future__ import annotations

import numpy as np


def batch_clicks_for_state_update(clicked_items, clicked_timestamps, last_item_label):
    """ItemsStateUpdaterHook.after_run, train-mode part (nar_model.py:1635-1646): [B,T] ids / timestamps + [B,1] last
    label -> (items_nonzero, timestamps_nonzero) row-major, padding dropped; the last label inherits the session's
    max timestamp."""
    batch_clicked_items = np.concatenate([clicked_items, last_item_label], axis=1)
    flat = batch_clicked_items.reshape(-1)
    nz = np.nonzero(flat)
    last_ts = np.max(clicked_timestamps, axis=1).reshape(-1, 1)
    ts = np.concatenate([clicked_timestamps, last_ts], axis=1).reshape(-1)
    return flat[nz], ts[nz]


class ClickedItemsStateRef:
    """Same constructor / getters / arrays as the reference class; numpy only."""

    def __init__(self, recent_clicks_buffer_hours, recent_clicks_buffer_max_size, recent_clicks_for_normalization, num_items):
        self.recent_clicks_buffer_hours = recent_clicks_buffer_hours
        self.recent_clicks_buffer_max_size = recent_clicks_buffer_max_size
        self.recent_clicks_for_normalization = recent_clicks_for_normalization
        self.num_items = num_items
        self.reset_state()

    def reset_state(self):
        self.articles_pop = np.zeros(shape=[self.num_items], dtype=np.int64)
        self.articles_recent_pop = np.zeros(shape=[self.num_items], dtype=np.int64)
        self._update_recent_pop_norm(self.articles_recent_pop)
        self.pop_recent_clicks_buffer = np.zeros(shape=[self.recent_clicks_buffer_max_size, 2], dtype=np.int64)
        self.current_step = 0

    def get_articles_pop(self):
        return self.articles_pop

    def get_articles_recent_pop(self):
        return self.articles_recent_pop

    def get_articles_recent_pop_norm(self):
        return self.articles_recent_pop_norm

    def get_recent_clicks_buffer(self):
        return self.pop_recent_clicks_buffer[:, 0]

    # clicked_items_state.py:187-194
    def update_items_state(self, batch_clicked_items, batch_clicked_timestamps):
        self._update_recently_clicked_items_buffer(batch_clicked_items, batch_clicked_timestamps)
        self._update_recent_pop_items()
        self._update_pop_items(batch_clicked_items)

    def update_from_batch(self, clicked_items, clicked_timestamps, last_item_label):
        items, ts = batch_clicks_for_state_update(clicked_items, clicked_timestamps, last_item_label)
        if items.size:
            self.update_items_state(items, ts)

    # :206-228
    def _update_recently_clicked_items_buffer(self, batch_clicked_items, batch_clicked_timestamps):
        batch = np.hstack([np.asarray(batch_clicked_items, dtype=np.int64).reshape(-1, 1),
                           np.asarray(batch_clicked_timestamps, dtype=np.int64).reshape(-1, 1)])
        batch = batch[::-1]                      # newest click first
        thr = np.min(batch_clicked_timestamps) - int(self.recent_clicks_buffer_hours * 1000 * 60 * 60)
        kept = self.pop_recent_clicks_buffer[self.pop_recent_clicks_buffer[:, 1] >= thr]
        buf = np.vstack([batch, kept])[:self.recent_clicks_buffer_max_size]
        if buf.shape[0] < self.recent_clicks_buffer_max_size:
            buf = np.vstack([buf, np.zeros(shape=[self.recent_clicks_buffer_max_size - buf.shape[0], 2], dtype=np.int64)])
        self.pop_recent_clicks_buffer = buf

    # :231-246
    def _update_recent_pop_items(self):
        items = self.pop_recent_clicks_buffer[:, 0]
        self.articles_recent_pop = np.bincount(items[np.nonzero(items)], minlength=self.num_items).astype(np.int64)
        self._update_recent_pop_norm(self.articles_recent_pop)

    def _update_recent_pop_norm(self, articles_recent_pop):
        self.articles_recent_pop_norm = np.maximum(articles_recent_pop / (articles_recent_pop.sum() + 1),
                                                   [1.0 / self.recent_clicks_for_normalization])

    def _update_pop_items(self, batch_items_nonzero):
        self.articles_pop += np.bincount(np.asarray(batch_items_nonzero, dtype=np.int64),
                                         minlength=self.num_items).astype(np.int64)
