"""Host state of the recent-clicks buffer and recent popularity (numpy).

Mirror of the hot-path part of the reference class of the same name
(nar_module/nar/clicked_items_state.py:10-250): same constructor, same method
names, same arrays.  Out of scope: the co-occurrence
CSR matrix (:252-255, benchmarks only), cold-start bookkeeping (:97-123, :196-203).

Buffer semantics (clicked_items_state.py:206-228):
  * ``[max_size, 2]`` int64 rows (article_id, click_timestamp_ms), newest first;
  * on update the batch is reversed and prepended, rows older than
    ``min(batch ts) - hours`` are dropped first (that also drops zero padding rows),
    then the buffer is clipped / zero padded back to ``max_size``.
Popularity (clicked_items_state.py:231-246):
  * ``articles_recent_pop`` = bincount of nonzero buffer ids,
  * ``articles_recent_pop_norm`` = max(pop / (sum(pop)+1), 1/recent_clicks_for_normalization) (float64).
"""
from __future__ import annotations

import numpy as np


class ClickedItemsState:

    def __init__(self, recent_clicks_buffer_hours, recent_clicks_buffer_max_size,
                 recent_clicks_for_normalization, num_items):
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
        self.pop_recent_buffer_article_id_column = 0
        self.pop_recent_buffer_timestamp_column = 1
        self.current_step = 0

    # -- checkpoint around eval (clicked_items_state.py:49-79), hot-path fields only
    def save_state_checkpoint(self):
        self.articles_pop_chkp = np.copy(self.articles_pop)
        self.pop_recent_clicks_buffer_chkp = np.copy(self.pop_recent_clicks_buffer)
        self.current_step_chkp = self.current_step

    def restore_state_checkpoint(self):
        self.articles_pop = self.articles_pop_chkp
        del self.articles_pop_chkp
        self.pop_recent_clicks_buffer = self.pop_recent_clicks_buffer_chkp
        del self.pop_recent_clicks_buffer_chkp
        self.current_step = self.current_step_chkp
        # NB: like the reference, recent_pop / recent_pop_norm are NOT restored here;
        # they are recomputed by the next update_items_state().

    # -- getters (clicked_items_state.py:81-108)
    def get_articles_pop(self):
        return self.articles_pop

    def get_articles_recent_pop(self):
        return self.articles_recent_pop

    def get_articles_recent_pop_norm(self):
        return self.articles_recent_pop_norm

    def get_recent_clicks_buffer(self):
        return self.pop_recent_clicks_buffer[:, self.pop_recent_buffer_article_id_column]

    def increment_current_step(self):
        self.current_step += 1

    def get_current_step(self):
        return self.current_step

    def get_max_timestamp_recent_clicks(self):
        return np.max(self.pop_recent_clicks_buffer[:, self.pop_recent_buffer_timestamp_column])

    # -- update (clicked_items_state.py:187-250)
    def update_items_state(self, batch_clicked_items, batch_clicked_timestamps):
        self._update_recently_clicked_items_buffer(batch_clicked_items, batch_clicked_timestamps)
        self._update_recent_pop_items()
        self._update_pop_items(batch_clicked_items)

    def _update_recently_clicked_items_buffer(self, batch_clicked_items, batch_clicked_timestamps):
        batch = np.hstack([np.asarray(batch_clicked_items, dtype=np.int64).reshape(-1, 1),
                           np.asarray(batch_clicked_timestamps, dtype=np.int64).reshape(-1, 1)])
        batch = batch[::-1]                      # newest click first
        min_timestamp_batch = np.min(batch_clicked_timestamps)
        self.truncate_last_hours_recent_clicks_buffer(min_timestamp_batch)
        buf = np.vstack([batch, self.pop_recent_clicks_buffer])[:self.recent_clicks_buffer_max_size]
        if buf.shape[0] < self.recent_clicks_buffer_max_size:
            buf = np.vstack([buf, np.zeros(shape=[self.recent_clicks_buffer_max_size - buf.shape[0], 2],
                                           dtype=np.int64)])
        self.pop_recent_clicks_buffer = buf

    def truncate_last_hours_recent_clicks_buffer(self, reference_timestamp):
        MILISECS_BY_HOUR = 1000 * 60 * 60
        thr = reference_timestamp - int(self.recent_clicks_buffer_hours * MILISECS_BY_HOUR)
        ts = self.pop_recent_clicks_buffer[:, self.pop_recent_buffer_timestamp_column]
        self.pop_recent_clicks_buffer = self.pop_recent_clicks_buffer[ts >= thr]

    def _update_recent_pop_items(self):
        items = self.pop_recent_clicks_buffer[:, self.pop_recent_buffer_article_id_column]
        nz = items[np.nonzero(items)]
        self.articles_recent_pop = np.bincount(nz, minlength=self.num_items).astype(np.int64)
        self._update_recent_pop_norm(self.articles_recent_pop)

    def _update_recent_pop_norm(self, articles_recent_pop):
        min_norm_pop = 1.0 / self.recent_clicks_for_normalization
        self.articles_recent_pop_norm = np.maximum(articles_recent_pop / (articles_recent_pop.sum() + 1),
                                                   [min_norm_pop])

    def _update_pop_items(self, batch_items_nonzero):
        self.articles_pop += np.bincount(np.asarray(batch_items_nonzero, dtype=np.int64),
                                         minlength=self.num_items).astype(np.int64)


def batch_clicks_for_state_update(clicked_items, clicked_timestamps, last_item_label):
    """ItemsStateUpdaterHook.after_run, train-mode part (nar_model.py:1635-1646).

    clicked_items [B,T] i64, clicked_timestamps [B,T] i64, last_item_label [B,1] i64
    -> (items_nonzero, timestamps_nonzero) row-major flattened, padding dropped; the
    last label inherits the session's max timestamp.
    """
    batch_clicked_items = np.concatenate([clicked_items, last_item_label], axis=1)
    flat = batch_clicked_items.reshape(-1)
    nz = np.nonzero(flat)
    last_ts = np.max(clicked_timestamps, axis=1).reshape(-1, 1)
    ts = np.concatenate([clicked_timestamps, last_ts], axis=1).reshape(-1)
    return flat[nz], ts[nz]
