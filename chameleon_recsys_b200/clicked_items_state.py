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
        """One call per step.  Runs the single-pass C implementation in libnar_b200 (``nar_host_state_update``, host
        code) when the library is built - the numpy restatement below is the specification it is tested against
        (tests/test_host_state.py) and costs ~0.8 ms per G1 step, which made the end-to-end loop host bound."""
        if self._native_update(batch_clicked_items, batch_clicked_timestamps):
            return
        self.update_items_state_numpy(batch_clicked_items, batch_clicked_timestamps)

    def update_items_state_numpy(self, batch_clicked_items, batch_clicked_timestamps):
        self._update_recently_clicked_items_buffer(batch_clicked_items, batch_clicked_timestamps)
        self._update_recent_pop_items()
        self._update_pop_items(batch_clicked_items)

    _lib = None           # class-level cache: ctypes handle, or False when the library is not available

    @classmethod
    def _native(cls):
        if cls._lib is None:
            try:
                from . import _lib as nl
                cls._lib = nl.load()
            except Exception:  # noqa: BLE001  (library not built: the numpy path is complete)
                cls._lib = False
        return cls._lib

    def update_from_batch(self, clicked_items, clicked_timestamps, last_item_label):
        """ItemsStateUpdaterHook.after_run in one call: ``batch_clicks_for_state_update`` + ``update_items_state``
        (both stay as the numpy specification; the C path does the same in one pass over the padded batch)."""
        lib = self._native()
        ci = np.ascontiguousarray(clicked_items, dtype=np.int64)
        if lib is False or ci.ndim != 2:
            items, ts = batch_clicks_for_state_update(clicked_items, clicked_timestamps, last_item_label)
            if items.size:
                self.update_items_state(items, ts)
            return
        ct = np.ascontiguousarray(clicked_timestamps, dtype=np.int64)
        ll = np.ascontiguousarray(last_item_label, dtype=np.int64).reshape(-1)
        B, T = ci.shape
        bs = getattr(self, '_batch_scratch', None)
        if bs is None or bs.size < 2 * B * (T + 1):
            bs = self._batch_scratch = np.empty(2 * B * (T + 1), dtype=np.int64)
        ok = self._native_call(lambda buf, scratch, recent, norm, pop, hours_ms: lib.nar_host_state_update_batch(
            buf.ctypes.data, buf.shape[0], ci.ctypes.data, ct.ctypes.data, ll.ctypes.data, B, T, hours_ms, bs.ctypes.data,
            scratch.ctypes.data, recent.ctypes.data, norm.ctypes.data, pop.ctypes.data, self.num_items,
            1.0 / self.recent_clicks_for_normalization), keep_pop_on_empty=not (ci.any() or ll.any()))
        if not ok:                            # unusual array layout: the numpy specification handles everything
            items, ts = batch_clicks_for_state_update(clicked_items, clicked_timestamps, last_item_label)
            if items.size:
                self.update_items_state_numpy(items, ts)

    def _native_update(self, batch_clicked_items, batch_clicked_timestamps) -> bool:
        lib = self._native()
        if lib is False:
            return False
        items = np.ascontiguousarray(batch_clicked_items, dtype=np.int64).reshape(-1)
        ts = np.ascontiguousarray(batch_clicked_timestamps, dtype=np.int64).reshape(-1)
        if items.size == 0 or items.size != ts.size:
            return False
        return self._native_call(lambda buf, scratch, recent, norm, pop, hours_ms: lib.nar_host_state_update(
            buf.ctypes.data, buf.shape[0], items.ctypes.data, ts.ctypes.data, items.size, hours_ms, scratch.ctypes.data,
            recent.ctypes.data, norm.ctypes.data, pop.ctypes.data, self.num_items,
            1.0 / self.recent_clicks_for_normalization))

    def _native_call(self, fn, keep_pop_on_empty: bool = False) -> bool:
        if keep_pop_on_empty:
            return True                       # nothing but padding in the batch: the hook does not touch the state
        buf = self.pop_recent_clicks_buffer
        if buf.dtype != np.int64 or not buf.flags['C_CONTIGUOUS'] or not buf.flags['WRITEABLE'] or \
                buf.shape != (self.recent_clicks_buffer_max_size, 2):
            buf = np.ascontiguousarray(buf, dtype=np.int64).copy()
            if buf.shape != (self.recent_clicks_buffer_max_size, 2):
                return False
        scratch = getattr(self, '_scratch', None)
        if scratch is None or scratch.shape != buf.shape:
            scratch = self._scratch = np.empty_like(buf)
        # two alternating output sets: fresh 368 KB arrays per step cost more (page faults) than the update itself, and
        # whoever still holds the previous step's arrays (a feed dict) keeps seeing that step's values
        flip = self._flip = 1 - getattr(self, '_flip', 0)
        outs = getattr(self, '_outs', None)
        if outs is None or outs[0][0].size != self.num_items:
            outs = self._outs = [(np.empty(self.num_items, dtype=np.int64), np.empty(self.num_items, dtype=np.float64))
                                 for _ in range(2)]
        recent, norm = outs[flip]
        pop = self.articles_pop
        if pop.dtype != np.int64 or not pop.flags['C_CONTIGUOUS'] or not pop.flags['WRITEABLE']:
            pop = np.ascontiguousarray(pop, dtype=np.int64).copy()
        hours_ms = int(self.recent_clicks_buffer_hours * 1000 * 60 * 60)
        rc = fn(buf, scratch, recent, norm, pop, hours_ms)
        if rc != 0:
            raise ValueError('nar_host_state_update failed (%d): article id outside [0, num_items)?' % rc)
        self.pop_recent_clicks_buffer = buf
        self.articles_recent_pop = recent
        self.articles_recent_pop_norm = norm
        self.articles_pop = pop
        return True

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
