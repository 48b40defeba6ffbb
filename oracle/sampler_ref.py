"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the NAR negative sampler.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path never does.

Follows the reference TF graph
  nar_module/nar/nar_model.py:1220-1233  get_sample_from_recently_clicked_items_buffer
  nar_module/nar/nar_model.py:1281-1304  get_batch_negative_samples
  nar_module/nar/nar_model.py:1271-1279  get_negative_samples
  nar_module/nar/nar_model.py:1257-1268  get_neg_items_session   (ListDiff keeps order + dups)
  nar_module/nar/nar_model.py:1239-1254  get_neg_items_click     (shuffle, first-occurrence unique, first K, zero pad)
and its numpy twin nar_module/nar/benchmarks/candidate_sampling.py:13-90.

RNG.  TF's stateful tf.random_shuffle inside nested map_fn is not reproducible even by
TF, so "bit-exact negatives" is defined against THIS
specification (identical code runs in the CUDA kernel):

  Philox4x32-10, key = 64-bit seed, counter = (idx >> 2, ctx, stream, step); the 32-bit
  draw for element idx is output word (idx & 3).

  "shuffle x, keep the first n"            == give element idx the 64-bit key
  (rand32 << 32) | idx and keep the n smallest keys (ascending key = shuffled order).
  "shuffle, first-occurrence unique, first K" == an item's key is the min over its
  occurrences; output the K items with the smallest keys, ascending, zero padded.
  Both are exactly a uniform shuffle (ties are impossible: idx is unique).

  stream 1: buffer sample      idx = buffer position,            ctx = 0
  stream 2: candidate pool     idx = flat position in all_clicked_items [B,T+1] for batch
                               clicks, B*(T+1) + buffer position for buffer samples, ctx = 0
  stream 3: per click          idx = position in the pool (shuffled order), ctx = b*(T+1)+p
  (b = GLOBAL session index, so data-parallel ranks draw identical negatives).

parity: the reference holds no golden vectors for sampled indices; the
pin is (i) the reference's 8 property tests (ported in tests/test_sampler_oracle.py),
(ii) inclusion-frequency fixtures generated from the reference's own numpy sampler
(tests/golden/make_sampler_golden.py), (iii) inclusion frequencies of the reference's TF sampler
code itself (nar_model.py:1220-1304 imported unmodified, run on the TF-API stand-in
tests/golden/tf1_shim.py: tests/golden/make_sampler_tf_golden.py).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)
KEY_INVALID = np.uint64(0xFFFFFFFFFFFFFFFF)

STREAM_BUFFER = 1
STREAM_POOL = 2
STREAM_CLICK = 3
FIRST_SAMPLING_MULTIPLYING_FACTOR = 20      # nar_model.py:1282


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32); c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32); c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * PHILOX_M0
            p1 = c2.astype(np.uint64) * PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def rand32(seed: int, step: int, stream: int, ctx, idx):
    """32-bit draw for element ``idx`` (array) in context ``ctx`` (scalar or array)."""
    idx = np.asarray(idx, dtype=np.uint64)
    blk = (idx >> np.uint64(2)).astype(np.uint32)
    w = philox4x32_10(blk, np.asarray(ctx, dtype=np.uint32), np.uint32(stream), np.uint32(step & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (idx & np.uint64(3)).astype(np.int64)
    out = np.where(sel == 0, w[0], np.where(sel == 1, w[1], np.where(sel == 2, w[2], w[3])))
    return out.astype(np.uint32)


def keys64(seed, step, stream, ctx, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    return (rand32(seed, step, stream, ctx, idx).astype(np.uint64) << np.uint64(32)) | idx


def sample_buffer_positions(buffer: np.ndarray, sample_size: int, seed: int, step: int) -> np.ndarray:
    """nar_model.py:1220-1233 -> buffer POSITIONS selected (ascending key order)."""
    buffer = np.asarray(buffer, dtype=np.int64).ravel()
    pos = np.flatnonzero(buffer != 0)
    if pos.size == 0:
        return pos
    k = keys64(seed, step, STREAM_BUFFER, 0, pos)
    order = np.argsort(k, kind='stable')
    return pos[order[:sample_size]]


def build_pool(all_clicked_items: np.ndarray, buffer: np.ndarray, negative_samples: int,
               negative_sample_from_buffer: int, seed: int, step: int) -> np.ndarray:
    """nar_model.py:1286-1300 -> pool of candidate item ids in shuffled order (<= K*20, repeats kept)."""
    allf = np.asarray(all_clicked_items, dtype=np.int64).ravel()
    buffer = np.asarray(buffer, dtype=np.int64).ravel()
    nb = allf.size
    bpos = sample_buffer_positions(buffer, negative_sample_from_buffer, seed, step)
    j = np.flatnonzero(allf != 0)
    ids = np.concatenate([j.astype(np.uint64), (nb + bpos).astype(np.uint64)])
    items = np.concatenate([allf[j], buffer[bpos]])
    if ids.size == 0:
        return np.zeros(0, dtype=np.int64)
    k = keys64(seed, step, STREAM_POOL, 0, ids)
    order = np.argsort(k, kind='stable')[:negative_samples * FIRST_SAMPLING_MULTIPLYING_FACTOR]
    return items[order]


def neg_items_click(pool_valid_idx: np.ndarray, pool: np.ndarray, ctx: int, num_neg_samples: int,
                    seed: int, step: int) -> np.ndarray:
    """nar_model.py:1239-1254 for one click.  ``pool_valid_idx`` = pool positions that survived ListDiff."""
    out = np.zeros(num_neg_samples, dtype=np.int64)
    if pool_valid_idx.size == 0:
        return out
    k = keys64(seed, step, STREAM_CLICK, ctx, pool_valid_idx)
    items = pool[pool_valid_idx]
    order = np.argsort(k, kind='stable')
    items_sorted = items[order]
    _, first = np.unique(items_sorted, return_index=True)
    first_unique = items_sorted[np.sort(first)][:num_neg_samples]
    out[:first_unique.size] = first_unique
    return out


def sample_negatives(all_clicked_items: np.ndarray, buffer: np.ndarray, negative_samples: int,
                     negative_sample_from_buffer: int, seed: int, step: int,
                     session_offset: int = 0, all_clicked_items_global: np.ndarray = None) -> np.ndarray:
    """Whole sampler: all_clicked_items [B,T+1] i64, buffer [buf] i64 -> negatives [B,T,K] i64.

    Data parallel: ``all_clicked_items_global`` (every rank's sessions, [Bg,T+1]) builds the
    pool; ``all_clicked_items`` are this rank's rows, ``session_offset`` its first global row.
    """
    all_clicked_items = np.asarray(all_clicked_items, dtype=np.int64)
    glob = all_clicked_items if all_clicked_items_global is None else np.asarray(all_clicked_items_global, np.int64)
    B, T1 = all_clicked_items.shape
    K = int(negative_samples)
    pool = build_pool(glob, buffer, K, negative_sample_from_buffer, seed, step)
    neg = np.zeros((B, T1 - 1, K), dtype=np.int64)
    for b in range(B):
        sess = all_clicked_items[b]
        # tf.setdiff1d (ListDiff): keep order and duplicates of pool entries not in the session
        valid_idx = np.flatnonzero(~np.isin(pool, sess))
        for p in range(T1 - 1):        # the last position (label_last_item) is dropped (nar_model.py:275)
            if sess[p] == 0:
                continue
            ctx = (session_offset + b) * T1 + p
            neg[b, p] = neg_items_click(valid_idx, pool, ctx, K, seed, step)
    return neg


# ---------------------------------------------------------------------------
# Reference-API-shaped wrapper so the reference's own unit tests can run against the spec
# (candidate_sampling.py:7-90 method names / signatures).
# ---------------------------------------------------------------------------
class CandidateSamplingManager:
    def __init__(self, get_recent_clicks_buffer_fn, ignore_session_items_on_sampling=True, seed=42):
        self.get_recent_clicks_buffer_fn = get_recent_clicks_buffer_fn
        self.ignore_session_items_on_sampling = ignore_session_items_on_sampling
        self.seed = seed
        self.step = 0

    def _tick(self):
        self.step += 1
        return self.step

    def get_sample_from_recently_clicked_items_buffer(self, sample_size):
        buf = np.asarray(self.get_recent_clicks_buffer_fn(), dtype=np.int64).ravel()
        return buf[sample_buffer_positions(buf, sample_size, self.seed, self._tick())]

    def get_neg_items_click(self, valid_samples_session, num_neg_samples, ctx=0, step=None):
        pool = np.asarray(valid_samples_session, dtype=np.int64)
        return neg_items_click(np.arange(pool.size), pool, ctx, num_neg_samples, self.seed,
                               self._tick() if step is None else step)

    def get_neg_items_session(self, session_item_ids, candidate_samples, num_neg_samples, b=0, step=None):
        session_item_ids = np.asarray(session_item_ids, dtype=np.int64)
        pool = np.asarray(candidate_samples, dtype=np.int64)
        step = self._tick() if step is None else step
        if self.ignore_session_items_on_sampling:
            valid_idx = np.flatnonzero(~np.isin(pool, session_item_ids))
        else:
            valid_idx = np.arange(pool.size)
        T1 = session_item_ids.size
        return np.vstack([neg_items_click(valid_idx, pool, b * T1 + p, num_neg_samples, self.seed, step)
                          if click_id != 0 else np.zeros(num_neg_samples, np.int64)
                          for p, click_id in enumerate(session_item_ids)])

    def get_negative_samples(self, all_clicked_items, candidate_samples, num_neg_samples):
        step = self._tick()
        return np.vstack([np.expand_dims(self.get_neg_items_session(s, candidate_samples, num_neg_samples, b, step), 0)
                          for b, s in enumerate(np.asarray(all_clicked_items))])

    def get_batch_negative_samples_by_session(self, all_clicked_items, additional_samples, num_negative_samples,
                                              first_sampling_multiplying_factor=20):
        allf = np.asarray(all_clicked_items, dtype=np.int64)
        step = self._tick()
        j = np.flatnonzero(allf.ravel() != 0)
        add = np.asarray(additional_samples, dtype=np.int64)
        ids = np.concatenate([j, allf.size + np.arange(add.size)]).astype(np.uint64)
        items = np.concatenate([allf.ravel()[j], add])
        k = keys64(self.seed, step, STREAM_POOL, 0, ids)
        pool = items[np.argsort(k, kind='stable')[:num_negative_samples * first_sampling_multiplying_factor]]
        return np.vstack([np.expand_dims(self.get_neg_items_session(s, pool, num_negative_samples, b, step), 0)
                          for b, s in enumerate(allf)])

    def get_batch_negative_samples(self, all_clicked_items, negative_samples_by_session, negative_sample_from_buffer):
        buf = np.asarray(self.get_recent_clicks_buffer_fn(), dtype=np.int64).ravel()
        # the reference twin keeps every position (no label column is appended there): add a dummy
        # last column so that sample_negatives' "drop the last position" leaves the same shape
        allf = np.asarray(all_clicked_items, dtype=np.int64)
        ext = np.concatenate([allf, np.zeros((allf.shape[0], 1), np.int64)], axis=1)
        return sample_negatives(ext, buf, negative_samples_by_session, negative_sample_from_buffer,
                                self.seed, self._tick())
