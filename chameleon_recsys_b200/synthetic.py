"""Synthetic G1/Adressa-shaped inputs.  numpy only, seed 42.

* ACR table [V,E] f32 ~ N(0,1), rows l2-normalised x content_embedding_scale_factor
  (mirrors nar_trainer_gcom.py:470-474); row 0 = padding article.
* metadata: created_at_ts[V] i64 uniform over the 16 days before the first click,
  category_id[V] uniform < cardinality.
* sessions: session_size = min(2 + Geom(p=.53), S) ('g1', mean ~2.9 clicks) or S ('dense');
  item ids Zipf(1.1) over a sliding "alive" window of 2000 ids, no repeats inside a
  session; event_timestamp i64 ms increasing with 30 s mean gaps; batches advance ~10 s;
  context ids uniform within the G1 cardinalities; time floats in [-1,1].
The stream is emitted in the on-disk SequenceExample schema's terms
(nar_preprocess_gcom.py:75-108): one dict per session holding the *full* click list;
``datasets.py`` applies the reference's truncate / shift-by-one / padding.
"""
from __future__ import annotations

from typing import Dict, Iterator, List

import numpy as np

MS_DAY = 86400000
T0_MS = 1506826800000      # 2017-10-01 03:00 UTC, first G1 click hour


def make_catalog(num_items: int, acr_dim: int, articles_features_config: dict,
                 content_embedding_scale_factor: float = 1.0, seed: int = 42):
    """-> (content_article_embeddings_matrix [V,E] f32, articles_metadata dict of [V] arrays)."""
    rs = np.random.RandomState(seed)
    acr = rs.standard_normal((num_items, acr_dim)).astype(np.float32)
    # sklearn Normalizer('l2') per row, then scale (nar_trainer_gcom.py:470-474)
    norms = np.sqrt((acr.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    norms[norms == 0.0] = 1.0
    acr = (acr / norms).astype(np.float32) * np.float32(content_embedding_scale_factor)
    meta = {}
    for fname, fc in articles_features_config.items():
        if fname == 'article_id':
            meta[fname] = np.arange(num_items, dtype=np.int64)
        elif fname == 'created_at_ts':
            meta[fname] = (T0_MS - rs.randint(0, 16 * MS_DAY, size=num_items)).astype(np.int64)
        elif fc['type'] == 'categorical':
            meta[fname] = rs.randint(0, fc['cardinality'], size=num_items).astype(np.int64)
        else:
            meta[fname] = rs.uniform(-1, 1, size=num_items).astype(np.float32)
    return acr.astype(np.float32), meta


class SessionStream:
    """Endless chronological stream of synthetic sessions."""

    def __init__(self, num_items: int, session_features_config: dict, max_session_len: int,
                 length_dist: str = 'g1', seed: int = 42, alive_window: int = 2000,
                 sessions_per_tick: int = 256):
        self.V = int(num_items)
        self.cfg = session_features_config
        self.S = int(max_session_len)
        self.length_dist = length_dist
        self.rs = np.random.RandomState(seed)
        self.alive = min(alive_window, self.V - 1)
        self.now = T0_MS
        self.sessions_per_tick = sessions_per_tick
        self.n_emitted = 0
        # Zipf(1.1) weights over the alive window ranks
        ranks = np.arange(1, self.alive + 1, dtype=np.float64)
        w = ranks ** -1.1
        self.cdf = np.cumsum(w / w.sum())
        self.window_start = 1

    def _draw_items(self, n: int) -> np.ndarray:
        out: List[int] = []
        seen = set()
        while len(out) < n:
            u = self.rs.random_sample(2 * n)
            r = np.searchsorted(self.cdf, u)
            for x in r:
                item = 1 + (self.window_start - 1 + int(x)) % (self.V - 1)
                if item not in seen:
                    seen.add(item); out.append(item)
                    if len(out) == n:
                        break
        return np.asarray(out, dtype=np.int64)

    def next_session(self) -> Dict[str, np.ndarray]:
        if self.length_dist == 'dense':
            size = self.S
        else:
            size = int(min(2 + self.rs.geometric(0.53) - 1, self.S))   # numpy geometric is >=1
        # long sessions exist on disk; the reader truncates. Emit up to S+2 to exercise that.
        items = self._draw_items(size)
        gaps = self.rs.exponential(30000.0, size=size).astype(np.int64) + 1
        start = self.now + int(self.rs.randint(0, 10000))
        ts = start + np.cumsum(gaps) - gaps[0]
        s = {'user_id': np.int64(self.rs.randint(1, 341193)),
             'session_id': np.int64(start * 1000 + self.n_emitted % 1000),
             'session_start': np.int64(start),
             'session_size': np.int64(size),
             'item_clicked': items,
             'event_timestamp': ts.astype(np.int64)}
        for fname, fc in self.cfg['sequence_features'].items():
            if fname in ('item_clicked', 'event_timestamp'):
                continue
            if fc['type'] == 'categorical':
                # one device/location per session; ids in [1, card)
                v = self.rs.randint(1, fc['cardinality'])
                s[fname] = np.full(size, v, dtype=np.int64)
            else:
                s[fname] = self.rs.uniform(-1, 1, size=size).astype(np.float32)
        self.n_emitted += 1
        if self.n_emitted % self.sessions_per_tick == 0:
            self.now += 10000                       # batches advance ~10 s
            self.window_start = 1 + (self.window_start + 3) % (self.V - 1)   # news recency drift
        return s

    def __iter__(self) -> Iterator[Dict[str, np.ndarray]]:
        while True:
            yield self.next_session()
