"""Data-parallel host logic (numpy only; no CUDA needed, so it is testable with gloo on CPU).

The reference has no distribution at all (README.md:252: single worker on purpose).  The B200 build
shards the *sessions* of one global batch contiguously across ranks (chronological order per rank is
kept) and keeps global-batch semantics identical to one GPU:
  * every rank sees the ids of the whole global batch, so the candidate pool and the per-click draws
    (counter = global session index) are identical for any world size;
  * the loss normaliser sum(mask) is the GLOBAL count (known from session_size, no collective);
  * gradients are sum-allreduced (NCCL) - the only collective on the path;
  * the host ClickedItemsState update is a deterministic function of the global ids: every rank
    computes it redundantly.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def shard_sessions(session_size: np.ndarray, T: int, world: int, rank: int) -> Dict[str, np.ndarray]:
    """-> dict(s0, per, lens[per], L, L_global, sess_off[per+1] int32, pos_idx[L] int32 (flat b*T+t, global b))."""
    Bg = int(session_size.shape[0])
    lens_g = np.clip(np.asarray(session_size, dtype=np.int64) - 1, 0, T)      # seq_lengths, nar_model.py:227
    per = Bg // world
    if per * world != Bg:
        raise ValueError('global batch %d not divisible by world size %d' % (Bg, world))
    s0 = rank * per
    lens = lens_g[s0:s0 + per]
    sess_off = np.zeros(per + 1, dtype=np.int32)
    np.cumsum(lens, out=sess_off[1:])
    tt = np.arange(T, dtype=np.int64)[None, :]
    valid = tt < lens[:, None]                                                # tf.sequence_mask, nar_model.py:231
    bb = (np.arange(per, dtype=np.int64) + s0)[:, None]
    pos_idx = (bb * T + tt)[valid].astype(np.int32)
    return {'s0': s0, 'per': per, 'lens': lens, 'L': int(lens.sum()), 'L_global': int(lens_g.sum()),
            'sess_off': sess_off, 'pos_idx': pos_idx}
