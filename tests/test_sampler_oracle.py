"""Oracle sampler (oracle/sampler_ref.py) pinned against
(1) Philox4x32-10 known-answer vectors (Random123 kat_vectors),
(2) the reference's own 8 unit tests, ported from
    nar_module/nar/benchmarks/candidate_sampling_tests.py:17-100 (same inputs, same assertions),
(3) inclusion-frequency fixtures generated from the reference's numpy sampler
    (tests/golden/make_sampler_golden.py -> tests/golden/sampler_freq.json)."""
import json
import os

import numpy as np
import pytest

from oracle import sampler_ref
from oracle.sampler_ref import CandidateSamplingManager

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = sampler_ref.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in out) == exp


@pytest.fixture
def mgr():
    buf = np.array([1, 2, 3, 1, 2, 3, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 9, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    return CandidateSamplingManager(lambda: buf)


def test_get_sample_from_recently_clicked_items_buffer(mgr):
    sample = mgr.get_sample_from_recently_clicked_items_buffer(5)
    assert sample.shape == (5,)
    assert 0 not in sample


def test_get_neg_items_click(mgr):
    sample = mgr.get_neg_items_click([1, 2, 2, 4, 4, 5, 4, 3, 2, 16, 4, 8, 6], num_neg_samples=5)
    assert sample.shape == (5,)
    assert np.unique(sample).shape == (5,)


def test_get_neg_items_click_padding(mgr):
    sample = mgr.get_neg_items_click([1, 2, 2], num_neg_samples=10)
    assert sample.shape == (10,)
    assert np.count_nonzero(sample) == 2
    assert not sample[2:].any()


def test_get_neg_items_session(mgr):
    session_item_ids = [1, 2, 3]
    candidate_samples = [1, 3, 5, 7, 9, 11, 13, 15, 18, 20, 9, 11]
    samples = mgr.get_neg_items_session(session_item_ids, candidate_samples, 10)
    assert samples.shape == (3, 10)
    assert np.count_nonzero(samples == 0) == 2 * 3
    assert not samples[:, -2:].any()
    for i in session_item_ids:
        assert i not in samples


def test_get_neg_items_session_not_ignore_session_items():
    buf = np.array([1, 2, 3, 0])
    m = CandidateSamplingManager(lambda: buf, ignore_session_items_on_sampling=False)
    session_item_ids = [1, 2, 3]
    samples = m.get_neg_items_session(session_item_ids, [1, 3, 5, 7, 9, 2, 13, 15, 18, 20, 9], 10)
    assert samples.shape == (3, 10)
    for i in session_item_ids:
        assert i in samples


def test_get_negative_samples(mgr):
    sessions = np.array([[1, 2, 3], [4, 0, 0]])
    samples = mgr.get_negative_samples(sessions, [1, 3, 5, 7, 9, 11, 13, 15, 18, 20, 9, 11], 10)
    assert samples.shape == (2, 3, 10)
    assert np.count_nonzero(samples == 0) == 2 * 10 + 2 * 3
    assert not samples[1, -2:].any()
    for session, neg in zip(sessions, samples):
        assert len(set(session.ravel()).intersection(set(neg.ravel())).difference({0})) == 0


def test_get_batch_negative_samples_by_session(mgr):
    sessions = np.array([[1, 2, 3, 4, 5], [4, 5, 6, 7, 0]])
    cand = [2, 2, 3, 3, 4, 4, 4, 5, 5, 5, 5, 6, 7, 10, 10, 10, 11, 11, 12, 12, 13, 15]
    samples = mgr.get_batch_negative_samples_by_session(sessions, cand, num_negative_samples=3,
                                                        first_sampling_multiplying_factor=2)
    assert samples.shape == (2, 5, 3)
    assert not samples[1, -1].any()
    for session, neg in zip(sessions, samples):
        assert len(set(session.ravel()).intersection(set(neg.ravel())).difference({0})) == 0


def test_get_batch_negative_samples(mgr):
    sessions = np.array([[1, 2, 3, 4, 5], [4, 5, 6, 7, 0]])
    samples = mgr.get_batch_negative_samples(sessions, 4, 10)
    assert samples.shape == (2, 5, 4)
    assert not samples[1, -1].any()
    for session, neg in zip(sessions, samples):
        assert len(set(session.ravel()).intersection(set(neg.ravel())).difference({0})) == 0


def test_listdiff_keeps_duplicates():
    """tf.setdiff1d (ListDiff) keeps order and duplicates (nar_model.py:1259); np.setdiff1d(assume_unique)
    of the numpy twin does not always (SURVEY.md section 4 caveat)."""
    pool = np.array([7, 7, 8, 9, 9, 9, 3])
    sess = np.array([3, 0, 0])
    valid = np.flatnonzero(~np.isin(pool, sess))
    assert list(pool[valid]) == [7, 7, 8, 9, 9, 9]


def test_determinism_and_step_dependence():
    rs = np.random.RandomState(1)
    allc = rs.randint(1, 100, size=(6, 5)); allc[2, 2:4] = 0
    buf = rs.randint(0, 100, size=200)
    a = sampler_ref.sample_negatives(allc, buf, 7, 50, 42, 3)
    b = sampler_ref.sample_negatives(allc, buf, 7, 50, 42, 3)
    c = sampler_ref.sample_negatives(allc, buf, 7, 50, 42, 4)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # data-parallel slice == rows of the global result
    d = sampler_ref.sample_negatives(allc[3:], buf, 7, 50, 42, 3, session_offset=3, all_clicked_items_global=allc)
    assert np.array_equal(d, a[3:])
    # per click: unique, no session items, padded positions empty
    for bi in range(6):
        for p in range(4):
            row = a[bi, p]
            nz = row[row != 0]
            assert len(set(nz)) == len(nz)
            assert not set(nz) & set(allc[bi])
            if allc[bi, p] == 0:
                assert not row.any()


def test_inclusion_frequencies_match_reference_sampler():
    """Popularity-proportional sampling without replacement: the spec's per-item inclusion frequencies
    must agree with the reference numpy sampler's (candidate_sampling.py) within sampling noise."""
    with open(os.path.join(HERE, 'golden', 'sampler_freq.json')) as f:
        g = json.load(f)
    pool = np.array(g['pool'], dtype=np.int64)
    K, trials = g['K'], 4000
    items = np.array(g['items'])
    counts = np.zeros(len(items))
    m = CandidateSamplingManager(lambda: np.zeros(1, np.int64))
    for t in range(trials):
        s = m.get_neg_items_click(pool, K, ctx=t, step=1 + t // 1000)
        counts += np.isin(items, s)
    freq = counts / trials
    ref = np.array(g['freq'])
    se = np.sqrt(ref * (1 - ref) / trials + ref * (1 - ref) / g['trials']) + 1e-9
    z = np.abs(freq - ref) / se
    assert z.max() < 4.5, (z.max(), freq, ref)


def test_inclusion_frequencies_match_reference_tf_sampler():
    """Distribution of the whole training-graph sampler (buffer sample -> candidate pool with repetitions, truncated to
    K*20 -> per-session exclusion -> per-click shuffle / unique / first K) against the REFERENCE's TF sampler code run
    3000 times on the TF-API stand-in (tests/golden/make_sampler_tf_golden.py): same support, and per (click, item)
    inclusion frequencies within sampling noise (two independent 3000-trial estimates: z = diff / sqrt(2 p (1-p) / N))."""
    d = np.load(os.path.join(HERE, 'golden', 'sampler_tf_freq.npz'))
    allc, buf, V = d['all_clicked'], d['buffer'], int(d['V'])
    B, T1 = allc.shape
    for ci in range(2):
        K, nfb, n_ref = (int(v) for v in d['c%d_cfg' % ci])
        N = 3000
        counts = np.zeros((B, T1 - 1, V))
        for step in range(1, N + 1):
            neg = sampler_ref.sample_negatives(allc, buf, K, nfb, 1234, step)
            for b in range(B):
                for t in range(T1 - 1):
                    row = neg[b, t]
                    counts[b, t, row[row != 0]] += 1
        f, g = counts / N, d['c%d_freq' % ci]
        assert np.array_equal(f > 0, g > 0)                      # same candidates are reachable for every click
        p = (f + g) / 2
        z = np.abs(f - g)[p > 0] / np.sqrt(p * (1 - p) * (1.0 / N + 1.0 / n_ref))[p > 0]
        assert z.max() < 4.5 and (z ** 2).mean() < 1.6, (ci, z.max(), (z ** 2).mean())
        assert float(d['c%d_pad_per_trial' % ci]) == 0.0          # (the scenario never runs out of candidates)
