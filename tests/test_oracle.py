"""Pins for the float oracle (oracle/nar_oracle.py).  The reference has no test or golden vector for
logits / loss / gradients / Adam; besides the comparison with the reference's own graph code run on a
TF-API stand-in (tests/test_oracle_reference_model.py) the oracle is pinned here by
(1) an independent scalar restatement of Appendix A on a hand-sized case with integer-valued weights,
(2) closed-form known answers (UGRNN cell, TF-Adam first step, l2 regulariser),
(3) finite-difference gradients in float64, (4) invariants of the reference code."""
import math

import numpy as np
import pytest
import torch

from chameleon_recsys_b200.harness import make_problem, warm_state
from chameleon_recsys_b200.hparams import (NARHParams, get_articles_features_config,
                                           get_internal_enabled_features_config, get_session_features_config)
from chameleon_recsys_b200.plan import FeaturePlan, ParamLayout
from oracle import sampler_ref
from oracle.nar_oracle import NarOracle


def _oracle(pb, dtype=torch.float64, **kw):
    hp = pb.hp
    args = dict(negative_samples=hp.train_total_negative_samples, softmax_temperature=hp.softmax_temperature,
                reg_weight_decay=hp.reg_l2, recent_clicks_for_normalization=hp.recent_clicks_for_normalization,
                CAR_embedding_size=hp.CAR_embedding_size, rnn_units=hp.rnn_units, rnn_num_layers=hp.rnn_num_layers,
                lr=hp.learning_rate, dtype=dtype)
    args.update(kw)
    return NarOracle(pb.session_features_config, pb.articles_features_config, pb.internal_features_config,
                     pb.content_article_embeddings_matrix, pb.articles_metadata, **args)


def _batch(pb, warm=3, step=1):
    if warm:
        warm_state(pb, warm)
    f, l = pb.input_fn().get_next()
    allc = np.concatenate([f['item_clicked'], l['label_last_item']], axis=1)
    buf = pb.clicked_items_state.get_recent_clicks_buffer().copy()
    pop = pb.clicked_items_state.get_articles_recent_pop_norm().copy()
    neg = sampler_ref.sample_negatives(allc, buf, pb.hp.train_total_negative_samples,
                                       pb.hp.train_negative_samples_from_buffer, 42, step)
    return f, l, neg, buf, pop


def leaky(x):
    return x if x > 0 else 0.2 * x


def test_hand_sized_case_matches_scalar_restatement():
    """1 session, 2 input clicks, K=2, ACR only (E=2), C=2, H=1: every number recomputed with plain python floats
    following SURVEY.md Appendix A."""
    V, E, C, H, K = 6, 2, 2, 1, 2
    acfg = get_articles_features_config(V, ['NONE'])
    scfg = get_session_features_config(V, ['NONE'])
    icfg = get_internal_enabled_features_config(['article_content_embeddings'])
    acr = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [2, -1], [-1, 2]], dtype=np.float32)
    meta = {'article_id': np.arange(V), 'created_at_ts': np.zeros(V, np.int64)}
    p = 'main/'
    shapes = {'user_items_contextual_features/input_features_center_scale/gamma_scale': (3,),
              'user_items_contextual_features/input_features_center_scale/beta_center': (3,),
              'CAR/PreCAR_representation/kernel': (3, C), 'CAR/PreCAR_representation/bias': (C,),
              'CAR/CAR_representation/kernel': (C, C), 'CAR/CAR_representation/bias': (C,),
              'RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/kernel': (C + H, 2 * H),
              'RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/bias': (2 * H,),
              'session_representation/FC1/kernel': (H, 512), 'session_representation/FC1/bias': (512,),
              'session_representation/FC2/kernel': (512, C), 'session_representation/FC2/bias': (C,)}
    for i, (a, b) in enumerate([(C, 128), (128, 64), (64, 32), (32, 1)]):
        shapes['recommendations_ranking/matching_dense_layer_%d/kernel' % (i + 1)] = (a, b)
        shapes['recommendations_ranking/matching_dense_layer_%d/bias' % (i + 1)] = (b,)
    lg = {p + k: np.zeros(v, np.float32) for k, v in shapes.items()}
    lg[p + 'user_items_contextual_features/input_features_center_scale/gamma_scale'][:] = [1, 2, 1]     # F = 1 (dummy ctx) + 2
    lg[p + 'user_items_contextual_features/input_features_center_scale/beta_center'][:] = [0, 0, 1]
    W1 = np.array([[5, 5], [1, -1], [2, 1]], np.float32); lg[p + 'CAR/PreCAR_representation/kernel'][:] = W1
    lg[p + 'CAR/PreCAR_representation/bias'][:] = [0, 1]
    W2 = np.array([[1, 0], [1, 1]], np.float32) * 0.5; lg[p + 'CAR/CAR_representation/kernel'][:] = W2
    Wr = np.array([[1, -1], [0, 1], [2, 1]], np.float32) * 0.5; lg[p + 'RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/kernel'][:] = Wr
    lg[p + 'RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/bias'][:] = [0.5, 0]
    W3 = np.zeros((1, 512), np.float32); W3[0, :2] = [1, -2]; lg[p + 'session_representation/FC1/kernel'][:] = W3
    W4 = np.zeros((512, 2), np.float32); W4[0] = [1, 1]; W4[1] = [0, 1]; lg[p + 'session_representation/FC2/kernel'][:] = W4
    M1 = np.zeros((2, 128), np.float32); M1[:, 0] = [1, 2]; M1[:, 1] = [-1, 1]; lg[p + 'recommendations_ranking/matching_dense_layer_1/kernel'][:] = M1
    M2 = np.zeros((128, 64), np.float32); M2[0, 0] = 1; M2[1, 0] = 1; lg[p + 'recommendations_ranking/matching_dense_layer_2/kernel'][:] = M2
    M3 = np.zeros((64, 32), np.float32); M3[0, 0] = 2; lg[p + 'recommendations_ranking/matching_dense_layer_3/kernel'][:] = M3
    M4 = np.zeros((32, 1), np.float32); M4[0, 0] = 1; lg[p + 'recommendations_ranking/matching_dense_layer_4/kernel'][:] = M4
    lg[p + 'recommendations_ranking/matching_dense_layer_4/bias'][:] = [0.25]
    tau, reg = 0.5, 0.01
    o = NarOracle(scfg, acfg, icfg, acr, meta, negative_samples=K, softmax_temperature=tau, reg_weight_decay=reg,
                  CAR_embedding_size=C, rnn_units=H, lr=0.1, dtype=torch.float64)
    o.set_params(lg)
    feats = {'item_clicked': np.array([[1, 2, 0]]), 'event_timestamp': np.array([[10, 20, 0]]), 'session_size': np.array([3]),
             'user_id': np.array([1]), 'session_id': np.array([1]), 'session_start': np.array([1])}
    labels = {'label_next_item': np.array([[2, 3, 0]]), 'label_last_item': np.array([[3]])}
    neg = np.array([[[4, 5], [5, 0], [0, 0]]])
    out = o.forward(feats, labels, neg, np.zeros(4, np.int64), np.full(V, 0.1))

    # ---- scalar restatement
    gamma, beta = [1, 2, 1], [0, 0, 1]

    def car(item):
        x = [0 * gamma[0] + beta[0], acr[item][0] * gamma[1] + beta[1], acr[item][1] * gamma[2] + beta[2]]
        h1 = [leaky(sum(x[i] * W1[i][j] for i in range(3)) + [0, 1][j]) for j in range(2)]
        return [math.tanh(sum(h1[i] * W2[i][j] for i in range(2))) for j in range(2)]
    h = 0.0
    logits_ref, loss = [], 0.0
    for t, (clicked, pos, negs) in enumerate([(1, 2, [4, 5]), (2, 3, [5, 0])]):
        e = car(clicked)
        v = e + [h]
        a = [sum(v[i] * Wr[i][j] for i in range(3)) + [0.5, 0][j] for j in range(2)]
        g = 1 / (1 + math.exp(-(a[0] + 1.0))); c = math.tanh(a[1])
        h = g * h + (1 - g) * c
        f1 = [leaky(h * 1), leaky(h * -2)]
        pred = [math.tanh(f1[0] * 1 + f1[1] * 0), math.tanh(f1[0] * 1 + f1[1] * 1)]
        sc = []
        for cand in [pos] + negs:
            ec = car(cand)
            z = [ec[0] * pred[0], ec[1] * pred[1]]
            z1 = [leaky(z[0] * 1 + z[1] * 2), leaky(-z[0] + z[1])]
            z2 = leaky(z1[0] + z1[1]); z3 = leaky(2 * z2)
            sc.append((z3 + 0.25) / tau)
        logits_ref.append(sc)
        m = max(sc); lse = m + math.log(sum(math.exp(s - m) for s in sc))
        loss += -(sc[0] - lse)
    loss /= 2
    regv = reg * sum(float((w.astype(np.float64) ** 2).sum()) / 2 for k, w in lg.items() if o.regularised(k))
    got = out['logits'].detach().numpy()[0, :2]
    assert np.allclose(got, np.array(logits_ref), atol=1e-6), (got, logits_ref)   # fp32 ACR/pop inputs, fp64 math
    assert abs(float(out['xe_loss']) - loss) < 1e-6
    assert abs(float(out['reg_loss']) - regv) < 1e-9
    assert out['mask'].numpy().tolist() == [[True, True, False]]
    assert np.allclose(out['probs'].sum(-1).detach().numpy(), 1.0)


def test_ugrnn_zero_kernel_and_adam_first_step():
    pb = make_problem('tiny', profile='A')
    o = _oracle(pb)
    lg = pb.layout.init_logical(0)
    lg['main/RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/kernel'][:] = 0
    o.set_params(lg)
    x = torch.randn(3, 4, 64, dtype=torch.float64)
    r = o.rnn(x, torch.tensor([4, 2, 0])).detach()
    assert float(r.abs().max()) == 0.0            # c = tanh(0) = 0, h' = sigmoid(1)*0
    lg['main/RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/bias'][64:] = 0.3     # candidate bias
    o.set_params(lg)
    r = o.rnn(x, torch.tensor([4, 2, 0])).detach()
    g, c = 1 / (1 + math.exp(-1.0)), math.tanh(0.3)
    h1 = (1 - g) * c; h2 = g * h1 + (1 - g) * c
    assert np.allclose(r[0, 0].numpy(), h1) and np.allclose(r[0, 1].numpy(), h2)
    assert float(r[1, 2:].abs().max()) == 0.0 and float(r[2].abs().max()) == 0.0     # zero output past the length
    # TF-Adam, step 1: update = lr * g/(|g| + eps*sqrt(1-b2)) ~ lr*sign(g)
    f, l, neg, buf, pop = _batch(pb)
    before = o.get_params()
    out, grads = o.train_step(f, l, neg, buf, pop)
    k = 'main/CAR/CAR_representation/kernel'
    gk = grads[k].numpy()
    upd = o.get_params()[k] - before[k]
    big = np.abs(gk) > 1e-6
    exact = -o.lr * gk / (np.abs(gk) + 1e-8 / math.sqrt(1 - 0.999))      # closed form of TF-Adam at t = 1
    assert np.allclose(upd[big], exact[big], rtol=1e-6, atol=1e-12)


def test_finite_difference_gradients():
    pb = make_problem('tiny', profile='B', batch_size=6, train_total_negative_samples=4)
    o = _oracle(pb)
    o.set_params(pb.layout.init_logical(5))
    f, l, neg, buf, pop = _batch(pb)
    out = o.forward(f, l, neg, buf, pop)
    grads = o.compute_gradients(out)
    rs = np.random.RandomState(0)
    names = ['main/CAR/PreCAR_representation/kernel', 'main/CAR/CAR_representation/bias',
             'main/RNN/rnn/multi_rnn_cell/cell_0/ugrnn_cell/kernel', 'main/session_representation/FC2/kernel',
             'main/recommendations_ranking/matching_dense_layer_2/kernel',
             'main/user_items_contextual_features/input_features_center_scale/gamma_scale',
             'main/user_items_contextual_features/item_features/item_cat_embedding/items_embedding']
    for n in names:
        w = o.params[n]
        g = grads[n]
        nzi = torch.nonzero(g.abs() > 1e-9)
        assert len(nzi) > 0, n
        for idx in nzi[rs.choice(len(nzi), size=min(3, len(nzi)), replace=False)]:
            idx = tuple(int(i) for i in idx)
            eps = 1e-5
            with torch.no_grad():
                old = float(w[idx]); w[idx] = old + eps
                lp = float(o.forward(f, l, neg, buf, pop)['total_loss'])
                w[idx] = old - eps
                lm = float(o.forward(f, l, neg, buf, pop)['total_loss'])
                w[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - float(g[idx])) <= 1e-5 + 1e-4 * abs(fd), (n, idx, fd, float(g[idx]))


def test_invariants_padding_and_regulariser():
    pb = make_problem('tiny', profile='B', batch_size=8)
    o = _oracle(pb)
    o.set_params(pb.layout.init_logical(2))
    f, l, neg, buf, pop = _batch(pb)
    a = o.forward(f, l, neg, buf, pop)
    # garbage in padded positions of every per-click input must not move the loss (mask, nar_model.py:660-664)
    f2 = {k: v.copy() for k, v in f.items()}
    T = f['item_clicked'].shape[1]
    pad = np.arange(T)[None, :] >= (f['session_size'] - 1)[:, None]
    f2['local_hour_sin'][pad] = 0.77
    f2['os'][pad] = 3
    b = o.forward(f2, l, neg, buf, pop)
    assert abs(float(a['xe_loss']) - float(b['xe_loss'])) < 1e-12
    assert float(a['reg_loss']) > 0
    o0 = _oracle(pb, reg_weight_decay=0.0)
    o0.set_params(pb.layout.init_logical(2))
    assert float(o0.forward(f, l, neg, buf, pop)['reg_loss']) == 0.0
    # recency / novelty features live in [-1, 1] for items inside the statistics' support
    assert np.isfinite(a['x_in'].detach().numpy()).all()


def test_rank_and_metrics_known_answer():
    """top_k order (descending, ties to the lower index), HR@n and MRR@n on a hand-made batch."""
    import torch
    from oracle.nar_oracle import NarOracle
    probs = torch.tensor([[[0.1, 0.5, 0.4], [0.4, 0.4, 0.2]],
                          [[0.6, 0.3, 0.1], [0.2, 0.3, 0.5]]], dtype=torch.float64)
    mask = torch.tensor([[True, True], [True, False]])
    labels = {'label_next_item': np.array([[7, 8], [9, 0]])}
    negatives = np.array([[[1, 2], [3, 4]], [[5, 6], [0, 0]]])
    ids, pr, hits, rr, cnt = NarOracle.rank_and_metrics({'probs': probs, 'mask': mask}, labels, negatives, top_n=2)
    assert ids[0, 0].tolist() == [1, 2, 7] and ids[0, 1].tolist() == [8, 3, 4] and ids[1, 0].tolist() == [9, 5, 6]
    assert np.allclose(pr[0, 0], [0.5, 0.4, 0.1])
    # positive ranks (0-based): 2, 0, 0 over the three valid positions -> two hits at n=2, rr = 1 + 1
    assert (hits, rr, cnt) == (2.0, 2.0, 3.0)


def test_eval_metrics_match_reference_metric_classes():
    """HR@n / MRR@n streaming values of the oracle's rank_and_metrics == the reference's own HitRate / MRR classes
    (tests/golden/metrics_golden.npz, generated by tests/golden/make_metrics_golden.py from nar/metrics.py)."""
    import os
    import torch
    from oracle.nar_oracle import NarOracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics_golden.npz'))
    for topn in (1, 3, 5):
        hits = rr = cnt = 0.0
        for b in range(4):
            probs = g['top%d/b%d/probs' % (topn, b)]
            labels = g['top%d/b%d/labels' % (topn, b)]
            negatives = g['top%d/b%d/negatives' % (topn, b)]
            out = {'probs': torch.from_numpy(probs), 'mask': torch.from_numpy(labels != 0)}
            _, _, h, r, c = NarOracle.rank_and_metrics(out, {'label_next_item': labels}, negatives, topn)
            hits += h; rr += r; cnt += c
            assert abs(hits / cnt - g['top%d/hitrate' % topn][b]) < 1e-12
            assert abs(rr / cnt - g['top%d/mrr' % topn][b]) < 1e-12
