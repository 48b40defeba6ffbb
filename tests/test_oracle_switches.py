"""Oracle branches behind the hparam switches (CPU): novelty regulariser (nar_model.py:517, :531-544, :673-683) and
dropout (:338-340, :417-419, :1330-1333) - hand-computed answers and invariants of the restatement itself."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')


def _problem(**hp):
    from chameleon_recsys_b200.harness import make_problem, warm_state
    pb = make_problem('tiny', profile='B', batch_size=8, **hp)
    warm_state(pb, 3)
    f, l = pb.input_fn().get_next()
    buf = pb.clicked_items_state.get_recent_clicks_buffer().copy()
    pop = pb.clicked_items_state.get_articles_recent_pop_norm().copy()
    return pb, f, l, buf, pop


def _negatives(pb, f, l, buf, step=1):
    from oracle import sampler_ref
    hp = pb.hp
    allc = np.concatenate([f['item_clicked'], l['label_last_item']], axis=1)
    return sampler_ref.sample_negatives(allc, buf, hp.train_total_negative_samples, hp.train_negative_samples_from_buffer,
                                        hp.sampler_seed, step)


def test_novelty_regulariser_known_answer():
    from tools.gpu_step_check import make_oracle
    pb, f, l, buf, pop = _problem(novelty_reg_factor=0.7)
    neg = _negatives(pb, f, l, buf)
    orc = make_oracle(pb, torch.float64)
    orc.set_params(pb.layout.init_logical(3))
    o = orc.forward(f, l, neg, buf, pop)
    base = make_oracle(pb.__class__(**{**pb.__dict__, 'hp': pb.hp.copy(novelty_reg_factor=0.0)}), torch.float64)
    base.set_params(pb.layout.init_logical(3))
    o0 = base.forward(f, l, neg, buf, pop)
    # scalar restatement: softmax over the negatives only, novelty = -log2(pop_norm[id]), masked mean over valid positions
    lg = o0['logits'].detach().numpy()[:, :, 1:]                      # already / temperature
    q = np.exp(lg - lg.max(-1, keepdims=True)); q /= q.sum(-1, keepdims=True)
    nov = -np.log(np.asarray(pop, dtype=np.float32).astype(np.float64)[neg]) / np.log(pb.hp.popularity_smooth_log_base)
    mask = o0['mask'].numpy()
    want = 0.7 * ((q * nov).sum(-1) * mask).sum() / mask.sum()
    assert abs(float(o['nov_reg_loss']) - want) < 1e-9
    assert abs(float(o['total_loss']) - (float(o0['total_loss']) - want)) < 1e-9
    assert float(o0['nov_reg_loss']) == 0.0


def test_dropout_masks_and_scaling():
    from oracle import dropout_ref
    from tools.gpu_step_check import make_oracle
    pb, f, l, buf, pop = _problem(dropout_keep_prob=0.6)
    neg = _negatives(pb, f, l, buf)
    orc = make_oracle(pb, torch.float64)
    orc.set_params(pb.layout.init_logical(3))
    plain = orc.forward(f, l, neg, buf, pop)                           # no train_step: inference, no dropout
    a = orc.forward(f, l, neg, buf, pop, train_step=4)
    b = orc.forward(f, l, neg, buf, pop, train_step=4)
    c = orc.forward(f, l, neg, buf, pop, train_step=5)
    assert torch.equal(a['logits'], b['logits'])                       # counter based: same step, same masks
    assert not torch.equal(a['logits'], c['logits'])
    xi, xp = plain['x_in'].detach().numpy(), a['x_in'].detach().numpy()
    B, T = f['item_clicked'].shape
    pos = np.arange(B)[:, None] * T + np.arange(T)[None, :]
    mi = dropout_ref.keep_mask(pb.hp.sampler_seed, 4, 1, pos, pb.plan.Fp, 0.6)
    valid = pb.plan.int2log >= 0
    m = np.zeros(xi.shape, dtype=bool)
    m[..., pb.plan.int2log[valid]] = mi[..., valid]
    assert np.allclose(xp, xi * m / 0.6)
    assert 0.5 < m.mean() < 0.7
    # a data-parallel shard draws the masks of its own rows (row key = global position)
    half = {k: v[B // 2:] for k, v in f.items()}
    lh = {k: v[B // 2:] for k, v in l.items()}
    sh = orc.forward(half, lh, neg[B // 2:], buf, pop, train_step=4, session0=B // 2)
    assert np.allclose(sh['x_in'].detach().numpy(), xp[B // 2:])
