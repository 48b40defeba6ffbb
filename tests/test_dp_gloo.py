"""Data-parallel host logic with two real processes over gloo (CPU): session sharding, global loss
normaliser, rank-independent negatives, gradient sum-allreduce == single-process gradient (oracle math)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from chameleon_recsys_b200.dp import shard_sessions
    from chameleon_recsys_b200.harness import make_problem, warm_state
    from oracle import sampler_ref
    from tools.gpu_step_check import make_oracle
    torch.set_num_threads(2)
    pb = make_problem('tiny', profile='B', batch_size=16)
    hp = pb.hp
    warm_state(pb, 3)
    f, l = pb.input_fn().get_next()                       # every rank builds the same global batch
    T = f['item_clicked'].shape[1]
    sh = shard_sessions(f['session_size'], T, world, rank)
    s0, per = sh['s0'], sh['per']
    # L is additive over ranks and the normaliser every rank uses is the global one
    Lsum = torch.tensor([sh['L']], dtype=torch.int64)
    dist.all_reduce(Lsum)
    assert int(Lsum) == sh['L_global'] == int(np.clip(f['session_size'] - 1, 0, T).sum())
    # negatives of the local shard == rows of the global result
    allc = np.concatenate([f['item_clicked'], l['label_last_item']], axis=1)
    buf = pb.clicked_items_state.get_recent_clicks_buffer().copy()
    pop = pb.clicked_items_state.get_articles_recent_pop_norm().copy()
    K, nfb = hp.train_total_negative_samples, hp.train_negative_samples_from_buffer
    neg_local = sampler_ref.sample_negatives(allc[s0:s0 + per], buf, K, nfb, 42, 1, session_offset=s0,
                                             all_clicked_items_global=allc)
    neg_global = sampler_ref.sample_negatives(allc, buf, K, nfb, 42, 1)
    assert np.array_equal(neg_local, neg_global[s0:s0 + per])
    # local loss with the global normaliser; gradients sum-allreduced == global gradients
    orc = make_oracle(pb, torch.float64)
    orc.reg = 0.0                                          # the regulariser is added once, after the allreduce
    orc.set_params(pb.layout.init_logical(42))
    fl = {k: v[s0:s0 + per] for k, v in f.items()}
    ll = {k: v[s0:s0 + per] for k, v in l.items()}
    out = orc.forward(fl, ll, neg_local, buf, pop, sum_mask_global=sh['L_global'])
    grads = orc.compute_gradients(out)
    flat = torch.cat([g.reshape(-1) for g in grads.values()])
    dist.all_reduce(flat)
    loss = out['xe_loss'].detach().clone()
    dist.all_reduce(loss)
    full = orc.forward(f, l, neg_global, buf, pop)
    gfull = torch.cat([g.reshape(-1) for g in orc.compute_gradients(full).values()])
    assert abs(float(loss) - float(full['xe_loss'])) < 1e-10
    assert float((flat - gfull).abs().max()) < 1e-10
    # pos_idx addresses the global arrays
    pos = sh['pos_idx']
    assert (pos // T >= s0).all() and (pos // T < s0 + per).all()
    assert (f['item_clicked'].reshape(-1)[pos] != 0).all()
    ret[rank] = 1
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}
