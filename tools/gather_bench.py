"""Time the feature-gather kernel (forward + backward) alone on one G1-shaped step: CUDA events, L2 flushed.
   python tools/gather_bench.py [--profile B] [--session-len g1]
Prints one JSON line per kernel (algorithmic bytes per SURVEY.md 8(d), and all bytes the kernel moves)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='g1')
    ap.add_argument('--profile', default='B')
    ap.add_argument('--session-len', default='g1')
    ap.add_argument('--iters', type=int, default=30)
    args = ap.parse_args()
    import torch
    import bench
    from chameleon_recsys_b200 import ops
    from chameleon_recsys_b200.estimator import build_estimator
    from chameleon_recsys_b200.harness import make_problem, warm_state
    torch.cuda.set_device(0)
    pb = make_problem(args.workload, profile=args.profile, session_len=args.session_len)
    warm_state(pb, 50)
    batches = bench.make_batches(pb, 2, pb.hp.batch_size)
    est = build_estimator(None, pb.content_article_embeddings_matrix, pb.articles_metadata, pb.articles_features_config,
                          pb.session_features_config, pb.hp, pb.clicked_items_state, device=0)
    eng = est._ensure_spec(None, None).model.engine
    f, l, buf, pop = batches[1]
    st = eng.stage(f, l, buf, pop, slot='gb')
    eng.step(st, train=False, keep=True)
    L, K = st['L'], eng.K
    R = L + L * (K + 1)
    plan = eng.plan
    planc = eng.feature_plan_c(st)
    t = st['t']
    X = torch.empty(R, plan.Fp, device='cuda')
    rows = ops.row_layout(R, L, K + 1, ctx_col0=plan.ctx_col0)
    dX = torch.randn(R, plan.Fp, device='cuda')
    row_pos, row_item = eng.last['row_pos'], eng.last['row_item']
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')

    def timeit(fn):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))

    E = plan.acr_dim if plan.use_acr else 0
    Di = plan.item_emb_dim if plan.use_item_emb else 0
    gbytes = L * ((2 + K) * (E + Di) * 4 * 2 + (2 + K) * 8)
    moved = R * plan.Fp * 4 + R * (E + Di) * 4 + R * 12
    med, mn = timeit(lambda: ops.gather_features(planc, row_pos, row_item, rows, t['event_ts'], t['max_ts'], X))
    print(json.dumps({'kernel': 'gather_features', 'rows': R, 'Fp': plan.Fp, 'us_median': med * 1e3, 'us_min': mn * 1e3,
                      'algorithmic_GBps': gbytes / (med * 1e-3) / 1e9, 'moved_GBps': moved / (med * 1e-3) / 1e9,
                      'minb': os.environ.get('NAR_GATHER_MINB', 'default')}))
    # calibration: plain row gather of an ACR-shaped table with the same ids, a same-size device copy, an empty launch
    V = int(plan.num_items)
    tab = torch.randn(V, 252, device='cuda')
    med, mn = timeit(lambda: ops.gather_rows(tab, row_item, X, 248))
    print(json.dumps({'kernel': 'gather_rows(248 of ld 252 -> ld %d)' % plan.Fp, 'us_median': med * 1e3, 'GBps': R * 248 * 8 / (med * 1e-3) / 1e9}))
    tab2 = torch.randn(V, 480, device='cuda')
    med, mn = timeit(lambda: ops.gather_rows(tab2, row_item, X, 480))
    print(json.dumps({'kernel': 'gather_rows(480 -> 480)', 'us_median': med * 1e3, 'GBps': R * 480 * 8 / (med * 1e-3) / 1e9}))
    X2 = torch.empty_like(X)
    med, mn = timeit(lambda: X2.copy_(X))
    print(json.dumps({'kernel': 'copy R x Fp', 'us_median': med * 1e3, 'GBps': R * plan.Fp * 8 / (med * 1e-3) / 1e9}))
    z = torch.zeros(4, device='cuda')
    med, mn = timeit(lambda: z.zero_())
    print(json.dumps({'kernel': 'empty launch', 'us_median': med * 1e3}))
    dg = torch.zeros(plan.Fp, device='cuda'); db = torch.zeros(plan.Fp, device='cuda')
    planb = eng.feature_plan_c(st)
    med, mn = timeit(lambda: ops.gather_features_bwd(planb, row_pos, row_item, rows, t['event_ts'], t['max_ts'], dX, dg, db))
    print(json.dumps({'kernel': 'gather_features_bwd', 'rows': R, 'us_median': med * 1e3, 'us_min': mn * 1e3,
                      'read_GBps': R * plan.Fp * 4 / (med * 1e-3) / 1e9}))


if __name__ == '__main__':
    main()
