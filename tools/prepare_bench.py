"""Time the weight-independent front of a step (sampler, row lists, statistics) alone, for a global batch of
N x 256 sessions of which this rank owns 256 (what rank 0 of an N-GPU data-parallel job runs).
   python tools/prepare_bench.py 1 8"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from chameleon_recsys_b200 import ops
    from chameleon_recsys_b200.estimator import build_estimator
    from chameleon_recsys_b200.harness import make_problem, warm_state
    torch.cuda.set_device(0)
    for world in [int(x) for x in sys.argv[1:]] or [1, 8]:
        pb = make_problem('g1', profile='B', session_len='g1')
        warm_state(pb, 50)
        batches = bench.make_batches(pb, 3, pb.hp.batch_size * world)
        est = build_estimator(None, pb.content_article_embeddings_matrix, pb.articles_metadata, pb.articles_features_config,
                              pb.session_features_config, pb.hp, pb.clicked_items_state, device=0)
        eng = est._ensure_spec(None, None).model.engine
        eng.world, eng.rank = world, 0          # shard like rank 0 of `world` (no collective is issued by prepare)
        f, l, buf, pop = batches[1]
        st = eng.stage(f, l, buf, pop, slot='pb')
        ts = []
        for it in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.prepare(st, it + 1); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(json.dumps({'world': world, 'Bg': st['Bg'], 'B': st['B'], 'L': st['L'], 'prepare_ms_median': float(np.median(ts[2:])),
                          'all': [round(x, 3) for x in ts]}))


if __name__ == '__main__':
    main()
