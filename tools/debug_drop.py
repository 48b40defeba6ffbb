import sys, json
sys.path.insert(0, '/root/repo')
import torch
from tools import gpu_step_check as g
for name, hp, ekw in [('u1_k8', dict(dropout_keep_prob=0.8), dict(bwd_precision=3)),
                      ('u1_k7', dict(dropout_keep_prob=0.7), dict(bwd_precision=3)),
                      ('u2_k8', dict(rnn_num_layers=2, dropout_keep_prob=0.8), dict(bwd_precision=3)),
                      ('u2_k7', dict(rnn_num_layers=2, dropout_keep_prob=0.7), dict(bwd_precision=3)),
                      ('u2_k1', dict(rnn_num_layers=2), dict(bwd_precision=3))]:
    res = g.run_case('tiny', 'B', 5, 2, hp_over=hp, oracle_dtype=torch.float64, engine_kw=ekw)
    for s in res['steps']:
        worst = sorted(s['grad_rel'].items(), key=lambda kv: -kv[1])[:4]
        print(name, 'step', s['step'], 'L', s['L'], 'fwd', '%.1e %.1e %.1e' % (s['rnn'], s['pred'], s['logits_rel_max']), 'grad_max %.2e' % s['grad_rel_max'], [(k, '%.1e' % v) for k, v in worst])
