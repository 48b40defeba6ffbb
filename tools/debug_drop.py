import sys, json
sys.path.insert(0, '/root/repo')
import torch
from tools import gpu_step_check as g
for name, hp, ekw in [
                      ('u2_k7', dict(rnn_num_layers=2, dropout_keep_prob=0.7), dict(bwd_precision=3)),
                      ('u2_k7_noaux', dict(rnn_num_layers=2, dropout_keep_prob=0.7), dict(bwd_precision=3)),
                      ('g2_k7', dict(rnn_num_layers=2, dropout_keep_prob=0.7, rnn_cell='gru'), dict(bwd_precision=3))]:
    import os
    os.environ['NAR_AUX_STREAM'] = '0' if name.endswith('noaux') else '1'
    res = g.run_case('tiny', 'B', 5, 3, hp_over=hp, oracle_dtype=torch.float64, engine_kw=ekw)
    for s in res['steps']:
        worst = sorted(s['grad_rel'].items(), key=lambda kv: -kv[1])[:4]
        print(name, 'step', s['step'], 'L', s['L'], 'fwd', '%.1e %.1e %.1e' % (s['rnn'], s['pred'], s['logits_rel_max']), 'grad_max %.2e' % s['grad_rel_max'], [(k, '%.1e' % v, 'abs %.1e of %.1e' % s['grad_abs'][k]) for k, v in worst])
