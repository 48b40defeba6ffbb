"""Diagnostic used while chasing the sporadic gradient discrepancies of steps with dropout (resolved: leaky_relu kink flips,
DESIGN.md section 3 - gradients are now compared at identical slope choices).  Prints, per case and step, the forward
errors and the worst gradient tensors (relative, absolute error of absolute max); NAR_DEBUG_DROP_ONLY isolates one site."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tools import gpu_step_check as g  # noqa: E402

CASES = [('u1_k7', dict(dropout_keep_prob=0.7)), ('u2_k7', dict(rnn_num_layers=2, dropout_keep_prob=0.7)),
         ('u2_k7_only4', dict(rnn_num_layers=2, dropout_keep_prob=0.7)), ('u2_k7_only0', dict(rnn_num_layers=2, dropout_keep_prob=0.7)),
         ('u2_k7_only8', dict(rnn_num_layers=2, dropout_keep_prob=0.7)), ('u2_k7_only9', dict(rnn_num_layers=2, dropout_keep_prob=0.7))]
for name, hp in CASES:
    if '_only' in name:
        os.environ['NAR_DEBUG_DROP_ONLY'] = name.split('_only')[1]
    res = g.run_case('tiny', 'B', 5, 6, hp_over=hp, oracle_dtype=torch.float64, engine_kw=dict(bwd_precision=3, fwd_precision=3))
    os.environ.pop('NAR_DEBUG_DROP_ONLY', None)
    for s in res['steps']:
        worst = sorted(s['grad_rel'].items(), key=lambda kv: -kv[1])[:3]
        print(name, 'step', s['step'], 'L', s['L'], 'T', s['T'], 'fwd %.1e %.1e %.1e' % (s['rnn'], s['pred'], s['logits_rel_max']),
              'grad_max %.2e' % s['grad_rel_max'], [(k, '%.1e' % v, 'abs %.1e of %.1e' % s['grad_abs'][k]) for k, v in worst])
