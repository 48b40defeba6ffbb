"""Per-op device time of one training step (NAR_PROFILE=1 wrappers in ops.py).  python tools/step_profile.py [dense]"""
import json
import os
import sys

os.environ['NAR_PROFILE'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from chameleon_recsys_b200 import ops  # noqa: E402
from chameleon_recsys_b200.harness import make_problem, warm_state  # noqa: E402
from tools.gpu_step_check import make_engine  # noqa: E402

sl = 'dense' if len(sys.argv) > 1 and sys.argv[1] == 'dense' else 'g1'
pb = make_problem('g1', profile='B', session_len=sl)
warm_state(pb, 50)
batches = bench.make_batches(pb, 8, pb.hp.batch_size)
eng = make_engine(pb)
eng.set_params(pb.layout.init_logical(42))
staged = [eng.stage(f, l, b, p, slot='s%d' % i) for i, (f, l, b, p) in enumerate(batches)]
for st in staged[:3]:
    eng.grads.zero_(); eng.step(st); eng.apply_gradients()
torch.cuda.synchronize()
ops.profile_reset()
n = 0
for st in staged[3:]:
    eng.grads.zero_(); eng.step(st); eng.apply_gradients(); n += 1
rows, tot = ops.profile_report(n)
print('session_len', sl, 'L', [s['L'] for s in staged[3:]], 'sum of op times per step: %.1f us' % tot)
for r in rows[:45]:
    print('%-46s x%4.1f %9.1f us %5.1f%%' % (r['op'], r['calls_per_step'], r['us_per_step'], r['pct']))
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'step_profile_%s.json' % sl), 'w'), indent=1)
