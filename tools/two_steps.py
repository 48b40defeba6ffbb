"""Two training steps of the G1 workload (used under `ncu --set full -k regex:<kernel>` for the captures in profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from chameleon_recsys_b200.harness import make_problem, warm_state  # noqa: E402
from tools.gpu_step_check import make_engine  # noqa: E402

pb = make_problem('g1', profile='B')
warm_state(pb, 50)
batches = bench.make_batches(pb, 2, pb.hp.batch_size)
eng = make_engine(pb)
eng.set_params(pb.layout.init_logical(42))
for i, (f, l, b, p) in enumerate(batches):
    out = eng.train_step(f, l, b, p)
print('loss', out['total_loss'], 'L', out['L'])
