"""Generate tests/golden/sampler_freq.json from the REFERENCE's numpy sampler
(/root/reference/nar_module/nar/benchmarks/candidate_sampling.py, loaded by file path because
the package __init__ imports tensorflow).  Run once in the build container; the JSON is committed."""
import importlib.util
import json
import os

import numpy as np

REF = '/root/reference/nar_module/nar/benchmarks/candidate_sampling.py'
spec = importlib.util.spec_from_file_location('ref_candidate_sampling', REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

np.random.seed(42)
# a pool with repeated items (popularity): item i appears w_i times
weights = {1: 12, 2: 8, 3: 6, 4: 4, 5: 3, 6: 2, 7: 2, 8: 1, 9: 1, 10: 1, 11: 1, 12: 1}
pool = np.concatenate([[i] * w for i, w in weights.items()]).astype(np.int64)
np.random.shuffle(pool)
K, trials = 5, 40000
m = mod.CandidateSamplingManager(lambda: np.zeros(1, dtype=np.int64))
items = np.array(sorted(weights))
counts = np.zeros(len(items))
for _ in range(trials):
    s = m.get_neg_items_click(pool, K)
    counts += np.isin(items, s)
out = {'pool': pool.tolist(), 'K': K, 'trials': trials, 'items': items.tolist(), 'freq': (counts / trials).tolist(),
       'source': REF + ':25-37 get_neg_items_click, np.random.seed(42)'}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sampler_freq.json'), 'w') as f:
    json.dump(out, f)
print(out['freq'])
