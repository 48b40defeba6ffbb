"""Generate tests/golden/sampler_tf_freq.npz: inclusion frequencies of the REFERENCE's training-graph sampler
(/root/reference/nar_module/nar/nar_model.py:1220-1304: get_sample_from_recently_clicked_items_buffer ->
get_batch_negative_samples -> get_negative_samples -> get_neg_items_session -> get_neg_items_click, imported unmodified)
executed on the TF-API stand-in tests/golden/tf1_shim.py, over many independent shuffles of one small scenario.  TF's
shuffles cannot be reproduced draw by draw; what oracle/sampler_ref.py must share with them is the DISTRIBUTION: which ids
end up among a click's K negatives how often (popularity weighting through repeated pool entries, the size of the buffer
sample, the K*20 truncation of the pool, per-session exclusion, no repetition within a click).
Run once in the build container; the .npz is committed."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf1_shim as shim  # noqa: E402
import pandas  # noqa: E402,F401

sys.modules.setdefault('pytz', types.ModuleType('pytz'))
_ua = types.ModuleType('ua_parser')
_ua.user_agent_parser = types.ModuleType('ua_parser.user_agent_parser')
sys.modules.setdefault('ua_parser', _ua)
sys.modules.setdefault('ua_parser.user_agent_parser', _ua.user_agent_parser)
pkg = types.ModuleType('refnar')
pkg.__path__ = ['/root/reference/nar_module/nar']
sys.modules['refnar'] = pkg
ref = importlib.import_module('refnar.nar_model')

rs = np.random.RandomState(5)
V = 40
# sessions: [clicked ..., last label], zero padded; popular items repeat across sessions
all_clicked = np.array([[3, 7, 3, 9], [5, 3, 0, 7], [11, 0, 0, 3], [7, 12, 5, 13], [3, 14, 0, 5], [15, 7, 16, 3]], dtype=np.int64)
buffer = np.concatenate([rs.zipf(1.5, 30) % (V - 1) + 1, np.zeros(10)]).astype(np.int64)
rs.shuffle(buffer)
out = {'all_clicked': all_clicked, 'buffer': buffer, 'V': np.array(V)}
for ci, (K, n_from_buffer, trials) in enumerate([(3, 12, 3000), (1, 12, 3000)]):      # K=1: the pool (cap 20) is truncated
    shim.configure(float64=False, seed=100 + ci)
    m = object.__new__(ref.NARModuleModel)
    m.pop_recent_items_buffer = shim._t(buffer)
    B, T1 = all_clicked.shape
    counts = np.zeros((B, T1 - 1, V), dtype=np.int64)
    pad = 0
    for _ in range(trials):
        extra = m.get_sample_from_recently_clicked_items_buffer(n_from_buffer)
        neg = m.get_batch_negative_samples(shim._t(all_clicked), additional_samples=extra, num_negative_samples=K)
        neg = neg[:, :-1, :].numpy()
        for b in range(B):
            for t in range(T1 - 1):
                row = neg[b, t]
                counts[b, t, row[row != 0]] += 1
                pad += int((row == 0).sum()) if all_clicked[b, t] != 0 else 0
    out['c%d_cfg' % ci] = np.array([K, n_from_buffer, trials])
    out['c%d_freq' % ci] = counts / float(trials)
    out['c%d_pad_per_trial' % ci] = np.array(pad / float(trials))
np.savez_compressed(os.path.join(HERE, 'sampler_tf_freq.npz'), **out)
print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})
print(out['c0_freq'][0, 0], out['c1_freq'][0, 0], out['c0_pad_per_trial'], out['c1_pad_per_trial'])
