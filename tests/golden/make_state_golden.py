"""Generate tests/golden/state_golden.npz by running the REFERENCE ClickedItemsState
(/root/reference/nar_module/nar/clicked_items_state.py) on seeded click batches.  The reference
package is loaded through a shim package so that its relative imports resolve (numpy / scipy /
sklearn only; no tensorflow).  Run once in the build container; the .npz is committed."""
import importlib
import os
import sys
import types

import numpy as np

REF_DIR = '/root/reference/nar_module/nar'
pkg = types.ModuleType('refnar')
pkg.__path__ = [REF_DIR]
sys.modules['refnar'] = pkg
ref = importlib.import_module('refnar.clicked_items_state')

rs = np.random.RandomState(7)
V = 300
cases = {}
for ci, (hours, max_size, n_norm) in enumerate([(1.0, 64, 20), (0.05, 200, 50), (1.0, 1000, 500)]):
    st = ref.ClickedItemsState(hours, max_size, n_norm, V)
    t = 1506826800000
    out = {}
    for step in range(6):
        n = int(rs.randint(5, 60))
        items = (rs.zipf(1.4, n) % (V - 1) + 1).astype(np.int64)
        ts = (t + np.sort(rs.randint(0, 240000, n))).astype(np.int64)
        t += int(rs.randint(60000, 400000))
        st.update_items_state(items, ts)
        out['items_%d' % step] = items
        out['ts_%d' % step] = ts
        out['buffer_%d' % step] = st.pop_recent_clicks_buffer.copy()
        out['recent_pop_%d' % step] = st.get_articles_recent_pop().copy()
        out['pop_norm_%d' % step] = st.get_articles_recent_pop_norm().copy()
        out['pop_%d' % step] = st.get_articles_pop().copy()
    for k, v in out.items():
        cases['c%d_%s' % (ci, k)] = v
    cases['c%d_cfg' % ci] = np.array([hours, max_size, n_norm, V], dtype=np.float64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'state_golden.npz'), **cases)
print('wrote', len(cases), 'arrays')
