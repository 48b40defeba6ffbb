"""Generate tests/golden/hook_golden.npz by running the REFERENCE ItemsStateUpdaterHook (nar_model.py:1370-1650, imported
unmodified on the TF-API stand-in tests/golden/tf1_shim.py: the hook only needs tf.train.SessionRunHook / SessionRunArgs
and tf.estimator.ModeKeys) together with the REFERENCE ClickedItemsState over a few training batches: before_run's feed
(recent-clicks buffer, recent popularity) and after_run's state update (flattening of [clicked | last label], the
timestamp borrowed for the label click, nar_model.py:1635-1650).  Run once in the build container; the .npz is committed."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf1_shim as shim  # noqa: E402,F401
import pandas  # noqa: E402,F401

sys.modules.setdefault('pytz', types.ModuleType('pytz'))
_ua = types.ModuleType('ua_parser')
_ua.user_agent_parser = types.ModuleType('ua_parser.user_agent_parser')
sys.modules.setdefault('ua_parser', _ua)
sys.modules.setdefault('ua_parser.user_agent_parser', _ua.user_agent_parser)
pkg = types.ModuleType('refnar')
pkg.__path__ = ['/root/reference/nar_module/nar']
sys.modules['refnar'] = pkg
ref_model = importlib.import_module('refnar.nar_model')
ref_state = importlib.import_module('refnar.clicked_items_state')

from chameleon_recsys_b200.harness import make_problem  # noqa: E402

out = {}
for ci, (hours, max_size, n_norm) in enumerate([(1.0, 2000, 500), (0.02, 150, 40)]):
    pb = make_problem('tiny', profile='B')
    V = pb.plan.num_items
    state = ref_state.ClickedItemsState(hours, max_size, n_norm, V)
    model = types.SimpleNamespace(item_clicked='item_clicked', event_timestamp='event_timestamp', next_item_label='next_item_label',
                                  label_last_item='label_last_item', session_id='session_id', user_id='user_id',
                                  articles_recent_pop_norm='articles_recent_pop_norm', pop_recent_items_buffer='pop_recent_items_buffer',
                                  content_article_embeddings_matrix='content_article_embeddings_matrix', articles_metadata={})
    hook = ref_model.ItemsStateUpdaterHook('train', model, 3, state, [], None, None, None, {}, 0.1)
    hook.begin()
    it = pb.input_fn()
    for step in range(5):
        f, l = it.get_next()
        args = hook.before_run(None)
        feed = args.kwargs['feed_dict']
        out['c%d_feed_buffer_%d' % (ci, step)] = np.asarray(feed['pop_recent_items_buffer']).copy()
        out['c%d_feed_pop_norm_%d' % (ci, step)] = np.asarray(feed['articles_recent_pop_norm']).copy()
        results = {'clicked_items': f['item_clicked'], 'clicked_timestamps': f['event_timestamp'][..., None],
                   'next_item_labels': l['label_next_item'], 'last_item_label': l['label_last_item'],
                   'user_id': f['user_id'], 'session_id': f['session_id']}
        hook.after_run(None, types.SimpleNamespace(results=results))
        for k in ('item_clicked', 'event_timestamp'):
            out['c%d_%s_%d' % (ci, k, step)] = f[k]
        out['c%d_label_last_item_%d' % (ci, step)] = l['label_last_item']
        out['c%d_buffer_%d' % (ci, step)] = state.pop_recent_clicks_buffer.copy()
        out['c%d_recent_pop_%d' % (ci, step)] = state.get_articles_recent_pop().copy()
        out['c%d_pop_norm_%d' % (ci, step)] = state.get_articles_recent_pop_norm().copy()
        out['c%d_pop_%d' % (ci, step)] = state.get_articles_pop().copy()
    out['c%d_cfg' % ci] = np.array([hours, max_size, n_norm, V], dtype=np.float64)
np.savez_compressed(os.path.join(HERE, 'hook_golden.npz'), **out)
print('wrote', len(out), 'arrays', os.path.getsize(os.path.join(HERE, 'hook_golden.npz')) // 1024, 'KB')
