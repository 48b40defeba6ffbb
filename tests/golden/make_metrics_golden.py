"""Generate tests/golden/metrics_golden.npz with the REFERENCE's own streaming metric classes
(/root/reference/nar_module/nar/metrics.py: HitRate :109-134, MRR :40-66; numpy + sklearn only, loaded by file path).
Inputs: per-batch candidate probabilities [B,T,1+K], labels [B,T] (0 = padding), negatives [B,T,K]; the predictions fed
to the reference classes are the candidates sorted by descending probability with ties to the lower index - the order
tf.nn.top_k produces in rank_items_by_predicted_prob (nar_model.py:777-795).  Run once in the build container; the
.npz is committed."""
import importlib.util
import os

import numpy as np

REF = '/root/reference/nar_module/nar/metrics.py'
spec = importlib.util.spec_from_file_location('ref_metrics', REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

rs = np.random.RandomState(42)
B, T, K = 6, 5, 9
out = {}
for topn in (1, 3, 5):
    hr, mrr = mod.HitRate(topn), mod.MRR(topn)
    stream_hr, stream_mrr = [], []
    for batch in range(4):
        lengths = rs.randint(0, T + 1, B)
        labels = np.zeros((B, T), dtype=np.int64)
        negatives = np.zeros((B, T, K), dtype=np.int64)
        probs = np.zeros((B, T, 1 + K))
        for b in range(B):
            for t in range(T):
                ids = rs.permutation(np.arange(1, 500))[:1 + K]
                p = rs.dirichlet(np.ones(1 + K))
                if rs.rand() < 0.3:                      # exact ties, including with the positive
                    p[rs.randint(1, 1 + K)] = p[0]
                    p = p / p.sum()
                probs[b, t] = p
                negatives[b, t] = ids[1:]
                if t < lengths[b]:
                    labels[b, t] = ids[0]
                else:
                    negatives[b, t, rs.randint(0, K):] = 0   # padded positions / short pools carry zero ids
        ids_all = np.concatenate([labels[..., None], negatives], axis=2)
        order = np.argsort(-probs, axis=2, kind='stable')
        preds = np.take_along_axis(ids_all, order, axis=2)
        hr.add(preds, labels); mrr.add(preds, labels)
        stream_hr.append(hr.result()); stream_mrr.append(mrr.result())
        out['top%d/b%d/probs' % (topn, batch)] = probs
        out['top%d/b%d/labels' % (topn, batch)] = labels
        out['top%d/b%d/negatives' % (topn, batch)] = negatives
    out['top%d/hitrate' % topn] = np.asarray(stream_hr)
    out['top%d/mrr' % topn] = np.asarray(stream_mrr)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'metrics_golden.npz'), **out)
print({k: v for k, v in out.items() if k.endswith('hitrate') or k.endswith('mrr')})
