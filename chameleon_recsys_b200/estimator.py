"""Estimator-shaped boundary of the NAR hot path (the reference's drop-in surface).

* ``nar_module_model_fn(features, labels, mode, params)`` - nar_trainer_gcom.py:234-332: picks the
  train / eval negative-sampling hparams by mode (:237-242), forces keep_prob 1 in eval (:245),
  builds ``NARModuleModel`` (:252-275) and returns an ``EstimatorSpec`` with ``loss``, ``train_op``
  and the ``ItemsStateUpdaterHook`` as training chief hook (:305-322).
* ``build_estimator`` / ``Estimator.train`` - nar_trainer_gcom.py:335-386, :511-517: the minimal
  MonitoredTrainingSession loop: hook.before_run -> train_op -> hook.after_run per batch.

Differences that are inherent to not being a TF graph: ``features``/``labels`` are numpy dicts
(one padded batch, same keys / dtypes / padding as datasets.py), ``train_op`` is a callable, and
the model object is cached across ``train()`` calls instead of being rebuilt from a checkpoint.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import checkpoint as ckpt
from .clicked_items_state import ClickedItemsState
from .datasets import OutOfRangeError
from .hparams import ModeKeys, get_internal_enabled_features_config
from .nar_model import ItemsStateUpdaterHook, NARModuleModel

# Global vars updated by the Estimator hook (nar_trainer_gcom.py:410-415)
clicked_items_state: Optional[ClickedItemsState] = None
eval_sessions_metrics_log: list = []


@dataclass
class EstimatorSpec:
    mode: str
    loss: Optional[float] = None
    train_op: Optional[Callable] = None
    training_chief_hooks: List = field(default_factory=list)
    eval_metric_ops: Optional[dict] = None
    evaluation_hooks: List = field(default_factory=list)
    model: Optional[NARModuleModel] = None


def nar_module_model_fn(features, labels, mode, params) -> EstimatorSpec:
    if mode == ModeKeys.TRAIN:
        negative_samples = params['train_total_negative_samples']
        negative_sample_from_buffer = params['train_negative_samples_from_buffer']
    elif mode == ModeKeys.EVAL:
        negative_samples = params['eval_total_negative_samples']
        negative_sample_from_buffer = params['eval_negative_samples_from_buffer']
    else:
        raise ValueError('mode %r' % (mode,))
    dropout_keep_prob = params['dropout_keep_prob'] if mode == ModeKeys.TRAIN else 1.0
    internal_features_config = params.get('internal_features_config') or get_internal_enabled_features_config()
    eval_metrics_top_n = params['eval_metrics_top_n']

    model = NARModuleModel(mode, features, labels,
                           session_features_config=params['session_features_config'],
                           articles_features_config=params['articles_features_config'],
                           batch_size=params['batch_size'],
                           lr=params['lr'],
                           keep_prob=dropout_keep_prob,
                           negative_samples=negative_samples,
                           negative_sample_from_buffer=negative_sample_from_buffer,
                           reg_weight_decay=params['reg_weight_decay'],
                           softmax_temperature=params['softmax_temperature'],
                           articles_metadata=params['articles_metadata'],
                           content_article_embeddings_matrix=params['content_article_embeddings_matrix'],
                           recent_clicks_buffer_hours=params['recent_clicks_buffer_hours'],
                           recent_clicks_buffer_max_size=params['recent_clicks_buffer_max_size'],
                           recent_clicks_for_normalization=params['recent_clicks_for_normalization'],
                           CAR_embedding_size=params['CAR_embedding_size'],
                           rnn_units=params['rnn_units'],
                           rnn_num_layers=params.get('rnn_num_layers', 1),
                           metrics_top_n=eval_metrics_top_n,
                           plot_histograms=params['save_histograms'],
                           novelty_reg_factor=params['novelty_reg_factor'],
                           diversity_reg_factor=params['diversity_reg_factor'],
                           internal_features_config=internal_features_config,
                           eval_cold_start=params['eval_cold_start'],
                           elapsed_days_smooth_log_base=params.get('elapsed_days_smooth_log_base', 1.3),
                           popularity_smooth_log_base=params.get('popularity_smooth_log_base', 2.0),
                           max_cardinality_for_ohe=params.get('max_cardinality_for_ohe', 10),
                           rnn_cell=params.get('rnn_cell', 'ugrnn'), ranking=params.get('ranking', 'mlp'),
                           sampler_seed=params.get('sampler_seed', 42), init_seed=params.get('init_seed', 42),
                           process_group=params.get('process_group'), device=params.get('device'))

    state = params.get('clicked_items_state') or clicked_items_state
    if state is None:
        raise RuntimeError('clicked_items_state is not set (nar_trainer_gcom.py:486-489 creates it before the Estimator)')
    hooks = [ItemsStateUpdaterHook(mode, model, eval_metrics_top_n=eval_metrics_top_n, clicked_items_state=state,
                                   eval_sessions_metrics_log=eval_sessions_metrics_log,
                                   content_article_embeddings_matrix=params['content_article_embeddings_matrix'],
                                   articles_metadata=params['articles_metadata'])]
    if mode == ModeKeys.TRAIN:
        def train_op(feats, labs, feed, sync=True):
            return model.train(feats, labs, feed['pop_recent_items_buffer'], feed['articles_recent_pop_norm'], sync=sync)
        return EstimatorSpec(mode, loss=None, train_op=train_op, training_chief_hooks=hooks, model=model)

    # ModeKeys.EVAL (nar_trainer_gcom.py:323-332): loss + eval_metric_ops {hitrate_at_n, mrr_at_n}; each "update op" is
    # one call of model.evaluate, the values are read from the device accumulator at the end
    def eval_update(feats, labs, feed, metrics, step_id=None):
        return model.evaluate(feats, labs, feed['pop_recent_items_buffer'], feed['articles_recent_pop_norm'],
                              metrics=metrics, step_id=step_id)
    return EstimatorSpec(mode, loss=None, eval_metric_ops={'hitrate_at_n': eval_update, 'mrr_at_n': eval_update},
                         evaluation_hooks=hooks, model=model)


class Estimator:
    """tf.estimator.Estimator stand-in: ``train(input_fn, steps=None)`` runs the hook/train_op loop."""

    def __init__(self, model_fn, params, model_dir=None, config=None):
        self.model_fn = model_fn
        self.params = params
        self.model_dir = model_dir
        self._spec: Optional[EstimatorSpec] = None
        self._eval_spec: Optional[EstimatorSpec] = None
        self.last_loss = None
        self.interactions = 0
        self.h2d_bytes_per_step = 0

    def _ensure_spec(self, features, labels) -> EstimatorSpec:
        if self._spec is None:
            self._spec = self.model_fn(features, labels, ModeKeys.TRAIN, self.params)
            latest = ckpt.latest_checkpoint(self.model_dir)          # Estimator semantics: warm-start from model_dir
            if latest is not None:
                ckpt.restore(latest, self._spec.model.engine, self._state())
        return self._spec

    def _state(self) -> Optional[ClickedItemsState]:
        return self.params.get('clicked_items_state') or clicked_items_state

    def save_checkpoint(self, path: Optional[str] = None) -> str:
        """Write weights + TF-Adam slots + global_step + ClickedItemsState (checkpoint.py); default name
        ``<model_dir>/model.ckpt-<global_step>.npz``."""
        eng = self._spec.model.engine
        if path is None:
            if not self.model_dir:
                raise ValueError('no model_dir and no path')
            path = ckpt.checkpoint_path(self.model_dir, eng.global_step)
        return ckpt.save(path, eng, self._state())

    def restore_checkpoint(self, path: str) -> int:
        return ckpt.restore(path, self._spec.model.engine, self._state())

    def train(self, input_fn, steps: Optional[int] = None, hooks=None):
        """hook.before_run -> train_op -> hook.after_run per batch (MonitoredTrainingSession), software-pipelined:
        while the GPU runs step n, the host folds batch n into ClickedItemsState (it depends on the batch's ids only,
        nar_model.py:1635-1650), fetches batch n+1 (tf.data prefetch(1), datasets.py:142) and stages it - H2D copy,
        negative sampling, row lists, normalisation statistics - on a side stream; the loss of step n is read only after
        step n+1 has been queued (two pinned loss slots, one event per step), so the GPU never waits for the host."""
        it = input_fn()

        def fetch():
            try:
                return it.get_next() if hasattr(it, 'get_next') else next(it)
            except (OutOfRangeError, StopIteration):
                return None

        def feed_of(spec):
            feed = {}
            for h in spec.training_chief_hooks:
                feed.update(h.before_run(None))
            return feed

        n = 0
        nxt = fetch() if (steps is None or steps > 0) else None
        if nxt is None:
            return self
        spec = self._ensure_spec(*nxt)
        eng = spec.model.engine
        for h in spec.training_chief_hooks:
            h.begin()
        # Device-resident ClickedItemsState (default; NAR_DEVICE_STATE=0 keeps the hook's host update + per-step upload):
        # the recent-clicks buffer / popularity live in HBM for the duration of train() and are advanced by one kernel per
        # step on the side stream (what hook.after_run does on the host, nar_model.py:1635-1650); the host object is
        # brought up to date when train() returns.
        use_ds = os.environ.get('NAR_DEVICE_STATE', '1') == '1' and self._state() is not None
        prev_st = None
        if use_ds:
            eng.attach_device_state(self._state())
            st_next = eng.stage_ahead_device_state(nxt[0], nxt[1], 'pipe0', None)
        else:
            feed = feed_of(spec)
            st_next = eng.stage_ahead(nxt[0], nxt[1], feed['pop_recent_items_buffer'], feed['articles_recent_pop_norm'], 'pipe0')
        pending = None                                              # (features, labels, out) of the step whose loss is unread

        def finish(p):
            f_, l_, o_ = p
            o_ = eng.result(o_)                                     # waits for that step only (event), reads its loss
            spec.model._publish(f_, l_, o_)
            self.last_loss = o_.get('total_loss')
            self.interactions += int(o_['stage']['L_global'])

        while nxt is not None:
            features, labels = nxt
            out = eng.submit(st_next)                               # step n queued on the main stream
            prev_st = st_next
            self.h2d_bytes_per_step = st_next['h2d_bytes']           # bytes of the one pinned H2D copy of this step
            if not use_ds:
                run_values = {'clicked_items': features['item_clicked'], 'clicked_timestamps': features['event_timestamp'],
                              'last_item_label': labels['label_last_item']}
                for h in spec.training_chief_hooks:
                    h.after_run(None, run_values)                    # host state now describes "before step n+1"
            n += 1
            nxt = fetch() if (steps is None or n < steps) else None
            if nxt is not None:
                # slot (n & 1) was last read by step n-1 (= pending): its event gates the side-stream copy
                after = pending[2]['done'] if pending is not None else None
                if use_ds:
                    st_next = eng.stage_ahead_device_state(nxt[0], nxt[1], 'pipe%d' % (n & 1), prev_st, after=after)
                else:
                    feed = feed_of(spec)
                    st_next = eng.stage_ahead(nxt[0], nxt[1], feed['pop_recent_items_buffer'],
                                              feed['articles_recent_pop_norm'], 'pipe%d' % (n & 1), after=after)
            elif use_ds:
                eng.advance_device_state(prev_st)                    # the last batch of this train() call
            if pending is not None:
                finish(pending)                                      # loss of step n-1: the GPU already runs step n
            pending = (features, labels, out)
        if pending is not None:
            finish(pending)
        if use_ds:
            eng.detach_device_state()                                # host ClickedItemsState = the device state (sync)
        for h in spec.training_chief_hooks:
            h.end()
        if self.model_dir:
            # CheckpointSaverHook at the end of train().  Data parallel: weights, Adam slots and the host state are
            # identical on every rank, so rank 0 alone writes and the others wait for the file to be complete.
            if eng.world > 1:
                import torch
                if eng.rank == 0:
                    self.save_checkpoint()
                torch.distributed.barrier(group=eng.pg)
            else:
                self.save_checkpoint()
        return self

    def evaluate(self, input_fn, steps: Optional[int] = None, hooks=None, name=None) -> dict:
        """tf.estimator.Estimator.evaluate: runs the EVAL graph over ``input_fn`` with the current weights and returns
        ``{'loss', 'hitrate_at_n', 'mrr_at_n', 'global_step'}`` (streaming means over all valid labels, nar_model.py:
        835-885; ``loss`` = mean of the per-batch total_loss like Estimator does).  The hook snapshots ClickedItemsState at
        ``begin`` and restores it at ``end`` (nar_model.py:1415, :1693), and keeps updating it batch by batch in between."""
        import torch
        it = input_fn()

        def fetch():
            try:
                return it.get_next() if hasattr(it, 'get_next') else next(it)
            except (OutOfRangeError, StopIteration):
                return None

        nxt = fetch() if (steps is None or steps > 0) else None
        if nxt is None:
            return {'loss': float('nan'), 'hitrate_at_n': float('nan'), 'mrr_at_n': float('nan'),
                    'global_step': 0 if self._spec is None else self._spec.model.global_step()}
        if self._eval_spec is None:
            self._eval_spec = self.model_fn(nxt[0], nxt[1], ModeKeys.EVAL, self.params)
        spec = self._eval_spec
        if self._spec is not None:
            spec.model.engine.share_params(self._spec.model.engine)        # "restore the latest checkpoint"
        else:
            # fresh Estimator over an existing model_dir: tf.estimator.Estimator.evaluate restores the latest checkpoint
            # (nar_trainer_gcom.py:523) and fails when there is none - never evaluate randomly initialised weights
            latest = ckpt.latest_checkpoint(self.model_dir)
            if latest is None:
                raise ValueError('Estimator.evaluate: no trained model - call train() first or point model_dir at a '
                                 'directory holding a checkpoint (model_dir=%r)' % (self.model_dir,))
            if getattr(self, '_eval_restored', None) != latest:
                ckpt.restore(latest, spec.model.engine, None)               # weights + step; the hook snapshots the state
                self._eval_restored = latest
        for h in spec.evaluation_hooks:
            h.begin()
        metrics = torch.zeros(3, device=spec.model.engine.dev, dtype=torch.float64)
        update = spec.eval_metric_ops['hitrate_at_n']
        n, loss_sum = 0, 0.0
        while nxt is not None:
            features, labels = nxt
            feed = {}
            for h in spec.evaluation_hooks:
                feed.update(h.before_run(None))
            out = update(features, labels, feed, metrics, step_id=n + 1)
            run_values = {'clicked_items': features['item_clicked'], 'clicked_timestamps': features['event_timestamp'],
                          'last_item_label': labels['label_last_item']}
            for h in spec.evaluation_hooks:
                h.after_run(None, run_values)
            loss_sum += out['total_loss']
            n += 1
            nxt = fetch() if (steps is None or n < steps) else None
        for h in spec.evaluation_hooks:
            h.end()
        eng = spec.model.engine
        if eng.world > 1:                                           # data parallel: every rank ranked its own sessions
            torch.distributed.all_reduce(metrics, group=eng.pg)
        m = metrics.cpu().numpy()
        cnt = max(float(m[2]), 1.0)
        return {'loss': loss_sum / max(n, 1), 'hitrate_at_n': float(m[0]) / cnt, 'mrr_at_n': float(m[1]) / cnt,
                'global_step': spec.model.global_step()}

    @property
    def model(self) -> Optional[NARModuleModel]:
        return None if self._spec is None else self._spec.model


def build_estimator(model_dir, content_article_embeddings_matrix, articles_metadata, articles_features_config,
                    session_features_config, hparams, state: ClickedItemsState, **extra) -> Estimator:
    """nar_trainer_gcom.py:335-386 with the flags carried by ``hparams`` (NARHParams)."""
    params = hparams.to_params(session_features_config, articles_features_config, articles_metadata,
                               content_article_embeddings_matrix)
    params['clicked_items_state'] = state
    params.update(extra)
    return Estimator(model_fn=nar_module_model_fn, params=params, model_dir=model_dir)
