"""``NARModuleModel`` and ``ItemsStateUpdaterHook`` - the reference's model-side API
(nar_module/nar/nar_model.py:100-129 ctor, :1370-1470 / :1504-1511 / :1635-1650 hook) on top of
the B200 engine.

TF builds a symbolic graph per ``model_fn`` call and the hook feeds placeholders per step; here
the object is built once (weights + ACR table resident in HBM) and ``run(features, labels)``
plays the role of one ``session.run(train_op)``.  Attribute names the hook fetches in the
reference (``item_clicked``, ``event_timestamp``, ``next_item_label``, ``label_last_item``,
``session_id``, ``user_id``, ``batch_negative_items``, ``total_loss``, ``train``) are kept.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .clicked_items_state import ClickedItemsState, batch_clicks_for_state_update
from .hparams import ModeKeys
from .plan import FeaturePlan, ParamLayout


class NARModuleModel:

    def __init__(self, mode, inputs, labels,
                 session_features_config,
                 articles_features_config,
                 batch_size,
                 lr, keep_prob, negative_samples, negative_sample_from_buffer,
                 content_article_embeddings_matrix,
                 rnn_num_layers=1,
                 softmax_temperature=1.0,
                 reg_weight_decay=0.0,
                 recent_clicks_buffer_hours=1.0,
                 recent_clicks_buffer_max_size=1000,
                 recent_clicks_for_normalization=1000,
                 articles_metadata=None,
                 plot_histograms=False,
                 metrics_top_n=5,
                 elapsed_days_smooth_log_base=1.3,
                 popularity_smooth_log_base=2.0,
                 CAR_embedding_size=256,
                 rnn_units=256,
                 max_cardinality_for_ohe=10,
                 novelty_reg_factor=0.0,
                 diversity_reg_factor=0.0,
                 internal_features_config={'recency': True,
                                           'novelty': True,
                                           'article_content_embeddings': True,
                                           'item_clicked_embeddings': True},
                 eval_cold_start=False,
                 # --- extensions (not in the reference signature) ---
                 rnn_cell='ugrnn', ranking='mlp', sampler_seed=42, init_seed=42, device=None,
                 process_group=None, fwd_precision=None, bwd_precision=1):
        from .engine import NarEngine          # imports torch + the CUDA library; fails loudly without them
        self.mode = mode
        self.lr = lr
        self.keep_prob = keep_prob
        self.is_training = (mode == ModeKeys.TRAIN)
        self.negative_samples = negative_samples
        self.negative_sample_from_buffer = negative_sample_from_buffer
        self.rnn_num_layers = rnn_num_layers
        self.metrics_top_n = metrics_top_n
        self.reg_weight_decay = reg_weight_decay
        self.batch_size = batch_size
        self.session_features_config = session_features_config
        self.articles_features_config = articles_features_config
        self.internal_features_config = internal_features_config
        self.items_vocab_size = articles_features_config['article_id']['cardinality']
        self.content_article_embeddings_matrix = content_article_embeddings_matrix
        self.articles_metadata = articles_metadata
        self.plan = FeaturePlan(session_features_config, articles_features_config, internal_features_config,
                                max_cardinality_for_ohe, content_article_embeddings_matrix.shape[1],
                                self.items_vocab_size)
        self.layout = ParamLayout(self.plan, CAR_embedding_size, rnn_units, rnn_num_layers, rnn_cell=rnn_cell)
        self.engine = NarEngine(self.plan, self.layout, content_article_embeddings_matrix, articles_metadata,
                                negative_samples=negative_samples,
                                negative_sample_from_buffer=negative_sample_from_buffer,
                                softmax_temperature=softmax_temperature, reg_weight_decay=reg_weight_decay, lr=lr,
                                recent_clicks_buffer_max_size=recent_clicks_buffer_max_size,
                                recent_clicks_for_normalization=recent_clicks_for_normalization,
                                elapsed_days_smooth_log_base=elapsed_days_smooth_log_base,
                                popularity_smooth_log_base=popularity_smooth_log_base, ranking=ranking,
                                rnn_cell=rnn_cell, sampler_seed=sampler_seed, device=device,
                                process_group=process_group, fwd_precision=fwd_precision,
                                bwd_precision=bwd_precision,
                                keep_prob=keep_prob if mode == ModeKeys.TRAIN else 1.0,
                                novelty_reg_factor=novelty_reg_factor)
        self.engine.set_params(self.layout.init_logical(init_seed))
        # fetch targets of the reference hook (numpy after each run)
        self.item_clicked = None
        self.event_timestamp = None
        self.next_item_label = None
        self.label_last_item = None
        self.session_id = None
        self.user_id = None
        self.batch_negative_items = None
        self.total_loss = None
        self.predicted_item_ids = None
        self.predicted_item_probs = None
        self._features = inputs
        self._labels = labels
        self._last = None

    # ``train`` is the train_op: call it with the hook's feed (state arrays) to run one step
    def train(self, features: Dict[str, np.ndarray], labels: Dict[str, np.ndarray], pop_recent_items_buffer: np.ndarray,
              articles_recent_pop_norm: np.ndarray, sync: bool = True) -> dict:
        out = self.engine.train_step(features, labels, pop_recent_items_buffer, articles_recent_pop_norm, sync=sync)
        self._publish(features, labels, out)
        return out

    def evaluate(self, features: Dict[str, np.ndarray], labels: Dict[str, np.ndarray], pop_recent_items_buffer: np.ndarray,
                 articles_recent_pop_norm: np.ndarray, metrics=None, step_id=None) -> dict:
        """One EVAL batch (the eval_metric_ops update): loss, ``predicted_item_ids`` / ``predicted_item_probs``
        (nar_model.py:520-524) and the HR@n / MRR@n accumulators (:835-885) for ``metrics_top_n``."""
        out = self.engine.eval_step(features, labels, pop_recent_items_buffer, articles_recent_pop_norm,
                                    top_n=self.metrics_top_n, metrics=metrics, step_id=step_id)
        self._publish(features, labels, out)
        self.predicted_item_ids = out.get('predicted_item_ids')
        self.predicted_item_probs = out.get('predicted_item_probs')
        return out

    def _publish(self, features, labels, out):
        self.item_clicked = features['item_clicked']
        self.event_timestamp = features['event_timestamp'][..., None]
        self.next_item_label = labels['label_next_item']
        self.label_last_item = labels['label_last_item']
        self.session_id = features.get('session_id')
        self.user_id = features.get('user_id')
        self.batch_negative_items = out['negatives']        # device tensor [B,T,K] int64
        self.total_loss = out.get('total_loss')
        self._last = out

    def global_step(self) -> int:
        return self.engine.global_step


class ItemsStateUpdaterHook:
    """Train-mode parts of the reference SessionRunHook (nar_model.py:1370-1470, :1635-1650):
    ``before_run`` hands the per-step host state to the graph (the 46 MB ACR matrix is NOT re-fed:
    it lives in HBM), ``after_run`` folds the batch's clicks back into ``ClickedItemsState``."""

    def __init__(self, mode, model: NARModuleModel, eval_metrics_top_n, clicked_items_state: ClickedItemsState,
                 eval_sessions_metrics_log=None, sessions_negative_items_log=None,
                 sessions_chameleon_recommendations_log=None, content_article_embeddings_matrix=None,
                 articles_metadata=None, eval_negative_sample_relevance=None, eval_benchmark_classifiers=(),
                 eval_metrics_by_session_position=False, eval_cold_start=False):
        self.mode = mode
        self.model = model
        self.eval_metrics_top_n = eval_metrics_top_n
        self.clicked_items_state = clicked_items_state
        self.eval_sessions_metrics_log = eval_sessions_metrics_log
        if eval_benchmark_classifiers:
            raise NotImplementedError('benchmark recommenders are out of scope (SURVEY.md section 2, rows 9-11)')

    def begin(self):
        if self.mode == ModeKeys.EVAL:
            self.clicked_items_state.save_state_checkpoint()        # nar_model.py:1415

    def before_run(self, run_context=None) -> dict:
        """-> feed dict (nar_model.py:1458-1467)."""
        return {'articles_recent_pop_norm': self.clicked_items_state.get_articles_recent_pop_norm(),
                'pop_recent_items_buffer': self.clicked_items_state.get_recent_clicks_buffer()}

    def after_run(self, run_context, run_values: dict):
        """run_values: {'clicked_items','clicked_timestamps','last_item_label'} (nar_model.py:1505-1508)."""
        self.clicked_items_state.update_from_batch(run_values['clicked_items'], run_values['clicked_timestamps'],
                                                   run_values['last_item_label'])

    def end(self, session=None):
        if self.mode == ModeKeys.EVAL:
            self.clicked_items_state.restore_state_checkpoint()     # nar_model.py:1693
