"""ACR-module resource loaders of the NAR trainer, without TensorFlow / GCS.

* G1 (nar_trainer_gcom.py:131-139, :142-150, :466-474): ``articles_metadata.csv`` + a pickled ``[V, E]`` ndarray of
  article content embeddings; the trainer l2-normalises every row and multiplies by ``content_embedding_scale_factor``.
* Adressa (nar_utils.py:9-17): one pickle holding ``(acr_label_encoders, articles_metadata_df, content_article_embeddings)``.

The pickles are the reference's own artefacts (``utils.serialize`` = ``pickle.dump``); loading a pickle executes code,
so only open files you produced or trust - same caveat as the reference.
"""
from __future__ import annotations

import pickle
from typing import Dict, Tuple

import numpy as np


def deserialize(path: str):
    """utils.py deserialize(): plain pickle."""
    with open(path, 'rb') as f:
        return pickle.load(f)


def load_acr_module_resources(articles_metadata_csv_path: str, articles_content_embeddings_pickle_path: str):
    """G1 form (nar_trainer_gcom.py:131-139) -> (articles_metadata_df, content_article_embeddings [V, E])."""
    import pandas as pd
    content_article_embeddings = np.asarray(deserialize(articles_content_embeddings_pickle_path))
    articles_metadata_df = pd.read_csv(articles_metadata_csv_path)
    return articles_metadata_df, content_article_embeddings


def load_acr_module_resources_adressa(acr_module_resources_path: str):
    """Adressa form (nar_utils.py:9-17) -> (acr_label_encoders, articles_metadata_df, content_article_embeddings)."""
    acr_label_encoders, articles_metadata_df, content_article_embeddings = deserialize(acr_module_resources_path)
    return acr_label_encoders, articles_metadata_df, np.asarray(content_article_embeddings)


def process_articles_metadata(articles_metadata_df, articles_features_config: dict) -> Dict[str, np.ndarray]:
    """nar_trainer_gcom.py:142-150: one ``[V]`` array per configured article feature, row i = article id i
    (the CSV is sorted by ``article_id``; id 0 is the padding article)."""
    out = {}
    for name in articles_features_config:
        out[name] = articles_metadata_df[name].values
        if out[name].dtype.kind in 'iu':
            out[name] = out[name].astype(np.int64)
    return out


def normalize_content_embeddings(content_article_embeddings: np.ndarray, content_embedding_scale_factor: float) -> np.ndarray:
    """nar_trainer_gcom.py:469-474: sklearn ``Normalizer(norm='l2')`` per row (all-zero rows stay zero), then scale."""
    m = np.asarray(content_article_embeddings, dtype=np.float32)
    norms = np.sqrt((m.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    norms[norms == 0.0] = 1.0
    return (m / norms).astype(np.float32) * np.float32(content_embedding_scale_factor)
