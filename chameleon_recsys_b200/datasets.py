"""``input_fn`` of the NAR Estimator: sessions -> padded (features, labels) batches.

Mirror of nar_module/nar/datasets.py: ``parse_sequence_example`` (:35-82) truncates to
``truncate_session_length``, builds ``label_next_item = item_clicked[1:]`` and
``label_last_item = item_clicked[-1:]`` and drops the last element of every sequence
feature; ``make_dataset`` (:100-143) zero-pads to the batch maximum (``padded_batch``),
no shuffling; ``deflate_and_split_features_label`` (:84-97) splits features / labels;
dtypes follow utils.get_tf_dtype (:59-68): int -> int64, float -> float32.

The source is either TFRecord files of ``SequenceExample`` protos in the reference's on-disk schema
(nar_preprocess_gcom.py:75-108; decoded by ``tfrecords.py`` - framing, CRC-32C, gzip and the protobuf wire format,
no TensorFlow) or any iterable of already-decoded per-session dicts, e.g. ``synthetic.SessionStream``.
"""
from __future__ import annotations

import itertools
from typing import Dict, Iterable, Iterator, Tuple

import numpy as np


class OutOfRangeError(StopIteration):
    """Raised at the end of a one-shot iterator (tf.errors.OutOfRangeError stand-in)."""


def get_np_dtype(dtype: str):
    """utils.py:59-68."""
    if dtype == 'int':
        return np.int64
    if dtype == 'float':
        return np.float32
    raise Exception('Dtype not supported: {}'.format(dtype))


def parse_sequence_example(example: Dict[str, np.ndarray], features_config: dict,
                           truncate_sequence_length: int = 20) -> Dict[str, np.ndarray]:
    """datasets.py:35-82 on an already-decoded example."""
    out: Dict[str, np.ndarray] = {}
    for name, fc in features_config['single_features'].items():
        out[name] = np.asarray(example[name], dtype=get_np_dtype(fc['dtype'])).reshape(())
    out['session_size'] = np.minimum(out['session_size'], truncate_sequence_length).astype(np.int64)
    seq = {}
    for name, fc in features_config['sequence_features'].items():
        seq[name] = np.asarray(example[name], dtype=get_np_dtype(fc['dtype']))[:truncate_sequence_length]
    out['label_next_item'] = seq['item_clicked'][1:]
    out['label_last_item'] = seq['item_clicked'][-1:]
    for name in seq:
        out[name] = seq[name][:-1]
    return out


def _pad_stack(rows, dtype) -> np.ndarray:
    n = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), n), dtype=dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def make_batch(parsed, features_config: dict) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """padded_batch + deflate_and_split_features_label (datasets.py:84-97, :134-140)."""
    feats: Dict[str, np.ndarray] = {}
    for name, fc in features_config['single_features'].items():
        feats[name] = np.asarray([p[name] for p in parsed], dtype=get_np_dtype(fc['dtype']))
    for name, fc in features_config['sequence_features'].items():
        feats[name] = _pad_stack([p[name] for p in parsed], get_np_dtype(fc['dtype']))
    labels = {
        'label_next_item': _pad_stack([p['label_next_item'] for p in parsed], np.int64),
        'label_last_item': _pad_stack([p['label_last_item'] for p in parsed], np.int64),
    }
    return feats, labels


def make_dataset(source: Iterable[Dict[str, np.ndarray]], features_config: dict, batch_size: int = 128,
                 truncate_sequence_length: int = 20) -> Iterator:
    it = iter(source)
    while True:
        chunk = list(itertools.islice(it, batch_size))
        if not chunk:
            return
        parsed = [parse_sequence_example(e, features_config, truncate_sequence_length) for e in chunk]
        yield make_batch(parsed, features_config)


class OneShotIterator:
    def __init__(self, gen):
        self._gen = gen

    def get_next(self):
        try:
            return next(self._gen)
        except StopIteration:
            raise OutOfRangeError()

    def __iter__(self):
        return self

    def __next__(self):
        return self.get_next()


def prepare_dataset_iterator(files, features_config, batch_size=128, truncate_session_length=20):
    """datasets.py:166-179.  ``files``: a TFRecord path / glob pattern / list of paths like the reference takes
    (nar_trainer_gcom.py:511-516), or an iterable of decoded session dicts (see module doc).
    Returns a one-shot iterator whose ``get_next()`` yields ``(features, labels)``."""
    if isinstance(files, str) or (isinstance(files, (list, tuple)) and files and all(isinstance(f, str) for f in files)):
        from .tfrecords import read_sequence_examples
        files = read_sequence_examples(files)
    return OneShotIterator(make_dataset(files, features_config, batch_size=batch_size,
                                        truncate_sequence_length=truncate_session_length))
