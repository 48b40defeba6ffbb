"""Generate tests/golden/dataset_golden.npz + tests/golden/sessions_golden.tfrecord.gz: batches produced by the REFERENCE
input pipeline (/root/reference/nar_module/nar/datasets.py: make_dataset = TFRecordDataset -> parse_sequence_example ->
padded_batch -> deflate_and_split_features_label, imported unmodified) running on the tf.data subset of the TF-API
stand-in tests/golden/tf1_shim.py.  The protobuf decoding inside tf.parse_single_sequence_example is done by Google's
protobuf runtime on the tensorflow/core/example .proto schema (tests/test_tfrecords.py::_example_classes), NOT by this
repo's decoder.  The TFRecord file is written by chameleon_recsys_b200.tfrecords (itself checked against the protobuf
runtime in tests/test_tfrecords.py) from synthetic sessions; it is input data of the test.
Run once in the build container; both files are committed."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
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
ref_ds = importlib.import_module('refnar.datasets')

from chameleon_recsys_b200 import tfrecords as tfr  # noqa: E402
from chameleon_recsys_b200.harness import make_problem  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_tfrecords import _example_classes  # noqa: E402

SequenceExample = _example_classes()


def decode(data):
    ex = SequenceExample()
    ex.ParseFromString(bytes(data))

    def vals(feature):
        kind = feature.WhichOneof('kind')
        return list(getattr(feature, kind).value)
    ctx = {k: vals(v) for k, v in ex.context.feature.items()}
    seqs = {k: [x for f in fl.feature for x in vals(f)] for k, fl in ex.feature_lists.feature_list.items()}
    return ctx, seqs


shim.configure(float64=False)
shim.S.example_decoder = decode
pb = make_problem('tiny', profile='B')
cfg = pb.session_features_config
src = iter(pb.stream)
sessions = [next(src) for _ in range(150)]
path = os.path.join(HERE, 'sessions_golden.tfrecord.gz')
tfr.write_sequence_examples(path, sessions, cfg)
out = {}
for ci, (batch_size, trunc) in enumerate([(64, 5), (32, 3), (150, 20)]):
    ds = ref_ds.make_dataset(path, cfg, batch_size=batch_size, truncate_sequence_length=trunc)
    nb = 0
    for feats, labels in ds:
        for k, v in feats.items():
            out['c%d_b%d_feat/%s' % (ci, nb, k)] = v.numpy()
        for k, v in labels.items():
            out['c%d_b%d_label/%s' % (ci, nb, k)] = v.numpy()
        nb += 1
    out['c%d_cfg' % ci] = np.array([batch_size, trunc, nb])
np.savez_compressed(os.path.join(HERE, 'dataset_golden.npz'), **out)
print('wrote', len(out), 'arrays;', os.path.getsize(path) // 1024, 'KB tfrecord,',
      os.path.getsize(os.path.join(HERE, 'dataset_golden.npz')) // 1024, 'KB npz')
