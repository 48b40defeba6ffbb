"""TFRecord / SequenceExample reader (chameleon_recsys_b200/tfrecords.py) against published known answers and against
Google's protobuf runtime as an independent encoder / decoder of the same .proto schema."""
import gzip
import os
import struct

import numpy as np
import pytest

from chameleon_recsys_b200 import tfrecords as tfr


def test_crc32c_rfc3720_vectors():
    assert tfr.crc32c(b'123456789') == 0xE3069283
    assert tfr.crc32c(b'\x00' * 32) == 0x8A9136AA
    assert tfr.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tfr.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfr.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    c = tfr.crc32c(b'foo')
    assert tfr.masked_crc32c(b'foo') == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _example_classes():
    """tensorflow/core/example/{feature,example}.proto rebuilt with the protobuf runtime (no TF import)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='nar_test_example.proto', package='nartest', syntax='proto3')
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add(); m.name = name; return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        f = m.field.add(); f.name = name; f.number = num; f.type = typ; f.label = label
        if type_name:
            f.type_name = '.nartest.' + type_name
        if packed is not None:
            f.options.packed = packed
        if oneof is not None:
            f.oneof_index = oneof
        return f

    m = msg('BytesList'); field(m, 'value', 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    m = msg('FloatList'); field(m, 'value', 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    m = msg('Int64List'); field(m, 'value', 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = msg('Feature'); m.oneof_decl.add().name = 'kind'
    field(m, 'bytes_list', 1, F.TYPE_MESSAGE, type_name='BytesList', oneof=0)
    field(m, 'float_list', 2, F.TYPE_MESSAGE, type_name='FloatList', oneof=0)
    field(m, 'int64_list', 3, F.TYPE_MESSAGE, type_name='Int64List', oneof=0)
    m = msg('FeatureList'); field(m, 'feature', 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name='Feature')
    for outer, val in (('Features', 'Feature'), ('FeatureLists', 'FeatureList')):
        m = msg(outer)
        e = m.nested_type.add(); e.name = 'MapEntry'; e.options.map_entry = True
        k = e.field.add(); k.name = 'key'; k.number = 1; k.type = F.TYPE_STRING; k.label = F.LABEL_OPTIONAL
        v = e.field.add(); v.name = 'value'; v.number = 2; v.type = F.TYPE_MESSAGE; v.label = F.LABEL_OPTIONAL
        v.type_name = '.nartest.' + val
        f = m.field.add(); f.name = 'feature' if outer == 'Features' else 'feature_list'; f.number = 1
        f.type = F.TYPE_MESSAGE; f.label = F.LABEL_REPEATED; f.type_name = '.nartest.%s.MapEntry' % outer
    m = msg('SequenceExample')
    field(m, 'context', 1, F.TYPE_MESSAGE, type_name='Features')
    field(m, 'feature_lists', 2, F.TYPE_MESSAGE, type_name='FeatureLists')
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:
        fac = message_factory.MessageFactory(pool)
        return fac.GetPrototype(pool.FindMessageTypeByName('nartest.SequenceExample'))
    return get(pool.FindMessageTypeByName('nartest.SequenceExample'))


def _session(rs, n):
    return ({'user_id': int(rs.randint(1, 1 << 40)), 'session_id': int(rs.randint(1, 1 << 50)),
             'session_start': 1506826800000 + int(rs.randint(0, 10 ** 9)), 'session_size': n},
            {'event_timestamp': 1506826800000 + np.cumsum(rs.randint(1, 10 ** 5, n)).astype(np.int64),
             'item_clicked': rs.randint(1, 46000, n).astype(np.int64), 'environment': rs.randint(0, 5, n).astype(np.int64),
             'local_hour_sin': rs.uniform(-1, 1, n).astype(np.float32), 'local_weekday': rs.uniform(0, 1, n).astype(np.float32)})


def test_parser_matches_google_protobuf_encoder_and_decoder():
    SE = _example_classes()
    rs = np.random.RandomState(0)
    for n in (1, 2, 7, 20):
        ctx, fl = _session(rs, n)
        # (1) Google's encoder -> our parser (unpacked-per-step layout exactly like make_sequential_feature)
        m = SE()
        for k, v in ctx.items():
            m.context.feature[k].int64_list.value.append(v)
        m.context.feature['neg'].int64_list.value.append(-5)              # negative int64 = 10-byte varint
        for k, arr in fl.items():
            for v in arr:
                f = m.feature_lists.feature_list[k].feature.add()
                if arr.dtype == np.float32:
                    f.float_list.value.append(float(v))
                else:
                    f.int64_list.value.append(int(v))
        got = tfr.parse_sequence_example_bytes(m.SerializeToString())
        assert int(got['neg']) == -5
        for k, v in ctx.items():
            assert got[k].shape == () and int(got[k]) == v
        for k, arr in fl.items():
            assert got[k].dtype == arr.dtype and np.array_equal(got[k], arr), k
        # (2) our encoder -> Google's decoder
        m2 = SE()
        m2.ParseFromString(tfr.encode_sequence_example(ctx, fl))
        for k, v in ctx.items():
            assert list(m2.context.feature[k].int64_list.value) == [v]
        for k, arr in fl.items():
            feats = m2.feature_lists.feature_list[k].feature
            vals = [(f.float_list.value[0] if arr.dtype == np.float32 else f.int64_list.value[0]) for f in feats]
            assert np.array_equal(np.asarray(vals, dtype=arr.dtype), arr)


def test_record_framing_gzip_and_corruption(tmp_path):
    recs = [b'', b'a', os.urandom(1000), b'x' * 70000]
    for compress in (True, False):
        p = str(tmp_path / ('r%d.tfrecord' % compress))
        tfr.write_records(p, recs, compress=compress)
        assert (open(p, 'rb').read(2) == b'\x1f\x8b') == compress
        assert list(tfr.read_records(p)) == recs
    # layout of the first record of the plain file: length | masked crc | data | masked crc
    raw = open(str(tmp_path / 'r0.tfrecord'), 'rb').read()
    assert struct.unpack('<Q', raw[:8])[0] == 0 and struct.unpack('<I', raw[8:12])[0] == tfr.masked_crc32c(raw[:8])
    bad = bytearray(raw); bad[100] ^= 1                                  # inside the payload of the third record
    p = str(tmp_path / 'bad.tfrecord'); open(p, 'wb').write(bytes(bad))
    with pytest.raises(tfr.TFRecordError):
        list(tfr.read_records(p))
    assert len(list(tfr.read_records(p, check_crc=False))) == len(recs)
    bad = bytearray(raw); bad[40] ^= 0x80                                 # inside the LENGTH field of the third record
    p = str(tmp_path / 'badlen.tfrecord'); open(p, 'wb').write(bytes(bad))
    for chk in (True, False):
        with pytest.raises(tfr.TFRecordError):
            list(tfr.read_records(p, check_crc=chk))
    p = str(tmp_path / 'trunc.tfrecord'); open(p, 'wb').write(raw[:-3])
    with pytest.raises(tfr.TFRecordError):
        list(tfr.read_records(p))


def test_input_fn_from_tfrecord_files_equals_in_memory(tmp_path):
    """prepare_dataset_iterator(<file pattern>) == prepare_dataset_iterator(<decoded sessions>) batch for batch."""
    from chameleon_recsys_b200.datasets import prepare_dataset_iterator
    from chameleon_recsys_b200.harness import make_problem
    pb = make_problem('tiny', profile='B')
    src = iter(pb.stream)
    sessions = [next(src) for _ in range(150)]
    cfg = pb.session_features_config
    half = len(sessions) // 2
    tfr.write_sequence_examples(str(tmp_path / 'sessions_0000.tfrecord.gz'), sessions[:half], cfg)
    tfr.write_sequence_examples(str(tmp_path / 'sessions_0001.tfrecord.gz'), sessions[half:], cfg)
    a = prepare_dataset_iterator(str(tmp_path / 'sessions_*.tfrecord.gz'), cfg, batch_size=64, truncate_session_length=5)
    b = prepare_dataset_iterator(iter(sessions), cfg, batch_size=64, truncate_session_length=5)
    n = 0
    for (fa, la), (fb, lb) in zip(a, b):
        assert fa.keys() == fb.keys() and la.keys() == lb.keys()
        for k in fa:
            assert fa[k].dtype == fb[k].dtype and np.array_equal(fa[k], fb[k]), k
        for k in la:
            assert np.array_equal(la[k], lb[k]), k
        n += 1
    assert n == 3


def test_input_pipeline_matches_reference_pipeline():
    """prepare_dataset_iterator on a TFRecord file against the batches the REFERENCE pipeline produced from the same file
    (datasets.py make_dataset imported unmodified, run on the tf.data subset of tests/golden/tf1_shim.py with Google's
    protobuf runtime as the example decoder; tests/golden/make_dataset_golden.py): keys, dtypes, shapes, padding, the
    label shift, truncation and the last partial batch."""
    from chameleon_recsys_b200.datasets import prepare_dataset_iterator
    from chameleon_recsys_b200.harness import make_problem
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    d = np.load(os.path.join(here, 'dataset_golden.npz'))
    cfg = make_problem('tiny', profile='B').session_features_config
    for ci in range(3):
        batch_size, trunc, nb = (int(v) for v in d['c%d_cfg' % ci])
        it = prepare_dataset_iterator(os.path.join(here, 'sessions_golden.tfrecord.gz'), cfg, batch_size=batch_size,
                                      truncate_session_length=trunc)
        n = 0
        for feats, labels in it:
            ref_f = {k.split('/', 1)[1]: d[k] for k in d.files if k.startswith('c%d_b%d_feat/' % (ci, n))}
            ref_l = {k.split('/', 1)[1]: d[k] for k in d.files if k.startswith('c%d_b%d_label/' % (ci, n))}
            assert set(feats) == set(ref_f) and set(labels) == set(ref_l)
            for k in ref_f:
                assert feats[k].dtype == ref_f[k].dtype and feats[k].shape == ref_f[k].shape, (k, feats[k].shape, ref_f[k].shape)
                assert np.array_equal(feats[k], ref_f[k]), k
            for k in ref_l:
                assert labels[k].dtype == ref_l[k].dtype and np.array_equal(labels[k], ref_l[k]), k
            n += 1
        assert n == nb
