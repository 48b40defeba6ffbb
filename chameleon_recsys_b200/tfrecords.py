"""TFRecord ``SequenceExample`` files without TensorFlow.

The reference writes its session files with ``tf.python_io.TFRecordWriter(..., GZIP)`` and one serialized
``tf.train.SequenceExample`` per record (nar_module/nar/tf_records_management.py:12-32; schema
preprocessing/nar_preprocess_gcom.py:75-108) and reads them with ``tf.data.TFRecordDataset(compression_type='GZIP')`` +
``tf.parse_single_sequence_example`` (datasets.py:35-56, :100-143).  This module restates the three public formats
involved; nothing here touches the GPU.

* TFRecord framing (tensorflow/core/lib/io/record_writer.cc): per record
  ``uint64 length | uint32 masked_crc32c(length) | bytes data[length] | uint32 masked_crc32c(data)``, little endian,
  ``masked = ((crc >> 15) | (crc << 17)) + 0xa282ead8`` with CRC-32C (Castagnoli, reflected 0x82F63B78).
  The whole file may be gzip-compressed (detected by the 1f 8b magic).
* protobuf wire format (varint / 64-bit / length-delimited / 32-bit fields; packed and unpacked repeated scalars).
* ``tensorflow/core/example/{example,feature}.proto``:
  ``SequenceExample{1: Features context, 2: FeatureLists feature_lists}``, ``Features{1: map<string, Feature>}``,
  ``FeatureLists{1: map<string, FeatureList>}``, ``FeatureList{1: repeated Feature}``,
  ``Feature{1: BytesList | 2: FloatList | 3: Int64List}``, each list ``{1: repeated value}``.

``read_sequence_examples(paths)`` yields per-session dicts in exactly the shape ``datasets.parse_sequence_example``
consumes: context features as scalars, feature lists as 1-D arrays (one value per step).  The writer half exists for
fixtures, tests and for exporting synthetic sessions in the reference's on-disk format.
"""
from __future__ import annotations

import glob
import gzip
import io
import struct
from typing import Dict, Iterable, Iterator, List, Tuple, Union

import numpy as np

# ----------------------------------------------------------------------------------------------- CRC-32C
_CRC_TABLE = None


def _crc_table() -> np.ndarray:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if (c & 1) else 0)
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------- record framing
class TFRecordError(ValueError):
    pass


def _open(path: str):
    f = open(path, 'rb')
    magic = f.read(2)
    f.seek(0)
    return gzip.GzipFile(fileobj=f) if magic == b'\x1f\x8b' else f


def read_records(path: str, check_crc: bool = True) -> Iterator[bytes]:
    """Yield the payload of every record of one TFRecord file (plain or gzip)."""
    with _open(path) as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) != 12:
                raise TFRecordError('%s: truncated record header' % path)
            (length,), (len_crc,) = struct.unpack('<Q', head[:8]), struct.unpack('<I', head[8:])
            if check_crc and masked_crc32c(head[:8]) != len_crc:
                raise TFRecordError('%s: corrupted record length' % path)
            if length > (1 << 31):
                raise TFRecordError('%s: implausible record length %d' % (path, length))
            data = f.read(length)
            tail = f.read(4)
            if len(data) != length or len(tail) != 4:
                raise TFRecordError('%s: truncated record' % path)
            if check_crc and masked_crc32c(data) != struct.unpack('<I', tail)[0]:
                raise TFRecordError('%s: corrupted record data' % path)
            yield data


def write_records(path: str, records: Iterable[bytes], compress: bool = True):
    f = gzip.open(path, 'wb') if compress else open(path, 'wb')
    with f:
        for data in records:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', masked_crc32c(head)) + data + struct.pack('<I', masked_crc32c(data)))


# ----------------------------------------------------------------------------------------------- protobuf wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7
        if shift > 63:
            raise TFRecordError('varint too long')


def _fields(buf: bytes) -> Iterator[Tuple[int, int, Union[int, bytes]]]:
    """(field number, wire type, value) for every field of one message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
            if len(v) != ln:
                raise TFRecordError('truncated length-delimited field')
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise TFRecordError('unsupported wire type %d' % wt)
        yield field, wt, v


def _signed64(u: int) -> int:
    return u - (1 << 64) if u >= (1 << 63) else u


def _parse_feature(buf: bytes):
    """Feature -> ('bytes', [bytes]) | ('float', float32 array) | ('int64', int64 array)."""
    kind, values = None, None
    for field, wt, v in _fields(buf):
        if field == 1:                                    # BytesList
            kind, values = 'bytes', [b for f2, _, b in _fields(v) if f2 == 1]
        elif field == 2:                                  # FloatList
            out: List[np.ndarray] = []
            for f2, wt2, b in _fields(v):
                if f2 != 1:
                    continue
                out.append(np.frombuffer(b, dtype='<f4') if wt2 in (2, 5) else np.zeros(0, dtype='<f4'))
            kind, values = 'float', (np.concatenate(out).astype(np.float32) if out else np.zeros(0, dtype=np.float32))
        elif field == 3:                                  # Int64List
            vals: List[int] = []
            for f2, wt2, b in _fields(v):
                if f2 != 1:
                    continue
                if wt2 == 0:
                    vals.append(_signed64(b))
                else:                                     # packed varints
                    p = 0
                    while p < len(b):
                        x, p = _varint(b, p)
                        vals.append(_signed64(x))
            kind, values = 'int64', np.asarray(vals, dtype=np.int64)
    return kind, values


def _parse_map(buf: bytes, value_parser) -> Dict[str, object]:
    """Features / FeatureLists: field 1 = repeated map entry {1: key, 2: value}."""
    out = {}
    for field, _, entry in _fields(buf):
        if field != 1:
            continue
        key, val = None, b''
        for f2, _, v in _fields(entry):
            if f2 == 1:
                key = v.decode('utf-8')
            elif f2 == 2:
                val = v
        if key is not None:
            out[key] = value_parser(val)
    return out


def _parse_feature_list(buf: bytes):
    return [_parse_feature(v) for field, _, v in _fields(buf) if field == 1]


def parse_sequence_example_bytes(data: bytes) -> Dict[str, np.ndarray]:
    """One serialized SequenceExample -> {context name: scalar/array, feature-list name: 1-D array (one value per step)}.
    Scalars of a context feature with one value become 0-d arrays (``FixedLenFeature([])``); a feature list must hold
    exactly one value per step (``FixedLenSequenceFeature([])``), like the reference's parser requires."""
    out: Dict[str, np.ndarray] = {}
    for field, _, v in _fields(data):
        if field == 1:                                    # context
            for name, (kind, values) in _parse_map(v, _parse_feature).items():
                if kind == 'bytes':
                    out[name] = np.asarray(values, dtype=object)
                else:
                    out[name] = values.reshape(()) if values.size == 1 else values
        elif field == 2:                                  # feature_lists
            for name, steps in _parse_map(v, _parse_feature_list).items():
                kinds = {k for k, _ in steps}
                if not steps:
                    out[name] = np.zeros(0, dtype=np.int64)
                    continue
                if len(kinds) != 1:
                    raise TFRecordError('feature list %r mixes value types' % name)
                kind = kinds.pop()
                if kind == 'bytes':
                    if any(len(vals) != 1 for _, vals in steps):
                        raise TFRecordError('feature list %r: one value per step expected' % name)
                    out[name] = np.asarray([vals[0] for _, vals in steps], dtype=object)
                else:
                    if any(vals.size != 1 for _, vals in steps):
                        raise TFRecordError('feature list %r: one value per step expected' % name)
                    out[name] = np.concatenate([vals for _, vals in steps])
    return out


def expand_files(files: Union[str, Iterable[str]]) -> List[str]:
    """A path, a glob pattern or a list of either -> sorted list of files (the reference passes chunked file lists,
    nar_trainer_gcom.py:496-509, and resolves patterns with tf.gfile / glob)."""
    pats = [files] if isinstance(files, str) else list(files)
    out: List[str] = []
    for p in pats:
        hits = sorted(glob.glob(p)) if any(ch in p for ch in '*?[') else [p]
        out.extend(hits)
    return out


def read_sequence_examples(files: Union[str, Iterable[str]], check_crc: bool = True) -> Iterator[Dict[str, np.ndarray]]:
    for path in expand_files(files):
        for rec in read_records(path, check_crc=check_crc):
            yield parse_sequence_example_bytes(rec)


# ----------------------------------------------------------------------------------------------- writer (fixtures / export)
def _enc_varint(x: int) -> bytes:
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_ld(field: int, payload: bytes) -> bytes:
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def _enc_feature(values, kind: str) -> bytes:
    if kind == 'int64':
        packed = b''.join(_enc_varint(int(v)) for v in values)
        return _enc_ld(3, _enc_ld(1, packed))
    if kind == 'float':
        return _enc_ld(2, _enc_ld(1, np.asarray(values, dtype='<f4').tobytes()))
    return _enc_ld(1, b''.join(_enc_ld(1, (v if isinstance(v, bytes) else str(v).encode())) for v in values))


def _kind_of(arr) -> str:
    a = np.asarray(arr)
    if a.dtype.kind in 'iub':
        return 'int64'
    if a.dtype.kind == 'f':
        return 'float'
    return 'bytes'


def encode_sequence_example(context: Dict[str, object], feature_lists: Dict[str, object]) -> bytes:
    """Inverse of ``parse_sequence_example_bytes`` (same layout as make_sequence_example, nar_preprocess_gcom.py:75-108:
    one-value context features, one Feature per step in every feature list)."""
    ctx = b''
    for name in sorted(context):
        vals = np.atleast_1d(np.asarray(context[name]))
        entry = _enc_ld(1, name.encode()) + _enc_ld(2, _enc_feature(vals, _kind_of(vals)))
        ctx += _enc_ld(1, entry)
    fl = b''
    for name in sorted(feature_lists):
        vals = np.asarray(feature_lists[name])
        kind = _kind_of(vals)
        steps = b''.join(_enc_ld(1, _enc_feature([v], kind)) for v in vals)
        fl += _enc_ld(1, _enc_ld(1, name.encode()) + _enc_ld(2, steps))
    return _enc_ld(1, ctx) + _enc_ld(2, fl)


def write_sequence_examples(path: str, sessions: Iterable[Dict[str, np.ndarray]], features_config: dict, compress: bool = True):
    """Export decoded session dicts (e.g. synthetic.SessionStream) as a reference-format TFRecord file."""
    def rows():
        for s in sessions:
            ctx = {k: s[k] for k in features_config['single_features']}
            fl = {k: s[k] for k in features_config['sequence_features']}
            yield encode_sequence_example(ctx, fl)
    write_records(path, rows(), compress=compress)


def _selftest():  # pragma: no cover
    buf = io.BytesIO()
    assert crc32c(b'123456789') == 0xE3069283
    return buf


if __name__ == '__main__':  # pragma: no cover
    _selftest()
    print('ok')
