"""TFRecord files and tf.train.Example messages without TensorFlow.

The reference reads its datasets through tf.data.TFRecordDataset + tf.parse_single_example
(video_prediction/datasets/base_dataset.py:129-151, 314-353, 415-453).  Both formats are small and public:

* a TFRecord file is a sequence of  [length u64 LE][masked crc32c(length) u32][data][masked crc32c(data) u32];
  masked crc = rotate_right(crc, 15) + 0xa282ead8;
* tf.train.Example is the protobuf  Example{1: Features{1: map<string, Feature>}} with
  Feature{1: BytesList{1: repeated bytes} | 2: FloatList{1: repeated float, packed} | 3: Int64List{1: repeated int64, packed}}.

The writer half exists for the tests (round trips, fixture files) and for converting data into the format."""
from __future__ import annotations

import struct

import numpy as np

_CRC_TABLE = None


def _table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78                                  # CRC-32C (Castagnoli), reflected
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = [int(v) for v in tab]
    return _CRC_TABLE


def crc32c(data):
    tab = _table()
    c = 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


class TFRecordError(IOError):
    pass


def read_records(path, verify_data_crc=False):
    """Yields the raw records of a TFRecord file.  The 12-byte header's CRC is always checked (it is what tells a TFRecord
    from garbage); the data CRC only on request (pure-Python CRC over megabytes of frames is slow)."""
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise TFRecordError('%s: truncated record header' % path)
            length, = struct.unpack('<Q', head[:8])
            if struct.unpack('<I', head[8:])[0] != masked_crc32c(head[:8]):
                raise TFRecordError('%s: corrupted record header (length crc)' % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise TFRecordError('%s: truncated record' % path)
            if verify_data_crc and struct.unpack('<I', tail)[0] != masked_crc32c(data):
                raise TFRecordError('%s: corrupted record (data crc)' % path)
            yield data


def write_records(path, records):
    with open(path, 'wb') as f:
        for data in records:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', masked_crc32c(head)) + data + struct.pack('<I', masked_crc32c(data)))


# ------------------------------------------------------------------ protobuf wire format (the subset tf.train.Example uses)
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """(field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield num, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_feature(buf):
    for num, wt, val in _fields(buf):
        if num == 1:                                        # BytesList
            return 'bytes', [bytes(v) for n2, _, v in _fields(val) if n2 == 1]
        if num == 2:                                        # FloatList: packed (wire type 2) or one fixed32 per element
            parts = []
            for n2, w2, v in _fields(val):
                if n2 == 1:
                    parts.append(np.frombuffer(bytes(v), dtype='<f4'))
            return 'float', (np.concatenate(parts) if parts else np.zeros(0, np.float32)).astype(np.float32)
        if num == 3:                                        # Int64List: packed varints or one varint per element
            vals = []
            for n2, w2, v in _fields(val):
                if n2 != 1:
                    continue
                if w2 == 0:
                    vals.append(_signed64(v))
                else:
                    p = 0
                    while p < len(v):
                        x, p = _varint(v, p)
                        vals.append(_signed64(x))
            return 'int64', np.asarray(vals, dtype=np.int64)
    return 'bytes', []                                      # an empty Feature


def parse_example(data):
    """tf.train.Example bytes -> {feature name: ('bytes', [bytes, ...]) | ('float', float32 array) | ('int64', int64 array)}."""
    buf = memoryview(data)
    out = {}
    for num, _, features in _fields(buf):
        if num != 1:
            continue
        for n2, _, entry in _fields(features):              # map<string, Feature> entries
            if n2 != 1:
                continue
            key, feat = None, None
            for n3, _, v in _fields(entry):
                if n3 == 1:
                    key = bytes(v).decode('utf-8')
                elif n3 == 2:
                    feat = _parse_feature(v)
            if key is not None:
                out[key] = feat if feat is not None else ('bytes', [])
    return out


def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def make_example(features):
    """{name: bytes | [bytes, ...] | float array | int array} -> serialized tf.train.Example (floats / ints packed, as TF writes them)."""
    entries = b''
    for name in sorted(features):
        v = features[name]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, (list, tuple)) and (not v or isinstance(v[0], (bytes, bytearray))):
            feat = _ld(1, b''.join(_ld(1, bytes(b)) for b in v))
        else:
            arr = np.asarray(v)
            if arr.dtype.kind == 'f':
                feat = _ld(2, _ld(1, arr.astype('<f4').tobytes()))
            else:
                feat = _ld(3, _ld(1, b''.join(_enc_varint(int(x)) for x in arr.reshape(-1))))
        entries += _ld(1, _ld(1, name.encode('utf-8')) + _ld(2, feat))
    return _ld(1, entries)
