"""TF tensor-bundle reader (video_prediction_b200/tf_checkpoint.py) against an independent WRITER of the same published
format (LevelDB table with prefix compression and restart points, BundleEntryProto / BundleHeaderProto by hand).  No
TensorFlow-written file exists in this environment: self-consistency only (see the module docstring)."""
import os
import struct

import numpy as np
import pytest

from video_prediction_b200 import tf_checkpoint as T


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def proto_varint(field, v):
    return varint(field << 3) + varint(v)


def proto_bytes(field, b):
    return varint((field << 3) | 2) + varint(len(b)) + b


def build_block(items, restart_interval=3):
    """LevelDB block builder: shared-prefix compression, a restart point every `restart_interval` entries."""
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_bundle(prefix, tensors, entries_per_block=4):
    dtype_code = {np.dtype(np.float32): 1, np.dtype(np.int64): 9, np.dtype(np.int32): 3}
    data, entries = bytearray(), []
    for name in sorted(tensors):
        a = np.array(tensors[name], order='C')          # (np.ascontiguousarray would turn 0-d scalars into shape (1,))
        shape = b''.join(proto_bytes(2, proto_varint(1, int(d))) for d in a.shape)
        e = proto_varint(1, dtype_code[a.dtype]) + proto_bytes(2, shape) + proto_varint(3, 0) + proto_varint(4, len(data)) + \
            proto_varint(5, a.nbytes) + varint((6 << 3) | 5) + struct.pack('<I', 0)
        entries.append((name.encode(), e))
        data += a.tobytes()
    header = proto_varint(1, 1) + proto_varint(2, 0) + proto_bytes(3, proto_varint(1, 1))
    items = [(b'', header)] + entries
    table, index_items = bytearray(), []
    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        block = build_block(chunk)
        handle = varint(len(table)) + varint(len(block))
        table += block + b'\x00' + struct.pack('<I', 0)            # compression type none + (unchecked) crc
        index_items.append((chunk[-1][0] + b'\x00', handle))       # a separator >= the last key of the block
    meta = build_block([])
    meta_handle = varint(len(table)) + varint(len(meta))
    table += meta + b'\x00' + struct.pack('<I', 0)
    index = build_block(index_items, restart_interval=1)
    index_handle = varint(len(table)) + varint(len(index))
    table += index + b'\x00' + struct.pack('<I', 0)
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', T.TABLE_MAGIC)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(table) + footer)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))


def test_reader_round_trip_and_name_mapping(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {
        'generator/rnn/dna_cell/h0/conv_pool2d/kernel': rng.standard_normal((5, 5, 14, 32)).astype(np.float32),   # historical scope name
        'generator/rnn/dna_cell/h0/conv_pool2d/kernel/Adam': rng.standard_normal((5, 5, 14, 32)).astype(np.float32),
        'generator/rnn/dna_cell/h0/conv_pool2d/kernel/Adam_1': rng.random((5, 5, 14, 32)).astype(np.float32),
        'generator/rnn/dna_cell/h0/conv_pool2d/bias': rng.standard_normal((32,)).astype(np.float32),
        'generator/encoder/z_mu/dense/kernel': rng.standard_normal((256, 8)).astype(np.float32),
        'discriminator/video/sn_conv0_0/conv3d/u': rng.standard_normal((1, 32)).astype(np.float32),
        'global_step': np.array(300000, np.int64),
        'beta1_power': np.array(0.5 ** 3, np.float32),
        'unrelated/variable': np.arange(7, dtype=np.int32),
    }
    prefix = str(tmp_path / 'model-300000')
    write_bundle(prefix, tensors)
    with open(os.path.join(str(tmp_path), 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "model-300000"\nall_model_checkpoint_paths: "model-300000"\n')
    assert T.latest_checkpoint(str(tmp_path)) == prefix and T.is_tf_checkpoint(prefix)
    r = T.CheckpointReader(prefix)
    assert r.header == dict(num_shards=1, endianness=0)
    shapes = r.get_variable_to_shape_map()
    assert set(shapes) == set(tensors) and shapes['generator/encoder/z_mu/dense/kernel'] == [256, 8] and shapes['global_step'] == []
    for k, v in tensors.items():
        assert np.array_equal(r.get_tensor(k), v), k
    # the SAVP model's restore mapping (savp_model.py:848-855): savp_cell falls back to dna_cell
    from video_prediction_b200.models.savp_model import SAVPVideoPredictionModel as M
    wanted = ['generator/rnn/savp_cell/h0/conv_pool2d/kernel', 'generator/rnn/savp_cell/h0/conv_pool2d/bias',
              'generator/encoder/z_mu/dense/kernel', 'generator/rnn/savp_cell/h1/conv_pool2d/kernel']
    found, missing, unused, extra = T.load_variables(prefix, wanted, M.restore_to_checkpoint_mapping)
    assert missing == ['generator/rnn/savp_cell/h1/conv_pool2d/kernel']
    assert np.array_equal(found[wanted[0]], tensors['generator/rnn/dna_cell/h0/conv_pool2d/kernel'])
    assert extra['global_step'] == 300000 and abs(extra['beta1_power'] - 0.125) < 1e-7
    assert np.array_equal(extra['slots']['m'][wanted[0]], tensors['generator/rnn/dna_cell/h0/conv_pool2d/kernel/Adam'])
    assert 'unrelated/variable' in unused and 'discriminator/video/sn_conv0_0/conv3d/u' in unused
    with pytest.raises(ValueError):
        open(str(tmp_path / 'bad.index'), 'wb').write(b'x' * 64)
        T.CheckpointReader(str(tmp_path / 'bad'))
