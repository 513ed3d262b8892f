"""CPU tests of the TFRecord input side: record framing + CRC, tf.train.Example codec, the BAIR
(`SoftmotionVideoDataset`) and KTH (`KTHVideoDataset`) datasets on fixture files written here in the reference's formats
(softmotion_dataset.py:10-82, kth_dataset.py:16-50, base_dataset.py:129-453)."""
import os
import struct

import numpy as np
import pytest

from video_prediction_b200 import datasets
from video_prediction_b200.datasets import tfrecord as R


def test_crc32c_known_answers_and_record_framing(tmp_path):
    assert R.crc32c(b'123456789') == 0xE3069283                          # the CRC-32C check value
    assert R.crc32c(b'') == 0
    assert R.masked_crc32c(b'\x00' * 8) == ((((R.crc32c(b'\x00' * 8) >> 15) | (R.crc32c(b'\x00' * 8) << 17)) + 0xA282EAD8) & 0xFFFFFFFF)
    path = str(tmp_path / 'a.tfrecords')
    recs = [b'', b'x', os.urandom(1000)]
    R.write_records(path, recs)
    assert list(R.read_records(path, verify_data_crc=True)) == recs
    raw = bytearray(open(path, 'rb').read())
    raw[-6] ^= 0xFF                                                      # flip a data byte of the last record
    open(path, 'wb').write(bytes(raw))
    assert len(list(R.read_records(path))) == 3                          # header CRCs still fine
    with pytest.raises(R.TFRecordError):
        list(R.read_records(path, verify_data_crc=True))
    raw[0] ^= 0x01                                                       # corrupt the first length field
    open(path, 'wb').write(bytes(raw))
    with pytest.raises(R.TFRecordError):
        list(R.read_records(path))


def test_example_codec_round_trip_and_unpacked_lists():
    feats = {'a/bytes': [b'\x00\x01\xff', b''], 'b': np.array([1.5, -2.25, 3e-8], np.float32), 'c': np.array([0, -1, 2 ** 40, 7]),
             'single': b'frame'}
    ex = R.parse_example(R.make_example(feats))
    assert ex['a/bytes'] == ('bytes', [b'\x00\x01\xff', b''])
    assert ex['single'] == ('bytes', [b'frame'])
    assert ex['b'][0] == 'float' and np.array_equal(ex['b'][1], feats['b'])
    assert ex['c'][0] == 'int64' and ex['c'][1].tolist() == [0, -1, 2 ** 40, 7]
    # writers other than TF's C++ one may emit repeated scalars unpacked (one tag per element): same parse
    unpacked_f = b''.join(b'\x0d' + struct.pack('<f', v) for v in (1.0, 2.0))          # field 1, wire type 5
    unpacked_i = b'\x08\x03\x08\x04'                                                  # field 1, wire type 0
    def entry(name, feature_field, payload):
        feat = bytes([feature_field << 3 | 2, len(payload)]) + payload
        e = b'\x0a' + bytes([len(name)]) + name.encode() + b'\x12' + bytes([len(feat)]) + feat
        return b'\x0a' + bytes([len(e)]) + e
    body = entry('f', 2, unpacked_f) + entry('i', 3, unpacked_i)
    ex = R.parse_example(b'\x0a' + bytes([len(body)]) + body)
    assert ex['f'][1].tolist() == [1.0, 2.0] and ex['i'][1].tolist() == [3, 4]


def _write_bair(root, split, ntraj=5, length=30, side=8):
    d = os.path.join(root, split)
    os.makedirs(d)
    recs = []
    for k in range(ntraj):
        f = {}
        for t in range(length):
            f['%d/image_aux1/encoded' % t] = bytes([(k * 40 + t) % 256]) * (side * side * 3)
            f['%d/endeffector_pos' % t] = np.array([k, t, 0.5], np.float32)
            if t < length - 1:
                f['%d/action' % t] = np.array([k, t, -t, 1.0], np.float32)
        recs.append(R.make_example(f))
    R.write_records(os.path.join(d, 'traj_0_to_%d.tfrecords' % (ntraj - 1)), recs)


def test_bair_dataset_shapes_time_shift_and_epochs(tmp_path):
    root = str(tmp_path / 'bair')
    _write_bair(root, 'train')
    ds = datasets.get_dataset_class('bair')(root, mode='train', num_epochs=1, seed=3, hparams_dict={'use_state': True})
    assert ds.hparams.sequence_length == 12 and ds.hparams.context_frames == 2 and ds.hparams.time_shift == 2
    assert ds.num_examples_per_epoch() == 5 and ds._max_sequence_length == 30
    assert ds.state_like_names_and_shapes['images'] == ('%d/image_aux1/encoded', (8, 8, 3))
    seen = []
    for _ in range(2):
        b = ds.make_batch(2)
        assert b['images'].shape == (2, 12, 8, 8, 3) and b['images'].dtype == np.float32
        assert b['actions'].shape == (2, 11, 4) and b['states'].shape == (2, 12, 3)
        for i in range(2):
            vals = np.rint(b['images'][i, :, 0, 0, 0] * 255).astype(int)
            k, t0 = int(b['states'][i, 0, 0]), int(b['states'][i, 0, 1])
            assert t0 % 2 == 0 and 0 <= t0 <= 18                          # a multiple of time_shift that leaves 12 frames
            assert vals.tolist() == [k * 40 + t0 + j for j in range(12)]
            assert b['actions'][i, :, 1].tolist() == [t0 + j for j in range(11)]
            assert float(b['images'][i].max()) <= 1.0
            seen.append(k)
    assert len(set(seen)) == 4                                            # four different trajectories, shuffled
    with pytest.raises(StopIteration):                                    # 5 examples, batch 2: the remainder is dropped
        ds.make_batch(2)


def test_bair_frame_skip_macro_actions_and_test_mode_order(tmp_path):
    root = str(tmp_path / 'bair')
    _write_bair(root, 'test', ntraj=3)
    ds = datasets.SoftmotionVideoDataset(root, mode='test', hparams_dict={'use_state': True, 'frame_skip': 1, 'sequence_length': 5})
    b = ds.make_batch(3)
    assert b['states'][:, 0, 0].tolist() == [0.0, 1.0, 2.0]               # file order, no shuffling, no time shift in test mode
    assert b['states'][0, :, 1].tolist() == [0, 2, 4, 6, 8]               # every other frame
    assert b['actions'].shape == (3, 4, 8)                                # two skipped-frame actions concatenated per step
    assert b['actions'][1, 2].tolist() == [1, 4, -4, 1, 1, 5, -5, 1]
    ds2 = datasets.SoftmotionVideoDataset(root, mode='test', hparams_dict={'sequence_length': 0})
    assert ds2.hparams.sequence_length == 30 and 'actions' not in ds2.make_batch(1)
    with pytest.raises(FileNotFoundError):
        datasets.SoftmotionVideoDataset(str(tmp_path / 'nowhere'))
    with pytest.raises(ValueError):
        datasets.SoftmotionVideoDataset(root, mode='eval')
    with pytest.raises(ValueError):                                       # 30 frames cannot give 20 frames at stride 2
        datasets.SoftmotionVideoDataset(root, mode='test', hparams_dict={'frame_skip': 1, 'sequence_length': 20,
                                                                         'force_time_shift': True}).make_batch(1)


def test_kth_dataset_filters_short_clips_and_samples_a_window(tmp_path):
    root = str(tmp_path / 'kth')
    d = os.path.join(root, 'val')
    os.makedirs(d)
    lengths, recs = [15, 25, 30], []
    for k, n in enumerate(lengths):
        frames = [bytes([(k * 50 + t) % 256]) * (4 * 6 * 1) for t in range(n)]
        recs.append(R.make_example({'sequence_length': np.array([n]), 'height': np.array([4]), 'width': np.array([6]),
                                    'channels': np.array([1]), 'images/encoded': frames}))
    R.write_records(os.path.join(d, 'sequence_0_to_2.tfrecords'), recs)
    open(os.path.join(d, 'sequence_lengths.txt'), 'w').write('\n'.join(map(str, lengths)) + '\n')
    ds = datasets.get_dataset_class('kth')(root, mode='val', num_epochs=1, seed=0)
    assert ds.hparams.sequence_length == 20 and ds.hparams.context_frames == 10 and ds.hparams.force_time_shift
    assert ds.num_examples_per_epoch() == 2                               # the 15-frame clip is filtered out
    b = ds.make_batch(2)
    assert b['images'].shape == (2, 20, 4, 6, 1)
    for i in range(2):
        vals = np.rint(b['images'][i, :, 0, 0, 0] * 255).astype(int)
        k = vals[0] // 50
        assert k in (1, 2) and vals.tolist() == list(range(vals[0], vals[0] + 20))
        assert vals[0] - k * 50 <= lengths[k] - 20                        # the window fits inside the clip
    with pytest.raises(StopIteration):
        ds.make_batch(1)


def test_registry_names():
    assert datasets.get_dataset_class('softmotion') is datasets.SoftmotionVideoDataset
    assert datasets.get_dataset_class('KTHVideoDataset') is datasets.KTHVideoDataset
    assert datasets.get_dataset_class('synthetic') is datasets.SyntheticVideoDataset
    with pytest.raises(NotImplementedError):
        datasets.get_dataset_class('ucf101')
    with pytest.raises(ValueError):
        datasets.get_dataset_class('imagenet')
