"""TF-V2 checkpoint bundle reader/writer (spatialaudiogen_b200/tf_checkpoint.py): checksums against the RFC 3720
vectors, the table reader against a table laid out byte by byte from the LevelDB format description, round trips."""
import os
import struct

import numpy as np
import pytest

from spatialaudiogen_b200 import tf_checkpoint as T


def test_crc32c_rfc3720_vectors_and_masking():
    assert T.crc32c(b'123456789') == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA
    assert T.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b'6789', T.crc32c(b'12345')) == 0xE3069283          # incremental
    for c in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert T.unmask_crc(T.mask_crc(c)) == c
    assert T.mask_crc(0) == 0xa282ead8                                  # ((0 >> 15) | (0 << 17)) + kMaskDelta


def _varint(v):
    out = b''
    while v >= 0x80:
        out += bytes([(v & 0x7F) | 0x80])
        v >>= 7
    return out + bytes([v])


def test_reader_on_a_hand_laid_table(tmp_path):
    """One data block with two prefix-compressed entries + index block + footer, written field by field from the
    LevelDB table format (not through write_table)."""
    payload = np.arange(6, dtype='<f4').reshape(2, 3)
    raw = payload.tobytes()
    # BundleEntryProto{dtype=1 (DT_FLOAT), shape{dim{size:2} dim{size:3}}, size=24, crc32c}
    shape = b'\x12\x02\x08\x02' + b'\x12\x02\x08\x03'
    entry = b'\x08\x01' + b'\x12' + _varint(len(shape)) + shape + b'\x28' + _varint(len(raw)) + b'\x35' + struct.pack('<I', T.mask_crc(T.crc32c(raw)))
    raw2 = (payload * 2).tobytes()                                      # second tensor at offset 24
    entry2 = (b'\x08\x01' + b'\x12' + _varint(len(shape)) + shape + b'\x20' + _varint(24) + b'\x28' + _varint(len(raw2)) + b'\x35'
              + struct.pack('<I', T.mask_crc(T.crc32c(raw2))))
    header = b'\x08\x01'                                                 # BundleHeaderProto{num_shards: 1}
    k1, k2 = b'net/w', b'net/w_1'
    block = (b'\x00\x00' + _varint(len(header)) + header                 # key "": shared 0, non-shared 0
             + b'\x00' + _varint(len(k1)) + _varint(len(entry)) + k1 + entry
             + _varint(5) + _varint(2) + _varint(len(entry2)) + k2[5:] + entry2      # shares "net/w"
             + struct.pack('<I', 0) + struct.pack('<I', 1))              # one restart at 0
    def with_trailer(b):
        return b + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(b'\x00', T.crc32c(b))))
    data_handle = _varint(0) + _varint(len(block))
    index_block = b'\x00' + _varint(len(k2)) + _varint(len(data_handle)) + k2 + data_handle + struct.pack('<I', 0) + struct.pack('<I', 1)
    meta_block = struct.pack('<I', 0) + struct.pack('<I', 1)
    body = with_trailer(block)
    meta_off = len(body)
    body += with_trailer(meta_block)
    idx_off = len(body)
    body += with_trailer(index_block)
    footer = _varint(meta_off) + _varint(len(meta_block)) + _varint(idx_off) + _varint(len(index_block))
    body += footer + bytes(40 - len(footer)) + struct.pack('<Q', T.TABLE_MAGIC)
    prefix = str(tmp_path / 'model.ckpt-7')
    open(prefix + '.index', 'wb').write(body)
    open(prefix + '.data-00000-of-00001', 'wb').write(raw + (payload * 2).tobytes())
    open(str(tmp_path / 'checkpoint'), 'w').write('model_checkpoint_path: "model.ckpt-7"\n')
    assert T.latest_checkpoint(str(tmp_path)) == prefix
    assert [v[:2] for v in T.list_variables(prefix)] == [('net/w', (2, 3)), ('net/w_1', (2, 3))]
    got = T.read_bundle(prefix, verify_data=True)
    np.testing.assert_array_equal(got['net/w'], payload)
    np.testing.assert_array_equal(got['net/w_1'], payload * 2)
    # a flipped byte in the block is caught by the block checksum
    bad = bytearray(body)
    bad[10] ^= 1
    open(prefix + '.index', 'wb').write(bytes(bad))
    with pytest.raises(ValueError):
        T.read_bundle(prefix)


def test_bundle_round_trip_with_many_variables(tmp_path):
    from spatialaudiogen_b200 import weights as Wt
    shapes = Wt.variable_shapes(['audio', 'video'])
    rng = np.random.RandomState(0)
    tensors = {}
    for k, sh in shapes.items():                                        # small stand-ins with the real names (many index blocks)
        tensors[k] = rng.randn(*[min(d, 3) for d in sh]).astype(np.float32)
        tensors[k + '/Adam'] = np.zeros_like(tensors[k])                 # optimizer slots live in the same bundle
    tensors['global_step'] = np.asarray(12345, dtype=np.int64)
    tensors['beta1_power'] = np.asarray(0.9, dtype=np.float32)
    prefix = str(tmp_path / 'model.ckpt-12345')
    T.write_bundle(prefix, tensors)
    assert T.latest_checkpoint(str(tmp_path)) == prefix
    got = T.read_bundle(prefix, verify_data=True)
    assert list(got) == sorted(tensors)                                 # table order = key order
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
        np.testing.assert_array_equal(got[k], v)
    only = T.read_bundle(prefix, names={'global_step', 'audio_encoder/conv1/weights'})
    assert set(only) == {'global_step', 'audio_encoder/conv1/weights'} and int(only['global_step']) == 12345
    assert T.load_model_dir(str(tmp_path), names={'beta1_power'})['beta1_power'] == np.float32(0.9)


def test_load_model_dir_falls_back_to_npz(tmp_path):
    np.savez(str(tmp_path / 'weights.npz'), **{'a/weights': np.ones((2, 2), np.float32)})
    assert T.load_model_dir(str(tmp_path))['a/weights'].shape == (2, 2)
    os.remove(str(tmp_path / 'weights.npz'))
    with pytest.raises(IOError):
        T.load_model_dir(str(tmp_path))
