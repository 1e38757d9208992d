"""TensorFlow V2 checkpoint bundles without TensorFlow.

The reference restores its variables with `tf.train.Saver().restore(sess, tf.train.latest_checkpoint(model_dir))`
(reference deploy.py:79-87, eval.py:98-118; the published models of README.md:70-78 ship as
`checkpoint` + `model.ckpt-N.index` + `model.ckpt-N.data-00000-of-00001`).  This module reads (and, for tests and
for exporting, writes) that format so the same `model_dir` drops in:

  <prefix>.index                 an SSTable in the LevelDB table format (tensorflow/core/lib/io/table): prefix-compressed
                                 key/value blocks + restart arrays, each block followed by a 1-byte compression type and a
                                 masked crc32c, an index block, and a 48-byte footer ending in the magic
                                 0xdb4775248b80fb57.  Key "" holds a BundleHeaderProto, every other key is a tensor name
                                 whose value is a BundleEntryProto (dtype, shape, shard_id, offset, size, crc32c).
  <prefix>.data-SSSSS-of-NNNNN   the raw little-endian tensor bytes at [offset, offset + size) of shard SSSSS.
  checkpoint                     text proto naming the latest prefix (`model_checkpoint_path: "model.ckpt-N"`).

Third-party format, restated from its published specification; TensorFlow itself is not installable here, so the
reader is pinned by round trips against the writer and by the format's own checksums (parity unpinned against real
TF output -- DESIGN.md section 5).  Pure host code: no kernels, no oracle.
"""
import ctypes as C
import os
import re
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'), 6: np.dtype('i1'),
           9: np.dtype('<i8'), 10: np.dtype('?'), 17: np.dtype('<u2'), 19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


# ---- crc32c (Castagnoli), masked the way LevelDB / TensorFlow store it ----------------------------------------------
def _make_table():
    poly = 0x82F63B78
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t[i] = c
    return t


_CRC_TABLE = _make_table()
_CRC_TABLE_L = [int(x) for x in _CRC_TABLE]


_native_crc = None


def crc32c(data, crc=0):
    """CRC-32C of `data` (bytes-like), continuing from `crc`.  Long buffers (tensor payloads: 25 MB for `video-fc`) go through
    libsag.so's slicing-by-8 routine when the library is built; otherwise, and for the short index blocks, a byte-at-a-time
    table walk in Python."""
    global _native_crc
    mv = memoryview(data).cast('B')
    if len(mv) >= 1024 and _native_crc is not False:
        if _native_crc is None:
            try:
                from . import _lib
                _native_crc = _lib.lib().sag_crc32c
            except Exception:
                _native_crc = False
        if _native_crc:
            return int(_native_crc(bytes(mv) if mv.readonly else (C.c_char * len(mv)).from_buffer(mv), len(mv), crc))
    crc ^= 0xFFFFFFFF
    t = _CRC_TABLE_L
    for b in memoryview(data).cast('B'):
        crc = t[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf wire format --------------------------------------------------------------------------
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError('malformed varint')


def _put_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf):
    """{field number: [values]}; varints as int, length-delimited as bytes, fixed32/64 as int."""
    fields, pos = {}, 0
    buf = bytes(buf)
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        fields.setdefault(num, []).append(v)
    return fields


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _to_signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(value):
    """BundleEntryProto -> dict (tensorflow/core/protobuf/tensor_bundle.proto)."""
    f = _parse_message(value)
    shape = []
    if 2 in f:
        sp = _parse_message(f[2][0])                       # TensorShapeProto: repeated Dim dim = 2; Dim.size = 1
        for d in sp.get(2, []):
            dm = _parse_message(d)
            shape.append(_to_signed64(dm.get(1, [0])[0]))
    return {'dtype': f.get(1, [0])[0], 'shape': tuple(shape), 'shard_id': f.get(3, [0])[0], 'offset': f.get(4, [0])[0],
            'size': f.get(5, [0])[0], 'crc32c': f.get(6, [None])[0], 'sliced': 7 in f}


def _encode_entry(dtype_id, shape, shard_id, offset, size, crc):
    dims = b''.join(_field(2, 2, _put_varint(len(m)) + m) for m in (_field(1, 0, _put_varint(int(d))) for d in shape))
    out = _field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if shard_id:
        out += _field(3, 0, _put_varint(shard_id))
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack('<I', crc))
    return out


# ---- LevelDB table ----------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    contents = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(contents) != size or len(trailer) != 5:
        raise ValueError('table block [%d, +%d) runs past the end of the index file' % (offset, size))
    if trailer[0] != 0:
        raise ValueError('compressed table block (type %d): TensorFlow writes bundles uncompressed' % trailer[0])
    if verify:
        want = unmask_crc(struct.unpack('<I', trailer[1:])[0])
        if crc32c(trailer[:1], crc32c(contents)) != want:
            raise ValueError('crc32c mismatch in table block at offset %d' % offset)
    return contents


def _block_entries(block):
    """(key, value) pairs of one block, undoing the shared-prefix compression; the restart array is skipped."""
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    data = open(path, 'rb').read()
    if len(data) < 48:
        raise ValueError('%s is too short to be a table' % path)
    footer = data[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise ValueError('%s: bad table magic' % path)
    pos = 0
    _, pos = _get_varint(footer, pos)                      # metaindex handle (unused by bundles)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00' + struct.pack('<I', mask_crc(crc32c(b'\x00', crc32c(block)))))
    return _put_varint(off) + _put_varint(len(block))


def write_table(path, items, block_size=4096):
    """Write sorted (key, value) pairs as an uncompressed SSTable (data blocks of ~block_size, index block, footer)."""
    items = sorted(items)
    with open(path, 'wb') as f:
        index, cur, cur_bytes = [], [], 0
        for k, v in items:
            cur.append((k, v))
            cur_bytes += len(k) + len(v) + 3
            if cur_bytes >= block_size:
                index.append((cur[-1][0], _emit_block(f, _build_block(cur))))
                cur, cur_bytes = [], 0
        if cur or not index:
            index.append((cur[-1][0] if cur else b'', _emit_block(f, _build_block(cur))))
        meta = _emit_block(f, _build_block([]))
        idx = _emit_block(f, _build_block(index, restart_interval=1))
        footer = meta + idx
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC))


# ---- bundles -----------------------------------------------------------------------------------------------------------
def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: the prefix named by `<model_dir>/checkpoint`, or None."""
    path = os.path.join(model_dir, 'checkpoint')
    if not os.path.exists(path):
        return None
    for line in open(path):
        m = re.match(r'\s*model_checkpoint_path\s*:\s*"(.*)"', line)
        if m:
            p = m.group(1)
            if not os.path.isabs(p):
                return os.path.join(model_dir, p)
            # an absolute path written on the training machine: when it does not exist here, the bundle of that name next to
            # the `checkpoint` file is the one meant (model directories are copied around)
            local = os.path.join(model_dir, os.path.basename(p))
            return p if os.path.exists(p + '.index') or not os.path.exists(local + '.index') else local
    return None


def list_variables(prefix, verify=True):
    """[(name, shape, numpy dtype)] of a bundle, like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + '.index', verify):
        if k == b'':
            continue
        e = _parse_entry(v)
        out.append((k.decode('utf-8'), e['shape'], _DTYPES.get(e['dtype'])))
    return out


def read_bundle(prefix, names=None, verify_data=True):
    """OrderedDict name -> ndarray of the tensors in `<prefix>.index` / `.data-*` (all, or those in `names`).
    Index blocks are always checksummed; tensor payloads too unless verify_data=False."""
    entries = read_table(prefix + '.index', verify=True)
    header = dict(entries).get(b'')
    num_shards = 1
    if header is not None:
        h = _parse_message(header)
        num_shards = h.get(1, [1])[0] or 1
        if h.get(2, [0])[0] != 0:
            raise ValueError('big-endian bundle')
    want = None if names is None else set(names)
    shards, out = {}, OrderedDict()
    for k, v in entries:
        name = k.decode('utf-8')
        if k == b'' or (want is not None and name not in want):
            continue
        e = _parse_entry(v)
        if e['sliced']:
            raise ValueError('variable %s is stored as slices (partitioned variable): not supported' % name)
        if e['dtype'] not in _DTYPES:
            continue                                        # strings / resources: nothing the model restores
        dt = _DTYPES[e['dtype']]
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        n = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if raw.size != e['size'] or n * dt.itemsize != e['size']:
            raise ValueError('variable %s: %d bytes on disk, shape %s needs %d' % (name, raw.size, e['shape'], n * dt.itemsize))
        if verify_data and e['crc32c'] is not None and crc32c(raw.tobytes()) != unmask_crc(e['crc32c']):
            raise ValueError('crc32c mismatch in the data of variable %s' % name)
        out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e['shape'])
    return out


def write_bundle(prefix, tensors, update_checkpoint_file=True):
    """Write name -> ndarray as a single-shard V2 bundle (`<prefix>.index`, `<prefix>.data-00000-of-00001`) and, like
    tf.train.Saver.save, point `<dir>/checkpoint` at it."""
    items, offset = [], 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])
            if not a.flags.c_contiguous:                    # (ascontiguousarray would turn scalars into shape (1,))
                a = np.ascontiguousarray(a)
            dt = a.dtype.newbyteorder('<') if a.dtype.byteorder == '>' else a.dtype
            if np.dtype(dt) not in _DTYPE_IDS:
                raise ValueError('variable %s: dtype %s has no bundle encoding here' % (name, a.dtype))
            raw = a.astype(dt, copy=False).tobytes()
            f.write(raw)
            items.append((name.encode('utf-8'), _encode_entry(_DTYPE_IDS[np.dtype(dt)], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # num_shards=1, version.producer=1
    write_table(prefix + '.index', [(b'', header)] + items)
    if update_checkpoint_file:
        base = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def load_model_dir(model_dir, names=None):
    """Weights of a reference model directory: the latest TF bundle when there is a `checkpoint` file, else
    `weights.npz` (np.savez of the same name -> array mapping)."""
    prefix = latest_checkpoint(model_dir)
    if prefix is not None and os.path.exists(prefix + '.index'):
        return read_bundle(prefix, names)
    path = os.path.join(model_dir, 'weights.npz')
    if os.path.exists(path):
        with np.load(path) as z:
            return OrderedDict((k, z[k]) for k in z.files if names is None or k in names)
    raise IOError('%s holds neither a TensorFlow checkpoint (checkpoint + *.index) nor weights.npz' % model_dir)
