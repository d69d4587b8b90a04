"""Dependency-free reader / writer of TensorFlow checkpoint V2 bundles (`<prefix>.index` + `<prefix>.data-00000-of-00001`),
the format `tf.train.Saver` writes in the reference (code/homography_CNN_synthetic.py:303,359-360,389) and restores from
(:314-317, :517) — SURVEY §8f-2.  TensorFlow itself is not needed (and is not installable here).

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table — LevelDB's table format):
  .index  an SSTable: sorted key -> value entries in prefix-compressed blocks, each block followed by a 1-byte compression
          type and a masked CRC32C; an index block; a 48-byte footer with the magic 0xdb4775248b80fb57.  Key "" holds a
          BundleHeaderProto, every other key is a variable name holding a BundleEntryProto (dtype, shape, shard_id, offset,
          size, masked crc32c of the tensor bytes).
  .data-00000-of-00001   the tensors' raw little-endian bytes at the recorded offsets.
Variable names of the reference graph: `model/conv_block{1..4}/conv{1,2}/{weights,biases}`, `model/fc{1,2}/fc{1,2}/...`
(TF-Slim scopes, homography_model.py:97-98,108-131,358), Adam slots `<var>/Adam` (m) and `<var>/Adam_1` (v), `beta1_power`,
`beta2_power`, and the unnamed `global_step = tf.Variable(0, trainable=False)` (:154) which TF calls `Variable` (int32).
The writer emits uncompressed blocks (what TF's BundleWriter does); the reader rejects snappy-compressed blocks loudly.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_NP = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}

# ---------------------------------------------------------------------------------------------------------- CRC32C
_POLY = 0x82F63B78
_TABLE = np.zeros(256, dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE[_i] = _c
_TL = [int(x) for x in _TABLE]


def _crc_update_small(state, data):
    for b in data:
        state = _TL[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def _mat_apply(M, v):
    r, k = 0, 0
    while v:
        if v & 1:
            r ^= M[k]
        v >>= 1; k += 1
    return r


def _mat_mul(A, B):
    return [_mat_apply(A, b) for b in B]


def _zero_shift_operator(nbytes):
    """GF(2) matrix (32 columns) advancing the CRC register through `nbytes` zero bytes."""
    M = [_TL[(1 << k) & 0xFF] ^ ((1 << k) >> 8) for k in range(32)]          # one zero byte
    R = [1 << k for k in range(32)]                                           # identity
    while nbytes:
        if nbytes & 1:
            R = _mat_mul(M, R)
        M = _mat_mul(M, M)
        nbytes >>= 1
    return R


_native_crc = None        # udh_crc32c of the in-tree libudh.so (hardware crc32 instruction); False = not available


def _native():
    """The C ABI's host CRC (include/udh.h udh_crc32c) when libudh.so is built — a 410 MB checkpoint takes ~0.1 s instead of
    seconds; this module stays importable and correct without it (tools/tf_ckpt_to_npz.py on a machine without nvcc)."""
    global _native_crc
    if _native_crc is None:
        try:
            from . import _lib
            _native_crc = _lib.lib.udh_crc32c
        except Exception:
            _native_crc = False
    return _native_crc


def crc32c(data):
    """CRC-32C (Castagnoli) of bytes / a uint8 array: udh_crc32c of the C ABI when available, else crc32c_numpy."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).reshape(-1).view(np.uint8)
    fn = _native()
    if fn and a.size:
        import ctypes
        return int(fn(ctypes.c_void_p(a.ctypes.data), a.size, 0))
    return crc32c_numpy(a)


def crc32c_numpy(data):
    """Pure numpy CRC-32C.  Large buffers are processed as many equal chunks in parallel numpy lanes (the register update is
    linear over GF(2)): crc(s, a||b) = shift(crc(s, a), len b) xor crc(0, b)."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    n = a.size
    if n < (1 << 16):
        return _crc_update_small(0xFFFFFFFF, a.tolist()) ^ 0xFFFFFFFF
    lanes = 1 << 14
    L = n // lanes
    head = n - L * lanes                                       # leading bytes processed serially (fewer than `lanes`)
    state = _crc_update_small(0xFFFFFFFF, a[:head].tolist())
    body = a[head:].reshape(lanes, L)
    s = np.zeros(lanes, dtype=np.uint32)
    cols = np.ascontiguousarray(body.T)                        # [L, lanes]: one contiguous row per byte position
    for j in range(L):
        s = _TABLE[(s ^ cols[j]) & 0xFF] ^ (s >> 8)
    Z = _zero_shift_operator(L)
    for v in s.tolist():
        state = _mat_apply(Z, state) ^ v
    return state ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xa282ead8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------- protobuf bits
def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]; pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def _entry_proto(dtype, shape, offset, size, crc_masked):
    dims = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(int(s))) for s in shape))
    msg = _field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _varint(offset))
    msg += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc_masked))
    return msg


def _parse_message(buf):
    """-> {field: [values]} with varints as ints, length-delimited as bytes, fixed32 as ints."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        num, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(num, []).append(v)
    return out


# ---------------------------------------------------------------------------------------------------------- SSTable
def _build_block(items, restart_interval=16):
    """LevelDB block: entries (shared, non_shared, value_len, key delta, value), restart offsets, restart count."""
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _parse_block(buf):
    n_restarts = struct.unpack_from("<I", buf, len(buf) - 4)[0]
    end = len(buf) - 4 - 4 * n_restarts
    pos, key, items = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(buf, pos)
        non_shared, pos = _read_varint(buf, pos)
        vlen, pos = _read_varint(buf, pos)
        key = key[:shared] + bytes(buf[pos:pos + non_shared]); pos += non_shared
        items.append((key, bytes(buf[pos:pos + vlen]))); pos += vlen
    return items


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    block, ctype, crc = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if ctype != 0:
        raise ValueError("checkpoint index uses block compression type %d (snappy); only uncompressed tables are supported" % ctype)
    if verify and unmask_crc(crc) != crc32c(raw[:size + 1]):
        raise ValueError("checkpoint index block at offset %d fails its CRC32C" % offset)
    return block


# ---------------------------------------------------------------------------------------------------------- public API
def write_checkpoint(prefix, tensors):
    """tensors: {name: ndarray (float32 / int32 / int64)} -> <prefix>.index and <prefix>.data-00000-of-00001."""
    names = sorted(tensors, key=lambda s: s.encode())
    entries, offset = [], 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])
            if not a.flags.c_contiguous:
                a = a.copy(order="C")                            # (np.ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in _DT:
                raise TypeError("variable %s has unsupported dtype %s" % (n, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            entries.append((n.encode(), _entry_proto(_DT[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(np.frombuffer(raw, np.uint8))))))
            offset += len(raw)
    header = _field(1, 0, _varint(1)) + _field(3, 2, _varint(2) + _field(1, 0, _varint(1)))     # num_shards = 1, version.producer = 1
    items = [(b"", header)] + entries
    data_block = _build_block(items)
    meta_block = _build_block([])
    out = bytearray()
    data_handle = _varint(0) + _varint(len(data_block))
    out += _with_trailer(data_block)
    meta_off = len(out)
    out += _with_trailer(meta_block)
    index_block = _build_block([(items[-1][0] + b"\x00", data_handle)], restart_interval=1)
    index_off = len(out)
    out += _with_trailer(index_block)
    footer = _varint(meta_off) + _varint(len(meta_block)) + _varint(index_off) + _varint(len(index_block))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def read_checkpoint(prefix, verify=True):
    """-> {name: ndarray}.  Verifies the table CRCs and every tensor's CRC32C when verify is True."""
    with open(prefix + ".index", "rb") as f:
        f.seek(0, 2)
        size = f.tell()
        if size < 48:
            raise ValueError("%s.index is too short to be a TensorFlow checkpoint index" % prefix)
        f.seek(size - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != MAGIC:
            raise ValueError("%s.index: bad table magic (not a TensorFlow V2 checkpoint index)" % prefix)
        pos = 0
        _, pos = _read_varint(footer, pos); _, pos = _read_varint(footer, pos)
        ioff, pos = _read_varint(footer, pos); isz, pos = _read_varint(footer, pos)
        entries = []
        for _, handle in _parse_block(_read_block(f, ioff, isz, verify)):
            boff, p = _read_varint(handle, 0); bsz, p = _read_varint(handle, p)
            entries += _parse_block(_read_block(f, boff, bsz, verify))
    out, shards = {}, {}
    num_shards = 1
    for key, val in entries:
        msg = _parse_message(val)
        if key == b"":
            num_shards = msg.get(1, [1])[0]
            if msg.get(2, [0])[0] != 0:
                raise ValueError("big-endian checkpoints are not supported")
            continue
        dtype = msg.get(1, [0])[0]
        if dtype not in _NP:
            continue                                              # e.g. string tensors of a MetaGraph: not variables we map
        shape = []
        if 2 in msg:
            for d in _parse_message(msg[2][0]).get(2, []):
                shape.append(_parse_message(d).get(1, [0])[0])
        shard, off, sz = msg.get(3, [0])[0], msg.get(4, [0])[0], msg.get(5, [0])[0]
        if 7 in msg:
            raise ValueError("variable %s is stored as slices (partitioned variable): not supported" % key.decode())
        if shard not in shards:
            shards[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), "rb")
        fh = shards[shard]
        fh.seek(off)
        raw = fh.read(sz)
        if verify and 6 in msg and unmask_crc(msg[6][0]) != crc32c(np.frombuffer(raw, np.uint8)):
            raise ValueError("variable %s fails its CRC32C" % key.decode())
        out[key.decode()] = np.frombuffer(raw, dtype=_NP[dtype]).reshape(tuple(int(d) for d in shape)).copy()
    for fh in shards.values():
        fh.close()
    return out


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: the `model_checkpoint_path` of <model_dir>/checkpoint."""
    p = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(p):
        return None
    for ln in open(p):
        if ln.startswith("model_checkpoint_path:"):
            path = ln.split(":", 1)[1].strip().strip('"')
            return path if os.path.isabs(path) else os.path.join(model_dir, path)
    return None


def update_checkpoint_state(model_dir, prefix, keep=5):
    """The `checkpoint` state file tf.train.Saver maintains (max_to_keep = 5 in the reference, :303)."""
    p = os.path.join(model_dir, "checkpoint")
    old = []
    if os.path.exists(p):
        old = [ln.split(":", 1)[1].strip().strip('"') for ln in open(p) if ln.startswith("all_model_checkpoint_paths:")]
    name = os.path.relpath(prefix, model_dir) if os.path.isabs(prefix) or os.path.dirname(prefix) else prefix
    allp = [x for x in old if x != name] + [name]
    for dead in allp[:-keep]:
        for ext in (".index", ".data-00000-of-00001"):
            try:
                os.remove(os.path.join(model_dir, dead) + ext)
            except OSError:
                pass
    allp = allp[-keep:]
    with open(p, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % name)
        for x in allp:
            f.write('all_model_checkpoint_paths: "%s"\n' % x)


# ---------------------------------------------------------------------------------------------------------- the model's variables
def engine_state_to_variables(params_flat, adam_m, adam_v, global_step, specs, beta1=0.9, beta2=0.999):
    """Flat buffers (package layout) -> the variable dictionary of the reference graph."""
    out = {}
    for n, s in specs.items():
        sl = slice(s.offset, s.offset + s.size)
        out[n] = np.asarray(params_flat[sl], np.float32).reshape(s.shape)
        if adam_m is not None:
            out[n + "/Adam"] = np.asarray(adam_m[sl], np.float32).reshape(s.shape)
            out[n + "/Adam_1"] = np.asarray(adam_v[sl], np.float32).reshape(s.shape)
    out["Variable"] = np.array(int(global_step), dtype=np.int32)              # global_step = tf.Variable(0, trainable=False)
    if adam_m is not None:
        out["beta1_power"] = np.array(beta1 ** (int(global_step) + 1), dtype=np.float32)     # AdamOptimizer's accumulators after t updates
        out["beta2_power"] = np.array(beta2 ** (int(global_step) + 1), dtype=np.float32)
    return out


def variables_to_engine_state(variables, specs, total_floats):
    """Inverse mapping; Adam slots / global_step are optional (a checkpoint saved for inference may lack them)."""
    flat = np.zeros(total_floats, np.float32); m = np.zeros(total_floats, np.float32); v = np.zeros(total_floats, np.float32)
    have_slots = True
    for n, s in specs.items():
        if n not in variables:
            raise KeyError("checkpoint is missing variable %s" % n)
        a = np.asarray(variables[n], np.float32)
        if tuple(a.shape) != tuple(s.shape):
            raise ValueError("variable %s has shape %s, expected %s" % (n, a.shape, s.shape))
        sl = slice(s.offset, s.offset + s.size)
        flat[sl] = a.reshape(-1)
        if n + "/Adam" in variables and n + "/Adam_1" in variables:
            m[sl] = np.asarray(variables[n + "/Adam"], np.float32).reshape(-1)
            v[sl] = np.asarray(variables[n + "/Adam_1"], np.float32).reshape(-1)
        else:
            have_slots = False
    step = None
    for k in ("Variable", "global_step"):
        if k in variables:
            step = int(np.asarray(variables[k]).reshape(-1)[0])
            break
    return flat, (m if have_slots else None), (v if have_slots else None), step
