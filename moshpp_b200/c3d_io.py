"""Minimal C3D point-data reader / writer (float or scaled-int16 frames; Intel, DEC and SGI/MIPS processor formats).

The reference reads .c3d through ezc3d (tools/mocap_interface.py:120-128), which is not
installable here; this module covers what that call site consumes: ``points`` (F x N x 3, invalid
samples as NaN), ``POINT:RATE`` and ``POINT:LABELS``.  Format facts follow the public C3D
specification (the reference's unused tools/c3d.py:35-190,191-424,1118-1393 served as the format spec):
the fourth byte of the parameter section names the processor format of the whole file -- 84 Intel
(little-endian, IEEE floats), 85 DEC (little-endian integers, VAX F-floating numbers), 86 SGI/MIPS
(big-endian, IEEE floats).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import numpy as np

_BLOCK = 512
PROCESSOR_INTEL, PROCESSOR_DEC, PROCESSOR_MIPS = 84, 85, 86
_PROCESSORS = {'intel': PROCESSOR_INTEL, 'dec': PROCESSOR_DEC, 'mips': PROCESSOR_MIPS}


# --------------------------------------------------------------------------------------------------------------------
# VAX F-floating <-> IEEE single.  A DEC number is stored as two little-endian 16-bit words, the word with sign,
# exponent and the high fraction bits FIRST; with the words swapped the bits read sign | exponent | fraction like an
# IEEE single whose value is four times the DEC number (excess-128 exponent and a 0.1f mantissa against excess-127
# and 1.f).  Exponent 0 is zero in DEC.  (The reference's reader decrements the exponent field by two instead,
# tools/c3d.py:157-190: the same numbers wherever the IEEE result is a normal number.)
# --------------------------------------------------------------------------------------------------------------------
def dec_to_ieee(buf: bytes) -> np.ndarray:
    w = np.frombuffer(buf, dtype='<u2').reshape(-1, 2)
    bits = (w[:, 0].astype(np.uint32) << 16) | w[:, 1].astype(np.uint32)
    out = bits.view(np.float32) * np.float32(0.25)
    return np.where((bits & 0x7F800000) == 0, np.float32(0), out).astype(np.float32)


def ieee_to_dec(values: np.ndarray) -> bytes:
    v = np.asarray(values, dtype=np.float32).ravel()
    bits = (v * np.float32(4.0)).astype(np.float32).view(np.uint32)
    bits = np.where(v == 0, np.uint32(0), bits)
    w = np.empty((len(v), 2), dtype='<u2')
    w[:, 0] = (bits >> 16).astype(np.uint16)
    w[:, 1] = (bits & 0xFFFF).astype(np.uint16)
    return w.tobytes()


class _Fmt:
    """Scalar and array codecs of one processor format."""

    def __init__(self, processor: int):
        if processor not in (PROCESSOR_INTEL, PROCESSOR_DEC, PROCESSOR_MIPS):
            raise ValueError(f'unknown C3D processor type {processor}')
        self.processor = processor
        self.e = '>' if processor == PROCESSOR_MIPS else '<'

    def i16(self, b: bytes, off: int = 0) -> int:
        return struct.unpack_from(self.e + 'h', b, off)[0]

    def u16(self, b: bytes, off: int = 0) -> int:
        return struct.unpack_from(self.e + 'H', b, off)[0]

    def f32(self, b: bytes, off: int = 0) -> float:
        if self.processor == PROCESSOR_DEC:
            return float(dec_to_ieee(b[off:off + 4])[0])
        return struct.unpack_from(self.e + 'f', b, off)[0]

    def pack_i16(self, v: int) -> bytes:
        return struct.pack(self.e + 'h', v)

    def pack_u16(self, v: int) -> bytes:
        return struct.pack(self.e + 'H', v)

    def pack_f32(self, v: float) -> bytes:
        if self.processor == PROCESSOR_DEC:
            return ieee_to_dec(np.array([v], dtype=np.float32))
        return struct.pack(self.e + 'f', v)

    def floats(self, raw: bytes, count: int, offset: int) -> np.ndarray:
        if self.processor == PROCESSOR_DEC:
            return dec_to_ieee(raw[offset: offset + 4 * count])
        return np.frombuffer(raw, dtype=self.e + 'f4', count=count, offset=offset)

    def pack_floats(self, a: np.ndarray) -> bytes:
        if self.processor == PROCESSOR_DEC:
            return ieee_to_dec(a)
        return np.asarray(a, dtype=self.e + 'f4').tobytes()


def _group(fm: _Fmt, gid: int, name: str, desc: str = '') -> bytes:
    nb, db = name.encode(), desc.encode()
    body = struct.pack('<bb', len(nb), -gid) + nb
    body += fm.pack_i16(3 + len(db)) + struct.pack('<B', len(db)) + db
    return body


def _param(fm: _Fmt, gid: int, name: str, dtype: int, dims: Sequence[int], data: bytes, desc: str = '') -> bytes:
    nb, db = name.encode(), desc.encode()
    payload = struct.pack('<bB', dtype, len(dims)) + bytes(dims) + data + struct.pack('<B', len(db)) + db
    return struct.pack('<bb', len(nb), gid) + nb + fm.pack_i16(2 + len(payload)) + payload


def write_c3d(fname: str, markers: np.ndarray, labels: List[str], frame_rate: float = 120.0,
              units: str = 'mm', processor: str = 'intel') -> None:
    """markers: F x N x 3 (NaN = missing) in ``units``; written as float32 frames in the given processor format."""
    fm = _Fmt(_PROCESSORS[processor])
    markers = np.asarray(markers, dtype=np.float64)
    F, N, _ = markers.shape
    if N > 255 or F > 65535:
        raise ValueError('minimal writer: at most 255 points and 65535 frames')
    lab_len = max(4, max(len(l) for l in labels))
    lab_bytes = b''.join(l.encode().ljust(lab_len) for l in labels)
    recs = _group(fm, 1, 'POINT', 'point data')
    recs += _param(fm, 1, 'USED', 2, [], fm.pack_i16(N))
    recs += _param(fm, 1, 'FRAMES', 2, [], fm.pack_u16(F))
    recs += _param(fm, 1, 'SCALE', 4, [], fm.pack_f32(-1.0))
    recs += _param(fm, 1, 'RATE', 4, [], fm.pack_f32(float(frame_rate)))
    recs += _param(fm, 1, 'UNITS', -1, [len(units)], units.encode())
    recs += _param(fm, 1, 'LABELS', -1, [lab_len, N], lab_bytes)
    n_param_blocks = (4 + len(recs) + 64 + _BLOCK - 1) // _BLOCK
    data_start = 2 + n_param_blocks
    recs += _param(fm, 1, 'DATA_START', 2, [], fm.pack_i16(data_start))
    recs += struct.pack('<bb', 0, 0)
    psec = struct.pack('<BBBB', 1, 80, n_param_blocks, fm.processor) + recs
    assert len(psec) <= n_param_blocks * _BLOCK
    psec = psec.ljust(n_param_blocks * _BLOCK, b'\0')

    hdr = struct.pack('<BB', 2, 0x50) + fm.pack_u16(N) + fm.pack_u16(0) + fm.pack_u16(1) + fm.pack_u16(F) + fm.pack_u16(0) \
        + fm.pack_f32(-1.0) + fm.pack_u16(data_start) + fm.pack_u16(0) + fm.pack_f32(float(frame_rate))
    hdr = hdr.ljust(_BLOCK, b'\0')

    valid = ~np.isnan(markers).any(-1)
    frames = np.zeros((F, N, 4), dtype=np.float32)
    frames[..., :3] = np.where(valid[..., None], markers, 0.0)
    frames[..., 3] = np.where(valid, 0.0, -1.0)
    data = fm.pack_floats(frames)
    data = data.ljust(((len(data) + _BLOCK - 1) // _BLOCK) * _BLOCK, b'\0')
    with open(fname, 'wb') as f:
        f.write(hdr + psec + data)


def _read_params(buf: bytes, fm: _Fmt):
    params = {}
    groups = {}
    pos = 4
    while pos < len(buf):
        nlen, gid = struct.unpack_from('<bb', buf, pos)
        if nlen == 0:
            break
        nlen = abs(nlen)
        name = buf[pos + 2: pos + 2 + nlen].decode('latin-1').upper()
        off_pos = pos + 2 + nlen
        offset = fm.i16(buf, off_pos)
        if gid < 0:
            groups[-gid] = name
        else:
            q = off_pos + 2
            dtype, ndim = struct.unpack_from('<bB', buf, q)
            dims = list(buf[q + 2: q + 2 + ndim])
            q += 2 + ndim
            count = int(np.prod(dims)) if dims else 1
            nbytes = abs(dtype) * count
            params[(gid, name)] = (dtype, dims, buf[q: q + nbytes])
        if offset <= 0:
            break
        pos = off_pos + offset
    return {(groups.get(g, str(g)), n): v for (g, n), v in params.items()}


def read_c3d(fname: str) -> Tuple[np.ndarray, List[str], float]:
    """Returns (points F x N x 3 with NaN for invalid samples, labels, frame_rate)."""
    with open(fname, 'rb') as f:
        raw = f.read()
    pblock, magic = struct.unpack_from('<BB', raw, 0)
    if magic != 0x50:
        raise ValueError(f'{fname}: not a C3D file')
    pstart = (pblock - 1) * _BLOCK
    try:
        fm = _Fmt(raw[pstart + 3])
    except ValueError as e:
        raise ValueError(f'{fname}: {e}')
    n_pts, n_analog, first, last, _gap = (fm.u16(raw, 2 + 2 * i) for i in range(5))
    scale = fm.f32(raw, 12)
    data_block, _apf = fm.u16(raw, 16), fm.u16(raw, 18)
    rate = fm.f32(raw, 20)
    n_pblocks = raw[pstart + 2]
    P = _read_params(raw[pstart: pstart + n_pblocks * _BLOCK], fm)

    def scalar(key, kind, default):
        if key in P:
            return getattr(fm, kind)(P[key][2], 0)
        return default

    n_pts = scalar(('POINT', 'USED'), 'i16', n_pts)
    n_frames = scalar(('POINT', 'FRAMES'), 'u16', last - first + 1)
    scale = scalar(('POINT', 'SCALE'), 'f32', scale)
    rate = scalar(('POINT', 'RATE'), 'f32', rate)
    data_block = scalar(('POINT', 'DATA_START'), 'i16', data_block)
    labels: List[str] = []
    if ('POINT', 'LABELS') in P:
        _, dims, blob = P[('POINT', 'LABELS')]
        ll, cnt = dims[0], (dims[1] if len(dims) > 1 else 1)
        labels = [blob[i * ll:(i + 1) * ll].decode('latin-1').strip() for i in range(cnt)]
    ofs = (data_block - 1) * _BLOCK
    if scale < 0:
        stride = 4 * (4 * n_pts + n_analog)
        arr = fm.floats(raw, n_frames * stride // 4, ofs)
        arr = arr.reshape(n_frames, -1)[:, :4 * n_pts].reshape(n_frames, n_pts, 4).astype(np.float64)
        pts, resid = arr[..., :3], arr[..., 3]
    else:
        stride = 2 * (4 * n_pts + n_analog)
        arr = np.frombuffer(raw, dtype=fm.e + 'i2', count=n_frames * stride // 2, offset=ofs)
        arr = arr.reshape(n_frames, -1)[:, :4 * n_pts].reshape(n_frames, n_pts, 4).astype(np.float64)
        pts, resid = arr[..., :3] * scale, arr[..., 3]
    pts = np.where((resid < 0)[..., None], np.nan, pts)
    return pts, labels, float(rate)
