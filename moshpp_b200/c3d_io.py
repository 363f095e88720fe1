"""Minimal C3D point-data reader / writer (Intel byte order, float or scaled-int16 frames).

The reference reads .c3d through ezc3d (tools/mocap_interface.py:120-128), which is not
installable here; this module covers what that call site consumes: ``points`` (F x N x 3, invalid
samples as NaN), ``POINT:RATE`` and ``POINT:LABELS``.  Format facts follow the public C3D
specification (the reference's unused tools/c3d.py:191-424,1118-1393 served as the format spec).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import numpy as np

_BLOCK = 512


def _group(gid: int, name: str, desc: str = '') -> bytes:
    nb, db = name.encode(), desc.encode()
    body = struct.pack('<bb', len(nb), -gid) + nb
    body += struct.pack('<h', 3 + len(db)) + struct.pack('<B', len(db)) + db
    return body


def _param(gid: int, name: str, dtype: int, dims: Sequence[int], data: bytes, desc: str = '') -> bytes:
    nb, db = name.encode(), desc.encode()
    payload = struct.pack('<bB', dtype, len(dims)) + bytes(dims) + data + struct.pack('<B', len(db)) + db
    return struct.pack('<bb', len(nb), gid) + nb + struct.pack('<h', 2 + len(payload)) + payload


def write_c3d(fname: str, markers: np.ndarray, labels: List[str], frame_rate: float = 120.0,
              units: str = 'mm') -> None:
    """markers: F x N x 3 (NaN = missing) in ``units``; written as float32 frames."""
    markers = np.asarray(markers, dtype=np.float64)
    F, N, _ = markers.shape
    if N > 255 or F > 65535:
        raise ValueError('minimal writer: at most 255 points and 65535 frames')
    lab_len = max(4, max(len(l) for l in labels))
    lab_bytes = b''.join(l.encode().ljust(lab_len) for l in labels)
    recs = _group(1, 'POINT', 'point data')
    recs += _param(1, 'USED', 2, [], struct.pack('<h', N))
    recs += _param(1, 'FRAMES', 2, [], struct.pack('<H', F))
    recs += _param(1, 'SCALE', 4, [], struct.pack('<f', -1.0))
    recs += _param(1, 'RATE', 4, [], struct.pack('<f', float(frame_rate)))
    recs += _param(1, 'UNITS', -1, [len(units)], units.encode())
    recs += _param(1, 'LABELS', -1, [lab_len, N], lab_bytes)
    n_param_blocks = (4 + len(recs) + 64 + _BLOCK - 1) // _BLOCK
    data_start = 2 + n_param_blocks
    recs += _param(1, 'DATA_START', 2, [], struct.pack('<h', data_start))
    recs += struct.pack('<bb', 0, 0)
    psec = struct.pack('<BBBB', 1, 80, n_param_blocks, 84) + recs
    assert len(psec) <= n_param_blocks * _BLOCK
    psec = psec.ljust(n_param_blocks * _BLOCK, b'\0')

    hdr = struct.pack('<BBHHHHHfHHf', 2, 0x50, N, 0, 1, F, 0, -1.0, data_start, 0, float(frame_rate))
    hdr = hdr.ljust(_BLOCK, b'\0')

    valid = ~np.isnan(markers).any(-1)
    frames = np.zeros((F, N, 4), dtype='<f4')
    frames[..., :3] = np.where(valid[..., None], markers, 0.0)
    frames[..., 3] = np.where(valid, 0.0, -1.0)
    data = frames.tobytes()
    data = data.ljust(((len(data) + _BLOCK - 1) // _BLOCK) * _BLOCK, b'\0')
    with open(fname, 'wb') as f:
        f.write(hdr + psec + data)


def _read_params(buf: bytes):
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
        (offset,) = struct.unpack_from('<h', buf, off_pos)
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
    pblock, magic, n_pts, n_analog, first, last, _gap, scale, data_block, _apf, rate = \
        struct.unpack_from('<BBHHHHHfHHf', raw, 0)
    if magic != 0x50:
        raise ValueError(f'{fname}: not a C3D file')
    pstart = (pblock - 1) * _BLOCK
    if raw[pstart + 3] != 84:
        raise ValueError(f'{fname}: only Intel-format C3D files are supported by the minimal reader')
    n_pblocks = raw[pstart + 2]
    P = _read_params(raw[pstart: pstart + n_pblocks * _BLOCK])

    def scalar(key, fmt, default):
        if key in P:
            return struct.unpack_from(fmt, P[key][2], 0)[0]
        return default

    n_pts = scalar(('POINT', 'USED'), '<h', n_pts)
    n_frames = scalar(('POINT', 'FRAMES'), '<H', last - first + 1)
    scale = scalar(('POINT', 'SCALE'), '<f', scale)
    rate = scalar(('POINT', 'RATE'), '<f', rate)
    data_block = scalar(('POINT', 'DATA_START'), '<h', data_block)
    labels: List[str] = []
    if ('POINT', 'LABELS') in P:
        _, dims, blob = P[('POINT', 'LABELS')]
        ll, cnt = dims[0], (dims[1] if len(dims) > 1 else 1)
        labels = [blob[i * ll:(i + 1) * ll].decode('latin-1').strip() for i in range(cnt)]
    ofs = (data_block - 1) * _BLOCK
    if scale < 0:
        stride = 4 * (4 * n_pts + n_analog)
        arr = np.frombuffer(raw, dtype='<f4', count=n_frames * stride // 4, offset=ofs)
        arr = arr.reshape(n_frames, -1)[:, :4 * n_pts].reshape(n_frames, n_pts, 4).astype(np.float64)
        pts, resid = arr[..., :3], arr[..., 3]
    else:
        stride = 2 * (4 * n_pts + n_analog)
        arr = np.frombuffer(raw, dtype='<i2', count=n_frames * stride // 2, offset=ofs)
        arr = arr.reshape(n_frames, -1)[:, :4 * n_pts].reshape(n_frames, n_pts, 4).astype(np.float64)
        pts, resid = arr[..., :3] * scale, arr[..., 3]
    pts = np.where((resid < 0)[..., None], np.nan, pts)
    return pts, labels, float(rate)
