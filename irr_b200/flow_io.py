"""On-disk formats either side of the hot path: Middlebury ``.flo`` and KITTI 16-bit PNG flow.

Same functions and conventions as the reference's readers / writers (datasets/common.py:19-27, utils/flow.py:11-62,
datasets/kitti_combined.py:19-34); the PNG codec is a self-contained 16-bit RGB reader / writer on zlib (the
reference depends on pypng, which is not needed here).  Host-side byte shuffling: numpy, no GPU involved.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

TAG_FLOAT = 202021.25  # b"PIEH" read as little-endian float32 (utils/flow.py:7)


def read_flo_as_float32(filename) -> np.ndarray:
    """(H, W, 2) float32, u then v interleaved per pixel (datasets/common.py:19-27)."""
    with open(filename, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != np.float32(TAG_FLOAT):
            raise ValueError("Magic number incorrect. Invalid .flo file")
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * h * w)
    if data.size != 2 * h * w:
        raise ValueError("truncated .flo file")
    return data.reshape(h, w, 2)


def write_flow(filename, uv, v=None) -> None:
    """utils/flow.py:11-34: header PIEH, int32 width, int32 height, then rows of interleaved (u, v) float32."""
    if v is None:
        uv = np.asarray(uv)
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(uv), np.asarray(v)
    assert u.shape == v.shape
    h, w = u.shape
    out = np.empty((h, w, 2), np.float32)
    out[:, :, 0] = u
    out[:, :, 1] = v
    with open(filename, "wb") as f:
        f.write(np.array([TAG_FLOAT], np.float32).tobytes())
        f.write(struct.pack("<ii", w, h))
        f.write(out.tobytes())


# ---------------------------------------------------------------------------------------------------- 16-bit PNG
_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png16_rgb(filename, arr: np.ndarray) -> None:
    """arr: (H, W, 3) uint16 -> 16-bit truecolour PNG (filter 0 on every row)."""
    arr = np.ascontiguousarray(arr, dtype=np.uint16)
    h, w, c = arr.shape
    assert c == 3
    raw = np.empty((h, 1 + w * 6), np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = arr.astype(">u2").view(np.uint8).reshape(h, w * 6)
    with open(filename, "wb") as f:
        f.write(_PNG_SIG)
        f.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)))
        f.write(_chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(_chunk(b"IEND", b""))


def _unfilter(raw: np.ndarray, h: int, stride: int, bpp: int) -> np.ndarray:
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    pos = 0
    for y in range(h):
        ft = raw[pos]
        line = raw[pos + 1:pos + 1 + stride].astype(np.int32)
        pos += 1 + stride
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            cur = line.copy()
            for i in range(bpp, stride):
                cur[i] = (cur[i] + cur[i - bpp]) & 255
        elif ft == 3:
            cur = line.copy()
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                cur[i] = (cur[i] + ((a + prev[i]) >> 1)) & 255
        elif ft == 4:
            cur = line.copy()
            for i in range(stride):
                a = int(cur[i - bpp]) if i >= bpp else 0
                b = int(prev[i])
                c = int(prev[i - bpp]) if i >= bpp else 0
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pr = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (cur[i] + pr) & 255
        else:
            raise ValueError(f"bad PNG filter type {ft}")
        out[y] = cur
        prev = cur
    return out


def read_png16_rgb(filename) -> np.ndarray:
    """16-bit truecolour, non-interlaced PNG -> (H, W, 3) uint16 (all five filter types)."""
    data = open(filename, "rb").read()
    if data[:8] != _PNG_SIG:
        raise ValueError("not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        (n,) = struct.unpack(">I", data[pos:pos + 4])
        tag = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    if (depth, ctype, interlace) != (16, 2, 0):
        raise ValueError("expected a 16-bit RGB non-interlaced PNG (KITTI flow format)")
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8)
    rows = _unfilter(raw, h, w * 6, 6)
    return rows.reshape(h, w * 3, 2).astype(np.uint16).dot(np.array([256, 1], np.uint16)).astype(np.uint16).reshape(h, w, 3)


def read_png_flow(filename):
    """datasets/kitti_combined.py:19-34: flow = (uint16 - 2**15) / 64 (float64), zero where the valid channel is 0;
    returns (flow (H, W, 2), valid (H, W, 1) int)."""
    img = read_png16_rgb(filename).astype(np.float64)
    invalid = img[:, :, 2] == 0
    flow = (img[:, :, 0:2] - 2 ** 15) / 64.0
    flow[invalid, 0] = 0
    flow[invalid, 1] = 0
    return flow, (1 - invalid * 1)[:, :, None]


def write_flow_png(filename, uv, v=None, mask=None) -> None:
    """utils/flow.py:37-62: u, v -> clip(x * 64 + 2**15, 0, 65535) as uint16 (truncation), third channel = valid mask."""
    if v is None:
        uv = np.asarray(uv)
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(uv), np.asarray(v)
    assert u.shape == v.shape
    h, w = u.shape
    valid = np.ones((h, w)) if mask is None else np.asarray(mask).reshape(h, w)
    fu = np.clip(u * 64 + 2 ** 15, 0.0, 65535.0).astype(np.uint16)
    fv = np.clip(v * 64 + 2 ** 15, 0.0, 65535.0).astype(np.uint16)
    write_png16_rgb(filename, np.stack((fu, fv, valid.astype(np.uint16)), axis=-1))
