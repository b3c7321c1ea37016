"""CPU restatement of the patch resize of the reference's loader (TEST INFRASTRUCTURE -- never imported by the
product).  ExtractFeatureDataset.resize_data (MyUtils2.py:362-376) calls, per band,
`cv2.resize(band_u8, (t, t), interpolation=cv2.INTER_AREA)` and then `astype(float32) / 255`.

The arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV (unpinned by the
reference; the container's 4.13.0 is the de-facto pin).  Its published algorithm for 8-bit single-channel input,
restated here and pinned bit-exactly against cv2 itself (tests/test_oracle_golden.py, tests/golden/resize.npz):

  * integer shrink factor k = s / t ("ResizeAreaFast"): k == 1 copies; k == 2 is (sum of the 2x2 block + 2) >> 2;
    any other k is the integer block sum times float32(1 / k^2), rounded half to even;
  * fractional shrink ("ResizeArea"): per axis a table of (dst, src, weight) with float32 weights
    (partial left cell, whole cells 1 / cellWidth, partial right cell); rows are reduced horizontally into float32
    buffers in table order, then vertically, then rounded half to even;
  * enlargement (INTER_AREA falls back to its "area-mode" bilinear): sx = floor(dx * s / t),
    fx = (dx + 1) - (sx + 1) * t / s (clamped to 0, fractional part), 11-bit fixed-point coefficients, and the
    fixed-point vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def _area_table(ssize, dsize):
    scale = ssize / dsize
    di, si, al = [], [], []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            di.append(dx); si.append(sx1 - 1); al.append(f32((sx1 - fsx1) / cell))
        for sx in range(sx1, sx2):
            di.append(dx); si.append(sx); al.append(f32(1.0 / cell))
        if fsx2 - sx2 > 1e-3:
            di.append(dx); si.append(sx2); al.append(f32(min(min(fsx2 - sx2, 1.0), cell) / cell))
    return di, si, al


def _shrink_fractional(a, t):
    H, W = a.shape
    xd, xs, xa = _area_table(W, t)
    yd, ys, ya = _area_table(H, t)
    buf = np.zeros((H, t), f32)
    for d, s, w in zip(xd, xs, xa):                       # sequential float32 multiply-adds, table order
        buf[:, d] = buf[:, d] + a[:, s].astype(f32) * w
    out = np.zeros((t, t), f32)
    prev, acc = -1, None
    for d, s, w in zip(yd, ys, ya):
        if d != prev:
            if prev >= 0:
                out[prev] = acc
            acc, prev = w * buf[s], d
        else:
            acc = acc + w * buf[s]
    out[prev] = acc
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def _linear_table(ssize, dsize):
    inv = dsize / ssize
    scale = 1.0 / inv
    sx, a0, a1, xmax = [], [], [], dsize
    for dx in range(dsize):
        s = math.floor(dx * scale)
        fx = f32((dx + 1) - (s + 1) * inv)
        fx = f32(0) if fx <= 0 else f32(fx - math.floor(fx))
        if s < 0:
            fx, s = f32(0), 0
        if s + 1 >= ssize:
            xmax = min(xmax, dx)
            if s >= ssize - 1:
                fx, s = f32(0), ssize - 1
        sx.append(s)
        a0.append(int(np.clip(np.rint((f32(1) - fx) * f32(2048)), -32768, 32767)))
        a1.append(int(np.clip(np.rint(fx * f32(2048)), -32768, 32767)))
    return np.asarray(sx), np.asarray(a0, np.int64), np.asarray(a1, np.int64), xmax


def _enlarge(a, t):
    H, W = a.shape
    sx, a0, a1, xmax = _linear_table(W, t)
    sy, b0, b1, _ = _linear_table(H, t)
    A = a.astype(np.int64)
    rows = A[:, sx] * a0 + A[:, np.minimum(sx + 1, W - 1)] * a1
    edge = np.arange(t) >= xmax
    rows[:, edge] = A[:, sx[edge]] * 2048
    S0, S1 = rows[sy], rows[np.minimum(sy + 1, H - 1)]
    out = (((b0[:, None] * (S0 >> 4)) >> 16) + ((b1[:, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area_u8(a, t):
    """cv2.resize(a, (t, t), interpolation=cv2.INTER_AREA) for a square uint8 [s, s] band."""
    a = np.ascontiguousarray(a, np.uint8)
    s = a.shape[0]
    if a.ndim != 2 or a.shape[1] != s:
        raise ValueError("square single-band patches only (what the reference's cutter produces)")
    if s % t == 0:
        k = s // t
        if k == 1:
            return a.copy()
        blk = a.reshape(t, k, t, k).astype(np.int64).sum(axis=(1, 3))
        if k == 2:
            return ((blk + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(blk.astype(f32) * f32(1.0 / (k * k))), 0, 255).astype(np.uint8)
    return _shrink_fractional(a, t) if s > t else _enlarge(a, t)


def resize_data(data_chw, t):
    """ExtractFeatureDataset.resize_data, MyUtils2.py:362-376: uint8 [C, s, s] -> float32 [C, t, t] in [0, 1]."""
    return np.stack([resize_area_u8(b, t) for b in data_chw]).astype(np.float32) / np.float32(255.0)
