"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the offset-correcting main scaler.

The reference declares ``//!OFFSET -0.5 -0.5`` for ravu and nnedi3 (``ravu-r2.hook:325``,
``nnedi3-nns16-win8x4.hook:95,185``) and leaves the correction to the host's main scaler.  That scaler lives in mpv
(``video/out/filter_kernels.c``, ``video/out/gpu/video.c``: third-party code that is NOT in the reference snapshot; no
version is pinned by the reference), so this file restates its published algorithm -- separable polyphase resampling with
the named filter kernels, weights normalised per output coordinate, clamp-to-edge, no kernel widening when downscaling
(mpv's default ``--correct-downscaling=no``) -- and the geometry the hook files imply:

    s(o) = (o + 0.5) * I / O - 0.5 + offset,   taps at floor(s) - R + 1 ... floor(s) + R,   w_k = K(k - s) / sum

PARITY PIN: none (no golden vectors exist for this step); checked by properties in tests/test_resample.py (constants are
preserved, integer offsets shift exactly, ravu + correction lands on ravu-lite's sample grid).

Only ``tests/`` may import this.
"""
from __future__ import annotations

import numpy as np

RADIUS = {"bilinear": 1, "catmull_rom": 2, "mitchell": 2, "spline36": 3, "lanczos": 3}


def _cubic(x, B, C):
    x = np.abs(x)
    a = ((12 - 9 * B - 6 * C) * x**3 + (-18 + 12 * B + 6 * C) * x**2 + (6 - 2 * B)) / 6.0
    b = ((-B - 6 * C) * x**3 + (6 * B + 30 * C) * x**2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) / 6.0
    return np.where(x < 1, a, np.where(x < 2, b, 0.0))


def _sinc(x):
    px = np.pi * x
    with np.errstate(all="ignore"):
        return np.where(np.abs(x) < 1e-8, 1.0, np.sin(px) / np.where(px == 0, 1.0, px))


def _spline36(x):
    x = np.abs(x)
    a = ((13.0 / 11.0 * x - 453.0 / 209.0) * x - 3.0 / 209.0) * x + 1.0
    y = x - 1.0
    b = ((-6.0 / 11.0 * y + 270.0 / 209.0) * y - 156.0 / 209.0) * y
    z = x - 2.0
    c = ((1.0 / 11.0 * z - 45.0 / 209.0) * z + 26.0 / 209.0) * z
    return np.where(x < 1, a, np.where(x < 2, b, np.where(x < 3, c, 0.0)))


def kernel(name: str, x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float64)
    if name == "bilinear":
        return np.maximum(0.0, 1.0 - np.abs(x))
    if name == "catmull_rom":
        return _cubic(x, 0.0, 0.5)
    if name == "mitchell":
        return _cubic(x, 1.0 / 3.0, 1.0 / 3.0)
    if name == "spline36":
        return _spline36(x)
    if name == "lanczos":
        return np.where(np.abs(x) < 3.0, _sinc(x) * _sinc(x / 3.0), 0.0)
    raise ValueError(name)


def axis_table(I: int, O: int, offset: float, name: str):
    """(base [O] int, weights [O, 2R] float32) of one axis."""
    R = RADIUS[name]
    o = np.arange(O, dtype=np.float64)
    s = (o + 0.5) * (float(I) / float(O)) - 0.5 + float(np.float32(offset))
    base = np.floor(s).astype(np.int64) - R + 1
    k = base[:, None] + np.arange(2 * R)[None, :]
    w = kernel(name, k - s[:, None])
    w = w / w.sum(axis=1, keepdims=True)
    return base, w.astype(np.float32)


def resample(img: np.ndarray, out_size=None, offset=(0.0, 0.0), name: str = "lanczos") -> np.ndarray:
    """[H, W] -> [OH, OW] float32; out_size=(OH, OW); offset=(x, y) in input texels (the accumulated //!OFFSET)."""
    img = np.asarray(img, np.float32)
    H, W = img.shape
    OH, OW = (H, W) if out_size is None else out_size
    bx, wx = axis_table(W, OW, offset[0], name)
    by, wy = axis_table(H, OH, offset[1], name)
    T = wx.shape[1]
    rows = np.zeros((H, OW), np.float32)
    for i in range(T):  # same tap order as the device: fp32 accumulation of w * sample, tap by tap
        rows = rows + img[:, np.clip(bx + i, 0, W - 1)] * wx[None, :, i]
    out = np.zeros((OH, OW), np.float32)
    for j in range(T):
        out = out + rows[np.clip(by + j, 0, H - 1), :] * wy[:, j, None]
    return out
