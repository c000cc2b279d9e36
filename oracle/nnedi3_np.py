"""TEST INFRASTRUCTURE ONLY -- NumPy (float32) restatement of the NNEDI3 shader math.

Follows ``nnedi3-nns16-win8x4.hook:20-104`` (double_y + combine_y) and ``:110-194`` (double_x +
combine_x): window mean / variance, per-neuron ``exp(W1.x * inv + b1)`` softmax weights times the
elliott-activated ``W2.x * inv + b2`` predictions, ``clamp(mean + 5 * vsum / wsum * sd, 0, 1)``.
The per-neuron dot products are restated as one ``[pixels x K] @ [K x nns]`` contraction per
weight set (<= 5e-6 from the literal neuron-serial order).

PARITY PIN: checked against ``oracle/glsl_exec.py`` (literal execution of the shader text) in
``tests/test_oracle.py`` with tolerance 2e-5; no reference golden vectors exist
(parity against the reference's own outputs is UNPINNED).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline leg may import this.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

F32 = np.float32
EPS = F32(1.192092896e-7)


def _windows(img: np.ndarray, S: int, direction: str) -> np.ndarray:
    """[H, W, 8, S] windows: index a = long axis offset -3..4, b = short axis offset -(S/2-1)..S/2."""
    H, W = img.shape
    pad = 4
    p = np.pad(img, pad, mode="edge")
    out = np.empty((H, W, 8, S), F32)
    for a in range(8):
        for b in range(S):
            lng, sht = a - 3, b - (S // 2 - 1)
            dx, dy = (lng, sht) if direction == "y" else (sht, lng)
            out[:, :, a, b] = p[pad + dy : pad + dy + H, pad + dx : pad + dx + W]
    return out


def predict(img: np.ndarray, nn, direction: str, chunk_rows: int = 64) -> np.ndarray:
    """The interpolated plane of one pass: value between y and y+1 (direction 'y') or x and x+1."""
    img = np.asarray(img, dtype=F32)
    nns, _, S = nn.w1.shape
    K = 8 * S
    W1 = np.ascontiguousarray(nn.w1.reshape(nns, K).T)  # [K, nns]
    W2 = np.ascontiguousarray(nn.w2.reshape(nns, K).T)
    H, W = img.shape
    out = np.empty((H, W), F32)
    win_all = _windows(img, S, direction).reshape(H, W, K)
    for y0 in range(0, H, chunk_rows):
        x = win_all[y0 : y0 + chunk_rows].reshape(-1, K)
        s = x.sum(axis=1, dtype=F32)
        sq = (x * x).sum(axis=1, dtype=F32)
        m0 = s / F32(K)
        m1 = sq / F32(K) - m0 * m0
        with np.errstate(all="ignore"):
            m2 = np.where(m1 >= EPS, F32(1.0) / np.sqrt(m1), F32(0.0)).astype(F32)
        m1 = m1 * m2
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            s1 = np.exp((x @ W1) * m2[:, None] + nn.b1[None, :])
            s2 = (x @ W2) * m2[:, None] + nn.b2[None, :]
            wsum = s1.sum(axis=1, dtype=F32)
            vsum = (s1 * (s2 / (F32(1.0) + np.abs(s2)))).sum(axis=1, dtype=F32)
            r = m0 + F32(5.0) * vsum / wsum * m1
        out[y0 : y0 + chunk_rows] = np.clip(r, F32(0.0), F32(1.0)).reshape(-1, W)
    return out


def nnedi3(img: np.ndarray, v, double_y: bool = True, double_x: bool = True) -> Tuple[np.ndarray, Tuple[float, float]]:
    """Full application: [H, W] -> [2H, 2W] (per-axis WHEN handled by the flags)."""
    cur = np.asarray(img, dtype=F32)
    off = [0.0, 0.0]
    if double_y:
        interp = predict(cur, v.nn_y, "y")
        H, W = cur.shape
        nxt = np.empty((2 * H, W), F32)
        nxt[0::2] = cur  # out(x, 2y) = in(x, y); out(x, 2y+1) = interp   (nnedi3-nns16-win8x4.hook:97-104)
        nxt[1::2] = interp
        cur = nxt
        off[1] -= 0.5
    if double_x:
        interp = predict(cur, v.nn_x, "x")
        H, W = cur.shape
        nxt = np.empty((H, 2 * W), F32)
        nxt[:, 0::2] = cur
        nxt[:, 1::2] = interp
        cur = nxt
        off[0] -= 0.5
    return cur, (off[0], off[1])
