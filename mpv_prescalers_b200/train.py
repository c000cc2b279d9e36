"""LUT trainer for the single-pass RAVU families, RAVU-Lite and RAVU-3x: the piece of the
absent upstream ``source`` branch (``README.md:3-4``) that produces a ``//!TEXTURE`` weight LUT from example images.

RAVU is "rapid and accurate" learned upscaling: the 2x2 output pixels of every source pixel are a linear filter of the
source window, with one filter per (angle, strength, coherence) bucket of the window's structure tensor
(``ravu-lite-r3.hook:15-120``).  Training = one least-squares problem per bucket: minimise, over all source pixels of the
training planes that fall into the bucket, the squared error between the filtered window and the true high-resolution
pixels.  The shipped LUTs are point-symmetric -- tap ``N-1-t`` of phase ``c`` carries the weight of tap ``t`` of phase
``3-c``, which is what lets the shader store half the taps and fetch ``w.wzyx`` for the mirrored one
(``ravu-lite-r3.hook:97-118``; verified on the shipped payloads: centre texel ``.x == .w``, ``.y == .z`` exactly) -- so
each bucket is solved for two weight vectors ``W_0, W_1`` on the data augmented with its 180-degree rotation, and
``W_3 = reverse(W_0)``, ``W_2 = reverse(W_1)``.  RAVU-3x has the same structure with a 3x3 block of output pixels: the
centre is the source pixel itself, the other eight are filters ``W_0..W_7`` in output order, stored as two texels per tap
(``res0`` / ``res1``), with ``W_{7-p} = reverse(W_p)`` (``compute/ravu-3x-r2.hook:84-114``: the mirrored tap reads
``w1.wzyx`` into ``res0`` and ``w0.wzyx`` into ``res1``), so four weight vectors are solved per bucket.

The three-pass RAVU (``ravu-rN.hook``) applies ONE LUT three times: to the source lattice (pass 1: the pixel at
(x+1/2, y+1/2)) and twice to the 45-degree lattice of source pixels and pass-1 results (passes 2 / 3: (x+1/2, y) and
(x, y+1/2); ``ravu-r2.hook:15-338``), with one weight per mirrored tap pair (``res += (s_k + s_{N-1-k}) * w_k``).
:func:`train_ravu_chain` pools the windows of all three passes per bucket; the windows and buckets of passes 2 / 3 are
those the hook itself produces with its current LUT (they contain its own pass-1 results, exactly as at run time), so one
call is one step of a fixed-point iteration -- ``rounds`` > 1 re-runs the hook with the LUT of the previous round.

RAVU-Zoom (``ravu-zoom-rN.hook``) stores, per bucket, a 9 x 9 grid of filters over the sub-pixel phase and lets the sampler
blend the four nodes around (8 sx, 8 sy) (``LUTPOS``, ``ravu-zoom-r2.hook:23,112-131``); the first half of the taps is
weighted at phase (sx, sy), the mirrored half at (1 - sx, 1 - sy).  The output is still linear in the node filters, so
:func:`train_ravu_zoom` solves one least-squares problem per bucket in the 81 * N/2 node weights, from training pairs at
several scale factors (the phases must cover the grid), regularised towards the hook's own LUT so that nodes no sample
reaches keep their values.  The anti-ringing LUTs (``ravu_zoom_lutN_ar``) are not trained: their objective is stated
nowhere in the reference.

Everything runs on the GPU: the buckets come from the same CUDA key kernel the hook uses at run time
(``prescale(..., return_buckets=True)``), the normal equations are accumulated per bucket in float64, and the result is
written back as a complete ``.hook`` file (the original GLSL, a new payload line), which ``prescale()`` accepts like a
shipped file.  The reference's own training set and hyper-parameters are not in the snapshot, so a LUT trained here is a
NEW filter, not a reconstruction of the shipped one; the self-consistency test trains on planes upscaled by a shipped LUT
and recovers that LUT.  The anti-ringing LUT of ravu-zoom (``ravu_zoom_lut3_ar``, ``.MISSING_LARGE_BLOBS``) has no
training objective stated anywhere in the reference and is not covered.
"""
from __future__ import annotations

import re
from typing import Optional, Tuple

import numpy as np
import torch

from .hookfile import HookError, HookFile

__all__ = ["train_ravu", "train_ravu_lite", "train_ravu_chain", "train_ravu_zoom", "write_hook_with_lut", "lut_to_hex"]


def _windows(lr: torch.Tensor, radius: int) -> torch.Tensor:
    """[H, W] -> [H*W, N] source windows in the shader's tap order t = i*n + j (i <-> dx, j <-> dy), clamp-to-edge."""
    n, o = 2 * radius - 1, radius - 1
    h, w = lr.shape
    p = torch.nn.functional.pad(lr[None, None], (o, o, o, o), mode="replicate")[0, 0]
    cols = []
    for i in range(n):          # dx = i - o
        for j in range(n):      # dy = j - o
            cols.append(p[j:j + h, i:i + w].reshape(-1))
    return torch.stack(cols, dim=1)


def train_ravu(hook: HookFile, lr: torch.Tensor, hr: torch.Tensor, ridge: float = 1e-9, min_samples: Optional[int] = None,
               exclude_clipped: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Least-squares LUT of a RAVU-Lite or (luma) RAVU-3x hook from training pairs.

    hook   a ``ravu-lite(-ar)-rN.hook`` or ``compute/ravu-3x-rN.hook``: its key constants decide the buckets, its LUT fills
           buckets that see too few samples;
    lr     ``[F, H, W]`` float32 CUDA planes in [0, 1];  hr  ``[F, sH, sW]`` (s = 2 or 3): the true planes, aligned like the
           shader's output (output pixel (i, j) of source pixel (x, y) is hr[s*y + j, s*x + i]; ``ravu-lite-r3.hook:128-134``,
           ``compute/ravu-3x-r2.hook:106-114``);
    ridge  Tikhonov term relative to the mean diagonal of the normal matrix;
    exclude_clipped  drop source pixels with a target at exactly 0 or 1 (the shader clamps its result to [0, 1], which
           makes such pixels uninformative about the linear filter).

    Returns ``(lut [rows, LW, 4] float32, samples_per_bucket [rows])``."""
    from .api import prescale

    v = hook.variant
    if v.family not in ("ravu-lite", "ravu-3x") or v.plane != "luma":
        raise HookError("train_ravu() trains the single-pass luma families: ravu-lite(-ar)-rN and ravu-3x-rN "
                        "(the -yuv / -rgb files of ravu-3x apply the same LUT per channel)")
    if lr.device.type != "cuda" or hr.device.type != "cuda":
        raise ValueError("training planes must be CUDA tensors")
    S = 2 if v.family == "ravu-lite" else 3
    f, h, w = lr.shape
    if tuple(hr.shape) != (f, S * h, S * w):
        raise ValueError(f"hr must be {(f, S * h, S * w)}, got {tuple(hr.shape)}")
    r = v.radius
    n = 2 * r - 1
    N, half = n * n, (n * n - 1) // 2
    rows = int(v.lut.height)
    # trained output pixels in the LUT's phase order p (lite: all four; 3x: the eight around the copied centre), q = i*S + j
    cells = [q for q in range(S * S) if not (S == 3 and q == 4)]
    P, U = len(cells), len(cells) // 2           # phases, unknown weight vectors (W_{P-1-p} = reverse(W_p))
    dev = lr.device
    A = torch.zeros((rows, N, N), dtype=torch.float64, device=dev)
    B = torch.zeros((rows, N, U), dtype=torch.float64, device=dev)
    count = torch.zeros(rows, dtype=torch.int64, device=dev)
    rev = torch.arange(N - 1, -1, -1, device=dev)
    out_size = None if S == 2 else (S * h, S * w)
    _, buckets = prescale(lr, hook, out_size, return_buckets=True)       # the hook's own key kernel
    for k in range(f):
        X = _windows(lr[k], r)                                   # [P, N]
        H = hr[k]
        Y = torch.stack([H[(q % S)::S, (q // S)::S] for q in cells], dim=-1).reshape(-1, P)
        b = buckets[k].reshape(-1).long()
        if exclude_clipped:
            keep = ((Y > 0.0) & (Y < 1.0)).all(dim=1)
            X, Y, b = X[keep], Y[keep], b[keep]
        order = torch.argsort(b)
        X, Y, b = X[order], Y[order], b[order]
        cnt = torch.bincount(b, minlength=rows)
        count += cnt
        starts = torch.cumsum(cnt, 0) - cnt
        for row in torch.nonzero(cnt).reshape(-1).tolist():
            s, m = int(starts[row]), int(cnt[row])
            xb = X[s:s + m].double()
            yb = Y[s:s + m].double()
            xr = xb[:, rev]                                      # the window rotated by 180 degrees
            A[row] += xb.T @ xb + xr.T @ xr
            # W_p sees (x, y_p) and (rev x, y_{P-1-p})
            B[row] += xb.T @ yb[:, :U] + xr.T @ yb[:, P - 1 - torch.arange(U, device=dev)]
    lut_old = np.asarray(v.lut.data, dtype=np.float32)           # [rows, LW, 4]
    lut = lut_old.copy()
    need = (4 * N) if min_samples is None else int(min_samples)
    count_h = count.cpu().numpy()
    eye = torch.eye(N, dtype=torch.float64, device=dev)
    for row in range(rows):
        if count_h[row] < need:
            continue                                             # too few samples: the hook's own row stays
        a = A[row]
        lam = ridge * float(torch.diagonal(a).mean())
        Wsol = torch.linalg.solve(a + lam * eye, B[row]).cpu().numpy()   # [N, U]
        Wfull = np.concatenate([Wsol, Wsol[::-1, ::-1]], axis=1)         # [N, P]: W_{P-1-p}[t] = W_p[N-1-t]
        for t in range(half + 1):
            if S == 2:
                lut[row, t] = Wfull[t]
            else:                                                # two texels per tap: res0 = phases 0..3, res1 = phases 4..7
                lut[row, 2 * t] = Wfull[t, :4]
                lut[row, 2 * t + 1] = Wfull[t, 4:]
    return lut.astype(np.float32), count_h


def train_ravu_lite(hook: HookFile, lr: torch.Tensor, hr: torch.Tensor, ridge: float = 1e-9, min_samples: Optional[int] = None,
                    exclude_clipped: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """:func:`train_ravu` restricted to the RAVU-Lite family (the round-2 entry point; kept for callers)."""
    if hook.variant.family != "ravu-lite":
        raise HookError("train_ravu_lite() trains the RAVU-Lite family")
    return train_ravu(hook, lr, hr, ridge, min_samples, exclude_clipped)


def _chain_taps(r: int):
    """Tap positions of the three passes on the 2x OUTPUT grid, relative to (2x, 2y): a list per pass of ``(dx2, dy2)`` in
    tap order t = i*n + j.  Pass 1 taps HOOKED at (x + i - (r-1), y + j - (r-1)); passes 2 / 3 tap the lattice point at twice
    the real position ``(tx2 - (2r-1) + i + j, ty2 - i + j)`` -- even = HOOKED, odd = the saved pass-1 texture
    (``ravu-r2.hook:137-180``: the 45-degree lattice, i along (1/2, -1/2), j along (1/2, 1/2))."""
    n = 2 * r
    p1 = [(2 * (t // n - (r - 1)), 2 * (t % n - (r - 1))) for t in range(n * n)]
    out = [p1]
    for tx2, ty2 in ((1, 0), (0, 1)):
        out.append([(tx2 - (2 * r - 1) + t // n + t % n, ty2 - t // n + t % n) for t in range(n * n)])
    return out


def _gather_lattice(out2: torch.Tensor, dx2: int, dy2: int) -> torch.Tensor:
    """Sample of every source pixel (x, y) at output-grid offset (dx2, dy2) from (2x, 2y), clamp-to-edge applied per texture
    like the host does: even offsets address HOOKED, odd ones the half-resolution pass-1 texture."""
    h2, w2 = out2.shape
    h, w = h2 // 2, w2 // 2
    par = dx2 & 1                                                 # dy2 has the same parity by construction
    ys = torch.arange(h, device=out2.device) + (dy2 - par) // 2
    xs = torch.arange(w, device=out2.device) + (dx2 - par) // 2
    ys = ys.clamp_(0, h - 1) * 2 + par
    xs = xs.clamp_(0, w - 1) * 2 + par
    return out2[ys[:, None], xs[None, :]]


def train_ravu_chain(hook: HookFile, lr: torch.Tensor, hr: torch.Tensor, ridge: float = 1e-9, min_samples: Optional[int] = None,
                     exclude_clipped: bool = True, rounds: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Least-squares LUT of a three-pass luma RAVU hook (``ravu-rN.hook``) from training pairs.

    lr ``[F, H, W]``, hr ``[F, 2H, 2W]`` float32 CUDA planes; hr is aligned like the hook's output: hr[2y, 2x] is the
    source pixel, hr[2y+1, 2x+1] / hr[2y, 2x+1] / hr[2y+1, 2x] the targets of passes 1 / 2 / 3 (``ravu-r2.hook:327-338``).
    Returns ``(lut [rows, LW, 4] float32, samples_per_bucket [rows])`` (all three passes pooled)."""
    import os
    import tempfile

    from .api import prescale

    v = hook.variant
    if v.family != "ravu" or v.plane != "luma":
        raise HookError("train_ravu_chain() trains the three-pass luma RAVU hooks (ravu-rN.hook)")
    if lr.device.type != "cuda" or hr.device.type != "cuda":
        raise ValueError("training planes must be CUDA tensors")
    f, h, w = lr.shape
    if tuple(hr.shape) != (f, 2 * h, 2 * w):
        raise ValueError(f"hr must be {(f, 2 * h, 2 * w)}, got {tuple(hr.shape)}")
    r = v.radius
    N = (2 * r) ** 2
    half = N // 2
    rows = int(v.lut.height)
    taps = _chain_taps(r)
    tgt = ((1, 1), (1, 0), (0, 1))                                 # (dx, dy) of the pass's target inside the 2x2 cell
    dev = lr.device
    cur = hook
    lut = np.asarray(v.lut.data, dtype=np.float32).copy()
    count_h = np.zeros(rows, np.int64)
    tmpdir = tempfile.mkdtemp(prefix="mpvp_train_") if rounds > 1 else None
    for rnd in range(max(1, int(rounds))):
        A = torch.zeros((rows, half, half), dtype=torch.float64, device=dev)
        B = torch.zeros((rows, half), dtype=torch.float64, device=dev)
        count = torch.zeros(rows, dtype=torch.int64, device=dev)
        out, buckets = prescale(lr, cur, return_buckets=True)      # [F, 2H, 2W], [F, 3, H, W]: the hook's own lattice and keys
        for k in range(f):
            for ps in range(3):
                cols = [_gather_lattice(out[k], dx2, dy2).reshape(-1) for dx2, dy2 in taps[ps]]
                X = torch.stack([cols[t] + cols[N - 1 - t] for t in range(half)], dim=1)     # one weight per mirrored pair
                y = hr[k][tgt[ps][1]::2, tgt[ps][0]::2].reshape(-1)
                b = buckets[k, ps].reshape(-1).long()
                if exclude_clipped:
                    keep = (y > 0.0) & (y < 1.0)
                    X, y, b = X[keep], y[keep], b[keep]
                order = torch.argsort(b)
                X, y, b = X[order], y[order], b[order]
                cnt = torch.bincount(b, minlength=rows)
                count += cnt
                starts = torch.cumsum(cnt, 0) - cnt
                for row in torch.nonzero(cnt).reshape(-1).tolist():
                    s0, m = int(starts[row]), int(cnt[row])
                    xb, yb = X[s0:s0 + m].double(), y[s0:s0 + m].double()
                    A[row] += xb.T @ xb
                    B[row] += xb.T @ yb
        need = (8 * half) if min_samples is None else int(min_samples)
        count_h = count.cpu().numpy()
        eye = torch.eye(half, dtype=torch.float64, device=dev)
        flat = lut.reshape(rows, -1).copy()                        # weight k of a row sits at texel k // 4, component k % 4
        for row in range(rows):
            if count_h[row] < need:
                continue
            lam = ridge * float(torch.diagonal(A[row]).mean())
            flat[row, :half] = torch.linalg.solve(A[row] + lam * eye, B[row]).cpu().numpy()
        lut = flat.reshape(lut.shape).astype(np.float32)
        if rnd + 1 < rounds:
            path = os.path.join(tmpdir, f"round{rnd}.hook")
            write_hook_with_lut(hook, lut, path)
            cur = HookFile.parse(path)
    return lut, count_h


def _zoom_axis(in_size: int, out_size: int, dev):
    """Base texel, node index and node blend weight of every output coordinate (``pos = (o + 1/2) / O * I`` in float32 like
    the shader; node coordinate 8 * sub)."""
    o = torch.arange(out_size, dtype=torch.float32, device=dev)
    pos = ((o + 0.5) / float(out_size)) * float(in_size)
    t = pos - 0.5
    sub = t - torch.floor(t)
    base = torch.floor(pos - sub).long()
    return base, sub.double()


def _node_basis(t: torch.Tensor):
    """(i0, i1, w0, w1): the two LUT nodes around 8 t and their blend weights (t in [0, 1], node 8 reached only at t = 1)."""
    u = (8.0 * t).clamp(0.0, 8.0)
    i0 = torch.floor(u).clamp(max=8.0)
    f = u - i0
    i0 = i0.long()
    return i0, (i0 + 1).clamp(max=8), 1.0 - f, f


def train_ravu_zoom(hook: HookFile, pairs, ridge: float = 1e-6, min_samples: Optional[int] = None,
                    exclude_clipped: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Least-squares LUT of a luma RAVU-Zoom hook (``ravu-zoom-rN.hook``, not the -ar files).

    pairs  a list of ``(lr [F, H, W], hr [F, OH, OW])`` float32 CUDA planes, ONE scale factor per pair and several pairs:
           hr[oy, ox] is the true value at the position the hook computes for output pixel (ox, oy) (centre-aligned,
           ``pos = (o + 1/2) * I / O``);
    ridge  pulls the solution towards the hook's own LUT (relative to the mean diagonal of the normal matrix): nodes that
           no training phase reaches keep their values.

    Returns ``(lut [288 * 9, B * 9, 4] float32, samples_per_bucket [288])``."""
    from .api import prescale

    v = hook.variant
    if v.family != "ravu-zoom" or v.plane != "luma" or v.ar:
        raise HookError("train_ravu_zoom() trains the luma RAVU-Zoom hooks without anti-ringing (ravu-zoom-rN.hook)")
    r = v.radius
    n = 2 * r
    N, H2 = n * n, n * n // 2
    B = (H2 + 3) // 4
    rows = int(v.lut.height) // 9
    D = 81 * H2
    lut_old = np.asarray(v.lut.data, dtype=np.float32)                       # [rows * 9, B * 9, 4]
    if lut_old.shape != (rows * 9, B * 9, 4):
        raise HookError(f"unexpected LUT geometry {lut_old.shape}")
    # node-major weight vector of a bucket: W[ny, nx, k], k = blk * 4 + comp
    w_old = lut_old.reshape(rows, 9, B, 9, 4).transpose(0, 1, 3, 2, 4).reshape(rows, 9, 9, B * 4)[..., :H2]
    dev = pairs[0][0].device
    W0 = torch.from_numpy(np.ascontiguousarray(w_old)).to(dev).double().reshape(rows, D)
    A = torch.zeros((rows, D, D), dtype=torch.float64, device=dev)
    G = torch.zeros((rows, D), dtype=torch.float64, device=dev)               # X^T (y - X w_old)
    count = torch.zeros(rows, dtype=torch.int64, device=dev)
    for lr, hr in pairs:
        if lr.device.type != "cuda" or hr.device.type != "cuda":
            raise ValueError("training planes must be CUDA tensors")
        f, h, w = lr.shape
        oh, ow = int(hr.shape[1]), int(hr.shape[2])
        if hr.shape[0] != f or oh <= h or ow <= w:
            raise ValueError("every pair needs hr planes larger than its lr planes on both axes")
        _, buckets = prescale(lr, hook, output_size=(oh, ow), return_buckets=True)      # [F, OH, OW]
        bx, sx = _zoom_axis(w, ow, dev)
        by, sy = _zoom_axis(h, oh, dev)
        nodes = {}
        for tag, tx, ty in (("d", sx, sy), ("m", 1.0 - sx, 1.0 - sy)):
            x0, x1, wx0, wx1 = _node_basis(tx)
            y0, y1, wy0, wy1 = _node_basis(ty)
            nodes[tag] = [((ya[:, None] * 9 + xa[None, :]).reshape(-1), (wya[:, None] * wxa[None, :]).reshape(-1))
                          for ya, wya in ((y0, wy0), (y1, wy1)) for xa, wxa in ((x0, wx0), (x1, wx1))]
        for k in range(f):
            src = lr[k]
            taps = []
            for t in range(N):                                               # t = i * n + j, i <-> dx, j <-> dy
                yy = (by + (t % n - (r - 1))).clamp(0, h - 1)
                xx = (bx + (t // n - (r - 1))).clamp(0, w - 1)
                taps.append(src[yy[:, None], xx[None, :]].reshape(-1))
            S = torch.stack(taps, dim=1).double()                            # [P, N]
            Sd, Sm = S[:, :H2], S[:, torch.arange(N - 1, N - 1 - H2, -1, device=dev)]
            y = hr[k].reshape(-1).double()
            b = buckets[k].reshape(-1).long()
            keep = ((y > 0.0) & (y < 1.0)) if exclude_clipped else torch.ones_like(b, dtype=torch.bool)
            order = torch.argsort(b[keep])
            sel = torch.nonzero(keep).reshape(-1)[order]
            bs = b[sel]
            cnt = torch.bincount(bs, minlength=rows)
            count += cnt
            starts = torch.cumsum(cnt, 0) - cnt
            for row in torch.nonzero(cnt).reshape(-1).tolist():
                ids = sel[int(starts[row]):int(starts[row]) + int(cnt[row])]
                m = ids.numel()
                X = torch.zeros((m, 81, H2), dtype=torch.float64, device=dev)
                ar = torch.arange(m, device=dev)
                for tag, Sv in (("d", Sd), ("m", Sm)):
                    for idx, wgt in nodes[tag]:
                        X[ar, idx[ids]] += wgt[ids][:, None] * Sv[ids]
                X = X.reshape(m, D)
                A[row] += X.T @ X
                G[row] += X.T @ (y[ids] - X @ W0[row])
    need = (4 * H2) if min_samples is None else int(min_samples)
    count_h = count.cpu().numpy()
    eye = torch.eye(D, dtype=torch.float64, device=dev)
    Wn = W0.clone()
    for row in range(rows):
        if count_h[row] < need:
            continue
        lam = ridge * float(torch.diagonal(A[row]).mean())
        Wn[row] = W0[row] + torch.linalg.solve(A[row] + lam * eye, G[row])
    full = np.zeros((rows, 9, 9, B * 4), np.float32)
    full[..., :H2] = Wn.reshape(rows, 9, 9, H2).cpu().numpy()
    full[..., H2:] = lut_old.reshape(rows, 9, B, 9, 4).transpose(0, 1, 3, 2, 4).reshape(rows, 9, 9, B * 4)[..., H2:]
    lut = full.reshape(rows, 9, 9, B, 4).transpose(0, 1, 3, 2, 4).reshape(rows * 9, B * 9, 4)
    return np.ascontiguousarray(lut, dtype=np.float32), count_h


def lut_to_hex(lut: np.ndarray) -> str:
    """``[h, w, 4]`` float32 -> the payload line of a ``//!TEXTURE`` block (little-endian float32, lowercase hex)."""
    return np.ascontiguousarray(lut, dtype="<f4").tobytes().hex()


def write_hook_with_lut(hook: HookFile, lut: np.ndarray, path: str) -> None:
    """Write a complete hook file: the text of ``hook`` with the payload of its (single) LUT replaced by ``lut``."""
    v = hook.variant
    if lut.shape != (v.lut.height, v.lut.width, 4):
        raise ValueError(f"LUT must be {(v.lut.height, v.lut.width, 4)}, got {lut.shape}")
    with open(hook.path) as f:
        text = f.read()
    pat = re.compile(r"(//!TEXTURE " + re.escape(v.lut.name) + r"\n(?://![^\n]*\n)+)([0-9a-f]+)")
    if len(pat.findall(text)) != 1:
        raise HookError(f"{hook.path}: cannot locate the payload of {v.lut.name}")
    with open(path, "w") as f:
        f.write(pat.sub(lambda m: m.group(1) + lut_to_hex(lut), text))
