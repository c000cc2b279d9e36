"""TEST INFRASTRUCTURE ONLY -- parametric NumPy (float32) restatement of the RAVU shader math.

Follows, operation by operation, the GLSL of the reference's root variants:

* key (structure tensor -> eigen -> angle/strength/coherence -> LUT row):
  ``ravu-lite-ar-r3.hook:48-88`` (lite / 3x stencils), ``ravu-r3.hook:58-119`` (ravu / zoom
  4th-order stencils), ``ravu-r2.hook:97`` (log2 strength), ``compute/ravu-3x-r2.hook:79``;
* ravu-lite convolution, anti-ringing and phase write: ``ravu-lite-ar-r3.hook:89-197``;
* ravu 3-convolution chain and merge: ``ravu-r2.hook:15-338`` (rgb: ``ravu-r2-rgb.hook:21-131``);
* ravu-zoom: ``ravu-zoom-r2.hook:23-134``, AR ``ravu-zoom-ar-r2.hook:24-208``;
* ravu-3x: ``compute/ravu-3x-r2.hook:15-115``.

PARITY PIN: ``tests/test_oracle.py`` checks this file against ``oracle/glsl_exec.py``
(literal execution of the shader text) -- bit-exact for every RAVU family.  There are no golden
vectors in the reference and it cannot be run here, so against the reference's own outputs parity is
UNPINNED.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline leg may import this.
Parameters (Gaussian weights, thresholds, LUT payloads ...) are passed in as a duck-typed
``variant`` object (``mpv_prescalers_b200.hookfile.Variant`` in the tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32
EPS = F32(1.192092896e-7)
PI = F32(3.141592653589793)
COLOR_PRIMARY = (F32(0.2126), F32(0.7152), F32(0.0722))


@dataclass
class KeyInfo:
    """Bucket decision of one key evaluation (arrays shaped like the pass output)."""

    row: np.ndarray  # int32 LUT row
    angle_f: np.ndarray  # theta*24/pi before floor
    lam: np.ndarray  # sqrt(L1)
    mu: np.ndarray


@dataclass
class OracleResult:
    out: np.ndarray
    keys: List[KeyInfo] = field(default_factory=list)
    offset: Tuple[float, float] = (0.0, 0.0)


def lut_array(tex, lut_precision: str = "fp16") -> np.ndarray:
    """LUT payload as the GPU texture holds it: rgba16f storage = RNE to binary16 (App. D.1)."""
    data = np.asarray(tex.data, dtype=F32)
    if lut_precision == "fp16":
        data = data.astype(np.float16).astype(F32)
    elif lut_precision != "fp32":
        raise ValueError(lut_precision)
    return data


class _Plane:
    """Clamp-to-edge access to a [H, W] (or [H, W, C]) image by integer offset."""

    def __init__(self, img: np.ndarray, pad: int):
        self.h, self.w = img.shape[:2]
        self.pad = pad
        pw = ((pad, pad), (pad, pad)) + ((0, 0),) * (img.ndim - 2)
        self.p = np.pad(img, pw, mode="edge")

    def at(self, dx: int, dy: int) -> np.ndarray:
        p = self.pad
        return self.p[p + dy : p + dy + self.h, p + dx : p + dx + self.w]


def _window_geometry(family: str, r: int) -> Tuple[int, int, int]:
    if family in ("ravu-lite", "ravu-3x"):
        return 2 * r - 1, r - 1, {2: 3, 3: 3, 4: 5}[r]
    return 2 * r, r - 1, {2: 4, 3: 4, 4: 6}[r]


def _grad(family: str, n: int, s, i: int, j: int, axis: int) -> np.ndarray:
    """One finite difference in the exact operand order of the shader lines."""

    def S(d):
        return s[(i + d) * n + j] if axis == 0 else s[i * n + j + d]

    k = i if axis == 0 else j
    if family in ("ravu", "ravu-zoom") and k - 2 >= 0 and k + 2 <= n - 1:
        # (-s[+2] + 8.0*s[+1] - 8.0*s[-1] + s[-2]) / 12.0   (ravu-r3.hook:64)
        return (((-S(2)) + F32(8.0) * S(1)) - F32(8.0) * S(-1) + S(-2)) / F32(12.0)
    if k - 1 >= 0 and k + 1 <= n - 1:
        return (S(1) - S(-1)) / F32(2.0)
    if k - 1 < 0:
        return S(1) - S(0)
    return S(0) - S(-1)


def compute_key(v, s: List[np.ndarray]) -> KeyInfo:
    """s: list of n*n key samples (x-major, t = i*n + j) as float32 arrays."""
    family, r = v.family, v.radius
    n, _, g = _window_geometry(family, r)
    o = (n - g) // 2
    gauss = np.asarray(v.gauss, dtype=F32)
    a = np.zeros_like(s[0])
    b = np.zeros_like(s[0])
    d = np.zeros_like(s[0])
    q = 0
    for i in range(o, o + g):
        for j in range(o, o + g):
            gx = _grad(family, n, s, i, j, 0)
            gy = _grad(family, n, s, i, j, 1)
            gw = gauss[q]
            q += 1
            a = a + (gx * gx) * gw
            b = b + (gx * gy) * gw
            d = d + (gy * gy) * gw
    with np.errstate(all="ignore"):
        T = a + d
        D = a * d - b * b
        delta = np.sqrt(np.maximum(T * T / F32(4.0) - D, F32(0.0)))
        L1 = T / F32(2.0) + delta
        L2 = T / F32(2.0) - delta
        sqrtL1 = np.sqrt(L1)
        sqrtL2 = np.sqrt(L2)
        at = np.arctan2(L1 - a, b).astype(F32) + PI
        theta = at - PI * np.floor(at / PI)
        theta = np.where(np.abs(b) < EPS, F32(0.0), theta).astype(F32)
        lam = sqrtL1
        mu = ((sqrtL1 - sqrtL2) / (sqrtL1 + sqrtL2)).astype(F32)
        mu = np.where(sqrtL1 + sqrtL2 < EPS, F32(0.0), mu).astype(F32)
        angle_f = (theta * F32(24.0) / PI).astype(F32)
        angle = np.floor(angle_f)
        if v.strength_thr:
            strength = np.zeros_like(lam)
            for t in v.strength_thr:
                strength = strength + (lam >= F32(t)).astype(F32)
        else:
            strength = np.clip(np.floor(np.log2(lam * F32(v.strength_log2_scale) + EPS)), F32(0.0), F32(v.n_strength - 1))
        c0, c1 = v.coherence_thr
        coh = (mu >= F32(c0)).astype(F32) + (mu >= F32(c1)).astype(F32)
        rowf = (angle * F32(v.n_strength) + strength) * F32(v.n_coherence) + coh
    nrows = v.n_angle * v.n_strength * v.n_coherence
    rowf = np.where(np.isnan(rowf), F32(0.0), rowf)
    row = np.clip(rowf, 0, nrows - 1).astype(np.int32)
    return KeyInfo(row, angle_f, lam, mu)


def _as_planes(img: np.ndarray) -> np.ndarray:
    img = np.asarray(img, dtype=F32)
    return img[..., None] if img.ndim == 2 else img


def _key_plane(v, img3: np.ndarray) -> np.ndarray:
    if v.plane == "luma" or v.plane == "yuv":
        return img3[..., 0]
    cp = COLOR_PRIMARY
    return (img3[..., 0] * cp[0] + img3[..., 1] * cp[1]) + img3[..., 2] * cp[2]


# ----------------------------------------------------------------------------------------------
# ravu-lite
# ----------------------------------------------------------------------------------------------


def _pow32(c: np.ndarray) -> np.ndarray:
    for _ in range(5):
        c = c * c
    return c


def ravu_lite(img: np.ndarray, v, lut_precision: str = "fp16") -> OracleResult:
    """2x luma upscale, [H, W] -> [2H, 2W]  (``ravu-lite-ar-r3.hook:15-197``)."""
    img = np.asarray(img, dtype=F32)
    r = v.radius
    n, o, _ = _window_geometry("ravu-lite", r)
    N = n * n
    pl = _Plane(img, o)
    s = [pl.at(t // n - o, t % n - o) for t in range(N)]
    key = compute_key(v, s)
    lut = lut_array(v.lut, lut_precision)  # [288, (N+1)/2, 4]
    w_all = lut[key.row]  # [H, W, (N+1)/2, 4]
    res = np.zeros(img.shape + (4,), F32)
    ar = bool(v.ar)
    if ar:
        hi = np.zeros_like(res)
        lo = np.zeros_like(res)
        hi2 = np.zeros_like(res)
        lo2 = np.zeros_like(res)
        ar_taps = set(v.ar_taps)
    half = (N - 1) // 2
    for t in range(half):
        w = w_all[:, :, t, :]
        wr = w[..., ::-1]
        la, lb = s[t][..., None], s[N - 1 - t][..., None]
        res = res + (la * w + lb * wr)
        if ar and t in ar_taps:
            wg = np.maximum(F32(0.0), w)
            wgr = wg[..., ::-1]
            ca, cb = F32(0.1) + la, F32(0.1) + lb
            da, db = F32(1.1) - la, F32(1.1) - lb
            pa, pb, qa, qb = _pow32(ca), _pow32(cb), _pow32(da), _pow32(db)
            hi = hi + (pa * wg + pb * wgr)
            lo = lo + (qa * wg + qb * wgr)
            pa, pb, qa, qb = pa * ca, pb * cb, qa * da, qb * db
            hi2 = hi2 + (pa * wg + pb * wgr)
            lo2 = lo2 + (qa * wg + qb * wgr)
    w = w_all[:, :, half, :]
    lc = s[half][..., None]
    res = res + lc * w
    if ar:
        wg = np.maximum(F32(0.0), w)
        c, dd = F32(0.1) + lc, F32(1.1) - lc
        p, q = _pow32(c), _pow32(dd)
        hi = hi + p * wg
        lo = lo + q * wg
        p, q = p * c, q * dd
        hi2 = hi2 + p * wg
        lo2 = lo2 + q * wg
        with np.errstate(all="ignore"):
            lo = F32(1.1) - lo2 / lo
            hi = hi2 / hi - F32(0.1)
        st = F32(v.ar_strength)
        res = res * (F32(1.0) - st) + np.minimum(np.maximum(res, lo), hi) * st
    else:
        res = np.minimum(np.maximum(res, F32(0.0)), F32(1.0))
    H, W = img.shape
    out = np.empty((2 * H, 2 * W), F32)
    for c in range(4):  # phase c -> (2x + c/2, 2y + c%2)   (ravu-lite-ar-r3.hook:194-196)
        out[c % 2 :: 2, c // 2 :: 2] = res[..., c]
    return OracleResult(out, [key], (0.0, 0.0))


# ----------------------------------------------------------------------------------------------
# ravu (3 convolutions + merge)
# ----------------------------------------------------------------------------------------------


def _ravu_conv(v, lut: np.ndarray, key_s: List[np.ndarray], col_s: List[np.ndarray]) -> Tuple[np.ndarray, KeyInfo]:
    """key_s: n*n key samples [H,W]; col_s: n*n colour samples [H,W,C]."""
    key = compute_key(v, key_s)
    N = len(key_s)
    w_all = lut[key.row].reshape(key.row.shape + (-1,))  # [H, W, 4*lutw]
    res = np.zeros_like(col_s[0])
    for k in range(N // 2):
        res = res + (col_s[k] + col_s[N - 1 - k]) * w_all[..., k][..., None]
    res = np.minimum(np.maximum(res, F32(0.0)), F32(1.0))
    return res, key


def ravu(img: np.ndarray, v, lut_precision: str = "fp16", return_intermediates: bool = False, int11_override=None):
    """2x upscale, [H, W(, 3)] -> [2H, 2W(, 3)], offset (-0.5, -0.5)  (``ravu-r2.hook:15-338``).

    ``int11_override`` ([H, W(, 3)]): run steps 2-4 on THIS saved ``ravu_int11`` texture instead of the one step 1
    produced.  The parity tests pass the device's own int11 here: steps 2/3 are then evaluated on bit-identical
    inputs, so their buckets can be compared without the cascade that last-bit differences of the int11 VALUES
    (different summation order in the convolution) send through the ill-conditioned parts of the key."""
    img3 = _as_planes(img)
    r = v.radius
    n, o, _ = _window_geometry("ravu", r)
    N = n * n
    lut = lut_array(v.lut, lut_precision)
    H, W, C = img3.shape
    hooked = _Plane(img3, 2 * r)
    hooked_key = _Plane(_key_plane(v, img3), 2 * r)
    # step 1: int11 = value at (x + 1/2, y + 1/2)
    col = [hooked.at(t // n - o, t % n - o) for t in range(N)]
    ks = [hooked_key.at(t // n - o, t % n - o) for t in range(N)]
    int11, key1 = _ravu_conv(v, lut, ks, col)
    if int11_override is not None:
        int11 = _as_planes(np.asarray(int11_override, dtype=F32))
        assert int11.shape == img3.shape
    i11 = _Plane(int11, 2 * r)
    i11_key = _Plane(_key_plane(v, int11), 2 * r)
    keys = [key1]
    outs = []
    for tx2, ty2 in ((1, 0), (0, 1)):  # step 2: int10 at (x+1/2, y); step 3: int01 at (x, y+1/2)
        col, ks = [], []
        for t in range(N):
            i, j = t // n, t % n
            # twice the real position: P = (t.x - (r - 1/2), t.y) + i*(1/2, -1/2) + j*(1/2, 1/2)
            px2 = tx2 - (2 * r - 1) + i + j
            py2 = ty2 - i + j
            if px2 % 2 == 0:  # integer position -> HOOKED
                assert py2 % 2 == 0
                col.append(hooked.at(px2 // 2, py2 // 2))
                ks.append(hooked_key.at(px2 // 2, py2 // 2))
            else:  # half-integer position -> int11(x, y) ~ (x + 1/2, y + 1/2)
                assert py2 % 2 != 0
                col.append(i11.at((px2 - 1) // 2, (py2 - 1) // 2))
                ks.append(i11_key.at((px2 - 1) // 2, (py2 - 1) // 2))
        res, key = _ravu_conv(v, lut, ks, col)
        outs.append(res)
        keys.append(key)
    int10, int01 = outs
    out = np.empty((2 * H, 2 * W, C), F32)
    out[0::2, 0::2] = img3  # (2x, 2y)     = HOOKED   (ravu-r2.hook:327-338)
    out[1::2, 0::2] = int01  # (2x, 2y+1)   = int01
    out[0::2, 1::2] = int10  # (2x+1, 2y)   = int10
    out[1::2, 1::2] = int11  # (2x+1, 2y+1) = int11
    if np.asarray(img).ndim == 2:
        out = out[..., 0]
    res = OracleResult(out, keys, (-0.5, -0.5))
    if return_intermediates:
        return res, {"ravu_int11": int11, "ravu_int10": int10, "ravu_int01": int01}
    return res


# ----------------------------------------------------------------------------------------------
# ravu-3x
# ----------------------------------------------------------------------------------------------


def ravu_3x(img: np.ndarray, v, lut_precision: str = "fp16") -> OracleResult:
    """3x upscale (``compute/ravu-3x-r2.hook:15-115``; rgb ``compute/ravu-3x-r2-rgb.hook``)."""
    img3 = _as_planes(img)
    r = v.radius
    n, o, _ = _window_geometry("ravu-3x", r)
    N = n * n
    pl = _Plane(img3, o)
    plk = _Plane(_key_plane(v, img3), o)
    s = [plk.at(t // n - o, t % n - o) for t in range(N)]
    col = [pl.at(t // n - o, t % n - o) for t in range(N)]
    key = compute_key(v, s)
    lut = lut_array(v.lut, lut_precision)  # [216, N+1, 4]
    w_all = lut[key.row]
    H, W, C = img3.shape
    res0 = np.zeros((H, W, C, 4), F32)
    res1 = np.zeros((H, W, C, 4), F32)
    half = (N - 1) // 2
    for t in range(half):
        w0 = w_all[:, :, None, 2 * t, :]
        w1 = w_all[:, :, None, 2 * t + 1, :]
        la, lb = col[t][..., None], col[N - 1 - t][..., None]
        res0 = res0 + (la * w0 + lb * w1[..., ::-1])
        res1 = res1 + (la * w1 + lb * w0[..., ::-1])
    lc = col[half][..., None]
    res0 = res0 + lc * w_all[:, :, None, 2 * half, :]
    res1 = res1 + lc * w_all[:, :, None, 2 * half + 1, :]
    res0 = np.minimum(np.maximum(res0, F32(0.0)), F32(1.0))
    res1 = np.minimum(np.maximum(res1, F32(0.0)), F32(1.0))
    out = np.empty((3 * H, 3 * W, C), F32)
    # imageStore(gid*3 + ivec2(i, j)): x = 3x + i, y = 3y + j   (compute/ravu-3x-r2.hook:106-114)
    vals = [res0[..., 0], res0[..., 1], res0[..., 2], res0[..., 3], col[half], res1[..., 0], res1[..., 1], res1[..., 2], res1[..., 3]]
    for p, val in enumerate(vals):
        i, j = p // 3, p % 3
        out[j::3, i::3] = val
    if np.asarray(img).ndim == 2:
        out = out[..., 0]
    return OracleResult(out, [key], (0.0, 0.0))


# ----------------------------------------------------------------------------------------------
# ravu-zoom
# ----------------------------------------------------------------------------------------------


def _bilinear(lut: np.ndarray, cx: np.ndarray, cy: np.ndarray) -> np.ndarray:
    """GL LINEAR fetch with clamp-to-edge of lut [h, w, 4] at normalised coordinates (float32)."""
    h, w = lut.shape[:2]
    u = cx * F32(w) - F32(0.5)
    vv = cy * F32(h) - F32(0.5)
    u0, v0 = np.floor(u), np.floor(vv)
    fu, fv = (u - u0)[..., None], (vv - v0)[..., None]
    x0, y0 = u0.astype(np.int64), v0.astype(np.int64)

    def f(ix, iy):
        return lut[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]

    one = F32(1.0)
    top = f(x0, y0) * (one - fu) + f(x0 + 1, y0) * fu
    bot = f(x0, y0 + 1) * (one - fu) + f(x0 + 1, y0 + 1) * fu
    return (top * (one - fv) + bot * fv).astype(F32)


def zoom_positions(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical ``pos`` arithmetic of App. D.6: base texel index and subpixel phase (float32)."""
    o = np.arange(out_size, dtype=F32)
    pos = ((o + F32(0.5)) / F32(out_size)) * F32(in_size)
    t = pos - F32(0.5)
    sub = t - np.floor(t)
    pos = pos - sub
    return np.floor(pos).astype(np.int64), sub.astype(F32)


def ravu_zoom(img: np.ndarray, v, out_size: Tuple[int, int], lut_precision: str = "fp16") -> OracleResult:
    """Arbitrary-ratio upscale to out_size=(OW, OH)  (``ravu-zoom-r2.hook:23-134``)."""
    img3 = _as_planes(img)
    H, W, C = img3.shape
    OW, OH = out_size
    r = v.radius
    n, o, _ = _window_geometry("ravu-zoom", r)
    N = n * n
    bx, sx = zoom_positions(W, OW)
    by, sy = zoom_positions(H, OH)
    keyp = _key_plane(v, img3)

    def gather(plane, dx, dy):
        yy = np.clip(by + dy, 0, H - 1)[:, None]
        xx = np.clip(bx + dx, 0, W - 1)[None, :]
        return plane[yy, xx]

    ks = [gather(keyp, t // n - o, t % n - o) for t in range(N)]
    col = [gather(img3, t // n - o, t % n - o) for t in range(N)]
    key = compute_key(v, ks)
    nrows = v.n_angle * v.n_strength * v.n_coherence
    B = (N // 2 + 3) // 4
    lut = lut_array(v.lut, lut_precision)
    lut_ar = lut_array(v.lut_ar, lut_precision) if v.ar else None
    lutpos_a = F32(0.5) / F32(9.0)
    lutpos_b = F32(1.0) - F32(0.5) / F32(9.0)

    def lutpos(t):
        return lutpos_a * (F32(1.0) - t) + lutpos_b * t

    spx, spy = lutpos(sx)[None, :], lutpos(sy)[:, None]
    ipx, ipy = F32(1.0) - spx, F32(1.0) - spy
    spx, ipx = spx / F32(B), ipx / F32(B)
    spy, ipy = spy / F32(nrows), ipy / F32(nrows)
    coord_y = key.row.astype(F32) / F32(nrows)
    res = np.zeros((OH, OW, C), F32)
    if v.ar:
        hi = np.zeros_like(res)
        lo = np.zeros_like(res)
        hi2 = np.zeros_like(res)
        lo2 = np.zeros_like(res)
    fetched = []
    for mirrored in (False, True):
        for blk in range(B):
            cx = F32(repr(blk / B)) + (ipx if mirrored else spx)
            cy = coord_y + (ipy if mirrored else spy)
            cxb = np.broadcast_to(cx, (OH, OW))
            w = _bilinear(lut, cxb, cy)
            for c in range(4):
                k = blk * 4 + c
                if k >= N // 2:
                    break
                t = (N - 1 - k) if mirrored else k
                res = res + col[t] * w[..., c][..., None]
            if v.ar:
                fetched.append((mirrored, blk, _bilinear(lut_ar, cxb, cy)))
    if v.ar:
        for mirrored, blk, w in fetched:
            for pair in range(2):
                ka, kb = blk * 4 + 2 * pair, blk * 4 + 2 * pair + 1
                ta, tb = ((N - 1 - ka), (N - 1 - kb)) if mirrored else (ka, kb)
                wa, wb = w[..., 2 * pair][..., None], w[..., 2 * pair + 1][..., None]
                sa, sb = col[ta], col[tb]
                ca, da, cb, db = F32(0.1) + sa, F32(1.1) - sa, F32(0.1) + sb, F32(1.1) - sb
                pa, qa, pb, qb = _pow32(ca), _pow32(da), _pow32(cb), _pow32(db)
                hi = hi + (pa * wa + pb * wb)
                lo = lo + (qa * wa + qb * wb)
                pa, qa, pb, qb = pa * ca, qa * da, pb * cb, qb * db
                hi2 = hi2 + (pa * wa + pb * wb)
                lo2 = lo2 + (qa * wa + qb * wb)
        with np.errstate(all="ignore"):
            hi = hi2 / hi - F32(0.1)
            lo = F32(1.1) - lo2 / lo
        st = F32(v.ar_strength)
        res = res * (F32(1.0) - st) + np.minimum(np.maximum(res, lo), hi) * st
    else:
        res = np.minimum(np.maximum(res, F32(0.0)), F32(1.0))
    if np.asarray(img).ndim == 2:
        res = res[..., 0]
    return OracleResult(res, [key], (0.0, 0.0))


def run(img: np.ndarray, v, out_size: Optional[Tuple[int, int]] = None, lut_precision: str = "fp16") -> OracleResult:
    """Dispatch on the variant's family."""
    if v.family == "ravu-lite":
        return ravu_lite(img, v, lut_precision)
    if v.family == "ravu":
        return ravu(img, v, lut_precision)
    if v.family == "ravu-3x":
        return ravu_3x(img, v, lut_precision)
    if v.family == "ravu-zoom":
        if out_size is None:
            raise ValueError("ravu-zoom needs out_size=(OW, OH)")
        return ravu_zoom(img, v, out_size, lut_precision)
    raise ValueError(v.family)
