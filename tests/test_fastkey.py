"""CPU check of the sqrt/division/atan-free bucket decision used by the CUDA kernels
(mpv_prescalers_b200/csrc/common.cuh::key_from_abd_fast): a NumPy emulation of exactly that code is
compared with the oracle's op-for-op key on the same structure tensors."""
import numpy as np
import pytest

import oracle.ravu_np as R
from mpv_prescalers_b200.api import _strength_of_l1, l1_thresholds
from mpv_prescalers_b200.hookfile import HookFile
from mpv_prescalers_b200.synth import batch
from tests.conftest import hook_path

F32 = np.float32
EPS = F32(1.192092896e-7)


def fast_rows(v, a, b, d):
    thr = l1_thresholds(v)
    with np.errstate(all="ignore"):
        T = a + d
        D = a * d - b * b
        delta = np.sqrt(np.maximum(T * T * F32(0.25) - D, F32(0)))
        hT = T * F32(0.5)
        L1, L2 = hT + delta, hT - delta
        strength = sum((L1 >= t).astype(np.int32) for t in thr)
        r = [F32(((1 + c) / (1 - c)) ** 2) for c in v.coherence_thr]
        coh_fast = (L1 >= L2 * r[0]).astype(np.int32) + (L1 >= L2 * r[1]).astype(np.int32)
        s1, s2 = np.sqrt(L1), np.sqrt(L2)
        ss = s1 + s2
        mu = np.where(ss < EPS, F32(0), (s1 - s2) / ss)
        coh_slow = (mu >= F32(v.coherence_thr[0])).astype(np.int32) + (mu >= F32(v.coherence_thr[1])).astype(np.int32)
        coh = np.where((L1 < F32(3.5e-15)) | (L2 < 0), 0, np.where(L1 < F32(1.5e-14), coh_slow, coh_fast))
        X, Y = b.copy(), L1 - a
        flip = Y < 0
        X, Y = np.where(flip, -X, X), np.where(flip, -Y, Y)
        neg = X < 0
        X = np.abs(X)
        sw = Y > X
        lo, hi = np.where(sw, X, Y), np.where(sw, Y, X)
        s = np.zeros(a.shape, np.int32)
        for k in range(1, 6):
            s += (lo >= hi * F32(np.tan(k * np.pi / 24))).astype(np.int32)
        s = np.where(sw, 11 - s, s)
        ang = np.where(neg, 23 - s, s)
        ang = np.where((np.abs(b) < EPS) | (L1 - a == 0), 0, ang)
    return (ang * v.n_strength + strength) * 3 + coh


def _abd(v, s):
    n, _, g = R._window_geometry(v.family, v.radius)
    o = (n - g) // 2
    gauss = np.asarray(v.gauss, dtype=F32)
    a = np.zeros_like(s[0]); b = np.zeros_like(s[0]); d = np.zeros_like(s[0])
    q = 0
    for i in range(o, o + g):
        for j in range(o, o + g):
            gx, gy, gw = R._grad(v.family, n, s, i, j, 0), R._grad(v.family, n, s, i, j, 1), gauss[q]
            q += 1
            a = a + (gx * gx) * gw
            b = b + (gx * gy) * gw
            d = d + (gy * gy) * gw
    return a, b, d


@pytest.mark.parametrize("name", ["ravu-lite-ar-r3.hook", "ravu-r3.hook", "compute/ravu-3x-r2.hook", "ravu-zoom-r2.hook"])
def test_l1_thresholds_are_exactly_equivalent(name):
    v = HookFile.parse(hook_path(name)).variant
    thr = l1_thresholds(v)
    assert len(thr) == v.n_strength - 1 and list(thr) == sorted(thr)
    rng = np.random.default_rng(1)
    x = np.exp(rng.uniform(np.log(1e-10), np.log(4.0), 400000)).astype(F32)
    edge = np.concatenate([np.nextafter(np.asarray(thr, F32), F32(0)), np.asarray(thr, F32), np.nextafter(np.asarray(thr, F32), F32(9))])
    x = np.concatenate([x, edge, np.zeros(1, F32)])
    assert np.array_equal(_strength_of_l1(v, x), sum((x >= t).astype(np.int32) for t in thr))


@pytest.mark.parametrize("name", ["ravu-lite-ar-r3.hook", "ravu-lite-r2.hook", "ravu-r4.hook", "compute/ravu-3x-r3.hook"])
def test_fast_key_matches_oracle_key(name, monkeypatch):
    v = HookFile.parse(hook_path(name)).variant
    captured = []
    orig = R.compute_key

    def spy(vv, s):
        captured.append(_abd(vv, s))
        return orig(vv, s)

    monkeypatch.setattr(R, "compute_key", spy)
    img = batch(1, 1, 360, 640, config=41)[0, 0]
    res = R.run(img, v)
    total = bad = 0
    for (a, b, d), key in zip(captured, res.keys):
        rows = fast_rows(v, a, b, d)
        bad += int((rows != key.row).sum())
        total += rows.size
    assert bad / total <= 2e-5, f"{name}: fast key differs on {bad}/{total}"


def test_fast_key_degenerate_planes():
    """flat, axis-aligned step and exact-diagonal planes: the shader's special cases must be kept."""
    v = HookFile.parse(hook_path("ravu-lite-r3.hook")).variant
    y, x = np.mgrid[0:24, 0:32]
    for img in (np.full((24, 32), 0.5, F32), (x >= 16).astype(F32) * 0.5 + 0.25, (y >= 12).astype(F32) * 0.5 + 0.25):
        n, o, _ = R._window_geometry("ravu-lite", 3)
        pl = R._Plane(img, o)
        s = [pl.at(t // n - o, t % n - o) for t in range(n * n)]
        key = R.compute_key(v, s)
        assert np.array_equal(fast_rows(v, *_abd(v, s)), key.row)
