"""Oracle pinning (CPU): parametric NumPy oracle vs literal execution of the shader text
(oracle/glsl_exec.py) and vs the committed golden fixtures (tests/golden, made by tools/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from mpv_prescalers_b200.hookfile import HookFile
from mpv_prescalers_b200.synth import batch, natural, pathological
from oracle import nnedi3_np, ravu_np
from tests.conftest import hook_path

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _img(x):
    return x[0] if x.shape[0] == 1 else np.moveaxis(x, 0, -1)


@pytest.mark.parametrize("fn", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_oracle_matches_golden_fixture(fn):
    g = np.load(fn)
    name = str(g["hook"])
    v = HookFile.parse(hook_path(name)).variant
    img = _img(g["input"])
    if v.family == "nnedi3":
        out, off = nnedi3_np.nnedi3(img, v)
        assert np.abs(out - g["output"]).max() <= 2e-5  # GEMM-order vs neuron-serial fp32 summation
    else:
        osz = tuple(int(t) for t in g["out_size"]) if g["out_size"][0] else None
        r = ravu_np.run(img, v, osz)
        out, off = r.out, r.offset
        assert np.array_equal(out, g["output"]), "parametric oracle must be bit-exact with the literal shader run"
    assert tuple(off) == tuple(g["offset"])


LITERAL = [
    ("ravu-lite-ar-r3.hook", None), ("ravu-lite-r2.hook", None), ("ravu-r3.hook", None), ("ravu-r2-rgb.hook", None),
    ("ravu-r4-yuv.hook", None), ("compute/ravu-3x-r3.hook", None), ("compute/ravu-3x-r2-rgb.hook", None),
    ("ravu-zoom-r3.hook", (93, 71)), ("ravu-zoom-ar-r2.hook", (96, 72)), ("ravu-zoom-ar-r2-rgb.hook", (80, 59)),
]


@pytest.mark.parametrize("name,out_size", LITERAL)
def test_oracle_bit_exact_vs_literal_shader(name, out_size):
    """Run the reference's GLSL text literally (needs /root/reference or baseline/_ref/hooks)."""
    from oracle.glsl_exec import run_hook

    path = hook_path(name)
    v = HookFile.parse(path).variant
    x = batch(1, v.channels, 24, 32, config=31)[0]
    img = _img(x)
    kw = {"out_size": out_size} if out_size else {}
    ref, off, applied = run_hook(path, img, **kw)
    r = ravu_np.run(img, v, out_size)
    assert applied and np.array_equal(r.out, ref) and tuple(off) == tuple(r.offset)


@pytest.mark.parametrize("name,out_size", [("ravu-lite-ar-r3.hook", None), ("ravu-r2.hook", None), ("ravu-zoom-ar-r2.hook", (70, 55))])
def test_oracle_bit_exact_vs_literal_shader_on_a_natural_plane(name, out_size):
    """The same pinning on a 1/f-spectrum plane: other buckets, other clamps than the synthetic mixture."""
    from oracle.glsl_exec import run_hook

    path = hook_path(name)
    v = HookFile.parse(path).variant
    img = natural(22, 30, seed=5)
    assert img.dtype == np.float32 and 0.0 <= img.min() and img.max() <= 1.0 and np.array_equal(img, natural(22, 30, seed=5))
    ref, off, applied = run_hook(path, img, **({"out_size": out_size} if out_size else {}))
    r = ravu_np.run(img, v, out_size)
    assert applied and np.array_equal(r.out, ref) and tuple(off) == tuple(r.offset)


def test_compute_and_root_flavours_agree_literally():
    """Three statements of the same math (root / gather / compute) must agree."""
    from oracle.glsl_exec import run_hook

    img = batch(1, 1, 40, 70, config=32)[0, 0]
    a, _, _ = run_hook(hook_path("ravu-lite-ar-r3.hook"), img)
    b, _, _ = run_hook(hook_path("compute/ravu-lite-ar-r3.hook"), img)
    assert np.abs(a - b).max() <= 1e-6
    a, _, _ = run_hook(hook_path("nnedi3-nns16-win8x4.hook"), img)
    b, _, _ = run_hook(hook_path("compute/nnedi3-nns16-win8x4.hook"), img)
    assert np.abs(a - b).max() <= 1e-5


@pytest.mark.parametrize("name", ["ravu-lite-r3.hook", "ravu-lite-ar-r4.hook", "ravu-r3.hook", "ravu-r2-rgb.hook", "nnedi3-nns32-win8x6.hook"])
def test_flat_in_flat_out(name):
    """LUT rows are partitions of unity; NNEDI3 weights are mean-removed."""
    v = HookFile.parse(hook_path(name)).variant
    img = np.full((12, 17) if v.channels == 1 else (12, 17, 3), 0.3125, np.float32)
    out = nnedi3_np.nnedi3(img, v)[0] if v.family == "nnedi3" else ravu_np.run(img, v).out
    assert np.abs(out - 0.3125).max() <= 3e-4


def test_zoom_flat_and_positions():
    v = HookFile.parse(hook_path("ravu-zoom-r3.hook")).variant
    out = ravu_np.ravu_zoom(np.full((10, 14), 0.5, np.float32), v, (42, 30)).out
    assert np.abs(out - 0.5).max() <= 2e-3  # zoom LUT sums are 0.9985..1.00007
    base, sub = ravu_np.zoom_positions(1280, 3840)
    assert base[0] == -1 or base[0] == 0
    assert np.all((sub >= 0) & (sub < 1)) and np.all(np.diff(base) >= 0)


def test_anti_ringing_limits_overshoot():
    """a 0.2 -> 0.8 step overshoots without AR and is cut by ~80 % with AR 0.8."""
    step = pathological(24, 32)["step"]
    plain = ravu_np.run(step, HookFile.parse(hook_path("ravu-lite-r3.hook")).variant).out
    ar = ravu_np.run(step, HookFile.parse(hook_path("ravu-lite-ar-r3.hook")).variant).out
    over_plain = max(plain.max() - 0.8, 0.2 - plain.min())
    over_ar = max(ar.max() - 0.8, 0.2 - ar.min())
    assert over_plain > 0.02 and over_ar < 0.35 * over_plain


def test_key_histogram_covers_buckets():
    v = HookFile.parse(hook_path("ravu-lite-r3.hook")).variant
    r = ravu_np.run(batch(1, 1, 135, 240, config=2)[0, 0], v)
    rows = r.keys[0].row
    assert rows.min() >= 0 and rows.max() < 288
    assert len(np.unique(rows // 12)) >= 20  # angles
    assert len(np.unique(rows % 3)) == 3  # coherence classes
