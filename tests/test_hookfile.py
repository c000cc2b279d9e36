"""Parser / planner tests (CPU).  Known answers: SURVEY.md App. C."""
import glob
import os

import numpy as np
import pytest

from mpv_prescalers_b200.hookfile import HookError, HookFile, eval_rpn, gaussian_weights
from tests.conftest import hook_path

REF = "/root/reference"
ALL = sorted(glob.glob(REF + "/*.hook") + glob.glob(REF + "/*/*.hook"))


@pytest.mark.skipif(not ALL, reason="reference snapshot not mounted")
def test_every_shipped_file_parses_and_classifies():
    assert len(ALL) == 99
    fams = {}
    for f in ALL:
        v = HookFile.parse(f).variant
        fams[v.family] = fams.get(v.family, 0) + 1
        rel = os.path.relpath(f, REF)
        assert v.flavour == (rel.split("/")[0] if "/" in rel else "root")
    assert fams == {"ravu-lite": 18, "ravu": 21, "ravu-zoom": 21, "ravu-3x": 9, "nnedi3": 30}


LUT_KAT = {  # file -> (texture, w, h, sha256[:16], first texel)
    "ravu-lite-r2.hook": ("ravu_lite_lut2", 5, 288, "a44a58ef7c8c9167", (0.0355936550, -0.0111581720, -0.0123693654, -0.0027147743)),
    "ravu-lite-ar-r3.hook": ("ravu_lite_lut3", 13, 288, "12ca04cfc91ab259", (0.000373310177, -6.13176599e-05, -8.34040184e-05, -0.000314347912)),
    "ravu-lite-r4.hook": ("ravu_lite_lut4", 25, 288, "619b849ad8ce5064", None),
    "ravu-r2.hook": ("ravu_lut2", 2, 648, "a5ab218aed83c093", (0.00698380871, 0.0125906318, 0.0115921479, 0.00971504394)),
    "ravu-r3.hook": ("ravu_lut3", 5, 648, "107b97c0e21d992e", None),
    "ravu-r4.hook": ("ravu_lut4", 8, 648, "170aa0f3e782e390", None),
    "compute/ravu-3x-r2.hook": ("ravu_3x_lut2", 10, 216, "2fd4ff5f0b689e66", None),
    "compute/ravu-3x-r3.hook": ("ravu_3x_lut3", 26, 216, "bbecba83cfe98ea4", None),
    "compute/ravu-3x-r4.hook": ("ravu_3x_lut4", 50, 216, "cca3995890861371", None),
    "ravu-zoom-r2.hook": ("ravu_zoom_lut2", 18, 2592, "565c6af51bf23827", (0, 0, 0, 0)),
    "ravu-zoom-r3.hook": ("ravu_zoom_lut3", 45, 2592, "e0091b14bc99c447", (0, 0, 0, 0)),
}


@pytest.mark.parametrize("name", sorted(LUT_KAT))
def test_lut_known_answers(name):
    tex_name, w, h, sha, first = LUT_KAT[name]
    hk = HookFile.parse(hook_path(name))
    t = hk.textures[tex_name]
    assert (t.width, t.height, t.sha16) == (w, h, sha)
    if first is not None:
        assert np.allclose(t.data[0, 0], first, rtol=1e-6, atol=0)


def test_zoom_ar_has_two_luts():
    v = HookFile.parse(hook_path("ravu-zoom-ar-r2.hook")).variant
    assert v.ar and v.lut.sha16 == "565c6af51bf23827" and v.lut_ar.sha16 == "75dec6c5bfc4cea3"
    assert v.ar_strength == pytest.approx(0.8) and len(v.ar_taps) == 16


def test_nnedi3_known_answers_and_transpose_identity():
    v = HookFile.parse(hook_path("nnedi3-nns16-win8x4.hook")).variant
    # first W(0,...) of neuron 0 = samples (-3,-1..2): a=0, b=0..3   (nnedi3-nns16-win8x4.hook:34)
    assert np.allclose(v.nn_y.w1[0, 0, :], (-0.0339266136, -0.0881003812, 0.216080278, -0.00652594119), rtol=1e-6)
    assert np.allclose((v.nn_y.b1[0], v.nn_y.b2[0]), (0.114879131, 0.147320032), rtol=1e-6)
    # double_x weights are the transpose of double_y (identical in canonical long/short order); biases equal
    assert np.array_equal(v.nn_x.w1, v.nn_y.w1) and np.array_equal(v.nn_x.w2, v.nn_y.w2)
    assert np.array_equal(v.nn_x.b1, v.nn_y.b1) and np.array_equal(v.nn_x.b2, v.nn_y.b2)
    # mean-removed weights: sum ~ 0 per neuron
    assert np.abs(v.nn_y.w1.reshape(16, -1).sum(1)).max() < 1e-5


def test_nnedi3_flavours_decode_to_the_same_weights():
    root = HookFile.parse(hook_path("nnedi3-nns16-win8x4.hook")).variant
    for other in ("gather/nnedi3-nns16-win8x4.hook", "compute/nnedi3-nns16-win8x4.hook"):
        o = HookFile.parse(hook_path(other)).variant
        assert np.array_equal(o.nn_y.w1, root.nn_y.w1) and np.array_equal(o.nn_x.w2, root.nn_x.w2)
        assert np.array_equal(o.nn_y.b2, root.nn_y.b2)


def test_variant_constants():
    v = HookFile.parse(hook_path("ravu-lite-ar-r3.hook")).variant
    assert (v.family, v.radius, v.ar, v.scale, v.taps) == ("ravu-lite", 3, True, 2, 25)
    assert v.strength_thr == (0.004, 0.016, 0.05) and v.coherence_thr == (0.25, 0.5)
    assert v.ar_strength == pytest.approx(0.8) and len(v.ar_taps) == 13
    assert np.allclose(v.gauss, gaussian_weights(3), atol=1e-7)
    v = HookFile.parse(hook_path("ravu-r4.hook")).variant
    assert v.strength_log2_scale == 2000.0 and v.n_strength == 9 and v.offset == (-0.5, -0.5) and len(v.gauss) == 36
    v = HookFile.parse(hook_path("compute/ravu-3x-r2.hook")).variant
    assert v.scale == 3 and v.strength_thr == (0.005, 0.02) and v.n_strength == 3


def test_rpn_truth_table():
    env = {"HOOKED": (960, 540), "OUTPUT": (1920, 1080), "LUMA": (960, 540)}
    when = "HOOKED.w OUTPUT.w / 0.707106 < HOOKED.h OUTPUT.h / 0.707106 < *".split()
    assert eval_rpn(when, env) == 1.0
    assert eval_rpn(when, {**env, "OUTPUT": (1200, 1080)}) == 0.0
    assert eval_rpn(when, {**env, "OUTPUT": (1920, 700)}) == 0.0
    assert eval_rpn("2 HOOKED.w *".split(), env) == 1920
    assert eval_rpn("HOOKED.w OUTPUT.w < HOOKED.h OUTPUT.h < * LUMA.w 0 > *".split(), env) == 1.0
    assert eval_rpn("LUMA.w 0 >".split(), {**env, "LUMA": (0, 0)}) == 0.0
    assert eval_rpn("3 4 + 2 - 5 =".split(), env) == 1.0
    with pytest.raises(HookError):
        eval_rpn("1 +".split(), env)
    with pytest.raises(HookError):
        eval_rpn("FOO.w".split(), env)


def test_plan_semantics():
    from mpv_prescalers_b200.api import plan

    hk = HookFile.parse(hook_path("ravu-lite-ar-r3.hook"))
    p = plan(hk, (540, 960))
    assert p.applied and p.out_size == (1080, 1920) and p.offset == (0.0, 0.0) and len(p.passes) == 2
    assert not plan(hk, (540, 960), (700, 1200)).applied  # ratio > 0.707106 -> WHEN false
    hk = HookFile.parse(hook_path("ravu-r3.hook"))
    assert plan(hk, (540, 960)).offset == (-0.5, -0.5)
    hk = HookFile.parse(hook_path("nnedi3-nns16-win8x4.hook"))
    p = plan(hk, (540, 960), (1080, 1000))
    assert p.double_y and not p.double_x and p.out_size == (1080, 960) and p.offset == (0.0, -0.5)
    hk = HookFile.parse(hook_path("ravu-r2-yuv.hook"))
    assert plan(hk, (64, 64), is_yuv=True).applied and not plan(hk, (64, 64), is_yuv=False).applied
    hk = HookFile.parse(hook_path("ravu-zoom-r2.hook"))
    with pytest.raises(HookError):
        plan(hk, (64, 64))
    assert plan(hk, (64, 64), (100, 150)).out_size == (100, 150)
    assert not plan(hk, (64, 64), (64, 150)).applied


def test_rejects_modified_or_foreign_files(tmp_path):
    src = open(hook_path("ravu-lite-r2.hook")).read()
    # hand-edited Gaussian weight
    bad = tmp_path / "bad-gauss.hook"
    bad.write_text(src.replace("0.13080118386382833", "0.23080118386382833"))
    with pytest.raises(HookError, match="Gaussian"):
        HookFile.parse(bad).variant
    # re-associated stencil
    bad2 = tmp_path / "bad-stencil.hook"
    bad2.write_text(src.replace("gx = (luma7-luma1)/2.0;", "gx = (luma7-luma1)*0.5;"))
    with pytest.raises(HookError):
        HookFile.parse(bad2).variant
    # unknown texture format (README.md:19-21)
    bad3 = tmp_path / "bad-format.hook"
    bad3.write_text(src.replace("//!FORMAT rgba16f", "//!FORMAT rgba16hf"))
    with pytest.raises(HookError, match="FORMAT"):
        HookFile.parse(bad3)
    # truncated payload
    bad4 = tmp_path / "bad-payload.hook"
    bad4.write_text(src.rstrip()[:-8] + "\n")
    with pytest.raises(HookError, match="payload"):
        HookFile.parse(bad4)
    # not a hook at all
    bad5 = tmp_path / "x.hook"
    bad5.write_text("//!DESC something else\n//!HOOK LUMA\n//!BIND HOOKED\nvec4 hook() { return HOOKED_tex(HOOKED_pos); }\n")
    with pytest.raises(HookError, match="unsupported shader"):
        HookFile.parse(bad5).variant


def test_missing_zoom_ar_r3_is_named():
    from mpv_prescalers_b200.hookfile import find_hook

    with pytest.raises(HookError, match="MISSING_LARGE_BLOBS"):
        find_hook("ravu-zoom-ar-r3.hook")


@pytest.mark.skipif(not ALL, reason="reference snapshot not mounted")
def test_flavours_carry_identical_payloads_and_constants():
    """root / gather / compute files of one variant ship the same LUT bytes and the same key constants."""
    by_name = {}
    for f in ALL:
        by_name.setdefault(os.path.basename(f), []).append(f)
    checked = 0
    for name, files in by_name.items():
        if len(files) < 2:
            continue
        vs = [HookFile.parse(f).variant for f in files]
        base = vs[0]
        for v in vs[1:]:
            assert (v.family, v.radius, v.ar, v.plane, v.nns, v.win) == (base.family, base.radius, base.ar, base.plane, base.nns, base.win)
            if base.lut is not None:
                assert v.lut.sha16 == base.lut.sha16
                assert np.array_equal(v.gauss, base.gauss) and v.strength_thr == base.strength_thr
            if base.nn_y is not None:
                assert np.array_equal(v.nn_y.w1, base.nn_y.w1) and np.array_equal(v.nn_x.w2, base.nn_x.w2)
            checked += 1
    assert checked >= 40


def test_content_key_identifies_the_weights_not_the_path(hooks):
    """The device weight cache is keyed on content: in-memory hooks (all of path '<string>') must not share entries
    unless they really hold the same weights and constants."""
    from mpv_prescalers_b200.hookfile import HookFile

    ta = open(hooks("ravu-lite-r3.hook")).read()
    tb = open(hooks("ravu-lite-ar-r3.hook")).read()
    a, a2, b = HookFile.parse_text(ta), HookFile.parse_text(ta), HookFile.parse_text(tb)
    assert a.path == b.path == "<string>"
    assert a.content_key == a2.content_key
    assert a.content_key != b.content_key            # same LUT payload, different kernel parameters (anti-ringing)
    assert HookFile.parse(hooks("ravu-lite-r3.hook")).content_key == a.content_key
    c = HookFile.parse_text(open(hooks("ravu-lite-r2.hook")).read())
    assert c.content_key != a.content_key


def test_nnedi3_weights_must_be_mean_removed(hooks):
    """The tensor-core path feeds (x - mean) / sigma into the contraction, which equals the shader's dot(x, W) / sigma
    only for weight rows that sum to zero: a file that breaks this is refused, not silently mis-run."""
    import re
    import struct

    from mpv_prescalers_b200.hookfile import HookError, HookFile

    text = open(hooks("nnedi3-nns16-win8x4.hook")).read()
    HookFile.parse_text(text).variant  # the shipped file passes
    one = struct.unpack("<i", struct.pack("<f", 1.0))[0]
    bad, n = re.subn(r"sum1=W\(0,(-?\d+),", f"sum1=W(0,{one},", text, count=1)
    assert n == 1
    with pytest.raises(HookError, match="mean-removed"):
        HookFile.parse_text(bad).variant
