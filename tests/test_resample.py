"""The offset-correcting main scaler: oracle properties on the CPU, device parity on the GPU."""
import numpy as np
import pytest

from oracle import resample_np
from tests.conftest import hook_path

KERNELS = ["bilinear", "catmull_rom", "mitchell", "spline36", "lanczos"]


def _img(h, w, seed=0):
    from mpv_prescalers_b200.synth import plane

    return plane(h, w, seed)


@pytest.mark.parametrize("name", KERNELS)
def test_oracle_preserves_constants_and_integer_shifts(name):
    img = _img(37, 53, 1)
    flat = np.full((20, 31), 0.3125, np.float32)
    assert np.abs(resample_np.resample(flat, (33, 47), (-0.5, -0.5), name) - 0.3125).max() <= 1e-6
    if name == "mitchell":
        return  # not an interpolating kernel (K(0) = 8/9): it smooths even at zero offset, as in mpv
    ident = resample_np.resample(img, None, (0.0, 0.0), name)
    assert np.abs(ident - img).max() <= 1e-6
    sh = resample_np.resample(img, None, (2.0, -1.0), name)      # out(x, y) = in(x + 2, y - 1), clamp-to-edge
    want = img[np.clip(np.arange(37) - 1, 0, 36)][:, np.clip(np.arange(53) + 2, 0, 52)]
    assert np.abs(sh - want).max() <= 1e-6


def test_oracle_half_texel_weights_are_the_textbook_ones():
    _, w = resample_np.axis_table(16, 16, -0.5, "catmull_rom")
    assert np.allclose(w[5], [-1 / 16, 9 / 16, 9 / 16, -1 / 16], atol=1e-7)
    _, w = resample_np.axis_table(16, 16, -0.5, "bilinear")
    assert np.allclose(w[5], [0.5, 0.5], atol=1e-7)


def test_correction_puts_ravu_on_the_ravu_lite_grid():
    """ravu's result is half a texel off (OFFSET -0.5 -0.5, ravu-r2.hook:325), ravu-lite's is not: after the correction
    the two 2x upscales of the same plane must agree far better than before (both on the CPU oracle)."""
    from mpv_prescalers_b200 import HookFile
    from oracle import ravu_np
    from tests.parity import psnr

    y, x = np.mgrid[0:72, 0:96].astype(np.float64)
    img = (0.5 + 0.3 * np.sin(0.21 * x + 0.13 * y) + 0.15 * np.sin(0.05 * x * y / 40)).astype(np.float32).clip(0, 1)
    lite = ravu_np.run(img, HookFile.parse(hook_path("ravu-lite-r3.hook")).variant).out
    rv = ravu_np.run(img, HookFile.parse(hook_path("ravu-r3.hook")).variant)
    assert rv.offset == (-0.5, -0.5)
    fixed = resample_np.resample(rv.out, None, rv.offset, "lanczos")
    before, after = psnr(rv.out[8:-8, 8:-8], lite[8:-8, 8:-8]), psnr(fixed[8:-8, 8:-8], lite[8:-8, 8:-8])
    assert after >= before + 10.0 and after >= 40.0, (before, after)


# ---- device ------------------------------------------------------------------------------------------------------------

torch = pytest.importorskip("torch")
CASES = [((37, 53), None, (-0.5, -0.5)), ((64, 96), (96, 144), (-0.5, 0.0)), ((50, 70), (40, 61), (0.0, -0.5)),
         ((1, 9), None, (-0.5, -0.5)), ((130, 200), (173, 267), (0.25, -0.75)), ((33, 31), (66, 62), (0.0, 0.0))]


@pytest.mark.gpu
@pytest.mark.parametrize("name", KERNELS)
@pytest.mark.parametrize("hw,out_hw,off", CASES)
def test_device_resampler_matches_oracle(name, hw, out_hw, off):
    """Tolerance: 2e-6 max abs (float32 tables computed independently in double on both sides, fp32 accumulation)."""
    from mpv_prescalers_b200 import resample

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    x = np.stack([_img(hw[0], hw[1], 7), _img(hw[0], hw[1], 8)])
    got = resample(torch.from_numpy(x).cuda(), out_hw, off, name).cpu().numpy()
    for f in range(2):
        ref = resample_np.resample(x[f], out_hw, off, name)
        assert got[f].shape == ref.shape
        assert np.abs(got[f] - ref).max() <= 2e-6, f"{name} {hw}->{out_hw} offset {off}: {np.abs(got[f] - ref).max():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("hook", ["ravu-r3.hook", "nnedi3-nns32-win8x4.hook", "ravu-r2-rgb.hook"])
def test_prescale_correct_offset(hook):
    """prescale(correct_offset=...) = hook, then the scaler at the accumulated offset: equal to doing the two steps by hand,
    offset reported as (0, 0); integer planes are quantised once, at the end."""
    from mpv_prescalers_b200 import HookFile, prescale, resample
    from mpv_prescalers_b200.synth import batch

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    hk = HookFile.parse(hook_path(hook))
    x = torch.from_numpy(batch(2, hk.variant.channels, 60, 84, config=61)).cuda()
    raw = prescale(x, hk)
    assert tuple(raw.offset) == (-0.5, -0.5)
    fixed = prescale(x, hk, correct_offset="catmull_rom")
    assert tuple(fixed.offset) == (0.0, 0.0) and fixed.shape == raw.shape and fixed.applied
    assert torch.equal(fixed, resample(raw, None, raw.offset, "catmull_rom"))
    ref = resample_np.resample(raw[0, 0].cpu().numpy(), None, (-0.5, -0.5), "catmull_rom")
    assert np.abs(fixed[0, 0].cpu().numpy() - ref).max() <= 2e-6
    x8 = torch.round(x.clamp(0, 1) * 255).to(torch.uint8)
    f8 = prescale(x8, hk, correct_offset=True)
    f32 = prescale(x8, hk, correct_offset=True, out_dtype=torch.float32)
    assert f8.dtype == torch.uint8 and torch.equal(f8.float(), torch.round(f32.clamp(0, 1) * 255))
    lite = prescale(x[:, :1], hook_path("ravu-lite-r3.hook"), correct_offset=True)     # nothing to correct
    assert tuple(lite.offset) == (0.0, 0.0) and torch.equal(lite, prescale(x[:, :1], hook_path("ravu-lite-r3.hook")))
