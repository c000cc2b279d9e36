"""Generic hook runner: GLSL -> CUDA transpilation is checked on the CPU box (NVRTC needs no
GPU), execution against the literal CPU execution of the same shader text on the GPU box."""
import re

import numpy as np
import pytest

from tests.conftest import hook_path

COMPILE = ["ravu-lite-ar-r3.hook", "ravu-r2.hook", "ravu-r2-rgb.hook", "ravu-zoom-ar-r2-rgb.hook", "gather/ravu-lite-ar-r2.hook",
           "gather/ravu-r3.hook", "nnedi3-nns16-win8x4.hook", "gather/nnedi3-nns16-win8x4.hook", "ravu-zoom-r3.hook",
           "compute/ravu-lite-r2.hook", "compute/ravu-lite-ar-r3.hook", "compute/ravu-r3-rgb.hook", "compute/ravu-3x-r2.hook",
           "compute/ravu-3x-r3-rgb.hook", "compute/ravu-zoom-r2.hook", "compute/nnedi3-nns16-win8x4.hook"]


@pytest.mark.parametrize("name", COMPILE)
def test_hook_transpiles_and_compiles_for_sm100a(name):
    from mpv_prescalers_b200 import HookFile
    from mpv_prescalers_b200.hookrunner import GenericHook

    gh = GenericHook(HookFile.parse(hook_path(name)))
    assert len(gh.cubin) > 1000 and b"pass_0" in gh.cubin
    body = gh.source.split("// pass 0")[1]
    assert re.search(r"(?<![\w.])\d+\.\d+(?:[eE][-+]?\d+)?(?![\w.])", body) is None   # every literal became float32


def test_compute_pass_keeps_the_shaders_own_work_group_structure():
    """//!COMPUTE passes: shared arrays become raw __shared__ storage, barrier() a block barrier, imageStore a bounds-checked
    store; every thread of the work group reaches the barrier (no early return)."""
    from mpv_prescalers_b200 import HookFile
    from mpv_prescalers_b200.hookrunner import GenericHook

    gh = GenericHook(HookFile.parse(hook_path("compute/ravu-lite-r2.hook")))
    body = gh.source.split("// pass 0")[1]
    assert "__shared__ float _sh_inp[340];" in body and "barrier();" in body and "imageStore(out_image" in body
    assert "if (_ox >= _ow || _oy >= _oh) return;" not in body
    assert re.search(r"^\s*shared\b", body, flags=re.M) is None


def test_unsupported_shared_declaration_is_refused():
    from mpv_prescalers_b200 import HookError, HookFile
    from mpv_prescalers_b200.hookrunner import GenericHook

    text = open(hook_path("compute/ravu-lite-r2.hook")).read().replace("shared float inp[340];", "shared int inp[340];")
    with pytest.raises(HookError, match="shared"):
        GenericHook(HookFile.parse_text(text))


def _modified_hook_text():
    """ravu-lite-r3 with a hand edit the fused kernels cannot express: the result is attenuated before the clamp."""
    text = open(hook_path("ravu-lite-r3.hook")).read()
    assert text.count("res = clamp(res, 0.0, 1.0);") == 1
    return text.replace("res = clamp(res, 0.0, 1.0);", "res = clamp(res * 0.75 + 0.125, 0.0, 1.0);")


def test_hand_edited_hook_is_refused_by_the_fused_path():
    from mpv_prescalers_b200 import HookError, HookFile

    with pytest.raises(HookError):
        HookFile.parse_text(_modified_hook_text()).variant


# ---- device ------------------------------------------------------------------------------------------------------------
torch = pytest.importorskip("torch")

RUN = [("ravu-lite-ar-r3.hook", None), ("ravu-r3.hook", None), ("ravu-r2-rgb.hook", None), ("ravu-r2-yuv.hook", None),
       ("ravu-zoom-ar-r2.hook", (96, 72)), ("nnedi3-nns16-win8x4.hook", None), ("gather/ravu-lite-ar-r2.hook", None),
       ("gather/ravu-r3.hook", None), ("gather/nnedi3-nns16-win8x4.hook", None), ("compute/ravu-lite-r2.hook", None),
       ("compute/ravu-lite-ar-r3.hook", None), ("compute/ravu-r3-rgb.hook", None), ("compute/ravu-3x-r2.hook", None),
       ("compute/ravu-3x-r3-rgb.hook", None), ("compute/nnedi3-nns16-win8x4.hook", None)]


def _close(got, ref, what):
    """The runner executes the same float32 operations in the same order as the literal CPU execution; what differs is
    the libm behind exp / log2 / atan (<= 2 ulp), which can move a key across a bucket edge in a rare pixel."""
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert got.shape == ref.shape, what
    assert np.mean(d > 1e-5) <= 5e-3 and np.median(d) <= 1e-6, f"{what}: {np.mean(d > 1e-5):.4f} of the pixels differ (max {d.max():.3e})"


@pytest.mark.gpu
@pytest.mark.parametrize("name,out_size", RUN)
def test_generic_runner_matches_literal_execution(name, out_size):
    from mpv_prescalers_b200 import HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from oracle.glsl_exec import run_hook

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    path = hook_path(name)
    hk = HookFile.parse(path)
    # the CPU interpreter covers the root flavours; a gather/ file is the same math with textureGatherOffset addressing,
    # so its device run is compared with the literal execution of its root twin
    ref_path = hook_path(name.split("/")[-1]) if name.startswith("gather/") else path
    c = 3 if ("rgb" in name or "yuv" in name) else 1
    x = batch(2, c, 24, 32, config=81)
    osz = None if out_size is None else (out_size[1], out_size[0])
    got = prescale(torch.from_numpy(x).cuda() if c == 3 else torch.from_numpy(x[:, 0]).cuda(), hk, output_size=osz, runner="generic")
    for f in range(2):
        img = x[f, 0] if c == 1 else np.moveaxis(x[f], 0, -1)
        ref, off, applied = run_hook(ref_path, img, **({"out_size": out_size} if out_size else {}))
        g = got[f].cpu().numpy()
        g = g if c == 1 else np.moveaxis(g, 0, -1)
        _close(g, ref, f"{name} frame {f}")
        assert tuple(got.offset) == tuple(off) and got.applied


@pytest.mark.gpu
def test_hand_edited_hook_runs_through_the_generic_runner(tmp_path):
    from mpv_prescalers_b200 import HookError, HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from oracle.glsl_exec import run_hook

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    fn = tmp_path / "ravu-lite-r3-edited.hook"
    fn.write_text(_modified_hook_text())
    hk = HookFile.parse(str(fn))
    x = batch(1, 1, 40, 56, config=82)
    xt = torch.from_numpy(x[:, 0]).cuda()
    with pytest.raises(HookError):
        prescale(xt, hk)
    got = prescale(xt, hk, runner="auto")
    ref, off, applied = run_hook(str(fn), x[0, 0])
    _close(got[0].cpu().numpy(), ref, "edited hook")
    plain = prescale(xt, hook_path("ravu-lite-r3.hook"), runner="auto")          # a known form: 'auto' takes the fused kernels
    assert plain.plan is not None and not torch.equal(plain, got)


@pytest.mark.gpu
def test_generic_runner_cross_checks_the_fused_kernels():
    """The file's own GLSL on the device against the hand-written kernels: same tolerances as the oracle tests."""
    from mpv_prescalers_b200 import prescale
    from mpv_prescalers_b200.synth import batch
    from tests.parity import psnr

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    x = torch.from_numpy(batch(2, 1, 135, 240, config=83)[:, 0]).cuda()
    for name, osz in (("ravu-lite-ar-r3.hook", None), ("ravu-r4.hook", None), ("nnedi3-nns32-win8x6.hook", None),
                      ("compute/ravu-3x-r4.hook", None), ("compute/ravu-zoom-r2.hook", (311, 541))):
        # (the compute flavour of ravu-zoom maps output texels through NAME_map(id), which the CPU interpreter does not model)
        a = prescale(x, hook_path(name), output_size=osz).cpu().numpy()
        b = prescale(x, hook_path(name), output_size=osz, runner="generic").cpu().numpy()
        d = np.abs(a - b)
        assert psnr(a, b) >= 60.0 and np.mean(d > 1e-3) <= 2e-4, f"{name}: {np.mean(d > 1e-3):.2e} of pixels differ, PSNR {psnr(a, b):.1f}"
