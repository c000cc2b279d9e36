"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/mpvp.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mpvp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpvp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for want in ("mpvp_last_error", "mpvp_weights_create_lut", "mpvp_weights_create_nnedi3", "mpvp_weights_destroy",
                 "mpvp_ravu_lite_launch", "mpvp_ravu_launch", "mpvp_ravu3x_launch", "mpvp_ravu_zoom_launch",
                 "mpvp_nnedi3_launch", "mpvp_ravu_lite_host"):
        assert want in names


def test_library_builds_loads_and_exports_everything():
    from mpv_prescalers_b200 import _native

    _native.build()
    lib = _native.lib()
    assert lib.mpvp_abi_version() == 1
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"libmpvp.so does not export {name}"
    assert sorted(_native.SIGNATURES) == _declared()
    assert lib.mpvp_last_error() is not None


def test_struct_layout_matches_header():
    from mpv_prescalers_b200 import _native

    # float[36] + int + float[3] + int + float + int + float[2] + float[8] + int + float[2] = 56 4-byte fields
    assert ctypes.sizeof(_native.KeyParams) == 56 * 4


def test_product_path_fails_loudly_without_gpu():
    import torch

    from mpv_prescalers_b200 import _native, prescale
    from tests.conftest import hook_path

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        prescale(torch.zeros(1, 16, 16), hook_path("ravu-lite-r3.hook"))


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under mpv_prescalers_b200/ may reference it."""
    pkg = os.path.join(ROOT, "mpv_prescalers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_plane_io_descriptor_rules():
    """PlaneIO (the Python face of mpvp_io): default output dtype follows the input, depths are validated."""
    import pytest
    import torch

    from mpv_prescalers_b200 import _native
    from mpv_prescalers_b200.api import PlaneIO

    io = PlaneIO()
    assert io.is_default and io.desc().in_format == _native.FMT_F32 and io.desc().out_format == _native.FMT_F32
    io = PlaneIO(torch.uint8)
    assert io.out_dtype == torch.uint8 and io.in_max == 255.0 and io.out_max == 255.0
    io = PlaneIO(torch.uint16, bit_depth=10)
    assert io.in_max == 1023.0 and io.out_max == 1023.0 and io.desc().in_format == _native.FMT_U16
    io = PlaneIO(torch.uint16, torch.uint8, bit_depth=10)
    assert io.out_max == 255.0                      # a 10-bit depth does not fit uint8: full 8-bit range
    io = PlaneIO(torch.uint8, torch.float16)
    d = io.desc()
    assert (d.in_format, d.out_format) == (_native.FMT_U8, _native.FMT_F16)
    # chains: only the first launch reads the integer plane, only the last one writes the output format
    mid = io.desc(first=True, last=False)
    assert (mid.in_format, mid.out_format) == (_native.FMT_U8, _native.FMT_F32)
    with pytest.raises(ValueError):
        PlaneIO(torch.uint8, bit_depth=10)
    with pytest.raises(TypeError):
        PlaneIO(torch.int32)


@pytest.mark.parametrize("name", ["ravu-lite-ar-r3.hook", "ravu-r3.hook", "ravu-r4.hook", "compute/ravu-3x-r2.hook", "ravu-zoom-r2.hook"])
def test_key_params_finalize_reproduces_the_python_tables(name):
    """mpvp_key_params_finalize (pure host arithmetic) derives the same l1_thr[] / coh_ratio[] from the shader constants
    as the NumPy bisection the Python host uses, so a C caller needs nothing from the Python package."""
    from mpv_prescalers_b200 import HookFile, _native
    from mpv_prescalers_b200.api import _key_params
    from tests.conftest import hook_path

    v = HookFile.parse(hook_path(name)).variant
    want = _key_params(v)
    kp = _key_params(v)
    kp.n_l1_thr = 0
    for i in range(8):
        kp.l1_thr[i] = 0.0
    kp.coh_ratio[0] = kp.coh_ratio[1] = 0.0
    rc = _native.lib().mpvp_key_params_finalize(ctypes.byref(kp))
    assert rc == 0, _native.lib().mpvp_last_error()
    assert kp.n_l1_thr == want.n_l1_thr == v.n_strength - 1
    assert list(kp.l1_thr) == list(want.l1_thr)
    assert list(kp.coh_ratio) == list(want.coh_ratio)


def test_key_params_finalize_rejects_nonsense():
    from mpv_prescalers_b200 import _native

    kp = _native.KeyParams()
    assert _native.lib().mpvp_key_params_finalize(ctypes.byref(kp)) < 0
    assert b"n_strength" in _native.lib().mpvp_last_error()
