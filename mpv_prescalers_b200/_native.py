"""Build and load ``libmpvp.so`` (the C-ABI declared in ``include/mpvp.h``) through ctypes.

There is deliberately no CPU fallback: if the shared library is missing or cannot be loaded every
entry point raises :class:`NativeError`.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MPVP_LIB: load an experiment build (tools/build_variant.py) instead of the product library
LIB_PATH = os.environ.get("MPVP_LIB") or os.path.join(HERE, "libmpvp.so")
SOURCES = ["abi.cu", "ravu_lite.cu", "ravu_lite_ar.cu", "ravu_3x.cu", "ravu.cu", "ravu_zoom.cu", "nnedi3.cu", "nnedi3_tc.cu", "resample.cu", "host_entry.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class NativeError(RuntimeError):
    pass


def sources() -> List[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers() -> List[str]:
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "mpvp.h")]


def _obj_path(src: str) -> str:
    return os.path.join(HERE, "..", "build", "obj", os.path.basename(src) + ".o")


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def needs_build() -> bool:
    return _stale(LIB_PATH, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libmpvp.so`` (nvcc cross-compiles without a GPU).

    Each translation unit is compiled to an object in parallel (build/obj) and only when it or a header it may include
    is newer than its object; then everything is linked."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise NativeError("nvcc not found; cannot build libmpvp.so")
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(os.path.join(HERE, "..", "build", "obj"), exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    hdrs = _headers()

    def compile_one(src: str):
        obj = _obj_path(src)
        if not force and not _stale(obj, [src] + hdrs):
            return src, obj, None
        tmp_obj = obj + ".tmp"
        proc = subprocess.run([nvcc] + cflags + ["-c", "-o", tmp_obj, src], cwd=CSRC, capture_output=True, text=True)
        if proc.returncode == 0:
            os.replace(tmp_obj, obj)
        return src, obj, proc

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    for src, _, proc in results:
        if proc is None:
            continue
        if proc.returncode != 0:
            raise NativeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stderr)
    tmp = LIB_PATH + ".tmp"
    link = subprocess.run([nvcc, "-shared", "-o", tmp] + [o for _, o, _ in results] + ["-lcuda"], capture_output=True, text=True)
    if link.returncode != 0:
        raise NativeError("link failed:\n" + link.stdout + link.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class KeyParams(ctypes.Structure):
    _fields_ = [
        ("gauss", ctypes.c_float * 36),
        ("n_gauss", ctypes.c_int32),
        ("strength_thr", ctypes.c_float * 3),
        ("n_strength_thr", ctypes.c_int32),
        ("strength_log2_scale", ctypes.c_float),
        ("n_strength", ctypes.c_int32),
        ("coherence_thr", ctypes.c_float * 2),
        ("l1_thr", ctypes.c_float * 8),
        ("n_l1_thr", ctypes.c_int32),
        ("coh_ratio", ctypes.c_float * 2),
    ]


class IoDesc(ctypes.Structure):
    """mpvp_io: plane formats either side of a launch (include/mpvp.h)."""

    _fields_ = [("in_format", ctypes.c_int32), ("out_format", ctypes.c_int32), ("in_max", ctypes.c_float), ("out_max", ctypes.c_float)]


FMT_F32, FMT_F16, FMT_U8, FMT_U16 = 0, 1, 2, 3
SCALERS = {"bilinear": 0, "catmull_rom": 1, "mitchell": 2, "spline36": 3, "lanczos": 4}

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float
_kp = ctypes.POINTER(KeyParams)
_iop = ctypes.POINTER(IoDesc)

# name -> (restype, argtypes): must list every symbol include/mpvp.h declares
SIGNATURES = {
    "mpvp_last_error": (ctypes.c_char_p, []),
    "mpvp_abi_version": (_i, []),
    "mpvp_launch_count": (ctypes.c_uint64, []),
    "mpvp_debug_set_grid_limit": (_i, [_i]),
    "mpvp_key_params_finalize": (_i, [_kp]),
    "mpvp_weights_create_lut": (_i, [_i, _vp, _i, _i, _i, ctypes.POINTER(_vp)]),
    "mpvp_weights_create_nnedi3": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, ctypes.POINTER(_vp)]),
    "mpvp_weights_destroy": (_i, [_vp]),
    "mpvp_ravu_lite_launch": (_i, [_vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _vp]),
    "mpvp_ravu_launch": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "mpvp_ravu3x_launch": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "mpvp_ravu_zoom_launch": (_i, [_vp, _vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "mpvp_nnedi3_launch": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _vp]),
    "mpvp_ravu_lite_launch_io": (_i, [_vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _iop, _vp]),
    "mpvp_ravu_launch_io": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _iop, _vp]),
    "mpvp_ravu3x_launch_io": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _iop, _vp]),
    "mpvp_ravu_zoom_launch_io": (_i, [_vp, _vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _iop, _vp]),
    "mpvp_nnedi3_launch_io": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _iop, _vp]),
    "mpvp_resample_launch_io": (_i, [_i, _i, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _i64, _i64, _i64, _i64, _iop, _vp]),
    "mpvp_ravu_lite_host": (_i, [_vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i]),
    "mpvp_ravu_host": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _iop]),
    "mpvp_ravu3x_host": (_i, [_vp, _kp, _i, _i, _vp, _vp, _i, _i, _i, _iop]),
    "mpvp_ravu_zoom_host": (_i, [_vp, _vp, _kp, _i, _i, _f, _vp, _vp, _i, _i, _i, _i, _i, _iop]),
    "mpvp_nnedi3_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _iop]),
}

_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """The loaded library; raises NativeError (never falls back) if it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise NativeError(f"cannot load {LIB_PATH}: {e}") from None
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                raise NativeError(f"{LIB_PATH} does not export {name}") from None
            fn.restype = res
            fn.argtypes = args
        if handle.mpvp_abi_version() != 1:
            raise NativeError("libmpvp.so ABI version mismatch; rebuild")
        _lib = handle
        return handle


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mpvp_last_error()
        raise NativeError(f"{what} failed (code {rc}): {msg.decode(errors='replace') if msg else '?'}")
