"""``prescale()`` -- the Python face of the drop-in boundary.

The reference is used as ``mpv --glsl-shader=ravu-lite-ar-r3.hook`` (``README.md:33-37``): the host
parses the file, uploads its ``//!TEXTURE`` LUTs once and runs the passes whose ``//!WHEN`` holds on
every frame.  ``prescale(frames, hook=...)`` does the same for a batch of frames held in a
``torch.Tensor``: parse (cached) -> plan (WHEN / WIDTH / HEIGHT / OFFSET evaluation) -> upload weights
(cached per device) -> one fused CUDA kernel per family through the C ABI of ``include/mpvp.h``.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in ``libmpvp.so``.
There is no CPU fallback: without a CUDA device or without the library this raises.
"""
from __future__ import annotations

import ctypes
import dataclasses
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native
from .hookfile import HookError, HookFile, Variant, find_hook

__all__ = ["prescale", "resample", "plan", "Plan", "upload_weights", "clear_weight_cache"]


# ----------------------------------------------------------------------------------------------
# planning (host logic, no GPU needed)
# ----------------------------------------------------------------------------------------------


@dataclass
class Plan:
    """What one application of a hook file does to a (h, w) plane."""

    family: str
    in_size: Tuple[int, int]  # (h, w)
    out_size: Tuple[int, int]  # (h, w) actually produced
    applied: bool
    offset: Tuple[float, float]  # accumulated //!OFFSET (x, y) in output pixels
    double_y: bool = False  # nnedi3 only
    double_x: bool = False
    passes: Tuple[str, ...] = ()  # DESC strings of the passes that fire


def plan(hook: HookFile, in_hw: Tuple[int, int], output_size: Optional[Tuple[int, int]] = None, is_yuv: bool = True) -> Plan:
    """Evaluate WHEN / WIDTH / HEIGHT / OFFSET of every pass for an (h, w) input.

    ``output_size`` is mpv's OUTPUT (the final target, (h, w)); ``None`` means the hook's natural
    factor, for which every WHEN of the shipped files holds.  mpv semantics: a pass whose WHEN is false
    is skipped silently; if none fires the input is returned unchanged (``applied=False``).
    """
    v = hook.variant
    h, w = int(in_hw[0]), int(in_hw[1])
    if output_size is None:
        if v.scale is None:
            raise HookError(f"{hook.name}: ravu-zoom needs output_size=(h, w)")
        oh, ow = h * v.scale, w * v.scale
    else:
        oh, ow = int(output_size[0]), int(output_size[1])
    # a player asks the same question for every frame: the answer is kept on the (immutable) parsed file
    cache = hook.__dict__.setdefault("_plan_cache", {})
    ckey = (h, w, oh, ow, bool(is_yuv))
    hit = cache.get(ckey)
    if hit is not None:
        return hit
    pl = _plan_uncached(hook, v, h, w, oh, ow, is_yuv)
    if len(cache) < 64:
        cache[ckey] = pl
    return pl


def _plan_uncached(hook: HookFile, v: Variant, h: int, w: int, oh: int, ow: int, is_yuv: bool) -> Plan:
    env = {"HOOKED": (w, h), "OUTPUT": (ow, oh), "LUMA": (w if is_yuv else 0, h if is_yuv else 0), "NATIVE": (w, h), "MAIN": (w, h)}
    saved: Dict[str, Tuple[int, int]] = {}
    fired: List[str] = []
    off = [0.0, 0.0]
    dy = dx = False
    for p in hook.passes:
        e = dict(env)
        e.update(saved)
        if not p.enabled(e):
            continue
        size = p.output_size(e)
        fired.append(p.desc)
        if p.save:
            saved[p.save] = size
        else:
            env["HOOKED"] = size
        if isinstance(p.offset, tuple):
            off[0] += p.offset[0]
            off[1] += p.offset[1]
        if "double_y" in p.desc:
            dy = True
        if "double_x" in p.desc:
            dx = True
    cw, ch = env["HOOKED"]
    applied = bool(fired)
    # a chain whose convolution passes were skipped (e.g. the LUMA.w guard of -yuv hooks on non-YUV
    # input, ravu-r2-yuv.hook:20 vs :326) has nothing to merge: mpv would fail the pass, we skip the hook
    need = {"ravu": 4, "ravu-lite": 2}.get(v.family)
    if need is not None and v.flavour == "root" and len(fired) != need:
        applied = False
    if not applied:
        return Plan(v.family, (h, w), (h, w), False, (0.0, 0.0), False, False, ())
    return Plan(v.family, (h, w), (ch, cw), True, (off[0], off[1]), dy, dx, tuple(fired))


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------


class _Weights:
    """Device-resident weights of one hook on one device (opaque C handles)."""

    def __init__(self, device: int):
        self.device = device
        self.handles: Dict[str, ctypes.c_void_p] = {}
        self.key = _native.KeyParams()

    def close(self):
        lib = _native.lib()
        for h in self.handles.values():
            lib.mpvp_weights_destroy(h)
        self.handles.clear()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass


_wcache: Dict[Tuple[str, int, str], _Weights] = {}
_wlock = threading.Lock()


def _key_params(v: Variant) -> _native.KeyParams:
    kp = _native.KeyParams()
    g = np.asarray(v.gauss, dtype=np.float32)
    for i, val in enumerate(g):
        kp.gauss[i] = float(val)
    kp.n_gauss = len(g)
    for i, t in enumerate(v.strength_thr):
        kp.strength_thr[i] = float(t)
    kp.n_strength_thr = len(v.strength_thr)
    kp.strength_log2_scale = float(v.strength_log2_scale)
    kp.n_strength = int(v.n_strength)
    kp.coherence_thr[0], kp.coherence_thr[1] = (float(t) for t in v.coherence_thr)
    thr = l1_thresholds(v)
    for i, t in enumerate(thr):
        kp.l1_thr[i] = float(t)
    kp.n_l1_thr = len(thr)
    for i, c in enumerate(v.coherence_thr):
        kp.coh_ratio[i] = float(np.float32(((1.0 + c) / (1.0 - c)) ** 2))
    return kp


def _strength_of_l1(v: Variant, x: np.ndarray) -> np.ndarray:
    """The shader's strength expression (float32, op for op) as a function of the eigenvalue L1."""
    f32 = np.float32
    lam = np.sqrt(x.astype(f32))
    if v.strength_thr:
        s = np.zeros(lam.shape, np.int32)
        for t in v.strength_thr:
            s += lam >= f32(t)
        return s
    with np.errstate(divide="ignore"):
        val = np.floor(np.log2(lam * f32(v.strength_log2_scale) + f32(1.192092896e-7)))
    return np.clip(val, 0, v.n_strength - 1).astype(np.int32)


def l1_thresholds(v: Variant) -> List[np.float32]:
    """For each strength level k >= 1 the smallest positive float32 L1 whose strength is >= k.

    The strength quantiser is monotone in lambda = sqrt(L1) and sqrt is correctly rounded, so the set
    {L1 : strength(L1) >= k} is an interval [x_k, inf): comparing L1 against x_k reproduces the shader's
    decision exactly, without taking the square root on the device.  Found by bisection over the
    float32 bit patterns."""
    out = []
    for k in range(1, v.n_strength):
        lo, hi = 0, 0x7F7FFFFF  # bit patterns of +0 and FLT_MAX
        while lo < hi:
            mid = (lo + hi) // 2
            x = np.array([mid], dtype=np.uint32).view(np.float32)
            if _strength_of_l1(v, x)[0] >= k:
                hi = mid
            else:
                lo = mid + 1
        out.append(np.array([lo], dtype=np.uint32).view(np.float32)[0])
    return out


def upload_weights(hook: HookFile, device: int, lut_precision: str = "fp16") -> _Weights:
    if lut_precision not in ("fp16", "fp32"):
        raise ValueError("lut_precision must be 'fp16' or 'fp32'")
    # keyed on CONTENT (LUT payloads / NNEDI3 weights + key constants), not on the path: every parse_text() hook has
    # the path '<string>', and an edited file keeps its path
    key = (hook.content_key, device, lut_precision)
    with _wlock:
        hit = _wcache.get(key)
        if hit is not None:
            return hit
        lib = _native.lib()
        v = hook.variant
        W = _Weights(device)
        if v.family == "nnedi3":
            for name, nn in (("y", v.nn_y), ("x", v.nn_x)):
                arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (nn.w1, nn.w2, nn.b1, nn.b2)]
                h = ctypes.c_void_p()
                rc = lib.mpvp_weights_create_nnedi3(device, *(a.ctypes.data for a in arrs), v.nns, v.win[1], ctypes.byref(h))
                _native.check(rc, "mpvp_weights_create_nnedi3")
                W.handles[name] = h
        else:
            W.key = _key_params(v)
            for name, tex in (("lut", v.lut), ("lut_ar", v.lut_ar)):
                if tex is None:
                    continue
                data = np.ascontiguousarray(tex.data, dtype=np.float32)
                h = ctypes.c_void_p()
                rc = lib.mpvp_weights_create_lut(device, data.ctypes.data, tex.width, tex.height, 1 if lut_precision == "fp16" else 0, ctypes.byref(h))
                _native.check(rc, "mpvp_weights_create_lut")
                W.handles[name] = h
        _wcache[key] = W
        return W


def clear_weight_cache() -> None:
    with _wlock:
        for W in _wcache.values():
            W.close()
        _wcache.clear()


# ----------------------------------------------------------------------------------------------
# launch
# ----------------------------------------------------------------------------------------------


_FMT_OF = {torch.float32: _native.FMT_F32, torch.float16: _native.FMT_F16, torch.uint8: _native.FMT_U8, torch.uint16: _native.FMT_U16}


class PlaneIO:
    """Plane formats either side of the path (``mpvp_io`` in include/mpvp.h).

    Integer planes are UNORM video planes: a raw sample r of a ``bit_depth``-bit plane stands for
    ``r / (2**bit_depth - 1)`` -- the host's texture normalisation times ``HOOKED_mul``
    (gather/ravu-lite-ar-r3.hook:23) -- and an integer result is ``rint(clamp(v, 0, 1) * (2**bit_depth - 1))``.
    float16 stands for the rgba16f FBO precision mpv keeps between passes."""

    def __init__(self, in_dtype=torch.float32, out_dtype=None, bit_depth: Optional[int] = None, out_bit_depth: Optional[int] = None):
        out_dtype = in_dtype if out_dtype is None else out_dtype
        for d in (in_dtype, out_dtype):
            if d not in _FMT_OF:
                raise TypeError(f"plane dtype {d} not supported (float32, float16, uint8, uint16)")
        self.in_dtype, self.out_dtype = in_dtype, out_dtype

        def depth(dt, bits, what):
            if dt not in (torch.uint8, torch.uint16):
                return 0
            full = 8 if dt == torch.uint8 else 16
            bits = full if bits is None else int(bits)
            if not 1 <= bits <= full:
                raise ValueError(f"{what} {bits} does not fit {dt}")
            return bits

        self.in_bits = depth(in_dtype, bit_depth, "bit_depth")
        if out_bit_depth is None and out_dtype in (torch.uint8, torch.uint16):
            out_bit_depth = self.in_bits if (self.in_bits and self.in_bits <= (8 if out_dtype == torch.uint8 else 16)) else None
        self.out_bits = depth(out_dtype, out_bit_depth, "out_bit_depth")
        self.in_max = float((1 << self.in_bits) - 1) if self.in_bits else 1.0
        self.out_max = float((1 << self.out_bits) - 1) if self.out_bits else 1.0

    @property
    def is_default(self) -> bool:
        return self.in_dtype == torch.float32 and self.out_dtype == torch.float32

    def desc(self, first: bool = True, last: bool = True) -> "_native.IoDesc":
        """mpvp_io of one launch of a chain: only the first launch reads in_dtype, only the last writes out_dtype
        (NNEDI3's W x 2H image between its two launches stays float32)."""
        return _native.IoDesc(
            _FMT_OF[self.in_dtype] if first else _native.FMT_F32, _FMT_OF[self.out_dtype] if last else _native.FMT_F32,
            self.in_max if first else 1.0, self.out_max if last else 1.0)


_IO_F32 = PlaneIO()


def _launch(hook: HookFile, pl: Plan, x: torch.Tensor, W: _Weights, want_buckets: bool, io: Optional[PlaneIO] = None):
    """x: CUDA tensor [N, C, H, W] (dtype io.in_dtype) contiguous in (H, W).  Returns (out, buckets|None)."""
    io = _IO_F32 if io is None else io
    if x.dtype != io.in_dtype:
        raise TypeError(f"frames are {x.dtype}, the plane format says {io.in_dtype}")
    iod = io.desc()
    lib = _native.lib()
    v = hook.variant
    dev = x.device.index
    n, c, h, w = x.shape
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    oh, ow = pl.out_size
    key_mode = {"luma": 0, "yuv": 1, "rgb": 2}[v.plane]
    bk = None
    bptr = None

    def out_tensor(hh, ww):
        return torch.empty((n, c, hh, ww), dtype=io.out_dtype, device=x.device)

    if v.family == "ravu-lite":
        out = out_tensor(oh, ow)
        if want_buckets:
            bk = torch.empty((n, h, w), dtype=torch.int32, device=x.device)
            bptr = bk.data_ptr()
        rc = lib.mpvp_ravu_lite_launch_io(
            W.handles["lut"], ctypes.byref(W.key), v.radius, 1 if v.ar else 0, float(v.ar_strength),
            x.data_ptr(), out.data_ptr(), n, h, w, x.stride(0), x.stride(2), out.stride(0), out.stride(2), bptr,
            ctypes.byref(iod), stream)
        _native.check(rc, "mpvp_ravu_lite_launch_io")
        return out, bk
    if v.family in ("ravu", "ravu-3x"):
        out = out_tensor(oh, ow)
        if want_buckets:
            bk = torch.empty((n, 3, h, w) if v.family == "ravu" else (n, h, w), dtype=torch.int32, device=x.device)
            bptr = bk.data_ptr()
        fn = lib.mpvp_ravu_launch_io if v.family == "ravu" else lib.mpvp_ravu3x_launch_io
        rc = fn(W.handles["lut"], ctypes.byref(W.key), v.radius, key_mode, x.data_ptr(), out.data_ptr(), n, h, w,
                x.stride(0), x.stride(1), x.stride(2), out.stride(0), out.stride(1), out.stride(2), bptr,
                ctypes.byref(iod), stream)
        _native.check(rc, "mpvp_ravu_launch_io" if v.family == "ravu" else "mpvp_ravu3x_launch_io")
        return out, bk
    if v.family == "ravu-zoom":
        out = out_tensor(oh, ow)
        if want_buckets:
            bk = torch.empty((n, oh, ow), dtype=torch.int32, device=x.device)
            bptr = bk.data_ptr()
        rc = lib.mpvp_ravu_zoom_launch_io(
            W.handles["lut"], W.handles.get("lut_ar"), ctypes.byref(W.key), v.radius, key_mode, float(v.ar_strength),
            x.data_ptr(), out.data_ptr(), n, h, w, oh, ow, x.stride(0), x.stride(1), x.stride(2),
            out.stride(0), out.stride(1), out.stride(2), bptr, ctypes.byref(iod), stream)
        _native.check(rc, "mpvp_ravu_zoom_launch_io")
        return out, bk
    if v.family == "nnedi3":
        cur = x.reshape(n * c, h, w)
        passes = [d for d, on in ((0, pl.double_y), (1, pl.double_x)) if on]
        for k, d in enumerate(passes):
            first, last = k == 0, k == len(passes) - 1
            shape = (n * c, 2 * cur.shape[1], cur.shape[2]) if d == 0 else (n * c, cur.shape[1], 2 * cur.shape[2])
            nxt = torch.empty(shape, dtype=io.out_dtype if last else torch.float32, device=x.device)
            pio = io.desc(first, last)
            rc = lib.mpvp_nnedi3_launch_io(W.handles["y" if d == 0 else "x"], d, cur.data_ptr(), nxt.data_ptr(), cur.shape[0],
                                           cur.shape[1], cur.shape[2], cur.stride(0), cur.stride(1), nxt.stride(0), nxt.stride(1),
                                           ctypes.byref(pio), stream)
            _native.check(rc, f"mpvp_nnedi3_launch_io({'y' if d == 0 else 'x'})")
            cur = nxt
        return cur.reshape(n, c, cur.shape[1], cur.shape[2]), None
    raise HookError(f"unsupported family {v.family}")


def _normalise_input(frames: torch.Tensor, v: Variant) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    if not isinstance(frames, torch.Tensor):
        raise TypeError("frames must be a torch.Tensor")
    if frames.dtype not in _FMT_OF:
        raise TypeError(f"frames must be float32 / float16 in [0, 1] or uint8 / uint16 video planes (got {frames.dtype})")
    shape = tuple(frames.shape)
    c = v.channels
    if c == 1:
        if frames.dim() == 2:
            x = frames[None, None]
        elif frames.dim() == 3:
            x = frames[:, None]
        elif frames.dim() == 4 and frames.shape[1] == 1:
            x = frames
        else:
            raise ValueError(f"luma hook expects [H,W], [N,H,W] or [N,1,H,W]; got {shape}")
    else:
        if frames.dim() == 3 and frames.shape[0] == 3:
            x = frames[None]
        elif frames.dim() == 4 and frames.shape[1] == 3:
            x = frames
        else:
            raise ValueError(f"{v.plane} hook expects planar [3,H,W] or [N,3,H,W]; got {shape}")
    if x.shape[2] < 1 or x.shape[3] < 1:
        raise ValueError(f"empty plane {shape}")
    return x, shape


def _restore_shape(out: torch.Tensor, in_shape: Tuple[int, ...], c: int) -> torch.Tensor:
    if c == 1:
        if len(in_shape) == 2:
            return out[0, 0]
        if len(in_shape) == 3:
            return out[:, 0]
        return out
    return out[0] if len(in_shape) == 3 else out


def prescale(
    frames: torch.Tensor,
    hook: Union[str, "HookFile"] = "ravu-lite-ar-r3.hook",
    output_size: Optional[Tuple[int, int]] = None,
    devices: Optional[Sequence[Union[int, str, torch.device]]] = None,
    lut_precision: str = "fp16",
    return_buckets: bool = False,
    is_yuv: bool = True,
    out: Optional[torch.Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    bit_depth: Optional[int] = None,
    out_bit_depth: Optional[int] = None,
    split: str = "frames",
    correct_offset: Union[bool, str] = False,
    runner: str = "fused",
):
    """Apply one mpv-prescalers hook file to a batch of frames.

    frames        float32 in [0,1]; luma hooks: ``[H,W]``, ``[N,H,W]`` or ``[N,1,H,W]``; ``-yuv``/``-rgb``
                  hooks: planar ``[3,H,W]`` or ``[N,3,H,W]``.  CUDA tensors are processed in place on their
                  device and the result stays there; CPU tensors are staged through the GPU (host->device
                  copy, kernel, device->host copy) and a CPU tensor is returned.
    hook          path or bare name of a shipped ``.hook`` file (root, ``gather/`` or ``compute/`` flavour).
    output_size   mpv's OUTPUT size ``(h, w)`` used by the ``//!WHEN`` conditions; required for ravu-zoom
                  (it is the size produced); ``None`` = the hook's natural 2x / 3x.
    devices       shard the batch dimension over these GPUs (frames are independent, no collective);
                  returns a list with one output tensor per device (outputs stay on their GPU).
    split         with ``devices``: 'frames' (default) or 'rows' -- split every frame into row bands with a halo,
                  one band per listed device, and return the re-assembled result on ``devices[0]`` (for a single
                  large frame; halo rows travel peer-to-peer over NVLink; bit-identical to the unsplit result).
    lut_precision 'fp16' reproduces the reference's rgba16f LUT storage; 'fp32' keeps the file's floats.
    out           optional destination for CPU inputs: a (pinned) CPU tensor ``[N,C,OH,OW]`` of the output dtype
                  that receives the result, so that a video loop does not re-allocate pinned memory per batch.
    out_dtype     dtype of the result (default: the dtype of ``frames``).  Besides float32 the path reads and writes
                  the wire formats of video directly: uint8 / uint16 UNORM planes (``bit_depth`` significant bits,
                  e.g. 10 for 10-bit video in 16-bit containers; a raw sample r means r / (2**bit_depth - 1), the
                  host's texture normalisation times ``HOOKED_mul``) and float16 (mpv's rgba16f FBO precision).
                  Integer results are ``rint(clamp(v, 0, 1) * (2**out_bit_depth - 1))``.

    correct_offset  ravu and nnedi3 leave their result half a texel off (``//!OFFSET -0.5 -0.5``, ravu-r2.hook:325) and
                  rely on the host's main scaler to compensate.  ``True`` (= 'lanczos') or the name of an mpv scaler
                  kernel ('bilinear', 'catmull_rom', 'mitchell', 'spline36', 'lanczos') runs that step (``resample()``,
                  same size, shifted by the accumulated offset) so that the result is aligned like ravu-lite's and
                  ``.offset`` is (0, 0); hooks without an offset are returned as they are.

    runner        'fused' (default): the hand-written kernels, which accept the shipped files and refuse anything else
                  (``HookError``); 'generic': transpile the file's GLSL to CUDA at load time (NVRTC) and run it pass by
                  pass (``hookrunner.GenericHook``: any fragment-shader hook within the GLSL subset, e.g. a hand-modified
                  RAVU; float32 planes only, slow); 'auto': fused if the file is a known form, else generic.

    Returns the output tensor (same rank as the input) carrying ``.offset`` (accumulated ``//!OFFSET``,
    (x, y) in output pixels), ``.applied`` and ``.plan``; with ``return_buckets=True`` a pair
    ``(out, buckets)`` where ``buckets`` holds the LUT row of every key evaluation (RAVU families).
    """
    hk = hook if isinstance(hook, HookFile) else HookFile.parse(find_hook(hook))
    if runner not in ("fused", "generic", "auto"):
        raise ValueError("runner must be 'fused', 'generic' or 'auto'")
    if runner == "auto":
        try:
            hk.variant
            runner = "fused"
        except HookError:
            runner = "generic"
    if runner == "generic":
        return _prescale_generic(frames, hk, output_size, lut_precision, is_yuv, correct_offset)
    v = hk.variant
    scaler = None
    if correct_offset:
        scaler = "lanczos" if correct_offset is True else str(correct_offset)
        if scaler not in _native.SCALERS:
            raise ValueError(f"correct_offset must be True or one of {sorted(_native.SCALERS)}")
    if devices is not None:
        from .sharding import prescale_rowsplit, prescale_sharded

        if return_buckets or out is not None:
            raise ValueError("return_buckets / out cannot be combined with devices=[...] (sharded results stay on their GPUs)")
        if split == "rows":
            # the bands are gathered first: the correction filter must see real neighbouring rows across the seams
            mid_dtype = torch.float32 if scaler else out_dtype
            res = prescale_rowsplit(frames, hk, output_size, list(devices), lut_precision, is_yuv,
                                    out_dtype=mid_dtype, bit_depth=bit_depth, out_bit_depth=None if scaler else out_bit_depth)
            if scaler and getattr(res, "applied", False) and tuple(res.offset) != (0.0, 0.0):
                pl0 = res.plan
                want = out_dtype if out_dtype is not None else frames.dtype
                res = resample(res, None, res.offset, scaler, out_dtype=want,
                               out_bit_depth=out_bit_depth if out_bit_depth is not None else bit_depth)
                res.offset, res.applied, res.plan = (0.0, 0.0), True, pl0
            return res
        if split != "frames":
            raise ValueError("split must be 'frames' or 'rows'")
        return prescale_sharded(frames, hk, output_size, list(devices), lut_precision, is_yuv,
                                out_dtype=out_dtype, bit_depth=bit_depth, out_bit_depth=out_bit_depth,
                                correct_offset=correct_offset)
    x, in_shape = _normalise_input(frames, v)
    n, c, h, w = x.shape
    io = PlaneIO(x.dtype, out_dtype, bit_depth, out_bit_depth)
    pl = plan(hk, (h, w), output_size, is_yuv)
    if not pl.applied:
        # mpv semantics: no pass fired, the plane goes on unchanged.  The caller still gets a tensor of its own (never
        # an alias it could mutate the input through), of the dtype / in the destination it asked for.
        if io.out_dtype != io.in_dtype:
            raise ValueError(f"{hk.name}: no pass fires for this geometry (//!WHEN), so the plane stays {io.in_dtype}; "
                             f"it cannot be returned as {io.out_dtype}")
        if out is not None:
            if tuple(out.shape) != tuple(x.shape) or out.dtype != x.dtype:
                raise ValueError(f"out must be a {x.dtype} tensor of shape {tuple(x.shape)} (no pass fires: the result is the input)")
            out.copy_(x)
            res = _restore_shape(out, in_shape, c)
        else:
            res = frames.clone()
        res.offset, res.applied, res.plan = (0.0, 0.0), False, pl
        return (res, None) if return_buckets else res
    if not torch.cuda.is_available():
        raise _native.NativeError("prescale() needs a CUDA device: there is no CPU fallback")
    host_input = x.device.type != "cuda"
    if scaler and tuple(pl.offset) != (0.0, 0.0):
        # hook -> float32 plane -> offset-correcting scaler -> the requested plane format (one quantisation, at the end)
        dev = torch.device("cuda", torch.cuda.current_device()) if host_input else x.device
        xd = x.to(dev, non_blocking=True) if host_input else (x if (x.stride(3) == 1 and x.stride(2) >= w) else x.contiguous())
        W = upload_weights(hk, dev.index, lut_precision)
        mid = PlaneIO(x.dtype, torch.float32, bit_depth, None)
        with torch.cuda.device(dev):
            res, bk = _launch(hk, pl, xd, W, return_buckets, mid)
            res = resample(res, None, pl.offset, scaler, out_dtype=io.out_dtype, out_bit_depth=io.out_bits or None)
        if host_input:
            if out is not None:
                out.copy_(res)
                res = out
            else:
                res = res.cpu()
            bk = bk.cpu() if bk is not None else None
        res = _restore_shape(res, in_shape, c)
        res.offset, res.applied, res.plan = (0.0, 0.0), True, pl
        return (res, bk) if return_buckets else res
    if host_input:
        dev = torch.device("cuda", torch.cuda.current_device())
        W = upload_weights(hk, dev.index, lut_precision)
        res, bk = _prescale_host(hk, pl, x, W, dev, return_buckets, out, io)
    else:
        dev = x.device
        xd = x if (x.stride(3) == 1 and x.stride(2) >= w) else x.contiguous()
        W = upload_weights(hk, dev.index, lut_precision)
        with torch.cuda.device(dev):
            res, bk = _launch(hk, pl, xd, W, return_buckets, io)
    res = _restore_shape(res, in_shape, c)
    res.offset, res.applied, res.plan = pl.offset, True, pl
    return (res, bk) if return_buckets else res


def _prescale_generic(frames: torch.Tensor, hk: HookFile, output_size, lut_precision: str, is_yuv: bool, correct_offset):
    """runner='generic': the file's own GLSL, transpiled and compiled at load time (hookrunner.py)."""
    from .hookrunner import GenericHook

    if not torch.cuda.is_available():
        raise _native.NativeError("prescale() needs a CUDA device: there is no CPU fallback")
    if frames.dtype != torch.float32:
        raise TypeError("runner='generic' takes float32 planes")
    gh = getattr(hk, "_generic_runner", None)
    if gh is None or gh.lut_precision != lut_precision:
        gh = hk._generic_runner = GenericHook(hk, lut_precision)
    host = frames.device.type != "cuda"
    x = frames.cuda() if host else frames
    squeeze = x.dim() == 2
    if squeeze:
        x = x[None]
    if output_size is None and any("3x" in (p.desc or "").lower() for p in hk.passes):
        output_size = (3 * x.shape[-2], 3 * x.shape[-1])     # the natural target of a tripling hook (run() assumes 2x)
    res, off = gh.run(x, output_size, is_yuv)
    applied = tuple(res.shape) != tuple(x.shape) or off != (0.0, 0.0) or not torch.equal(res, x)
    if correct_offset and off != (0.0, 0.0):
        res = resample(res, None, off, "lanczos" if correct_offset is True else str(correct_offset))
        off = (0.0, 0.0)
    if squeeze:
        res = res[0]
    if host:
        res = res.cpu()
    res.offset, res.applied, res.plan = off, applied, None
    return res


def resample(frames: torch.Tensor, output_size: Optional[Tuple[int, int]] = None, offset: Tuple[float, float] = (0.0, 0.0),
             kernel: str = "lanczos", out_dtype: Optional[torch.dtype] = None, bit_depth: Optional[int] = None,
             out_bit_depth: Optional[int] = None) -> torch.Tensor:
    """The step after the hook: mpv's main scaler, which also absorbs the accumulated
    ``//!OFFSET`` of the hooked plane (ravu-r2.hook:325, nnedi3-nns16-win8x4.hook:95,185).

    frames       ``[H,W]``, ``[N,H,W]`` or ``[N,C,H,W]`` planes (float32 / float16 / uint8 / uint16) on a CUDA device, or
                 on the host (staged through the current GPU);
    output_size  ``(h, w)``; ``None`` keeps the size (pure offset correction);
    offset       ``(x, y)`` in texels of ``frames``: the ``.offset`` a ``prescale()`` result carries;
    kernel       'bilinear', 'catmull_rom', 'mitchell', 'spline36' or 'lanczos' (mpv's filter kernels, not widened when
                 downscaling, which is mpv's default).

    Output pixel o of an axis samples the input at ``(o + 0.5) * in / out - 0.5 + offset`` (clamp-to-edge)."""
    if kernel not in _native.SCALERS:
        raise ValueError(f"kernel must be one of {sorted(_native.SCALERS)}")
    if not isinstance(frames, torch.Tensor) or frames.dtype not in _FMT_OF:
        raise TypeError("frames must be a float32 / float16 / uint8 / uint16 torch.Tensor")
    if frames.dim() not in (2, 3, 4):
        raise ValueError(f"expected [H,W], [N,H,W] or [N,C,H,W]; got {tuple(frames.shape)}")
    if not torch.cuda.is_available():
        raise _native.NativeError("resample() needs a CUDA device: there is no CPU fallback")
    host = frames.device.type != "cuda"
    dev = torch.device("cuda", torch.cuda.current_device()) if host else frames.device
    x = frames.to(dev, non_blocking=True) if host else frames
    h, w = x.shape[-2], x.shape[-1]
    if h < 1 or w < 1:
        raise ValueError("empty plane")
    lead = tuple(x.shape[:-2])
    x3 = x.reshape((-1, h, w))
    if not (x3.stride(2) == 1 and x3.stride(1) >= w):
        x3 = x3.contiguous()
    oh, ow = (h, w) if output_size is None else (int(output_size[0]), int(output_size[1]))
    io = PlaneIO(x3.dtype, out_dtype, bit_depth, out_bit_depth)
    outp = torch.empty((x3.shape[0], oh, ow), dtype=io.out_dtype, device=dev)
    iod = io.desc()
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = _native.lib().mpvp_resample_launch_io(
            dev.index, _native.SCALERS[kernel], x3.data_ptr(), outp.data_ptr(), x3.shape[0], h, w, oh, ow,
            float(offset[0]), float(offset[1]), x3.stride(0), x3.stride(1), outp.stride(0), outp.stride(1),
            ctypes.byref(iod), stream)
    _native.check(rc, "mpvp_resample_launch_io")
    res = outp.reshape(lead + (oh, ow))
    return res.cpu() if host else res


_host_streams: Dict[int, List[torch.cuda.Stream]] = {}


def _prescale_host(hk: HookFile, pl: Plan, x: torch.Tensor, W: _Weights, dev: torch.device, want_buckets: bool,
                   out: Optional[torch.Tensor], io: Optional["PlaneIO"] = None):
    """CPU tensor in, CPU tensor out: frames are staged through the GPU in chunks on two streams so that the
    host->device copy of chunk k+1 and the device->host copy of chunk k-1 overlap the kernel of chunk k
    (effective only with pinned host memory)."""
    n, c, h, w = x.shape
    oh, ow = pl.out_size
    if hk.variant.family == "nnedi3":
        oh, ow = h * (2 if pl.double_y else 1), w * (2 if pl.double_x else 1)
    x = x.contiguous()
    io = _IO_F32 if io is None else io
    if out is None:
        out = torch.empty((n, c, oh, ow), dtype=io.out_dtype, pin_memory=True)
    elif tuple(out.shape) != (n, c, oh, ow) or out.dtype != io.out_dtype or out.device.type != "cpu":
        raise ValueError(f"out must be a {io.out_dtype} CPU tensor of shape {(n, c, oh, ow)}")
    bks = torch.empty((n,) + _bucket_shape(hk.variant, h, w, oh, ow), dtype=torch.int32) if want_buckets else None
    if bks is not None and hk.variant.family == "nnedi3":
        bks = None
    streams = _host_streams.get(dev.index)
    if streams is None:
        streams = _host_streams[dev.index] = [torch.cuda.Stream(dev) for _ in range(2)]
    frame_bytes = c * (x.element_size() * h * w + out.element_size() * oh * ow)
    bands = _host_row_bands(hk, n, c, h, oh, frame_bytes) if bks is None else None
    if bands is not None:
        # a few large luma frames: the pipeline's unit is a ROW BAND (plus the halo the kernels read, cropped after the
        # launch -- the arithmetic of sharding.prescale_rowsplit, bit-identical to the whole frame), so that the copies
        # of one band overlap the kernel of the next even when the call holds a single frame
        sy = oh // h
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            for s in streams:
                s.wait_stream(cur)
            k = 0
            for f in range(n):
                for a, b, s0, s1 in bands:
                    s = streams[k & 1]
                    k += 1
                    bpl = dataclasses.replace(pl, in_size=(s1 - s0, w), out_size=((s1 - s0) * sy, ow))
                    with torch.cuda.stream(s):
                        xd = x[f:f + 1, :, s0:s1, :].to(dev, non_blocking=True)          # contiguous rows of one plane
                        od, _ = _launch(hk, bpl, xd, W, False, io)
                        out[f:f + 1, :, a * sy:b * sy, :].copy_(od[:, :, (a - s0) * sy:(b - s0) * sy, :], non_blocking=True)
                        del xd, od
            for s in streams:
                s.synchronize()
        return out, None
    # chunks of 32..256 MB, about ten per call: the first upload and the last download are the only copies nothing hides
    chunk_bytes = min(256 << 20, max(32 << 20, n * frame_bytes // 10))
    chunk = max(1, min(n, chunk_bytes // max(frame_bytes, 1)))
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)
        # the first chunks are small (an eighth, a quarter, a half of the steady size), so the download engine starts early
        bounds, f0 = [], 0
        while f0 < n:
            step = max(1, chunk >> max(0, 3 - len(bounds)))
            bounds.append((f0, min(n, f0 + step)))
            f0 += step
        for k, (f0, f1) in enumerate(bounds):
            s = streams[k & 1]
            with torch.cuda.stream(s):
                xd = x[f0:f1].to(dev, non_blocking=True)
                od, bd = _launch(hk, pl, xd, W, want_buckets and bks is not None, io)
                out[f0:f1].copy_(od, non_blocking=True)
                if bd is not None:
                    bks[f0:f1].copy_(bd, non_blocking=True)
                del xd, od, bd
        for s in streams:
            s.synchronize()
    return out, bks


_HOST_BAND_MIN_BYTES = 8 << 20   # a band moves at least this much over PCIe (in + out); tests lower it
_HOST_BAND_CHUNKS = 16           # pipeline steps a call is cut into when it holds fewer frames than that


def _host_row_bands(hk: HookFile, n: int, c: int, h: int, oh: int, frame_bytes: int):
    """Row bands ``(a, b, s0, s1)`` (output rows of source rows [a, b) come from a launch on source rows [s0, s1)) for the
    host pipeline, or None when whole frames already give it enough chunks.  Luma planes only (a band of one plane is one
    contiguous piece of host memory); ravu-zoom is excluded like in the multi-GPU row split."""
    from .sharding import row_bands

    if c != 1 or hk.variant.family == "ravu-zoom" or oh % h != 0 or n >= _HOST_BAND_CHUNKS // 2 or h < 512:
        return None
    parts = min(-(-_HOST_BAND_CHUNKS // n), h // 256, max(1, frame_bytes // _HOST_BAND_MIN_BYTES))
    if parts < 2:
        return None
    return [bd for bd in row_bands(h, parts) if bd[1] > bd[0]]


def _bucket_shape(v: Variant, h: int, w: int, oh: int, ow: int) -> Tuple[int, ...]:
    if v.family == "ravu":
        return (3, h, w)
    if v.family == "ravu-zoom":
        return (oh, ow)
    return (h, w)
