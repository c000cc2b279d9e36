"""Generic hook runner: run an mpv user-shader ``.hook`` file that is NOT one of the shipped
prescalers' known forms -- e.g. a hand-modified RAVU, or any fragment-shader hook within the GLSL subset the shipped files
use -- by transpiling each pass body to a CUDA kernel at load time (NVRTC, sm_100a) and executing the pass chain with the
host semantics the reference relies on (``//!HOOK / BIND / SAVE / WIDTH / HEIGHT / WHEN / OFFSET / COMPONENTS``, textures
with NEAREST / LINEAR filtering, clamp-to-edge).

This is the slow, general path: one thread per output texel, one launch per pass, intermediates in HBM -- exactly the
structure of the reference's root variants.  The fused kernels behind :func:`prescale` remain the product path for the
shipped files; this runner exists so that a hook which those kernels refuse (``HookError``: "differs from the supported
form") still runs on the GPU, and it doubles as an on-device cross-check of the fused kernels.

Supported: fragment passes (``vec4 hook()``), helper functions, ``#define`` macros, the types / built-ins listed in
``glsl_prelude.cuh``, ``NAME_tex / NAME_texOff / NAME_pos / NAME_size / NAME_pt / NAME_raw / NAME_mul``,
``texture(lut, vec2)``, ``textureGatherOffset``, and compute passes (``//!COMPUTE bw bh [tw th]``: ``shared`` arrays,
``barrier()``, the ``gl_*InvocationID`` / ``gl_WorkGroup*`` built-ins, ``imageStore(out_image, ...)``, ``NAME_map``): one
CUDA block per work group, the shader's own shared-memory staging and thread mapping kept as written.
There is no CPU fallback; nothing here imports ``oracle/``.
"""
from __future__ import annotations

import os
import re
import threading
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .hookfile import HookError, HookFile, Pass

__all__ = ["GenericHook", "transpile_pass"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
_SWIZZLE = re.compile(r"\.([xyzw]{2,4})\b(?!\s*\()")
_FUNC_DEF = re.compile(r"^\s*(float|int|vec2|vec3|vec4|mat4x3|void)\s+([A-Za-z_]\w*)\s*\(([^)]*)\)\s*\{\s*$")


def _convert_functions(lines: List[str]) -> List[str]:
    """GLSL function definitions -> C++ lambdas (so that they can live inside the kernel and see its context)."""
    out: List[str] = []
    depth = 0
    open_at: List[int] = []  # brace depth at which a converted function body was opened
    for ln in lines:
        if ln.lstrip().startswith("#"):
            out.append(ln)
            continue
        m = _FUNC_DEF.match(ln) if depth == 0 else None
        if m:
            rtype, name, params = m.groups()
            out.append(f"auto {name} = [&]({params.strip()}) -> {rtype} {{")
            open_at.append(depth)
            depth += 1
            continue
        opens, closes = ln.count("{"), ln.count("}")
        depth += opens - closes
        if open_at and depth == open_at[-1] and closes > opens:
            open_at.pop()
            idx = ln.rfind("}")
            ln = ln[:idx] + "};" + ln[idx + 1:]
        out.append(ln)
    if depth != 0:
        raise HookError("unbalanced braces in shader body")
    return out


_SHARED_DECL = re.compile(r"^\s*shared\s+(float|vec2|vec3|vec4)\s+([A-Za-z_]\w*)\s*\[\s*(\d+)\s*\]\s*;\s*$")
_SHARED_WORDS = {"float": 1, "vec2": 2, "vec3": 3, "vec4": 4}


def _convert_shared(lines: List[str]) -> List[str]:
    """``shared T name[n];`` -> raw ``__shared__`` storage plus a typed pointer (the vector types have constructors, which
    CUDA does not allow on ``__shared__`` objects)."""
    out = []
    for ln in lines:
        m = _SHARED_DECL.match(ln)
        if m:
            t, name, n = m.group(1), m.group(2), int(m.group(3))
            out.append(f"__shared__ float _sh_{name}[{n * _SHARED_WORDS[t]}]; {t}* const {name} = reinterpret_cast<{t}*>(_sh_{name});")
        elif re.match(r"^\s*shared\b", ln):
            raise HookError(f"unsupported shared declaration: {ln.strip()!r}")
        else:
            out.append(ln)
    return out


def transpile_pass(p: Pass, index: int, bound: List[str]) -> str:
    """CUDA source of one pass: ``extern "C" __global__ void pass_<index>(...)``.  Fragment passes (``vec4 hook()``) run one
    thread per output texel; compute passes (``void hook()`` + ``imageStore``) run one block per work group."""
    body = p.body
    entry = "void hook()" if p.compute else "vec4 hook()"
    if entry not in body:
        raise HookError(f"pass {p.desc!r}: no `{entry}` entry point")
    body = _FLOAT_LIT.sub(lambda m: m.group(1) + "f", body)
    body = _SWIZZLE.sub(lambda m: "." + m.group(1) + "()", body)
    lines = _convert_functions(body.split("\n"))
    if p.compute:
        lines = _convert_shared(lines)
    macros = []
    for k, name in enumerate(bound):
        macros += [
            f"#define {name}_raw (_tex[{k}])",
            f"#define {name}_pos _pos",
            f"#define {name}_size vec2((float)_tex[{k}].w, (float)_tex[{k}].h)",
            f"#define {name}_pt vec2(1.0f / (float)_tex[{k}].w, 1.0f / (float)_tex[{k}].h)",
            f"#define {name}_mul 1.0f",
            f"#define {name}_tex(p) tex_sample(_tex[{k}], _frame, (p))",
            f"#define {name}_texOff(o) tex_sample(_tex[{k}], _frame, _pos + {name}_pt * (o))",
            f"#define {name}_map(id) _texmap(id)",
            f"#define {name} (_tex[{k}])",
        ]
    undef = [f"#undef {name}{sfx}" for name in bound for sfx in ("_raw", "_pos", "_size", "_pt", "_mul", "_tex", "_texOff", "_map", "")]
    # user macros must not leak into the next pass
    user = re.findall(r"^\s*#define\s+([A-Za-z_]\w*)", p.body, flags=re.M)
    undef += [f"#undef {u}" for u in user]
    head = [
        f"// pass {index}: {p.desc}",
        *macros,
        "#define texture(t, c) tex_sample((t), _frame, (c))",
        "#define textureGatherOffset(t, c, o, comp) tex_gather((t), _frame, (c), (o), (comp))",
        f'extern "C" __global__ void pass_{index}(const Tex* __restrict__ _tex, float* __restrict__ _out, int _ow, int _oh, int _oc, int _n) {{',
        "  const int _ox = blockIdx.x * blockDim.x + threadIdx.x, _oy = blockIdx.y * blockDim.y + threadIdx.y, _frame = blockIdx.z;",
    ]
    if p.compute:
        # every thread of the work group stays alive (barrier()); out-of-range imageStore()s are dropped like the host does.
        # NAME_map(id): output texel -> normalised coordinate, (id + 0.5) * (1 / output size) as the host's prelude defines it
        kernel = [
            "  const uvec3 gl_WorkGroupID(blockIdx.x, blockIdx.y, 0u), gl_WorkGroupSize(blockDim.x, blockDim.y, 1u);",
            "  const uvec3 gl_LocalInvocationID(threadIdx.x, threadIdx.y, 0u), gl_GlobalInvocationID((uint)_ox, (uint)_oy, 0u);",
            "  const uint gl_LocalInvocationIndex = threadIdx.y * blockDim.x + threadIdx.x;",
            "  const vec2 _pos = vec2(((float)_ox + 0.5f) / (float)_ow, ((float)_oy + 0.5f) / (float)_oh);",
            "  const vec2 _out_scale = vec2(1.0f / (float)_ow, 1.0f / (float)_oh);",
            "  auto _texmap = [&](ivec2 id) -> vec2 { return (vec2(id) + vec2(0.5f)) * _out_scale; };",
            "  const int out_image = 0;",
            "  auto imageStore = [&](int, ivec2 _p, vec4 _r) {",
            "    if (_p.x < 0 || _p.y < 0 || _p.x >= _ow || _p.y >= _oh) return;",
            "    float* _q = _out + ((i64)_frame * _oh + _p.y) * (i64)_ow * _oc + (i64)_p.x * _oc;",
            "    _q[0] = _r.x; if (_oc > 1) _q[1] = _r.y; if (_oc > 2) _q[2] = _r.z; if (_oc > 3) _q[3] = _r.w;",
            "  };",
            *lines,
            "  hook();",
            "}",
        ]
    else:
        kernel = [
            "  if (_ox >= _ow || _oy >= _oh) return;",
            "  const vec2 _pos = vec2(((float)_ox + 0.5f) / (float)_ow, ((float)_oy + 0.5f) / (float)_oh);",
            *lines,
            "  const vec4 _r = hook();",
            "  float* _q = _out + ((i64)_frame * _oh + _oy) * (i64)_ow * _oc + (i64)_ox * _oc;",
            "  _q[0] = _r.x; if (_oc > 1) _q[1] = _r.y; if (_oc > 2) _q[2] = _r.z; if (_oc > 3) _q[3] = _r.w;",
            "}",
        ]
    src = head + kernel + ["#undef texture", "#undef textureGatherOffset", *undef]
    return "\n".join(src) + "\n"


# ---- NVRTC / driver plumbing ------------------------------------------------------------------------------------------------

_nv_lock = threading.Lock()


def _check(res, what):
    err = res[0]
    if int(err) != 0:
        raise HookError(f"{what} failed: {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


def compile_cuda(source: str, name: str = "hook.cu") -> bytes:
    """NVRTC: CUDA source -> cubin for sm_100a (works without a GPU)."""
    from cuda.bindings import nvrtc

    with _nv_lock:
        prog = _check(nvrtc.nvrtcCreateProgram(source.encode(), name.encode(), 0, [], []), "nvrtcCreateProgram")
        opts = [b"--gpu-architecture=sm_100a", b"--fmad=false", b"-std=c++17", b"-default-device", b"--prec-div=true", b"--prec-sqrt=true"]
        res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if int(res[0]) != 0:
            size = _check(nvrtc.nvrtcGetProgramLogSize(prog), "nvrtcGetProgramLogSize")
            log = b" " * size
            nvrtc.nvrtcGetProgramLog(prog, log)
            nvrtc.nvrtcDestroyProgram(prog)
            raise HookError("the hook's GLSL is outside the supported subset (NVRTC):\n" + log.decode(errors="replace")[-3000:])
        size = _check(nvrtc.nvrtcGetCUBINSize(prog), "nvrtcGetCUBINSize")
        cubin = b" " * size
        _check(nvrtc.nvrtcGetCUBIN(prog, cubin), "nvrtcGetCUBIN")
        nvrtc.nvrtcDestroyProgram(prog)
        return cubin


class GenericHook:
    """A hook file compiled pass by pass; ``run()`` executes the chain on a batch of frames."""

    def __init__(self, hook: HookFile, lut_precision: str = "fp16"):
        self.hook = hook
        self.lut_precision = lut_precision
        with open(os.path.join(_HERE, "glsl_prelude.cuh")) as f:
            prelude = f.read()
        self.binds: List[List[str]] = []
        parts = [prelude]
        for k, p in enumerate(hook.passes):
            bound = list(dict.fromkeys(["HOOKED"] + [b for b in p.binds if b != "HOOKED"]))
            self.binds.append(bound)
            parts.append(transpile_pass(p, k, bound))
        self.source = "\n".join(parts)
        self.cubin = compile_cuda(self.source, os.path.basename(hook.path) + ".cu")
        self._modules: Dict[int, Tuple[object, List[object]]] = {}
        self._luts: Dict[Tuple[int, str], torch.Tensor] = {}

    # -- device state -----------------------------------------------------------------------------------------------
    def _functions(self, dev: int):
        from cuda.bindings import driver

        hit = self._modules.get(dev)
        if hit is None:
            mod = _check(driver.cuModuleLoadData(self.cubin), "cuModuleLoadData")
            fns = [_check(driver.cuModuleGetFunction(mod, f"pass_{k}".encode()), "cuModuleGetFunction") for k in range(len(self.hook.passes))]
            hit = self._modules[dev] = (mod, fns)
        return hit[1]

    def _lut(self, name: str, dev: torch.device) -> torch.Tensor:
        key = (dev.index, name)
        t = self._luts.get(key)
        if t is None:
            data = np.ascontiguousarray(self.hook.textures[name].data, dtype=np.float32)
            if self.lut_precision == "fp16":  # rgba16f storage
                data = data.astype(np.float16).astype(np.float32)
            t = self._luts[key] = torch.from_numpy(data).to(dev)
        return t

    # -- execution ----------------------------------------------------------------------------------------------------
    def run(self, frames: torch.Tensor, output_size: Optional[Tuple[int, int]] = None, is_yuv: bool = True):
        """frames: CUDA float32 ``[N,H,W]`` / ``[N,1,H,W]`` (LUMA hooks) or planar ``[N,3,H,W]`` (NATIVE / MAIN hooks).

        Returns ``(result, offset)``: the hooked plane after the chain (same layout as the input) and the accumulated
        ``//!OFFSET``.  ``output_size=(h, w)`` is mpv's OUTPUT for the ``WHEN`` / ``WIDTH`` / ``HEIGHT`` expressions
        (default: twice the input, what the shipped 2x prescalers assume)."""
        from cuda.bindings import driver

        if not torch.cuda.is_available() or frames.device.type != "cuda":
            raise HookError("the generic hook runner needs CUDA tensors: there is no CPU fallback")
        if frames.dtype != torch.float32:
            raise TypeError("the generic hook runner takes float32 planes")
        x = frames
        if x.dim() == 3:
            x = x[:, None]
        n, c, h, w = x.shape
        dev = x.device
        oh, ow = (2 * h, 2 * w) if output_size is None else (int(output_size[0]), int(output_size[1]))
        # textures are interleaved [n][h][w][comps]
        hooked = x.permute(0, 2, 3, 1).contiguous() if c > 1 else x.reshape(n, h, w, 1).contiguous()
        if c == 3:  # alpha = 1 like the host's RGBA textures
            hooked = torch.cat([hooked, torch.ones((n, h, w, 1), device=dev)], dim=3).contiguous()
        saved: Dict[str, torch.Tensor] = {}
        env = {"HOOKED": (w, h), "OUTPUT": (ow, oh), "LUMA": (w if is_yuv else 0, h if is_yuv else 0), "NATIVE": (w, h), "MAIN": (w, h)}
        off = [0.0, 0.0]
        fns = self._functions(dev.index)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for k, p in enumerate(self.hook.passes):
                e = dict(env)
                e.update({name: (t.shape[2], t.shape[1]) for name, t in saved.items()})
                if not p.enabled(e):
                    continue
                pw, ph = p.output_size(e)
                comps = int(p.components) if p.components else hooked.shape[3]
                texs = []
                for name in self.binds[k]:
                    if name == "HOOKED":
                        t, linear, per_frame = hooked, 0, True
                    elif name in saved:
                        t, linear, per_frame = saved[name], 0, True
                    elif name in self.hook.textures:
                        t, linear, per_frame = self._lut(name, dev), int(self.hook.textures[name].filter.upper() == "LINEAR"), False
                    else:
                        raise HookError(f"pass {p.desc!r} binds {name}, which no earlier pass saved")
                    th, tw, tc = (t.shape[1], t.shape[2], t.shape[3]) if per_frame else (t.shape[0], t.shape[1], t.shape[2])
                    texs.append((t, tw, th, tc, linear, (th * tw * tc) if per_frame else 0))
                # struct Tex { const float* p; int w, h, comps, linear; i64 stride_n; }  (32 bytes)
                desc = np.zeros((len(texs), 4), dtype=np.int64)
                for i, (t, tw, th, tc, linear, sn) in enumerate(texs):
                    desc[i, 0] = t.data_ptr()
                    desc[i, 1] = (th << 32) | tw
                    desc[i, 2] = (linear << 32) | tc
                    desc[i, 3] = sn
                d_desc = torch.from_numpy(desc).to(dev)
                if p.compute:   # //!COMPUTE bw bh [tw th]: one work group per bw x bh output block, tw x th threads
                    bw, bh = p.compute[0], p.compute[1]
                    tw, th = (p.compute[2], p.compute[3]) if len(p.compute) >= 4 else (bw, bh)
                    grid, block = ((pw + bw - 1) // bw, (ph + bh - 1) // bh, n), (tw, th, 1)
                    out = torch.zeros((n, ph, pw, comps), dtype=torch.float32, device=dev)
                else:
                    grid, block = ((pw + 31) // 32, (ph + 7) // 8, n), (32, 8, 1)
                    out = torch.empty((n, ph, pw, comps), dtype=torch.float32, device=dev)
                args = [np.array([d_desc.data_ptr()], dtype=np.uint64), np.array([out.data_ptr()], dtype=np.uint64),
                        np.array([pw], dtype=np.int32), np.array([ph], dtype=np.int32), np.array([comps], dtype=np.int32),
                        np.array([n], dtype=np.int32)]
                argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
                _check(driver.cuLaunchKernel(fns[k], *grid, *block, 0, stream, argp.ctypes.data, 0), f"launch of pass {k}")
                torch.cuda.current_stream(dev).synchronize()  # keeps d_desc / args alive; this is the slow general path
                if p.save:
                    saved[p.save] = out
                else:
                    hooked = out
                    env["HOOKED"] = (pw, ph)
                if isinstance(p.offset, tuple):
                    off[0] += p.offset[0]
                    off[1] += p.offset[1]
        res = hooked[..., : max(1, min(c, hooked.shape[3]))].permute(0, 3, 1, 2).contiguous()
        if frames.dim() == 3:
            res = res[:, 0]
        return res, (off[0], off[1])
