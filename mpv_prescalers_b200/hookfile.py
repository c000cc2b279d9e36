"""Parser for mpv user-shader ``.hook`` files as shipped by bjin/mpv-prescalers.

This is the drop-in boundary of the project: the reference's "API" is the hook
file itself (``README.md:33-37`` -- ``glsl-shader="~~/shaders/ravu-lite-ar-r3.hook"``).
We parse the shipped files unchanged:

* block structure and ``//!`` directives (HOOK BIND SAVE DESC WIDTH HEIGHT OFFSET
  WHEN COMPONENTS COMPUTE TEXTURE SIZE FORMAT FILTER), e.g.
  ``ravu-lite-ar-r3.hook:15-21,185-193,198-201``;
* the RPN expressions of WIDTH/HEIGHT/WHEN (``ravu-lite-ar-r3.hook:20``);
* the ``//!TEXTURE`` hex payloads (4 x float32 per texel, ``ravu-lite-ar-r3.hook:202``);
* the NNEDI3 inline weights ``W(i,w0..w3)`` / ``WS(b1,b2)``
  (``nnedi3-nns16-win8x4.hook:31-49``) for the root, gather and compute orderings;
* the numeric constants the kernels take as parameters (Gaussian weights of the
  structure tensor, strength / coherence thresholds, anti-ringing tap set and
  strength), each validated against the structural rule the CUDA kernels
  implement so that a hand-edited hook is refused instead of silently mis-run.

Nothing here touches the GPU.
"""
from __future__ import annotations

import hashlib
import os
import re
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np


class HookError(ValueError):
    """Raised for files that are not (supported) mpv-prescalers hook files."""


# ----------------------------------------------------------------------------------------------
# RPN expressions  (mpv user-shader semantics: postfix, operands are numbers or TEX.w / TEX.h,
# operators + - * / < > = !, comparisons give 1/0, '*' doubles as logical AND)
# ----------------------------------------------------------------------------------------------

RpnEnv = Dict[str, Tuple[float, float]]


def eval_rpn(tokens: Sequence[str], env: RpnEnv) -> float:
    """Evaluate an mpv RPN size/condition expression.

    ``env`` maps texture names (``HOOKED``, ``OUTPUT``, ``LUMA`` ...) to ``(w, h)``.
    Follows the semantics of the WHEN/WIDTH/HEIGHT lines used by the reference, e.g.
    ``HOOKED.w OUTPUT.w / 0.707106 < HOOKED.h OUTPUT.h / 0.707106 < *``
    (``ravu-lite-ar-r3.hook:20``).
    """
    stack: List[float] = []
    for tok in tokens:
        if tok in ("+", "-", "*", "/", "<", ">", "="):
            if len(stack) < 2:
                raise HookError(f"RPN stack underflow at {tok!r} in {' '.join(tokens)!r}")
            b = stack.pop()
            a = stack.pop()
            if tok == "+":
                stack.append(a + b)
            elif tok == "-":
                stack.append(a - b)
            elif tok == "*":
                stack.append(a * b)
            elif tok == "/":
                stack.append(a / b)
            elif tok == "<":
                stack.append(1.0 if a < b else 0.0)
            elif tok == ">":
                stack.append(1.0 if a > b else 0.0)
            else:
                stack.append(1.0 if a == b else 0.0)
        elif tok == "!":
            if not stack:
                raise HookError("RPN stack underflow at '!'")
            stack.append(0.0 if stack.pop() else 1.0)
        elif re.fullmatch(r"[A-Za-z_][A-Za-z0-9_]*\.(w|h|width|height)", tok):
            name, comp = tok.split(".")
            if name not in env:
                raise HookError(f"RPN references unknown texture {name!r}")
            stack.append(float(env[name][0 if comp[0] == "w" else 1]))
        else:
            try:
                stack.append(float(tok))
            except ValueError:
                raise HookError(f"bad RPN token {tok!r}") from None
    if len(stack) != 1:
        raise HookError(f"RPN expression {' '.join(tokens)!r} leaves {len(stack)} values")
    return stack[0]


# ----------------------------------------------------------------------------------------------
# Blocks
# ----------------------------------------------------------------------------------------------


@dataclass
class Texture:
    """A ``//!TEXTURE`` block (``ravu-lite-ar-r3.hook:198-202``)."""

    name: str
    width: int
    height: int
    format: str
    filter: str
    data: np.ndarray  # float32 [height, width, 4], as written in the file (NOT yet fp16-rounded)
    line: int = 0

    @property
    def sha16(self) -> str:
        return hashlib.sha256(self.data.tobytes()).hexdigest()[:16]


@dataclass
class Pass:
    """One shader pass: a run of ``//!`` directives followed by GLSL."""

    desc: str = ""
    hook: List[str] = field(default_factory=list)
    binds: List[str] = field(default_factory=list)
    save: Optional[str] = None
    width: Optional[List[str]] = None
    height: Optional[List[str]] = None
    when: Optional[List[str]] = None
    offset: Union[None, str, Tuple[float, float]] = None
    components: Optional[int] = None
    compute: Optional[Tuple[int, ...]] = None
    body: str = ""
    line: int = 0  # 1-based line of the first directive

    def output_size(self, env: RpnEnv) -> Tuple[int, int]:
        w, h = env["HOOKED"]
        if self.width is not None:
            w = eval_rpn(self.width, env)
        if self.height is not None:
            h = eval_rpn(self.height, env)
        return int(w), int(h)

    def enabled(self, env: RpnEnv) -> bool:
        return True if self.when is None else eval_rpn(self.when, env) != 0.0


_KNOWN_FORMATS = {"rgba16f": 16, "rgba32f": 16}  # bytes of HOST data per texel (4 x float32)


def _decode_texture(header: Dict[str, List[str]], payload: str, line: int) -> Texture:
    name = header["TEXTURE"][0]
    size = [int(v) for v in header.get("SIZE", [""])[0].split()]
    if len(size) != 2:
        raise HookError(f"texture {name}: only 2-D textures are supported (SIZE {size})")
    fmt = header.get("FORMAT", ["?"])[0].strip()
    if fmt not in _KNOWN_FORMATS:
        # mirrors the host's "Unrecognized/unavailable FORMAT name" failure (README.md:19-21)
        raise HookError(f"texture {name}: unrecognized/unavailable FORMAT name: {fmt!r}")
    filt = header.get("FILTER", ["NEAREST"])[0].strip().upper()
    if filt not in ("NEAREST", "LINEAR"):
        raise HookError(f"texture {name}: bad FILTER {filt!r}")
    w, h = size
    payload = payload.strip()
    want = w * h * _KNOWN_FORMATS[fmt] * 2
    if len(payload) != want:
        raise HookError(f"texture {name}: payload has {len(payload)} hex chars, expected {want}")
    try:
        raw = bytes.fromhex(payload)
    except ValueError as e:
        raise HookError(f"texture {name}: payload is not hex ({e})") from None
    data = np.frombuffer(raw, dtype="<f4").reshape(h, w, 4).astype(np.float32)
    return Texture(name, w, h, fmt, filt, data, line)


_cache_lock = threading.Lock()
_cache: Dict[Tuple[str, float, int], "HookFile"] = {}


class HookFile:
    """A parsed hook file: ordered passes + named textures."""

    def __init__(self, path: str, passes: List[Pass], textures: Dict[str, Texture]):
        self.path = path
        self.name = os.path.basename(path)
        self.passes = passes
        self.textures = textures
        self._variant = None
        self._content_key = None

    # -- parsing ---------------------------------------------------------------------------
    @classmethod
    def parse(cls, path: Union[str, os.PathLike]) -> "HookFile":
        path = os.fspath(path)
        try:
            st = os.stat(path)
        except OSError as e:
            raise HookError(f"cannot read hook file {path!r}: {e}") from None
        key = (os.path.abspath(path), st.st_mtime, st.st_size)
        with _cache_lock:
            hit = _cache.get(key)
        if hit is not None:
            return hit
        with open(path, "r", encoding="utf-8", errors="replace") as f:
            text = f.read()
        hook = cls.parse_text(text, path)
        with _cache_lock:
            _cache[key] = hook
        return hook

    @classmethod
    def parse_text(cls, text: str, path: str = "<string>") -> "HookFile":
        lines = text.split("\n")
        passes: List[Pass] = []
        textures: Dict[str, Texture] = {}
        i, n = 0, len(lines)
        # skip everything before the first directive (licence header, lines 1-14)
        while i < n and not lines[i].startswith("//!"):
            i += 1
        while i < n:
            start = i
            header: Dict[str, List[str]] = {}
            while i < n and lines[i].startswith("//!"):
                parts = lines[i][3:].strip().split(None, 1)
                if parts:
                    header.setdefault(parts[0].upper(), []).append(parts[1] if len(parts) > 1 else "")
                i += 1
            body_start = i
            while i < n and not lines[i].startswith("//!"):
                i += 1
            body = "\n".join(lines[body_start:i])
            if "TEXTURE" in header:
                tex = _decode_texture(header, body, start + 1)
                if tex.name in textures:
                    raise HookError(f"{path}: duplicate texture {tex.name}")
                textures[tex.name] = tex
                continue
            p = Pass(body=body, line=start + 1)
            for key, vals in header.items():
                if key == "DESC":
                    p.desc = vals[-1]
                elif key == "HOOK":
                    p.hook = [v.strip() for v in vals]
                elif key == "BIND":
                    p.binds = [v.strip() for v in vals]
                elif key == "SAVE":
                    p.save = vals[-1].strip()
                elif key == "WIDTH":
                    p.width = vals[-1].split()
                elif key == "HEIGHT":
                    p.height = vals[-1].split()
                elif key == "WHEN":
                    p.when = vals[-1].split()
                elif key == "OFFSET":
                    v = vals[-1].split()
                    if len(v) == 1 and v[0].upper() == "ALIGN":
                        p.offset = "ALIGN"
                    elif len(v) == 2:
                        p.offset = (float(v[0]), float(v[1]))
                    else:
                        raise HookError(f"{path}:{start + 1}: bad OFFSET {vals[-1]!r}")
                elif key == "COMPONENTS":
                    p.components = int(vals[-1])
                elif key == "COMPUTE":
                    p.compute = tuple(int(t) for t in vals[-1].split())
                else:
                    raise HookError(f"{path}:{start + 1}: unsupported directive //!{key}")
            if not p.hook:
                raise HookError(f"{path}:{start + 1}: pass without //!HOOK")
            passes.append(p)
        if not passes:
            raise HookError(f"{path}: no shader passes found (not an mpv hook file?)")
        return cls(path, passes, textures)

    # -- classification --------------------------------------------------------------------
    @property
    def variant(self) -> "Variant":
        if self._variant is None:
            self._variant = classify(self)
        return self._variant

    @property
    def content_key(self) -> str:
        """Identity of everything the kernels consume from this file (LUT payloads or NNEDI3 weights, key constants,
        family parameters): two hooks with the same key are interchangeable on the device, whatever their path --
        which is what the device weight cache is keyed on (every ``parse_text()`` hook has the path '<string>')."""
        if self._content_key is None:
            v = self.variant
            h = hashlib.sha256()
            h.update(repr((v.family, v.plane, v.radius, v.ar, v.ar_strength, v.ar_taps, v.scale, v.strength_thr,
                           v.strength_log2_scale, v.coherence_thr, v.n_strength, v.nns, v.win)).encode())
            if v.gauss is not None:
                h.update(np.ascontiguousarray(v.gauss, dtype=np.float32).tobytes())
            for tex in (v.lut, v.lut_ar):
                if tex is not None:
                    h.update(repr((tex.name, tex.width, tex.height)).encode())
                    h.update(np.ascontiguousarray(tex.data, dtype=np.float32).tobytes())
            for nn in (v.nn_y, v.nn_x):
                if nn is not None:
                    for a in (nn.w1, nn.w2, nn.b1, nn.b2):
                        h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
            self._content_key = h.hexdigest()
        return self._content_key


# ----------------------------------------------------------------------------------------------
# Variant extraction
# ----------------------------------------------------------------------------------------------

_FLT = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"
EPS_LITERAL = "1.192092896e-7"


@dataclass
class Nnedi3Weights:
    """Weights of one NNEDI3 direction in canonical order.

    ``w1``/``w2`` are ``[nns, 8, S]``: index ``[n, a, b]`` multiplies the sample at offset
    ``a - 3`` along the LONG window axis and ``b - (S/2 - 1)`` along the SHORT axis
    (long axis = x for double_y, y for double_x; ``nnedi3-nns16-win8x4.hook:55-86,145-176``).
    """

    w1: np.ndarray
    w2: np.ndarray
    b1: np.ndarray
    b2: np.ndarray


@dataclass
class Variant:
    family: str  # 'ravu-lite' | 'ravu' | 'ravu-zoom' | 'ravu-3x' | 'nnedi3'
    flavour: str  # 'root' | 'gather' | 'compute'
    plane: str  # 'luma' | 'yuv' | 'rgb'
    radius: int = 0
    ar: bool = False
    ar_strength: float = 0.0
    ar_taps: Tuple[int, ...] = ()  # tap indices t (x-major) that take part in anti-ringing
    scale: Optional[int] = None  # 2, 3 or None (zoom: arbitrary)
    gauss: Optional[np.ndarray] = None  # float32 [g*g], x-major over the inner g x g points
    strength_thr: Tuple[float, ...] = ()  # lite/zoom/3x thresholds; () for ravu (log2 form)
    strength_log2_scale: float = 0.0  # ravu: clamp(floor(log2(lambda*scale + eps)), 0, 8)
    coherence_thr: Tuple[float, ...] = (0.25, 0.5)
    n_angle: int = 24
    n_strength: int = 0
    n_coherence: int = 3
    lut: Optional[Texture] = None
    lut_ar: Optional[Texture] = None
    offset: Tuple[float, float] = (0.0, 0.0)  # accumulated //!OFFSET of a full application
    hook_point: str = "LUMA"
    # nnedi3
    nns: int = 0
    win: Tuple[int, int] = (0, 0)
    nn_y: Optional[Nnedi3Weights] = None
    nn_x: Optional[Nnedi3Weights] = None

    @property
    def taps(self) -> int:
        if self.family in ("ravu-lite", "ravu-3x"):
            return (2 * self.radius - 1) ** 2
        if self.family in ("ravu", "ravu-zoom"):
            return (2 * self.radius) ** 2
        return self.win[0] * self.win[1]

    @property
    def channels(self) -> int:
        return 1 if self.plane == "luma" else 3


_DESC_PATTERNS = [
    ("ravu-lite", re.compile(r"RAVU-Lite(?P<ar>-AR)? \((?:step\d, )?r(?P<r>\d)(?P<c>, compute)?\)")),
    ("ravu", re.compile(r"RAVU \(step\d, (?P<plane>luma|yuv|rgb), r(?P<r>\d)(?P<c>, compute)?\)")),
    ("ravu-zoom", re.compile(r"RAVU-Zoom(?P<ar>-AR)? \((?P<plane>luma|yuv|rgb), r(?P<r>\d)(?P<c>, compute)?\)")),
    ("ravu-3x", re.compile(r"RAVU-3x \((?P<plane>luma|yuv|rgb), r(?P<r>\d)\)")),
    ("nnedi3", re.compile(r"NNEDI3 \(double_[xy], nns(?P<nns>\d+), win(?P<wl>\d)x(?P<ws>\d)\)")),
]


def gaussian_weights(g: int) -> np.ndarray:
    """sigma=2 Gaussian over the inner g x g gradient points, normalised (float64)."""
    c = (g - 1) / 2.0
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    w = np.exp(-((ii - c) ** 2 + (jj - c) ** 2) / 8.0)
    return (w / w.sum()).reshape(-1)


def _window(family: str, r: int) -> Tuple[int, int, int]:
    """(n, lo, g): window side, lowest tap offset, side of the inner gradient square."""
    if family in ("ravu-lite", "ravu-3x"):
        return 2 * r - 1, -(r - 1), {2: 3, 3: 3, 4: 5}[r]
    return 2 * r, -(r - 1), {2: 4, 3: 4, 4: 6}[r]


def gradient_points(family: str, r: int) -> List[Tuple[int, int]]:
    """Window indices (i, j) of the inner g x g gradient points, x-major (file order)."""
    n, _, g = _window(family, r)
    o = (n - g) // 2
    return [(i, j) for i in range(o, o + g) for j in range(o, o + g)]


def stencil_kind(family: str, n: int, i: int) -> str:
    """Which finite-difference form the reference uses along an axis at window index ``i``.

    lite/3x (``ravu-lite-r2.hook:34-60``): central if both neighbours are inside the window else
    one-sided.  ravu/zoom (``ravu-r3.hook:60-106``): 4th order where +-2 are inside, else central,
    else one-sided.
    """
    if family in ("ravu", "ravu-zoom") and i - 2 >= 0 and i + 2 <= n - 1:
        return "o4"
    if i - 1 >= 0 and i + 1 <= n - 1:
        return "central"
    return "fwd" if i - 1 < 0 else "bwd"


def _expected_gradient_lines(family: str, r: int, var: str) -> List[Tuple[str, str]]:
    n, _, _ = _window(family, r)

    def expr(kind: str, idx, i: int, j: int, axis: int) -> str:
        def s(d: int) -> str:
            ii, jj = (i + d, j) if axis == 0 else (i, j + d)
            return f"{var}{idx(ii, jj)}"

        if kind == "o4":
            return f"(-{s(2)}+8.0*{s(1)}-8.0*{s(-1)}+{s(-2)})/12.0"
        if kind == "central":
            return f"({s(1)}-{s(-1)})/2.0"
        if kind == "fwd":
            return f"({s(1)}-{s(0)})"
        return f"({s(0)}-{s(-1)})"

    idx = lambda ii, jj: ii * n + jj
    out = []
    for (i, j) in gradient_points(family, r):
        out.append((expr(stencil_kind(family, n, i), idx, i, j, 0), expr(stencil_kind(family, n, j), idx, i, j, 1)))
    return out


def _extract_key_constants(v: Variant, body: str, where: str, deep: bool) -> None:
    """Pull the structure-tensor / bucket constants out of a RAVU pass body and validate them."""
    n, _, g = _window(v.family, v.radius)
    gauss = [float(m) for m in re.findall(r"abd \+= vec3\(gx \* gx, gx \* gy, gy \* gy\) \* (" + _FLT + r");", body)]
    if len(gauss) != g * g:
        raise HookError(f"{where}: expected {g * g} structure-tensor terms, found {len(gauss)}")
    ref = gaussian_weights(g)
    if not np.allclose(gauss, ref, rtol=0, atol=1e-12):
        raise HookError(f"{where}: structure-tensor weights are not the sigma=2 Gaussian the kernels implement")
    v.gauss = np.asarray(gauss, dtype=np.float64).astype(np.float32)

    if deep:
        got = re.findall(r"^gx = (.*);\ngy = (.*);$", body, flags=re.M)
        # key samples are called lumaN (lite, 3x, rgb), sampleN (luma ravu/zoom) or sampleN.x (yuv)
        got = [tuple(re.sub(r"(?:luma|sample)(\d+)(?:\.x)?", r"s\1", e) for e in pair) for pair in got]
        if got != _expected_gradient_lines(v.family, v.radius, "s"):
            raise HookError(f"{where}: gradient stencils differ from the rule the kernels implement")

    must = [
        "float T = a + d, D = a * d - b * b;",
        "float delta = sqrt(max(T * T / 4.0 - D, 0.0));",
        "float L1 = T / 2.0 + delta, L2 = T / 2.0 - delta;",
        "float sqrtL1 = sqrt(L1), sqrtL2 = sqrt(L2);",
        "float theta = mix(mod(atan(L1 - a, b) + 3.141592653589793, 3.141592653589793), 0.0, abs(b) < 1.192092896e-7);",
        "float mu = mix((sqrtL1 - sqrtL2) / (sqrtL1 + sqrtL2), 0.0, sqrtL1 + sqrtL2 < 1.192092896e-7);",
        "float angle = floor(theta * 24.0 / 3.141592653589793);",
    ]
    for line in must:
        if line not in body:
            raise HookError(f"{where}: key section differs from the supported form (missing {line!r})")

    m = re.search(r"float strength = (.*);", body)
    if not m:
        raise HookError(f"{where}: no strength line")
    s = m.group(1)
    m3 = re.fullmatch(r"mix\(mix\(0\.0, 1\.0, lambda >= (" + _FLT + r")\), mix\(2\.0, 3\.0, lambda >= (" + _FLT + r")\), lambda >= (" + _FLT + r")\)", s)
    m2 = re.fullmatch(r"mix\(mix\(0\.0, 1\.0, lambda >= (" + _FLT + r")\), 2\.0, lambda >= (" + _FLT + r")\)", s)
    ml = re.fullmatch(r"clamp\(floor\(log2\(lambda \* (" + _FLT + r") \+ 1\.192092896e-7\)\), 0\.0, (" + _FLT + r")\)", s)
    if m3:
        t0, t2, t1 = (float(x) for x in m3.groups())
        v.strength_thr, v.n_strength = (t0, t1, t2), 4
    elif m2:
        t0, t1 = (float(x) for x in m2.groups())
        v.strength_thr, v.n_strength = (t0, t1), 3
    elif ml:
        v.strength_log2_scale = float(ml.group(1))
        v.n_strength = int(float(ml.group(2))) + 1
    else:
        raise HookError(f"{where}: unsupported strength quantiser {s!r}")
    if list(v.strength_thr) != sorted(v.strength_thr):
        raise HookError(f"{where}: strength thresholds are not increasing")

    m = re.search(r"float coherence = mix\(mix\(0\.0, 1\.0, mu >= (" + _FLT + r")\), 2\.0, mu >= (" + _FLT + r")\);", body)
    if not m:
        raise HookError(f"{where}: unsupported coherence quantiser")
    v.coherence_thr = (float(m.group(1)), float(m.group(2)))

    m = re.search(r"float coord_y = \(\(angle \* (" + _FLT + r") \+ strength\) \* (" + _FLT + r") \+ coherence( \+ 0\.5)?\) / (" + _FLT + r");", body)
    if not m:
        raise HookError(f"{where}: unsupported LUT row formula")
    ns, nc, rows = int(float(m.group(1))), int(float(m.group(2))), int(float(m.group(4)))
    if ns != v.n_strength or nc != v.n_coherence or rows != v.n_angle * ns * nc:
        raise HookError(f"{where}: LUT row formula ({ns},{nc},{rows}) inconsistent with the quantisers")
    if (m.group(3) is None) != (v.family == "ravu-zoom"):
        raise HookError(f"{where}: unexpected LUT row centring for family {v.family}")


def _ar_from_body(v: Variant, body: str, where: str) -> None:
    m = re.search(r"res = mix\(res, clamp\(res, lo, hi\), (" + _FLT + r")\);", body)
    if not m:
        raise HookError(f"{where}: anti-ringing variant without the mix(res, clamp(res, lo, hi), S) line")
    v.ar_strength = float(m.group(1))
    if not (0.0 <= v.ar_strength <= 1.0):
        raise HookError(f"{where}: anti-ringing strength {v.ar_strength} outside [0,1]")
    n, lo, _ = _window(v.family, v.radius)
    if v.family == "ravu-lite":
        # diamond dx^2+dy^2 <= 4 (r2: all 9 taps); validate against the cg4/cg2 lines of root files
        taps = tuple(t for t in range(n * n) if ((t // n) + lo) ** 2 + ((t % n) + lo) ** 2 <= 4)
        v.ar_taps = taps
        if v.flavour == "root":
            seen = set()
            for a, b in re.findall(r"cg4 = vec4\(0\.1 \+ luma(\d+), 1\.1 - luma\d+, 0\.1 \+ luma(\d+), 1\.1 - luma\d+\);", body):
                seen.update((int(a), int(b)))
            for a in re.findall(r"vec2 cg2 = vec2\(0\.1 \+ luma(\d+), 1\.1 - luma\d+\);", body):
                seen.add(int(a))
            if seen != set(taps):
                raise HookError(f"{where}: anti-ringing tap set differs from the dx^2+dy^2<=4 diamond")
            if body.count("cg4 *= cg4;cg4 *= cg4;cg4 *= cg4;cg4 *= cg4;cg4 *= cg4;") != (len(taps) - 1) // 2:
                raise HookError(f"{where}: anti-ringing power chain is not x^32 / x^33")
    else:
        v.ar_taps = tuple(range(n * n))


def _check_lut(v: Variant, tex: Texture, where: str) -> None:
    n = {"ravu-lite": (2 * v.radius - 1) ** 2, "ravu-3x": (2 * v.radius - 1) ** 2}.get(v.family, (2 * v.radius) ** 2)
    rows = v.n_angle * v.n_strength * v.n_coherence
    if v.family == "ravu-lite":
        want = ((n + 1) // 2, rows)
    elif v.family == "ravu-3x":
        want = (n + 1, rows)
    elif v.family == "ravu":
        want = ((n // 2 + 3) // 4, rows)
    else:  # zoom
        want = (((n // 2 + 3) // 4) * 9, rows * 9)
    if (tex.width, tex.height) != want:
        raise HookError(f"{where}: LUT {tex.name} is {tex.width}x{tex.height}, kernels expect {want[0]}x{want[1]}")
    want_filter = "LINEAR" if v.family == "ravu-zoom" else "NEAREST"
    if tex.filter != want_filter:
        raise HookError(f"{where}: LUT {tex.name} FILTER {tex.filter}, expected {want_filter}")


_W_RE = re.compile(r"W\((\d+),(-?\d+),(-?\d+),(-?\d+),(-?\d+)\)")
_WS_RE = re.compile(r"WS\((-?\d+),(-?\d+)\)")


def _bits_to_float(vals) -> np.ndarray:
    return np.asarray(vals, dtype=np.int64).astype(np.int32).view(np.float32)


def _nnedi3_sample_map(body: str, where: str) -> Dict[Tuple[int, int], Tuple[int, int]]:
    """(vec4 index, component) -> (dx, dy) for root / gather / compute bodies."""
    smap: Dict[Tuple[int, int], Tuple[int, int]] = {}
    for i, j, dx, dy in re.findall(r"samples\[(\d+)\]\[(\d)\] = HOOKED_texOff\(vec2\((" + _FLT + r"), (" + _FLT + r")\)\)\.x;", body):
        smap[(int(i), int(j))] = (int(float(dx)), int(float(dy)))
    if smap:
        return smap
    for i, ox, oy in re.findall(r"samples\[(\d+)\] = HOOKED_mul \* textureGatherOffset\(HOOKED_raw, HOOKED_pos, ivec2\((-?\d+), (-?\d+)\), 0\);", body):
        i, ox, oy = int(i), int(ox), int(oy)
        # textureGather component order: x=(0,1) y=(1,1) z=(1,0) w=(0,0)
        for comp, (ax, ay) in enumerate(((0, 1), (1, 1), (1, 0), (0, 0))):
            smap[(i, comp)] = (ox + ax, oy + ay)
    if smap:
        return smap
    m = re.search(r"int local_pos = int\(gl_LocalInvocationID\.x\) \* (\d+) \+ int\(gl_LocalInvocationID\.y\);", body)
    mb = re.search(r"group_base\.x\+x-\((\d+)\)\)\+0\.5,float\(group_base\.y\+y-\((\d+)\)\)\+0\.5", body)
    if m and mb:
        stride, bx, by = int(m.group(1)), int(mb.group(1)), int(mb.group(2))
        for i, j, k in re.findall(r"samples\[(\d+)\]\[(\d)\] = inp\[local_pos \+ (\d+)\];", body):
            k = int(k)
            smap[(int(i), int(j))] = (k // stride - bx, k % stride - by)
    if not smap:
        raise HookError(f"{where}: cannot find the NNEDI3 sample window")
    return smap


def _nnedi3_pass(p: Pass, nns: int, win: Tuple[int, int], direction: str, where: str) -> Nnedi3Weights:
    K = win[0] * win[1]
    S = win[1]
    nvec = K // 4
    body = p.body
    for line in (
        f"float mstd0 = sum / {K}.0;",
        f"float mstd1 = sumsq / {K}.0 - mstd0 * mstd0;",
        "float mstd2 = mix(0.0, inversesqrt(mstd1), mstd1 >= 1.192092896e-7);",
        "mstd1 *= mstd2;",
        "#define WS(w0,w1) sum1 = exp(sum1 * mstd2 + T(w0)); sum2 = sum2 * mstd2 + T(w1); wsum += sum1; vsum += sum1*(sum2/(1.0+abs(sum2)));",
        "return clamp(mstd0 + 5.0 * vsum / wsum * mstd1, 0.0, 1.0);",
    ):
        if line not in body:
            raise HookError(f"{where}: NNEDI3 predictor differs from the supported form (missing {line!r})")
    smap = _nnedi3_sample_map(body, where)
    if len(smap) != K:
        raise HookError(f"{where}: NNEDI3 window has {len(smap)} samples, expected {K}")
    # canonical index: a along the long axis (8), b along the short axis (S)
    canon = np.zeros((nvec, 4), dtype=np.int64)
    seen = set()
    for (i, j), (dx, dy) in smap.items():
        lng, sht = (dx, dy) if direction == "y" else (dy, dx)
        a, b = lng + 3, sht + (S // 2 - 1)
        if not (0 <= a < 8 and 0 <= b < S):
            raise HookError(f"{where}: NNEDI3 sample offset ({dx},{dy}) outside the {win[0]}x{win[1]} window")
        canon[i, j] = a * S + b
        seen.add(a * S + b)
    if len(seen) != K:
        raise HookError(f"{where}: NNEDI3 window does not cover {win[0]}x{win[1]}")
    neuron_lines = [ln for ln in body.split("\n") if ln.startswith("sum1=W(")]
    if len(neuron_lines) != nns:
        raise HookError(f"{where}: {len(neuron_lines)} neuron lines, expected {nns}")
    w1 = np.zeros((nns, K), np.float32)
    w2 = np.zeros((nns, K), np.float32)
    b1 = np.zeros(nns, np.float32)
    b2 = np.zeros(nns, np.float32)
    flat = canon.reshape(-1)
    for nidx, ln in enumerate(neuron_lines):
        try:
            part1, rest = ln.split(";sum2=", 1)
            part2, ws = rest.split(";WS(", 1)
        except ValueError:
            raise HookError(f"{where}: malformed neuron line {nidx}") from None
        for dst, part in ((w1, part1), (w2, part2)):
            toks = _W_RE.findall(part)
            if [int(t[0]) for t in toks] != list(range(nvec)):
                raise HookError(f"{where}: neuron {nidx}: W() terms out of order")
            vals = _bits_to_float([[int(x) for x in t[1:]] for t in toks]).reshape(-1)
            dst[nidx, flat] = vals
        m = _WS_RE.fullmatch("WS(" + ws.rstrip().rstrip(";"))
        if not m:
            raise HookError(f"{where}: neuron {nidx}: malformed WS()")
        bb = _bits_to_float([int(m.group(1)), int(m.group(2))])
        b1[nidx], b2[nidx] = bb[0], bb[1]
    # The tensor-core path feeds (x - mean) * inv_std into the contraction while the shader computes dot(x, W) * inv_std
    # (nnedi3-nns16-win8x4.hook:30-50): the two agree only for mean-removed weights (sum_k W[n][k] = 0; the shipped
    # files hold |sum| <= 1.4e-6).  A file that breaks this would run silently wrong: refuse it.
    for nm, wmat in (("W1", w1), ("W2", w2)):
        worst = float(np.abs(wmat.astype(np.float64).sum(axis=1)).max())
        if worst > 1e-5:
            raise HookError(f"{where}: NNEDI3 {nm} weights are not mean-removed (|sum| = {worst:.3e} > 1e-5); "
                            "this is not a shipped mpv-prescalers weight set")
    return Nnedi3Weights(w1.reshape(nns, 8, S), w2.reshape(nns, 8, S), b1, b2)


_AR_MIX_RE = re.compile(r"(mix\(res, clamp\(res, lo, hi\), )" + _FLT + r"\)")


def body_signature(body: str) -> str:
    """Identity of a pass body for the shipped-form check: sha256 of its text with the one documented tunable (the
    anti-ringing strength literal, README.md:64) masked."""
    return hashlib.sha256(_AR_MIX_RE.sub(r"\1S)", body).encode()).hexdigest()[:20]


def classify(hook: HookFile) -> Variant:
    """Work out which prescaler a parsed file is and extract everything the kernels need."""
    from .known_bodies import KNOWN_BODIES

    v = _classify(hook)   # structural checks first: they name what differs (Gaussian weights, stencils, LUT geometry ...)
    # The fused kernels implement the shipped shaders.  A file whose GLSL differs from every shipped pass body (a hand
    # edit the structural checks did not catch) must not be run as if it were the original: it is refused here
    # (prescale(..., runner='generic') runs the file's own GLSL instead).
    for p in hook.passes:
        if body_signature(p.body) not in KNOWN_BODIES:
            raise HookError(f"{hook.path}:{p.line}: pass {p.desc!r} differs from the supported form (its GLSL is not the text of "
                            "a shipped mpv-prescalers pass); the fused kernels refuse it -- use prescale(..., runner='generic')")
    return v


def _classify(hook: HookFile) -> Variant:
    first = hook.passes[0]
    fam = None
    for name, pat in _DESC_PATTERNS:
        m = pat.fullmatch(first.desc.strip())
        if m:
            fam = name
            break
    if fam is None:
        raise HookError(f"{hook.path}: unsupported shader {first.desc!r} (not an mpv-prescalers RAVU/NNEDI3 hook)")
    gd = m.groupdict()
    uses_gather = any("textureGather" in p.body for p in hook.passes)
    flavour = "compute" if any(p.compute for p in hook.passes) else ("gather" if uses_gather else "root")
    where = f"{hook.path}:{first.line}"
    hook_point = first.hook[0]

    if fam == "nnedi3":
        nns, win = int(gd["nns"]), (int(gd["wl"]), int(gd["ws"]))
        if win[0] != 8 or win[1] not in (4, 6) or nns not in (16, 32, 64, 128, 256):
            raise HookError(f"{where}: unsupported NNEDI3 geometry nns={nns} win={win}")
        py = [p for p in hook.passes if "double_y" in p.desc]
        px = [p for p in hook.passes if "double_x" in p.desc]
        if len(py) != 1 or len(px) != 1:
            raise HookError(f"{where}: expected one double_y and one double_x pass")
        v = Variant("nnedi3", flavour, "luma", nns=nns, win=win, scale=2, hook_point=hook_point)
        v.nn_y = _nnedi3_pass(py[0], nns, win, "y", f"{hook.path}:{py[0].line}")
        v.nn_x = _nnedi3_pass(px[0], nns, win, "x", f"{hook.path}:{px[0].line}")
        v.offset = (-0.5, -0.5)
        return v

    r = int(gd["r"])
    plane = gd.get("plane") or "luma"
    ar = bool(gd.get("ar"))
    if fam in ("ravu-zoom",) and r not in (2, 3) or r not in (2, 3, 4):
        raise HookError(f"{where}: unsupported radius {r}")
    want_hook = {"luma": "LUMA", "yuv": "NATIVE", "rgb": "MAIN"}[plane]
    if hook_point != want_hook:
        raise HookError(f"{where}: plane mode {plane} expects //!HOOK {want_hook}, file has {hook_point}")
    v = Variant(fam, flavour, plane, radius=r, ar=ar, hook_point=hook_point)
    v.scale = {"ravu-lite": 2, "ravu": 2, "ravu-3x": 3, "ravu-zoom": None}[fam]
    deep = flavour == "root" or fam == "ravu-3x"
    _extract_key_constants(v, first.body, where, deep)
    if fam == "ravu" and flavour == "root":
        # steps 2 and 3 must carry the same constants
        for p in hook.passes[1:3]:
            v2 = Variant(fam, flavour, plane, radius=r)
            _extract_key_constants(v2, p.body, f"{hook.path}:{p.line}", deep)
            if not np.array_equal(v2.gauss, v.gauss) or v2.strength_log2_scale != v.strength_log2_scale:
                raise HookError(f"{hook.path}:{p.line}: constants differ between RAVU steps")
    lut_names = [b for b in first.binds if b in hook.textures]
    if fam == "ravu-zoom" and ar:
        main = [b for b in lut_names if not b.endswith("_ar")]
        arl = [b for b in lut_names if b.endswith("_ar")]
        if len(main) != 1 or len(arl) != 1:
            raise HookError(f"{where}: RAVU-Zoom-AR needs a main and an _ar LUT (found {lut_names})")
        v.lut, v.lut_ar = hook.textures[main[0]], hook.textures[arl[0]]
        _check_lut(v, v.lut_ar, where)
    else:
        if len(lut_names) != 1:
            raise HookError(f"{where}: expected exactly one LUT texture, found {lut_names}")
        v.lut = hook.textures[lut_names[0]]
    _check_lut(v, v.lut, where)
    if ar:
        _ar_from_body(v, first.body, where)
    elif "clamp(res" not in first.body and "clamp(res0" not in first.body:
        raise HookError(f"{where}: non-AR variant without the final clamp")
    if fam == "ravu":
        v.offset = (-0.5, -0.5)
    return v


MISSING_IN_SNAPSHOT = (
    "ravu-zoom-ar-r3.hook",
    "ravu-zoom-ar-r3-yuv.hook",
    "ravu-zoom-ar-r3-rgb.hook",
)


def find_hook(name: Union[str, os.PathLike]) -> str:
    """Resolve a hook name to a path.

    An existing path is used as is.  Bare names (``'ravu-lite-ar-r3.hook'``,
    ``'compute/ravu-3x-r2.hook'``) are searched in ``$MPV_PRESCALERS_HOOKS`` (os.pathsep
    separated), ``/root/reference`` and ``<repo>/baseline/_ref/hooks``.
    """
    name = os.fspath(name)
    if os.path.isfile(name):
        return name
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dirs = [d for d in os.environ.get("MPV_PRESCALERS_HOOKS", "").split(os.pathsep) if d]
    dirs += ["/root/reference", os.path.join(here, "baseline", "_ref", "hooks")]
    for d in dirs:
        cand = os.path.join(d, name)
        if os.path.isfile(cand):
            return cand
    extra = ""
    if os.path.basename(name) in MISSING_IN_SNAPSHOT:
        extra = " (this variant is listed in the reference's .MISSING_LARGE_BLOBS: its anti-ringing LUT ravu_zoom_lut3_ar is not shipped)"
    raise HookError(f"hook file {name!r} not found in {dirs}{extra}")
