"""TEST INFRASTRUCTURE ONLY -- literal executor for the GLSL inside mpv-prescalers ``.hook`` files.

This module runs the reference's shader text *as written* (``vec4 hook()`` fragment bodies and the
``void hook()`` + ``imageStore`` compute bodies) on NumPy arrays, vectorised over all output
texels, in float32.  It shares no code with ``mpv_prescalers_b200`` (it has its own block
splitter, RPN evaluator and texture decoder) so that it can serve as an independent check of
both the product parser and the parametric NumPy oracle (``oracle/ravu_np.py`` etc.).

It is the closest thing to "running the reference" that this environment allows: there is no
mpv / libplacebo / GL / Vulkan here and the reference ships no tests or golden
vectors.  PARITY PIN: the shader text itself; host-defined semantics (clamp-to-edge addressing,
``HOOKED_pos`` = output texel centre / output size, SAVE/replace chaining, rgba16f storage
rounding) are restated from the mpv manual, not from executable reference
code -- in that sense parity is *unpinned by reference golden vectors*.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline leg may import this.

Supported GLSL subset: SURVEY.md App. A.4 (declarations, compound assignment, swizzles,
``if/else`` with early ``return``, one uniform ``for``, function-like macros, user functions,
``shared`` arrays + cooperative load loop + ``barrier()`` + ``imageStore`` for compute passes).
"""
from __future__ import annotations

import math
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32
I32 = np.int32


class GlslError(RuntimeError):
    pass


# ==============================================================================================
# hook-file container (independent of mpv_prescalers_b200.hookfile)
# ==============================================================================================


class XPass:
    def __init__(self, directives: Dict[str, List[str]], body: str, line: int):
        self.d = directives
        self.body = body
        self.line = line

    def one(self, key: str, default=None):
        return self.d[key][-1] if key in self.d else default

    @property
    def desc(self) -> str:
        return self.one("DESC", "")


class XTexture:
    def __init__(self, name, w, h, filt, data):
        self.name, self.w, self.h, self.filter, self.data = name, w, h, filt, data


def split_hook(text: str) -> Tuple[List[XPass], Dict[str, XTexture]]:
    """Split a hook file into passes and textures (``ravu-lite-ar-r3.hook:15-202``)."""
    lines = text.split("\n")
    passes, textures = [], {}
    i = 0
    while i < len(lines) and not lines[i].startswith("//!"):
        i += 1
    while i < len(lines):
        start = i
        d: Dict[str, List[str]] = {}
        while i < len(lines) and lines[i].startswith("//!"):
            k, _, v = lines[i][3:].strip().partition(" ")
            d.setdefault(k, []).append(v.strip())
            i += 1
        b0 = i
        while i < len(lines) and not lines[i].startswith("//!"):
            i += 1
        body = "\n".join(lines[b0:i])
        if "TEXTURE" in d:
            w, h = (int(t) for t in d["SIZE"][0].split())
            raw = np.frombuffer(bytes.fromhex(body.strip()), dtype="<f4")
            textures[d["TEXTURE"][0]] = XTexture(d["TEXTURE"][0], w, h, d.get("FILTER", ["NEAREST"])[0], raw.reshape(h, w, 4).astype(F32))
        else:
            passes.append(XPass(d, body, start + 1))
    return passes, textures


def rpn(expr: str, sizes: Dict[str, Tuple[int, int]]) -> float:
    st: List[float] = []
    for t in expr.split():
        if t in "+-*/<>=" and len(t) == 1:
            b, a = st.pop(), st.pop()
            st.append({"+": a + b, "-": a - b, "*": a * b, "/": a / b if t == "/" else 0.0, "<": float(a < b), ">": float(a > b), "=": float(a == b)}[t])
        elif "." in t and t.split(".")[0] in sizes:
            n, c = t.split(".")
            st.append(float(sizes[n][0 if c == "w" else 1]))
        else:
            st.append(float(t))
    assert len(st) == 1
    return st[0]


# ==============================================================================================
# preprocessing: function-like macros
# ==============================================================================================

_DEFINE_RE = re.compile(r"^#define\s+(\w+)\(([^)]*)\)\s+(.*)$")


def _split_args(s: str, pos: int) -> Tuple[List[str], int]:
    """s[pos] == '(' ; returns (args, index after the closing paren)."""
    depth, i, cur, args = 0, pos, [], []
    while True:
        c = s[i]
        if c == "(":
            depth += 1
            if depth > 1:
                cur.append(c)
        elif c == ")":
            depth -= 1
            if depth == 0:
                args.append("".join(cur).strip())
                return args, i + 1
            cur.append(c)
        elif c == "," and depth == 1:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
        i += 1


def _expand(text: str, macros: Dict[str, Tuple[List[str], str]]) -> str:
    if not macros:
        return text
    name_re = re.compile(r"\b(" + "|".join(map(re.escape, macros)) + r")\s*\(")
    out, pos = [], 0
    while True:
        m = name_re.search(text, pos)
        if not m:
            out.append(text[pos:])
            break
        out.append(text[pos : m.start()])
        params, body = macros[m.group(1)]
        args, end = _split_args(text, m.end() - 1)
        args = [_expand(a, macros) for a in args]
        rep = body
        if params:
            pr = re.compile(r"\b(" + "|".join(map(re.escape, params)) + r")\b")
            amap = dict(zip(params, args))
            rep = pr.sub(lambda mm: amap[mm.group(1)], body)
        out.append(_expand(rep, macros))
        pos = end
    return "".join(out)


def preprocess(src: str) -> str:
    macros: Dict[str, Tuple[List[str], str]] = {}
    out = []
    for line in src.split("\n"):
        s = line.strip()
        if s.startswith("#define"):
            m = _DEFINE_RE.match(s)
            if not m:
                raise GlslError(f"unsupported #define: {s}")
            macros[m.group(1)] = ([p.strip() for p in m.group(2).split(",") if p.strip()], m.group(3))
            continue
        if s.startswith("#pragma") or s.startswith("//"):
            continue
        line = re.sub(r"//.*$", "", line)
        out.append(_expand(line, macros))
    return "\n".join(out)


# ==============================================================================================
# tokenizer / parser
# ==============================================================================================

_TOK_RE = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|\d+u?)|(?P<id>[A-Za-z_]\w*)|(?P<op>\+\+|--|\+=|-=|\*=|/=|<=|>=|==|!=|&&|\|\||[-+*/%<>=!(){}\[\],;.?:]))"
)

TYPES = {"void", "float", "int", "uint", "bool", "vec2", "vec3", "vec4", "ivec2", "ivec3", "uvec2", "uvec3", "mat4x3"}


def tokenize(src: str) -> List[Tuple[str, str]]:
    toks, pos, n = [], 0, len(src)
    while pos < n:
        m = _TOK_RE.match(src, pos)
        if not m or m.end() == pos:
            if src[pos:].strip() == "":
                break
            raise GlslError(f"cannot tokenize at {src[pos:pos + 40]!r}")
        pos = m.end()
        if m.group("num") is not None:
            toks.append(("num", m.group("num")))
        elif m.group("id") is not None:
            toks.append(("id", m.group("id")))
        else:
            toks.append(("op", m.group("op")))
    toks.append(("eof", ""))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.p = toks, 0

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def accept(self, val):
        if self.t[self.p][1] == val and self.t[self.p][0] != "num":
            self.p += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise GlslError(f"expected {val!r}, got {self.t[self.p]!r} near token {self.p}")

    # ---- top level --------------------------------------------------------------------------
    def parse_unit(self):
        items = []
        while self.peek()[0] != "eof":
            items.append(self.parse_global())
        return items

    def parse_global(self):
        quals = []
        while self.peek()[1] in ("const", "shared", "uniform"):
            quals.append(self.next()[1])
        typ = self.next()[1]
        if typ not in TYPES:
            raise GlslError(f"expected a type, got {typ!r}")
        name = self.next()[1]
        if self.accept("("):
            params = []
            if not self.accept(")"):
                while True:
                    ptyp = self.next()[1]
                    pname = self.next()[1]
                    plen = None
                    if self.accept("["):
                        plen = int(self.next()[1])
                        self.expect("]")
                    params.append((ptyp, pname, plen))
                    if self.accept(")"):
                        break
                    self.expect(",")
            body = self.parse_block()
            return ("func", typ, name, params, body)
        self.p -= 1
        decl = self.parse_decl_rest(typ, quals)
        return decl

    def parse_decl_rest(self, typ, quals):
        decls = []
        while True:
            name = self.next()[1]
            alen = None
            if self.accept("["):
                alen = int(self.next()[1])
                self.expect("]")
            init = self.parse_assign_expr() if self.accept("=") else None
            decls.append((name, alen, init))
            if self.accept(";"):
                break
            self.expect(",")
        return ("decl", typ, tuple(quals), decls)

    def parse_block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.parse_stmt())
        return ("block", stmts)

    def parse_stmt(self):
        k, v = self.peek()
        if v == "{" and k == "op":
            return self.parse_block()
        if v == ";" and k == "op":
            self.next()
            return ("block", [])
        if k == "id":
            if v == "if":
                self.next()
                self.expect("(")
                c = self.parse_expr()
                self.expect(")")
                a = self.parse_stmt()
                b = self.parse_stmt() if self.accept("else") else None
                return ("if", c, a, b)
            if v == "for":
                self.next()
                self.expect("(")
                init = self.parse_stmt()
                cond = self.parse_expr()
                self.expect(";")
                step = self.parse_simple()
                self.expect(")")
                return ("for", init, cond, step, self.parse_stmt())
            if v == "return":
                self.next()
                e = None if self.peek()[1] == ";" else self.parse_expr()
                self.expect(";")
                return ("return", e)
            if v in ("const",) or (v in TYPES and self.peek(1)[0] == "id"):
                quals = []
                while self.peek()[1] == "const":
                    quals.append(self.next()[1])
                typ = self.next()[1]
                return self.parse_decl_rest(typ, quals)
        s = self.parse_simple()
        self.expect(";")
        return s

    def parse_simple(self):
        lhs = self.parse_expr()
        k, v = self.peek()
        if k == "op" and v in ("=", "+=", "-=", "*=", "/="):
            self.next()
            rhs = self.parse_assign_expr()
            return ("assign", v, lhs, rhs)
        if k == "op" and v in ("++", "--"):
            self.next()
            return ("assign", "+=" if v == "++" else "-=", lhs, ("int", 1))
        return ("expr", lhs)

    def parse_assign_expr(self):
        return self.parse_expr()

    # ---- expressions ------------------------------------------------------------------------
    _BIN = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]

    def parse_expr(self, level=0):
        if level == len(self._BIN):
            return self.parse_unary()
        lhs = self.parse_expr(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self._BIN[level]:
            op = self.next()[1]
            rhs = self.parse_expr(level + 1)
            lhs = ("bin", op, lhs, rhs)
        return lhs

    def parse_unary(self):
        k, v = self.peek()
        if k == "op" and v in ("-", "+", "!"):
            self.next()
            return ("un", v, self.parse_unary())
        return self.parse_postfix()

    def parse_postfix(self):
        k, v = self.next()
        if k == "num":
            if v.endswith("u"):
                e = ("int", int(v[:-1]))
            elif re.fullmatch(r"\d+", v):
                e = ("int", int(v))
            else:
                e = ("float", v)
        elif k == "id":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.parse_expr())
                        if self.accept(")"):
                            break
                        self.expect(",")
                e = ("call", v, args)
            else:
                e = ("var", v)
        elif v == "(":
            e = self.parse_expr()
            self.expect(")")
        else:
            raise GlslError(f"unexpected token {v!r}")
        while True:
            if self.accept("["):
                idx = self.parse_expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.peek() == ("op", "."):
                self.next()
                e = ("field", e, self.next()[1])
            else:
                return e


# ==============================================================================================
# values
# ==============================================================================================


class V:
    """kind: 'f' float scalar, 'i' int scalar, 'b' bool scalar, 'v' float vector, 'iv' int vector,
    'm' matrix [cols, rows].  Array layout: lead + (n,) (scalars: n = 1), matrices lead + (c, r)."""

    __slots__ = ("k", "a")

    def __init__(self, k, a):
        self.k, self.a = k, a

    @property
    def lead(self):
        return self.a.shape[: -2 if self.k == "m" else -1]

    @property
    def n(self):
        return self.a.shape[-1]


def fconst(x) -> V:
    return V("f", np.asarray([x], dtype=F32))


def iconst(x) -> V:
    return V("i", np.asarray([x], dtype=I32))


_SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


def _is_float(v: V) -> bool:
    return v.k in ("f", "v", "m")


def _tofloat(v: V) -> V:
    if v.k in ("i", "b"):
        return V("f", v.a.astype(F32))
    if v.k == "iv":
        return V("v", v.a.astype(F32))
    return v


def _toint(v: V) -> V:
    if v.k == "f":
        return V("i", np.trunc(v.a).astype(I32))
    if v.k == "b":
        return V("i", v.a.astype(I32))
    if v.k == "v":
        return V("iv", np.trunc(v.a).astype(I32))
    return v


def _binop(op: str, x: V, y: V) -> V:
    if op in ("&&", "||"):
        f = np.logical_and if op == "&&" else np.logical_or
        return V("b", f(x.a.astype(bool), y.a.astype(bool)))
    if _is_float(x) != _is_float(y):
        # GLSL has no implicit int->float in these shaders except literals like `x / 10` on ints;
        # promote the int side (e.g. `float * int-literal` never occurs, but be permissive)
        x, y = _tofloat(x), _tofloat(y)
    xa, ya = x.a, y.a
    if x.k == "m" and y.k != "m":
        ya = ya[..., None]
    elif y.k == "m" and x.k != "m":
        xa = xa[..., None]
    isint = not _is_float(x)
    if op in ("<", ">", "<=", ">=", "==", "!="):
        f = {"<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal, "==": np.equal, "!=": np.not_equal}[op]
        return V("b", f(xa, ya))
    if op == "+":
        r = xa + ya
    elif op == "-":
        r = xa - ya
    elif op == "*":
        if x.k == "m" and y.k == "m":
            raise GlslError("matrix*matrix not supported (use matrixCompMult)")
        r = xa * ya
    elif op == "/":
        if isint:
            # C-style truncating division (operands are non-negative in these shaders)
            r = (np.trunc(xa.astype(np.float64) / ya.astype(np.float64))).astype(I32)
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                r = xa / ya
    elif op == "%":
        r = np.fmod(xa, ya)
    else:
        raise GlslError(f"operator {op}")
    if isint:
        k = "iv" if (x.k == "iv" or y.k == "iv") else "i"
        return V(k, r.astype(I32))
    k = "m" if "m" in (x.k, y.k) else ("v" if "v" in (x.k, y.k) else "f")
    return V(k, r.astype(F32, copy=False))


def _where(mask, new: V, old: V) -> V:
    if mask is None:
        return new
    m = mask[..., None, None] if new.k == "m" else mask[..., None]
    return V(new.k, np.where(m, new.a, old.a))


def _concat(args: List[V], n: int, kind: str) -> V:
    parts = []
    for a in args:
        a = _tofloat(a) if kind == "v" else _toint(a)
        parts.append(a.a)
    if len(parts) == 1 and parts[0].shape[-1] == 1:
        return V(kind, np.repeat(parts[0], n, axis=-1))
    lead = np.broadcast_shapes(*[p.shape[:-1] for p in parts])
    parts = [np.broadcast_to(p, lead + p.shape[-1:]) for p in parts]
    r = np.concatenate(parts, axis=-1)
    if r.shape[-1] < n:
        raise GlslError(f"constructor with {r.shape[-1]} components for {kind}{n}")
    return V(kind, r[..., :n])


# ==============================================================================================
# texture sampling (host semantics)
# ==============================================================================================


class Tex:
    def __init__(self, data: np.ndarray, filt: str = "NEAREST"):
        self.data = np.ascontiguousarray(data, dtype=F32)  # [h, w, 4]
        self.h, self.w = data.shape[:2]
        self.filter = filt

    def size_v(self) -> V:
        return V("v", np.asarray([self.w, self.h], dtype=F32))

    def pt_v(self) -> V:
        return V("v", np.asarray([F32(1) / F32(self.w), F32(1) / F32(self.h)], dtype=F32))

    def fetch(self, ix, iy):
        ix = np.clip(ix, 0, self.w - 1)
        iy = np.clip(iy, 0, self.h - 1)
        return self.data[iy, ix]

    def sample(self, coord: V, linear: bool) -> V:
        c = coord.a
        u = c[..., 0] * F32(self.w)
        v = c[..., 1] * F32(self.h)
        if not linear:
            return V("v", self.fetch(np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)))
        u = u - F32(0.5)
        v = v - F32(0.5)
        u0, v0 = np.floor(u), np.floor(v)
        fu, fv = (u - u0)[..., None], (v - v0)[..., None]
        x0, y0 = u0.astype(np.int64), v0.astype(np.int64)
        t00, t10 = self.fetch(x0, y0), self.fetch(x0 + 1, y0)
        t01, t11 = self.fetch(x0, y0 + 1), self.fetch(x0 + 1, y0 + 1)
        one = F32(1)
        top = t00 * (one - fu) + t10 * fu
        bot = t01 * (one - fu) + t11 * fu
        return V("v", (top * (one - fv) + bot * fv).astype(F32))


# ==============================================================================================
# interpreter
# ==============================================================================================


class _Frame:
    def __init__(self):
        self.scopes: List[Dict[str, Any]] = [{}]
        self.ret: Optional[V] = None
        self.ret_mask = None  # lanes that already returned (None = none yet)
        self.done = False


class Shared:
    def __init__(self, n: int, comps: int):
        self.n, self.comps, self.a = n, comps, None  # a: [GY, GX, n, comps]


class Interp:
    def __init__(self, src: str, textures: Dict[str, Tex], hooked: str = "HOOKED"):
        self.unit = Parser(tokenize(preprocess(src))).parse_unit()
        self.funcs = {it[2]: it for it in self.unit if it[0] == "func"}
        self.tex = textures
        self.globals: Dict[str, Any] = {}
        self.types: Dict[str, str] = {}
        self.out_image = None
        self.coop = False  # inside the cooperative shared-memory load loop

    # ---- environment ------------------------------------------------------------------------
    def lookup(self, fr: _Frame, name: str):
        for sc in reversed(fr.scopes):
            if name in sc:
                return sc[name]
        if name in self.globals:
            return self.globals[name]
        raise GlslError(f"undefined identifier {name!r}")

    def store(self, fr: _Frame, name: str, val):
        for sc in reversed(fr.scopes):
            if name in sc:
                sc[name] = val
                return
        if name in self.globals:
            self.globals[name] = val
            return
        raise GlslError(f"assignment to undeclared {name!r}")

    def zero(self, typ: str) -> V:
        if typ == "float":
            return fconst(0.0)
        if typ in ("int", "uint"):
            return iconst(0)
        if typ == "bool":
            return V("b", np.asarray([False]))
        if typ.startswith("vec"):
            return V("v", np.zeros(int(typ[3]), F32))
        if typ.startswith(("ivec", "uvec")):
            return V("iv", np.zeros(int(typ[4]), I32))
        if typ == "mat4x3":
            return V("m", np.zeros((4, 3), F32))
        raise GlslError(f"type {typ}")

    def coerce(self, typ: str, v: V) -> V:
        z = self.zero(typ)
        if z.k in ("f", "v", "m") and not _is_float(v):
            v = _tofloat(v)
        if z.k in ("i", "iv") and _is_float(v):
            raise GlslError(f"implicit float->int conversion to {typ}")
        if z.k == "v" and v.k == "f":
            raise GlslError(f"scalar assigned to {typ}")
        return V(z.k, v.a)

    # ---- execution --------------------------------------------------------------------------
    def active(self, fr: _Frame, mask):
        if fr.ret_mask is None:
            return mask
        nm = ~fr.ret_mask
        return nm if mask is None else (mask & nm)

    def exec_block(self, fr, stmts, mask):
        fr.scopes.append({})
        for s in stmts:
            if fr.done:
                break
            self.exec_stmt(fr, s, mask)
        fr.scopes.pop()

    def exec_stmt(self, fr: _Frame, s, mask):
        kind = s[0]
        if kind == "block":
            self.exec_block(fr, s[1], mask)
        elif kind == "decl":
            self.exec_decl(fr, s, fr.scopes[-1])
        elif kind == "assign":
            self.exec_assign(fr, s, self.active(fr, mask))
        elif kind == "expr":
            self.eval(fr, s[1], self.active(fr, mask))
        elif kind == "return":
            act = self.active(fr, mask)
            val = self.eval(fr, s[1], act) if s[1] is not None else None
            if act is None:
                fr.ret, fr.done = val, True
            else:
                if val is not None:
                    fr.ret = val if fr.ret is None else _where(act, val, fr.ret)
                fr.ret_mask = act if fr.ret_mask is None else (fr.ret_mask | act)
        elif kind == "if":
            c = self.eval(fr, s[1], mask)
            if c.a.size == 1:
                if bool(c.a.reshape(-1)[0]):
                    self.exec_stmt(fr, s[2], mask)
                elif s[3] is not None:
                    self.exec_stmt(fr, s[3], mask)
            else:
                cm = c.a[..., 0].astype(bool)
                self.exec_stmt(fr, s[2], cm if mask is None else (mask & cm))
                if s[3] is not None:
                    self.exec_stmt(fr, s[3], ~cm if mask is None else (mask & ~cm))
        elif kind == "for":
            self.exec_for(fr, s, mask)
        else:
            raise GlslError(f"statement {kind}")

    def exec_decl(self, fr, s, scope):
        _, typ, quals, decls = s
        for name, alen, init in decls:
            if "shared" in quals:
                comps = {"float": 1, "vec3": 3, "vec4": 4, "vec2": 2}[typ]
                scope[name] = Shared(alen, comps)
                continue
            if alen is not None:
                scope[name] = [self.zero(typ) for _ in range(alen)]
                self.types[name] = typ
                continue
            val = self.coerce(typ, self.eval(fr, init, None)) if init is not None else self.zero(typ)
            scope[name] = val
            self.types[name] = typ

    def exec_for(self, fr, s, mask):
        _, init, cond, step, body = s
        fr.scopes.append({})
        coop = init[0] == "decl" and "gl_LocalInvocationIndex" in repr(init)
        if coop:
            # cooperative tile load: every id in [0, bound) is written by exactly one invocation of
            # the group; emulate with a uniform loop over id (compute/ravu-3x-r2.hook:27-30)
            name = init[3][0][0]
            fr.scopes[-1][name] = iconst(0)
            self.coop = True
        else:
            self.exec_stmt(fr, init, mask)
        guard = 0
        while True:
            c = self.eval(fr, cond, mask)
            if c.a.size != 1:
                raise GlslError("non-uniform for-loop condition")
            if not bool(c.a.reshape(-1)[0]):
                break
            self.exec_stmt(fr, body, mask)
            if coop:
                self.store(fr, name, _binop("+", self.lookup(fr, name), iconst(1)))
            else:
                self.exec_stmt(fr, step, mask)
            guard += 1
            if guard > 100000:
                raise GlslError("runaway loop")
        self.coop = False
        fr.scopes.pop()

    # ---- assignment -------------------------------------------------------------------------
    def exec_assign(self, fr, s, mask):
        _, op, lhs, rhs = s
        val = self.eval(fr, rhs, mask)
        if op != "=":
            cur = self.eval(fr, lhs, mask)
            val = _binop(op[0], cur, val)
        self.assign(fr, lhs, val, mask)

    def assign(self, fr, lhs, val: V, mask):
        kind = lhs[0]
        if kind == "var":
            old = self.lookup(fr, lhs[1])
            if isinstance(old, V):
                if old.k in ("f", "v", "m") and not _is_float(val):
                    val = _tofloat(val)
                if old.k != val.k and not (old.k == "v" and val.k == "v"):
                    if not (old.k in ("i",) and val.k == "i"):
                        raise GlslError(f"type mismatch assigning {val.k} to {lhs[1]} ({old.k})")
                if old.k in ("v", "iv") and old.n != val.n:
                    raise GlslError(f"vector size mismatch assigning to {lhs[1]}")
                self.store(fr, lhs[1], _where(mask, val, old) if mask is not None else val)
            else:
                raise GlslError(f"cannot assign whole array {lhs[1]}")
            return
        if kind == "index":
            base = lhs[1]
            container = self.eval_lvalue_container(fr, base)
            idx = self.eval(fr, lhs[2], mask)
            if isinstance(container, Shared):
                if not self.coop or idx.a.size != 1:
                    raise GlslError("shared array store outside the cooperative load loop")
                i = int(idx.a.reshape(-1)[0])
                v = _tofloat(val).a
                if container.a is None:
                    container.a = np.zeros(self.group_shape + (container.n, container.comps), F32)
                container.a[:, :, i, :] = np.broadcast_to(v, self.group_shape + (1, 1) + v.shape[-1:])[:, :, 0, 0, :]
                return
            if idx.a.size != 1:
                raise GlslError("varying index on the left-hand side")
            i = int(idx.a.reshape(-1)[0])
            if isinstance(container, list):
                old = container[i]
                new = list(container)
                new[i] = _where(mask, self._match(old, val), self._bc(old, val)) if mask is not None else self._match(old, val)
                self.assign_container(fr, base, new)
                return
            old = container
            if old.k == "m":
                if val.k != "v":
                    raise GlslError("matrix column assignment needs a vector")
                lead = np.broadcast_shapes(old.lead, val.lead)
                a = np.broadcast_to(old.a, lead + old.a.shape[-2:]).copy()
                nv = np.broadcast_to(val.a, lead + val.a.shape[-1:])
                a[..., i, :] = nv if mask is None else np.where(mask[..., None], nv, a[..., i, :])
                self.assign(fr, base, V("m", a), None)
                return
            val = _tofloat(val) if old.k == "v" else val
            lead = np.broadcast_shapes(old.lead, val.lead)
            a = np.broadcast_to(old.a, lead + (old.n,)).copy()
            nv = np.broadcast_to(val.a, lead + (1,))[..., 0]
            a[..., i] = nv if mask is None else np.where(mask, nv, a[..., i])
            self.assign(fr, base, V(old.k, a), None)
            return
        if kind == "field":
            old = self.eval(fr, lhs[1], mask)
            idxs = [_SWZ[c] for c in lhs[2]]
            lead = np.broadcast_shapes(old.lead, val.lead)
            a = np.broadcast_to(old.a, lead + (old.n,)).copy()
            nv = np.broadcast_to(_tofloat(val).a if old.k == "v" else val.a, lead + (len(idxs),))
            for j, i in enumerate(idxs):
                a[..., i] = nv[..., j] if mask is None else np.where(mask, nv[..., j], a[..., i])
            self.assign(fr, lhs[1], V(old.k, a), None)
            return
        raise GlslError(f"bad lvalue {kind}")

    @staticmethod
    def _match(old: V, val: V) -> V:
        if old.k in ("f", "v", "m") and not _is_float(val):
            val = _tofloat(val)
        return V(old.k, val.a)

    @staticmethod
    def _bc(old: V, val: V) -> V:
        return old

    def eval_lvalue_container(self, fr, e):
        if e[0] == "var":
            return self.lookup(fr, e[1])
        if e[0] == "index":
            c = self.eval_lvalue_container(fr, e[1])
            idx = self.eval(fr, e[2], None)
            i = int(idx.a.reshape(-1)[0])
            if isinstance(c, list):
                return c[i]
            return self.eval(fr, e, None)
        return self.eval(fr, e, None)

    def assign_container(self, fr, e, new):
        if e[0] == "var":
            self.store(fr, e[1], new)
        else:
            raise GlslError("nested array assignment")

    # ---- expressions ------------------------------------------------------------------------
    def eval(self, fr, e, mask) -> V:
        k = e[0]
        if k == "float":
            return fconst(F32(e[1]))
        if k == "int":
            return iconst(e[1])
        if k == "var":
            v = self.lookup(fr, e[1])
            return v
        if k == "un":
            x = self.eval(fr, e[2], mask)
            if e[1] == "-":
                return V(x.k, -x.a)
            if e[1] == "!":
                return V("b", ~x.a.astype(bool))
            return x
        if k == "bin":
            return _binop(e[1], self.eval(fr, e[2], mask), self.eval(fr, e[3], mask))
        if k == "field":
            x = self.eval(fr, e[1], mask)
            idxs = [_SWZ[c] for c in e[2]]
            if len(idxs) == 1:
                return V("f" if x.k == "v" else "i", x.a[..., idxs[0] : idxs[0] + 1])
            return V(x.k, x.a[..., idxs])
        if k == "index":
            base = self.eval_any(fr, e[1], mask)
            idx = self.eval(fr, e[2], mask)
            if isinstance(base, Shared):
                ii = idx.a[..., 0]
                if ii.ndim == 0:
                    r = base.a[:, :, int(ii), :][:, :, None, None, :]
                else:
                    # ii: [1,1,LY,LX] (or broadcastable) -> gather per group
                    ii2 = np.broadcast_to(ii, (1, 1) + self.local_shape)[0, 0]
                    r = base.a[:, :, ii2, :]
                return V("f" if base.comps == 1 else "v", r)
            if isinstance(base, list):
                if idx.a.size != 1:
                    raise GlslError("varying index into a local array")
                return base[int(idx.a.reshape(-1)[0])]
            if base.k == "m":
                return V("v", base.a[..., int(idx.a.reshape(-1)[0]), :])
            if idx.a.size == 1:
                i = int(idx.a.reshape(-1)[0])
                return V("f" if base.k == "v" else "i", base.a[..., i : i + 1])
            lead = np.broadcast_shapes(base.lead, idx.lead)
            a = np.broadcast_to(base.a, lead + (base.n,))
            ii = np.broadcast_to(idx.a, lead + (1,)).astype(np.int64)
            return V("f" if base.k == "v" else "i", np.take_along_axis(a, ii, axis=-1))
        if k == "call":
            return self.call(fr, e[1], e[2], mask)
        raise GlslError(f"expression {k}")

    def eval_any(self, fr, e, mask):
        if e[0] == "var":
            return self.lookup(fr, e[1])
        return self.eval(fr, e, mask)

    # ---- calls ------------------------------------------------------------------------------
    def call(self, fr, name: str, argexprs, mask) -> V:
        if name in self.funcs:
            _, rtyp, _, params, body = self.funcs[name]
            nf = _Frame()
            for (ptyp, pname, plen), ae in zip(params, argexprs):
                nf.scopes[0][pname] = self.eval_any(fr, ae, mask)
            self.exec_stmt(nf, body, mask)
            return nf.ret
        if name == "imageStore":
            pos = self.eval(fr, argexprs[1], mask)
            val = self.eval(fr, argexprs[2], mask)
            self.image_store(pos, val, mask)
            return None
        if name == "barrier":
            return None
        if name == "texture":
            t = self.tex[argexprs[0][1]]
            return t.sample(self.eval(fr, argexprs[1], mask), t.filter == "LINEAR")
        args = [self.eval_any(fr, a, mask) for a in argexprs]
        m = re.fullmatch(r"(\w+?)_(texOff|tex)", name)
        if m and m.group(1) in self.tex:
            t = self.tex[m.group(1)]
            if m.group(2) == "tex":
                return t.sample(args[0], False)
            # NAME_texOff(off) = NAME_tex(NAME_pos + NAME_pt * off)
            pos = self.globals[m.group(1) + "_pos"]
            coord = _binop("+", pos, _binop("*", t.pt_v(), _tofloat(args[0])))
            return t.sample(coord, False)
        return self.builtin(name, args)

    def builtin(self, name: str, a: List[V]) -> V:
        if name in ("vec2", "vec3", "vec4"):
            return _concat(a, int(name[3]), "v")
        if name in ("ivec2", "ivec3", "uvec2", "uvec3"):
            return _concat(a, int(name[4]), "iv")
        if name == "float":
            return _tofloat(a[0])
        if name in ("int", "uint"):
            return _toint(a[0])
        if name == "mat4x3":
            cols = [np.asarray(_tofloat(c).a) for c in a]
            if len(cols) == 1 and cols[0].shape[-1] == 1:  # mat4x3(s): s on the diagonal
                m = np.zeros(cols[0].shape[:-1] + (4, 3), F32)
                for i in range(3):
                    m[..., i, i] = cols[0][..., 0]
                return V("m", m)
            if len(cols) != 4:
                raise GlslError("mat4x3 needs 4 column vectors")
            lead = np.broadcast_shapes(*[c.shape[:-1] for c in cols])
            return V("m", np.stack([np.broadcast_to(c, lead + (3,)) for c in cols], axis=-2).astype(F32))
        if name == "matrixCompMult":
            return V("m", a[0].a * a[1].a)
        if name == "outerProduct":  # outerProduct(c: vec3, r: vec4) -> mat4x3: m[col j][row i] = c[i]*r[j]
            c, r = a[0].a, a[1].a
            return V("m", (r[..., :, None] * c[..., None, :]).astype(F32))
        if name == "dot":
            x, y = _tofloat(a[0]), _tofloat(a[1])
            prod = x.a * y.a
            # left-to-right summation of the component products (matches a scalar GLSL expansion)
            acc = prod[..., 0:1]
            for i in range(1, prod.shape[-1]):
                acc = acc + prod[..., i : i + 1]
            return V("f", acc.astype(F32))
        if name == "intBitsToFloat":
            return V("f", a[0].a.astype(I32).view(F32))
        x = a[0]
        one = {
            "floor": np.floor,
            "sqrt": lambda t: np.sqrt(t),
            "exp": np.exp,
            "log2": np.log2,
            "abs": np.abs,
        }
        with np.errstate(all="ignore"):
            if name in one:
                return V(x.k, one[name](x.a).astype(x.a.dtype))
            if name == "fract":
                return V(x.k, (x.a - np.floor(x.a)).astype(F32))
            if name == "inversesqrt":
                return V(x.k, (F32(1) / np.sqrt(x.a)).astype(F32))
            if name == "max":
                return self._k2(x, a[1], np.maximum)
            if name == "min":
                return self._k2(x, a[1], np.minimum)
            if name == "clamp":
                lo, hi = _tofloat(a[1]), _tofloat(a[2])
                # GLSL: min(max(x, lo), hi)
                r = np.minimum(np.maximum(x.a, self._mx(lo, x)), self._mx(hi, x))
                return V(x.k, r.astype(F32))
            if name == "mod":
                y = _tofloat(a[1])
                ya = self._mx(y, x)
                return V(x.k, (x.a - ya * np.floor(x.a / ya)).astype(F32))
            if name == "atan":
                if len(a) == 2:
                    return V(x.k, np.arctan2(x.a, a[1].a).astype(F32))
                return V(x.k, np.arctan(x.a).astype(F32))
            if name == "mix":
                y, t = a[1], a[2]
                x, y = _tofloat(x), _tofloat(y)
                kk = "m" if "m" in (x.k, y.k) else ("v" if "v" in (x.k, y.k, t.k) else "f")
                if t.k == "b":
                    return V(kk, np.where(t.a, y.a, x.a).astype(F32))
                t = _tofloat(t)
                # x*(1-a) + y*a
                return V(kk, (x.a * (F32(1) - t.a) + y.a * t.a).astype(F32))
        raise GlslError(f"unsupported function {name}")

    @staticmethod
    def _mx(y: V, x: V):
        return y.a[..., None] if (x.k == "m" and y.k != "m") else y.a

    def _k2(self, x: V, y: V, f) -> V:
        if _is_float(x) or _is_float(y):
            x, y = _tofloat(x), _tofloat(y)
        k = "m" if "m" in (x.k, y.k) else ("v" if "v" in (x.k, y.k) else ("iv" if "iv" in (x.k, y.k) else x.k))
        return V(k, f(self._mx(x, y), self._mx(y, x)))

    # ---- compute-mode image store -----------------------------------------------------------
    def image_store(self, pos: V, val: V, mask):
        full = self.group_shape + self.local_shape
        p = np.broadcast_to(pos.a, full + (2,))
        v = np.broadcast_to(_tofloat(val).a, full + (4,))
        H, W = self.out_image.shape[:2]
        ok = (p[..., 0] >= 0) & (p[..., 0] < W) & (p[..., 1] >= 0) & (p[..., 1] < H)
        if mask is not None:
            ok &= np.broadcast_to(mask, full)
        self.out_image[p[..., 1][ok], p[..., 0][ok]] = v[ok]


# ==============================================================================================
# running passes
# ==============================================================================================


def _pos_grid(ow: int, oh: int) -> V:
    """HOOKED_pos for every output texel: ((ox + 0.5) / OW, (oy + 0.5) / OH) in float32 (App. D.6)."""
    xs = (np.arange(ow, dtype=F32) + F32(0.5)) / F32(ow)
    ys = (np.arange(oh, dtype=F32) + F32(0.5)) / F32(oh)
    a = np.empty((oh, ow, 2), F32)
    a[..., 0] = xs[None, :]
    a[..., 1] = ys[:, None]
    return V("v", a)


def run_pass(p: XPass, textures: Dict[str, Tex], out_wh: Tuple[int, int]) -> np.ndarray:
    """Execute one pass; returns the [oh, ow, 4] float32 result."""
    ow, oh = out_wh
    binds = [b for b in p.d.get("BIND", [])]
    tex = {b: textures[b] for b in binds if b in textures}
    it = Interp(p.body, tex)
    for name, t in tex.items():
        it.globals[name + "_size"] = t.size_v()
        it.globals[name + "_pt"] = t.pt_v()
        it.globals[name + "_mul"] = fconst(1.0)
    compute = p.one("COMPUTE")
    fr = _Frame()
    # globals declared in the shader (const vec3 color_primary, shared arrays)
    for item in it.unit:
        if item[0] == "decl":
            it.exec_decl(fr, item, it.globals)
    if not compute:
        grid = _pos_grid(ow, oh)
        for name in tex:
            it.globals[name + "_pos"] = grid
        res = it.call(fr, "hook", [], None)
        out = np.broadcast_to(_tofloat(res).a, (oh, ow, 4)).astype(F32)
        return np.ascontiguousarray(out)
    c = [int(t) for t in compute.split()]
    bw, bh = c[0], c[1]
    tw, th = (c[2], c[3]) if len(c) == 4 else (bw, bh)
    gx, gy = -(-ow // bw), -(-oh // bh)
    it.group_shape, it.local_shape = (gy, gx), (th, tw)
    wid = np.zeros((gy, gx, 1, 1, 3), I32)
    wid[..., 0] = np.arange(gx)[None, :, None, None]
    wid[..., 1] = np.arange(gy)[:, None, None, None]
    lid = np.zeros((1, 1, th, tw, 3), I32)
    lid[..., 0] = np.arange(tw)[None, None, None, :]
    lid[..., 1] = np.arange(th)[None, None, :, None]
    wsz = np.asarray([tw, th, 1], I32)
    it.globals["gl_WorkGroupID"] = V("iv", wid)
    it.globals["gl_WorkGroupSize"] = V("iv", wsz)
    it.globals["gl_LocalInvocationID"] = V("iv", lid)
    it.globals["gl_LocalInvocationIndex"] = V("i", (lid[..., 1:2] * tw + lid[..., 0:1]).astype(I32))
    it.globals["gl_GlobalInvocationID"] = V("iv", (wid * wsz + lid).astype(I32))
    it.out_image = np.zeros((oh, ow, 4), F32)
    it.call(fr, "hook", [], None)
    return it.out_image


def run_hook(
    path: str,
    image: np.ndarray,
    out_size: Optional[Tuple[int, int]] = None,
    lut_precision: str = "fp16",
    hook_point: Optional[str] = None,
    return_saved: bool = False,
):
    """Run every pass of a hook file on ``image``.

    image: float32 [H, W] (luma) or [H, W, 3] (yuv / rgb).  ``out_size`` = (OW, OH) of the final
    target (mpv's OUTPUT); default is the hook's natural factor.  Returns (output [OH', OW', C],
    accumulated offset (x, y), applied flag).  Intermediates are kept in float32 (no FBO rounding).
    """
    with open(path) as f:
        passes, textures = split_hook(f.read())
    img = np.asarray(image, dtype=F32)
    ch = 1 if img.ndim == 2 else img.shape[2]
    h, w = img.shape[:2]
    rgba = np.zeros((h, w, 4), F32)
    rgba[..., :ch] = img.reshape(h, w, ch)
    rgba[..., 3] = 1.0
    texs: Dict[str, Tex] = {}
    for name, t in textures.items():
        data = t.data
        if lut_precision == "fp16":
            data = data.astype(np.float16).astype(F32)
        texs[name] = Tex(data, t.filter)
    if hook_point is None:
        hook_point = passes[0].d["HOOK"][0]
    texs["HOOKED"] = Tex(rgba)
    if out_size is None:
        d0 = passes[0].desc
        fac = 3 if "3x" in d0 else 2
        if "Zoom" in d0:
            raise ValueError("out_size is required for ravu-zoom hooks")
        out_size = (w * fac, h * fac)
    offset = [0.0, 0.0]
    applied = False
    saved = {}
    for p in passes:
        if hook_point not in p.d["HOOK"]:
            continue
        hk = texs["HOOKED"]
        sizes = {"HOOKED": (hk.w, hk.h), "OUTPUT": out_size, "LUMA": (w, h), "NATIVE": (w, h), "MAIN": (w, h)}
        for name, t in texs.items():
            sizes.setdefault(name, (t.w, t.h))
        when = p.one("WHEN")
        if when is not None and rpn(when, sizes) == 0.0:
            continue
        ow = int(rpn(p.one("WIDTH"), sizes)) if p.one("WIDTH") else hk.w
        oh = int(rpn(p.one("HEIGHT"), sizes)) if p.one("HEIGHT") else hk.h
        res = run_pass(p, texs, (ow, oh))
        applied = True
        save = p.one("SAVE")
        if save:
            texs[save] = Tex(res)
            saved[save] = res
        else:
            texs["HOOKED"] = Tex(res)
        off = p.one("OFFSET")
        if off and off != "ALIGN":
            ox, oy = (float(t) for t in off.split())
            # an offset declared by a pass is in units of that pass's output pixels; later passes
            # that scale the image scale earlier offsets with it
            offset[0] += ox
            offset[1] += oy
    out = texs["HOOKED"].data[..., :ch]
    if ch == 1:
        out = out[..., 0]
    if return_saved:
        return np.ascontiguousarray(out), tuple(offset), applied, saved
    return np.ascontiguousarray(out), tuple(offset), applied
