#!/usr/bin/env python
"""Build an experiment variant of libmpvp.so with extra -D flags (A/B runs on the GPU box):

    python tools/build_variant.py x1 ravu_lite.cu -DMPVP_X_FOO=1     -> build/libmpvp_x1.so
    MPVP_LIB=build/libmpvp_x1.so python tools/sweep_lite.py ...

Only the named translation units are recompiled; the other objects come from the product build.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpv_prescalers_b200 import _native

def main():
    tag = sys.argv[1]
    units = [a for a in sys.argv[2:] if a.endswith(".cu")]
    flags = [a for a in sys.argv[2:] if not a.endswith(".cu")]
    objdir = os.path.join(ROOT, "build", "obj")
    cflags = [f for f in _native.NVCC_FLAGS if f != "-shared"]
    objs = []
    for src in _native.sources():
        base = os.path.basename(src)
        obj = os.path.join(objdir, base + ".o")
        if base in units:
            obj = os.path.join(objdir, f"{base}.{tag}.o")
            subprocess.run(["nvcc"] + cflags + flags + ["-c", "-o", obj, src], cwd=_native.CSRC, check=True)
        objs.append(obj)
    out = os.path.join(ROOT, "build", f"libmpvp_{tag}.so")
    subprocess.run(["nvcc", "-shared", "-o", out] + objs + ["-lcuda"], check=True)
    print(out)

if __name__ == "__main__":
    main()
