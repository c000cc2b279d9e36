#!/usr/bin/env python
"""Generate tests/golden/*.npz: outputs of the reference's shader text, executed LITERALLY by
oracle/glsl_exec.py on small seeded planes, in this container (where /root/reference is mounted).

The fixtures travel with the repo (the GPU box has no /root/reference): `pytest -m "not gpu"` checks
the parametric oracle against them, `pytest -m gpu` checks the CUDA kernels against them.

    python tools/make_golden.py            # rewrites every fixture
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mpv_prescalers_b200.synth import batch  # noqa: E402
from oracle.glsl_exec import run_hook  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

CASES = (
    [(f"ravu-lite{ar}-r{r}.hook", None) for ar in ("", "-ar") for r in (2, 3, 4)]
    + [(f"ravu-r{r}{p}.hook", None) for r in (2, 3, 4) for p in ("", "-yuv", "-rgb")]
    + [(f"compute/ravu-3x-r{r}{p}.hook", None) for r in (2, 3, 4) for p in ("", "-rgb")]
    + [("ravu-zoom-r2.hook", (69, 47)), ("ravu-zoom-r3.hook", (84, 60)), ("ravu-zoom-ar-r2.hook", (61, 50)),
       ("ravu-zoom-ar-r2-rgb.hook", (61, 50)), ("ravu-zoom-r2-yuv.hook", (56, 40))]
    + [(f"nnedi3-nns{n}-win8x{s}.hook", None) for n, s in ((16, 4), (16, 6), (32, 4), (64, 6), (256, 6))]
)
H, W = 20, 28


def main():
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, out_size) in enumerate(CASES):
        ch = 3 if ("-yuv" in name or "-rgb" in name) else 1
        x = batch(1, ch, H, W, config=100 + idx)[0]
        img = x[0] if ch == 1 else np.moveaxis(x, 0, -1)
        kw = {"out_size": out_size} if out_size else {}
        out, off, applied = run_hook(os.path.join(REF, name), img, **kw)
        assert applied
        fn = os.path.join(OUT, name.replace("/", "__").replace(".hook", ".npz"))
        np.savez_compressed(fn, hook=name, input=x, output=out.astype(np.float32), offset=np.asarray(off, np.float32),
                            out_size=np.asarray(out_size if out_size else (0, 0)))
        print(f"{name:36s} -> {os.path.basename(fn)} {out.shape} {os.path.getsize(fn)} B")


if __name__ == "__main__":
    main()
