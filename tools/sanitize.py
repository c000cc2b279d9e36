#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpv_prescalers_b200 import _native, prescale, resample

torch.manual_seed(0)
cases = [
    ("ravu-lite-ar-r3.hook", (2, 70, 200), None, {}),          # TMA path (width % 4 == 0)
    ("ravu-lite-ar-r3.hook", (1, 37, 61), None, {}),           # plain staging (odd width)
    ("ravu-lite-r4.hook", (1, 40, 132), None, {}),
    ("ravu-lite-ar-r2.hook", (1, 33, 64), None, {"out_dtype": torch.float16}),
    ("compute/ravu-3x-r3.hook", (1, 30, 68), None, {}),
    ("compute/ravu-3x-r2-rgb.hook", (1, 3, 30, 50), None, {}),
    ("ravu-r4.hook", (1, 70, 90), None, {}),                   # plain staging (row pitch not a multiple of 16 bytes)
    ("ravu-r4.hook", (2, 70, 96), None, {}),                   # TMA staging, one buffer
    ("ravu-r3.hook", (2, 70, 200), None, {}),                  # TMA staging, double buffer
    ("ravu-r3-rgb.hook", (1, 3, 50, 70), None, {}),
    ("ravu-r2-rgb.hook", (2, 3, 50, 72), None, {}),            # TMA staging of the three colour planes
    ("ravu-zoom-r3.hook", (1, 40, 60), (97, 151), {}),         # general path
    ("ravu-zoom-r3.hook", (2, 40, 60), (120, 180), {}),        # exact 3x: key pre-pass + phase kernel (TMA)
    ("ravu-zoom-r2.hook", (1, 41, 61), (82, 122), {}),         # exact 2x: phase kernel, plain staging
    ("ravu-zoom-ar-r2-rgb.hook", (1, 3, 40, 60), (120, 180), {}),
    ("ravu-zoom-ar-r2.hook", (2, 120, 90), (360, 270), {}),    # exact 3x, anti-ringing: phase kernel with exact node classes
    ("ravu-zoom-ar-r2.hook", (1, 50, 70), (100, 140), {}),     # exact 2x, anti-ringing phase kernel
    ("nnedi3-nns256-win8x6.hook", (1, 40, 70), None, {}),
    ("nnedi3-nns32-win8x4.hook", (2, 40, 70), None, {}),
    ("nnedi3-nns64-win8x6.hook", (1, 40, 70), None, {}),
]
for hook, shape, osz, kw in cases:
    x = torch.rand(*shape, device="cuda")
    out = prescale(x, hook, output_size=osz, **kw)
    torch.cuda.synchronize()
    print(hook, tuple(out.shape), "ok")
_native.lib().mpvp_debug_set_grid_limit(2)   # every CTA walks many tiles: buffer hand-over, mbarrier phase flips
for hook, shape, osz in (("ravu-lite-ar-r3.hook", (2, 120, 200), None), ("ravu-r3.hook", (1, 150, 200), None),
                         ("ravu-zoom-r3.hook", (1, 60, 100), (180, 300)), ("nnedi3-nns256-win8x6.hook", (1, 40, 100), None),
                         ("nnedi3-nns16-win8x4.hook", (1, 40, 100), None)):
    out = prescale(torch.rand(*shape, device="cuda"), hook, output_size=osz)
    torch.cuda.synchronize()
    print(hook, "grid limit 2", tuple(out.shape), "ok")
_native.lib().mpvp_debug_set_grid_limit(0)
for shape, osz, kern in (((2, 50, 70), (75, 101), "lanczos"), ((2, 200, 300), (200, 300), "lanczos"),      # edge tiles only / interior tiles
                         ((1, 260, 280), (150, 170), "catmull_rom"), ((1, 130, 140), (260, 280), "bilinear")):
    out = resample(torch.rand(*shape, device="cuda"), osz, (-0.5, -0.5), kern)
    torch.cuda.synchronize()
    print("resample", kern, tuple(out.shape), "ok")
out = resample(torch.randint(0, 256, (2, 200, 300), dtype=torch.uint8, device="cuda"), (200, 300), (-0.5, -0.5), "spline36")
torch.cuda.synchronize()
print("resample uint8", tuple(out.shape), "ok")
u8 = torch.randint(0, 256, (2, 48, 80), dtype=torch.uint8, device="cuda")
for hook in ("ravu-lite-ar-r3.hook", "ravu-r3.hook", "nnedi3-nns32-win8x4.hook"):
    out = prescale(u8, hook)
    torch.cuda.synchronize()
    print(hook, "uint8", tuple(out.shape), "ok")
