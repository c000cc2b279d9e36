#!/usr/bin/env python
"""Train a RAVU-Lite, RAVU-3x or three-pass RAVU LUT on the GPU and write it as a complete hook file (SURVEY.md section 8f
rank 4).

    python tools/train_lut.py --hook ravu-lite-r3.hook --hr planes.npy --out my-ravu-lite-r3.hook
    python tools/train_lut.py --hook compute/ravu-3x-r3.hook --hr planes.npy --out my-ravu-3x-r3.hook
    python tools/train_lut.py --hook ravu-r3.hook --hr planes.npy --out my-ravu-r3.hook --rounds 3
    python tools/train_lut.py --hook ravu-zoom-r2.hook --hr planes.npy --out my-ravu-zoom-r2.hook   (ratios 1.5, 2, 2.5, 3)

``planes.npy``: float32 [F, sH, sW] high-resolution luma planes in [0, 1] (s = 2, or 3 for ravu-3x); the low-resolution
training input is their s x s box average (the usual RAVU training degradation).  Without --hr a synthetic set is used (a smoke run: the result
is a valid filter for synthetic statistics, not a replacement of the shipped weights).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpv_prescalers_b200 import HookFile, find_hook, prescale  # noqa: E402
from mpv_prescalers_b200.synth import batch  # noqa: E402
from mpv_prescalers_b200.train import train_ravu, train_ravu_chain, train_ravu_zoom, write_hook_with_lut  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hook", default="ravu-lite-r3.hook")
    ap.add_argument("--hr", default=None, help=".npy with float32 [F, 2H, 2W] planes")
    ap.add_argument("--out", required=True)
    ap.add_argument("--ridge", type=float, default=1e-6)
    ap.add_argument("--rounds", type=int, default=2, help="fixed-point rounds of the three-pass ravu trainer")
    args = ap.parse_args()
    hk = HookFile.parse(find_hook(args.hook))
    hr = np.load(args.hr).astype(np.float32) if args.hr else batch(8, 1, 720, 960, config=7)[:, 0]
    if hk.variant.family == "ravu-zoom":
        hrz = torch.from_numpy(np.ascontiguousarray(hr)).cuda()
        pairs = []
        for ratio in (1.5, 2.0, 2.5, 3.0):
            size = (int(hrz.shape[1] / ratio), int(hrz.shape[2] / ratio))
            pairs.append((torch.nn.functional.interpolate(hrz[:, None], size=size, mode="area")[:, 0].contiguous(), hrz))

        def mse(hook):
            return float(np.mean([float(((prescale(lr, hook, output_size=tuple(t.shape[1:])) - t) ** 2).mean()) for lr, t in pairs]))

        before = mse(hk)
        lut, count = train_ravu_zoom(hk, pairs, ridge=args.ridge, exclude_clipped=False)
        write_hook_with_lut(hk, lut, args.out)
        print(f"{args.out}: {int((count >= 4 * lut.shape[1]).sum())} of {count.shape[0]} buckets retrained on {int(count.sum())} output "
              f"pixels at 4 scale factors; training MSE {before:.3e} -> {mse(args.out):.3e}")
        return
    s = 3 if hk.variant.family == "ravu-3x" else 2
    hr = torch.from_numpy(np.ascontiguousarray(hr[:, : hr.shape[1] // s * s, : hr.shape[2] // s * s])).cuda()
    lr = torch.nn.functional.avg_pool2d(hr[:, None], s)[:, 0].contiguous()
    if hk.variant.family == "ravu":
        # the hook's output grid carries the source pixel at (2x, 2y): align the targets with it (the box average sits
        # half a high-resolution pixel further right / down, which is the hook's //!OFFSET -0.5 -0.5)
        lr = hr[:, 0::2, 0::2].contiguous()
    before = float(((prescale(lr, hk) - hr) ** 2).mean())
    if hk.variant.family == "ravu":
        lut, count = train_ravu_chain(hk, lr, hr, ridge=args.ridge, exclude_clipped=False, rounds=args.rounds)
    else:
        lut, count = train_ravu(hk, lr, hr, ridge=args.ridge, exclude_clipped=False)
    write_hook_with_lut(hk, lut, args.out)
    after = float(((prescale(lr, args.out) - hr) ** 2).mean())
    print(f"{args.out}: {int((count >= 4 * lut.shape[1] * 2).sum())} of {lut.shape[0]} buckets retrained on {int(count.sum())} windows; "
          f"training MSE {before:.3e} -> {after:.3e}")


if __name__ == "__main__":
    main()
