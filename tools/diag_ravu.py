#!/usr/bin/env python
"""GPU diagnostic: where do CUDA and oracle RAVU keys / outputs differ?  (writes to stdout)"""
import sys, os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpv_prescalers_b200 import HookFile, find_hook, prescale
from mpv_prescalers_b200.synth import batch
from oracle import ravu_np

np.set_printoptions(linewidth=200, precision=6)


def run(name, h, w, config, n=1):
    hk = HookFile.parse(find_hook(name))
    v = hk.variant
    x = batch(n, v.channels, h, w, config=config)
    xt = torch.from_numpy(x).cuda()
    if v.channels == 1:
        xt = xt[:, 0]
    out, bk = prescale(xt, hk, return_buckets=True)
    out, bk = out.cpu().numpy(), bk.cpu().numpy()
    for f in range(n):
        img = x[f, 0] if v.channels == 1 else np.moveaxis(x[f], 0, -1)
        ref = ravu_np.run(img, v)
        got = out[f] if v.channels == 1 else np.moveaxis(out[f], 0, -1)
        d = np.abs(got - ref.out)
        print(f"== {name} {h}x{w} frame {f}: max abs {d.max():.3e}, px > 1e-3: {(d > 1e-3).sum()}")
        keys = bk[f] if bk[f].ndim == 3 else bk[f][None]
        for k, key in enumerate(ref.keys):
            mism = np.argwhere(keys[k] != key.row)
            print(f"   key {k}: {len(mism)} mismatches of {key.row.size}")
            for (yy, xx) in mism[:12]:
                print(f"      (y={yy},x={xx}) cuda row {keys[k][yy, xx]} oracle row {key.row[yy, xx]}  angle_f {key.angle_f[yy, xx]:.7f} lam {key.lam[yy, xx]:.7e} "
                      f"log2 {np.log2(key.lam[yy, xx] * 2000 + 1.19e-7) if v.strength_log2_scale else 0:.6f} mu {key.mu[yy, xx]:.7f}")
        bad = np.argwhere(d > 1e-3)
        for p in bad[:10]:
            print("      out diff at", tuple(p), float(got[tuple(p)]), float(ref.out[tuple(p)]))


if __name__ == "__main__":
    run("ravu-r3.hook", 9, 1, 17)
    run("ravu-r3.hook", 101, 157, 11, n=2)
    run("ravu-r4.hook", 101, 157, 11, n=2)
    run("ravu-r2.hook", 101, 157, 11, n=2)
    run("ravu-r3-rgb.hook", 101, 157, 11, n=1)
    run("ravu-lite-ar-r3.hook", 540, 960, 3, n=1)
    run("ravu-r3.hook", 540, 960, 3, n=1)
