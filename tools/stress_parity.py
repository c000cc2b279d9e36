#!/usr/bin/env python
"""Randomised parity sweep (not part of the test suite: minutes of GPU + oracle time): many seeded plane sizes, batch sizes
and ratios per family, each compared with the oracle by the same helpers the GPU tests use.

    python tools/stress_parity.py [cases_per_family] [seed] [max_h max_w]
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_parity as T  # noqa: E402
from tests.parity import check_output  # noqa: E402
from tests.conftest import hook_path  # noqa: E402

RAVU = ["ravu-lite-r2.hook", "ravu-lite-r3.hook", "ravu-lite-r4.hook", "ravu-lite-ar-r2.hook", "ravu-lite-ar-r3.hook", "ravu-lite-ar-r4.hook",
        "ravu-r2.hook", "ravu-r3.hook", "ravu-r4.hook", "ravu-r2-rgb.hook", "ravu-r3-yuv.hook", "ravu-r4-rgb.hook",
        "compute/ravu-3x-r2.hook", "compute/ravu-3x-r3.hook", "compute/ravu-3x-r4.hook", "compute/ravu-3x-r2-rgb.hook",
        "ravu-zoom-r2.hook", "ravu-zoom-r3.hook", "ravu-zoom-ar-r2.hook", "ravu-zoom-r2-yuv.hook", "ravu-zoom-ar-r2-rgb.hook"]
NN = ["nnedi3-nns16-win8x4.hook", "nnedi3-nns32-win8x6.hook", "nnedi3-nns64-win8x4.hook", "nnedi3-nns128-win8x6.hook", "nnedi3-nns256-win8x4.hook"]


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    hmax, wmax = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (200, 330)
    rng = np.random.default_rng(seed)
    fails, runs, t0 = [], 0, time.time()
    for name in RAVU:
        for h, w in T._random_sizes(int(rng.integers(1 << 30)), per, hmax, wmax):
            n = int(rng.integers(1, 4))
            out_hw = None
            if "zoom" in name:
                ratios = [(2.0, 2.0), (3.0, 3.0), (1.5, 1.5), (rng.uniform(1.0, 3.5), rng.uniform(1.0, 3.5))]
                ry, rx = ratios[int(rng.integers(len(ratios)))]
                out_hw = (max(h + 1, int(round(h * ry))), max(w + 1, int(round(w * rx))))     # the hook only fires when both axes grow
            in_bits = None
            if "zoom" not in name and rng.random() < 0.3:
                in_bits = int(rng.choice([8, 10, 16]))           # UNORM integer planes in, float32 out
            cap = int(rng.integers(1, 4)) if rng.random() < 0.3 else 0   # persistent grid capped to 1..3 CTAs: many tiles per CTA
            try:
              with T.grid_limit(cap):
                # cascade_tol: the end-to-end count of differing ravu keys (last-bit int11 differences amplified by
                # ill-conditioned keys of passes 2 / 3) is a property of the algorithm, reported but not a parity failure;
                # the comparison on identical inputs inside the helper keeps its 99.99 % / boundary-only rule
                T._run_ravu_variant(name, n=n, h=h, w=w, config=int(rng.integers(100, 900)), out_hw=out_hw, cascade_tol=2e-3,
                                    in_bits=in_bits)
            except Exception as e:  # keep going: the point is the list of failures
                fails.append((name, n, h, w, out_hw, in_bits, cap, f"{type(e).__name__}: {str(e)[:200]}"))
            runs += 1
    from mpv_prescalers_b200 import HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from oracle import nnedi3_np

    for name in NN:
        hk = HookFile.parse(hook_path(name))
        for h, w in T._random_sizes(int(rng.integers(1 << 30)), max(2, per // 2), max(90, hmax // 2), max(220, wmax // 2)):
            n = int(rng.integers(1, 3))
            x = batch(n, 1, h, w, config=int(rng.integers(100, 900)))
            cap = int(rng.integers(1, 4)) if rng.random() < 0.3 else 0
            try:
                with T.grid_limit(cap):
                    out = prescale(torch.from_numpy(x).cuda(), hk).cpu().numpy()
                for f in range(n):
                    ref, _ = nnedi3_np.nnedi3(x[f, 0], hk.variant)
                    check_output(out[f, 0], ref, None, f"{name} {h}x{w}")
            except Exception as e:
                fails.append((name, n, h, w, None, f"{type(e).__name__}: {str(e)[:200]}"))
            runs += 1
    print(f"{runs} cases, {len(fails)} failures, {time.time() - t0:.0f} s")
    for f in fails:
        print("FAIL", f)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
