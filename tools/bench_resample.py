#!/usr/bin/env python
"""Time the offset-correcting scaler (csrc/resample.cu) with CUDA events and report it against the HBM roofline:

    python tools/bench_resample.py [out.json]

Algorithmic bytes = 4 B read per source pixel + 4 B written per output pixel (float32 planes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpv_prescalers_b200 import resample  # noqa: E402

PEAK = 6447.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbps"]["burst"])
except Exception:
    pass

CASES = [  # planes, in (h, w), out (h, w), offset, kernel
    (16, (2160, 3840), (2160, 3840), (-0.5, -0.5), "lanczos"),        # pure offset correction of a ravu / nnedi3 result
    (16, (2160, 3840), (2160, 3840), (-0.5, -0.5), "catmull_rom"),
    (16, (2160, 3840), (2160, 3840), (-0.5, -0.5), "bilinear"),
    (16, (2160, 3840), (1440, 2560), (-0.5, -0.5), "lanczos"),        # 4K hooked plane -> 1440p window
    (16, (1080, 1920), (2160, 3840), (0.0, 0.0), "spline36"),         # plain 2x main scaler
]


def main():
    rows = []
    for n, (h, w), (oh, ow), off, kern in CASES:
        x = torch.rand(n, h, w, device="cuda")
        for _ in range(3):
            y = resample(x, (oh, ow), off, kern)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in evs:
            a.record()
            y = resample(x, (oh, ow), off, kern)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
        gb = 4.0 * n * (h * w + oh * ow) / 1e9
        rows.append({"planes": n, "in": [h, w], "out": [oh, ow], "offset": off, "kernel": kern, "ms": round(ms, 4),
                     "out_gpix_s": round(n * oh * ow / ms / 1e6, 2), "gb_s": round(gb / ms * 1e3, 1), "hbm_frac": round(gb / ms * 1e3 / PEAK, 4)})
        print(rows[-1])
        del x, y
    if len(sys.argv) > 1:
        json.dump({"peak_gb_s": PEAK, "rows": rows}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
