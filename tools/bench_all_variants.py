#!/usr/bin/env python
"""Time every shipped variant family member once (CUDA events, device-resident float32 planes) and write a table:

    python tools/bench_all_variants.py gpurun_out/r01_variants.json

2x / 3x hooks: 8 frames of 1920x1080 (3-channel: 4); ravu-zoom: 8 frames 1280x720 -> 3840x2160; NNEDI3: 2 frames of
1920x1080.  Roofline fraction as in bench.py (HBM for RAVU, fp16 tensor peak for NNEDI3)."""
import json, os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpv_prescalers_b200 import HookFile, find_hook, prescale
from mpv_prescalers_b200.synth import torch_batch

HOOKS = (
    [f"ravu-lite{ar}-r{r}.hook" for ar in ("", "-ar") for r in (2, 3, 4)]
    + [f"ravu-r{r}{p}.hook" for r in (2, 3, 4) for p in ("", "-yuv", "-rgb")]
    + [f"compute/ravu-3x-r{r}{p}.hook" for r in (2, 3, 4) for p in ("", "-yuv", "-rgb")]
    + ["ravu-zoom-r2.hook", "ravu-zoom-r3.hook", "ravu-zoom-ar-r2.hook", "ravu-zoom-ar-r2-rgb.hook", "ravu-zoom-r2-yuv.hook"]
    + [f"nnedi3-nns{n}-win8x{s}.hook" for n in (16, 32, 64, 128, 256) for s in (4, 6)]
)

def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/variants.json"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tc = float(peaks.get("hbm_gbs", 6650.0)), float(peaks.get("bf16_tflops", 1590.0))
    rows = []
    for name in HOOKS:
        try:
            hk = HookFile.parse(find_hook(name))
        except Exception as e:  # hook not shipped to this box
            rows.append({"hook": name, "error": str(e)[:80]})
            continue
        v = hk.variant
        zoom, nn = v.family == "ravu-zoom", v.family == "nnedi3"
        h, w = (720, 1280) if zoom else (1080, 1920)
        n = 2 if nn else (4 if v.channels == 3 else 8)
        x = torch_batch(n, v.channels, h, w, "cuda", seed=5)
        osz = (2160, 3840) if zoom else None
        for _ in range(2):
            o = prescale(x, hk, output_size=osz)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            o = prescale(x, hk, output_size=osz)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        out_px = o.numel() // v.channels
        nbytes = 4.0 * (x.numel() + o.numel())
        row = {"hook": name, "frames": n, "in": [h, w], "out": list(o.shape[-2:]), "ms": round(ms, 4),
               "out_gpix_s": round(out_px / ms / 1e6, 2), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / hbm, 4)}
        if nn:
            flops = 2.0 * (8 * v.win[1]) * (2 * v.nns) * 3.0 * h * w * n
            row["tensor_frac"] = round(flops / (ms * 1e-3) / 1e12 / tc, 4)
        rows.append(row)
        print(row, flush=True)
        del x, o
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(rows, open(out_path, "w"), indent=1)

if __name__ == "__main__":
    main()
