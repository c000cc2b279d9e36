#!/usr/bin/env python
"""Time one hook on a 16-frame 1080p batch (CUDA events); used for A/B sweeps of kernel tuning knobs
selected through environment variables (each setting needs its own process)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpv_prescalers_b200 import HookFile, find_hook, prescale
from mpv_prescalers_b200.synth import torch_batch

def main():
    hook = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    hk = HookFile.parse(find_hook(hook))
    v = hk.variant
    x = torch_batch(n, v.channels, 1080, 1920, "cuda", seed=3)
    for _ in range(3):
        prescale(x, hk)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        prescale(x, hk)
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"hook": hook, "frames": n, "ms": a.elapsed_time(b) / 10, "env": {k: v for k, v in os.environ.items() if k.startswith("MPVP_")}}))

if __name__ == "__main__":
    main()
