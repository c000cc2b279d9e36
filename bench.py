#!/usr/bin/env python
"""Benchmark of the mpv-prescalers hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic frames.  Default workload is
BASELINE.json configs[1]: ravu-lite-ar-r3 1080p -> 2160p luma, 64-frame batch per GPU.  Prints ONE JSON
line (rank 0).  ``--impl reference`` times the CPU restatement of the reference (oracle/) on the host
cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (hook, channels, h, w, frames per GPU, output factor or (oh, ow), BASELINE.json config index)
    "ravu-lite-ar-r3": ("ravu-lite-ar-r3.hook", 1, 1080, 1920, 64, 2, 1),
    "ravu-lite-r3-540p": ("ravu-lite-r3.hook", 1, 540, 960, 1, 2, 0),
    "ravu-r4": ("ravu-r4.hook", 1, 1080, 1920, 64, 2, 2),
    "ravu-r3-rgb": ("compute/ravu-r3-rgb.hook", 3, 1080, 1920, 16, 2, 2),
    "ravu-zoom-r3": ("ravu-zoom-r3.hook", 1, 720, 1280, 8, (2160, 3840), 3),
    "ravu-zoom-ar-r2": ("ravu-zoom-ar-r2.hook", 1, 720, 1280, 8, (2160, 3840), 3),
    "ravu-3x-r3": ("compute/ravu-3x-r3.hook", 1, 720, 1280, 16, 3, 3),
    "nnedi3-nns256-win8x6": ("nnedi3-nns256-win8x6.hook", 1, 2160, 3840, 2, 2, 4),
    "nnedi3-nns32-win8x4": ("nnedi3-nns32-win8x4.hook", 1, 1080, 1920, 8, 2, 4),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), sampled in-process through
    NVML every few milliseconds (a fresh `nvidia-smi` per sample is too slow for a 40 ms region)."""

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.th = index, [], threading.Event(), None
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(reasons)))
            except Exception:
                pass
            self.stop_flag.wait(0.002)

    def start(self):
        if self.nv is None:
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        nv = self.nv
        bits = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        sm = sorted(s[0] for s in self.samples)
        seen = 0
        for _, r in self.samples:
            seen |= r
        reasons = sorted(k for k, b in bits.items() if seen & b)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


IO_MODES = {  # name: (input bytes / sample, output bytes / sample)
    "f32": (4, 4), "u8": (1, 1), "u10": (2, 2), "f16out": (4, 2),
}


def algorithmic_work(wl, n_frames, io="f32"):
    """(output Mpix, algorithmic bytes, algorithmic flops) of one step on one GPU."""
    hook, c, h, w, _, fac, _ = WORKLOADS[wl]
    oh, ow = (h * fac, w * fac) if isinstance(fac, int) else fac
    out_px = n_frames * oh * ow
    bi, bo = IO_MODES[io]
    nbytes = 1.0 * c * n_frames * (bi * h * w + bo * oh * ow)
    flops = 0.0
    if hook.startswith("nnedi3"):
        import re

        m = re.search(r"nns(\d+)-win8x(\d)", hook)
        nns, s = int(m.group(1)), int(m.group(2))
        flops = 2.0 * (8 * s) * (2 * nns) * (3.0 * h * w) * n_frames
    return out_px / 1e6, nbytes, flops


def cpu_frames(wl, count, seed0=0):
    import numpy as np
    from mpv_prescalers_b200.synth import batch

    hook, c, h, w, _, _, cfg = WORKLOADS[wl]
    x = batch(count, c, h, w, config=cfg + 10 * seed0)
    return x


def _cpu_one(args):
    """Worker: run the oracle on one frame (executed in a separate process for --impl reference)."""
    wl, frame = args
    import numpy as np

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from mpv_prescalers_b200 import HookFile, find_hook
    from oracle import nnedi3_np, ravu_np

    hook, c, h, w, _, fac, _ = WORKLOADS[wl]
    hk = HookFile.parse(find_hook(hook))
    v = hk.variant
    img = frame[0] if c == 1 else np.moveaxis(frame, 0, -1)
    t0 = time.perf_counter()
    if v.family == "nnedi3":
        out, _ = nnedi3_np.nnedi3(img, v)
    else:
        osz = None if isinstance(fac, int) else (fac[1], fac[0])
        out = ravu_np.run(img, v, osz).out
    return time.perf_counter() - t0, int(out.shape[0] * out.shape[1])


def cpu_baseline(wl, budget_s=20.0, procs=1):
    """Time the oracle (a port of the reference's shader math) on a bounded sample of the workload, one process."""
    hook, c, h, w, _, fac, _ = WORKLOADS[wl]
    # nnedi3 at 2160p is minutes per frame on a CPU: use a crop of the same planes
    crop = (270, 480) if hook.startswith("nnedi3") else (h, w)
    frames = cpu_frames(wl, 1)[:, :, : crop[0], : crop[1]]
    t0 = time.perf_counter()
    done_px, used = 0, 0
    while True:
        dt, px = _cpu_one((wl, frames[used % len(frames)]))
        done_px += px
        used += 1
        if time.perf_counter() - t0 > budget_s * 0.5 or used >= 4:
            break
    wall = time.perf_counter() - t0
    return {
        "value": done_px / 1e6 / wall,
        "unit": "Mpix/s",
        "cores": 1,
        "kind": "port",
        "sample": f"{used} frame(s) of {crop[1]}x{crop[0]} from the same synthetic workload, NumPy fp32 oracle, 1 process",
    }


def config_of(wl, nf, world, io):
    hook, c, h, w, _, fac, cfg_idx = WORKLOADS[wl]
    oh, ow = (h * fac, w * fac) if isinstance(fac, int) else fac
    big = 1.0 * c * nf * (IO_MODES[io][0] * h * w + IO_MODES[io][1] * oh * ow) > 2.5e8
    return {
        "workload": f"{hook} {w}x{h}->{ow}x{oh} {'luma' if c == 1 else '3ch'} {'fp32' if io == 'f32' else 'planes ' + io}, {nf} frames per GPU (BASELINE.json configs[{cfg_idx}])",
        "frames_per_gpu": nf,
        "parallelism": f"frame-sharded x{world}, no collective",
        "l2": "working set per step exceeds the 126 MB L2" if big else "L2 flushed between steps",
        "io": io,
    }


def reference_arm(args, wl, nf, world):
    """--impl reference: the reference's CPU implementation of the path = the NumPy restatement of its shader math
    (oracle/; the GLSL itself needs mpv + GL/Vulkan, which this image lacks), on all host cores.  A step = one bounded
    sample of the workload: every process runs the oracle on one crop of a synthetic frame of the workload."""
    import multiprocessing as mp

    hook, c, h, w, _, fac, _ = WORKLOADS[wl]
    procs = os.cpu_count() or 1
    # sized so that W warm-up + K timed steps end within a few minutes: half-height crops (quarter for nnedi3)
    crop = (270, 480) if hook.startswith("nnedi3") else (max(h // 2, 1), w)
    frames = cpu_frames(wl, procs)[:, :, : crop[0], : crop[1]]
    work = [(wl, frames[i]) for i in range(procs)]
    warm = max(args.warmup, 1)
    per_step = []
    with mp.get_context("spawn").Pool(procs) as pool:
        for _ in range(warm):
            pool.map(_cpu_one, work)
        for _ in range(max(args.steps, 1)):
            t0 = time.perf_counter()
            res = pool.map(_cpu_one, work)
            per_step.append((time.perf_counter() - t0, sum(r[1] for r in res)))
    wall = sum(t for t, _ in per_step)
    val = sum(px for _, px in per_step) / 1e6 / wall
    sample = (f"per step: {procs} crop(s) of {crop[1]}x{crop[0]} from the synthetic frames of the workload, NumPy fp32 oracle, "
              f"{procs} processes (one per host core)")
    return {
        "impl": "reference", "metric": "output Mpix/s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": len(per_step),
        "warmup": args.warmup, "ms_per_step": wall / len(per_step) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(wl, nf, world, args.io),
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = NumPy restatement of the reference's GLSL (oracle/); the GLSL itself needs mpv + GL/Vulkan, which this image lacks",
    }


class Runner:
    """One workload on this rank's GPU: device-timed steps, roofline, end-to-end through prescale()."""

    def __init__(self, wl, nf, io, rank, world, local_rank, dist):
        import torch

        from mpv_prescalers_b200 import HookFile, find_hook
        from mpv_prescalers_b200.api import PlaneIO, plan, upload_weights
        from mpv_prescalers_b200.synth import torch_batch

        self.torch, self.wl, self.nf, self.io, self.rank, self.world, self.dist = torch, wl, nf, io, rank, world, dist
        hook, c, h, w, _, fac, cfg_idx = WORKLOADS[wl]
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        self.hk = HookFile.parse(find_hook(hook))
        self.out_size = None if isinstance(fac, int) else fac
        self.oh, self.ow = (h * fac, w * fac) if isinstance(fac, int) else fac
        self.c = c
        self.pl = plan(self.hk, (h, w), self.out_size)
        self.W = upload_weights(self.hk, local_rank)
        x = torch_batch(nf, c, h, w, self.dev, seed=1000 * cfg_idx + rank)
        self.io_kw = {}
        if io == "u8":
            x = torch.round(x.clamp(0, 1) * 255.0).to(torch.uint8)
        elif io == "u10":
            x = torch.round(x.clamp(0, 1) * 1023.0).to(torch.int32).to(torch.uint16)
            self.io_kw = dict(bit_depth=10)
        elif io == "f16out":
            self.io_kw = dict(out_dtype=torch.float16)
        self.x = x
        self.pio = PlaneIO(x.dtype, self.io_kw.get("out_dtype"), self.io_kw.get("bit_depth"))
        self.config = config_of(wl, nf, world, io)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """(ms_per_step max over ranks, launches, clocks, wall seconds of the timed region)"""
        torch = self.torch
        from mpv_prescalers_b200 import _native
        from mpv_prescalers_b200.api import _launch
        from mpv_prescalers_b200.sharding import max_over_ranks

        lib = _native.lib()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev) if "flushed" in self.config["l2"] else None

        def step():
            out, _ = _launch(self.hk, self.pl, self.x, self.W, False, self.pio)
            return out

        for _ in range(warmup):
            out = step()
            if flush is not None:
                flush.zero_()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
        launches0 = lib.mpvp_launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        t_wall0 = time.perf_counter()
        for i in range(steps):
            if flush is not None:
                flush.zero_()
            evs[i][0].record()
            out = step()
            evs[i][1].record()
        self.barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = lib.mpvp_launch_count() - launches0
        clocks = sampler.stop() if self.rank == 0 else None
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs), self.dev)
        del out, flush
        return total_ms / steps, int(launches), clocks, t_wall

    def roofline(self, ms_per_step, launches_per_step):
        mpix, abytes, aflops = algorithmic_work(self.wl, self.nf, self.io)
        hbm_peak, tc_peak, src = peaks()
        if aflops > 0:
            ach = aflops / (ms_per_step / 1e3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "traffic": None,
                    "peak_source": src, "hbm_frac": abytes / (ms_per_step / 1e3) / 1e9 / hbm_peak}
        else:
            ach = abytes / (ms_per_step / 1e3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_source": src}
        roof["kernel"] = f"{self.pl.family} fused kernel(s), {launches_per_step} launch(es) per step, avg {ms_per_step / launches_per_step:.4f} ms"
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr) and self.io == "f32":
            try:
                with open(tr) as f:
                    roof["traffic"] = json.load(f).get(self.wl)
            except Exception:
                pass
        return roof

    def e2e(self, steps, x=None, io_kw=None, copy_floor=True):
        """The same step through the public prescale() with pinned HOST buffers: H2D, kernel(s), D2H inside the timed
        region.  copy_floor_ms = the bare pinned H2D + D2H copies of the same bytes on two streams (no kernel): what
        the box's PCIe path allows."""
        torch = self.torch
        from mpv_prescalers_b200 import prescale
        from mpv_prescalers_b200.sharding import max_over_ranks

        x = self.x if x is None else x
        io_kw = self.io_kw if io_kw is None else io_kw
        mpix = self.nf * self.oh * self.ow / 1e6
        xh = x.cpu().pin_memory()
        out_dtype = io_kw.get("out_dtype", x.dtype)
        oh_pinned = torch.empty((self.nf, self.c, self.oh, self.ow), dtype=out_dtype, pin_memory=True)
        o = prescale(xh, self.hk, self.out_size, out=oh_pinned, **io_kw)  # warm-up (allocator pools, streams)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            o = prescale(xh, self.hk, self.out_size, out=oh_pinned, **io_kw)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, self.dev)
        res = {"value": mpix * self.world * steps / dt, "unit": "Mpix/s", "h2d_bytes_per_step": int(xh.numel() * xh.element_size()),
               "d2h_bytes_per_step": int(o.numel() * o.element_size()), "steps": steps, "ms_per_step": dt / steps * 1e3}
        if copy_floor:
            dx = torch.empty_like(x)
            do = torch.empty((self.nf, self.c, self.oh, self.ow), dtype=out_dtype, device=self.dev)
            s1, s2 = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
            for rep in range(2):   # first pass warms up, second is timed
                self.barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s1):
                    dx.copy_(xh, non_blocking=True)
                with torch.cuda.stream(s2):
                    oh_pinned.copy_(do, non_blocking=True)
                s1.synchronize()
                s2.synchronize()
                fl = time.perf_counter() - t0
            res["copy_floor_ms"] = max_over_ranks(fl, self.dev) * 1e3
            del dx, do
        del o, xh, oh_pinned
        return res


def bind_to_gpu_cpus(torch, local_rank):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU BEFORE any pinned host buffer is
    allocated (first touch places the pages on that NUMA node), so the end-to-end copies of eight ranks do not all cross
    the socket interconnect.  Returns the number of CPUs bound to, or None when the GPU is local to every CPU the process
    may use (single-socket boxes, the 1-GPU lease) or NVML is not available."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus_id = f"{getattr(pr, 'pci_domain_id', 0):08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = (max(os.sched_getaffinity(0)) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
        cur = os.sched_getaffinity(0)
        local &= cur
        if local and local != cur:
            os.sched_setaffinity(0, local)
            return len(local)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ravu-lite-ar-r3", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the workload's)")
    ap.add_argument("--io", default="f32", choices=["f32", "u8", "u10", "f16out"],
                    help="plane formats either side of the path: f32 (BASELINE.json's), u8 / u10 = UNORM video planes in "
                         "and out (uint8, 10 bits in uint16), f16out = float32 in, binary16 out (SURVEY.md 8f rank 1)")
    ap.add_argument("--secondary", default="auto",
                    help="comma-separated workloads timed in the same process after the headline one and reported under "
                         "'secondary'; 'auto' = the other kernels BASELINE.json's metric and configs name (default workload "
                         "only), 'none' = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    nf = args.frames or WORKLOADS[wl][4]

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(args, wl, nf, world)))
        return 0

    import torch

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    host_cpus = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("MPVP_BENCH_BIND", "1") != "0":        # A/B switch for the NUMA binding
            host_cpus = bind_to_gpu_cpus(torch, local_rank)
        dist.init_process_group("nccl", device_id=dev)

    warm = max(args.warmup, 3)
    run = Runner(wl, nf, args.io, rank, world, local_rank, dist)
    ms_per_step, launches, clocks, t_wall = run.timed(args.steps, warm)
    mpix, _, _ = algorithmic_work(wl, nf, args.io)
    value = mpix * world / (ms_per_step / 1e3)
    roof = run.roofline(ms_per_step, max(1, launches // max(args.steps, 1)))

    e2e = e2e_u8 = None
    if not args.no_e2e:
        e2e = run.e2e(max(1, min(args.steps, 3)))
        if args.io == "f32" and wl == "ravu-lite-ar-r3":
            # the same frames as 8-bit video planes in and out (the wire format either side of the path)
            x8 = torch.round(run.x.clamp(0, 1) * 255.0).to(torch.uint8)
            e2e_u8 = run.e2e(max(1, min(args.steps, 3)), x=x8, io_kw={})
            del x8
    del run
    torch.cuda.empty_cache()

    # ---- the other kernels of BASELINE.json's metric / configs, in the same process ---------------------------------
    secondary = []
    names = []
    if args.secondary == "auto":
        if wl == "ravu-lite-ar-r3" and args.io == "f32" and not args.frames:
            names = ["nnedi3-nns256-win8x6", "ravu-r4", "ravu-r3-rgb", "ravu-zoom-r3", "ravu-zoom-ar-r2", "ravu-3x-r3"]
    elif args.secondary != "none":
        names = [n for n in args.secondary.split(",") if n]
    for name in names:
        entry = {"workload": name}
        try:
            r2 = Runner(name, WORKLOADS[name][4], "f32", rank, world, local_rank, dist)
            steps2 = max(5, min(args.steps, 20))
            ms2, l2, ck2, _ = r2.timed(steps2, warm)
            mp2, _, _ = algorithmic_work(name, r2.nf, "f32")
            entry.update({"config": r2.config, "value": mp2 * world / (ms2 / 1e3), "unit": "Mpix/s", "ms_per_step": ms2, "steps": steps2,
                          "warmup": warm, "roofline": r2.roofline(ms2, max(1, l2 // steps2)), "gpu_launches": l2, "clocks": ck2})
            if not args.no_e2e and name == "nnedi3-nns256-win8x6":
                entry["e2e"] = r2.e2e(2, copy_floor=False)
            del r2
            torch.cuda.empty_cache()
        except Exception as e:  # a secondary workload must never take the headline line down with it
            entry["error"] = f"{type(e).__name__}: {e}"[:300]
        secondary.append(entry)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(wl, nf, world, args.io), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if e2e is not None and host_cpus is not None:
            e2e["host_cpus_per_rank"] = host_cpus        # ranks are bound to the CPUs local to their GPU (NUMA)
        if e2e_u8 is not None:
            line["e2e_u8_planes"] = e2e_u8
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
