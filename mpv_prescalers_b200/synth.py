"""Deterministic synthetic planes: band-limited mixture that exercises every
angle / strength / coherence bucket, plus pathological planes for the edge-case tests."""
from __future__ import annotations

import numpy as np


def plane(h: int, w: int, seed: int = 0) -> np.ndarray:
    """0.5 + 0.25 sin(0.07x+0.03y) + 0.15 sin(1e-4 (x^2+y^2)) + 0.1 checker32 + N(0, 0.02), clipped."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    ph = 0.37 * (seed % 17)
    p = (
        0.5
        + 0.25 * np.sin(0.07 * x + 0.03 * y + ph)
        + 0.15 * np.sin(1e-4 * (x * x + y * y) + 2.0 * ph)
        + 0.1 * ((((x // 32) + (y // 32)) % 2) - 0.5)
        + rng.normal(0.0, 0.02, (h, w))
    )
    return np.clip(p, 0.0, 1.0).astype(np.float32)


def batch(n: int, c: int, h: int, w: int, config: int = 0) -> np.ndarray:
    """[n, c, h, w] float32; seed = 1000 * config + frame (+ 131 * channel)."""
    out = np.empty((n, c, h, w), np.float32)
    for f in range(n):
        for ch in range(c):
            out[f, ch] = plane(h, w, 1000 * config + f + 131 * ch)
    return out


def natural(h: int, w: int, seed: int = 0, c: int = 1) -> np.ndarray:
    """Natural-image statistics: 1/f amplitude spectrum with random phases, scaled to mean 0.5 and
    standard deviation 0.18, clipped to [0, 1].  ``[h, w]`` (c == 1) or ``[c, h, w]`` with correlated channels."""
    rng = np.random.default_rng(seed)
    fy, fx = np.fft.fftfreq(h)[:, None], np.fft.rfftfreq(w)[None, :]
    amp = 1.0 / np.maximum(np.sqrt(fx * fx + fy * fy), 1.0 / max(h, w))
    amp[0, 0] = 0.0

    def field():
        spec = amp * np.exp(2j * np.pi * rng.random(amp.shape))
        f = np.fft.irfft2(spec, s=(h, w))
        return f / f.std()

    base = field()
    out = np.stack([np.clip(0.5 + 0.18 * (base if k == 0 else 0.8 * base + 0.6 * field()), 0.0, 1.0) for k in range(c)])
    return out[0].astype(np.float32) if c == 1 else out.astype(np.float32)


def pathological(h: int, w: int) -> dict:
    y, x = np.mgrid[0:h, 0:w]
    return {
        "zeros": np.zeros((h, w), np.float32),
        "ones": np.ones((h, w), np.float32),
        "const": np.full((h, w), 0.3125, np.float32),
        "step": (x >= w // 2).astype(np.float32) * 0.6 + 0.2,
        "checker1": ((x + y) % 2).astype(np.float32),
        "ramp": (x / max(w - 1, 1)).astype(np.float32),
    }


def torch_batch(n: int, c: int, h: int, w: int, device, seed: int = 0):
    """Fast on-device synthetic batch for the bench (same recipe, torch RNG)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    y = torch.arange(h, device=device, dtype=torch.float32)[:, None]
    x = torch.arange(w, device=device, dtype=torch.float32)[None, :]
    base = 0.5 + 0.25 * torch.sin(0.07 * x + 0.03 * y) + 0.15 * torch.sin(1e-4 * (x * x + y * y))
    base = base + 0.1 * ((((x // 32) + (y // 32)) % 2) - 0.5)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
    for f in range(n):
        ph = 0.37 * (f % 17)
        fr = base * (0.9 + 0.01 * (f % 7)) + 0.02 * torch.sin(0.11 * x + ph) * torch.cos(0.05 * y - ph)
        out[f] = fr[None] + 0.02 * torch.randn((c, h, w), generator=g, device=device)
    return out.clamp_(0.0, 1.0)
