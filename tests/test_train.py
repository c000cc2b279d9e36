"""LUT trainer: host logic on the CPU, self-consistency on the GPU."""
import numpy as np
import pytest

from tests.conftest import hook_path


def test_lut_hex_round_trip_and_hook_rewrite(tmp_path):
    """A rewritten hook file keeps the GLSL byte for byte (so the fused kernels accept it) and carries the new payload."""
    pytest.importorskip("torch")
    from mpv_prescalers_b200 import HookFile
    from mpv_prescalers_b200.train import lut_to_hex, write_hook_with_lut

    hk = HookFile.parse(hook_path("ravu-lite-r2.hook"))
    lut = np.asarray(hk.variant.lut.data, np.float32)
    assert lut_to_hex(lut) in open(hk.path).read()          # the shipped payload is reproduced bit for bit
    new = (lut * np.float32(0.5)).astype(np.float32)
    out = tmp_path / "half.hook"
    write_hook_with_lut(hk, new, str(out))
    hk2 = HookFile.parse(str(out))
    assert np.array_equal(hk2.variant.lut.data, new)
    assert [p.body for p in hk2.passes] == [p.body for p in hk.passes]
    assert hk2.content_key != hk.content_key


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ravu-lite-r3.hook", "ravu-lite-r2.hook", "compute/ravu-3x-r2.hook", "compute/ravu-3x-r3.hook"])
def test_trainer_recovers_the_lut_that_made_the_targets(name, tmp_path):
    """Self-consistency: high-resolution targets produced by a shipped LUT are a per-bucket LINEAR function of the source
    windows, so least squares must give that LUT back (well-populated buckets), and the retrained hook file must
    reproduce the original's output on a plane it has not seen."""
    import torch

    from mpv_prescalers_b200 import HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from mpv_prescalers_b200.train import train_ravu, write_hook_with_lut
    from tests.parity import psnr

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    hk = HookFile.parse(hook_path(name))
    rng = np.random.default_rng(5)
    x = batch(6, 1, 360, 480, config=91)[:, 0]
    x = np.clip(0.2 + 0.6 * x + rng.normal(0, 0.04, x.shape), 0.02, 0.98).astype(np.float32)   # full-rank windows, few clipped targets
    lr = torch.from_numpy(x).cuda()
    hr = prescale(lr, hk)
    lut, count = train_ravu(hk, lr, hr)
    ref = np.asarray(hk.variant.lut.data, np.float32).astype(np.float16).astype(np.float32)      # what the kernels apply
    well = count >= 5000
    assert well.sum() >= 20, f"only {well.sum()} well-populated buckets"
    assert np.abs(lut[well] - ref[well]).max() <= 2e-3, f"recovered LUT differs by {np.abs(lut[well] - ref[well]).max():.2e}"
    starved = count < 4 * (2 * hk.variant.radius - 1) ** 2          # buckets that saw too few windows keep the hook's own row
    assert np.array_equal(lut[starved], np.asarray(hk.variant.lut.data, np.float32)[starved])
    out = tmp_path / "retrained.hook"
    write_hook_with_lut(hk, lut, str(out))
    fresh = torch.from_numpy(np.clip(0.2 + 0.6 * batch(1, 1, 200, 300, config=92)[:, 0], 0, 1).astype(np.float32)).cuda()
    a, b = prescale(fresh, hk), prescale(fresh, str(out))
    assert psnr(a.cpu().numpy(), b.cpu().numpy()) >= 60.0


@pytest.mark.parametrize("r", [2, 3, 4])
def test_chain_lattice_reproduces_the_three_passes(r):
    """The trainer's own tap geometry of ravu's passes (source lattice, then the 45-degree lattice of source pixels and pass-1
    results, clamp-to-edge per texture), applied with the shipped LUT to the oracle's output plane, gives the oracle's
    int11 / int10 / int01 planes back bit for bit."""
    torch = pytest.importorskip("torch")
    from mpv_prescalers_b200 import HookFile
    from mpv_prescalers_b200.synth import batch
    from mpv_prescalers_b200.train import _chain_taps, _gather_lattice
    from oracle import ravu_np

    v = HookFile.parse(hook_path(f"ravu-r{r}.hook")).variant
    x = batch(1, 1, 18, 23, config=93)[0, 0]
    res, inter = ravu_np.ravu(x, v, return_intermediates=True)
    out = torch.from_numpy(res.out)
    N = (2 * r) ** 2
    lut = ravu_np.lut_array(v.lut, "fp16")
    for ps, name in enumerate(("ravu_int11", "ravu_int10", "ravu_int01")):
        cols = [_gather_lattice(out, dx2, dy2).numpy() for dx2, dy2 in _chain_taps(r)[ps]]
        w_all = lut[res.keys[ps].row].reshape(res.keys[ps].row.shape + (-1,))
        acc = np.zeros_like(cols[0])
        for k in range(N // 2):
            acc = acc + (cols[k] + cols[N - 1 - k]) * w_all[..., k]
        assert np.array_equal(np.clip(acc, 0, 1), inter[name][..., 0]), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ravu-r2.hook", "ravu-r3.hook"])
def test_chain_trainer_recovers_the_lut_that_made_the_targets(name, tmp_path):
    """Targets made by the shipped three-pass hook are, per bucket, a linear function of the pooled windows of its three
    passes: one round of least squares is the fixed point and returns the shipped LUT; a second round stays there."""
    import torch

    from mpv_prescalers_b200 import HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from mpv_prescalers_b200.train import train_ravu_chain, write_hook_with_lut
    from tests.parity import psnr

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    hk = HookFile.parse(hook_path(name))
    rng = np.random.default_rng(6)
    x = batch(6, 1, 300, 400, config=94)[:, 0]
    x = np.clip(0.2 + 0.6 * x + rng.normal(0, 0.04, x.shape), 0.02, 0.98).astype(np.float32)
    lr = torch.from_numpy(x).cuda()
    hr = prescale(lr, hk)
    half = (2 * hk.variant.radius) ** 2 // 2
    ref = np.asarray(hk.variant.lut.data, np.float32).astype(np.float16).astype(np.float32).reshape(hk.variant.lut.height, -1)[:, :half]
    for rounds in (1, 2):
        lut, count = train_ravu_chain(hk, lr, hr, rounds=rounds)
        got = lut.reshape(lut.shape[0], -1)[:, :half]
        well = count >= 5000
        assert well.sum() >= 20, f"only {well.sum()} well-populated buckets"
        assert np.abs(got[well] - ref[well]).max() <= 3e-3, f"rounds={rounds}: recovered LUT differs by {np.abs(got[well] - ref[well]).max():.2e}"
    out = tmp_path / "retrained.hook"
    write_hook_with_lut(hk, lut, str(out))
    fresh = torch.from_numpy(np.clip(0.2 + 0.6 * batch(1, 1, 200, 300, config=95)[:, 0], 0, 1).astype(np.float32)).cuda()
    assert psnr(prescale(fresh, hk).cpu().numpy(), prescale(fresh, str(out)).cpu().numpy()) >= 55.0


def test_zoom_node_model_reproduces_the_oracle():
    """The trainer's linear model of ravu-zoom -- node filters blended with (1 - f, f) around 8 * sub, first half of the taps
    at (sx, sy), mirrored half at (1 - sx, 1 - sy) -- evaluated in float64 with the shipped LUT gives the oracle's output."""
    torch = pytest.importorskip("torch")
    from mpv_prescalers_b200 import HookFile
    from mpv_prescalers_b200.synth import batch
    from mpv_prescalers_b200.train import _node_basis, _zoom_axis
    from oracle import ravu_np

    v = HookFile.parse(hook_path("ravu-zoom-r2.hook")).variant
    h, w, OH, OW = 26, 34, 61, 83
    x = batch(1, 1, h, w, config=96)[0, 0]
    res = ravu_np.ravu_zoom(x, v, (OW, OH))
    r = v.radius
    n = 2 * r
    N, H2 = n * n, n * n // 2
    B = (H2 + 3) // 4
    lut = ravu_np.lut_array(v.lut, "fp16")
    wn = lut.reshape(288, 9, B, 9, 4).transpose(0, 1, 3, 2, 4).reshape(288, 9, 9, B * 4)[..., :H2].astype(np.float64)
    bx, sx = _zoom_axis(w, OW, "cpu")
    by, sy = _zoom_axis(h, OH, "cpu")
    src = torch.from_numpy(x)
    S = np.stack([src[(by + (t % n - (r - 1))).clamp(0, h - 1)[:, None], (bx + (t // n - (r - 1))).clamp(0, w - 1)[None, :]].double().numpy()
                  for t in range(N)], -1)
    row = res.keys[0].row
    out = np.zeros((OH, OW))
    for tx, ty, sel in ((sx, sy, list(range(H2))), (1 - sx, 1 - sy, [N - 1 - k for k in range(H2)])):
        x0, x1, wx0, wx1 = [a.numpy() for a in _node_basis(tx)]
        y0, y1, wy0, wy1 = [a.numpy() for a in _node_basis(ty)]
        for ya, wya in ((y0, wy0), (y1, wy1)):
            for xa, wxa in ((x0, wx0), (x1, wx1)):
                out += (wya[:, None] * wxa[None, :]) * np.sum(wn[row, ya[:, None], xa[None, :]] * S[..., sel], -1)
    assert np.abs(np.clip(out, 0, 1) - res.out).max() <= 1e-4


@pytest.mark.gpu
def test_zoom_trainer_recovers_the_lut_that_made_the_targets(tmp_path):
    """Targets made by the shipped ravu-zoom-r2 LUT at three scale factors are linear in its 81 node filters per bucket; the
    regularised least squares returns them (nodes of well-populated buckets), and the retrained file reproduces the
    original on an unseen plane at a fourth scale factor."""
    import torch

    from mpv_prescalers_b200 import HookFile, prescale
    from mpv_prescalers_b200.synth import batch
    from mpv_prescalers_b200.train import train_ravu_zoom, write_hook_with_lut
    from tests.parity import psnr

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device")
    hk = HookFile.parse(hook_path("ravu-zoom-r2.hook"))
    rng = np.random.default_rng(8)
    pairs = []
    for k, (ry, rx) in enumerate(((1.37, 1.61), (2.23, 1.93), (2.9, 3.1))):
        x = batch(4, 1, 200, 260, config=97 + k)[:, 0]
        x = np.clip(0.2 + 0.6 * x + rng.normal(0, 0.04, x.shape), 0.02, 0.98).astype(np.float32)
        lr = torch.from_numpy(x).cuda()
        pairs.append((lr, prescale(lr, hk, output_size=(int(200 * ry), int(260 * rx)))))
    lut, count = train_ravu_zoom(hk, pairs)
    ref = np.asarray(hk.variant.lut.data, np.float32).astype(np.float16).astype(np.float32)
    well = np.repeat(count >= 30000, 9)                                       # the nine node rows of a well-populated bucket
    assert well.sum() >= 9 * 10, f"only {well.sum() // 9} well-populated buckets"
    err = np.abs(lut[well] - ref[well]).max()
    assert err <= 5e-3, f"recovered node filters differ by {err:.2e}"
    out = tmp_path / "retrained.hook"
    write_hook_with_lut(hk, lut, str(out))
    fresh = torch.from_numpy(np.clip(0.2 + 0.6 * batch(1, 1, 150, 210, config=99)[:, 0], 0, 1).astype(np.float32)).cuda()
    a, b = prescale(fresh, hk, output_size=(330, 400)), prescale(fresh, str(out), output_size=(330, 400))
    assert psnr(a.cpu().numpy(), b.cpu().numpy()) >= 55.0


def test_trainer_refuses_families_it_does_not_cover():
    pytest.importorskip("torch")
    import torch

    from mpv_prescalers_b200 import HookError, HookFile
    from mpv_prescalers_b200.train import train_ravu, train_ravu_chain, train_ravu_lite, train_ravu_zoom

    z = torch.zeros((1, 4, 4))
    for name in ("ravu-r2.hook", "ravu-zoom-r2.hook", "compute/ravu-3x-r2-rgb.hook"):
        with pytest.raises(HookError, match="single-pass luma"):
            train_ravu(HookFile.parse(hook_path(name)), z, z)
    with pytest.raises(HookError, match="RAVU-Lite"):
        train_ravu_lite(HookFile.parse(hook_path("compute/ravu-3x-r2.hook")), z, z)
    for name in ("ravu-lite-r2.hook", "ravu-r2-rgb.hook"):
        with pytest.raises(HookError, match="three-pass luma"):
            train_ravu_chain(HookFile.parse(hook_path(name)), z, z)
    for name in ("ravu-zoom-ar-r2.hook", "ravu-zoom-r2-yuv.hook", "ravu-r2.hook"):
        with pytest.raises(HookError, match="without anti-ringing"):
            train_ravu_zoom(HookFile.parse(hook_path(name)), [(z, z)])
