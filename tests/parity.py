"""Shared parity metrics (north_star tolerances, written here once):

* RAVU bucket indices must match on >= 99.99 % of key evaluations, mismatches only at quantisation
  boundaries;
* output max abs error <= 1e-3 on [0,1] wherever the bucket agrees, PSNR >= 60 dB overall;
* NNEDI3 within the same bound.
"""
import numpy as np

BUCKET_MIN_AGREE = 0.9999
MAX_ABS = 1e-3
MIN_PSNR = 60.0


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


def boundary_distance(key, v):
    """Distance of every key evaluation to the nearest quantisation boundary, in the units of each
    quantiser (angle: sectors, cyclic -- sector 0 and sector 23 touch at theta = 0 = pi; strength /
    coherence: relative)."""
    af = np.asarray(key.angle_f, np.float64)
    d_angle = np.minimum(af - np.floor(af), np.ceil(af) - af)
    d_angle = np.where(np.isnan(d_angle), 0.0, d_angle)
    lam = np.asarray(key.lam, np.float64)
    if v.strength_thr:
        d_str = np.min([np.abs(lam - t) / t for t in v.strength_thr], axis=0)
    else:
        val = np.log2(lam * v.strength_log2_scale + 1.192092896e-7)
        d_str = np.abs(val - np.round(val))
    mu = np.asarray(key.mu, np.float64)
    d_coh = np.min([np.abs(mu - t) for t in v.coherence_thr], axis=0)
    d_coh = np.where(np.isnan(d_coh), 0.0, d_coh)
    return d_angle, d_str, d_coh


def on_edge_mask(key, v):
    """Key evaluations that sit on a bucket edge to within a few ulps."""
    d_angle, d_str, d_coh = boundary_distance(key, v)
    return (d_angle < 1e-4) | (d_str < 1e-5) | (d_coh < 1e-5)


def check_buckets(got_rows, key, v, what=""):
    """got_rows, key.row: same shape.  Returns the agreement mask.

    Every comparison made with this function runs the device and the oracle on IDENTICAL inputs (the RAVU steps 2 / 3,
    whose int11 inputs differ in the last bits between device and oracle, are checked against the oracle evaluated on
    the device's own int11), and up to the eigenvalues the device arithmetic is the shader's, operation for operation.
    What differs is only the last step -- sector tests on (b, L1 - a) instead of atan, eigenvalue thresholds instead of
    sqrt / division -- so a mismatch is "at a quantisation boundary" only if the oracle's own continuous coordinate
    lies within a few float32 ulps of a bucket edge; there is no allowance for badly conditioned keys.
    """
    same = np.asarray(got_rows) == key.row
    frac = float(same.mean())
    if not same.all():
        d_angle, d_str, d_coh = (d[~same] for d in boundary_distance(key, v))
        # mismatches that sit ON a bucket edge to within a few ulps (exact diagonals of degenerate planes:
        # 1-pixel-wide images, checkerboards) are not counted against the agreement rate
        on_edge = (d_angle < 1e-4) | (d_str < 1e-5) | (d_coh < 1e-5)
        counted = int((~on_edge).sum())
        assert frac >= BUCKET_MIN_AGREE or counted <= 1, f"{what}: bucket agreement {frac:.6f} < {BUCKET_MIN_AGREE}"
        near = (d_angle < 2e-3) | (d_str < 2e-4) | (d_coh < 2e-4)
        assert np.all(near), (
            f"{what}: bucket mismatch away from a quantisation boundary "
            f"(angle {d_angle[~near]}, strength {d_str[~near]}, coherence {d_coh[~near]})")
    return same


def check_output(got, ref, same_mask_out=None, what=""):
    got = np.asarray(got, np.float32)
    ref = np.asarray(ref, np.float32)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all() or not np.isfinite(ref).all(), f"{what}: non-finite output"
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    diff = np.where(np.isnan(ref), 0.0, diff)
    m = diff if same_mask_out is None else diff[same_mask_out]
    mx = float(m.max()) if m.size else 0.0
    assert mx <= MAX_ABS, f"{what}: max abs error {mx:.3e} > {MAX_ABS}"
    p = psnr(np.nan_to_num(got), np.nan_to_num(ref))
    assert p >= MIN_PSNR, f"{what}: PSNR {p:.1f} dB < {MIN_PSNR}"
    return mx, p
