"""Diagnostics for the full-size parity failures (run on the GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpv_prescalers_b200 import HookFile, find_hook, prescale
from mpv_prescalers_b200.synth import batch
from oracle import ravu_np
from tests.parity import boundary_distance

def zoom(name, h, w, oh, ow, cfg, c=1):
    hk = HookFile.parse(find_hook(name)); v = hk.variant
    x = batch(1, v.channels, h, w, config=cfg)
    xt = torch.from_numpy(x).cuda()
    if v.channels == 1: xt = xt[:, 0]
    out, bk = prescale(xt, hk, output_size=(oh, ow), return_buckets=True)
    out = out.cpu().numpy()[0]; bk = bk.cpu().numpy()[0]
    img = x[0, 0] if v.channels == 1 else np.moveaxis(x[0], 0, -1)
    ref = ravu_np.run(img, v, (ow, oh))
    got = out if v.channels == 1 else np.moveaxis(out, 0, -1)
    same = bk == ref.keys[0].row
    d = np.abs(got - ref.out)
    if d.ndim == 3: dm = d.max(-1)
    else: dm = d
    dm = np.where(same, dm, 0)
    print(f"== {name} {w}x{h}->{ow}x{oh} TEX={os.environ.get('MPVP_ZOOM_TEX','1')}: bucket agree {same.mean():.6f}, max err (agreeing) {dm.max():.3e}, n>1e-3: {(dm>1e-3).sum()}, n>3e-4: {(dm>3e-4).sum()}")
    idx = np.argsort(dm.reshape(-1))[::-1][:8]
    bx, sx = ravu_np.zoom_positions(w, ow); by, sy = ravu_np.zoom_positions(h, oh)
    for i in idx:
        oy, ox = divmod(int(i), ow)
        print(f"   (ox {ox}, oy {oy}) err {dm[oy,ox]:.3e} got {got[oy,ox]} ref {ref.out[oy,ox]} base ({bx[ox]},{by[oy]}) sub ({sx[ox]:.6f},{sy[oy]:.6f}) src {img[min(max(by[oy],0),h-1), min(max(bx[ox],0),w-1)]}")

def ravu(name, h, w, cfg):
    hk = HookFile.parse(find_hook(name)); v = hk.variant
    x = batch(1, 1, h, w, config=cfg)
    out, bk = prescale(torch.from_numpy(x).cuda()[:, 0], hk, return_buckets=True)
    out = out.cpu().numpy()[0]; bk = bk.cpu().numpy()[0]
    ref = ravu_np.run(x[0, 0], v)
    i11 = out[1::2, 1::2]; r11 = ref.out[1::2, 1::2]
    ref2 = ravu_np.ravu(x[0, 0], v, int11_override=i11)
    for k in (1, 2):
        same = bk[k] == ref2.keys[k].row
        print(f" oracle steps 2/3 on the DEVICE int11: key {k}: {int((~same).sum())} mismatches; out max diff {np.abs(out - ref2.out).max():.3e}")
    print(f"== {name} {w}x{h}: int11 max abs diff {np.abs(i11-r11).max():.3e}")
    for k in range(3):
        same = bk[k] == ref.keys[k].row
        da, ds, dc = boundary_distance(ref.keys[k], v)
        ys, xs = np.nonzero(~same)
        print(f" key {k}: {len(ys)} mismatches of {same.size}")
        for y, xx in list(zip(ys, xs))[:40]:
            g = int(bk[k][y, xx]); r = int(ref.keys[k].row[y, xx])
            nb = np.abs(i11[max(y-5,0):y+6, max(xx-5,0):xx+6] - r11[max(y-5,0):y+6, max(xx-5,0):xx+6]).max()
            k0 = (~(bk[0] == ref.keys[0].row))[max(y-6,0):y+7, max(xx-6,0):xx+7].any()
            print(f"   (x {xx}, y {y}) gpu row {g} = (a{g//27},s{(g//3)%9},c{g%3}) ref {r} = (a{r//27},s{(r//3)%9},c{r%3}) angle_f {ref.keys[k].angle_f[y,xx]:.5f} lam {ref.keys[k].lam[y,xx]:.5f} mu {ref.keys[k].mu[y,xx]:.5f} d=({da[y,xx]:.4f},{ds[y,xx]:.4f},{dc[y,xx]:.4f}) int11 nb diff {nb:.2e} key0 flip nearby {k0}")

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "zoomar": zoom("ravu-zoom-ar-r2.hook", 360, 640, 1080, 1920, 4)
    if what == "zoomaryuv": zoom("ravu-zoom-ar-r2-yuv.hook", 48, 60, 144, 180, 13)
    if what == "zoomr3": zoom("ravu-zoom-r3.hook", 720, 1280, 2160, 3840, 4)
    if what == "ravur4": ravu("ravu-r4.hook", 1080, 1920, 3)
