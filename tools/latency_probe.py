import torch, time, sys
sys.path.insert(0, "/root/repo")
from mpv_prescalers_b200 import prescale, HookFile, find_hook
for hook, shape, osz in (("ravu-lite-r3.hook", (1, 540, 960), None), ("ravu-lite-ar-r3.hook", (1, 1080, 1920), None), ("ravu-r3.hook", (1, 1080, 1920), None),
                         ("nnedi3-nns32-win8x4.hook", (1, 1080, 1920), None), ("ravu-zoom-r3.hook", (1, 720, 1280), (2160, 3840))):
    x = torch.rand(*shape, device="cuda")
    hk = HookFile.parse(find_hook(hook))
    for _ in range(20):
        y = prescale(x, hk, output_size=osz)
    torch.cuda.synchronize()
    n = 300
    t = time.perf_counter()
    for _ in range(n):
        y = prescale(x, hk, output_size=osz)
    t_issue = time.perf_counter() - t
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = prescale(x, hk, output_size=osz); b.record(); torch.cuda.synchronize()
    print(f"{hook:28s} host issue {t_issue / n * 1e6:7.1f} us/call, steady state {t_all / n * 1e6:7.1f} us/call, one call on the device {a.elapsed_time(b) * 1e3:7.1f} us")

if len(sys.argv) > 1 and sys.argv[1] == "--profile":
    import cProfile, pstats
    x = torch.rand(1, 540, 960, device="cuda")
    hk = HookFile.parse(find_hook("ravu-lite-r3.hook"))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        y = prescale(x, hk)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
