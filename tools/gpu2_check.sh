#!/bin/bash
# 2-GPU check (gpurun --gpus 2 -- bash tools/gpu2_check.sh): sharding tests, torchrun bench at N=2, 8K row split timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_split or sharding" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_ravu_lite_ar_r3_2gpu.json
python -c "import json; d=json.load(open('gpurun_out/r01_bench_ravu_lite_ar_r3_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
python - <<PY
import torch, time, json
from mpv_prescalers_b200 import prescale
x = torch.rand(1, 4320, 7680, device="cuda:0")
res = {}
for hook in ("ravu-lite-ar-r3.hook", "ravu-r4.hook", "nnedi3-nns256-win8x6.hook"):
  for devs in ([0], [0, 1]):
    run = (lambda: prescale(x, hook, devices=devs, split="rows")) if len(devs) > 1 else (lambda: prescale(x, hook))
    for _ in range(2):
        o = run()
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    t = time.perf_counter()
    for _ in range(5):
        o = run()
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    res[hook + " " + str(devs)] = (time.perf_counter() - t) / 5 * 1e3
    print("8K frame", hook, "on", devs, res[hook + " " + str(devs)], "ms")
json.dump(res, open("gpurun_out/r01_rowsplit_8k.json", "w"))
PY
