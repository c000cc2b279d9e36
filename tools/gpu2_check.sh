#!/bin/bash
# 2-GPU check (gpurun --gpus 2 -- bash tools/gpu2_check.sh r02): the multi-GPU parity tests (frame sharding and row split
# with REAL peer-to-peer bands), the torchrun bench at N=2, and the 8K row-split timing.  Logs go to gpurun_out/<tag>/.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu2_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -v -k "row_split or sharding" >> $OUT/${TAG}_gpu2_pytest.log 2>&1
tail -5 $OUT/${TAG}_gpu2_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --secondary nnedi3-nns256-win8x6 2>/dev/null | tail -1 > $OUT/${TAG}_bench_default_2gpu.json
python -c "import json; d=json.load(open('$OUT/${TAG}_bench_default_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], [ (s['workload'], s.get('value')) for s in d.get('secondary', [])])"
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
json.dump(res, open("$OUT/${TAG}_rowsplit_8k.json", "w"))
PY
