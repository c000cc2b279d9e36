#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/gpu_profile_round.sh r01): bench lines of every workload, the ncu launch
# lists and one `ncu --set full` capture of the two headline kernels.  Everything lands in gpurun_out/<tag>/.
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in ravu-lite-ar-r3 ravu-lite-r3-540p ravu-r4 ravu-r3-rgb ravu-zoom-r3 ravu-zoom-ar-r2 ravu-3x-r3 nnedi3-nns256-win8x6 nnedi3-nns32-win8x4; do
  extra="--no-cpu-baseline"
  [ "$wl" = "ravu-lite-ar-r3" ] && extra=""
  timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 $extra 2>$OUT/bench_$wl.err | tail -1 > $OUT/${TAG}_bench_$wl.json
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $OUT/${TAG}_bench_reference_cpu.json
# launch lists (per-launch durations, serialised and cold-cache: the kernel's SHARE of the step is what must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_ravu_lite_ar_r3.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_nnedi3_nns256_win8x6.csv \
  python bench.py --workload nnedi3-nns256-win8x6 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ravu_lite -s 3 -c 1 -f -o $OUT/${TAG}_ravu_lite_ar_r3 \
  python tools/sweep_lite.py ravu-lite-ar-r3.hook 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nnedi3_tc -s 2 -c 1 -f -o $OUT/${TAG}_nnedi3_nns256_win8x6 \
  python bench.py --workload nnedi3-nns256-win8x6 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ravu_zoom -s 2 -c 1 -f -o $OUT/${TAG}_ravu_zoom_r3 \
  python bench.py --workload ravu-zoom-r3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ravu_kernel -s 2 -c 1 -f -o $OUT/${TAG}_ravu_r4 \
  python tools/sweep_lite.py ravu-r4.hook 16 > /dev/null 2>&1
ls -la $OUT
