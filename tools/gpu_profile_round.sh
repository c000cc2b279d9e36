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
for io in u8 u10 f16out; do
  timeout 600 python bench.py --io $io --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/${TAG}_bench_ravu-lite-ar-r3_io-$io.json
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $OUT/${TAG}_bench_reference_cpu.json
# launch lists (per-launch durations, serialised and cold-cache: the kernel's SHARE of the step is what must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ravu_lite -c 10 --csv --log-file $OUT/${TAG}_launches_ravu_lite_ar_r3.csv \
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
# summaries are made on the box (gpurun copies back at most 64 MiB): keep only the two headline .ncu-rep files
python tools/ncu_summary.py $OUT/${TAG}_ravu_lite_ar_r3.ncu-rep "$TAG ravu-lite-ar-r3 1080p->2160p, 64-frame launch (TMA staging, binary16 LUT + AR power tile in smem, FFMA2 AR sums)" 4147200 > $OUT/${TAG}_ravu_lite_ar_r3_ncu.txt
KN=$(ncu -i $OUT/${TAG}_nnedi3_nns256_win8x6.ncu-rep --page raw --csv 2>/dev/null | python -c "import csv,sys; r=list(csv.reader(sys.stdin)); print(r[2][r[0].index('Kernel Name')])")
case "$KN" in *"<6, 0,"*) NU=518400; PASS="double_y";; *) NU=1036800; PASS="double_x";; esac
python tools/ncu_summary.py $OUT/${TAG}_nnedi3_nns256_win8x6.ncu-rep "$TAG nnedi3-nns256-win8x6, $PASS pass of two 2160p frames (tcgen05, one Newton reciprocal per 8 neurons)" $NU > $OUT/${TAG}_nnedi3_nns256_win8x6_ncu.txt
python tools/ncu_summary.py $OUT/${TAG}_ravu_r4.ncu-rep "$TAG ravu-r4 luma 1080p->2160p, 16-frame launch" 1036800 > $OUT/${TAG}_ravu_r4_ncu.txt
python tools/ncu_summary.py $OUT/${TAG}_ravu_zoom_r3.ncu-rep "$TAG ravu-zoom-r3 720p->2160p, 8-frame launch (texture-unit LUT fetch); unit = 32 OUTPUT pixels" 2073600 > $OUT/${TAG}_ravu_zoom_r3_ncu.txt
rm -f $OUT/${TAG}_ravu_r4.ncu-rep $OUT/${TAG}_ravu_zoom_r3.ncu-rep $OUT/${TAG}_ravu_lite_ar_r3.ncu-rep
ls -la $OUT; du -sh gpurun_out
