#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/gpu_profile_round2.sh r02): the driver-style default bench (headline + secondary
# workloads), the reference arm, per-workload bench lines, ncu launch lists and `ncu --set full` summaries of the
# dominant kernels (both NNEDI3 passes, for the traffic of the intermediate image), the all-variants table.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/bench_default.err | tail -1 > $OUT/${TAG}_bench_default.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 2>/dev/null | tail -1 > $OUT/${TAG}_bench_reference_cpu.json
for io in u8 u10 f16out; do
  timeout 600 python bench.py --io $io --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/${TAG}_bench_ravu-lite-ar-r3_io-$io.json
done
for wl in ravu-lite-r3-540p nnedi3-nns32-win8x4; do
  timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline --secondary none 2>/dev/null | tail -1 > $OUT/${TAG}_bench_$wl.json
done
MPVP_ZOOM_PHASE=0 timeout 600 python bench.py --workload ravu-zoom-r3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --secondary none 2>/dev/null | tail -1 > $OUT/${TAG}_bench_ravu-zoom-r3_general-path.json
MPVP_ZOOM_TEX=1 timeout 600 python bench.py --workload ravu-zoom-r3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --secondary none 2>/dev/null | tail -1 > $OUT/${TAG}_bench_ravu-zoom-r3_texture-unit.json
# launch lists (per-launch durations, serialised and cold-cache: the kernel's SHARE of the step is what must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ravu_lite -c 10 --csv --log-file $OUT/${TAG}_launches_ravu_lite_ar_r3.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nnedi3 -c 12 --csv --log-file $OUT/${TAG}_launches_nnedi3_nns256_win8x6.csv \
  python bench.py --workload nnedi3-nns256-win8x6 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:zoom -c 12 --csv --log-file $OUT/${TAG}_launches_ravu_zoom_r3.csv \
  python bench.py --workload ravu-zoom-r3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none > /dev/null 2>&1
# full captures
cap() {  # name regex skip title units cmd...
  local name=$1 rx=$2 skip=$3 title=$4 units=$5; shift 5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py $OUT/$name.ncu-rep "$title" $units > $OUT/${TAG}_${name}_ncu.txt
  rm -f $OUT/$name.ncu-rep
}
cap ravu_lite_ar_r3 ravu_lite 3 "$TAG ravu-lite-ar-r3 1080p->2160p, 64-frame launch (TMA staging, binary16 LUT + AR power tile in smem, FFMA2 AR sums)" 4147200 python tools/sweep_lite.py ravu-lite-ar-r3.hook 64
cap nnedi3_nns256_win8x6_pass1 nnedi3_tc 2 "$TAG nnedi3-nns256-win8x6, double_y pass of two 2160p frames (tcgen05, TMA-staged windows)" 518400 python bench.py --workload nnedi3-nns256-win8x6 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none
cap nnedi3_nns256_win8x6_pass2 nnedi3_tc 3 "$TAG nnedi3-nns256-win8x6, double_x pass of two 2160p frames (input = the 3840x4320 image of pass 1)" 1036800 python bench.py --workload nnedi3-nns256-win8x6 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none
cap ravu_zoom_r3_phase zoom_phase 2 "$TAG ravu-zoom-r3 720p->2160p, 8-frame launch, phase kernel; unit = 32 OUTPUT pixels" 2073600 python bench.py --workload ravu-zoom-r3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none
cap ravu_zoom_r3_key zoom_key 2 "$TAG ravu-zoom-r3 720p->2160p, 8-frame launch, key pre-pass; unit = 32 SOURCE cells" 231011 python bench.py --workload ravu-zoom-r3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none
cap ravu_r4 ravu_kernel 2 "$TAG ravu-r4 luma 1080p->2160p, 16-frame launch (TMA staging, one buffer)" 1036800 python tools/sweep_lite.py ravu-r4.hook 16
cap ravu_r3_rgb ravu_kernel 2 "$TAG compute/ravu-r3-rgb 1080p->2160p, 16-frame launch (TMA staging of the three planes)" 1036800 python bench.py --workload ravu-r3-rgb --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --secondary none
timeout 900 python tools/bench_all_variants.py $OUT/${TAG}_variants.json > /dev/null 2>&1
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize.py" >> $OUT/${TAG}_sanitizer.txt
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -v "^$" | tail -40 >> $OUT/${TAG}_sanitizer.txt
done
ls -la $OUT; du -sh gpurun_out
