# usage: bash tools/prof_round.sh <tag>   (e.g. r02a)  — launch list of the bench + ncu --set full captures
TAG=${1:-r02a}
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-other-precision"
export LOCR_BENCH_PASSES=1
# launches 0..515 are the three warm-up passes of the e2e leg (2 lanes x 86 launches per pass); the next 344 are its timed region
$BENCH > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 340 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
unset LOCR_BENCH_PASSES
export LOCR_BENCH_WARMUP=0
python tools/prof_kernels.py tc > gpurun_out/prof_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_tc|lstm_cluster' -o gpurun_out/prof_tc_${TAG} python tools/prof_kernels.py tc > gpurun_out/prof_tc_ncu.log 2>&1
python tools/prof_kernels.py mem > gpurun_out/prof_mem_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'pp_|crop_resize|tps_sample|decode|maxpool|upsample|direct_conv|loc_head|preproc|png_|jpeg_' -o gpurun_out/prof_mem_${TAG} python tools/prof_kernels.py mem > gpurun_out/prof_mem_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/prof_tc_plain.log; tail -n 3 gpurun_out/prof_mem_plain.log
