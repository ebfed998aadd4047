# usage: bash tools/prof_round.sh <tag> [list|tc|mem]...   (e.g. r02a list tc)
# list = launch list of the bench; tc / mem = ncu --set full captures of the tensor-core / the other kernels.
# gpurun brings back at most 64 MiB: run `tc` and `mem` in separate calls.
TAG=${1:-r02a}
shift
STAGES=${*:-list}
set -x
for ST in $STAGES; do
case $ST in
list)
    BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-other-precision --no-other-head"
    export LOCR_BENCH_PASSES=1
    export LOCR_BENCH_LANES=2     # the skip / count below are for two lanes
    # launches 0..515 are the three warm-up passes of the e2e leg (2 lanes x 86 launches per pass); the next 344 are its timed region
    $BENCH > gpurun_out/plain_${TAG}.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 340 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
    unset LOCR_BENCH_PASSES LOCR_BENCH_LANES
    ;;
tc)
    export LOCR_BENCH_WARMUP=0
    python tools/prof_kernels.py tc > gpurun_out/prof_tc_plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'conv_tc|lstm_cluster' -o gpurun_out/prof_tc_${TAG} python tools/prof_kernels.py tc > gpurun_out/prof_tc_ncu.log 2>&1
    tail -n 3 gpurun_out/prof_tc_plain.log
    ;;
mem)
    python tools/prof_kernels.py memi > gpurun_out/prof_mem_plain.log 2>&1 &&
    ncu --set full --clock-control none -k regex:'pp_|crop_resize|tps_sample|decode|maxpool|upsample|direct_conv|loc_head|preproc|png_|jpeg_' -c 48 -o gpurun_out/prof_mem_${TAG} python tools/prof_kernels.py memi > gpurun_out/prof_mem_ncu.log 2>&1
    tail -n 3 gpurun_out/prof_mem_plain.log
    ;;
esac
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
