# usage: bash tools/prof_round.sh <tag>   (e.g. r01e)  — launch list of the bench + ncu --set full captures
TAG=${1:-r01e}
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 720 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
export LOCR_BENCH_WARMUP=0
python tools/prof_kernels.py tc > gpurun_out/prof_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_tc|lstm_cluster' -o gpurun_out/prof_tc_${TAG} python tools/prof_kernels.py tc > gpurun_out/prof_tc_ncu.log 2>&1
python tools/prof_kernels.py mem > gpurun_out/prof_mem_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'pp_|crop_resize|tps_sample|decode|maxpool|upsample|direct_conv|loc_head|preproc' -o gpurun_out/prof_mem_${TAG} python tools/prof_kernels.py mem > gpurun_out/prof_mem_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/prof_tc_plain.log; tail -n 3 gpurun_out/prof_mem_plain.log
