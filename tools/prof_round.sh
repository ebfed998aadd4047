set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 720 -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r01d.log 2>&1
export LOCR_BENCH_WARMUP=0
python tools/prof_kernels.py tc > gpurun_out/prof_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_tc|lstm_cluster' -o gpurun_out/prof_tc_r01d python tools/prof_kernels.py tc > gpurun_out/prof_tc_ncu.log 2>&1
python tools/prof_kernels.py mem > gpurun_out/prof_mem_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'pp_|crop_resize|tps_sample|decode|maxpool|upsample|direct_conv|loc_head|preproc' -o gpurun_out/prof_mem_r01d python tools/prof_kernels.py mem > gpurun_out/prof_mem_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/prof_tc_plain.log gpurun_out/prof_mem_plain.log
