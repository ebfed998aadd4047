set -x
mkdir -p gpurun_out
bash tools/prof_round.sh r02c list
bash tools/conv_traffic.sh r02c
bash tools/prof_round.sh r02c tc
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,launch__cluster_dim_x"
ncu -i gpurun_out/prof_tc_r02c.ncu-rep --page raw --csv --metrics $M > gpurun_out/prof_tc_r02c_raw.csv 2> gpurun_out/prof_tc_r02c_raw.err
ls -la gpurun_out/ | tail -20
SZ=$(du -sm gpurun_out | cut -f1)
if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/prof_tc_r02c.ncu-rep; echo "ncu-rep dropped ($SZ MB), raw csv kept"; fi
du -sm gpurun_out
