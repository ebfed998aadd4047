mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4f_b1_launches.csv python tools/ncu_b1.py > gpurun_out/r4f_ncu.log 2>&1; echo rc=$?
wc -l gpurun_out/r4f_b1_launches.csv
