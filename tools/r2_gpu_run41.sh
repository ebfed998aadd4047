mkdir -p gpurun_out
timeout 300 python tools/prof_b1.py 20 > gpurun_out/r4b_b1.log 2>&1; echo rc=$?
