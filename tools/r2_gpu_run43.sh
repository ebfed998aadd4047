mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r4d_tests.log 2>&1; echo "tests rc=$? after ${SECONDS}s" >> gpurun_out/r4d_tests.log
tail -6 gpurun_out/r4d_tests.log
timeout 200 python tools/prof_dropin.py 8 > gpurun_out/r4d_dropin.log 2>&1; echo rc=$?
head -3 gpurun_out/r4d_dropin.log
