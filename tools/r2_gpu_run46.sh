mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r4g_tests.log 2>&1; echo "tests rc=$? after ${SECONDS}s" >> gpurun_out/r4g_tests.log
tail -6 gpurun_out/r4g_tests.log
timeout 300 python tools/prof_b1.py 20 > gpurun_out/r4g_b1.log 2>&1; echo rc=$?
grep "^==\|loc_head\|decode\|lstm " gpurun_out/r4g_b1.log
