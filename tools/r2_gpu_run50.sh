mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r4k_tests.log 2>&1; echo "tests rc=$? after ${SECONDS}s" >> gpurun_out/r4k_tests.log
tail -4 gpurun_out/r4k_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
