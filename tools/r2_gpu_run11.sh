set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -s > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -5 gpurun_out/r2j_tests.log
grep -i "validation loss\|CTC loss\|attention CE\|edge targets" gpurun_out/r2j_tests.log
