set -x
mkdir -p gpurun_out
python -m pytest tests/test_poly_gpu.py tests/test_png_gpu.py -q -s > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -3 gpurun_out/r2e_tests.log
bash tools/prof_round.sh r02a tc
du -sh gpurun_out
