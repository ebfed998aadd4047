set -x
mkdir -p gpurun_out
python -m pytest tests/test_png_gpu.py -q -s > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -3 gpurun_out/r2g_tests.log
python tools/bench_png.py 8 > gpurun_out/r2g_bench_png.log 2>&1
cat gpurun_out/r2g_bench_png.log
python tools/bench_jpeg.py 8 > gpurun_out/r2g_bench_jpeg.log 2>&1
tail -4 gpurun_out/r2g_bench_jpeg.log
