set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
python tools/prof_pipeline.py > gpurun_out/r2c_prof_fused.log 2>&1
LOCR_FIRST_FUSED=0 python tools/prof_pipeline.py > gpurun_out/r2c_prof_unfused.log 2>&1
python bench.py --steps 20 --warmup 3 --no-dropin --no-other-precision --no-cpu-baseline > gpurun_out/r2c_bench4.log 2> gpurun_out/r2c_bench4.err; echo "rc=$?" >> gpurun_out/r2c_bench4.err
