set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1 || { echo "smoke failed" >> gpurun_out/r2c_smoke.log; tail -5 gpurun_out/r2c_smoke.log; exit 1; }
python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
python tools/prof_pipeline.py > gpurun_out/r2c_prof_fused.log 2>&1
LOCR_FIRST_FUSED=0 python tools/prof_pipeline.py > gpurun_out/r2c_prof_unfused.log 2>&1
python tools/prof_pipeline.py 1 20 > gpurun_out/r2c_prof_b1.log 2>&1
python bench.py --steps 20 --warmup 3 --no-dropin --no-other-precision --no-cpu-baseline > gpurun_out/r2c_bench4.log 2> gpurun_out/r2c_bench4.err; echo "rc=$?" >> gpurun_out/r2c_bench4.err
