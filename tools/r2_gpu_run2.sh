set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_smoke.log
python bench.py --steps 20 --warmup 3 --no-dropin > gpurun_out/r2b_bench4.log 2> gpurun_out/r2b_bench4.err; echo "rc=$?" >> gpurun_out/r2b_bench4.err
LOCR_CRNN_PREC=exact python tools/prof_pipeline.py > gpurun_out/r2b_prof_exact.log 2>&1
python tools/prof_pipeline.py > gpurun_out/r2b_prof_fast.log 2>&1
