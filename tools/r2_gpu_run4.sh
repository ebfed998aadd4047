set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1 || { echo "smoke failed" >> gpurun_out/r2d_smoke.log; tail -5 gpurun_out/r2d_smoke.log; exit 1; }
python -m pytest tests -m gpu -q -s > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
grep -q "failed" gpurun_out/r2d_tests.log && tail -5 gpurun_out/r2d_tests.log
LOCR_CRNN_PREC=exact python tools/prof_pipeline.py > gpurun_out/r2d_prof_exact.log 2>&1
bash tools/prof_round.sh r02a list
bash tools/conv_traffic.sh r02a
du -sh gpurun_out
