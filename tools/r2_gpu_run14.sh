set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2m_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_conv.log
tail -15 gpurun_out/r2m_conv.log
grep -q "rc=0" gpurun_out/r2m_conv.log || exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
for g in 0 1 2 3; do LOCR_CONV_KGROUP=$g python tools/prof_pipeline.py > gpurun_out/r2m_prof_kg$g.log 2>&1; head -1 gpurun_out/r2m_prof_kg$g.log; done
