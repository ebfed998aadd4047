set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -s > gpurun_out/r2i_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_conv.log
tail -15 gpurun_out/r2i_conv.log
grep -q "rc=0" gpurun_out/r2i_conv.log || exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -2 gpurun_out/r2i_smoke.log
python tools/prof_pipeline.py > gpurun_out/r2i_prof_cta2.log 2>&1
LOCR_CONV_CTA2=0 python tools/prof_pipeline.py > gpurun_out/r2i_prof_nocta2.log 2>&1
head -3 gpurun_out/r2i_prof_cta2.log; head -3 gpurun_out/r2i_prof_nocta2.log
