set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2y_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_conv.log
tail -12 gpurun_out/r2y_conv.log
grep -q "rc=0" gpurun_out/r2y_conv.log || exit 1
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s -k "craft" > gpurun_out/r2y_craft.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_craft.log
grep "score max-abs\|passed\|failed\|rc=" gpurun_out/r2y_craft.log
python tools/prof_pipeline.py > gpurun_out/r2y_prof.log 2>&1
LOCR_CONV_KSPLIT=0 python tools/prof_pipeline.py > gpurun_out/r2y_prof_nok.log 2>&1
head -1 gpurun_out/r2y_prof.log gpurun_out/r2y_prof_nok.log
grep "slice1.3" gpurun_out/r2y_prof.log gpurun_out/r2y_prof_nok.log
