set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2o_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_conv.log
tail -12 gpurun_out/r2o_conv.log
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s -k "craft" > gpurun_out/r2o_craft.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_craft.log
tail -12 gpurun_out/r2o_craft.log
python tools/prof_pipeline.py > gpurun_out/r2o_prof.log 2>&1
LOCR_HEAD_HALO=0 python tools/prof_pipeline.py > gpurun_out/r2o_prof_nohalo.log 2>&1
head -1 gpurun_out/r2o_prof.log gpurun_out/r2o_prof_nohalo.log
grep "upconv4\|conv_cls" gpurun_out/r2o_prof.log gpurun_out/r2o_prof_nohalo.log
