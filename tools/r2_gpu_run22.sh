set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2s_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_conv.log
tail -12 gpurun_out/r2s_conv.log
grep -q "rc=0" gpurun_out/r2s_conv.log || exit 1
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s -k "craft" > gpurun_out/r2s_craft.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_craft.log
grep "score max-abs\|passed\|rc=" gpurun_out/r2s_craft.log
python tools/prof_pipeline.py > gpurun_out/r2s_prof.log 2>&1
LOCR_CONV_HALO_PAIR=0 python tools/prof_pipeline.py > gpurun_out/r2s_prof_nopair.log 2>&1
head -1 gpurun_out/r2s_prof.log gpurun_out/r2s_prof_nopair.log
grep "slice1.7\|slice1.10\|slice1.3\|upconv2.conv.3" gpurun_out/r2s_prof.log gpurun_out/r2s_prof_nopair.log
