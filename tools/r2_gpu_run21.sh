set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k "hstream or cta2" > gpurun_out/r2r_conv_hs.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_conv_hs.log
tail -12 gpurun_out/r2r_conv_hs.log
grep -q "rc=0" gpurun_out/r2r_conv_hs.log || exit 1
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2r_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_conv.log
tail -4 gpurun_out/r2r_conv.log
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s -k "craft" > gpurun_out/r2r_craft.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_craft.log
tail -3 gpurun_out/r2r_craft.log
python tools/prof_pipeline.py > gpurun_out/r2r_prof.log 2>&1
LOCR_CONV_HSTREAM=0 python tools/prof_pipeline.py > gpurun_out/r2r_prof_nohs.log 2>&1
head -1 gpurun_out/r2r_prof.log gpurun_out/r2r_prof_nohs.log
grep "slice1.7\|slice1.10" gpurun_out/r2r_prof.log gpurun_out/r2r_prof_nohs.log
