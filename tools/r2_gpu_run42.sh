mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -m gpu > gpurun_out/r4c_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4c_conv.log
tail -15 gpurun_out/r4c_conv.log
timeout 300 python tools/prof_b1.py 20 > gpurun_out/r4c_b1.log 2>&1; echo rc=$?
grep "^==" gpurun_out/r4c_b1.log
