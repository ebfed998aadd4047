set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py -q -x -k "craft or string or dropin" > gpurun_out/r2x_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
python tools/prof_pipeline.py > gpurun_out/r2x_prof.log 2>&1
head -1 gpurun_out/r2x_prof.log; grep "upsample" gpurun_out/r2x_prof.log
python tools/prof_pipeline.py 8 10 > gpurun_out/r2x_prof10.log 2>&1
head -1 gpurun_out/r2x_prof10.log
