set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py tests/test_pipeline_gpu.py -q -x > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests.log
tail -8 gpurun_out/r2n_tests.log
python tools/prof_pipeline.py > gpurun_out/r2n_prof.log 2>&1
LOCR_FIRST_CH8=0 python tools/prof_pipeline.py > gpurun_out/r2n_prof_ch16.log 2>&1
LOCR_TPS_PASSES=2 python tools/prof_pipeline.py > gpurun_out/r2n_prof_tps2.log 2>&1
head -1 gpurun_out/r2n_prof.log gpurun_out/r2n_prof_ch16.log gpurun_out/r2n_prof_tps2.log
grep "slice1.0\|preproc" gpurun_out/r2n_prof.log gpurun_out/r2n_prof_ch16.log
LOCR_TPS_PASSES=2 timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py -q -s > gpurun_out/r2n_tests_tps2.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests_tps2.log
tail -8 gpurun_out/r2n_tests_tps2.log
