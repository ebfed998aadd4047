set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_golden_gpu.py -q -x -s -k "craft or dropin or string or golden" > gpurun_out/r2v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_tests.log
grep "score max-abs\|passed\|failed\|rc=" gpurun_out/r2v_tests.log | tail -15
python tools/prof_pipeline.py > gpurun_out/r2v_prof.log 2>&1
head -1 gpurun_out/r2v_prof.log; grep "upsample" gpurun_out/r2v_prof.log
