mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_nets_gpu.py tests/test_dropin_golden_gpu.py -q -x -m gpu 2>&1 | tail -3
python tools/prof_pipeline.py 8 3 > gpurun_out/r4j_prof.log 2>&1
head -1 gpurun_out/r4j_prof.log; grep "loc_head\|decode\|lstm \|tps_sample" gpurun_out/r4j_prof.log
timeout 100 python tools/prof_b1.py 20 2>&1 | grep "^==\|loc_head"
