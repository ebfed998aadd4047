set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2p_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_conv.log
tail -4 gpurun_out/r2p_conv.log
grep -q "rc=0" gpurun_out/r2p_conv.log || exit 1
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s -k "craft" > gpurun_out/r2p_craft.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_craft.log
tail -3 gpurun_out/r2p_craft.log
python tools/prof_pipeline.py > gpurun_out/r2p_prof.log 2>&1
LOCR_CONV_HALO_STAGES=2 python tools/prof_pipeline.py > gpurun_out/r2p_prof_st2.log 2>&1
LOCR_CONV_HALO_STAGES=4 python tools/prof_pipeline.py > gpurun_out/r2p_prof_st4.log 2>&1
head -1 gpurun_out/r2p_prof.log gpurun_out/r2p_prof_st2.log gpurun_out/r2p_prof_st4.log
grep "upconv4\|conv_cls\|slice1.3" gpurun_out/r2p_prof.log gpurun_out/r2p_prof_st2.log gpurun_out/r2p_prof_st4.log
python tools/prof_pipeline.py 1 20 > gpurun_out/r2p_prof_b1.log 2>&1
head -1 gpurun_out/r2p_prof_b1.log
python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench2.log 2> gpurun_out/r2p_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_bench2.log").read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
