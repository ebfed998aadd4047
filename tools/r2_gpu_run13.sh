set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
tail -5 gpurun_out/r2l_tests.log
grep -q "rc=0" gpurun_out/r2l_tests.log || exit 1
python tools/prof_pipeline.py > gpurun_out/r2l_prof.log 2>&1
LOCR_CRNN_PREC=exact python tools/prof_pipeline.py > gpurun_out/r2l_prof_exact.log 2>&1
LOCR_CONV_CTA2_MINKB=8 python tools/prof_pipeline.py > gpurun_out/r2l_prof_kb8.log 2>&1
head -3 gpurun_out/r2l_prof.log gpurun_out/r2l_prof_exact.log gpurun_out/r2l_prof_kb8.log
python bench.py --steps 10 --warmup 3 --no-dropin --no-cpu-baseline > gpurun_out/r2l_bench.log 2>gpurun_out/r2l_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_bench.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), round(d["roofline"]["whole_step_tensor_frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["power_w"], d.get("other_precision"))
PY
