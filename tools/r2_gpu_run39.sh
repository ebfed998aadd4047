set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3h_tests.log
tail -4 gpurun_out/r3h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1; tail -1 gpurun_out/r3h_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3h_ref.log 2> gpurun_out/r3h_ref.err
python bench.py > gpurun_out/r3h_bench4.log 2> gpurun_out/r3h_bench4.err; echo "rc=$?" >> gpurun_out/r3h_bench4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3h_bench4.log").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r3h_ref.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["roofline"]["whole_step_tensor_frac"], d["roofline"]["kernel_share_of_step"], d["clocks"]["sm_mhz"], d["other_precision"]["value"], d["e2e_dropin"]["value"], d["cpu_baseline"]["value"])
print(d["config"]["workload"] == r["config"]["workload"], r["value"])
PY
