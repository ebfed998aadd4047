set -x
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r4i_tests.log 2>&1; echo "tests rc=$? after ${SECONDS}s" >> gpurun_out/r4i_tests.log
tail -4 gpurun_out/r4i_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4i_smoke.log 2>&1; tail -1 gpurun_out/r4i_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4i_ref.log 2> gpurun_out/r4i_ref.err
python bench.py > gpurun_out/r4i_bench4.log 2> gpurun_out/r4i_bench4.err; echo "rc=$?" >> gpurun_out/r4i_bench4.err
python tools/prof_pipeline.py 8 3 > gpurun_out/r4i_prof.log 2>&1
python bench.py --config 1 --no-cpu-baseline > gpurun_out/r4i_bench1.log 2>&1
python bench.py --config 2 --no-cpu-baseline > gpurun_out/r4i_bench2.log 2>&1
python bench.py --config 3 --no-cpu-baseline > gpurun_out/r4i_bench3.log 2>&1
bash tools/prof_round.sh r02d list > gpurun_out/r4i_list.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r4i_bench4.log").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r4i_ref.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["roofline"]["whole_step_tensor_frac"], d["roofline"]["kernel_share_of_step"], d["roofline"]["kernel_share_of_gpu_time"], d["clocks"]["sm_mhz"], d["other_precision"]["value"], d["other_head"]["value"], d["e2e_dropin"], d["cpu_baseline"]["value"])
print(d["config"]["workload"] == r["config"]["workload"], r["value"])
for k in (1,2,3):
    x=json.loads(open("gpurun_out/r4i_bench%d.log"%k).read().strip().splitlines()[-1])
    print(k, x["metric"], x["value"], x["e2e"]["value"], x["roofline"]["frac"])
PY
echo total ${SECONDS}s
