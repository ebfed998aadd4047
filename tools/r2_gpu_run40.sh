set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_edge_gpu.py -q -x -m gpu > gpurun_out/r4a_edge.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4a_edge.log
tail -5 gpurun_out/r4a_edge.log
timeout 400 python bench.py --no-cpu-baseline --no-dropin > gpurun_out/r4a_bench.log 2> gpurun_out/r4a_bench.err; echo "rc=$?" >> gpurun_out/r4a_bench.err
tail -3 gpurun_out/r4a_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r4a_bench.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["other_precision"], d["other_head"])
PY
