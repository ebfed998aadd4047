mkdir -p gpurun_out
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/r4l_bench.log 2> gpurun_out/r4l_bench.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r4l_bench.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["other_precision"]["value"], d["other_head"]["value"], d["clocks"])
PY
