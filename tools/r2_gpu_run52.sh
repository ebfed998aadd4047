mkdir -p gpurun_out
timeout 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-precision --no-other-head > gpurun_out/r4m_bench.log 2> gpurun_out/r4m_bench.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r4m_bench.log").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["e2e_dropin"])
PY
