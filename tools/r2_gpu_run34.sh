set -x
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
for cfg in "2 8" "3 8" "3 12" "4 8" "3 6" "2 8" "3 8"; do set -- $cfg; LOCR_BENCH_LANES=$1 LOCR_BENCH_PER_LANE=$2 $B > gpurun_out/r3c_l$1p$2.log 2>gpurun_out/r3c_l$1p$2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c_l$1p$2.log").read().strip().splitlines()[-1])
    print("l$1p$2", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
except Exception as e:
    print("l$1p$2 failed", e)
PY
done
