set -x
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
$B > gpurun_out/r3a_l2.log 2>gpurun_out/r3a_l2.err
LOCR_BENCH_LANES=3 $B > gpurun_out/r3a_l3.log 2>gpurun_out/r3a_l3.err
LOCR_BENCH_LANES=4 LOCR_BENCH_PER_LANE=4 $B > gpurun_out/r3a_l4p4.log 2>gpurun_out/r3a_l4p4.err
LOCR_BENCH_LANES=2 LOCR_BENCH_PER_LANE=12 $B > gpurun_out/r3a_l2p12.log 2>gpurun_out/r3a_l2p12.err
$B > gpurun_out/r3a_l2b.log 2>gpurun_out/r3a_l2b.err
for f in l2 l3 l4p4 l2p12 l2b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3a_$f.log").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f failed", e)
PY
done
