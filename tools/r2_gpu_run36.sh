set -x
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r3e_$tag.log 2>gpurun_out/r3e_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3e_$tag.log").read().strip().splitlines()[-1])
    print("$tag", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], round(d["crops_per_sec"]))
except Exception as e:
    print("$tag failed", e)
PY
}
run p8 LOCR_BENCH_PER_LANE=8
run p10c10 LOCR_BENCH_PER_LANE=10 LOCR_CRAFT_CHUNK=10
run p10c8 LOCR_BENCH_PER_LANE=10
run p8b LOCR_BENCH_PER_LANE=8
run p10c10b LOCR_BENCH_PER_LANE=10 LOCR_CRAFT_CHUNK=10
run p9c9 LOCR_BENCH_PER_LANE=9 LOCR_CRAFT_CHUNK=9
