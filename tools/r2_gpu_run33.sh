set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py tests/test_pipeline_gpu.py -q -x > gpurun_out/r3b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3b_tests.log
tail -4 gpurun_out/r3b_tests.log
grep -q "rc=0" gpurun_out/r3b_tests.log || exit 1
python tools/prof_pipeline.py > gpurun_out/r3b_prof.log 2>&1
LOCR_CONV_TAIL=0 python tools/prof_pipeline.py > gpurun_out/r3b_prof_notail.log 2>&1
head -1 gpurun_out/r3b_prof.log gpurun_out/r3b_prof_notail.log
grep "layer3.2.conv2\|layer4.0.conv1\|conv3 " gpurun_out/r3b_prof.log gpurun_out/r3b_prof_notail.log
B="python bench.py --steps 8 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
for cfg in "2 8" "3 8" "2 12" "3 12" "2 16" "4 8"; do set -- $cfg; LOCR_BENCH_LANES=$1 LOCR_BENCH_PER_LANE=$2 $B > gpurun_out/r3b_l$1p$2.log 2>gpurun_out/r3b_l$1p$2.err; done
LOCR_CONV_TAIL=0 LOCR_BENCH_LANES=3 LOCR_BENCH_PER_LANE=8 $B > gpurun_out/r3b_l3p8_notail.log 2>gpurun_out/r3b_l3p8_notail.err
for f in l2p8 l3p8 l2p12 l3p12 l2p16 l4p8 l3p8_notail; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3b_$f.log").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f failed", e)
PY
done
