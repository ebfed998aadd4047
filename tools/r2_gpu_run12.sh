set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x > gpurun_out/r2k_conv.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_conv.log
tail -15 gpurun_out/r2k_conv.log
grep -q "rc=0" gpurun_out/r2k_conv.log || exit 1
LOCR_CONV_CTA2_N256=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
LOCR_CONV_CTA2_N256=1 python tools/prof_pipeline.py > gpurun_out/r2k_prof_n256.log 2>&1
python tools/prof_pipeline.py > gpurun_out/r2k_prof_base.log 2>&1
head -12 gpurun_out/r2k_prof_n256.log; head -12 gpurun_out/r2k_prof_base.log
B="python bench.py --steps 10 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
$B > gpurun_out/r2k_bench_base.log 2>gpurun_out/r2k_bench_base.err
LOCR_CONV_CTA2_N256=1 $B > gpurun_out/r2k_bench_n256.log 2>gpurun_out/r2k_bench_n256.err
$B > gpurun_out/r2k_bench_base2.log 2>gpurun_out/r2k_bench_base2.err
LOCR_CONV_CTA2_N256=1 $B > gpurun_out/r2k_bench_n256b.log 2>gpurun_out/r2k_bench_n256b.err
for f in base n256 base2 n256b; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2k_bench_$f.log").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), round(d["roofline"]["whole_step_tensor_frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
PY
done
