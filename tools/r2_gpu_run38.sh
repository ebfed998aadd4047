mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline > gpurun_out/r3g_bench.log 2> gpurun_out/r3g_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r3g_bench.log").read().strip().splitlines()[-1])
r=d["roofline"]
print(round(d["value"],1), round(d["e2e"]["value"],1), r["frac"], r["kernel_share_of_step"], r["kernel_share_of_gpu_time"], r["all_kernels_ms"], r["kernel_ms"], r["pass_ms"])
PY
