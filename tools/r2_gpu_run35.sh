set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r3d_bench4.log 2> gpurun_out/r3d_bench4.err; echo "rc=$?" >> gpurun_out/r3d_bench4.err
python bench.py --steps 10 --warmup 3 --jpeg --no-dropin --no-other-precision --no-cpu-baseline > gpurun_out/r3d_bench_jpeg.log 2> gpurun_out/r3d_bench_jpeg.err
python bench.py --config 5 --steps 10 --warmup 3 --no-dropin --no-cpu-baseline > gpurun_out/r3d_bench5.log 2> gpurun_out/r3d_bench5.err
for f in 4 _jpeg 5; do python - <<PY
import json
d=json.loads(open("gpurun_out/r3d_bench$f.log").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["roofline"]["whole_step_tensor_frac"], d["roofline"]["kernel_share_of_step"], d["clocks"]["sm_mhz"], (d.get("other_precision") or {}).get("value"), (d.get("e2e_dropin") or {}).get("value"), (d.get("e2e_jpeg") or {}).get("value"), d["gpu_launches"], d["e2e"]["h2d_bytes_per_step"], d["config"]["batching"] if "batching" in d["config"] else "")
PY
done
tail -2 gpurun_out/r3d_bench4.err
