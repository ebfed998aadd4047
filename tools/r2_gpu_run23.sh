set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2t_tests.log
tail -4 gpurun_out/r2t_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2t_bench4.log 2> gpurun_out/r2t_bench4.err; echo "rc=$?" >> gpurun_out/r2t_bench4.err
python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench5.log 2> gpurun_out/r2t_bench5.err
for c in 1 2 3; do python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2t_bench$c.log 2> gpurun_out/r2t_bench$c.err; done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2t_ref.log 2> gpurun_out/r2t_ref.err
python bench.py --steps 10 --warmup 3 --jpeg --no-dropin --no-other-precision --no-cpu-baseline > gpurun_out/r2t_bench_jpeg.log 2> gpurun_out/r2t_bench_jpeg.err
for f in 4 5 1 2 3 _jpeg; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_bench$f.log").read().strip().splitlines()[-1])
print("$f", d["metric"], round(d["value"],3), round(d["e2e"]["value"],3), d["roofline"].get("frac"), d["roofline"].get("whole_step_tensor_frac"), d.get("clocks",{}).get("sm_mhz"), (d.get("other_precision") or {}).get("value"), (d.get("e2e_dropin") or {}).get("value"), (d.get("e2e_jpeg") or {}).get("value"))
PY
done
tail -1 gpurun_out/r2t_ref.log | cut -c1-400
du -sm gpurun_out
