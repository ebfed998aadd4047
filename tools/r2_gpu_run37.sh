set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 150 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline > gpurun_out/r3f_soak.log 2> gpurun_out/r3f_soak.err; echo "rc=$?" >> gpurun_out/r3f_soak.err
tail -2 gpurun_out/r3f_soak.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3f_soak.log").read().strip().splitlines()[-1])
print("soak", d["steps"], round(d["value"],1), round(d["e2e"]["value"],1), d["clocks"])
PY
for i in 1 2 3; do timeout 300 python -m pytest tests/test_conv_gpu.py -q -x -k "cta2 or hstream or halo or tail" > gpurun_out/r3f_conv$i.log 2>&1; tail -1 gpurun_out/r3f_conv$i.log; done
timeout 600 python bench.py --config 5 --precision exact --steps 30 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline > gpurun_out/r3f_soak5.log 2> gpurun_out/r3f_soak5.err; echo "rc=$?" >> gpurun_out/r3f_soak5.err; tail -1 gpurun_out/r3f_soak5.err
