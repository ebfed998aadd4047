set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_reference_pipeline_gpu.py tests/test_serve_gpu.py tests/test_dropin_golden_gpu.py tests/test_edge_gpu.py -q -x > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log
tail -4 gpurun_out/r2w_tests.log
python tools/prof_pipeline.py > gpurun_out/r2w_prof.log 2>&1
LOCR_SORT_OVERLAP=0 python tools/prof_pipeline.py > gpurun_out/r2w_prof_nosort.log 2>&1
head -1 gpurun_out/r2w_prof.log gpurun_out/r2w_prof_nosort.log
python tools/prof_pipeline.py 8 3 Attention > gpurun_out/r2w_prof_attn.log 2>&1
head -4 gpurun_out/r2w_prof_attn.log; grep -i "attention" gpurun_out/r2w_prof_attn.log
B="python bench.py --steps 10 --warmup 3 --no-other-precision --no-dropin --no-cpu-baseline"
$B > gpurun_out/r2w_bench.log 2>gpurun_out/r2w_bench.err
LOCR_SORT_OVERLAP=0 $B > gpurun_out/r2w_bench_nosort.log 2>gpurun_out/r2w_bench_nosort.err
for f in bench bench_nosort; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2w_$f.log").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4), round(d["roofline"]["whole_step_tensor_frac"],4), round(d["roofline"]["kernel_share_of_step"],4))
PY
done
