set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 --no-dropin --no-cpu-baseline > gpurun_out/r2u_n4.log 2> gpurun_out/r2u_n4.err; echo "rc=$?" >> gpurun_out/r2u_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2u_n4_ref.log 2> gpurun_out/r2u_n4_ref.err
tail -2 gpurun_out/r2u_n4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_n4.log").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["frac"], d["ranks"])
PY
