set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_n2.log 2> gpurun_out/r2h_n2.err; echo "rc=$?" >> gpurun_out/r2h_n2.err
LOCR_BENCH_SAMPLER=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-other-precision > gpurun_out/r2h_n2_nosampler.log 2> gpurun_out/r2h_n2_nosampler.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2h_n2_ref.log 2> gpurun_out/r2h_n2_ref.err
tail -2 gpurun_out/r2h_n2.err
