set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench4.log 2> gpurun_out/r2f_bench4.err; echo "rc=$?" >> gpurun_out/r2f_bench4.err
python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench5.log 2> gpurun_out/r2f_bench5.err
for c in 1 2 3; do python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2f_bench$c.log 2> gpurun_out/r2f_bench$c.err; done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_ref.log 2> gpurun_out/r2f_ref.err
python bench.py --steps 10 --warmup 3 --jpeg --no-dropin --no-other-precision --no-cpu-baseline > gpurun_out/r2f_bench_jpeg.log 2> gpurun_out/r2f_bench_jpeg.err
du -sm gpurun_out
