set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.log 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench4.log 2> gpurun_out/r2a_bench4.err; echo "rc=$?" >> gpurun_out/r2a_bench4.err
python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench5.log 2> gpurun_out/r2a_bench5.err
for c in 1 2 3; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2a_bench$c.log 2> gpurun_out/r2a_bench$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.log 2> gpurun_out/r2a_ref.err
# plain-fp32 synthetic checkpoint (CTC, then the attention decoder on top of it)
LOCR_TRAIN_MODE=fp32 timeout 1500 python tools/train_synth_crnn.py 2000 9000 > gpurun_out/r2a_train_ctc.log 2>&1
cp gpurun_out/calib_crnn_ctc_fp32.npz lightly_ocr_b200/synth/ && \
LOCR_TRAIN_MODE=fp32 LOCR_TRAIN_HEAD=Attention timeout 600 python tools/train_synth_crnn.py 2000 6000 > gpurun_out/r2a_train_attn.log 2>&1
ls -la gpurun_out/*.npz
