mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_gpu.py -q -x -m gpu > gpurun_out/r4e_lstm.log 2>&1; echo "tests rc=$?" >> gpurun_out/r4e_lstm.log
tail -12 gpurun_out/r4e_lstm.log
echo "--- push form"; timeout 120 python tools/bench_lstm.py 2>&1 | tail -5
echo "--- pull form"; LOCR_LSTM_MCAST=0 timeout 120 python tools/bench_lstm.py 2>&1 | tail -5
