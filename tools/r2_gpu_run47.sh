mkdir -p gpurun_out
timeout 300 python tools/prof_b1.py 20 > gpurun_out/r4h_b1.log 2>&1; echo rc=$?
grep "^==\|loc_head\|decode\|lstm \|tps\|crop_resize" gpurun_out/r4h_b1.log
timeout 200 python -m pytest tests/test_nets_gpu.py -q -x -m gpu -k "tps or loc or fid or crnn" 2>&1 | tail -3
