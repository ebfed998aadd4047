"""Per-launch CUDA-event times of one receipt (BASELINE config 2) and of one crop (config 1) on one lane: where the
single-unit latency goes.  python tools/prof_b1.py [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
img = receipts.receipt(0)
crop = np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)


def table(title, fn, res_fn=None):
    for _ in range(5):
        fn()
    f = res_fn or fn
    for _ in range(3):
        f()
    r.timer_start()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    dev = r.timer_stop() / reps
    wall = 1e3 * (time.perf_counter() - t0) / reps
    r.profile(True)
    r.profile_read()
    r.profile_layers()
    for _ in range(reps):
        f()
    r.profile_read()
    rows = r.profile_layers()
    r.profile(False)
    tot = sum(x[1] for x in rows) / reps
    print("== %s: %.3f ms device, %.3f ms wall per call; %.3f ms summed over %d launches" %
          (title, dev, wall, tot, sum(x[3] for x in rows) // reps))
    for name, ms, fl, n in rows:
        print("  %-34s %8.1f us  x%d  %8.1f TFLOP/s" % (name, 1e3 * ms / reps, n // reps, fl / ms / 1e9 if ms > 0 else 0.0))


table("config 2: one 1280x960 receipt, CRAFT + boxes (resident)", lambda: r.detect([img]), lambda: r.detect_resident(1))
table("config 1: one 32x100 crop, CRNN CTC", lambda: r.recognize([crop], want_logits=False))
r.close()
