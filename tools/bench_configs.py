"""Timings of the five BASELINE.json configs on one B200 (single lane, CUDA events on the handle's stream around whole
calls, host buffers in / host results out).  The headline metric (config 4) is bench.py's job; this fills the table in
DESIGN.md for the others."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights


def timed(r, fn, iters):
    fn()
    fn()
    r.timer_start()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    ms = r.timer_stop()
    return max(ms, 1e3 * (time.perf_counter() - t0)) / iters


def main():
    craft = weights.craft_calibrated(0, ink=True)
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    r.load_state_dict(bridge.MODEL_CRAFT, craft)
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
    # config 1: one 32x100 gray crop, CTC
    crop = np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)
    ms = timed(r, lambda: r.recognize([crop], want_logits=False), 20)
    print("config 1  CRNN (CTC) on one 32x100 crop:                 %8.3f ms per call" % ms)
    # config 2: CRAFT forward + boxes on one 1280x960 receipt
    img = receipts.receipt(0)
    ms = timed(r, lambda: r.detect([img]), 20)
    print("config 2  CRAFT + getDetBoxes, one 1280x960 receipt:     %8.3f ms per image (%d boxes)" %
          (ms, len(r.detect([img])[0][0])))
    # config 3: 512 ragged crops
    crops = receipts.crops(512, seed=3)
    ms = timed(r, lambda: r.recognize(crops, want_logits=False), 10)
    print("config 3  CRNN on 512 ragged crops (TPS+BiLSTM+CTC):     %8.3f ms per batch = %.0f crops/s" % (ms, 512e3 / ms))
    # config 4 (single lane, for reference): 8 receipts per call
    imgs = [receipts.receipt(i) for i in range(8)]
    ms = timed(r, lambda: r.ocr(imgs), 10)
    n = len(r.ocr(imgs)[1]["text"])
    print("config 4  end to end, 8 receipts per call, one lane:     %8.3f ms per call = %.0f receipts/s, %.0f crops/s" %
          (ms, 8e3 / ms, n * 1e3 / ms))
    r.close()
    # config 5: attention decoder end to end
    a = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="Attention")
    a.load_state_dict(bridge.MODEL_CRAFT, craft)
    a.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "Attention"))
    ms = timed(a, lambda: a.ocr(imgs), 10)
    print("config 5  attention decoder end to end, 8 receipts/call: %8.3f ms per call = %.0f receipts/s" % (ms, 8e3 / ms))
    a.close()


if __name__ == "__main__":
    main()
