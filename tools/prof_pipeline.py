"""Per-kernel time table of the whole pipeline (CUDA events around every launch, single lane, in-pipeline): which layers
are far from the roofline.  Usage: python tools/prof_pipeline.py [receipts_per_batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
head = sys.argv[3] if len(sys.argv) > 3 else "CTC"
r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head=head)
r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, head))
batch = [receipts.receipt(i) for i in range(n)]
for _ in range(2):
    _, out = r.ocr(batch)
crops = len(out["text"])
r.profile(True)
r.profile_read()
r.profile_layers()
r.timer_start()
for _ in range(steps):
    r.ocr_resident(n)
total = r.timer_stop()
r.profile_read()
rows = r.profile_layers()
r.profile(False)
ksum = sum(x[1] for x in rows)
print("%d receipts x %d steps, %d crops per step: %.3f ms per step (kernels %.3f ms)" % (n, steps, crops, total / steps, ksum / steps))
for name, ms, fl, cnt in sorted(rows, key=lambda x: -x[1]):
    print("%-55s %8.3f ms/step %5.1f%%  n=%-3d %s" % (name, ms / steps, 100 * ms / ksum, cnt // steps,
                                                     ("%7.1f TF/s" % (fl / ms / 1e9)) if fl > 0 else ""))
