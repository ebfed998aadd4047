"""One receipt through detection (and one crop through recognition) a few times: target of an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights

r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
img = receipts.receipt(0)
crop = np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)
for _ in range(3):
    r.detect([img])
for _ in range(3):
    r.recognize([crop], want_logits=False)
r.close()
