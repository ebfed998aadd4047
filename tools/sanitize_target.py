"""Small end-to-end workload for compute-sanitizer (memcheck): a few conv shapes incl. fused pool, the LSTM cluster
kernel, and one detect + recognise pass (CTC and Attention) on a 320x256 receipt window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights

rng = np.random.default_rng(0)
x = rng.standard_normal((2, 24, 40, 64)).astype(np.float32)
w = (rng.standard_normal((64, 3, 3, 64)) / 24).astype(np.float32)
b = rng.standard_normal(64).astype(np.float32)
bridge.test_conv(x, w, b, None, pad=(1, 1), relu=True, out_fp32=False, act_dtype=0)
bridge.test_conv_pool(x, w, b, pad=(1, 1), relu=True, act_dtype=0, want_full=False)
x2 = rng.standard_normal((3, 4, 26, 128)).astype(np.float32)
w2 = (rng.standard_normal((256, 3, 3, 128)) / 34).astype(np.float32)
bridge.test_conv(x2, w2, None, rng.standard_normal((3, 4, 26, 256)).astype(np.float32), pad=(1, 1), relu=True, act_dtype=0)
xp = rng.normal(0, 1, (130, 26, 2048)).astype(np.float32)
whh = (rng.uniform(-1, 1, (2, 1024, 256)) / 16).astype(np.float32)
bridge.test_lstm(xp, whh, 0)
img = np.ascontiguousarray(receipts.receipt(0)[40:360, 40:296])
for head in ("CTC", "Attention"):
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head=head)
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, head))
    per_image, out = r.ocr([img, img])
    print(head, len(out["text"]), out["text"][:4])
    r.close()
print("done")
