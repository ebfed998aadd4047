"""Times the BiLSTM recurrence kernel alone (CUDA events inside liblocr) for a few batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge

for B in (81, 512, 650, 1300):
    rng = np.random.default_rng(0)
    xp = rng.normal(0, 1, (B, 26, 2048)).astype(np.float32)
    whh = (rng.uniform(-1, 1, (2, 1024, 256)) / 16).astype(np.float32)
    _, ms = bridge.test_lstm(xp, whh, 0, iters=20)
    print("B=%4d  %.1f us per launch  (%.2f us per step)" % (B, ms * 1e3, ms * 1e3 / 26))
