"""Per-step timeline of the BiLSTM cluster kernel (needs liblocr built with LOCR_NVCC_EXTRA=-DLOCR_LSTM_TRACE=1)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge
rng = np.random.default_rng(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 650
xp = rng.normal(0, 1, (B, 26, 2048)).astype(np.float32)
whh = (rng.uniform(-1, 1, (2, 1024, 256)) / 16).astype(np.float32)
bridge.test_lstm(xp, whh, 0)
bridge.test_lstm(xp, whh, 0)
L = bridge.lib()
buf = (C.c_longlong * (64 * 8))()
assert L.locr_debug_lstm_trace(buf) == 0
t = np.array(buf, np.int64).reshape(64, 8)[:26]
names = ["hready seen (producer)", "chunk 0 landed (MMA warp)", "chunk 3 landed + MMAs issued", "tfull seen (epilogue)",
         "gate math done", "h stored + peers signalled"]
print("cycles relative to 'h stored' of the previous step (mean over steps 2..25):")
for s in range(2, 26):
    pass
prev = t[1:25, 5]
for i, n in enumerate(names):
    d = t[2:26, i] - prev
    print("  %-34s %8.0f  (min %d max %d)" % (n, d.mean(), d.min(), d.max()))
print("step period: %.0f cycles" % np.diff(t[2:26, 5]).mean())
