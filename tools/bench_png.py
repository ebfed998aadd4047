"""PNG ingest timing on the GPU box: cv2.imdecode on the host against the fused locr_detect_encoded path (zlib inflate on
host threads, un-filtering + sample conversion in CUDA).  Usage: python tools/bench_png.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rng = np.random.default_rng(0)
for label, noise in (("clean synthetic receipts", 0), ("receipts with sensor-like noise", 6)):
    blobs = []
    for i in range(n):
        img = receipts.receipt(i)
        if noise:
            img = np.clip(img.astype(np.int16) + rng.integers(-noise, noise + 1, img.shape), 0, 255).astype(np.uint8)
        blobs.append(cv2.imencode(".png", img)[1].tobytes())
    print("%s: %d files of 1280x960, %.0f KB per file" % (label, n, np.mean([len(b) for b in blobs]) / 1e3))
    arrs = [np.frombuffer(b, np.uint8) for b in blobs]
    for _ in range(2):
        t = time.perf_counter()
        dec = [cv2.imdecode(a, cv2.IMREAD_COLOR) for a in arrs]
        t_cv = time.perf_counter() - t
    print("  cv2.imdecode (host, serial): %.2f ms per image" % (t_cv / n * 1e3))
    t = time.perf_counter()
    for b in blobs:
        bridge.png_scanlines(b)
    print("  liblocr chunk walk + CRC + inflate alone (1 thread, incl. ctypes): %.2f ms per image" % ((time.perf_counter() - t) / n * 1e3))
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
    for _ in range(2):
        r.ocr_encoded(blobs)
        r.ocr(dec)
    reps = 5
    t = time.perf_counter()
    for _ in range(reps):
        r.ocr_encoded(blobs)
    t_enc = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        r.ocr([cv2.imdecode(a, cv2.IMREAD_COLOR) for a in arrs])
    t_cvocr = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        r.ocr(dec)
    t_ocr = (time.perf_counter() - t) / reps
    print("  one lane, %d files per call: ocr(decoded pixels) %.1f ms | cv2.imdecode + ocr %.1f ms | ocr_encoded (GPU decode) %.1f ms"
          % (n, t_ocr * 1e3, t_cvocr * 1e3, t_enc * 1e3))
    r.profile(True); r.profile_read(); r.profile_layers()
    r.ocr_encoded(blobs)
    r.profile_read()
    rows = {name: ms for name, ms, fl, cnt in r.profile_layers()}
    print("  GPU kernels per %d files: png_unfilter %.3f ms, png_color %.3f ms" % (n, rows.get("png_unfilter", 0), rows.get("png_color", 0)))
    r.close()
