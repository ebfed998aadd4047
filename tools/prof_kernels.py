"""Targets of the `ncu --set full` captures kept under profiles/ (one launch per kernel of interest).

  python tools/prof_kernels.py tc    representative tcgen05 conv layers (one launch each with LOCR_BENCH_WARMUP=0) and the
                                     BiLSTM cluster kernel at 650 crops
  python tools/prof_kernels.py mem   one detect+recognize pass over 8 synthetic receipts: the HBM-bound kernels
                                     (pre-processing, pooling, up-sampling, labelling, boxes, crops, TPS, decode)
  python tools/prof_kernels.py memi  the same plus a second pass with the receipts handed over as PNG / JPEG files
                                     (the ingest kernels: png_unfilter, png_color, jpeg_idct, jpeg_color)
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lightly_ocr_b200 import bridge

mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
if mode == "tc":
    L = bridge.lib()
    L.locr_bench_conv.restype = C.c_int
    L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]

    def run(name, B, H, W, Cin, Cout, k=3, pool=0):
        os.environ["LOCR_BENCH_POOL"] = str(pool)     # 2: only the fused 2x2 max-pooled tensor is written (as in the pipeline)
        pad = k // 2
        d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, pad, pad, 1, Cin, Cout, 1, 0, 0, 0)
        ms = C.c_float()
        rc = L.locr_bench_conv(C.byref(d), 1, C.byref(ms))
        fl = 2.0 * B * H * W * Cout * Cin * k * k
        print("%-12s %8.3f ms %8.1f TF/s rc=%d" % (name, ms.value, fl / ms.value / 1e9, rc))

    run("slice1.0", 8, 1280, 960, 16, 64)
    run("slice1.3", 8, 1280, 960, 64, 64, pool=2)
    run("slice1.10", 8, 640, 480, 128, 128)
    run("slice3.27", 8, 160, 120, 512, 512)
    run("cls.0", 8, 640, 480, 32, 32)
    run("upconv4.3", 8, 640, 480, 64, 32)
    run("slice2.17", 8, 320, 240, 256, 256)
    run("crnn512", 640, 4, 26, 512, 512)
    rng = np.random.default_rng(0)
    xp = rng.normal(0, 1, (650, 26, 2048)).astype(np.float32)
    whh = (rng.uniform(-1, 1, (2, 1024, 256)) / 16).astype(np.float32)
    bridge.test_lstm(xp, whh, 0)
    print("lstm ok")
else:
    from lightly_ocr_b200.synth import receipts, weights
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
    imgs = [receipts.receipt(i) for i in range(8)]
    per_image, out = r.ocr(imgs)
    print("%d crops, first strings %s" % (len(out["text"]), out["text"][:5]))
    # the ingest kernels: the same receipts as PNG and JPEG files through locr_detect_encoded
    import cv2
    if mode == "memi":
        r.ocr_encoded([cv2.imencode(".png", im)[1].tobytes() for im in imgs[:4]] +
                      [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in imgs[4:]])
