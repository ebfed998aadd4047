"""cProfile of the drop-in's literal path (net.CRAFT.process + one net.CRNN.process per crop, ocr/pipeline.py:70-79) over
a few receipts: where the host time of the one-image-at-a-time path goes.  python tools/prof_dropin.py [receipts]"""
import contextlib
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import torch
import yaml
from lightly_ocr_b200.synth import receipts, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tmp = tempfile.mkdtemp(prefix="locr_prof_")
dst = os.path.join(tmp, "ocr")
os.makedirs(os.path.join(dst, "save_models"))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(root, "lightly_ocr_b200", "config.yml")))
cfg["prediction"], cfg["num_classes"] = "CTC", 37
yaml.safe_dump(cfg, open(os.path.join(dst, "config.yml"), "w"))
torch.save(weights.craft_calibrated(0, ink=True), os.path.join(dst, "save_models", "CRAFT.pth"))
torch.save(weights.crnn_calibrated(1, "CTC"), os.path.join(dst, "save_models", "CRNN.pth"))
os.environ["LOCR_OCR_DIR"] = dst
import lightly_ocr_b200.net as net

det, rec = net.CRAFT(), net.CRNN()
images = [receipts.receipt(3000 + i) for i in range(n)]


def get_text(image):
    res = {}
    for img in det.process(image):
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        _, res = rec.process(res, gray)
    return res


sink = io.StringIO()
with contextlib.redirect_stdout(sink):
    for im in images[:2]:
        get_text(im)
    t0 = time.perf_counter()
    for im in images:
        get_text(im)
    dt = time.perf_counter() - t0
    pr = cProfile.Profile()
    pr.enable()
    for im in images:
        get_text(im)
    pr.disable()
print("%.2f ms per receipt without the profiler" % (1e3 * dt / n))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue())
