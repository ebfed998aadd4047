"""Top-level `net` module for the reference's `from net import CRAFT, CRNN` (ocr/pipeline.py:9, ocr/torch2onnx.py:12).

Put this directory ahead of the reference's ocr/ on sys.path (and the repository root on sys.path as well):
    PYTHONPATH=<repo>/lightly_ocr_b200/dropin:<repo>:<reference>/ocr  python <reference>/ocr/pipeline.py --img x.png
"""
from lightly_ocr_b200.net import (CONFIG, CRAFT, CRNN, DEVICE, MODEL_PATH,  # noqa: F401
                                  Placeholder, copyStateDict)
