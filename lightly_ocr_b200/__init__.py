"""B200-native detect-then-recognize path of lightly-ocr (CRAFT + CRNN) behind a C ABI.

`net` mirrors the reference's ocr/net.py (CRAFT, CRNN); `bridge` is the ctypes binding of include/locr.h.
"""
__all__ = ["bridge", "build"]
