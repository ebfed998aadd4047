"""ctypes binding of include/locr.h (liblocr.so).  No torch types cross this boundary: numpy / raw pointers only.

The library is built in-tree by lightly_ocr_b200/build.py; if it is missing it is built on first use, and if that is
impossible the import fails loudly (there is no CPU fallback for the product path).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblocr.so")

_lib = None


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "B", "H", "W", "Cin", "Cout", "KH", "KW", "dil_h", "dil_w", "pad_h", "pad_w", "stride_h",
        "x_pitch", "y_pitch", "relu", "out_fp32", "act_dtype", "n_tile")]


class LocrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        _lib = C.CDLL(LIB_PATH)
        _lib.locr_version.restype = C.c_char_p
        _lib.locr_last_error.restype = C.c_char_p
        _lib.locr_last_error.argtypes = [C.c_void_p]
        _lib.locr_test_conv.restype = C.c_int
        _lib.locr_test_conv.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    return _lib


def _check(rc, handle=None):
    if rc != 0:
        msg = lib().locr_last_error(handle)
        raise LocrError("liblocr error %d: %s" % (rc, (msg or b"").decode("utf-8", "replace")))


def _fptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def test_conv(x, w, bias=None, residual=None, *, dil=(1, 1), pad=(0, 0), stride_h=1, relu=False, out_fp32=False,
              act_dtype=1, n_tile=0, x_pitch=None, y_pitch=None):
    """x [B,H,W,x_pitch] fp32 NHWC (first Cin channels used), w [Cout,KH,KW,Cin] -> y [B,OH,OW,y_pitch] fp32."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    B, H, W, xp = x.shape
    Cout, KH, KW, Cin = w.shape
    if x_pitch is None:
        x_pitch = xp
    assert x_pitch == xp
    if y_pitch is None:
        y_pitch = Cout
    OH = (H + 2 * pad[0] - dil[0] * (KH - 1) - 1) // stride_h + 1
    OW = (W + 2 * pad[1] - dil[1] * (KW - 1) - 1) + 1
    d = ConvDesc(B, H, W, Cin, Cout, KH, KW, dil[0], dil[1], pad[0], pad[1], stride_h, x_pitch, y_pitch,
                 int(relu), int(out_fp32), act_dtype, n_tile)
    y = np.zeros((B, OH, OW, y_pitch), np.float32)
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32)
    if residual is not None:
        residual = np.ascontiguousarray(residual, np.float32)
    _check(lib().locr_test_conv(C.byref(d), _fptr(x), _fptr(w), _fptr(bias), _fptr(residual), _fptr(y)))
    return y
