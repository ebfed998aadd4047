"""Runs a few representative conv layers through locr_bench_conv (1 timed launch after 3 warm-ups each): the target of
the `ncu --set full` captures kept under profiles/."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge

L = bridge.lib()
L.locr_bench_conv.restype = C.c_int
L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]


def run(name, B, H, W, Cin, Cout, k=3):
    pad = k // 2
    d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, pad, pad, 1, Cin, Cout, 1, 0, 0, 0)
    ms = C.c_float()
    rc = L.locr_bench_conv(C.byref(d), 1, C.byref(ms))
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    print("%-12s %8.3f ms %8.1f TF/s rc=%d" % (name, ms.value, fl / ms.value / 1e9, rc))


run("slice1.0", 2, 1280, 960, 16, 64)
run("slice1.3", 2, 1280, 960, 64, 64)
run("cls.0", 2, 640, 480, 32, 32)
run("slice3.27", 2, 160, 120, 512, 512)
run("crnn512", 512, 4, 26, 512, 512)
