import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge
L = bridge.lib()
L.locr_bench_conv.restype = C.c_int
L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]
def run(name, B, H, W, Cin, Cout, k=3):
    pad = k // 2
    d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, pad, pad, 1, Cin, Cout, 1, 0, 0, 0)
    ms = C.c_float(); L.locr_bench_conv(C.byref(d), 1, C.byref(ms))
    print(name, ms.value, flush=True)
run("3x3 64->64", 2,1280,960,64,64)
run("3x3 256->256", 2,320,240,256,256)
run("crnn512", 512,4,26,512,512)
