import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge
L = bridge.lib()
L.locr_bench_conv.restype = C.c_int
L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]
def run(name, B, H, W, Cin, Cout, k=3, n_tile=0, iters=10):
    pad = k // 2
    d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, pad, pad, 1, Cin, Cout, 1, 0, 0, n_tile)
    ms = C.c_float(); rc = L.locr_bench_conv(C.byref(d), iters, C.byref(ms))
    tiles = ((W + 63)//64) * ((H+1)//2) * B * max(1, Cout // (n_tile or min(Cout,256)))
    kb = k*k*max(1,Cin//64)
    cyc = ms.value*1e-3*1.9e9 / (tiles/148.0) / kb
    fl = 2.0*B*H*W*Cout*Cin*k*k
    print("%-26s %8.3f ms %7.1f TF/s  ~%6.0f cyc/kblock (kb=%d)" % (name, ms.value, fl/ms.value/1e9, cyc, kb))
B=2
run("3x3 64->64", B,1280,960,64,64)
run("3x3 64->128", B,1280,960,64,128)
run("3x3 64->256", B,1280,960,64,256)
run("1x1 64->64", B,1280,960,64,64,k=1)
run("1x1 512->64", B,640,480,512,64,k=1)
run("1x1 512->256", B,640,480,512,256,k=1)
run("3x3 128->128", B,640,480,128,128)
run("3x3 256->256", B,320,240,256,256)
run("3x3 256->256 nt128", B,320,240,256,256,n_tile=128)
run("3x3 256->256 nt64", B,320,240,256,256,n_tile=64)
