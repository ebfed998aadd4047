"""Per-layer timing of the tcgen05 conv kernel on the CRAFT / CRNN layer shapes (run on the GPU box)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge

L = bridge.lib()
L.locr_bench_conv.restype = C.c_int
L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]

def run(name, B, H, W, Cin, Cout, k=3, dil=1, n_tile=0, iters=20):
    pad = dil * (k // 2)
    d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, dil, dil, pad, pad, 1, Cin, Cout, 1, 0, 1, n_tile)
    ms = C.c_float()
    rc = L.locr_bench_conv(C.byref(d), iters, C.byref(ms))
    if rc != 0:
        print(name, "ERR", L.locr_last_error(None)); return 0.0
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    print("%-28s B%-3d %4dx%-4d %4d->%-4d k%d nt%-3d %8.3f ms %8.1f TF/s" % (name, B, H, W, Cin, Cout, k, n_tile, ms.value, fl / ms.value / 1e9))
    return ms.value

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tot = 0
tot += run("slice1.3", B, 1280, 960, 64, 64)
tot += run("slice1.7", B, 640, 480, 64, 128)
tot += run("slice1.10", B, 640, 480, 128, 128)
tot += run("slice2.14", B, 320, 240, 128, 256)
tot += 2 * run("slice2.17/3.20", B, 320, 240, 256, 256)
tot += run("slice3.24", B, 160, 120, 256, 512)
tot += 2 * run("slice3.27/4.30", B, 160, 120, 512, 512)
tot += 2 * run("slice4.34/37", B, 80, 60, 512, 512)
tot += run("slice5.1 dil6", B, 80, 60, 512, 1024, dil=6)
tot += run("slice5.2 1x1", B, 80, 60, 1024, 1024, k=1)
tot += run("upconv1.0", B, 80, 60, 1536, 512, k=1)
tot += run("upconv1.3", B, 80, 60, 512, 256)
tot += run("upconv2.0", B, 160, 120, 768, 256, k=1)
tot += run("upconv2.3", B, 160, 120, 256, 128)
tot += run("upconv3.0", B, 320, 240, 384, 128, k=1)
tot += run("upconv3.3", B, 320, 240, 128, 64)
tot += run("upconv4.0", B, 640, 480, 192, 64, k=1)
tot += run("upconv4.3", B, 640, 480, 64, 32)
tot += 2 * run("cls.0/2", B, 640, 480, 32, 32)
tot += run("cls.4", B, 640, 480, 32, 16)
print("CRAFT TC-conv total %.3f ms per %d image(s) -> %.1f TF/s" % (tot, B, 874.2 * B / tot))
for nt in (0, 128):
    run("crnn 512 4x26 b64", 64, 4, 26, 512, 512, n_tile=nt)
    run("crnn 512 4x26 b512", 512, 4, 26, 512, 512, n_tile=nt)
run("crnn 256 8x25 b512", 512, 8, 25, 256, 256)
run("crnn 128 16x50 b512", 512, 16, 50, 128, 128)
for nt in (64, 128, 256):
    run("sweep 256ch 320x240", 1, 320, 240, 256, 256, n_tile=nt)
