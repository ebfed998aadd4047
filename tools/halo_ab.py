"""A/B of the haloed-patch path of conv_tc (3x3 / 64 -> 64): LOCR_CONV_HALO=0|1, with and without the fused pool."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from lightly_ocr_b200 import bridge
    L = bridge.lib()
    L.locr_bench_conv.restype = C.c_int
    L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]
    for name, B, H, W, Cin, Cout, k in (("slice1.3", 8, 1280, 960, 64, 64, 3), ("slice1.3 b1", 1, 1280, 960, 64, 64, 3)):
        d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, 1, 1, 1, Cin, Cout, 1, 0, 0, 0)
        ms = C.c_float()
        rc = L.locr_bench_conv(C.byref(d), 20, C.byref(ms))
        print("  %-12s %8.3f ms rc=%d %s" % (name, ms.value, rc, L.locr_last_error(None) if rc else ""), flush=True)
else:
    for halo in (0, 1):
        for pool in (0, 1, 2):
            print("LOCR_CONV_HALO=%d LOCR_BENCH_POOL=%d" % (halo, pool), flush=True)
            env = dict(os.environ, LOCR_CONV_HALO=str(halo), LOCR_BENCH_POOL=str(pool))
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
