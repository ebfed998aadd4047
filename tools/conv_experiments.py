"""Which resource bounds a conv layer?  Needs liblocr built with LOCR_NVCC_EXTRA=-DLOCR_CONV_EXPERIMENTS=1; re-runs
one layer with parts of the kernel switched off through LOCR_CONV_DBG (results are garbage, only the time matters)."""
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
    for name, B, H, W, Cin, Cout, k in (("slice1.0", 8, 1280, 960, 16, 64, 3), ("slice1.3", 8, 1280, 960, 64, 64, 3),
                                        ("slice1.10", 8, 640, 480, 128, 128, 3), ("cls.0", 8, 640, 480, 32, 32, 3),
                                        ("upconv4.0", 8, 640, 480, 192, 64, 1), ("slice3.27", 8, 160, 120, 512, 512, 3)):
        pad = k // 2
        d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, pad, pad, 1, Cin, Cout, 1, 0, 0, 0)
        ms = C.c_float()
        rc = L.locr_bench_conv(C.byref(d), 10, C.byref(ms))
        print("  %-10s %8.3f ms rc=%d" % (name, ms.value, rc))
else:
    for dbg, what in ((0, "full kernel"), (1, "no MMAs"), (2, "no A loads"), (4, "no B loads"), (8, "no output stores"),
                      (11, "no MMAs, no A loads, no stores"), (16, "no epilogue"), (27, "no MMAs, A loads, stores, epilogue"),
                      (31, "barrier chain only")):
        print("LOCR_CONV_DBG=%d (%s)" % (dbg, what), flush=True)
        env = dict(os.environ, LOCR_CONV_DBG=str(dbg))
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
