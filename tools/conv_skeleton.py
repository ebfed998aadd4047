"""Skeleton timings of conv_tc_kernel (experiments build): barrier chain only, with tcgen05.commit or a plain arrive."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LAYERS = (("slice1.0", 8, 1280, 960, 16, 64, 3), ("slice1.3", 8, 1280, 960, 64, 64, 3), ("slice1.7", 8, 640, 480, 64, 128, 3),
          ("slice1.10", 8, 640, 480, 128, 128, 3), ("cls.0", 8, 640, 480, 32, 32, 3), ("upconv4.0", 8, 640, 480, 192, 64, 1),
          ("upconv4.3", 8, 640, 480, 64, 32, 3), ("conv0_2", 640, 32, 100, 32, 64, 3), ("slice3.27", 8, 160, 120, 512, 512, 3))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from lightly_ocr_b200 import bridge
    L = bridge.lib()
    L.locr_bench_conv.restype = C.c_int
    L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]
    for name, B, H, W, Cin, Cout, k in LAYERS:
        d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, k // 2, k // 2, 1, Cin, Cout, 1, 0, 0, 0)
        ms = C.c_float()
        rc = L.locr_bench_conv(C.byref(d), 10, C.byref(ms))
        print("  %-10s %8.3f ms rc=%d" % (name, ms.value, rc), flush=True)
else:
    for dbg, what in ((0, "full kernel"), (31, "barrier chain only"), (95, "barrier chain, plain arrive instead of commit"),
                      (16, "no epilogue"), (7, "epilogue + stores only")):
        print("LOCR_CONV_DBG=%d (%s)" % (dbg, what), flush=True)
        env = dict(os.environ, LOCR_CONV_DBG=str(dbg), LOCR_CONV_HALO="0")
        if dbg == 0:
            env["LOCR_CONV_VERBOSE"] = "1"
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
