"""In-kernel timeline of conv_tc_kernel (block 0): where a tile's time goes in the producer, the MMA issuer and the
epilogue.  Needs liblocr built with LOCR_NVCC_EXTRA=-DLOCR_CONV_EXPERIMENTS=1.
Usage: python tools/conv_trace.py            (runs the layer list, one child process per layer / setting)"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

LAYERS = {
    "slice1.0": (8, 1280, 960, 16, 64, 3), "slice1.3": (8, 1280, 960, 64, 64, 3), "slice1.7": (8, 640, 480, 64, 128, 3),
    "slice1.10": (8, 640, 480, 128, 128, 3), "cls.0": (8, 640, 480, 32, 32, 3), "upconv4.0": (8, 640, 480, 192, 64, 1),
    "upconv4.3": (8, 640, 480, 64, 32, 3), "slice3.27": (8, 160, 120, 512, 512, 3),
}
EV = {0: {0: "tile", 1: "empty", 2: "issued"}, 1: {0: "tile", 1: "tempty", 2: "full", 3: "commit", 4: "tfull"},
      2: {0: "tile", 1: "tfull", 2: "bar0", 3: "ld", 4: "sts", 5: "pool", 6: "fence", 7: "bar1", 8: "store"}}


def child(name):
    from lightly_ocr_b200 import bridge
    L = bridge.lib()
    L.locr_bench_conv.restype = C.c_int
    L.locr_bench_conv.argtypes = [C.POINTER(bridge.ConvDesc), C.c_int, C.POINTER(C.c_float)]
    B, H, W, Cin, Cout, k = LAYERS[name]
    d = bridge.ConvDesc(B, H, W, Cin, Cout, k, k, 1, 1, k // 2, k // 2, 1, Cin, Cout, 1, 0, 0, 0)
    ms = C.c_float()
    buf = np.zeros((3, 8192), np.uint64)
    cnt = np.zeros(3, np.int32)
    rc = L.locr_bench_conv(C.byref(d), 1, C.byref(ms))          # LOCR_BENCH_WARMUP=0: exactly one launch
    L.locr_conv_trace(buf.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    print("%s: %.3f ms (single cold launch) rc=%d, trace entries %s" % (name, ms.value, rc, cnt.tolist()))
    for role, rname in ((0, "producer"), (1, "mma"), (2, "epilogue")):
        n = int(cnt[role])
        if n < 8:
            continue
        ev = (buf[role, :n] & 15).astype(np.int64)
        clk = (buf[role, :n] >> 4).astype(np.int64)
        starts = np.nonzero(ev == 0)[0]
        if len(starts) < 12:
            continue
        # steady state: tiles 8 .. last-2 of the recorded window
        lo, hi = starts[8], starts[-2]
        period = np.diff(clk[starts[8:-1]])
        print("  %-9s tiles recorded %d, tile period median %d cycles (p10 %d, p90 %d)" %
              (rname, len(starts), np.median(period), np.percentile(period, 10), np.percentile(period, 90)))
        # time attributed to the step that ENDS at each event (delta to the previous stamp), summed per tile
        dt = np.diff(clk[lo:hi + 1])
        e = ev[lo + 1:hi + 1]
        ntiles = int((ev[lo:hi] == 0).sum())
        for code in sorted(EV[role]):
            m = e == code
            if m.any():
                print("      -> %-7s %7.0f cycles/tile (%d stamps/tile, median %d each)" %
                      (EV[role][code], dt[m].sum() / ntiles, round(m.sum() / ntiles), np.median(dt[m])))


if len(sys.argv) > 2 and sys.argv[1] == "child":
    child(sys.argv[2])
else:
    runs = [("slice1.0", {}), ("slice1.0", {"LOCR_CONV_DBG": "63"}), ("slice1.3", {"LOCR_CONV_HALO": "0"}),
            ("slice1.3", {"LOCR_BENCH_POOL": "2"}), ("cls.0", {}), ("cls.0", {"LOCR_CONV_DBG": "63"}),
            ("upconv4.3", {}), ("upconv4.0", {}), ("slice1.7", {}), ("slice1.10", {}), ("slice3.27", {})]
    for name, extra in runs:
        env = dict(os.environ, LOCR_BENCH_WARMUP="0")
        env.update(extra)
        env["LOCR_CONV_DBG"] = str(int(env.get("LOCR_CONV_DBG", "0")) | 32)
        print("== %s %s" % (name, extra), flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", name], env=env)
