#!/usr/bin/env python
"""Throughput of the detect-then-recognize path (BASELINE.json metric: end-to-end receipts/sec at 1280 px,
CRAFT + CRNN, 1/2/4/8 B200; crops/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--head CTC|Attention]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default = BASELINE config 4.  One step = one pass of the whole path (CRAFT forward, thresholds + labelling + boxes, host
reading-order sort, GPU crops + BICUBIC, CRNN forward, CTC decode) over RECEIPTS_PER_STEP distinct synthetic 1280x960
receipts per GPU, issued as PASSES rounds of LANES x PER_LANE receipts (weak scaling: per-GPU work is fixed).  `value`
times the path with the receipts already resident in HBM; `e2e` times the same path through the C ABI with host buffers
(host->device copy of the images and device->host copy of rects, strings and confidences inside the timed region).
Receipts are independent units: no collective on the data path, torch.distributed only provides the barrier and the
max-over-ranks reduction of the elapsed time.

--head Attention (or --config 5) runs the same legs with the attention decoder (BASELINE config 5).  Every run of
config 4 also times the e2e leg once more with the attention decoder (`other_head`; the CTC head when the main legs are
config 5), at whatever N it was launched with, so a 1 / 2 / 4 / 8-GPU sweep of the default command carries both heads.
--config 1 / 2 / 3 time the other BASELINE configurations on one lane and print the same JSON shape:
  1  CRNN (CTC) on a single 32x100 gray crop          (ms per crop, lower is better)
  2  CRAFT forward + getDetBoxes on one 1280x960 receipt (ms per receipt, lower is better)
  3  CRNN on a batch of 512 ragged crops              (crops/s)
--impl reference times the reference's own CPU implementation of the chosen configuration: the UNMODIFIED reference
from baseline/_ref/ocr (oracle/ref_env.py) when it is installed, else the oracle port.
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LANES = int(os.environ.get("LOCR_BENCH_LANES", "3"))   # host threads per GPU, each with its own liblocr handle/stream:
                               # while one lane sorts rects / copies results on the host, the other lanes' kernels
                               # keep the GPU busy
PER_LANE = int(os.environ.get("LOCR_BENCH_PER_LANE", "8"))   # receipts per lane and pass (one CRAFT launch sequence)
PASSES = int(os.environ.get("LOCR_BENCH_PASSES", "4"))       # passes per step: a 20-step leg lasts > 2 s
PER_PASS = PER_LANE * LANES
RECEIPTS_PER_STEP = PER_PASS * PASSES
POOL = 2 * PER_PASS            # distinct receipts cycled through (3.7 MB of pixels each)
UNIT = "receipts/s"
CRAFT_FLOPS = 874.217e9        # per 1280x960 canvas
CRNN_FLOPS_PER_CROP = 10.593e9


def metric_name(head):
    return "receipts_per_sec_1280px_craft_crnn_%s" % ("ctc" if head == "CTC" else "attention")


def conv_traffic():
    """DRAM bytes per conv_tc launch (mean over one 8-receipt pass) from the newest committed ncu capture, or None."""
    try:
        import glob
        path = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_conv_traffic.json")))[-1]
        with open(path) as f:
            t = json.load(f)
        return float(t["bytes_per_launch"]), "profiles/%s: %s" % (os.path.basename(path), t["source"])
    except Exception:
        return None, None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe).  One sampler
    runs from before the first warm-up to after the last leg (nvidia-smi needs a few hundred ms to produce its first
    line); window(t0, t1) summarises the samples whose arrival time falls inside a timed window.  LOCR_BENCH_SAMPLER=0
    switches it off (used once to show that the sampler does not slow its rank: DESIGN.md 7)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        if os.environ.get("LOCR_BENCH_SAMPLER", "1") == "0":
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off or nvidia-smi unavailable"]}
        rows = [r for (t, r) in list(self.rows) if t0 <= t <= t1 + 0.05]
        window = "timed region"
        if not rows:
            rows = [r for (_, r) in list(self.rows)]
            window = "whole run (no sample fell inside the timed region)"
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "reasons": reasons, "samples": len(sm),
                "window": window}


def make_receipts(rank, count):
    """Decoded BGR receipts in PINNED host memory (the e2e leg copies them to the device every pass)."""
    import torch
    from lightly_ocr_b200.synth import receipts
    out = []
    for i in range(count):
        t = torch.from_numpy(receipts.receipt(1000 * rank + i)).pin_memory()
        out.append(t.numpy())
    make_receipts.keep = getattr(make_receipts, "keep", []) + out     # the arrays view the pinned tensors' storage
    return out


# ------------------------------------------------------------------------------------------------ the CPU reference
class CpuReference:
    """The reference's CPU implementation of the path: the unmodified reference package when baseline/_ref/ocr is
    installed (kind "reference"), else the oracle port (kind "port").  Same synthetic checkpoints either way."""

    def __init__(self, head, n_threads):
        import torch
        from oracle import ref_env, weights
        torch.set_num_threads(n_threads)
        self.head = head
        self.craft_sd = weights.craft_calibrated(0, ink=True)
        self.crnn_sd = weights.crnn_calibrated(1, head)
        self.kind = "port"
        self._stack = contextlib.ExitStack()
        if ref_env.source() is not None:
            try:
                self.tmp = tempfile.mkdtemp(prefix="locr_ref_")
                dst = ref_env.stage(head, self.craft_sd, self.crnn_sd, self.tmp)
                self._stack.enter_context(ref_env.imported(dst))
                import pipeline
                with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                    self.detector, self.recognizer = pipeline.prepModel(pipeline.CONFIG, docker=True)
                self.pipeline = pipeline
                self.kind = "reference"
            except Exception as e:                      # fall back to the port, but say why
                print("reference package unusable (%s: %s); timing the oracle port" % (type(e).__name__, e),
                      file=sys.stderr)
                self._stack.close()
                self._stack = contextlib.ExitStack()

    def close(self):
        self._stack.close()

    def get_text(self, image):
        """pipeline.getText (ocr/pipeline.py:65-87) on a decoded image; returns the number of results."""
        import cv2
        import torch
        if self.kind == "reference":
            res = {}
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                for img in self.detector.process(image):
                    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
                    try:
                        _, res = self.recognizer.process(res, gray)
                    except IndexError:          # reference quirk (Attention): [s] at position 0
                        pass
            return len(res)
        from oracle import ocr_ref
        return len(ocr_ref.get_text(self.craft_sd, self.crnn_sd, image, self.head))

    def detect(self, image):
        import torch
        if self.kind == "reference":
            with torch.no_grad():
                return len(self.detector.process(image))
        from oracle import ocr_ref
        return len(ocr_ref.craft_process(self.craft_sd, image))

    def recognize(self, crops):
        import torch
        if self.kind == "reference":
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                for g in crops:
                    self.recognizer.getPreds(g)
            return len(crops)
        from oracle import ocr_ref
        for g in crops:
            ocr_ref.crnn_get_preds(self.crnn_sd, g, self.head)
        return len(crops)


def config_workload(cfg, head):
    from lightly_ocr_b200.synth import receipts
    import numpy as np
    if cfg == 1:
        crop = np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)
        return dict(metric="crnn_%s_single_crop_latency" % head.lower(), unit="ms/crop", higher=False, units=1,
                    workload="CRNN (%s) on a single 32x100 gray crop (BASELINE config 1)" % head, data=[crop])
    if cfg == 2:
        return dict(metric="craft_detect_latency_1280px", unit="ms/receipt", higher=False, units=1,
                    workload="CRAFT forward + getDetBoxes on one synthetic 1280x960 receipt (BASELINE config 2)",
                    data=[receipts.receipt(0)])
    if cfg == 3:
        return dict(metric="crnn_%s_crops_per_sec_batch512" % head.lower(), unit="crops/s", higher=True, units=512,
                    workload="CRNN (%s) on a batch of 512 ragged gray crops, TPS + BiLSTM + decode (BASELINE config 3)"
                             % head, data=receipts.crops(512, seed=3))
    raise ValueError(cfg)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation on the box's host cores, all threads; rank 0 only."""
    if rank != 0:
        return
    import torch
    from lightly_ocr_b200.synth import receipts
    cores = os.cpu_count() or 1
    ref = CpuReference(args.head, cores)
    if args.config in (4, 5):
        pool = [receipts.receipt(i) for i in range(4)]
        ref.get_text(pool[0][:640, :480].copy())                     # quarter receipt: first warm-up
        for w in range(max(args.warmup - 1, 0)):
            ref.get_text(pool[w % len(pool)])
        t0 = time.perf_counter()
        crops = 0
        for k in range(args.steps):
            crops += ref.get_text(pool[k % len(pool)])
        dt = time.perf_counter() - t0
        value, unit, metric, higher = args.steps / dt, UNIT, metric_name(args.head), True
        workload = "end-to-end CRAFT+CRNN(%s) over synthetic 1280x960 receipts (BASELINE config %d)" % (
            args.head, 4 if args.head == "CTC" else 5)
        sample = "one whole 1280x960 receipt (~%d crops) per step, %d steps" % (crops // max(args.steps, 1), args.steps)
        extra = {"crops_per_sec": crops / dt}
    else:
        wl = config_workload(args.config, args.head)
        data = wl["data"]
        if args.config == 3:
            data = data[:32]                                          # bounded sample of the 512-crop batch
        fn = {1: lambda: ref.recognize(data), 2: lambda: ref.detect(data[0]), 3: lambda: ref.recognize(data)}[args.config]
        for _ in range(max(args.warmup, 1)):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = time.perf_counter() - t0
        per_call = dt / args.steps
        value = len(data) / per_call if wl["higher"] else 1e3 * per_call
        unit, metric, higher, workload = wl["unit"], wl["metric"], wl["higher"], wl["workload"]
        sample = ("%d of the 512 crops per step" % len(data)) if args.config == 3 else "the whole unit per step"
        extra = {}
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": higher, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": unit, "cores": torch.get_num_threads(), "kind": ref.kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    ref.close()
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ the drop-in leg
def dropin_leg(head, device_index, images, via_reference=False):
    """receipts/s through the reference-facing classes exactly as ocr/pipeline.py:65-87 drives them: one image at a
    time, `net.CRAFT.process(image)` then one `net.CRNN.process(result, gray)` per crop with cv2.cvtColor on the host
    (decoded images in, result dicts out).  Nothing under oracle/ is touched.  via_reference (--dropin-via-reference,
    needs baseline/_ref/ocr): additionally the unmodified reference's own `pipeline.getText` on PNG files (cv2.imread
    included) with the drop-in classes behind it, staged by the same helper the reference arm uses."""
    import cv2
    import torch
    import yaml
    from lightly_ocr_b200.synth import weights
    tmp = tempfile.mkdtemp(prefix="locr_dropin_")
    craft_sd, crnn_sd = weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, head)
    out = {}
    ref_env = None
    if via_reference:
        from oracle import ref_env as _ref_env
        if _ref_env.source() is not None:
            ref_env = _ref_env
    if ref_env is not None:
        dst = ref_env.stage(head, craft_sd, crnn_sd, tmp)
    else:
        dst = os.path.join(tmp, "ocr_" + head)
        os.makedirs(os.path.join(dst, "save_models"))
        with open(os.path.join(ROOT, "lightly_ocr_b200", "config.yml")) as f:
            cfg = yaml.safe_load(f)
        cfg["prediction"], cfg["num_classes"] = head, (37 if head == "CTC" else 38)
        with open(os.path.join(dst, "config.yml"), "w") as f:
            yaml.safe_dump(cfg, f)
        torch.save(craft_sd, os.path.join(dst, "save_models", "CRAFT.pth"))
        torch.save(crnn_sd, os.path.join(dst, "save_models", "CRNN.pth"))
    os.environ["LOCR_OCR_DIR"] = dst
    import importlib
    import lightly_ocr_b200.net as net
    net = importlib.reload(net)
    dev = torch.device("cuda", device_index)
    detector, recognizer = net.CRAFT(device=dev), net.CRNN(device=dev)

    def get_text(image):
        res = {}
        for img in detector.process(image):
            gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
            try:
                _, res = recognizer.process(res, gray)
            except IndexError:
                pass
        return res

    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        for im in images[:2]:
            get_text(im)
        torch.cuda.synchronize()
        lat = []
        t0 = time.perf_counter()
        crops = 0
        for im in images:
            t1 = time.perf_counter()
            crops += len(get_text(im))
            lat.append(time.perf_counter() - t1)
        dt = time.perf_counter() - t0
    out = {"value": len(images) / dt, "unit": UNIT, "ms_per_receipt": 1e3 * dt / len(images),
           "ms_per_receipt_median": 1e3 * statistics.median(lat), "crops_per_sec": crops / dt, "receipts": len(images),
           "path": "net.CRAFT.process(image) + net.CRNN.process(result, gray) per crop, one image at a time "
                   "(ocr/pipeline.py:70-79); decoded BGR arrays in, result dict out"}
    if ref_env is not None:
        paths = []
        for i, im in enumerate(images):
            p = os.path.join(tmp, "r%d.png" % i)
            cv2.imwrite(p, im)
            paths.append(p)
        for e in net._ENGINES.values():
            e.close()
        with ref_env.imported(dst, dropin=True):
            import pipeline
            with contextlib.redirect_stdout(sink):
                d2, r2 = pipeline.prepModel(pipeline.CONFIG)
                pipeline.getText(paths[0], d2, r2, write=False)
                t0 = time.perf_counter()
                for p in paths:
                    try:
                        pipeline.getText(p, d2, r2, write=False)
                    except IndexError:
                        pass
                dt2 = time.perf_counter() - t0
            import lightly_ocr_b200.net as net2
            for e in net2._ENGINES.values():
                e.close()
        out["via_reference_getText_png"] = {"value": len(paths) / dt2, "unit": UNIT,
                                            "ms_per_receipt": 1e3 * dt2 / len(paths),
                                            "path": "the unmodified reference pipeline.getText(path) over the drop-in "
                                                    "(cv2.imread of a PNG file included)"}
    else:
        for e in net._ENGINES.values():
            e.close()
    return out


# ------------------------------------------------------------------------------------------------ configs 1-3
def run_small_config(args, rank, local_rank):
    """BASELINE configs 1-3 on one lane of one GPU: `value` from CUDA events on the handle's stream around the calls
    (host->device copies of the inputs are on that stream), `e2e` from the wall clock around the same calls."""
    if rank != 0:
        return
    import torch
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import weights
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    wl = config_workload(args.config, args.head)
    r = bridge.OcrRunner(device_id=local_rank, act_dtype=bridge.ACT_F16, head=args.head)
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, args.head))
    data = wl["data"]
    if args.config == 2:
        h2d = int(data[0].nbytes)
        fn = lambda: r.detect(data)
        res_fn = lambda: r.detect_resident(1)
    else:
        h2d = int(sum(c.nbytes for c in data))
        fn = lambda: r.recognize(data, want_logits=False)
        res_fn = None
    reps = {1: 64, 2: 64, 3: 8}[args.config]            # calls per step: a 20-step leg lasts ~2 s
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        fn()
    l0 = r.launch_count()
    r.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps * reps):
        fn()
    dev_ms = r.timer_stop()
    wall = time.perf_counter() - t0
    launches = r.launch_count() - l0
    calls = args.steps * reps
    clk = clocks.window(t0, t0 + wall)
    res_ms = None
    if res_fn is not None:
        fn()
        for _ in range(3):
            res_fn()
        r.timer_start()
        for _ in range(calls):
            res_fn()
        res_ms = r.timer_stop()
    r.profile(True)
    r.profile_read()
    r.timer_start()
    for _ in range(args.steps):
        fn()
    prof_ms = r.timer_stop()
    conv_ms, conv_flops, conv_launches = r.profile_read()
    r.profile(False)
    clocks.stop()
    peak_tf, _, which = measured_peaks()

    def to_value(ms_per_call):
        return wl["units"] * 1e3 / ms_per_call if wl["higher"] else ms_per_call / wl["units"]

    dev_call = (res_ms if res_ms is not None else dev_ms) / calls
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    line = {"metric": wl["metric"], "value": to_value(dev_call), "unit": wl["unit"], "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_call * reps, "higher_is_better": wl["higher"],
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": wl["workload"], "calls_per_step": reps,
                       "l2": "back-to-back calls on one stream; activations of configs 2 / 3 exceed the 126 MB L2, "
                             "config 1 is latency-bound (weights stay in L2 as they would in a serving process)",
                       "value_is": ("device time with the receipt resident in HBM (locr_detect_resident)"
                                    if res_ms is not None else
                                    "device time (CUDA events on the handle's stream) around the host-buffer calls")},
            "e2e": {"value": to_value(1e3 * wall / calls), "unit": wl["unit"], "h2d_bytes_per_step": h2d * reps,
                    "d2h_bytes_per_step": None, "device_ms_per_call": dev_ms / calls,
                    "note": "wall clock around the same host-buffer calls"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all layers)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": None,
                         "peak_source": "bf16_tflops_sustained, %s" % which, "launches": conv_launches,
                         "kernel_ms": conv_ms, "pass_ms": prof_ms,
                         "kernel_share_of_step": conv_ms / prof_ms if prof_ms > 0 else None,
                         "algorithmic_flops": conv_flops},
            "clocks": clk}
    r.close()
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ref = CpuReference(args.head, cores)
        sample = data[:16] if args.config == 3 else data
        fnr = (lambda: ref.detect(sample[0])) if args.config == 2 else (lambda: ref.recognize(sample))
        fnr()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 5.0:
            fnr()
            n += 1
        per = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": len(sample) / per if wl["higher"] else 1e3 * per, "unit": wl["unit"],
                                "cores": cores, "kind": ref.kind,
                                "sample": "%d unit(s) per call, %d calls in ~5 s" % (len(sample), n)}
        ref.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--head", default=None, choices=["CTC", "Attention"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin leg (one image at a time through net.py)")
    ap.add_argument("--dropin-via-reference", action="store_true",
                    help="e2e_dropin additionally times the unmodified reference's pipeline.getText(path) over the drop-in "
                         "classes (needs baseline/_ref/ocr)")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"],
                    help="arithmetic of the recogniser for the main legs (include/locr.h LOCR_PREC_*); the other mode is "
                         "timed as an extra e2e leg (`other_precision`) unless --no-other-precision")
    ap.add_argument("--no-other-precision", action="store_true")
    ap.add_argument("--no-other-head", action="store_true",
                    help="skip the `other_head` leg (the e2e leg once more with the other prediction head: the attention "
                         "decoder of BASELINE config 5 when the main legs run config 4)")
    ap.add_argument("--jpeg", action="store_true",
                    help="extra leg: the same receipts handed over as JPEG files (q90, 4:2:0) through locr_detect_encoded; "
                         "adds an `e2e_jpeg` object to the JSON line (BASELINE's metric itself excludes the image decode)")
    args = ap.parse_args()
    if args.head is None:
        args.head = "Attention" if args.config == 5 else "CTC"
    if args.config == 5 and args.head != "Attention":
        raise SystemExit("--config 5 is the attention decoder")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config in (1, 2, 3):
        run_small_config(args, rank, local_rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from lightly_ocr_b200 import bridge, shard
    from lightly_ocr_b200.synth import weights
    # LOCR_BENCH_BACKEND=gloo exists for tests/test_bench_cpu.py only: the bench's multi-rank bookkeeping (barriers, per-rank
    # times, reductions, the one JSON line) run on CPU tensors against a stand-in engine.  The product path stays CUDA-only.
    backend = os.environ.get("LOCR_BENCH_BACKEND", "nccl")
    if backend == "nccl":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        if world > 1:
            dist.init_process_group(backend)
    head = args.head

    from concurrent.futures import ThreadPoolExecutor
    craft_sd, crnn_sd = weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, head)
    def make_runners(precision, head_=None):
        sd = crnn_sd if head_ in (None, head) else weights.crnn_calibrated(1, head_)
        out = []
        for _ in range(LANES):
            r = bridge.OcrRunner(device_id=local_rank, act_dtype=bridge.ACT_F16, head=head_ or head,
                                 precision=bridge.PREC_EXACT if precision == "exact" else bridge.PREC_FAST)
            r.load_state_dict(bridge.MODEL_CRAFT, craft_sd)
            r.load_state_dict(bridge.MODEL_CRNN, sd)
            out.append(r)
        return out

    runners = make_runners(args.precision)
    pool = make_receipts(rank, POOL)
    batches = [pool[i:i + PER_PASS] for i in range(0, POOL, PER_PASS)]
    ex = ThreadPoolExecutor(max_workers=LANES)

    def lanes(fn):
        """Run fn(lane index, runner) on every lane concurrently (ctypes releases the GIL inside liblocr)."""
        return [f.result() for f in [ex.submit(fn, i, r) for i, r in enumerate(runners)]]

    def barrier():
        if world > 1:
            dist.barrier()
        if backend == "nccl":
            torch.cuda.synchronize()

    def launches():
        return sum(r.launch_count() for r in runners)

    def per_rank(dev_s, wall_s):
        """elapsed per rank (device events / wall clock), gathered so that a slow rank is visible in the JSON line"""
        if world == 1:
            return [[dev_s, wall_s]]
        t = torch.tensor([dev_s, wall_s], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [[float(x[0]), float(x[1])] for x in out]

    def rank_stats(rows):
        e = [max(a, b) for a, b in rows]
        return {"elapsed_s": {"min": min(e), "median": statistics.median(e), "max": max(e)},
                "per_rank_dev_s": [round(a, 5) for a, _ in rows], "per_rank_wall_s": [round(b, 5) for _, b in rows]}

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                                   # before the warm-up: nvidia-smi takes a while to start

    # ---------------- e2e: host buffers in, host results out, every pass
    def e2e_pass(k):
        batch = batches[k % len(batches)]
        outs = lanes(lambda i, r: r.ocr(batch[i * PER_LANE:(i + 1) * PER_LANE]))
        crops_ = sum(len(o[1]["text"]) for o in outs)
        d2h_ = sum(sum(len(x) for x in o[0]) * 16 + len(o[1]["text"]) * (26 * 4 + bridge.TEXT_STRIDE + 8) for o in outs)
        return crops_, d2h_

    for w in range(max(args.warmup, 3)):
        e2e_pass(w)
    barrier()
    launches0 = launches()
    for r in runners:
        r.timer_start()
    t0 = time.perf_counter()
    crops_e2e = 0
    d2h = 0
    for k in range(args.steps * PASSES):
        c_, d_ = e2e_pass(k)
        crops_e2e += c_
        d2h += d_
    e2e_ms = max(r.timer_stop() for r in runners)
    e2e_wall = time.perf_counter() - t0
    e2e_window = (t0, t0 + e2e_wall)
    barrier()
    e2e_rows = per_rank(e2e_ms / 1e3, e2e_wall)
    e2e_s = max(max(a, b) for a, b in e2e_rows)
    launches_e2e = launches() - launches0

    # ---------------- value: the same path with the pass's receipts already resident in HBM
    e2e_pass(0)                                          # leaves batch 0 resident (PER_LANE receipts on each lane)
    for _ in range(max(args.warmup, 3)):
        lanes(lambda i, r: r.ocr_resident(PER_LANE))
    barrier()
    launches0 = launches()
    for r in runners:
        r.timer_start()
    t0 = time.perf_counter()
    crops = 0
    for k in range(args.steps * PASSES):
        outs = lanes(lambda i, r: r.ocr_resident(PER_LANE))
        crops += sum(len(o[1]["text"]) for o in outs)
    dev_ms = max(r.timer_stop() for r in runners)
    wall = time.perf_counter() - t0
    barrier()
    res_rows = per_rank(dev_ms / 1e3, wall)
    elapsed = max(max(a, b) for a, b in res_rows)
    clk = clocks.window(t0, t0 + wall) if rank == 0 else None
    clk_e2e = clocks.window(*e2e_window) if rank == 0 else None
    n_launches = launches() - launches0
    # ---------------- roofline pass: the same work on ONE lane with per-launch CUDA events around every conv_tc_kernel
    # (with two lanes a launch can queue behind the other lane's kernel and its event pair would over-count)
    r0 = runners[0]
    r0.profile(True)
    r0.profile_read()
    r0.profile_layers()        # clears the per-kernel totals
    r0.timer_start()
    for k in range(args.steps):
        r0.ocr_resident(PER_LANE)
    prof_ms = r0.timer_stop()
    conv_ms, conv_flops, conv_launches = r0.profile_read()
    all_kernels_ms = sum(row[1] for row in r0.profile_layers())      # every launch of the pass, conv or not
    r0.profile(False)
    barrier()
    if rank == 0:
        clocks.stop()
    # ---------------- optional, LAST so that it cannot disturb the legs above: encoded files in (Huffman decoding on host
    # threads, the rest of the JPEG reader on the GPU)
    e2e_jpeg = None
    if args.jpeg:
        import cv2
        blobs = [[cv2.imencode(".jpg", np.asarray(im), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in b] for b in batches]

        def jpeg_pass(k):
            bl = blobs[k % len(blobs)]
            return sum(len(o[1]["text"]) for o in lanes(lambda i, r: r.ocr_encoded(bl[i * PER_LANE:(i + 1) * PER_LANE])))

        for w in range(max(args.warmup, 3)):
            jpeg_pass(w)
        barrier()
        t0 = time.perf_counter()
        crops_jpeg = sum(jpeg_pass(k) for k in range(args.steps * PASSES))
        barrier()
        jpeg_s = shard.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_jpeg = {"value": world * RECEIPTS_PER_STEP * args.steps / jpeg_s, "unit": UNIT,
                    "crops_per_sec": world * crops_jpeg / jpeg_s,
                    "file_bytes_per_step": int(PASSES * sum(len(x) for b in blobs for x in b) / len(blobs)),
                    "h2d_bytes_per_step": int(PASSES * sum((im.shape[0] + 15) // 16 * ((im.shape[1] + 15) // 16) * 6 * 128
                                                           for im in batches[0])),   # quantised coefficients, 4:2:0
                    "input": "JPEG q90 4:2:0 files of the same receipts; entropy decoding on host threads (one per "
                             "image), coefficients -> pixels on the GPU (lightly_ocr_b200/csrc/jpeg.cu); wall-clock"}

    total_crops = crops
    if world > 1:
        t = torch.tensor([crops, crops_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_crops, crops_e2e = int(t[0].item()), int(t[1].item())
    for r in runners:
        r.close()
    # ---------------- two more e2e legs (host buffers in, host results out) at every N, so that the driver's 1 -> 8 sweep
    # carries them too: the other arithmetic of the recogniser, and the other prediction head (BASELINE config 5 when
    # the main legs are config 4 and vice versa)
    def extra_leg(precision, head_):
        nonlocal runners
        runners = make_runners(precision, head_)
        n_pass = max(args.steps // 2, 2) * PASSES
        for w in range(3):
            e2e_pass(w)
        barrier()
        for r in runners:
            r.timer_start()
        t0 = time.perf_counter()
        oc = 0
        for k in range(n_pass):
            oc += e2e_pass(k)[0]
        o_ms = max(r.timer_stop() for r in runners)
        o_wall = time.perf_counter() - t0
        barrier()
        o_rows = per_rank(o_ms / 1e3, o_wall)
        o_s = max(max(a, b) for a, b in o_rows)
        if world > 1:
            t = torch.tensor([oc], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            oc = int(t[0].item())
        for r in runners:
            r.close()
        runners = []
        return {"precision": precision, "head": head_, "value": world * PER_PASS * n_pass / o_s, "unit": UNIT,
                "crops_per_sec": oc / o_s, "passes": n_pass, "n_gpus": world,
                "leg": "e2e (host buffers in, host results out), same receipts and lanes as `e2e`"}

    other = None
    if not args.no_other_precision:
        other = extra_leg("exact" if args.precision == "fast" else "fast", head)
    other_head = None
    if not args.no_other_head:
        head2 = "Attention" if head == "CTC" else "CTC"
        try:
            other_head = extra_leg(args.precision, head2)
            other_head["metric"] = metric_name(head2)
            other_head["workload"] = ("end-to-end CRAFT+CRNN(%s) over synthetic 1280x960 receipts (BASELINE config %d)"
                                      % (head2, 4 if head2 == "CTC" else 5))
        except Exception as e:        # an extra leg must not cost the run its headline line (same on every rank)
            other_head = {"head": head2, "error": "%s: %s" % (type(e).__name__, e)}
    runners = []

    if rank == 0:
        peak_tf, peak_hbm, which = measured_peaks()
        receipts_total = world * RECEIPTS_PER_STEP * args.steps
        value = receipts_total / elapsed
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        line = {
            "metric": metric_name(head), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "precision": args.precision,
            # `workload` is the same string in both arms (--impl reference); how each arm batches it is its own key
            "config": {"workload": "end-to-end CRAFT+CRNN(%s) over synthetic 1280x960 receipts (BASELINE config %d)"
                                   % (head, 4 if head == "CTC" else 5),
                       "batching": "%d receipts per step per GPU in %d passes of %d; ~%d crops per receipt"
                                   % (RECEIPTS_PER_STEP, PASSES, PER_PASS, total_crops // max(receipts_total, 1)),
                       "l2": "inputs + activations per pass (~0.6 GB per receipt) far exceed the 126 MB L2",
                       "weights": "synthetic checkpoints (lightly_ocr_b200/synth: seed-generated, read-out trained on synthetic receipts), fp16 storage, fp32 accumulate",
                       "parallelism": "replicas x%d, receipts sharded, no collective; %d host lanes (handles/streams) per GPU" % (world, LANES)},
            "crops_per_sec": total_crops / elapsed,
            "ranks": rank_stats(res_rows),
            "e2e": {"value": world * RECEIPTS_PER_STEP * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(PASSES * sum(im.nbytes for im in batches[0])),
                    "d2h_bytes_per_step": int(d2h / max(args.steps, 1)), "crops_per_sec": crops_e2e / e2e_s,
                    "gpu_launches": launches_e2e, "ranks": rank_stats(e2e_rows), "clocks": clk_e2e},
            "gpu_launches": n_launches,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all layers)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": conv_traffic()[0],
                         "traffic_unit": "DRAM bytes per launch (mean over the conv launches of one 8-receipt pass)",
                         "traffic_source": conv_traffic()[1],
                         "peak_source": "bf16_tflops_sustained, %s (fp16 and bf16 share the tensor rate)" % which,
                         "launches": conv_launches,
                         "timed_region": "separate single-lane pass of %d x %d receipts right after the timed "
                                         "steps (per-launch CUDA events on the launching stream)" % (args.steps, PER_LANE),
                         "kernel_ms": conv_ms, "pass_ms": prof_ms,
                         "kernel_share_of_step": conv_ms / prof_ms if prof_ms > 0 else None,
                         # share of the GPU's busy time (what an ncu launch list measures: no host gaps in it)
                         "all_kernels_ms": all_kernels_ms,
                         "kernel_share_of_gpu_time": conv_ms / all_kernels_ms if all_kernels_ms > 0 else None,
                         "algorithmic_flops": conv_flops,
                         "whole_step_tensor_frac": ((CRAFT_FLOPS * receipts_total + CRNN_FLOPS_PER_CROP * total_crops)
                                                    / elapsed / 1e12 / world / peak_tf) if peak_tf else None},
            "clocks": clk,
        }
        if other is not None:
            line["other_precision"] = other
        if other_head is not None:
            line["other_head"] = other_head
        if e2e_jpeg is not None:
            line["e2e_jpeg"] = e2e_jpeg
        if world == 1 and not args.no_dropin:
            try:
                line["e2e_dropin"] = dropin_leg(head, local_rank, pool[:8], args.dropin_via_reference)
            except Exception as e:
                line["e2e_dropin"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_cpu_baseline and world == 1:     # rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            ref = CpuReference(head, cores)
            from lightly_ocr_b200.synth import receipts
            whole = receipts.receipt(0)
            ref.get_text(np.ascontiguousarray(whole[:640, :480]))
            t0 = time.perf_counter()
            ncrops = ref.get_text(whole)
            secs = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": ref.kind,
                                    "sample": "one whole 1280x960 receipt (%d crops) through %s (torch fp32 with %d "
                                              "threads + cv2 + PIL), timed once after a warm-up pass on a 640x480 window"
                                              % (ncrops, "the unmodified reference's CRAFT.process + CRNN.process "
                                                 "(baseline/_ref/ocr)" if ref.kind == "reference" else
                                                 "oracle/ocr_ref.get_text", cores)}
            ref.close()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
