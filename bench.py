#!/usr/bin/env python
"""Throughput of the detect-then-recognize path (BASELINE.json metric: end-to-end receipts/sec at 1280 px,
CRAFT + CRNN, 1/2/4/8 B200; crops/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one pass of the whole path (CRAFT forward, thresholds + labelling + boxes, host reading-order sort, GPU
crops + BICUBIC, CRNN forward, CTC decode) over a batch of RECEIPTS_PER_STEP distinct synthetic 1280x960 receipts per
GPU (weak scaling: per-GPU work is fixed).  `value` times the path with the receipts already resident in HBM;
`e2e` times the same path through the C ABI with host buffers (host->device copy of the images and device->host copy
of rects, strings and confidences inside the timed region).  Receipts are independent units: no collective on the
data path, torch.distributed only provides the barrier and the max-over-ranks reduction of the elapsed time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LANES = int(os.environ.get("LOCR_BENCH_LANES", "2"))   # host threads per GPU, each with its own liblocr handle/stream:
                               # while one lane sorts rects / copies results on the host, the other lanes' kernels
                               # keep the GPU busy
PER_LANE = int(os.environ.get("LOCR_BENCH_PER_LANE", "8"))
RECEIPTS_PER_STEP = PER_LANE * LANES  # receipts per lane and step (one CRAFT launch sequence over that many canvases)
POOL = 2 * RECEIPTS_PER_STEP   # distinct receipts cycled through (3.7 MB of pixels each)
METRIC = "receipts_per_sec_1280px_craft_crnn_ctc"
UNIT = "receipts/s"
CRAFT_FLOPS = 874.217e9        # per 1280x960 canvas
CRNN_FLOPS_PER_CROP = 10.593e9


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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  The sampler runs
    from before the warm-up (nvidia-smi needs a few hundred ms to produce its first line); stop(t0, t1) keeps the
    samples whose arrival time falls inside the timed window [t0, t1]."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03]
        window = "timed region"
        if not rows:
            rows = [r for (_, r) in self.rows]
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
    """Decoded BGR receipts in PINNED host memory (the e2e leg copies them to the device every step)."""
    import torch
    from lightly_ocr_b200.synth import receipts
    out = []
    for i in range(count):
        t = torch.from_numpy(receipts.receipt(1000 * rank + i)).pin_memory()
        out.append(t.numpy())
    make_receipts.keep = getattr(make_receipts, "keep", []) + out     # the arrays view the pinned tensors' storage
    return out


def cpu_oracle_sample(n_threads, full=False):
    """The oracle (CPU port of the reference path, torch fp32 + cv2 + PIL) on a bounded sample: one quarter receipt
    (640x480 window of receipt 0, ~24 words) as warm-up / reference-arm unit and, with full=True, one whole 1280x960
    receipt timed after it.  Returns (seconds, crops, state)."""
    import numpy as np
    import torch
    from lightly_ocr_b200.synth import receipts
    from oracle import ocr_ref, weights
    torch.set_num_threads(n_threads)
    craft_sd, crnn_sd = weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, "CTC")
    img = np.ascontiguousarray(receipts.receipt(0)[:640, :480])
    t0 = time.perf_counter()
    res = ocr_ref.get_text(craft_sd, crnn_sd, img, "CTC")
    dt = time.perf_counter() - t0
    if full:
        whole = receipts.receipt(0)
        t0 = time.perf_counter()
        res = ocr_ref.get_text(craft_sd, crnn_sd, whole, "CTC")
        dt = time.perf_counter() - t0
    return dt, len(res), (craft_sd, crnn_sd, img)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; /root/reference cannot travel to
    the GPU box) on the box's host cores, all threads, one whole 1280x960 receipt per step."""
    if rank != 0:
        return
    import torch
    from lightly_ocr_b200.synth import receipts
    from oracle import ocr_ref
    cores = os.cpu_count() or 1
    _, _, (craft_sd, crnn_sd, img) = cpu_oracle_sample(cores)          # quarter receipt: first warm-up
    pool = [receipts.receipt(i) for i in range(4)]
    for w in range(max(args.warmup - 1, 0)):
        ocr_ref.get_text(craft_sd, crnn_sd, pool[w % len(pool)], "CTC")
    t0 = time.perf_counter()
    crops = 0
    for k in range(args.steps):
        crops += len(ocr_ref.get_text(craft_sd, crnn_sd, pool[k % len(pool)], "CTC"))
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = "one whole 1280x960 receipt (~%d crops) per step, %d steps" % (crops // max(args.steps, 1), args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "end-to-end CRAFT+CRNN(CTC) over synthetic 1280x960 receipts (BASELINE config 4)",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "crops_per_sec": crops / dt}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--jpeg", action="store_true",
                    help="extra leg: the same receipts handed over as JPEG files (q90, 4:2:0) through locr_detect_encoded; "
                         "adds an `e2e_jpeg` object to the JSON line (BASELINE's metric itself excludes the image decode)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from lightly_ocr_b200 import bridge, shard
    from lightly_ocr_b200.synth import weights
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    from concurrent.futures import ThreadPoolExecutor
    craft_sd, crnn_sd = weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, "CTC")
    runners = []
    for _ in range(LANES):
        r = bridge.OcrRunner(device_id=local_rank, act_dtype=bridge.ACT_F16, head="CTC")
        r.load_state_dict(bridge.MODEL_CRAFT, craft_sd)
        r.load_state_dict(bridge.MODEL_CRNN, crnn_sd)
        runners.append(r)
    pool = make_receipts(rank, POOL)
    per_lane = RECEIPTS_PER_STEP // LANES
    batches = [pool[i:i + RECEIPTS_PER_STEP] for i in range(0, POOL, RECEIPTS_PER_STEP)]
    ex = ThreadPoolExecutor(max_workers=LANES)

    def lanes(fn):
        """Run fn(lane index, runner) on every lane concurrently (ctypes releases the GIL inside liblocr)."""
        return [f.result() for f in [ex.submit(fn, i, r) for i, r in enumerate(runners)]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        return sum(r.launch_count() for r in runners)

    # ---------------- e2e: host buffers in, host results out, every step
    def e2e_step(k):
        batch = batches[k % len(batches)]
        outs = lanes(lambda i, r: r.ocr(batch[i * per_lane:(i + 1) * per_lane]))
        crops_ = sum(len(o[1]["text"]) for o in outs)
        d2h_ = sum(sum(len(x) for x in o[0]) * 16 + len(o[1]["text"]) * (26 * 4 + bridge.TEXT_STRIDE + 8) for o in outs)
        return crops_, d2h_

    for w in range(max(args.warmup, 3)):
        e2e_step(w)
    barrier()
    launches0 = launches()
    for r in runners:
        r.timer_start()
    t0 = time.perf_counter()
    crops_e2e = 0
    d2h = 0
    for k in range(args.steps):
        c_, d_ = e2e_step(k)
        crops_e2e += c_
        d2h += d_
    e2e_ms = max(r.timer_stop() for r in runners)
    e2e_wall = time.perf_counter() - t0
    barrier()
    e2e_s = shard.max_over_ranks(max(e2e_ms / 1e3, e2e_wall), dev)
    launches_e2e = launches() - launches0

    # ---------------- value: the same path with the step's receipts already resident in HBM
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                                   # before the warm-up: nvidia-smi takes a while to start
    e2e_step(0)                                          # leaves batch 0 resident (per_lane receipts on each lane)
    for _ in range(max(args.warmup, 3)):
        lanes(lambda i, r: r.ocr_resident(per_lane))
    barrier()
    launches0 = launches()
    for r in runners:
        r.timer_start()
    t0 = time.perf_counter()
    crops = 0
    for k in range(args.steps):
        outs = lanes(lambda i, r: r.ocr_resident(per_lane))
        crops += sum(len(o[1]["text"]) for o in outs)
    dev_ms = max(r.timer_stop() for r in runners)
    wall = time.perf_counter() - t0
    barrier()
    clk = clocks.stop(t0, t0 + wall) if rank == 0 else None
    elapsed = shard.max_over_ranks(max(dev_ms / 1e3, wall), dev)
    n_launches = launches() - launches0
    # ---------------- roofline pass: the same steps on ONE lane with per-launch CUDA events around every conv_tc_kernel
    # (with two lanes a launch can queue behind the other lane's kernel and its event pair would over-count)
    r0 = runners[0]
    r0.profile(True)
    r0.profile_read()
    r0.timer_start()
    for k in range(args.steps):
        r0.ocr_resident(per_lane)
    prof_ms = r0.timer_stop()
    conv_ms, conv_flops, conv_launches = r0.profile_read()
    r0.profile(False)
    barrier()
    # ---------------- optional, LAST so that it cannot disturb the legs above: encoded files in (Huffman decoding on host
    # threads, the rest of the JPEG reader on the GPU)
    e2e_jpeg = None
    if args.jpeg:
        import cv2
        blobs = [[cv2.imencode(".jpg", np.asarray(im), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in b] for b in batches]

        def jpeg_step(k):
            bl = blobs[k % len(blobs)]
            return sum(len(o[1]["text"]) for o in lanes(lambda i, r: r.ocr_encoded(bl[i * per_lane:(i + 1) * per_lane])))

        for w in range(max(args.warmup, 3)):
            jpeg_step(w)
        barrier()
        t0 = time.perf_counter()
        crops_jpeg = sum(jpeg_step(k) for k in range(args.steps))
        barrier()
        jpeg_s = shard.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_jpeg = {"value": world * RECEIPTS_PER_STEP * args.steps / jpeg_s, "unit": UNIT,
                    "crops_per_sec": world * crops_jpeg / jpeg_s,
                    "file_bytes_per_step": int(sum(len(x) for b in blobs for x in b) / len(blobs)),
                    "h2d_bytes_per_step": int(sum((im.shape[0] + 15) // 16 * ((im.shape[1] + 15) // 16) * 6 * 128
                                                  for im in batches[0])),   # quantised coefficients, 4:2:0

                    "input": "JPEG q90 4:2:0 files of the same receipts; entropy decoding on host threads (one per "
                             "image), coefficients -> pixels on the GPU (lightly_ocr_b200/csrc/jpeg.cu); wall-clock"}

    total_crops = crops
    if world > 1:
        t = torch.tensor([crops, crops_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_crops, crops_e2e = int(t[0].item()), int(t[1].item())

    if rank == 0:
        peak_tf, peak_hbm, which = measured_peaks()
        receipts_total = world * RECEIPTS_PER_STEP * args.steps
        value = receipts_total / elapsed
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "end-to-end CRAFT+CRNN(CTC) over synthetic 1280x960 receipts, %d receipts per "
                                   "step per GPU (BASELINE config 4; ~%d crops per receipt)"
                                   % (RECEIPTS_PER_STEP, total_crops // max(receipts_total, 1)),
                       "l2": "inputs + activations per step (~0.6 GB per receipt) far exceed the 126 MB L2",
                       "weights": "synthetic checkpoints (lightly_ocr_b200/synth: seed-generated, CTC read-out trained on synthetic receipts), fp16 storage, fp32 accumulate",
                       "parallelism": "replicas x%d, receipts sharded, no collective; %d host lanes (handles/streams) per GPU" % (world, LANES)},
            "crops_per_sec": total_crops / elapsed,
            "e2e": {"value": world * RECEIPTS_PER_STEP * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(im.nbytes for im in batches[0])),
                    "d2h_bytes_per_step": int(d2h / max(args.steps, 1)), "crops_per_sec": crops_e2e / e2e_s,
                    "gpu_launches": launches_e2e},
            "gpu_launches": n_launches,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all layers)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": conv_traffic()[0],
                         "traffic_unit": "DRAM bytes per launch (mean over the conv launches of one 8-receipt pass)",
                         "traffic_source": conv_traffic()[1],
                         "peak_source": "bf16_tflops_sustained, %s (fp16 and bf16 share the tensor rate)" % which,
                         "launches": conv_launches,
                         "timed_region": "separate single-lane pass of %d steps x %d receipts right after the timed "
                                         "steps (per-launch CUDA events on the launching stream)" % (args.steps, per_lane),
                         "kernel_ms": conv_ms, "pass_ms": prof_ms,
                         "kernel_share_of_step": conv_ms / prof_ms if prof_ms > 0 else None,
                         "algorithmic_flops": conv_flops},
            "clocks": clk,
        }
        if e2e_jpeg is not None:
            line["e2e_jpeg"] = e2e_jpeg
        if not args.no_cpu_baseline and world == 1:     # rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            secs, ncrops, _ = cpu_oracle_sample(cores, full=True)
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "one whole 1280x960 receipt (%d crops) through oracle/ocr_ref.get_text "
                                              "(torch fp32 with %d threads + cv2 + PIL), timed once after a warm-up "
                                              "pass on a 640x480 window" % (ncrops, cores)}
        print(json.dumps(line))
    for r in runners:
        r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
