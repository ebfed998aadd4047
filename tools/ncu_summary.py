"""Text summary of the two `ncu --set full` captures of tools/prof_round.sh (read on the CPU box with `ncu -i`).
Usage: python tools/ncu_summary.py r01d [dir] > profiles/r01d_ncu_full_summary.txt
`dir` (default gpurun_out) holds prof_{tc,mem}_<tag>.ncu-rep; a capture too large to bring back from the GPU box can be
replaced by its raw-page csv made there (prof_{tc,mem}_<tag>_raw.csv: `ncu -i rep --page raw --csv --metrics ...`)."""
import csv
import os
import subprocess
import sys

tag = sys.argv[1]
root = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
M = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,"
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,"
     "l1tex__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,"
     "launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active")
print("ncu --set full --clock-control none (tools/prof_round.sh, capture %s), B200, fp16 storage" % tag)
print("targets: tools/prof_kernels.py tc (one launch per layer, LOCR_BENCH_WARMUP=0) and tools/prof_kernels.py mem "
      "(8 receipts, 635 crops)")
print("sources: gpurun_out/prof_tc_%s.ncu-rep, gpurun_out/prof_mem_%s.ncu-rep (scratch); cold-cache replays: compare "
      "shares, not absolutes\n" % (tag, tag))
labels_tc = ["slice1.0 16->64 @1280x960 x8, plain 9-tap form (first launch of the process: cold)",
             "slice1.3 64->64 @1280x960 x8, haloed-patch path, only the 2x2 max-pooled tensor written (as in the pipeline)",
             "slice1.10 128->128 @640x480 x8 (CTA pairs, N = 128)", "slice3.27 512->512 @160x120 x8 (CTA pairs, N = 256)",
             "conv_cls.0 32->32 @640x480 x8, haloed-patch form on 64-byte pixels",
             "upconv4.3 64->32 @640x480 x8, haloed-patch form",
             "slice2.17 256->256 @320x240 x8 (CTA pairs, N = 256)",
             "CRNN 512->512 @4x26 x640 crops (CTA pairs, N = 256)",
             "BiLSTM recurrence, 650 crops x 26 steps x 2 directions"]
if tag < "r02b":      # captures before the round-2 additions to tools/prof_kernels.py
    labels_tc = labels_tc[:5] + labels_tc[7:]
for rep, labels in (("%s/prof_tc_%s.ncu-rep" % (root, tag), labels_tc), ("%s/prof_mem_%s.ncu-rep" % (root, tag), None)):
    raw = rep.replace(".ncu-rep", "_raw.csv")
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", M], capture_output=True, text=True).stdout
    elif os.path.exists(raw):
        out = open(raw).read()
    else:
        print("(no capture %s)" % rep)
        continue
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    for i, r in enumerate(rows[2:]):
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("locr::<unnamed>::", "").replace("unnamed>::", "")

        def val(k):
            v = float(d[k].replace(",", ""))
            return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u[k], 1)

        t, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        print("=== %s%s" % (name, ("  [%s]" % labels[i]) if labels and i < len(labels) else ""))
        print("  gpu__time_duration.sum %.1f us | dram read %.2f MB write %.2f MB -> %.0f GB/s (%.2f of the measured "
              "6542 GB/s copy rate)" % (t, rd, wr, (rd + wr) / t * 1e3, (rd + wr) / t * 1e3 / 6542))
        print("  tensor pipe active %.1f %% | sm throughput %.1f %% | dram %.1f %% | lts %.1f %% | l1tex %.1f %% | warps "
              "active %.1f %%" % tuple(float(d[k]) for k in (
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                  "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__warps_active.avg.pct_of_peak_sustained_active")))
        print("  grid %s block %s regs/thread %s" % (d["launch__grid_size"], d["launch__block_size"],
                                                     d["launch__registers_per_thread"]))
