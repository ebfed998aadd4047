"""gpurun_out/conv_traffic_<tag>.csv (tools/conv_traffic.sh) -> profiles/<tag>_conv_traffic.json"""
import csv
import json
import sys

tag = sys.argv[1]
lines = [l for l in open("gpurun_out/conv_traffic_%s.csv" % tag) if not l.startswith("==")]
rd = wr = ns = 0.0
ids = set()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    m = row["Metric Name"]
    ids.add(row["ID"])
    if m == "dram__bytes_read.sum":
        rd += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif m == "dram__bytes_write.sum":
        wr += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif m == "gpu__time_duration.sum":
        ns += v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[u]
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/conv_traffic.sh %s), one pass of 8 receipts / "
                 "635 crops, every conv_tc_kernel launch" % tag,
       "launches": len(ids), "dram_read_bytes": rd, "dram_write_bytes": wr, "bytes_per_launch": (rd + wr) / max(len(ids), 1),
       "sum_duration_ms_under_ncu": ns / 1e6}
json.dump(out, open("profiles/%s_conv_traffic.json" % tag, "w"), indent=1)
print(json.dumps(out))
