"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON/TXT: duration, pipe utilisation, DRAM traffic per launch."""
import csv
import io
import json
import subprocess
import sys

rep, out_json = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__inst_executed.sum": "warp_instructions",
}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]}
    for k, name in want.items():
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if u in scale and name in ("duration", "dram_read", "dram_write"):
                v *= scale[u]
            d[name] = v
    if "dram_read" in d and "dram_write" in d:
        d["dram_traffic_bytes"] = d["dram_read"] + d["dram_write"]
    out.append(d)
json.dump(out, open(out_json, "w"), indent=1)
for d in out:
    print(json.dumps(d))
