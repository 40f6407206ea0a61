"""Per-launch DRAM traffic and duration of the kernels in a .ncu-rep as JSON (read by bench.py for roofline.traffic).

  python tools/ncu_to_json.py report.ncu-rep mesh_size out.json
"""
import csv
import json
import subprocess
import sys

rep, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def val(d, key):
    u = units[hdr.index(key)]
    v = float(d[key].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
             "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v * scale


kernels = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    kernels.append({
        "name": d.get("Kernel Name", "?"), "n": n,
        "dram_bytes_read": val(d, "dram__bytes_read.sum"), "dram_bytes_write": val(d, "dram__bytes_write.sum"),
        "duration_s": val(d, "gpu__time_duration.sum"),
        "registers": int(float(d["launch__registers_per_thread"])),
        "dram_pct": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
        "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
    })
json.dump({"report": rep, "kernels": kernels}, open(out, "w"), indent=1)
print(json.dumps(kernels, indent=1))
