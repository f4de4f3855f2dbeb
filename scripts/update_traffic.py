"""Writes profiles/traffic.json from `ncu --set full` reports of the fused kernel: python scripts/update_traffic.py 192x640=rep1.ncu-rep [375x1242=rep2.ncu-rep ...]

Per shape: dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of ONE fused_tile_kernel launch (B = 12,
T mode, white-noise flow, scripts/run_fused.py).  The file carries the hash of the kernel sources the reports were captured
from (bench.source_hash()); bench.py prints roofline.traffic / issue_roofline only when the hash matches the build it times.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def num(key):
        u, v = m[key]
        x = float(v.replace(",", ""))
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "inst": 1.0}.get(u, 1.0)
    return {"dram_bytes": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
            "warp_instructions": int(num("smsp__inst_executed.sum")),
            "kernel_us_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(m["gpu__time_duration.sum"][0], 1.0),
            "report": os.path.basename(rep)}


def main():
    out = {"source_hash": bench.source_hash(), "modes": ["T"],
           "source": "ncu --set full --clock-control none, one fused_tile_kernel launch, B=12, T mode, white-noise flow (scripts/run_fused.py); "
                     "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum, warp_instructions = smsp__inst_executed.sum"}
    for a in sys.argv[1:]:
        shape, rep = a.split("=", 1)
        out[shape] = metrics(rep)
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
