"""Write profiles/decode_traffic.json from an `ncu --set full` capture of bench.py: DRAM bytes per decode launch, keyed by the
hash of the decode kernel's source so that bench.py only reports it for the kernel it was measured on.

usage: python tools/update_traffic.py <report.ncu-rep> [capture description]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (decode_source_sha)


def main():
    rep = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "decode_kernel" not in d["Kernel Name"]:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k]) * scale[units[hdr.index(k)]]
        rec = {"source_sha16": bench.decode_source_sha(), "dram_bytes_per_launch": tot,
               "dram_read_bytes": float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]],
               "dram_write_bytes": float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]],
               "kernel": d["Kernel Name"], "gpu_time_us": float(d["gpu__time_duration.sum"]),
               "inst_executed": float(d["smsp__inst_executed.sum"]),
               "capture": "ncu --set full --clock-control none, %s (dram__bytes_read.sum + dram__bytes_write.sum of one launch)" % what}
        with open(os.path.join(ROOT, "profiles", "decode_traffic.json"), "w") as f:
            json.dump(rec, f, indent=1)
        print(rec)
        return
    raise SystemExit("no decode_kernel launch in %s" % rep)


if __name__ == "__main__":
    main()
