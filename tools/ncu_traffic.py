#!/usr/bin/env python
"""Turn an .ncu-rep (one ncu --set full capture of a bench workload) into an entry of profiles/ncu_traffic.json.
Usage: python tools/ncu_traffic.py report.ncu-rep <op>_<size>_<frames> [kernel-name-substring]
Runs where ncu is installed; no GPU needed.  The entry is keyed by the sha of the kernel sources in the tree NOW, so run
it on the same tree the capture was taken on (bench.py prints null for `roofline.traffic` when the sha differs)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit])


def main():
    rep, key = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want in name and "synth" not in name:
            best = r
    rd = to_bytes(best[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
    wr = to_bytes(best[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(path))
    db[key] = {"dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "kernel": best[hdr.index("Kernel Name")],
               "gpu_time_under_ncu": best[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
               "l2_hit_pct": best[hdr.index("lts__t_sector_hit_rate.pct")], "kernel_src_sha": kernel_source_sha(),
               "source": os.path.basename(rep) + " (ncu --set full --clock-control none, one launch)"}
    json.dump(db, open(path, "w"), indent=1)
    print(key, db[key])


if __name__ == "__main__":
    main()
