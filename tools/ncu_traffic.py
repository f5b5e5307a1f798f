"""DRAM bytes per launch of the MSDA kernel from `ncu --set full` reports -> profiles/msda_fwd_traffic.json,
the file bench.py's `roofline.traffic` / `roofline_hbm[*].traffic` read (runs anywhere; no GPU).

    python tools/ncu_traffic.py key=report.ncu-rep[:launch_index] ...

e.g.  dram_bytes_per_launch=gpurun_out/r2_msda_s512.ncu-rep xl_q256_dram_bytes_per_launch=gpurun_out/r2_msda_xl.ncu-rep:0
The file records the sha of csrc/msda.cu the capture was taken with; bench.py ignores a stale capture.
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(rep, which=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(vals[i].replace(",", "")) * UNIT[units[i]]
        res.append((vals[hdr.index("Kernel Name")], tot, float(vals[hdr.index("gpu__time_duration.sum")].replace(",", ""))))
    if which is not None:
        res = [res[which]]
    return res


def main():
    with open(os.path.join(ROOT, "demf_b200", "csrc", "msda.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()[:16]
    path = os.path.join(ROOT, "profiles", "msda_fwd_traffic.json")
    doc = {}
    if os.path.exists(path):
        with open(path) as f:
            doc = json.load(f)
    if doc.get("msda_cu_sha16") != sha:
        doc = {"msda_cu_sha16": sha, "sources": {}}
    for arg in sys.argv[1:]:
        key, _, spec = arg.partition("=")
        rep, _, idx = spec.partition(":")
        launches = dram_bytes(rep, int(idx) if idx else None)
        doc[key] = sum(b for _, b, _ in launches) / len(launches)
        doc["sources"][key] = {"report": os.path.basename(rep), "launches": len(launches),
                               "kernel": launches[0][0][:80], "ncu_time_us": [round(t, 2) for _, _, t in launches]}
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
