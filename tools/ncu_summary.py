#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box into small, tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md [--steps N]
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/r1_msda_fwd.json [--traffic-json profiles/msda_fwd_traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys


def launches(src, dst, forwards=None):
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = row["Kernel Name"]
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    total = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if "demf::" in k)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are "
                "cold-cache and serialised: read the SHARES, not the absolutes.\n\n")
        f.write(f"{n} launches, {total:.0f} us summed; demf:: kernels {mine:.0f} us "
                f"({100 * mine / total:.1f}%)")
        if forwards:
            f.write(f"; ~{forwards} forward passes captured -> {total / forwards:.0f} us and "
                    f"{n / forwards:.0f} launches per pass")
        f.write("\n\n| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"| {100 * v[1] / total:.1f}% | {v[1]:.1f} | {v[0]} | {v[1] / v[0]:.2f} | `{k[:120]}` |\n")
    print(f"wrote {dst}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def kernel(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{row[i]} {units[i]}".strip()
        res.append(d)
    with open(dst, "w") as f:
        json.dump({"source": src, "how": "ncu --set full --clock-control none --import-source on",
                   "launches": res}, f, indent=1)
    print(f"wrote {dst}")
    if traffic_json:
        def mb(s):
            val, unit = s.split()
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            return float(val) * scale
        tr = [mb(d["dram__bytes_read.sum"]) + mb(d["dram__bytes_write.sum"]) for d in res]
        with open(traffic_json, "w") as f:
            json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr),
                       "source": dst}, f, indent=1)
        print(f"wrote {traffic_json}")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    rest = sys.argv[4:]
    if mode == "launches":
        launches(src, dst, float(rest[1]) if rest[:1] == ["--steps"] else None)
    else:
        kernel(src, dst, rest[1] if rest[:1] == ["--traffic-json"] else None)
