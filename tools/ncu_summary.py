"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) as a
markdown table: per kernel share, total, launches, average. Development tool (runs anywhere).

    python tools/ncu_summary.py gpurun_out/launches.csv [passes] > profiles/rX_launches.md
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    passes = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        tot[r["Kernel Name"]] += us
        cnt[r["Kernel Name"]] += 1
    total = sum(tot.values())
    n = sum(cnt.values())
    demf = sum(v for k, v in tot.items() if "demf::" in k)
    print(f"# ncu launch list summary ({path})\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache "
          "and serialised: read the SHARES, not the absolutes.\n")
    line = f"{n} launches, {total:.0f} us summed; demf:: kernels {demf:.0f} us ({100 * demf / total:.1f}%)"
    if passes:
        line += f"; {passes:g} passes captured -> {total / passes:.0f} us and {n / passes:.0f} launches per pass"
    print(line + "\n")
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / total < 0.002:
            continue
        print(f"| {100 * v / total:.1f}% | {v:.1f} | {cnt[k]} | {v / cnt[k]:.2f} | `{k[:120]}` |")


if __name__ == "__main__":
    main()
