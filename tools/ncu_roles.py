"""Summarise an ncu report of one kernel (--set full --import-source on): issue utilisation, stall mix and, from the
source page, executed instructions / stall samples between marker instructions (waits, tensor-memory loads, MMAs).
usage: python tools/ncu_roles.py gpurun_out/r2_sa_pipe2.ncu-rep"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__icc_request_hit_rate.pct"]
for h, v in zip(hdr, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                     and float(v or 0) > 0.3):
        print(f"{h:75s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ie, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot_i = sum(int(r[ie] or 0) for r in data)
tot_s = sum(int(r[ismp] or 0) for r in data)
print("instructions", tot_i, "samples", tot_s, "sass lines", len(data))
prev = 0
for i, r in enumerate(data):
    if re.search(r"UTCHMMA|LDTM|UTCBAR|BAR.SYNC|TRYWAIT|MEMBAR|EXIT|STG", r[isrc]) and int(r[ie] or 0) > 0:
        seg_i = sum(int(data[k][ie] or 0) for k in range(prev, i + 1))
        seg_s = sum(int(data[k][ismp] or 0) for k in range(prev, i + 1))
        if seg_i > 0.004 * tot_i or seg_s > 0.004 * tot_s:
            print(f"{i:5d} inst {100 * seg_i / tot_i:5.1f}%  samples {100 * seg_s / tot_s:5.1f}%  {r[ie]:>8s}  {r[isrc][:60]}")
        prev = i + 1
