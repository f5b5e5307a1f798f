"""Timeline of ONE replay of the graphed training step (torch.profiler / CUPTI kernel records): busy time per stream,
the union of busy intervals, idle gaps on the union, and the kernels that occupy the timeline. Development tool (GPU).
usage: python tools/train_timeline.py [out.md]"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(99)
model = engine.build_demf_votenet(num_points=4).to(dev).train()
trainer = engine.Trainer(model, capturable=True)
sets = [engine.synthetic_batch(4, 20000, "S512", seed=777 + i, device=dev) for i in range(2)]
for ts in sets:
    ts["gt_bboxes_3d"], ts["gt_labels_3d"] = engine.pad_gt(ts["gt_bboxes_3d"], ts["gt_labels_3d"], 16, dev)
for i in range(3):
    trainer.step(sets[i % 2])
gstep = engine.GraphedTrainStep(trainer, sets[0], max_gt=16)
for i in range(4):
    gstep(sets[i % 2], next_batch=sets[(i + 1) % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        gstep(sets[i % 2], next_batch=sets[(i + 1) % 2])
    torch.cuda.synchronize()
ks = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range.end > ev.time_range.start:
        ks.append((ev.time_range.start, ev.time_range.end, ev.name, getattr(ev, "device_resource_id", getattr(ev, "thread", 0))))
ks.sort()
if not ks:
    print("no kernel records")
    sys.exit(0)
# split into replays by the largest gaps: take the middle third of the records by time
t0, t1 = ks[0][0], max(k[1] for k in ks)
span = (t1 - t0) / 3.0
mid = [k for k in ks if t0 + span <= k[0] < t0 + 2 * span]
a, b = mid[0][0], max(k[1] for k in mid)
out = []
out.append(f"records {len(ks)} total, {len(mid)} in the middle replay window of {(b - a):.0f} us")
busy = defaultdict(float)
for s, e, n, st in mid:
    busy[st] += e - s
for st, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    out.append(f"stream {st}: busy {v:.0f} us ({100 * v / (b - a):.0f}% of the window), {sum(1 for k in mid if k[3] == st)} kernels")
# union of intervals and gaps
iv = sorted((s, e) for s, e, _, _ in mid)
union, gaps, cur_s, cur_e = 0.0, [], iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        union += cur_e - cur_s
        gaps.append(s - cur_e)
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
union += cur_e - cur_s
out.append(f"union busy {union:.0f} us; idle {sum(gaps):.0f} us in {len(gaps)} gaps (median {sorted(gaps)[len(gaps) // 2] if gaps else 0:.1f} us, "
           f"{sum(1 for g in gaps if g > 5)} gaps > 5 us totalling {sum(g for g in gaps if g > 5):.0f} us)")
# exclusive timeline share: at each instant, attribute to the longest-running active kernel's name
tot = defaultdict(float)
cnt = defaultdict(int)
for s, e, n, st in mid:
    tot[n] += e - s
    cnt[n] += 1
out.append("")
out.append("| total us | launches | avg us | kernel |")
out.append("|---|---|---|---|")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    out.append(f"| {v:.1f} | {cnt[n]} | {v / cnt[n]:.2f} | `{n[:110]}` |")
out.append("")
out.append("duration histogram of our GEMM kernels in this replay (us): launches, total")
for key in ("rows_gemm_kernel<true>", "rows_gemm_kernel<false>", "wgrad_kernel"):
    ds = sorted(e - s for s, e, n, st in mid if key in n)
    for lo, hi in ((0, 8), (8, 12), (12, 20), (20, 40), (40, 1e9)):
        sel = [d for d in ds if lo <= d < hi]
        out.append(f"  {key:26s} [{lo:>3.0f}, {hi if hi < 1e9 else float('inf'):>4.0f}): {len(sel):3d} launches, {sum(sel):7.1f} us")
text = "\n".join(out)
print(text)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        f.write("# graphed training step: one replay, warm (torch.profiler kernel records)\n\n" + text + "\n")
