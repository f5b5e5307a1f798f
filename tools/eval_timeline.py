"""Warm timeline of the benchmark's forward pipeline (8 captured forwards in flight, torch.profiler / CUPTI kernel
records over 32 steps): device busy fraction, per-kernel totals PER STEP. Development tool (GPU).
usage: python tools/eval_timeline.py [out.md]"""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
sets = [engine.synthetic_batch(8, 20000, "S512", seed=1234 + i, device=dev, with_gt=False) for i in range(8)]
with torch.no_grad():
    for i in range(3):
        model.simple_test(points=sets[i]["points"], img=sets[i]["img"], img_metas=sets[i]["img_metas"], nms=False)
pipe = engine.ForwardPipeline(model, sets, lanes=8)
for _ in range(16):
    pipe.submit()
pipe.join()
torch.cuda.synchronize()
STEPS = 32
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        pipe.submit()
    pipe.join()
    torch.cuda.synchronize()
ks = [(ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
      if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range.end > ev.time_range.start]
ks.sort()
t0, t1 = ks[0][0], max(k[1] for k in ks)
iv = sorted((s, e) for s, e, _ in ks)
union, cs, ce = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > ce:
        union += ce - cs
        cs, ce = s, e
    else:
        ce = max(ce, e)
union += ce - cs
tot, cnt = defaultdict(float), defaultdict(int)
for s, e, n in ks:
    tot[n] += e - s
    cnt[n] += 1
out = [f"{STEPS} steps in {(t1 - t0):.0f} us = {(t1 - t0) / STEPS:.1f} us per step (under the profiler); device busy "
       f"{100 * union / (t1 - t0):.1f} %; kernel time summed over all streams {sum(tot.values()) / STEPS:.0f} us per step "
       f"({sum(cnt.values()) / STEPS:.0f} kernels)", "",
       "| us per step | launches per step | avg us | kernel |", "|---|---|---|---|"]
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    out.append(f"| {v / STEPS:.1f} | {cnt[n] / STEPS:.1f} | {v / cnt[n]:.2f} | `{n[:120]}` |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        f.write("# forward pipeline, warm: kernel time per step with 8 batches in flight (durations are CONCURRENT-execution\n"
                "# durations: a kernel sharing the GPU with seven other batches runs longer than alone)\n\n" + text + "\n")
