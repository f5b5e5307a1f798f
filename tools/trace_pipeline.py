"""CTA-level timeline of the instrumented kernels while ForwardPipeline runs (development tool)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import _lib, engine  # noqa: E402

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sets = [engine.synthetic_batch(8, 20000, "S512", seed=1234 + i, device=dev, with_gt=False) for i in range(max(lanes, 4))]
with torch.no_grad():
    for _ in range(3):
        model.simple_test(points=sets[0]["points"], img=sets[0]["img"], img_metas=sets[0]["img_metas"], nms=False)
pipe = engine.ForwardPipeline(model, sets, lanes=lanes)
for _ in range(8):
    pipe.submit()
pipe.join()
torch.cuda.synchronize()
cap = 400000
recs = torch.zeros(cap * 4, dtype=torch.int64, device=dev)   # 32-byte records
cnt = torch.zeros(1, dtype=torch.int32, device=dev)
lib = _lib.load()
lib.demf_trace_set(recs.data_ptr(), cnt.data_ptr(), cap)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    pipe.submit()
pipe.join()
b.record()
torch.cuda.synchronize()
lib.demf_trace_set(None, None, 0)
n = min(int(cnt.item()), cap)
raw = recs[:n * 4].cpu().numpy().view(np.int32).reshape(n, 8)
kern, block, smid = raw[:, 0], raw[:, 1], raw[:, 2]
t = recs[:n * 4].cpu().numpy().reshape(n, 4)
t0, t1 = t[:, 2].astype(np.int64), t[:, 3].astype(np.int64)
base = t0.min()
t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3   # us
print(f"lanes={lanes} steps={steps}: {a.elapsed_time(b) / steps:.3f} ms/step (traced), {n} CTA records, span {t1.max():.0f} us")
names = {1: "fps", 2: "sa_fused", 3: "ball_query_grid", 4: "ball_grid_build"}
for k, nm in names.items():
    m = kern == k
    if m.any():
        d = t1[m] - t0[m]
        print(f"  {nm:16s} CTAs {m.sum():7d}  mean {d.mean():8.1f} us  max {d.max():8.1f}  SM-time {d.sum() / 1e3:8.1f} SM*ms"
              f"  ({d.sum() / 1e3 / steps:.2f} per step)")
# occupancy over time: how many SMs host an FPS CTA / an SA CTA, sampled every 5 us
T = np.arange(0, t1.max(), 5.0)
def sm_busy(mask):
    out = np.zeros(len(T))
    for s in np.unique(smid[mask]):
        mm = mask & (smid == s)
        busy = np.zeros(len(T), dtype=bool)
        for x0, x1 in zip(t0[mm], t1[mm]):
            busy[int(x0 // 5):int(x1 // 5) + 1] = True
        out += busy
    return out
fps_sm = sm_busy(kern == 1)
sa_sm = sm_busy(kern == 2)
both = None
print(f"  SMs hosting >=1 FPS CTA: mean {fps_sm.mean():.1f}; >=1 SA CTA: mean {sa_sm.mean():.1f} (of {len(np.unique(smid))} SMs seen)")
# co-residency: SA CTAs whose interval overlaps an FPS CTA on the same SM
co = 0
tot = 0
for s in np.unique(smid):
    f = (kern == 1) & (smid == s)
    g = (kern == 2) & (smid == s)
    if not g.any():
        continue
    f0, f1 = t0[f], t1[f]
    for x0, x1 in zip(t0[g], t1[g]):
        tot += 1
        if f.any() and ((f0 < x1) & (f1 > x0)).any():
            co += 1
print(f"  SA CTAs that overlapped an FPS CTA on the same SM: {co} of {tot}")
np.savez_compressed("gpurun_out/trace_pipeline.npz", kern=kern, block=block, smid=smid, t0=t0, t1=t1)
# coarse timeline print: every 100 us the SM counts
for i in range(0, len(T), 20):
    print(f"   t={T[i]:7.0f} us  fps SMs {fps_sm[i]:5.0f}  sa SMs {sa_sm[i]:5.0f}")
    if i > 20 * 60:
        break
