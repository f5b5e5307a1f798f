"""Which Python lines launch the small library kernels of an eval forward? (development tool, GPU)"""
import os
import sys
from collections import Counter

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
batch = engine.synthetic_batch(8, 20000, "S512", seed=1234, device=dev, with_gt=False)
from demf_b200.mm import geometry  # noqa: E402
mats, affs = geometry.fold_projection(batch["img_metas"])
proj = (mats.to(dev), affs.to(dev))
with torch.no_grad():
    for _ in range(3):
        model.simple_test(points=batch["points"], img=batch["img"], img_metas=batch["img_metas"], projection=proj, nms=False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        model.simple_test(points=batch["points"], img=batch["img"], img_metas=batch["img_metas"], projection=proj, nms=False)
        torch.cuda.synchronize()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sites = Counter()
kernels = Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        continue
    n_k = len([k for k in ev.kernels]) if hasattr(ev, "kernels") else 0
    if n_k == 0:
        continue
    where = "?"
    for fr in (ev.stack or []):
        if "demf_b200" in fr and "site-packages" not in fr:
            where = fr.replace(root + "/", "")
            break
    sites[(where.split(" ")[0] if where != "?" else "?", ev.name)] += n_k
    kernels[ev.name] += n_k
print("---- launching ops in order (name, input shapes)")
seq = [ev for ev in prof.events() if ev.device_type != torch.autograd.DeviceType.CUDA and getattr(ev, "kernels", None)]
seq.sort(key=lambda e: e.time_range.start)
for ev in seq:
    print(f"  {ev.name:32s} {str(ev.input_shapes)[:110]}")
print("kernel launches by op:", sum(kernels.values()))
for k, v in kernels.most_common(25):
    print(f"  {v:4d}  {k}")
print("by call site:")
for (w, name), v in sorted(sites.items(), key=lambda kv: -kv[1])[:70]:
    print(f"  {v:3d}  {name:28s} {w}")
