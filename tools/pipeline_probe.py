"""Where does a forward's time go when several batches are in flight? (development tool, GPU only)

Times ForwardPipeline at the bench shapes for several lane counts, with the real sampling chain and
with furthest-point sampling replaced by cached indices (no FPS kernels in the graphs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import engine  # noqa: E402
from demf_b200.mm import point_ops as P  # noqa: E402

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sets = [engine.synthetic_batch(B, 20000, "S512", seed=1234 + i, device=dev, with_gt=False) for i in range(16)]


def run(lanes, steps=40, tag=""):
    pipe = engine.ForwardPipeline(model, sets[:max(lanes, 4)], lanes=lanes)
    for _ in range(8):
        pipe.submit()
    pipe.join()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        pipe.submit()
    pipe.join()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(f"{tag} lanes={lanes}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} scenes/s", flush=True)
    del pipe


with torch.no_grad():
    for _ in range(3):
        model.simple_test(points=sets[0]["points"], img=sets[0]["img"], img_metas=sets[0]["img_metas"], nms=False)
if os.environ.get("PROBE_SERIAL_FPS"):
    model.pts_backbone.overlap_sampling = False
quick = os.environ.get("PROBE_QUICK")
for lanes in ((4, 8, 12, 16) if quick else (1, 2, 4, 6, 8)):
    run(lanes, tag="real FPS  ")
if quick:
    sys.exit(0)

# cached indices: the same forward without any furthest-point-sampling kernel
real_fps = P.furthest_point_sample
cache = {}


def cached_fps(xyz, m):
    key = (xyz.shape[1], m)
    if key not in cache:
        cache[key] = real_fps(xyz, m)
    return cache[key]


real_fps_grid = P.furthest_point_sample_grid
P.furthest_point_sample_grid = lambda xyz, m, grid: cached_fps(xyz, m)
P.furthest_point_sample = cached_fps
with torch.no_grad():
    model.simple_test(points=sets[0]["points"], img=sets[0]["img"], img_metas=sets[0]["img_metas"], nms=False)
for lanes in (1, 2, 4, 8):
    run(lanes, tag="cached FPS")
# ... and without the fused set-abstraction kernels either (cached level outputs): what the
# vote module, FP modules, head and decoder cost on their own
real_sa = P.sa_fused
sa_cache = {}


def cached_sa(xyz, center_xyz, feat_rows, *a, **k):
    key = (xyz.shape[1], center_xyz.shape[1])
    if key not in sa_cache:
        sa_cache[key] = real_sa(xyz, center_xyz, feat_rows, *a, **k)
    return sa_cache[key]


P.sa_fused = cached_sa
with torch.no_grad():
    model.simple_test(points=sets[0]["points"], img=sets[0]["img"], img_metas=sets[0]["img_metas"], nms=False)
for lanes in (1, 4, 8):
    run(lanes, tag="cached FPS + cached SA")
P.sa_fused = real_sa
P.furthest_point_sample = real_fps
P.furthest_point_sample_grid = real_fps_grid
