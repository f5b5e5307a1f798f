"""Launch the fused set-abstraction kernel a few times at one level's geometry (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import synth  # noqa: E402
from demf_b200.mm import point_ops as ops  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
B = 8
geoms = [(20000, 2048, 0.2, 64, 1, (64, 64, 128)), (2048, 1024, 0.4, 32, 128, (128, 128, 256)),
         (1024, 512, 0.8, 16, 256, (128, 128, 256)), (512, 256, 1.2, 16, 256, (128, 128, 256)),
         (1024, 256, 0.3, 16, 256, (256, 256, 256))]
N, M, r, ns, C, widths = geoms[level]
pts = synth.make_points(B, 20000, seed=0, clustered=True)[..., :3].contiguous().to(dev)
x = pts[:, :N].contiguous()
c = ops.gather_rows(x, ops.furthest_point_sample(x, M)).contiguous()
f = torch.randn(B, N, C, device=dev)
cin = [ops.group_rows_width(C)] + list(widths[:-1])
ws = [torch.randn(co, ci, device=dev) / ci ** 0.5 for co, ci in zip(widths, cin)]
bs = [torch.randn(co, device=dev) * 0.1 for co in widths]
wpack, bias, wd = ops.sa_pack_mlp(ws, bs)
grid = ops.ball_grid(x, r) if N >= 4096 else None
for _ in range(reps):
    ops.sa_fused(x, c, f, 0.0, r, ns, True, wpack, bias, wd, grid=grid)
torch.cuda.synchronize()
print("done")
