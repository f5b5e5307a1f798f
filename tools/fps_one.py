"""Launch the grid-pruned FPS a few times at the SA1 geometry (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import synth  # noqa: E402
from demf_b200.mm import point_ops as ops  # noqa: E402

dev = torch.device("cuda:0")
x = synth.make_points(8, 20000, seed=0, clustered=True)[..., :3].contiguous().to(dev)
grid = ops.ball_grid(x, 0.2)
for _ in range(3):
    ops.furthest_point_sample_grid(x, 2048, grid)
torch.cuda.synchronize()
print("done")
