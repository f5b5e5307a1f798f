import os, sys
sys.path.insert(0, "/root/repo")
import torch
from demf_b200 import synth
from demf_b200.mm import point_ops as ops
dev = torch.device("cuda:0")
x = synth.make_points(8, 20000, seed=0, clustered=True)[..., :3].contiguous().to(dev)
c = ops.gather_rows(x, ops.furthest_point_sample(x, 2048)).contiguous()
grid = ops.ball_grid(x, 0.2)
for _ in range(3):
    ops.ball_query_grid(0.0, 0.2, 64, x, c, grid)
torch.cuda.synchronize()
print("done")
