"""Per-phase clock stamps of the fused set-abstraction kernel (development tool, GPU only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import _lib, synth  # noqa: E402
from demf_b200.mm import point_ops as ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
B = 8
pts = synth.make_points(B, 20000, seed=0, clustered=True)[..., :3].contiguous().to(dev)
geoms = [(20000, 2048, 0.2, 64, 1, (64, 64, 128)), (2048, 1024, 0.4, 32, 128, (128, 128, 256)),
         (1024, 512, 0.8, 16, 256, (128, 128, 256)), (1024, 256, 0.3, 16, 256, (256, 256, 256))]
names = ["gather", "wait0", "epi0", "wait1", "epi1", "wait2", "epi2"]
for N, M, r, ns, C, widths in geoms:
    x = pts[:, :N].contiguous()
    c = ops.gather_rows(x, ops.furthest_point_sample(x, M)).contiguous()
    f = torch.randn(B, N, C, device=dev)
    cin = [ops.group_rows_width(C)] + list(widths[:-1])
    ws = [torch.randn(co, ci, device=dev) / ci ** 0.5 for co, ci in zip(widths, cin)]
    bs = [torch.randn(co, device=dev) * 0.1 for co in widths]
    wpack, bias, wd = ops.sa_pack_mlp(ws, bs)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    grid = ops.ball_grid(x, r) if N >= 4096 else None
    ops.sa_fused(x, c, f, 0.0, r, ns, True, wpack, bias, wd, grid=grid)
    lib.demf_sa_fused_set_profile(buf.data_ptr())
    ops.sa_fused(x, c, f, 0.0, r, ns, True, wpack, bias, wd, grid=grid)
    torch.cuda.synchronize()
    lib.demf_sa_fused_set_profile(None)
    st = buf.cpu().tolist()
    n, t = st[0], st[1:]
    print(f"N={N} M={M} ns={ns} C={C} widths={widths}: query {t[1] - t[0]} clk")
    nch0 = (((C + 3) // 4 * 4 + 4 + 7) // 8 * 8 + 31) // 32
    print("   stamps:", n, " deltas:", [t[i] - t[i - 1] for i in range(2, n)], " total", t[n - 1] - t[0])
    iss = st[64:]
    base = t[0]
    ev = [(iss[3 * i], iss[3 * i + 1] - base, iss[3 * i + 2] - iss[3 * i + 1]) for i in range(60) if iss[3 * i + 1]]
    if ev:
        print("   issuer (lane*100+pos, t_detect-t0, issue clk):", ev[:40])
        print("   worker stamps rel t0:", [x - base for x in t[:n]])
