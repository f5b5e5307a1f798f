"""sa_fused vs sa_pipe on the first backbone level (B=8, 20000 -> 2048, r 0.2, ns 64), same ball-query rows."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import synth
from demf_b200.mm import point_ops as ops
dev = torch.device("cuda:0")
B, N, M, r, ns = 8, 20000, 2048, 0.2, 64
pts = synth.make_points(B, N, seed=0, clustered=True).to(dev)
xyz = pts[..., :3].contiguous(); f = pts[..., 3:].contiguous()
c = ops.gather_rows(xyz, ops.furthest_point_sample(xyz, M)).contiguous()
g = torch.Generator(device=dev).manual_seed(0)
ws = [torch.randn(co, ci, generator=g, device=dev) / ci ** 0.5 for co, ci in ((64, 8), (64, 64), (128, 64))]
bs = [torch.randn(co, generator=g, device=dev) * 0.1 for co in (64, 64, 128)]
wpack, bias, wd = ops.sa_pack_mlp(ws, bs)
grid = ops.ball_grid(xyz, r)
nbr = ops.ball_query_grid(0.0, r, ns, xyz, c, grid)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        junk.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
a = ops.sa_fused(xyz, c, f, 0.0, r, ns, True, wpack, bias, wd, idx=nbr)
b = ops.sa_pipe(xyz, c, f, r, ns, True, wpack, bias, nbr)
torch.cuda.synchronize()
print("max |diff|", (a - b).abs().max().item(), "scale", a.abs().max().item(), "errors", ops.sa_pipe_error())
print("sa_fused (given rows) ms", t(lambda: ops.sa_fused(xyz, c, f, 0.0, r, ns, True, wpack, bias, wd, idx=nbr)))
print("sa_pipe  (given rows) ms", t(lambda: ops.sa_pipe(xyz, c, f, r, ns, True, wpack, bias, nbr)))
pk = pts.contiguous()
b2 = ops.sa_pipe(xyz, c, f, r, ns, True, wpack, bias, nbr, packed=pk)
print("packed identical", torch.equal(b, b2))
print("sa_pipe  (packed rows) ms", t(lambda: ops.sa_pipe(xyz, c, f, r, ns, True, wpack, bias, nbr, packed=pk)))
print("ball_query_grid ms", t(lambda: ops.ball_query_grid(0.0, r, ns, xyz, c, grid)))

from demf_b200 import _lib
lib = _lib.load()
if hasattr(lib, "demf_sa_pipe_profile"):   # only in a -DDEMF_SAP_PROF build (DEMF_NVCC_EXTRA)
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    lib.demf_sa_pipe_profile(buf)
    ops.sa_pipe(xyz, c, f, r, ns, True, wpack, bias, nbr)
    lib.demf_sa_pipe_profile(buf)
    names = {0: "total", 1: "L0 wait a0_full", 2: "L0 wait d0_free", 3: "L1 wait a1_full", 4: "L1 wait d1_free",
             5: "L2 wait a2_full", 6: "L2 wait d2_free", 7: "gather wait a0_free", 8: "epi0 wait d0_full",
             9: "epi1 wait d1_full", 10: "epi0 wait a1_free", 11: "epi1 wait a2_free", 12: "max wait d2_full",
             15: "tiles"}
    for k in sorted(names):
        print("%-22s %10d  per tile %8.0f" % (names[k], buf[k], buf[k] / max(buf[15], 1)))
    for k, nm in enumerate(["epi0: tmem ld", "epi0: + store loop", "epi0: + fence/arrive", "L1 issue 8 MMA + 2 commits",
                            "max: whole tile"]):
        print("%-28s per tile %8.0f" % (nm, buf[16 + k] / max(buf[15], 1)))
