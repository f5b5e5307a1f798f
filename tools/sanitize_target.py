"""Small invocations of the protocol-heavy kernels (mbarrier / TMEM / DSMEM / TMA) for compute-sanitizer:
fps_cluster_kernel, fps_grid_kernel, sa_fused_fwd_kernel, sa_pipe_kernel, ball_query_grid_kernel (bitmap selection),
rows_gemm_kernel (fwd + statistics, dgrad through BatchNorm), wgrad_kernel, col_sum_add_kernel, mha_fwd_kernel, msda fwd/bwd. Results are checked so that a tool-induced failure is visible.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import _lib, synth
from demf_b200.mm import point_ops as ops
from demf_b200.mm.bricks import permute_weight_columns
from demf_b200.mm.ms_deform_attn import MultiScaleDeformableAttnFunction as MSDA

dev = torch.device("cuda:0")
_lib.load()
only = set(sys.argv[1:])
want = lambda k: not only or k in only  # noqa: E731
pts = synth.make_points(2, 4096, seed=0, clustered=True)[..., :3].contiguous().to(dev)
if want("fps"):
    idx = ops.furthest_point_sample(pts[:, :1024].contiguous(), 128)
    grid = ops.ball_grid(pts, 0.4)
    gidx, gxyz = ops.furthest_point_sample_xyz(pts, 96, grid)
    ref = ops.furthest_point_sample(pts, 96)
    torch.cuda.synchronize()
    assert torch.equal(gidx, ref), "grid FPS != cluster FPS"
    print("fps ok", int(idx.sum()))
if want("sa"):
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(2, 4096, 16, generator=g).to(dev)
    centres = ops.gather_rows(pts, ops.furthest_point_sample(pts, 64)).contiguous()
    ws = [torch.randn(co, ci, generator=g).to(dev) / ci ** 0.5 for co, ci in ((64, 19), (64, 64), (128, 64))]
    bs = [torch.randn(co, generator=g).to(dev) * 0.1 for co in (64, 64, 128)]
    w0 = permute_weight_columns(ws[0], ops.group_rows_columns(16))
    wpack, bias, widths = ops.sa_pack_mlp([w0, ws[1], ws[2]], bs)
    out = ops.sa_fused(pts, centres, feats, 0.0, 0.4, 32, True, wpack, bias, widths, grid=ops.ball_grid(pts, 0.4))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and int(_lib.load().demf_sa_fused_error()) == 0
    print("sa_fused ok", float(out.abs().mean()))
if want("pipe"):
    g = torch.Generator().manual_seed(2)
    f1 = torch.randn(2, 4096, 1, generator=g).to(dev)
    centres = ops.gather_rows(pts, ops.furthest_point_sample(pts, 64)).contiguous()
    ws = [torch.randn(co, ci, generator=g).to(dev) / ci ** 0.5 for co, ci in ((64, 4), (64, 64), (128, 64))]
    bs = [torch.randn(co, generator=g).to(dev) * 0.1 for co in (64, 64, 128)]
    w0 = permute_weight_columns(ws[0], ops.group_rows_columns(1))
    wpack, bias, widths = ops.sa_pack_mlp([w0, ws[1], ws[2]], bs)
    grid = ops.ball_grid(pts, 0.4)
    nbr = ops.ball_query_grid(0.0, 0.4, 32, pts, centres, grid)
    assert torch.equal(nbr, ops.ball_query(0.0, 0.4, 32, pts, centres)), "bitmap grid query != scan"
    packed = torch.cat([pts, f1], -1).contiguous()
    a = ops.sa_pipe(pts, centres, f1, 0.4, 32, True, wpack, bias, nbr, packed=packed)
    b = ops.sa_fused(pts, centres, f1, 0.0, 0.4, 32, True, wpack, bias, widths, idx=nbr)
    torch.cuda.synchronize()
    assert ops.sa_pipe_error() == 0 and (a - b).abs().max() <= 1e-5 * b.abs().max()
    print("sa_pipe + ball_query_grid ok", float(a.abs().mean()))
if want("gemm"):
    g = torch.Generator(device=dev).manual_seed(1)
    R, K, N = 640, 64, 128
    x = torch.randn(R, K, generator=g, device=dev); w = torch.randn(N, K, generator=g, device=dev) / 8
    st = ops.bn_rows_state(N, dev)
    y = ops.gemm_rows_fwd(x, w, bn_state=st)
    mean, invstd = ops.bn_finalize(st, R, N, 1e-5, 0.1, None, None)
    dy = torch.randn(R, N, generator=g, device=dev)
    st2 = ops.bn_rows_state(K, dev)
    gm = ops.gemm_rows_dgrad_bn(dy, w, x, x.mean(0), 1.0 / x.std(0), torch.ones(K, device=dev), torch.zeros(K, device=dev), st2)
    dw = torch.zeros(N, K, device=dev); ops.gemm_wgrad_(dw, dy, x)
    db = torch.zeros(N, device=dev); ops.col_sum_add_(db, dy)
    torch.cuda.synchronize()
    assert (db - dy.sum(0)).abs().max() < 1e-3
    assert (y - x @ w.t()).abs().max() < 2e-2 and (dw - dy.t() @ x).abs().max() < 0.2 and ops.gemm_error() == 0
    assert torch.allclose(mean, y.mean(0), atol=1e-4)
    print("gemm ok", float(gm.abs().mean()))
if want("mha"):
    g = torch.Generator(device=dev).manual_seed(3)
    Lq, Lk, B, H, D = 100, 77, 2, 4, 36
    q = torch.randn(Lq * B, H * D, generator=g, device=dev)
    k = torch.randn(Lk * B, H * D, generator=g, device=dev)
    v = torch.randn(Lk * B, H * D, generator=g, device=dev)
    o = ops.mha_rows(q, k, v, B, H)
    hd = lambda t, L: t.reshape(L, B, H, D).permute(1, 2, 0, 3)  # noqa: E731
    ref = (torch.softmax(hd(q, Lq) @ hd(k, Lk).transpose(-1, -2) * D ** -0.5, -1) @ hd(v, Lk)).permute(2, 0, 1, 3)
    torch.cuda.synchronize()
    assert (o - ref.reshape(Lq * B, H * D)).abs().max() < 1e-4
    print("mha ok", float(o.abs().mean()))
if want("msda"):
    v, sh, lsi, loc, at = (t.to(dev) for t in synth.make_msda_inputs(B=1, Q=64, name="S512", P=4, seed=0))
    v.requires_grad_(True); loc.requires_grad_(True); at.requires_grad_(True)
    o = MSDA.apply(v, sh, lsi, loc, at, 64)
    o.sum().backward()
    torch.cuda.synchronize()
    print("msda ok", float(o.abs().mean()))
print("sanitize target done")
