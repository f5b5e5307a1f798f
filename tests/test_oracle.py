"""CPU tests of the oracle itself (no GPU).

The reference ships no golden vectors for this path ("parity
unpinned"), so the oracle is pinned three ways instead: (1) hand-derived known
answers that follow from the upstream algorithm text; (2) agreement with an
independent second formulation (brute-force numpy for the point ops, the
grid_sample composition -- mmcv's own CPU path -- for MSDA); (3) the committed
fixtures in tests/golden reproduce bit for bit.
"""
import numpy as np
import pytest
import torch

from demf_b200 import synth
from oracle import cref, msda_torch, ops_numpy


# ----------------------------------------------------------------- known answers --
def test_fps_known_answer_line():
    # ten points on a line, N=10 -> upstream block of T=8 threads
    xyz = torch.zeros(1, 10, 3)
    xyz[0, :, 0] = torch.arange(10.0)
    # 0 -> 9 (farthest) -> 4 (4 and 5 tie at 16; slot 0 of the last tree level wins)
    assert cref.furthest_point_sample(xyz, 3).tolist() == [[0, 9, 4]]


def test_fps_tree_tie_rule_prefers_bit_reversed_thread_order():
    # k=1 and k=2 are equidistant from k=0. The shared-memory tree compares slot 0 (threads
    # 0,2) with slot 1 (threads 1,3) last and keeps slot 0 on a tie -> k=2, not argmax-first k=1.
    xyz = torch.tensor([[[0.0, 0, 0], [1.0, 0, 0], [0.0, 1.0, 0], [0.5, 0, 0]]])
    assert cref.furthest_point_sample(xyz, 2).tolist() == [[0, 2]]


def test_fps_m1_and_m_equals_n():
    xyz = torch.rand(2, 16, 3)
    assert cref.furthest_point_sample(xyz, 1).tolist() == [[0], [0]]
    idx = cref.furthest_point_sample(xyz, 16)
    assert sorted(idx[0].tolist()) == list(range(16))  # distinct points: a permutation


def test_ball_query_known_answer():
    xyz = torch.zeros(1, 6, 3)
    xyz[0, :, 0] = torch.tensor([0.0, 0.1, 0.2, 0.3, 5.0, 0.05])
    centre = torch.zeros(1, 2, 3)
    centre[0, 1, 0] = 100.0  # empty ball
    idx = cref.ball_query(0.0, 0.25, 4, xyz, centre)
    # hits in index order: 0 (d2==0), 1, 2, 5 -> stops at nsample
    assert idx[0, 0].tolist() == [0, 1, 2, 5]
    assert idx[0, 1].tolist() == [0, 0, 0, 0]
    idx = cref.ball_query(0.0, 0.25, 6, xyz, centre)
    assert idx[0, 0].tolist() == [0, 1, 2, 5, 0, 0]  # padded with the first hit
    # dilated ball: d2==0 is always accepted, [0.15, 0.25) keeps only x=0.2
    idx = cref.ball_query(0.15, 0.25, 3, xyz, centre)
    assert idx[0, 0].tolist() == [0, 2, 0]


def test_three_nn_known_answer():
    known = torch.tensor([[[0.0, 0, 0], [1.0, 0, 0], [0.0, 2.0, 0], [0.0, 0, 3.0], [1.0, 0, 0]]])
    unknown = torch.tensor([[[0.9, 0.0, 0.0]]])
    dist, idx = cref.three_nn(unknown, known)
    # duplicates 1 and 4 tie: strict '<' keeps the earlier index first
    assert idx.tolist() == [[[1, 4, 0]]]
    assert torch.allclose(dist, torch.tensor([[[0.1, 0.1, 0.9]]]), atol=1e-6)


def test_msda_known_answers():
    shapes = torch.tensor([[4, 6], [2, 3]])
    lsi = torch.tensor([0, 24])
    B, Q, H, D, L, P = 1, 3, 2, 4, 2, 2
    # (a) constant field: interior samples reproduce it, so out = const * sum(weights)
    value = torch.full((B, 30, H, D), 2.5)
    loc = torch.rand(B, Q, H, L, P, 2) * 0.5 + 0.25
    w = torch.rand(B, Q, H, L, P)
    out = cref.ms_deform_attn_forward(value, shapes, lsi, loc, w)
    expect = 2.5 * w.sum((-1, -2))  # (B,Q,H)
    assert torch.allclose(out.view(B, Q, H, D), expect[..., None].expand(-1, -1, -1, D), atol=1e-5)
    # (b) sampling exactly at a pixel centre returns that pixel
    value = torch.randn(B, 30, H, D)
    y, x = 2, 4
    loc = torch.zeros(B, 1, H, L, 1, 2)
    loc[..., 0, :, 0] = (x + 0.5) / 6
    loc[..., 0, :, 1] = (y + 0.5) / 4
    loc[..., 1, :, :] = 5.0  # level 1: far outside -> contributes nothing
    w = torch.ones(B, 1, H, L, 1)
    out = cref.ms_deform_attn_forward(value, shapes, lsi, loc, w)
    assert torch.allclose(out.view(B, 1, H, D), value[:, y * 6 + x][:, None], atol=1e-6)
    # (c) half a pixel outside the border: only the in-range corners count (zero padding)
    loc = torch.zeros(B, 1, H, L, 1, 2)  # (0,0) -> pixel coords (-0.5,-0.5): weight 1/4 on [0,0]
    loc[..., 1, :, :] = 5.0
    out = cref.ms_deform_attn_forward(value, shapes, lsi, loc, w)
    assert torch.allclose(out.view(B, 1, H, D), 0.25 * value[:, 0][:, None], atol=1e-6)


# ----------------------------------------------------------- second formulations --
@pytest.mark.parametrize("clustered", [False, True])
def test_point_ops_match_bruteforce(clustered):
    pts = synth.make_points(2, 3000, seed=3, clustered=clustered)[..., :3].contiguous()
    idx = cref.furthest_point_sample(pts, 200)
    assert np.array_equal(idx.numpy(), ops_numpy.furthest_point_sample(pts.numpy(), 200))
    centres = torch.gather(pts, 1, idx.long()[..., None].expand(-1, -1, 3)).contiguous()
    for r, ns in ((0.2, 64), (0.4, 32), (1.2, 16)):
        a = cref.ball_query(0.0, r, ns, pts, centres)
        b = ops_numpy.ball_query(0.0, r, ns, pts.numpy(), centres.numpy())
        assert np.array_equal(a.numpy(), b), (r, ns)
    dist, nidx = cref.three_nn(pts[:, :700].contiguous(), centres)
    d2, i2 = ops_numpy.three_nn(pts[:, :700].numpy(), centres.numpy())
    assert np.array_equal(nidx.numpy(), i2)
    assert np.allclose(dist.numpy(), d2, atol=1e-6)


def test_group_gather_interp_match_torch_indexing():
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 6, 100, generator=g)
    idx = torch.randint(0, 100, (2, 9, 4), generator=g, dtype=torch.int32)
    out = cref.grouping_operation(feat, idx)
    ref = torch.gather(feat[:, :, None].expand(-1, -1, 9, -1), 3,
                       idx.long()[:, None].expand(-1, 6, -1, -1))
    assert torch.equal(out, ref)
    go = torch.randn(2, 6, 9, 4, generator=g)
    feat_ag = feat.clone().requires_grad_()
    torch.gather(feat_ag[:, :, None].expand(-1, -1, 9, -1), 3,
                 idx.long()[:, None].expand(-1, 6, -1, -1)).backward(go)
    assert torch.allclose(cref.grouping_operation_backward(go, idx, 100), feat_ag.grad, atol=1e-5)
    w = torch.rand(2, 9, 3, generator=g)
    i3 = idx[:, :, :3].contiguous()
    out = cref.three_interpolate(feat, i3, w)
    ref = (torch.gather(feat[:, :, None].expand(-1, -1, 9, -1), 3,
                        i3.long()[:, None].expand(-1, 6, -1, -1)) * w[:, None]).sum(-1)
    assert torch.allclose(out, ref, atol=1e-6)


def test_query_and_group_matches_composition():
    pts = synth.make_points(2, 800, seed=1)[..., :3].contiguous()
    centres = pts[:, :50].contiguous()
    feat = torch.randn(2, 3, 800)
    idx, out = cref.query_and_group(pts, centres, feat, 0.0, 0.7, 8, True, True)
    assert torch.equal(idx, cref.ball_query(0.0, 0.7, 8, pts, centres))
    gx = cref.grouping_operation(pts.transpose(1, 2).contiguous(), idx)
    gx = (gx - centres.transpose(1, 2)[..., None]) * (torch.tensor(1.0) / torch.tensor(0.7))
    assert torch.equal(out[:, :3], gx)
    assert torch.equal(out[:, 3:], cref.grouping_operation(feat, idx))


@pytest.mark.parametrize("name,P", [("S512", 4), ("REAL", 2)])
def test_msda_matches_grid_sample_formulation(name, P):
    value, shapes, lsi, loc, attn = synth.make_msda_inputs(B=2, Q=64, name=name, P=P, seed=2)
    a = cref.ms_deform_attn_forward(value, shapes, lsi, loc, attn)
    b = msda_torch.multi_scale_deformable_attn_pytorch(value, shapes, loc, attn)
    assert (a - b).abs().max().item() < 2e-5  # two fp32 formulations; the parity budget is 1e-4
    c = msda_torch.multi_scale_deformable_attn_pytorch(value.double(), shapes, loc.double(),
                                                       attn.double())
    assert (a.double() - c).abs().max().item() < 2e-5


def test_msda_backward_matches_autograd_of_grid_sample_formulation():
    value, shapes, lsi, loc, attn = synth.make_msda_inputs(
        B=2, Q=20, shapes=((9, 12), (5, 6), (3, 3), (1, 2)), P=4, seed=4)
    value = value.double().requires_grad_()
    loc64 = loc.double().requires_grad_()
    attn64 = attn.double().requires_grad_()
    out = msda_torch.multi_scale_deformable_attn_pytorch(value, shapes, loc64, attn64)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(go.double())
    gv, gl, ga = cref.ms_deform_attn_backward(value.detach().float(), shapes, lsi, loc, attn, go)
    assert (gv.double() - value.grad).abs().max().item() < 1e-4
    assert (gl.double() - loc64.grad).abs().max().item() < 1e-3  # scaled by W,H up to 12
    assert (ga.double() - attn64.grad).abs().max().item() < 1e-4


# ------------------------------------------------------------------ golden files --
def test_oracle_reproduces_golden_fixtures(golden):
    z = golden("fps_n777_m64")
    assert torch.equal(cref.furthest_point_sample(z["xyz"], 64), z["idx"])
    z = golden("fps_dup_n300_m100")
    assert torch.equal(cref.furthest_point_sample(z["xyz"], 100), z["idx"])
    for name, ns in (("ball_query_r06_ns8", 8), ("ball_query_dilated", 16)):
        z = golden(name)
        got = cref.ball_query(z["min_radius"], z["max_radius"], ns, z["xyz"], z["new_xyz"])
        assert torch.equal(got, z["idx"])
    z = golden("group_c5")
    assert torch.equal(cref.grouping_operation(z["features"], z["idx"]), z["out"])
    z = golden("gather_c5")
    assert torch.equal(cref.gather_points(z["features"], z["idx"]), z["out"])
    z = golden("query_and_group")
    qidx, qout = cref.query_and_group(z["xyz"], z["new_xyz"], z["features"], 0.0, z["max_radius"], 8,
                                      True, True)
    assert torch.equal(qidx, z["idx"]) and torch.equal(qout, z["out"])
    z = golden("three_nn_interp")
    dist, idx = cref.three_nn(z["unknown"], z["known"])
    assert torch.equal(idx, z["idx"]) and torch.equal(dist, z["dist"])
    assert torch.equal(cref.three_interpolate(z["features"], z["idx"], z["weight"]), z["out"])
    for name in ("msda_mmcv_unit_shape", "msda_h8_d32_l4_p4", "msda_h8_d32_l4_p2"):
        z = golden(name)
        out = cref.ms_deform_attn_forward(z["value"], z["spatial_shapes"], z["level_start_index"],
                                          z["sampling_loc"], z["attn_weight"])
        assert torch.equal(out, z["out"])
        ref = msda_torch.multi_scale_deformable_attn_pytorch(
            z["value"], z["spatial_shapes"], z["sampling_loc"], z["attn_weight"])
        assert (out - ref).abs().max().item() < 1e-5


def test_fps_duplicate_points_golden_differs_from_index_order_rule(golden):
    # documents that the tie rule matters on duplicate points: the upstream tree order is NOT
    # "first maximum in index order", and the oracle follows the tree.
    z = golden("fps_dup_n300_m100")
    naive = ops_numpy.furthest_point_sample(z["xyz"].numpy(), 100)
    assert not np.array_equal(naive, z["idx"].numpy())


# ------------------------------------------------ set-abstraction module oracle --
def test_tf32_rna_emulation_bit_patterns():
    """cvt.rna.tf32.f32: 10 mantissa bits kept, round to nearest, ties AWAY from zero."""
    from oracle.sa_module import tf32_rna
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 - 2 ** -20, 1.0 + 3 * 2 ** -11, -1.0 - 2 ** -11,
                      0.0, 3.0e-39, 65504.0])
    y = tf32_rna(x)
    want = torch.tensor([1.0, 1.0 + 2 ** -10, 1.0, 1.0 + 2 ** -9, -1.0 - 2 ** -10, 0.0, 3.0e-39, 65504.0])
    # (the denormal only loses low bits; compare through the rounding rule)
    assert torch.equal(y[:6], want[:6]) and y[7] == want[7]
    assert (y.view(torch.int32) & 0x1FFF).abs().sum() == 0        # low 13 mantissa bits cleared
    r = torch.randn(10000) * 7
    assert ((tf32_rna(r) - r).abs() <= r.abs() * 2 ** -11 + 1e-45).all()


def test_sa_module_oracle_matches_the_mm_module_on_cpu():
    """Two formulations of one PointSAModule forward in eval mode: oracle/sa_module.py (grouped
    tensor in upstream channel order, folded (W, b) pairs) vs the package's own PointSAModule run
    on CPU through the oracle ops (rows layout, its own BN folding)."""
    from demf_b200.mm.bricks import _fold_conv_bn
    from demf_b200.mm.pointnet_modules import PointSAModule
    from oracle import sa_module
    from oracle.cpu_backend import oracle_ops
    torch.manual_seed(3)
    sa = PointSAModule(mlp_channels=[8, 32, 32, 64], num_point=40, radius=0.5, num_sample=16,
                       normalize_xyz=True).eval()
    for m in sa.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    xyz = synth.make_points(2, 600, seed=2)[..., :3].contiguous()
    feats = torch.randn(2, 8, 600)
    with oracle_ops(), torch.no_grad():
        new_xyz, out, idx = sa(xyz, feats)
    folded = [_fold_conv_bn(cm, None) for cm in sa.mlps[0]]
    ridx, ref = sa_module.sa_forward(xyz, new_xyz, feats, 0.0, 0.5, 16, True,
                                     [w for w, _ in folded], [b for _, b in folded])
    assert torch.equal(idx, cref.furthest_point_sample(xyz, 40))
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-5)
    _, ref32 = sa_module.sa_forward(xyz, new_xyz, feats, 0.0, 0.5, 16, True,
                                    [w for w, _ in folded], [b for _, b in folded], tf32=True)
    assert (ref32 - ref).abs().max() <= 1e-2 * ref.abs().max()     # TF32 operands: ~3 decimal digits
    assert not torch.equal(ref32, ref)
