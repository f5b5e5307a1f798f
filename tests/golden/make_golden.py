"""Generates tests/golden/*.npz from the CPU oracle on small seeded inputs.

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference (haoy945/DeMF) ships no fixtures and its native ops
(mmdet3d 0.18.1 / mmcv-full 1.3.18) cannot be imported here, so these vectors pin
the ORACLE (oracle/demf_oracle.c), not the reference. Every file stores the inputs
as well as the outputs, so the GPU tests do not depend on RNG reproducibility.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                           for k, v in arrays.items()})
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


def main():
    g = torch.Generator().manual_seed(20261017)
    rnd = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    rndn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731

    # ---- FPS: N not a power of two (T=512 threads upstream), and a duplicate-point set that
    # exercises the block-tree tie rule.
    xyz = rnd(2, 777, 3) * torch.tensor([6.0, 6.0, 2.5])
    save("fps_n777_m64", xyz=xyz, idx=cref.furthest_point_sample(xyz, 64))
    base = rnd(1, 100, 3)
    dup = torch.cat([base, base, base], 1)[:, torch.randperm(300, generator=g)].contiguous()
    save("fps_dup_n300_m100", xyz=dup, idx=cref.furthest_point_sample(dup, 100))

    # ---- ball query: ordinary, empty balls (far-away centres), and min_radius > 0
    xyz = rnd(2, 500, 3) * 2.0
    new_xyz = xyz[:, :32].clone()
    new_xyz[:, -4:] += 50.0  # nothing in range -> rows stay zero
    save("ball_query_r06_ns8", xyz=xyz, new_xyz=new_xyz, min_radius=0.0, max_radius=0.6,
         idx=cref.ball_query(0.0, 0.6, 8, xyz, new_xyz))
    save("ball_query_dilated", xyz=xyz, new_xyz=new_xyz, min_radius=0.3, max_radius=0.7,
         idx=cref.ball_query(0.3, 0.7, 16, xyz, new_xyz))

    # ---- grouping / gather forward + backward
    feat = rndn(2, 5, 500)
    idx = cref.ball_query(0.0, 0.6, 8, xyz, new_xyz)
    go = rndn(2, 5, 32, 8)
    save("group_c5", features=feat, idx=idx, out=cref.grouping_operation(feat, idx), grad_out=go,
         grad_features=cref.grouping_operation_backward(go, idx, 500))
    gidx = torch.randint(0, 500, (2, 40), generator=g, dtype=torch.int32)
    go = rndn(2, 5, 40)
    save("gather_c5", features=feat, idx=gidx, out=cref.gather_points(feat, gidx), grad_out=go,
         grad_features=cref.gather_points_backward(go, gidx, 500))

    # ---- fused QueryAndGroup (use_xyz, normalize_xyz as in demf_votenet.py:58-62)
    qidx, qout = cref.query_and_group(xyz, new_xyz, feat, 0.0, 0.6, 8, True, True)
    save("query_and_group", xyz=xyz, new_xyz=new_xyz, features=feat, max_radius=0.6, idx=qidx,
         out=qout)

    # ---- three_nn / three_interpolate
    unknown = rnd(2, 60, 3)
    known = rnd(2, 25, 3)
    dist, nidx = cref.three_nn(unknown, known)
    w = 1.0 / (dist + 1e-8)
    w = w / w.sum(2, keepdim=True)
    kfeat = rndn(2, 7, 25)
    go = rndn(2, 7, 60)
    save("three_nn_interp", unknown=unknown, known=known, dist=dist, idx=nidx, weight=w,
         features=kfeat, out=cref.three_interpolate(kfeat, nidx, w), grad_out=go,
         grad_features=cref.three_interpolate_backward(go, nidx, w, 25))

    # ---- MSDA: the shape of mmcv's own unit test (N,M,D=1,2,2; Lq,L,P=2,2,2; (6,4),(3,2)) ...
    def msda_case(name, B, Q, H, D, shapes, P, spread):
        shapes_t = torch.tensor(shapes, dtype=torch.int64)
        lsi = torch.cat([shapes_t.new_zeros(1), shapes_t.prod(1).cumsum(0)[:-1]])
        S = int(shapes_t.prod(1).sum())
        L = len(shapes)
        value = rndn(B, S, H, D) * 0.5
        loc = (rnd(B, Q, H, L, P, 2) * (1 + 2 * spread) - spread).contiguous()
        attn = rnd(B, Q, H, L, P) + 1e-5
        attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).contiguous()
        go = rndn(B, Q, H * D)
        out = cref.ms_deform_attn_forward(value, shapes_t, lsi, loc, attn)
        gv, gl, ga = cref.ms_deform_attn_backward(value, shapes_t, lsi, loc, attn, go)
        save(name, value=value, spatial_shapes=shapes_t, level_start_index=lsi, sampling_loc=loc,
             attn_weight=attn, out=out, grad_out=go, grad_value=gv, grad_sampling_loc=gl,
             grad_attn_weight=ga)

    msda_case("msda_mmcv_unit_shape", 1, 2, 2, 2, [(6, 4), (3, 2)], 2, 0.0)
    # ... and the DeMF head geometry (H=8, D=32, L=4) with samples falling off every border
    msda_case("msda_h8_d32_l4_p4", 2, 12, 8, 32, [(8, 10), (4, 5), (2, 3), (1, 2)], 4, 0.25)
    msda_case("msda_h8_d32_l4_p2", 1, 9, 8, 32, [(7, 9), (4, 5), (2, 3), (1, 1)], 2, 0.25)


if __name__ == "__main__":
    main()
