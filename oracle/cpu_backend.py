"""CPU execution of the DeMF(VoteNet) path with the oracle ops.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. PARITY UNPINNED (see oracle/demf_oracle.c).

`with oracle_ops():` swaps the CUDA-only op entry points of `demf_b200.mm.point_ops` and the
MSDA autograd function for CPU implementations built from the C oracle (oracle/cref.py: index
ops, same tie rules and rounding as the upstream kernels) plus plain torch indexing (so that
gradients flow for training-step checks). Inside the context the same nn.Modules that the
product runs on the GPU run on CPU tensors; this is (a) the whole-model checker for the GPU
parity tests and (b) the CPU baseline: upstream mmdet3d point ops have no CPU implementation at
all, so "the reference's CPU path" is this restatement for the point ops and
`multi_scale_deformable_attn_pytorch` (oracle/msda_torch.py, upstream's own CPU formulation) for
the attention core.
"""
import contextlib

import numpy as np
import torch

from . import cref
from .msda_torch import multi_scale_deformable_attn_pytorch


def _fps(points_xyz, num_points):
    return cref.furthest_point_sample(points_xyz.detach(), num_points)


def _ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    return cref.ball_query(min_radius, max_radius, sample_num, xyz.detach(), center_xyz.detach())


def _three_nn(target, source):
    return cref.three_nn(target.detach(), source.detach())


def _take(rows, idx):
    """rows (B,N,C), idx (B,...) -> (B,...,C) by torch indexing (differentiable)."""
    B = rows.shape[0]
    flat = idx.reshape(B, -1).long()
    out = torch.gather(rows, 1, flat.unsqueeze(-1).expand(-1, -1, rows.shape[-1]))
    return out.view(*idx.shape, rows.shape[-1])


def _query_and_group_rows(xyz, center_xyz, feat_rows, min_radius, max_radius, sample_num,
                          normalize_xyz, grid=None):
    idx = _ball_query(min_radius, max_radius, sample_num, xyz, center_xyz)
    diff = _take(xyz, idx) - center_xyz.unsqueeze(2)
    if normalize_xyz:  # the kernel multiplies by the float32 reciprocal of the radius
        diff = diff * float(np.float32(1.0) / np.float32(max_radius))
    B, M, ns, _ = diff.shape
    parts = []
    if feat_rows is not None:
        C = feat_rows.shape[-1]
        parts.append(_take(feat_rows, idx))
        pad = (-C) % 4
        if pad:
            parts.append(diff.new_zeros(B, M, ns, pad))
    parts += [diff, diff.new_zeros(B, M, ns, 1)]
    return idx, torch.cat(parts, -1)


def _three_interpolate_rows(feat_rows, idx, weight):
    f = _take(feat_rows, idx)                       # (B,n,3,C)
    w = weight.unsqueeze(-1)
    # upstream fma order: fma(w2,p2, fma(w0,p0, w1*p1)); torch has no fma, the sums below
    # differ from it by at most one rounding per term
    return (w[:, :, 1] * f[:, :, 1] + w[:, :, 0] * f[:, :, 0]) + w[:, :, 2] * f[:, :, 2]


def _grouping_operation(features, idx):
    return _take(features.transpose(1, 2), idx).permute(0, 3, 1, 2)


def _gather_points(features, idx):
    return _take(features.transpose(1, 2), idx).transpose(1, 2)


def _three_interpolate(features, idx, weight):
    return _three_interpolate_rows(features.transpose(1, 2), idx, weight).transpose(1, 2)


def _box_point_count(points, boxes, gravity_centre=False):
    from .postprocess import points_in_boxes_depth
    if gravity_centre:
        boxes = boxes.clone()
        boxes[..., 2] = boxes[..., 2] - boxes[..., 5] * 0.5
    return torch.stack([points_in_boxes_depth(points[b, :, :3], boxes[b]).sum(0).to(torch.int32)
                        for b in range(points.shape[0])])


def _aligned_3d_nms(minmax, scores, classes, valid, thresh):
    from .postprocess import aligned_3d_nms
    keep = torch.zeros_like(valid, dtype=torch.bool)
    for b in range(scores.shape[0]):
        inds = torch.nonzero(valid[b], as_tuple=False).flatten()
        picked = aligned_3d_nms(minmax[b][inds], scores[b][inds], classes[b][inds], thresh)
        keep[b, inds[picked]] = True
    return keep


def _nms_select(boxes, obj_scores, sem_scores, counts, min_points, nms_thr, score_thr):
    from .postprocess import box_corners_depth
    bc = boxes.clone()
    bc[..., 2] = bc[..., 2] - bc[..., 5] * 0.5
    corners = torch.stack([box_corners_depth(b) for b in bc])
    minmax = torch.cat([corners.min(2)[0], corners.max(2)[0]], -1)
    classes = torch.argmax(sem_scores, -1)
    keep = _aligned_3d_nms(minmax, obj_scores, classes, counts > min_points, nms_thr)
    selected = keep & (obj_scores > score_thr)
    return selected, classes, selected.sum(1).to(torch.int32)


class _MsdaCpu:
    """Stand-in for MultiScaleDeformableAttnFunction with the same .apply signature."""

    @staticmethod
    def apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
              im2col_step):
        return multi_scale_deformable_attn_pytorch(value, spatial_shapes, sampling_locations,
                                                   attention_weights)


@contextlib.contextmanager
def oracle_ops():
    from demf_b200.mm import ms_deform_attn as msda_mod
    from demf_b200.mm import point_ops as P
    patched = {
        "furthest_point_sample": _fps, "ball_query": _ball_query, "three_nn": _three_nn,
        "query_and_group_rows": _query_and_group_rows,
        "three_interpolate_rows": _three_interpolate_rows,
        "grouping_operation": _grouping_operation, "gather_points": _gather_points,
        "three_interpolate": _three_interpolate,
        "box_point_count": _box_point_count, "aligned_3d_nms": _aligned_3d_nms, "nms_select": _nms_select,
    }
    saved = {k: getattr(P, k) for k in patched}
    saved_fn = msda_mod.MultiScaleDeformableAttnFunction
    try:
        for k, v in patched.items():
            setattr(P, k, v)
        msda_mod.MultiScaleDeformableAttnFunction = _MsdaCpu
        yield
    finally:
        for k, v in saved.items():
            setattr(P, k, v)
        msda_mod.MultiScaleDeformableAttnFunction = saved_fn
