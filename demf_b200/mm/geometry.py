"""Coordinate helpers of mmdet3d that the DeMF head calls, batched for the GPU.

Stands in for `apply_3d_transformation`, `coord_2d_transform` (mmdet3d
models/fusion_layers/coord_transform.py), `points_cam2img`, `rotation_3d_in_axis`
(core/bbox/structures/utils.py) and `DepthInstance3DBoxes.points_in_boxes` -- reached from
demf/modeling/heads/class_agnostic_vote_head.py:524-547 (reference points) and :834,:903-906
(target assignment). Upstream loops over scenes in Python and launches a dozen tiny kernels per
scene; here the per-scene chain "undo 3D augmentation -> depth2img projection -> 2D
augmentation -> normalise" is folded ON THE HOST (float64) into one 3x4 matrix and a 2-D affine
per scene, so the device work for the whole batch is one bmm + a handful of element-wise ops.
"""
import math

import numpy as np
import torch


# ---------------------------------------------------------------- per-scene reference ---
def apply_3d_transformation(pcd, coord_type, img_meta, reverse=False):
    """Apply (or undo, reverse=True) the recorded 3D augmentation flow to pcd (N,3)."""
    assert coord_type == 'DEPTH', "the DeMF path only uses depth coordinates"
    dtype, device = pcd.dtype, pcd.device
    rot = (torch.tensor(img_meta['pcd_rotation'], dtype=dtype, device=device)
           if 'pcd_rotation' in img_meta else torch.eye(3, dtype=dtype, device=device))
    scale = img_meta.get('pcd_scale_factor', 1.)
    trans = (torch.tensor(img_meta['pcd_trans'], dtype=dtype, device=device)
             if 'pcd_trans' in img_meta else torch.zeros(3, dtype=dtype, device=device))
    hflip = img_meta.get('pcd_horizontal_flip', False)
    vflip = img_meta.get('pcd_vertical_flip', False)
    flow = list(img_meta.get('transformation_3d_flow', []))
    pcd = pcd.clone()
    if reverse:
        scale, trans, rot, flow = 1.0 / scale, -trans, rot.inverse(), flow[::-1]
    for op in flow:
        assert op in ('T', 'S', 'R', 'HF', 'VF'), f'This 3D data transformation op ({op}) is not supported'
        if op == 'T':
            pcd[:, :3] = pcd[:, :3] + trans
        elif op == 'S':
            pcd[:, :3] = pcd[:, :3] * scale
        elif op == 'R':
            pcd[:, :3] = pcd[:, :3] @ rot
        elif op == 'HF' and hflip:
            pcd[:, 0] = -pcd[:, 0]
        elif op == 'VF' and vflip:
            pcd[:, 1] = -pcd[:, 1]
    return pcd


def points_cam2img(points_3d, proj_mat, with_depth=False):
    """Project (…,3) points with a 3x3 / 3x4 / 4x4 matrix; returns (…,2) pixel coordinates."""
    points_shape = list(points_3d.shape)
    points_shape[-1] = 1
    assert len(proj_mat.shape) == 2
    d1, d2 = proj_mat.shape[:2]
    assert (d1 == 3 and d2 == 3) or (d1 == 3 and d2 == 4) or (d1 == 4 and d2 == 4)
    if d1 == 3:
        expanded = torch.eye(4, device=proj_mat.device, dtype=proj_mat.dtype)
        expanded[:d1, :d2] = proj_mat
        proj_mat = expanded
    points_4 = torch.cat([points_3d, points_3d.new_ones(points_shape)], dim=-1)
    point_2d = points_4 @ proj_mat.T
    res = point_2d[..., :2] / point_2d[..., 2:3]
    if with_depth:
        res = torch.cat([res, point_2d[..., 2:3]], dim=-1)
    return res


def coord_2d_transform(img_meta, coord_2d, ori2new, img_shape=None):
    """Apply the recorded 2D augmentation (scale, crop offset, flip) to (N,2k) pixel coords."""
    img_h, img_w = (img_meta['img_shape'][:2] if img_shape is None else img_shape)
    scale = img_meta['scale_factor'][:2] if 'scale_factor' in img_meta else [1.0, 1.0]
    flip = img_meta.get('flip', False)
    crop = img_meta.get('img_crop_offset', [0.0, 0.0])
    coord_2d = coord_2d.clone()
    if ori2new:
        coord_2d[:, 0::2] = coord_2d[:, 0::2] * scale[0]
        coord_2d[:, 1::2] = coord_2d[:, 1::2] * scale[1]
        coord_2d[:, 0::2] += crop[0]
        coord_2d[:, 1::2] += crop[1]
        if flip:
            coord_2d[:, 0::2] = img_w - coord_2d[:, 0::2]
    else:
        if flip:
            coord_2d[:, 0::2] = img_w - coord_2d[:, 0::2]
        coord_2d[:, 0::2] -= crop[0]
        coord_2d[:, 1::2] -= crop[1]
        coord_2d[:, 0::2] = coord_2d[:, 0::2] / scale[0]
        coord_2d[:, 1::2] = coord_2d[:, 1::2] / scale[1]
    return coord_2d


# ----------------------------------------------------------------------- batched form ---
def fold_projection(img_metas):
    """Host side, float64: for every scene fold [undo 3D aug] -> [depth2img] into M (3,4) with
    [x,y,z,1] @ M^T = (u*w, v*w, w), and [2D aug + normalisation] into (su, sv, ou, ov) with
    u_n = su*u + ou. Returns float32 tensors (B,3,4) and (B,4) on the CPU."""
    mats, affs = [], []
    for meta in img_metas:
        A = np.eye(3)
        t = np.zeros(3)
        rot = np.asarray(meta.get('pcd_rotation', np.eye(3)), dtype=np.float64)
        scale = float(meta.get('pcd_scale_factor', 1.0))
        trans = np.asarray(meta.get('pcd_trans', np.zeros(3)), dtype=np.float64)
        # x' = x @ A + t, ops of the flow undone last-to-first
        for op in list(meta.get('transformation_3d_flow', []))[::-1]:
            if op == 'T':
                t = t - trans
            elif op == 'S':
                A, t = A / scale, t / scale
            elif op == 'R':
                rinv = np.linalg.inv(rot.astype(np.float32)).astype(np.float64)
                A, t = A @ rinv, t @ rinv
            elif op == 'HF':
                if meta.get('pcd_horizontal_flip', False):
                    A = A.copy(); A[:, 0] = -A[:, 0]; t = t.copy(); t[0] = -t[0]
            elif op == 'VF':
                if meta.get('pcd_vertical_flip', False):
                    A = A.copy(); A[:, 1] = -A[:, 1]; t = t.copy(); t[1] = -t[1]
            else:
                raise AssertionError(f'This 3D data transformation op ({op}) is not supported')
        P = np.eye(4)
        d2i = np.asarray(meta['depth2img'], dtype=np.float64)
        P[:d2i.shape[0], :d2i.shape[1]] = d2i
        T = np.eye(4)
        T[:3, :3] = A.T
        T[:3, 3] = t
        mats.append((P @ T)[:3])
        img_h, img_w = meta['img_shape'][:2]
        sf = meta['scale_factor'][:2] if 'scale_factor' in meta else [1.0, 1.0]
        crop = meta.get('img_crop_offset', [0.0, 0.0])
        su, ou = float(sf[0]), float(crop[0])
        sv, ov = float(sf[1]), float(crop[1])
        if meta.get('flip', False):
            su, ou = -su, img_w - ou
        affs.append([su / (img_w - 1), sv / (img_h - 1), ou / (img_w - 1), ov / (img_h - 1)])
    return (torch.from_numpy(np.stack(mats)).float(), torch.tensor(affs, dtype=torch.float32))


def project_batched(xyz, mats, affs):
    """xyz (B,Q,3), mats (B,3,4), affs (B,4) on xyz.device -> normalised (B,Q,2) in [0,1]."""
    homo = torch.baddbmm(mats[:, :, 3].unsqueeze(1), xyz, mats[:, :, :3].transpose(1, 2))
    uv = homo[..., :2] / homo[..., 2:3]
    uv = uv * affs[:, None, 0:2] + affs[:, None, 2:4]
    return uv.clamp(0, 1)


# ------------------------------------------------------------------- boxes (targets) ---
def rotation_3d_in_axis(points, angles, axis=0):
    """Rotate (N,M,3) points by (N,) angles around `axis` (mmdet3d 0.18 convention)."""
    rot_sin, rot_cos = torch.sin(angles), torch.cos(angles)
    ones, zeros = torch.ones_like(rot_cos), torch.zeros_like(rot_cos)
    if axis == 1:
        rot_mat_T = torch.stack([torch.stack([rot_cos, zeros, -rot_sin]),
                                 torch.stack([zeros, ones, zeros]),
                                 torch.stack([rot_sin, zeros, rot_cos])])
    elif axis == 2 or axis == -1:
        rot_mat_T = torch.stack([torch.stack([rot_cos, -rot_sin, zeros]),
                                 torch.stack([rot_sin, rot_cos, zeros]),
                                 torch.stack([zeros, zeros, ones])])
    elif axis == 0:
        rot_mat_T = torch.stack([torch.stack([zeros, rot_cos, -rot_sin]),
                                 torch.stack([zeros, rot_sin, rot_cos]),
                                 torch.stack([ones, zeros, zeros])])
    else:
        raise ValueError(f'axis should in range [0, 1, 2], got {axis}')
    return torch.einsum('aij,jka->aik', (points, rot_mat_T))


class DepthBoxes:
    """The slice of mmdet3d DepthInstance3DBoxes the loss touches: tensor (G,7)
    = (x, y, z_bottom, dx, dy, dz, yaw), gravity_center, dims, yaw, points_in_boxes."""

    def __init__(self, tensor, box_dim=7, with_yaw=True, origin=(0.5, 0.5, 0)):
        tensor = torch.as_tensor(tensor, dtype=torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((0, box_dim))
        if tensor.shape[-1] == 6:
            tensor = torch.cat([tensor, tensor.new_zeros(tensor.shape[0], 1)], -1)
        self.tensor = tensor.clone()
        self.box_dim = box_dim
        self.with_yaw = with_yaw
        if tuple(origin) != (0.5, 0.5, 0):
            dst = self.tensor.new_tensor((0.5, 0.5, 0))
            src = self.tensor.new_tensor(origin)
            self.tensor[:, :3] += self.tensor[:, 3:6] * (dst - src)

    @property
    def gravity_center(self):
        g = self.tensor[:, :3].clone()
        g[:, 2] = g[:, 2] + self.tensor[:, 5] * 0.5
        return g

    @property
    def dims(self):
        return self.tensor[:, 3:6]

    @property
    def yaw(self):
        return self.tensor[:, 6]

    def to(self, device):
        out = DepthBoxes.__new__(DepthBoxes)
        out.tensor, out.box_dim, out.with_yaw = self.tensor.to(device), self.box_dim, self.with_yaw
        return out

    def new_box(self, data):
        return DepthBoxes(data, box_dim=self.box_dim, with_yaw=self.with_yaw)

    def __len__(self):
        return self.tensor.shape[0]

    def points_in_boxes(self, points):
        """points (N,3+) -> (N,G) int mask (mmdet3d 0.18 DepthInstance3DBoxes.points_in_boxes:
        rotate the points into each box frame, compare with the half sizes)."""
        return points_in_boxes_batch(points[None, :, :3], self.tensor[None])[0]


def box_corner_minmax(boxes):
    """boxes (...,7) bottom-centre (x,y,z,dx,dy,dz,yaw) -> (...,6) = (min xyz, max xyz) over the 8 corners
    of `DepthInstance3DBoxes.corners` (mmdet3d 0.18.1 core/bbox/structures/depth_box3d.py: dims * corner
    pattern relative to (0.5,0.5,0), rotated about z by rotation_3d_in_axis, + bottom centre), which is
    what multiclass_nms_single feeds to aligned_3d_nms."""
    dims, yaw = boxes[..., 3:6], boxes[..., 6]
    cosa, sina = torch.cos(yaw)[..., None], torch.sin(yaw)[..., None]
    sx = boxes.new_tensor([-0.5, -0.5, 0.5, 0.5]) * dims[..., 0:1]       # 4 footprint corners
    sy = boxes.new_tensor([-0.5, 0.5, 0.5, -0.5]) * dims[..., 1:2]
    rx = sx * cosa + sy * sina + boxes[..., 0:1]
    ry = -sx * sina + sy * cosa + boxes[..., 1:2]
    z0 = boxes[..., 2] + dims[..., 2] * 0.0
    z1 = boxes[..., 2] + dims[..., 2] * 1.0
    return torch.stack([rx.min(-1)[0], ry.min(-1)[0], torch.minimum(z0, z1),
                        rx.max(-1)[0], ry.max(-1)[0], torch.maximum(z0, z1)], -1)


def bbox3d2result(bboxes, scores, labels):
    """mmdet3d.core.bbox3d2result: detections of one scene as a dict of CPU tensors."""
    return dict(boxes_3d=bboxes.to('cpu'), scores_3d=scores.cpu(), labels_3d=labels.cpu())


def points_in_boxes_batch(xyz, boxes):
    """xyz (B,N,3), boxes (B,G,7) [bottom-centre] -> (B,N,G) int32 membership.
    Local frame test of mmdet3d roiaware_pool3d points_in_boxes: z within [bottom, bottom+dz]
    (centre +- dz/2), |x_local| < dx/2 and |y_local| < dy/2 in the box frame."""
    centre = boxes[..., :3].clone()
    centre[..., 2] = centre[..., 2] + boxes[..., 5] * 0.5
    d = xyz[:, :, None, :] - centre[:, None, :, :]                 # (B,N,G,3)
    # world -> box frame: the same rotation the head applies to its centre offsets,
    # rotation_3d_in_axis(d, -yaw, axis=2) (class_agnostic_vote_head.py:903-906)
    cosa, sina = torch.cos(boxes[..., 6]), torch.sin(boxes[..., 6])
    lx = d[..., 0] * cosa[:, None] - d[..., 1] * sina[:, None]
    ly = d[..., 0] * sina[:, None] + d[..., 1] * cosa[:, None]
    half = boxes[..., 3:6] * 0.5
    inside = (d[..., 2].abs() <= half[:, None, :, 2]) & (lx.abs() < half[:, None, :, 0]) & \
        (ly.abs() < half[:, None, :, 1])
    return inside.to(torch.int32)
