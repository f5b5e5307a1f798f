"""Bounding-box coder of the DeMF head (reference: demf/core/bbox/coders/
class_agnostic_bbox_coder.py:140-251, on top of mmdet3d PartialBinBasedBBoxCoder's
angle2class / class2angle)."""
import numpy as np
import torch

from ..mm.registry import BBOX_CODERS


@BBOX_CODERS.register_module()
class DeMFClassAgnosticBBoxCoder:
    """Class-agnostic coder: the head regresses centre offset (3), size (3) and a
    num_dir_bins-way heading class + per-bin residual; classification = objectness(2)+semantic."""

    def __init__(self, num_dir_bins, num_sizes, mean_sizes, with_rot=True):
        self.num_dir_bins = num_dir_bins
        self.num_sizes = num_sizes
        self.mean_sizes = mean_sizes
        self.with_rot = with_rot
        assert len(mean_sizes) == num_sizes

    # --- heading discretisation (mmdet3d PartialBinBasedBBoxCoder) ---
    def angle2class(self, angle):
        angle = angle % (2 * np.pi)
        angle_per_class = 2 * np.pi / float(self.num_dir_bins)
        shifted_angle = (angle + angle_per_class / 2) % (2 * np.pi)
        angle_cls = shifted_angle // angle_per_class
        angle_res = shifted_angle - (angle_cls * angle_per_class + angle_per_class / 2)
        return angle_cls.long(), angle_res

    def class2angle(self, angle_cls, angle_res, limit_period=True):
        angle_per_class = 2 * np.pi / float(self.num_dir_bins)
        angle = angle_cls.float() * angle_per_class + angle_res
        if limit_period:
            angle = torch.where(angle > np.pi, angle - 2 * np.pi, angle)
        return angle

    # --- targets (class_agnostic_bbox_coder.py:142-166) ---
    def encode(self, gt_bboxes_3d, gt_labels_3d, ret_dir_target=False):
        center_target = gt_bboxes_3d.gravity_center
        size_res_target = gt_bboxes_3d.dims
        box_num = gt_labels_3d.shape[0]
        if self.with_rot:
            dir_class_target, dir_res_target = self.angle2class(gt_bboxes_3d.yaw)
            dir_target = gt_bboxes_3d.yaw
        else:
            dir_class_target = gt_labels_3d.new_zeros(box_num)
            dir_res_target = gt_bboxes_3d.tensor.new_zeros(box_num)
            dir_target = gt_bboxes_3d.tensor.new_zeros(box_num)
        if ret_dir_target:
            return center_target, size_res_target, dir_class_target, dir_res_target, dir_target
        return center_target, size_res_target, dir_class_target, dir_res_target

    # --- predictions -> boxes (class_agnostic_bbox_coder.py:168-194) ---
    def decode(self, bbox_out, mode='rpn'):
        assert mode in ['rpn', 'rcnn']
        prefix = 'refined_' if mode == 'rcnn' else ''
        center = bbox_out['center']
        bbox_size = bbox_out['size']
        batch_size, num_proposal, _ = center.shape
        if self.with_rot:
            if mode == 'rpn':
                dir_class = torch.argmax(bbox_out['dir_class'], -1).detach()
                dir_res = torch.gather(bbox_out['dir_res'], -1, dir_class.unsqueeze(-1)).squeeze(-1)
                dir_angle = self.class2angle(dir_class, dir_res).reshape(batch_size, num_proposal, 1)
            else:
                dir_angle = bbox_out[prefix + 'angle'].reshape(batch_size, num_proposal, 1)
            dir_angle = dir_angle % (2 * np.pi)
        else:
            dir_angle = center.new_zeros(batch_size, num_proposal, 1)
        return torch.cat([center, bbox_size, dir_angle], dim=-1)

    # --- raw conv outputs -> named slices (class_agnostic_bbox_coder.py:196-240) ---
    def split_pred(self, cls_preds, reg_preds, base_xyz):
        """cls_preds (B,2+num_classes,Q), reg_preds (B,6+2*bins,Q), base_xyz (B,Q,3)."""
        cls_t = cls_preds.transpose(2, 1)
        reg_t = reg_preds.transpose(2, 1)
        nb = self.num_dir_bins
        # the slices stay VIEWS of the two row tensors (upstream copies each with .contiguous():
        # seven tiny launches per stage); every consumer here takes strided inputs
        results = dict(
            center=base_xyz + reg_t[..., 0:3],
            size=reg_t[..., 3:6],
            dir_class=reg_t[..., 6:6 + nb])
        dir_res_norm = reg_t[..., 6 + nb:6 + 2 * nb]
        results['dir_res_norm'] = dir_res_norm
        results['dir_res'] = dir_res_norm * (np.pi / nb)
        results['obj_scores'] = cls_t[..., 0:2]
        if cls_t.shape[-1] > 2:
            results['sem_scores'] = cls_t[..., 2:]
        return results

    def decode_corners(self, center, size):
        """(B,N,3),(B,N,3) -> axis-aligned corners (B,N,6)."""
        size_half = size / 2.0
        return torch.cat([center - size_half, center + size_half], dim=-1)
