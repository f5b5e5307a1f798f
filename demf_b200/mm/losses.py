"""Loss modules with the mmdet / mmdet3d registry names used by configs/demf/demf_votenet.py:113-154:
CrossEntropyLoss, SmoothL1Loss (mmdet), AxisAlignedIoULoss, ChamferDistance (mmdet3d).
Only the options those configs use are implemented (softmax CE, sum/mean/none reductions)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .registry import LOSSES


def reduce_loss(loss, reduction):
    if reduction == 'none':
        return loss
    if reduction == 'mean':
        return loss.mean()
    if reduction == 'sum':
        return loss.sum()
    raise ValueError(reduction)


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return reduce_loss(loss, reduction)
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None,
                 loss_weight=1.0):
        super().__init__()
        assert not use_sigmoid and not use_mask, "only softmax cross entropy is on the DeMF path"
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = class_weight
        self._class_weight_cache = {}

    def _class_weight(self, like):
        if self.class_weight is None:
            return None
        key = (str(like.device), like.dtype)
        if key not in self._class_weight_cache:  # once per device: no H2D copy in the step
            self._class_weight_cache[key] = like.new_tensor(self.class_weight)
        return self._class_weight_cache[key]

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        class_weight = self._class_weight(cls_score)
        loss = F.cross_entropy(cls_score, label, weight=class_weight, reduction='none')
        if weight is not None:
            weight = weight.float()
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


@LOSSES.register_module()
class SmoothL1Loss(nn.Module):
    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.beta = beta
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None,
                **kwargs):
        reduction = reduction_override if reduction_override else self.reduction
        assert self.beta > 0 and pred.size() == target.size()
        diff = torch.abs(pred - target)
        loss = torch.where(diff < self.beta, 0.5 * diff * diff / self.beta,
                           diff - 0.5 * self.beta)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


def axis_aligned_iou_aligned(b1, b2, eps=1e-6):
    """IoU of corresponding axis-aligned boxes (…,6) = (x1,y1,z1,x2,y2,z2)."""
    area1 = (b1[..., 3] - b1[..., 0]) * (b1[..., 4] - b1[..., 1]) * (b1[..., 5] - b1[..., 2])
    area2 = (b2[..., 3] - b2[..., 0]) * (b2[..., 4] - b2[..., 1]) * (b2[..., 5] - b2[..., 2])
    lt = torch.max(b1[..., :3], b2[..., :3])
    rb = torch.min(b1[..., 3:], b2[..., 3:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1] * wh[..., 2]
    union = (area1 + area2 - overlap).clamp(min=eps)
    return overlap / union


@LOSSES.register_module()
class AxisAlignedIoULoss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None,
                **kwargs):
        reduction = reduction_override if reduction_override else self.reduction
        loss = 1 - axis_aligned_iou_aligned(pred, target)
        # (upstream returns (pred*weight).sum() when no weight is positive, which needs a device
        # sync to decide; with all-zero weights the weighted sum below is 0 with zero gradient
        # for every finite box, which is what that branch guards.)
        return weight_reduce_loss(loss, weight, reduction, avg_factor) * self.loss_weight


def chamfer_distance(src, dst, src_weight=1.0, dst_weight=1.0, criterion_mode='l2',
                     reduction='mean'):
    """src (B,N,C), dst (B,M,C) -> (loss_src, loss_dst, idx_src->dst (B,N), idx_dst->src (B,M))."""
    if criterion_mode == 'smooth_l1':
        criterion = F.smooth_l1_loss
    elif criterion_mode == 'l1':
        criterion = F.l1_loss
    elif criterion_mode == 'l2':
        criterion = F.mse_loss
    else:
        raise NotImplementedError
    src_expand = src.unsqueeze(2).expand(-1, -1, dst.shape[1], -1)
    dst_expand = dst.unsqueeze(1).expand(-1, src.shape[1], -1, -1)
    distance = criterion(src_expand, dst_expand, reduction='none').sum(-1)
    src2dst_distance, indices1 = torch.min(distance, dim=2)
    dst2src_distance, indices2 = torch.min(distance, dim=1)
    loss_src = src2dst_distance * src_weight
    loss_dst = dst2src_distance * dst_weight
    if reduction == 'sum':
        loss_src, loss_dst = torch.sum(loss_src), torch.sum(loss_dst)
    elif reduction == 'mean':
        loss_src, loss_dst = torch.mean(loss_src), torch.mean(loss_dst)
    elif reduction != 'none':
        raise NotImplementedError
    return loss_src, loss_dst, indices1, indices2


@LOSSES.register_module()
class ChamferDistance(nn.Module):
    def __init__(self, mode='l2', reduction='mean', loss_src_weight=1.0, loss_dst_weight=1.0):
        super().__init__()
        assert mode in ['smooth_l1', 'l1', 'l2']
        assert reduction in ['none', 'sum', 'mean']
        self.mode = mode
        self.reduction = reduction
        self.loss_src_weight = loss_src_weight
        self.loss_dst_weight = loss_dst_weight

    def forward(self, source, target, src_weight=1.0, dst_weight=1.0, reduction_override=None,
                return_indices=False, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        loss_source, loss_target, indices1, indices2 = chamfer_distance(
            source, target, src_weight, dst_weight, self.mode, reduction)
        loss_source = loss_source * self.loss_src_weight
        loss_target = loss_target * self.loss_dst_weight
        if return_indices:
            return loss_source, loss_target, indices1, indices2
        return loss_source, loss_target
