"""Checkpoint files and SUN RGB-D samples in the formats the reference stack reads and writes.

* `load_checkpoint` / `save_checkpoint`: mmcv.runner's checkpoint file (reference eval.py:87,
  train.py:127-137 `checkpoint_config.meta`): a `torch.save`d dict `{'meta', 'state_dict',
  'optimizer'}`, loaded with `map_location='cpu'`, a leading `module.` (DDP wrapper) stripped from
  every key, non-strict with the mismatches reported. A stage-1 image-branch checkpoint
  (`load_from`, configs/demf/demf_votenet.py:8) goes through the same call: the detector's
  `_load_from_state_dict` renames `img_bbox_head.transformer.{encoder,level_embeds}` to
  `img_encoder.*` and drops the rest of the 2D head (demfnet.py:85-101).
* `SUNRGBDSamples`: the test-time data path of configs/demf/demf_votenet.py:224-253 over mmdet3d's
  SUN RGB-D layout (`sunrgbd_infos_*.pkl`, `points/*.bin`, `sunrgbd_trainval/image/*.jpg`):
  LoadPointsFromFile(load_dim=6, use_dim=[0,1,2], shift_height=True) -> LoadImageFromFile ->
  Resize((1333, 800), keep_ratio) -> Normalize(img_norm_cfg, to_rgb) -> Pad(size_divisor=32) ->
  PointSample(20000), and `collate` = the padded image batch + img_metas (`depth2img`, `img_shape`,
  `scale_factor`, `batch_input_shape`, identity 3D flow) that DeMFVoteNet.simple_test consumes.
  Training-time augmentation (RandomFlip3D / GlobalRotScaleTrans) is `augment=True`: the same
  img_meta keys (`pcd_horizontal_flip`, `pcd_rotation`, `pcd_scale_factor`, `transformation_3d_flow`)
  the head inverts when it projects proposals into the image (class_agnostic_vote_head.py:524-547).
"""
import os
import pickle
import re
from collections import OrderedDict

import numpy as np
import torch

IMG_NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
CLASSES = ('bed', 'table', 'sofa', 'chair', 'toilet', 'desk', 'dresser', 'night_stand', 'bookshelf',
           'bathtub')


# ------------------------------------------------------------------ checkpoints ---
def _model_of(model):
    return model.module if hasattr(model, 'module') and isinstance(model.module, torch.nn.Module) else model


def load_checkpoint(model, filename, map_location='cpu', strict=False, revise_keys=((r'^module\.', ''),),
                    logger=None):
    """mmcv.runner.load_checkpoint: returns the checkpoint dict (so callers can read `meta`)."""
    checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if not isinstance(checkpoint, dict):
        raise RuntimeError(f'No state_dict found in checkpoint file {filename}')
    state_dict = checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint
    metadata = getattr(state_dict, '_metadata', OrderedDict())
    for pattern, repl in revise_keys:
        state_dict = OrderedDict((re.sub(pattern, repl, k), v) for k, v in state_dict.items())
    state_dict._metadata = metadata
    target = _model_of(model)
    result = target.load_state_dict(state_dict, strict=strict)
    missing = [k for k in result.missing_keys if 'num_batches_tracked' not in k]
    report = []
    if result.unexpected_keys:
        report.append('unexpected key in source state_dict: ' + ', '.join(result.unexpected_keys))
    if missing:
        report.append('missing keys in source state_dict: ' + ', '.join(missing))
    if report:
        msg = 'The model and loaded state dict do not match exactly\n' + '\n'.join(report)
        (logger.warning if logger is not None else print)(msg)
    checkpoint['_load_result'] = dict(missing_keys=missing, unexpected_keys=list(result.unexpected_keys))
    return checkpoint


def save_checkpoint(model, filename, optimizer=None, meta=None):
    """mmcv.runner.save_checkpoint: CPU tensors, `meta` (versions, config text, CLASSES, epoch, iter)."""
    target = _model_of(model)
    meta = dict(meta or {})
    if hasattr(target, 'CLASSES') and target.CLASSES is not None:
        meta.setdefault('CLASSES', target.CLASSES)
    checkpoint = {'meta': meta,
                  'state_dict': OrderedDict((k, v.detach().cpu()) for k, v in target.state_dict().items())}
    if optimizer is not None:
        checkpoint['optimizer'] = optimizer.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    tmp = filename + '.tmp'
    torch.save(checkpoint, tmp)
    os.replace(tmp, filename)


# ----------------------------------------------------------------------- samples ---
def shift_height(points_xyz):
    """LoadPointsFromFile(shift_height=True): height above the 0.99th-percentile floor as 4th channel."""
    floor = np.percentile(points_xyz[:, 2], 0.99)
    return np.concatenate([points_xyz[:, :3], (points_xyz[:, 2] - floor)[:, None]], 1).astype(np.float32)


def point_sample(points, num_points, rng):
    """mmdet3d PointSample: without replacement when the cloud is large enough, else with."""
    n = points.shape[0]
    choice = rng.choice(n, num_points, replace=n < num_points)
    return points[choice]


def resize_keep_ratio(img, img_scale=(1333, 800)):
    """mmcv.imrescale with keep_ratio: the largest scale that fits (long, short) edges; bilinear.
    -> resized image, (w_scale, h_scale, w_scale, h_scale)."""
    import cv2
    h, w = img.shape[:2]
    long_edge, short_edge = max(img_scale), min(img_scale)
    scale = min(long_edge / max(h, w), short_edge / min(h, w))
    new_w, new_h = int(w * float(scale) + 0.5), int(h * float(scale) + 0.5)
    out = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    return out, np.array([new_w / w, new_h / h, new_w / w, new_h / h], dtype=np.float32)


def normalize_image(img_bgr, mean=IMG_NORM['mean'], std=IMG_NORM['std'], to_rgb=IMG_NORM['to_rgb']):
    img = img_bgr.astype(np.float32)
    if to_rgb:
        img = img[..., ::-1]
    return (img - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)


def sunrgbd_depth2img(calib):
    """mmdet3d SUNRGBDDataset.get_data_info: depth coordinates -> image plane, (3,3)."""
    rt = np.asarray(calib['Rt'], dtype=np.float32).reshape(3, 3)
    k = np.asarray(calib['K'], dtype=np.float32).reshape(3, 3)
    rt_mat = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32) @ rt.transpose(1, 0)
    return k @ rt_mat


class SUNRGBDSamples:
    """Indexable view of an mmdet3d SUN RGB-D info file: sample(i) -> dict(points (N,4) f32, img (3,H,W)
    f32 normalised, img_meta, gt_bboxes_3d (G,7), gt_labels_3d (G,))."""

    def __init__(self, data_root, ann_file, num_points=20000, img_scale=(1333, 800), augment=False, seed=0):
        self.data_root = data_root
        with open(ann_file if os.path.isabs(ann_file) else os.path.join(data_root, ann_file), 'rb') as f:
            self.infos = pickle.load(f)
        self.num_points, self.img_scale, self.augment = num_points, img_scale, augment
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return len(self.infos)

    def annotations(self, info):
        annos = info.get('annos', {})
        if annos.get('gt_num', 0) == 0:
            return np.zeros((0, 7), np.float32), np.zeros((0,), np.int64)
        boxes = np.asarray(annos['gt_boxes_upright_depth'], dtype=np.float32).reshape(-1, 7)
        labels = np.asarray(annos['class'], dtype=np.int64)
        return boxes, labels

    def sample(self, i):
        import cv2
        info = self.infos[i]
        raw = np.fromfile(os.path.join(self.data_root, info['pts_path']), dtype=np.float32).reshape(-1, 6)
        points = shift_height(raw[:, :3])
        img = cv2.imread(os.path.join(self.data_root, 'sunrgbd_trainval', info['image']['image_path']))
        if img is None:
            raise FileNotFoundError(info['image']['image_path'])
        img, scale_factor = resize_keep_ratio(img, self.img_scale)
        img_shape = img.shape
        img = normalize_image(img)
        boxes, labels = self.annotations(info)
        meta = dict(depth2img=sunrgbd_depth2img(info['calib']), img_shape=img_shape,
                    ori_shape=tuple(info['image'].get('image_shape', img_shape[:2])),
                    scale_factor=scale_factor, flip=False, sample_idx=info['image'].get('image_idx', i),
                    pcd_horizontal_flip=False, pcd_vertical_flip=False, pcd_scale_factor=1.0,
                    pcd_rotation=torch.eye(3), pcd_trans=np.zeros(3, np.float32),
                    transformation_3d_flow=['R', 'S', 'T'])
        if self.augment:
            points, boxes, meta = self._augment(points, boxes, meta)
        points = point_sample(points, self.num_points, self.rng)
        return dict(points=torch.from_numpy(points), img=torch.from_numpy(img.transpose(2, 0, 1).copy()),
                    img_meta=meta, gt_bboxes_3d=torch.from_numpy(boxes), gt_labels_3d=torch.from_numpy(labels))

    def _augment(self, points, boxes, meta):
        """RandomFlip3D(flip_ratio_bev_horizontal=0.5) then GlobalRotScaleTrans(rot +-0.5236, scale
        0.85-1.15, shift_height) in depth coordinates (configs/demf/demf_votenet.py:198-207)."""
        flow = []
        if self.rng.random() < 0.5:     # depth-frame horizontal flip: x -> -x, yaw -> pi - yaw
            points[:, 0] = -points[:, 0]
            boxes = boxes.copy()
            boxes[:, 0] = -boxes[:, 0]
            boxes[:, 6] = np.pi - boxes[:, 6]
            meta['pcd_horizontal_flip'] = True
        flow.append('HF')
        angle = self.rng.uniform(-0.523599, 0.523599)
        c, s = np.cos(angle), np.sin(angle)
        rot_t = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], dtype=np.float32)   # points @ rot_t: counter-clockwise, yaw += angle
        points[:, :3] = points[:, :3] @ rot_t
        boxes = boxes.copy()
        boxes[:, :3] = boxes[:, :3] @ rot_t
        boxes[:, 6] += angle
        meta['pcd_rotation'] = torch.from_numpy(rot_t)
        flow.append('R')
        scale = self.rng.uniform(0.85, 1.15)
        points[:, :3] *= scale
        points[:, 3] *= scale
        boxes[:, :6] *= scale
        meta['pcd_scale_factor'] = float(scale)
        flow += ['S', 'T']
        meta['transformation_3d_flow'] = flow
        return points, boxes, meta


def collate(samples, size_divisor=32):
    """Samples -> forward_train / simple_test keywords: points (B,N,4), img (B,3,H,W) zero-padded to the
    batch maximum rounded up to `size_divisor` (mmcv Pad + DataContainer stacking), img_metas with
    `batch_input_shape`, per-scene ground truth lists."""
    from .mm.geometry import DepthBoxes
    hs = [s['img'].shape[1] for s in samples]
    ws = [s['img'].shape[2] for s in samples]
    H = -(-max(hs) // size_divisor) * size_divisor
    W = -(-max(ws) // size_divisor) * size_divisor
    img = torch.zeros(len(samples), 3, H, W)
    metas = []
    for b, s in enumerate(samples):
        img[b, :, :hs[b], :ws[b]] = s['img']
        meta = dict(s['img_meta'])
        meta['pad_shape'] = (H, W, 3)
        meta['batch_input_shape'] = (H, W)
        metas.append(meta)
    return dict(points=torch.stack([s['points'] for s in samples]), img=img, img_metas=metas,
                gt_bboxes_3d=[DepthBoxes(s['gt_bboxes_3d']) for s in samples],
                gt_labels_3d=[s['gt_labels_3d'] for s in samples])
