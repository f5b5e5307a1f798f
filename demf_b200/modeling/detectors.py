"""DeMFVoteNet detector (reference: demf/modeling/detectors/demfnet.py:12-283).

Point branch = PointNet2SASSG backbone + DeMFVoteHead; image branch (ResNet-50 backbone,
ChannelMapper neck, Deformable-DETR encoder: mm/image_backbone.py, modeling/encoder.py) is frozen in
the reference and produces the 4-level feature pyramid the head samples. Every image module named in
the config is built through its registry exactly like the reference does (demfnet.py:42-49) -- an
unknown `type` raises KeyError, nothing is skipped. BASELINE.json's configs feed synthetic pyramids in
place of the branch: a model built WITHOUT an `img_backbone` accepts `img` as the list of
(B,256,H_l,W_l) level features -- the encoder's output (what `extract_img_feat` returns upstream), or
the neck's output when an `img_encoder` is built.
"""
import torch
import torch.nn as nn

from ..mm.bricks import BaseModule
from ..mm.geometry import bbox3d2result
from ..mm import image_backbone  # noqa: F401  (registers ResNet / ChannelMapper)
from ..mm.registry import DETECTORS, build_backbone, build_head, build_neck


@DETECTORS.register_module()
class DeMFVoteNet(BaseModule):

    def __init__(self, pts_backbone=None, pts_bbox_head=None, pts_neck=None, img_backbone=None,
                 img_neck=None, img_encoder=None, freeze_img_branch=False, num_sampled_seed=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, **kwargs):
        super().__init__(init_cfg=init_cfg)
        if pts_backbone is not None:
            self.pts_backbone = build_backbone(pts_backbone)
        if pts_neck is not None:
            self.pts_neck = build_neck(pts_neck)
        if pts_bbox_head is not None:
            pts_bbox_head = dict(pts_bbox_head)
            pts_bbox_head.update(train_cfg=train_cfg['pts'] if train_cfg is not None else None)
            pts_bbox_head.update(test_cfg=test_cfg['pts'] if test_cfg is not None else None)
            self.pts_bbox_head = build_head(pts_bbox_head)
        # image branch: frozen feature extractor; whatever the config names is built or raises
        if img_backbone:
            self.img_backbone = build_backbone(img_backbone)
        if img_neck is not None:
            self.img_neck = build_neck(img_neck)
        if img_encoder is not None:
            self.img_encoder = build_head(img_encoder)
        self.freeze_img_branch = freeze_img_branch
        if freeze_img_branch:
            self.freeze_img_branch_params()
        self.num_sampled_seed = num_sampled_seed
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        # the head's proposal FPS (sample_mod='seed') runs on the seed coordinates = xyz of SA
        # level num_sa - num_fp: let the backbone's sampling chain issue it early
        bb, head = getattr(self, 'pts_backbone', None), getattr(self, 'pts_bbox_head', None)
        if bb is not None and head is not None and hasattr(bb, 'num_fp'):
            self._seed_fps_level = (bb.num_sa - bb.num_fp, head.num_proposal)

    # --- with_* properties of mmdet3d's ImVoteNet / Base3DDetector
    @property
    def with_img_backbone(self):
        return getattr(self, 'img_backbone', None) is not None

    @property
    def with_img_neck(self):
        return getattr(self, 'img_neck', None) is not None

    @property
    def with_img_encoder(self):
        return getattr(self, 'img_encoder', None) is not None

    @property
    def with_pts_backbone(self):
        return getattr(self, 'pts_backbone', None) is not None

    @property
    def with_pts_neck(self):
        return getattr(self, 'pts_neck', None) is not None

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        """Stage-1 checkpoints keep the image encoder under img_bbox_head.transformer.*: remap
        those keys to img_encoder.* and drop the rest of img_bbox_head (demfnet.py:85-101)."""
        for key in list(state_dict):
            if not key.startswith('img_bbox_head'):
                continue
            if 'encoder' in key or 'level_embeds' in key:
                state_dict[key.replace('img_bbox_head.transformer', 'img_encoder')] = \
                    state_dict.pop(key)
            else:
                state_dict.pop(key)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def _img_modules(self):
        return [m for m in (getattr(self, n, None) for n in
                            ('img_encoder', 'img_backbone', 'img_neck')) if m is not None]

    def freeze_img_branch_params(self):
        for m in self._img_modules():
            for p in m.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        if self.freeze_img_branch:
            for m in self._img_modules():
                m.eval()
        return self

    @torch.no_grad()
    def extract_img_feat(self, img, img_metas):
        """img: (B,3,H,W) image batch when an image branch is built, else the list of pyramid
        levels itself."""
        if isinstance(img, (list, tuple)) and not self.with_img_backbone:
            # a pyramid handed in directly is the neck's output when the encoder is built
            # (demfnet.py:118-125 runs backbone -> neck -> encoder), else the encoder's output
            return self.img_encoder(list(img), img_metas) if self.with_img_encoder else list(img)
        x = self.img_backbone(img)
        if self.with_img_neck:
            x = self.img_neck(x)
        if self.with_img_encoder:
            x = self.img_encoder(x, img_metas)
        return x

    def presample(self, points, sample_mod):
        """Every index the forward of `points` will need from furthest point sampling (backbone levels,
        the head's seed sampling) plus the first level's ball-query grid, on the current stream.
        Weight-independent: hand the result to forward_train / simple_test as `presampled=` to take
        the sampling chain off that step's critical path."""
        prefetch = sample_mod == 'seed' and hasattr(self, '_seed_fps_level')
        self.pts_backbone.prefetch_seed_fps = self._seed_fps_level if prefetch else None
        points = torch.stack(list(points)) if not torch.is_tensor(points) else points
        return self.pts_backbone.sample(points)

    def extract_pts_feat(self, pts, sample_mod=None, presampled=None):
        prefetch = sample_mod == 'seed' and hasattr(self, '_seed_fps_level')
        self.pts_backbone.prefetch_seed_fps = self._seed_fps_level if prefetch else None
        self.pts_backbone.presampled = presampled
        try:
            x = self.pts_backbone(pts)
        finally:
            self.pts_backbone.presampled = None
        self._seed_fps_indices = x.get('seed_fps_indices') if isinstance(x, dict) else None
        if self.with_pts_neck:
            x = self.pts_neck(x)
        return x['fp_xyz'][-1], x['fp_features'][-1], x['fp_indices'][-1]

    @staticmethod
    def _batch_input_shape(img, img_metas):
        if isinstance(img, (list, tuple)):
            return  # pyramid given directly: metas already carry batch_input_shape
        shape = tuple(img[0].size()[-2:])
        for meta in img_metas:
            meta['batch_input_shape'] = shape

    def _forward_head(self, points, img, img_metas, sample_mod, projection=None, presampled=None, gt=None):
        self._batch_input_shape(img, img_metas)
        img_features = self.extract_img_feat(img, img_metas)
        points = torch.stack(list(points)) if not torch.is_tensor(points) else points
        seeds_3d, seed_3d_features, seed_indices = self.extract_pts_feat(points, sample_mod, presampled)
        feat_dict = dict(seed_points=seeds_3d, seed_features=seed_3d_features,
                         seed_indices=seed_indices)
        if self._seed_fps_indices is not None:
            feat_dict['seed_sample_indices'] = self._seed_fps_indices
        img_dict = dict(img_features=img_features, img_metas=img_metas)
        if projection is not None:
            img_dict['projection'] = projection
        if gt is not None:   # training: lets the head assign targets while its decoder still runs
            img_dict['gt'] = (points,) + tuple(gt)
        return points, self.pts_bbox_head(feat_dict, sample_mod, img_dict)

    def forward_train(self, points=None, img=None, img_metas=None, gt_bboxes_ignore=None,
                      gt_bboxes_3d=None, gt_labels_3d=None, pts_semantic_mask=None,
                      pts_instance_mask=None, projection=None, presampled=None, **kwargs):
        points, bbox_preds = self._forward_head(points, img, img_metas,
                                                self.train_cfg['pts']['sample_mod'], projection, presampled,
                                                gt=(gt_bboxes_3d, gt_labels_3d))
        loss_inputs = (points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask, pts_instance_mask,
                       img_metas)
        return self.pts_bbox_head.loss(bbox_preds, *loss_inputs, gt_bboxes_ignore=gt_bboxes_ignore)

    def forward_dummy(self, points=None, img=None, img_metas=None):
        """Head outputs without post-processing: what the forward benchmark times."""
        return self._forward_head(points, img, img_metas, self.test_cfg['pts']['sample_mod'])[1]

    def simple_test(self, points=None, img_metas=None, img=None, bboxes_2d=None, rescale=False,
                    projection=None, nms=True, **kwargs):
        """Forward + decoding of the ensemble layers. nms=True (default): the reference's return value
        (demfnet.py:254-283) -- per scene a dict of CPU boxes_3d / scores_3d / labels_3d after
        DeMFVoteHead.get_bboxes. nms=False (what bench.py times, BASELINE.json's scenes/s of the
        forward path): (boxes (B, len(ensemble)*Q, 7), objectness, semantic scores) on the device."""
        _, bbox_preds = self._forward_head(points, img, img_metas,
                                           self.test_cfg['pts']['sample_mod'], projection)
        head = self.pts_bbox_head
        if not nms:
            return head.decode_ensemble(bbox_preds)
        pts = torch.stack(points) if isinstance(points, (list, tuple)) else points
        bbox_list = head.get_bboxes(pts, bbox_preds, img_metas, rescale=rescale)
        return [bbox3d2result(b, s, l) for b, s, l in bbox_list]

    # simple_test(nms=False) in two parts, for callers that overlap the image features' arrival with the first
    # (engine.GraphedForward(split=True)): `state` = what the point branch produced, then the decoder + decoding.
    def simple_test_points(self, points, img_metas, projection=None):
        points = torch.stack(list(points)) if not torch.is_tensor(points) else points
        sample_mod = self.test_cfg['pts']['sample_mod']
        seeds_3d, seed_3d_features, seed_indices = self.extract_pts_feat(points, sample_mod, None)
        feat_dict = dict(seed_points=seeds_3d, seed_features=seed_3d_features, seed_indices=seed_indices)
        if self._seed_fps_indices is not None:
            feat_dict['seed_sample_indices'] = self._seed_fps_indices
        img_dict = dict(img_metas=img_metas)
        if projection is not None:
            img_dict['projection'] = projection
        return self.pts_bbox_head.forward_points(feat_dict, sample_mod, img_dict), img_dict

    def simple_test_images(self, state, img, img_metas):
        results, img_dict = state
        self._batch_input_shape(img, img_metas)
        img_dict = dict(img_dict, img_features=self.extract_img_feat(img, img_metas))
        return self.pts_bbox_head.decode_ensemble(self.pts_bbox_head.forward_images(results, img_dict))

    def forward(self, return_loss=True, **kwargs):
        return self.forward_train(**kwargs) if return_loss else self.simple_test(**kwargs)
