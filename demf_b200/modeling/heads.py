"""DeMFVoteHead: VoteNet head whose proposals are refined by deformable attention over image
features (reference: demf/modeling/heads/class_agnostic_vote_head.py:335-941).

Same constructor arguments, forward signature, result-dict keys, parameter names and loss keys
as the reference. What changed is how the work reaches the GPU:
  * vote aggregation and the prediction convs run on point-major rows (mm/pointnet_modules.py);
  * reference points for the whole batch come from one bmm (mm/geometry.py) instead of a Python
    loop over scenes (reference :524-547);
  * the image pyramid is laid out once as (B,S,C) -- the layout the MSDA kernel and the
    value_proj GEMM both want -- and handed on as an (S,B,C) view (reference :570-591 builds a
    (S,B,C) tensor that the attention module permutes back);
  * targets are assigned for the whole batch with padded (B,G,*) tensors instead of per-scene,
    per-box Python loops (reference :818-941).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..mm import geometry
from ..mm import point_ops as P
from ..mm.bricks import BaseModule
from ..mm.pointnet_modules import BaseConvBboxHead, VoteModule
from ..mm.registry import (HEADS, build_bbox_coder, build_loss, build_sa_module,
                           build_transformer_layer)


class LossDict(dict):
    """A loss dict that may carry `.total` (the sum of its entries as ONE tensor, built with fewer kernels than
    adding the entries) and, for one fused stage, `.vec` / `.index` (the kernel's (7,) vector and the entry -> slot
    map). Plain-dict consumers are unaffected."""
    total = None
    vec = None
    index = None


@HEADS.register_module()
class DeMFVoteHead(BaseModule):

    def __init__(self, num_classes, bbox_coder, train_cfg=None, test_cfg=None,
                 vote_module_cfg=None, vote_aggregation_cfg=None, pred_layer_cfg=None,
                 conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'), objectness_loss=None,
                 center_loss=None, dir_class_loss=None, dir_res_loss=None, size_class_loss=None,
                 size_res_loss=None, semantic_loss=None, iou_loss=None, decoder=None,
                 init_cfg=None):
        super().__init__(init_cfg=init_cfg)
        self.num_classes = num_classes
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.gt_per_seed = vote_module_cfg['gt_per_seed']
        self.num_proposal = vote_aggregation_cfg['num_point']

        self.objectness_loss = build_loss(objectness_loss)
        self.center_loss = build_loss(center_loss)
        self.dir_res_loss = build_loss(dir_res_loss)
        self.dir_class_loss = build_loss(dir_class_loss)
        self.size_res_loss = build_loss(size_res_loss)
        if size_class_loss is not None:
            self.size_class_loss = build_loss(size_class_loss)
        if semantic_loss is not None:
            self.semantic_loss = build_loss(semantic_loss)
        self.iou_loss = build_loss(iou_loss) if iou_loss is not None else None

        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.num_sizes = self.bbox_coder.num_sizes
        self.num_dir_bins = self.bbox_coder.num_dir_bins

        self.vote_module = VoteModule(**vote_module_cfg)
        self.vote_aggregation = build_sa_module(vote_aggregation_cfg)
        self.fp16_enabled = False

        self.num_decoder_layers = decoder['num_layers']
        self.num_fusion_layers = decoder['num_layers']
        self.decoder = nn.ModuleList()
        for _ in range(self.num_decoder_layers):
            self.decoder.append(build_transformer_layer(decoder))

        pred_layer_cfg = dict(pred_layer_cfg)
        self.conv_pred_layers = pred_layer_cfg.pop('conv_pred_layers')
        assert self.conv_pred_layers == self.num_decoder_layers + 1
        self.conv_preds = []
        for i in range(self.conv_pred_layers):
            conv_pred = BaseConvBboxHead(**pred_layer_cfg,
                                         num_cls_out_channels=self._get_cls_out_channels(),
                                         num_reg_out_channels=self._get_reg_out_channels())
            self.add_module('conv_pred' + str(i), conv_pred)
            self.conv_preds.append(conv_pred)

    def _get_cls_out_channels(self):
        return self.num_classes + 2 if hasattr(self, 'semantic_loss') else 2

    def _get_reg_out_channels(self):
        return 6 + self.num_dir_bins * 2

    def init_weights(self):
        for layer in self.decoder:
            layer.init_weights()
        self._is_init = True

    @staticmethod
    def _extract_input(feat_dict):
        if 'seed_points' in feat_dict and 'seed_features' in feat_dict \
                and 'seed_indices' in feat_dict:
            return feat_dict['seed_points'], feat_dict['seed_features'], feat_dict['seed_indices']
        return feat_dict['fp_xyz'][-1], feat_dict['fp_features'][-1], feat_dict['fp_indices'][-1]

    # ------------------------------------------------------------------ forward ---
    def forward(self, feat_dict, sample_mod, img_dict):
        """= forward_images(forward_points(...)): the part that only needs the point cloud (vote module, proposal
        sampling, vote aggregation; with ground truth also the early target assignment), then the decoder, which is
        the first reader of the image features. engine.GraphedForward captures the two parts as two graphs so that
        the image features can still be on their way to the device while the first runs."""
        return self.forward_images(self.forward_points(feat_dict, sample_mod, img_dict), img_dict)

    def forward_points(self, feat_dict, sample_mod, img_dict):
        assert sample_mod in ['vote', 'seed', 'random', 'spec']
        seed_points, seed_features, seed_indices = self._extract_input(feat_dict)

        vote_points, vote_features, vote_offset = self.vote_module(seed_points, seed_features)
        results = dict(seed_points=seed_points, seed_indices=seed_indices, vote_points=vote_points,
                       vote_features=vote_features, vote_offset=vote_offset)

        if sample_mod == 'vote':
            aggregation_inputs = dict(points_xyz=vote_points, features=vote_features)
        elif sample_mod == 'seed':
            # FPS on the seeds (reference :429-430); the backbone's sampling chain may already
            # have run it on its side stream (same input, same kernel)
            sample_indices = feat_dict.get('seed_sample_indices')
            if sample_indices is None:
                sample_indices = P.furthest_point_sample(seed_points.contiguous(),
                                                         self.num_proposal)
            aggregation_inputs = dict(points_xyz=vote_points, features=vote_features,
                                      indices=sample_indices)
        elif sample_mod == 'random':
            batch_size, num_seed = seed_points.shape[:2]
            sample_indices = torch.randint(0, num_seed, (batch_size, self.num_proposal),
                                           dtype=torch.int32, device=seed_points.device)
            aggregation_inputs = dict(points_xyz=vote_points, features=vote_features,
                                      indices=sample_indices)
        else:  # 'spec'
            aggregation_inputs = dict(points_xyz=seed_points, features=seed_features,
                                      target_xyz=vote_points)

        aggregated_points, features, aggregated_indices = self.vote_aggregation(
            **aggregation_inputs)
        results['aggregated_points'] = aggregated_points
        results['aggregated_indices'] = aggregated_indices
        if self.early_targets and img_dict.get('gt') is not None and aggregated_points.is_cuda \
                and torch.is_tensor(img_dict['gt'][1]):
            # Target assignment needs the proposals' positions but nothing the decoder computes: start it
            # now on a second stream, under the decoder and prediction heads (~120 small kernels of latency).
            dev = aggregated_points.device
            cur = torch.cuda.current_stream(dev)
            side = self._stage_stream(dev, 'targets')
            side.wait_stream(cur)
            common = {k: results[k].detach() for k in ('seed_points', 'seed_indices', 'aggregated_points',
                                                       'vote_points')}
            gt_points, gt_boxes, gt_labels = img_dict['gt']
            for t in list(common.values()) + [gt_points, gt_boxes, gt_labels]:
                t.record_stream(side)
            with torch.cuda.stream(side), torch.no_grad():
                targets = self.get_targets(gt_points, gt_boxes, gt_labels, bbox_preds=common)
                done = torch.cuda.Event()
                done.record(side)
            results['_targets'] = (targets, done)
        results['_aggregated_features'] = features
        return results

    def forward_images(self, results, img_dict):
        features = results.pop('_aggregated_features')
        results['decode_res_all'] = self.transformer_decoder(
            features, results['aggregated_points'], img_dict['img_features'], img_dict['img_metas'],
            projection=img_dict.get('projection'))
        return results

    def transformer_decoder(self, features, aggregated_points, img_features, img_metas,
                            projection=None):
        """features (B,C,Q) -> list of num_layers+1 prediction dicts (reference :468-512)."""
        decode_res_all = []
        cls_predictions, reg_predictions = self.conv_preds[0](features)
        decode_res = self.bbox_coder.split_pred(cls_predictions, reg_predictions, aggregated_points)
        decode_res_all.append(decode_res)

        feat_flatten, mask_flatten, reference_points, spatial_shapes, level_start_index, \
            valid_ratios = self.prepare_decoder_inputs(aggregated_points, img_features, img_metas,
                                                       projection=projection)

        query = features.permute(2, 0, 1)
        for i in range(self.num_decoder_layers):
            query_pos = torch.cat([decode_res['center'], decode_res['size']], dim=-1).detach()
            query = self.decoder[i](
                query=query, key=None, value=feat_flatten, query_pos=query_pos,
                key_padding_mask=mask_flatten, reference_points=reference_points,
                spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                valid_ratios=valid_ratios)
            cls_predictions, reg_predictions = self.conv_preds[i + 1](query.permute(1, 2, 0))
            decode_res = self.bbox_coder.split_pred(cls_predictions, reg_predictions,
                                                    aggregated_points)
            decode_res_all.append(decode_res)
        return decode_res_all

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def get_reference_points(self, seeds_3d_batch, img_metas, projection=None):
        """(B,Q,3) proposals -> (B,Q,2) normalised image coordinates, clamped to [0,1].
        Whole batch in one bmm; the per-scene matrices are folded on the host (geometry.py).
        `projection` = (mats (B,3,4), affs (B,4)) already on the device (CUDA-graph replay keeps
        them in static buffers, engine.GraphedForward)."""
        if projection is None:
            mats, affs = geometry.fold_projection(img_metas)
            dev = seeds_3d_batch.device
            projection = (mats.to(dev, non_blocking=True), affs.to(dev, non_blocking=True))
        if seeds_3d_batch.is_cuda and not torch.is_grad_enabled() and seeds_3d_batch.dtype == torch.float32:
            return P.project_points(seeds_3d_batch, *projection)
        return geometry.project_batched(seeds_3d_batch, *projection)

    _level_cache = {}

    @classmethod
    def _level_tensors(cls, shapes, starts, dev):
        key = (tuple(shapes), str(dev))
        if key not in cls._level_cache:
            cls._level_cache[key] = (
                torch.as_tensor(shapes, dtype=torch.long).to(dev),
                torch.as_tensor(starts[:-1], dtype=torch.long).to(dev))
        return cls._level_cache[key]

    def prepare_decoder_inputs(self, seeds_3d, mlvl_feats, img_metas, projection=None):
        """-> feat_flatten (S,B,C) [view of a (B,S,C) buffer], mask_flatten (B,S) bool or None,
        reference_points (B,Q,2), spatial_shapes (L,2) i64, level_start_index (L) i64,
        valid_ratios (B,L,2). Reference :549-594."""
        reference_points = self.get_reference_points(seeds_3d, img_metas, projection)
        dev = mlvl_feats[0].device
        batch_size, channels = mlvl_feats[0].shape[:2]
        shapes = [tuple(f.shape[-2:]) for f in mlvl_feats]
        starts = np.concatenate([[0], np.cumsum([h * w for h, w in shapes])])
        num_value = int(starts[-1])

        if dev.type == 'cuda' and len(mlvl_feats) <= 8 and all(
                f.is_contiguous() and f.dtype == torch.float32 and not f.requires_grad for f in mlvl_feats):
            pyramid = P.levels_to_rows(list(mlvl_feats))     # one coalesced transpose launch
        else:
            pyramid = mlvl_feats[0].new_empty(batch_size, num_value, channels)
            for lvl, feat in enumerate(mlvl_feats):
                pyramid[:, starts[lvl]:starts[lvl + 1]].copy_(feat.flatten(2).transpose(1, 2))
        feat_flatten = pyramid.permute(1, 0, 2)

        input_img_h, input_img_w = img_metas[0]['batch_input_shape']
        padded = any(tuple(m['img_shape'][:2]) != (input_img_h, input_img_w) for m in img_metas)
        if padded:
            img_masks = mlvl_feats[0].new_ones((batch_size, input_img_h, input_img_w))
            for img_id in range(batch_size):
                img_h, img_w = img_metas[img_id]['img_shape'][:2]
                img_masks[img_id, :img_h, :img_w] = 0
            mlvl_masks = [F.interpolate(img_masks[None], size=s).to(torch.bool).squeeze(0)
                          for s in shapes]
            mask_flatten = torch.cat([m.flatten(1) for m in mlvl_masks], 1)
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in mlvl_masks], 1)
        else:  # nothing is padding: no mask to apply, every valid ratio is exactly 1
            mask_flatten = None
            valid_ratios = reference_points.new_ones(batch_size, len(shapes), 2)

        spatial_shapes, level_start_index = self._level_tensors(shapes, starts, dev)
        return (feat_flatten, mask_flatten, reference_points, spatial_shapes, level_start_index,
                valid_ratios)

    # ------------------------------------------------------------ post-processing ---
    def decode_ensemble(self, bbox_preds):
        """Decoded boxes (B,R,7) [gravity centre], objectness (B,R) and semantic probabilities (B,R,C)
        of the `test_cfg.ensemble_layers` stages, R = len(layers) * num_proposal (reference :720-736)."""
        layers = list(self.test_cfg['ensemble_layers'])
        stages = bbox_preds['decode_res_all']
        first = stages[layers[0]]
        coder = self.bbox_coder
        if first['center'].is_cuda and coder.with_rot and 'sem_scores' in first:
            # one launch per ensembled stage: softmaxes, heading decode and the concatenations
            B, Q = first['center'].shape[:2]
            R = Q * len(layers)
            dev = first['center'].device
            box = torch.empty(B, R, 7, device=dev)
            obj = torch.empty(B, R, device=dev)
            sem = torch.empty(B, R, first['sem_scores'].size(-1), device=dev)
            for n, i in enumerate(layers):
                P.decode_boxes(stages[i], coder.num_dir_bins, box, obj, sem, n * Q)
            return box, obj, sem
        obj, sem, box = [], [], []
        for i in layers:
            res = stages[i]
            obj.append(torch.softmax(res['obj_scores'], dim=-1)[..., -1])
            sem.append(torch.softmax(res['sem_scores'], dim=-1))
            box.append(coder.decode(res))
        return torch.cat(box, 1), torch.cat(obj, 1), torch.cat(sem, 1)

    def get_bboxes(self, points, bbox_preds, input_metas, rescale=False, use_nms=True):
        """Reference :714-754. points (B,N,3+); returns per scene (boxes, scores, labels) after
        class-aware 3D NMS, or the raw (B,R,7) boxes when use_nms is False."""
        bbox3d, obj_scores, sem_scores = self.decode_ensemble(bbox_preds)
        if not use_nms:
            return bbox3d
        selected = self.multiclass_nms_batch(obj_scores, sem_scores, bbox3d, points[..., :3])
        results = []
        for b, (boxes, scores, labels) in enumerate(selected):
            box_type = input_metas[b].get('box_type_3d', geometry.DepthBoxes) if input_metas else geometry.DepthBoxes
            results.append((box_type(boxes, box_dim=boxes.shape[-1], with_yaw=self.bbox_coder.with_rot),
                            scores, labels))
        return results

    def multiclass_nms_single(self, obj_scores, sem_scores, bbox, points, input_meta=None):
        """One scene, upstream signature (mmdet3d 0.18.1 VoteHead.multiclass_nms_single):
        obj_scores (R,), sem_scores (R,C), bbox (R,7) gravity centre, points (N,3)."""
        return self.multiclass_nms_batch(obj_scores[None], sem_scores[None], bbox[None], points[None])[0]

    def multiclass_nms_batch(self, obj_scores, sem_scores, bbox, points):
        """multiclass_nms_single for the whole batch with ONE host synchronisation (upstream: a Python
        loop over scenes around a Python loop over kept boxes). Steps, in upstream's order:
        boxes to bottom-centre form; drop boxes holding <= 5 points; class-aware NMS on the
        axis-aligned hulls of the survivors by objectness; drop objectness <= score_thr; with
        per_class_proposal every kept box is emitted once per class with score obj * sem[k]."""
        cfg = self.test_cfg
        counts = P.box_point_count(points, bbox, gravity_centre=True)
        selected, classes, nsel = P.nms_select(bbox, obj_scores, sem_scores, counts, 5, cfg['nms_thr'],
                                               cfg['score_thr'])
        n_list = nsel.tolist()                                            # the ONE host sync
        total = sum(n_list)
        rows = torch.nonzero_static(selected, size=total)                 # (T,2), scene-major: no second sync
        b_idx, r_idx = rows[:, 0], rows[:, 1]
        box_sel = bbox[b_idx, r_idx]
        box_sel[:, 2] = box_sel[:, 2] - box_sel[:, 5] * 0.5              # origin (.5,.5,.5) -> (.5,.5,0)
        obj_sel = obj_scores[b_idx, r_idx]
        cls_sel = classes[b_idx, r_idx]
        starts = np.concatenate([[0], np.cumsum(n_list)]).astype(np.int64)
        if not cfg['per_class_proposal']:
            return [(box_sel[s:e], obj_sel[s:e], cls_sel[s:e]) for s, e in zip(starts[:-1], starts[1:])]
        # every kept box once per class, class-major inside a scene (upstream: a Python loop over classes
        # around boolean indexing, per scene). The gather indices are built on the host from the counts
        # (<= B*R*C small integers, one copy); every scene's result is a slice of three batched gathers.
        C = sem_scores.shape[-1]
        sem_sel = sem_scores[b_idx, r_idx]
        src = np.concatenate([np.tile(np.arange(s, e), C) for s, e in zip(starts[:-1], starts[1:])]
                             + [np.zeros(0, np.int64)])
        cls = np.concatenate([np.repeat(np.arange(C, dtype=np.int64), e - s) for s, e in zip(starts[:-1], starts[1:])]
                             + [np.zeros(0, np.int64)])
        idx = torch.from_numpy(np.stack([src, cls])).to(bbox.device)
        src_t, cls_t = idx[0], idx[1]
        boxes_out = box_sel[src_t]
        scores_out = obj_sel[src_t] * sem_sel[src_t, cls_t]
        return [(boxes_out[C * s:C * e], scores_out[C * s:C * e], cls_t[C * s:C * e])
                for s, e in zip(starts[:-1], starts[1:])]

    # --------------------------------------------------------------------- loss ---
    def loss(self, bbox_preds, *args, **kwargs):
        """Average of _loss over the num_fusion_layers+1 prediction stages (reference :596-620).
        Targets do not depend on the stage, so they are assigned once."""
        bbox_preds = dict(bbox_preds)
        decode_res_all = bbox_preds.pop('decode_res_all')
        common = {k: bbox_preds[k] for k in ('seed_points', 'seed_indices', 'aggregated_points',
                                             'vote_points')}
        early = bbox_preds.pop('_targets', None)
        if early is not None:   # assigned on a second stream while the decoder ran (forward())
            targets, done = early
            cur = torch.cuda.current_stream(decode_res_all[0]['center'].device)
            cur.wait_event(done)
            for t in targets:
                if torch.is_tensor(t):
                    t.record_stream(cur)
        else:
            targets = self.get_targets(args[0], args[1], args[2], bbox_preds=common)
        assert self.num_fusion_layers + 1 == len(decode_res_all)
        first = decode_res_all[0]['center']
        if self.parallel_stage_loss and first.is_cuda and len(decode_res_all) > 1:
            # Each stage's loss is ~150 tiny dependent kernels (and as many again in backward): latency, not
            # work. The stages are independent, so every stage after the first runs on its own stream --
            # parallel branches of the captured step graph; autograd replays each backward on the stream
            # its forward ran on.
            dev = first.device
            cur = torch.cuda.current_stream(dev)
            stages = [None] * len(decode_res_all)
            # the vote loss does not depend on the stage (upstream recomputes it per stage and averages
            # identical values): once, shared
            vote_loss = self._vote_loss(common, targets)
            for i in range(1, len(decode_res_all)):
                side = self._stage_stream(dev, i)
                side.wait_stream(cur)
                # tensors allocated on the main stream and read on the side stream: tell the allocator
                for t in list(common.values()) + list(decode_res_all[i].values()) + list(targets):
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(side)
                with torch.cuda.stream(side):
                    stages[i] = self._loss(dict(common, **decode_res_all[i]), *args, targets=targets,
                                           vote_loss=vote_loss, **kwargs)
            stages[0] = self._loss(dict(common, **decode_res_all[0]), *args, targets=targets,
                                   vote_loss=vote_loss, **kwargs)
            for i in range(1, len(decode_res_all)):
                cur.wait_stream(self._stage_stream(dev, i))
                for v in stages[i].values():
                    v.record_stream(cur)
                if getattr(stages[i], "vec", None) is not None:
                    stages[i].vec.record_stream(cur)
        else:
            stages = [self._loss(dict(common, **decode_res), *args, targets=targets, **kwargs)
                      for decode_res in decode_res_all]
        fused = self._average_fused_stages(stages)
        if fused is not None:
            return fused
        losses = dict()
        for stage in stages:
            for k, v in stage.items():
                losses[k] = losses.get(k, 0) + v / (self.num_fusion_layers + 1)
        return losses

    def _average_fused_stages(self, stages):
        """Stage average when every stage came out of the fused loss kernel as one (7,) vector: the vectors are
        averaged as vectors and the entries of the returned dict are views of the result; `.total` (what the
        trainer differentiates) is one more reduction -- instead of ~30 scalar kernels forward and ~50 backward
        for the same numbers term by term."""
        if not all(isinstance(st, LossDict) and st.vec is not None for st in stages):
            return None
        n = self.num_fusion_layers + 1
        vec = stages[0].vec if len(stages) == 1 else torch.stack([st.vec for st in stages]).sum(0)
        vec = vec / n
        votes = [st['vote_loss'] for st in stages]
        if len(stages) == n and all(v is votes[0] for v in votes):
            vote = votes[0]                       # the shared stage-independent term: n * (v / n)
        else:
            vote = sum(v / n for v in votes)
        out = LossDict(vote_loss=vote)
        for k, i in stages[0].index.items():
            out[k] = vec[i]
        out.total = vec.sum() + vote
        return out

    parallel_stage_loss = True
    early_targets = True
    _stage_streams = {}

    @classmethod
    def _stage_stream(cls, device, i):
        key = (str(device), i)
        if key not in cls._stage_streams:
            cls._stage_streams[key] = torch.cuda.Stream(device=device)
        return cls._stage_streams[key]

    def _vote_loss(self, bbox_preds, targets):
        return self.vote_module.get_loss(bbox_preds['seed_points'], bbox_preds['vote_points'],
                                         bbox_preds['seed_indices'], targets[1], targets[0])

    def _loss(self, bbox_preds, points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask=None,
              pts_instance_mask=None, img_metas=None, gt_bboxes_ignore=None, ret_target=False,
              targets=None, vote_loss=None):
        if targets is None:
            targets = self.get_targets(points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask,
                                       pts_instance_mask, bbox_preds)
        (vote_targets, vote_target_masks, dir_class_targets, dir_res_targets, mask_targets,
         objectness_targets, objectness_weights, box_loss_weights, distance_targets, dir_targets,
         size_targets, center_targets) = targets

        if vote_loss is None:
            vote_loss = self._vote_loss(bbox_preds, targets)
        fused = self._fused_stage_loss(bbox_preds, targets)
        if fused is not None:
            losses = LossDict(vote_loss=vote_loss, **fused)
            if not ret_target:
                losses.vec, losses.index = fused.vec, fused.index
            else:
                losses['targets'] = targets
            return losses
        objectness_loss = self.objectness_loss(bbox_preds['obj_scores'].transpose(2, 1),
                                               objectness_targets, weight=objectness_weights)
        w3 = box_loss_weights.unsqueeze(-1).expand(-1, -1, 3)
        size_reg_loss = self.size_res_loss(bbox_preds['size'], size_targets, weight=w3)
        center_loss = self.center_loss(bbox_preds['center'], center_targets, weight=w3)
        dir_class_loss = self.dir_class_loss(bbox_preds['dir_class'].transpose(2, 1),
                                             dir_class_targets, weight=box_loss_weights)
        dir_res_norm = torch.gather(bbox_preds['dir_res_norm'], 2,
                                    dir_class_targets.unsqueeze(-1)).squeeze(-1)
        dir_res_loss = self.dir_res_loss(dir_res_norm, dir_res_targets, weight=box_loss_weights)
        losses = dict(vote_loss=vote_loss, objectness_loss=objectness_loss,
                      dir_class_loss=dir_class_loss, dir_res_loss=dir_res_loss,
                      size_res_loss=size_reg_loss, center_loss=center_loss)
        if hasattr(self, 'semantic_loss'):
            losses['semantic_loss'] = self.semantic_loss(bbox_preds['sem_scores'].transpose(2, 1),
                                                         mask_targets, weight=box_loss_weights)
        if self.iou_loss:
            corners_pred = self.bbox_coder.decode_corners(bbox_preds['center'], bbox_preds['size'])
            corners_target = self.bbox_coder.decode_corners(center_targets, size_targets)
            losses['iou_loss'] = self.iou_loss(corners_pred, corners_target,
                                               weight=box_loss_weights)
        if ret_target:
            losses['targets'] = targets
        return losses

    fused_stage_loss = True

    def _fused_stage_loss(self, bbox_preds, targets):
        """The seven per-proposal loss sums of one stage in one launch (csrc/loss.cu) when every term is configured
        the way the reference config does (softmax CE / SmoothL1 / AxisAlignedIoU, reduction='sum'); None otherwise."""
        from ..mm.losses import AxisAlignedIoULoss, CrossEntropyLoss, SmoothL1Loss
        center = bbox_preds['center']
        if not (self.fused_stage_loss and center.is_cuda and center.dtype == torch.float32):
            return None
        ce = [self.objectness_loss, self.dir_class_loss] + ([self.semantic_loss] if hasattr(self, 'semantic_loss') else [])
        sl1 = [self.dir_res_loss, self.size_res_loss, self.center_loss]
        if not (all(type(m) is CrossEntropyLoss and m.reduction == 'sum' for m in ce)
                and all(type(m) is SmoothL1Loss and m.reduction == 'sum' for m in sl1)
                and (self.iou_loss is None or (type(self.iou_loss) is AxisAlignedIoULoss
                                               and self.iou_loss.reduction == 'sum'))
                and self.dir_class_loss.class_weight is None
                and (not hasattr(self, 'semantic_loss') or self.semantic_loss.class_weight is None)
                and bbox_preds['dir_class'].shape[-1] <= 16):
            return None
        cw = self.objectness_loss.class_weight or [1.0, 1.0]
        has_sem = hasattr(self, 'semantic_loss')
        cfg = [cw[0], cw[1], self.objectness_loss.loss_weight, self.dir_class_loss.loss_weight,
               self.dir_res_loss.loss_weight, self.size_res_loss.loss_weight, self.center_loss.loss_weight,
               self.semantic_loss.loss_weight if has_sem else 0.0,
               self.iou_loss.loss_weight if self.iou_loss is not None else 0.0,
               self.dir_res_loss.beta, self.size_res_loss.beta, self.center_loss.beta]
        (_, _, dir_class_targets, dir_res_targets, mask_targets, objectness_targets, objectness_weights,
         box_loss_weights, _, _, size_targets, center_targets) = targets
        sem = bbox_preds['sem_scores'].contiguous() if has_sem else None
        out = P.stage_loss(center.contiguous(), bbox_preds['size'].contiguous(), bbox_preds['dir_class'].contiguous(),
                           bbox_preds['dir_res_norm'].contiguous(), bbox_preds['obj_scores'].contiguous(), sem,
                           (objectness_targets, objectness_weights.float(), box_loss_weights.float(), size_targets,
                            center_targets, dir_class_targets, dir_res_targets, mask_targets if has_sem else None), cfg)
        index = dict(objectness_loss=0, dir_class_loss=1, dir_res_loss=2, size_res_loss=3, center_loss=4)
        if has_sem:
            index['semantic_loss'] = 5
        if self.iou_loss is not None:
            index['iou_loss'] = 6
        losses = LossDict((k, out[i]) for k, i in index.items())
        losses.vec, losses.index = out, index    # entries the config leaves out stay 0 in `out`
        return losses

    # ------------------------------------------------------------------ targets ---
    def get_targets(self, points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask=None,
                    pts_instance_mask=None, bbox_preds=None):
        """Batched target assignment (reference :756-941 loops scenes and boxes in Python).
        Scenes are padded to G boxes; an empty scene gets one all-zero box, as upstream."""
        if not self.bbox_coder.with_rot:
            raise NotImplementedError("only the with_rot (SUN RGB-D) target path is implemented")
        points = torch.stack(list(points)) if not torch.is_tensor(points) else points
        dev = points.device
        B = points.shape[0]
        if torch.is_tensor(gt_bboxes_3d):
            # already padded (engine.pad_gt): (B,G,7) boxes, (B,G) labels with -1 = padding
            box = gt_bboxes_3d.to(dev)
            valid = gt_labels_3d.to(dev) >= 0
            label = gt_labels_3d.to(dev).clamp(min=0)
            G = box.shape[1]
        else:
            boxes, labels = [], []
            for b in range(B):
                t = gt_bboxes_3d[b].tensor.to(dev)
                lab = gt_labels_3d[b].to(dev)
                if lab.numel() == 0:
                    t = t.new_zeros(1, 7)
                    lab = lab.new_zeros(1)
                boxes.append(t)
                labels.append(lab.long())
            G = max(t.shape[0] for t in boxes)
            box = points.new_zeros(B, G, 7)
            label = torch.zeros(B, G, dtype=torch.long, device=dev)
            valid = torch.zeros(B, G, dtype=torch.bool, device=dev)
            for b in range(B):
                g = boxes[b].shape[0]
                box[b, :g] = boxes[b]
                label[b, :g] = labels[b]
                valid[b, :g] = True

        centre = box[..., :3].clone()
        centre[..., 2] = centre[..., 2] + box[..., 5] * 0.5      # gravity centre
        size = box[..., 3:6]
        yaw = box[..., 6]

        # ---- vote targets: slot0 = first containing box, slot1 = second (else first),
        #      slot2 = last of >=3 (else first); reference :834-858
        inside = geometry.points_in_boxes_batch(points[..., :3], box).bool() & valid[:, None, :]
        count = inside.sum(-1, dtype=torch.int32)
        gidx = torch.arange(G, device=dev, dtype=torch.int32)
        big = gidx.new_full((), G)
        first = torch.where(inside, gidx, big).amin(-1)                      # lowest box index
        after_first = inside & (gidx > first.unsqueeze(-1))
        second = torch.where(after_first, gidx, big).amin(-1)                # next one
        last = torch.where(inside, gidx, gidx.new_full((), -1)).amax(-1)     # highest
        first = first.clamp(max=G - 1).long()
        second = torch.where(count >= 2, second.long(), first)
        third = torch.where(count >= 3, last.long(), first)
        xyz = points[..., :3]
        slots = []
        for sel in (first, second, third):
            c = torch.gather(centre, 1, sel.unsqueeze(-1).expand(-1, -1, 3))
            slots.append(c - xyz)
        vote_target_masks = (count > 0).long()
        vote_targets = torch.cat(slots, -1) * vote_target_masks.unsqueeze(-1).float()

        # ---- proposal -> nearest GT centre (chamfer l2, reference :868-873)
        agg = bbox_preds['aggregated_points']
        d2 = ((agg[:, :, None, :] - centre[:, None, :, :]) ** 2).sum(-1)
        d2 = d2.masked_fill(~valid[:, None, :], float('inf'))
        distance1, assignment = d2.min(-1)
        euclidean_distance1 = torch.sqrt(distance1 + 1e-6)
        pos_thr, neg_thr = self.train_cfg['pos_distance_thr'], self.train_cfg['neg_distance_thr']
        objectness_masks = ((euclidean_distance1 < pos_thr) | (euclidean_distance1 > neg_thr)).float()

        def take(t):
            idx = assignment if t.dim() == 2 else assignment.unsqueeze(-1).expand(-1, -1, t.shape[-1])
            return torch.gather(t, 1, idx)

        dir_class, dir_res = self.bbox_coder.angle2class(yaw)
        center_targets = take(centre)
        size_targets = take(size)
        dir_class_targets = take(dir_class)
        dir_res_targets = take(dir_res) / (np.pi / self.num_dir_bins)
        dir_targets = take(yaw)
        mask_targets = take(label)

        # ---- inside-the-box test in the box frame (reference :899-935)
        canonical = agg - center_targets
        cosa, sina = torch.cos(dir_targets), torch.sin(dir_targets)
        cx = canonical[..., 0] * cosa - canonical[..., 1] * sina   # rotation_3d_in_axis(-yaw, 2)
        cy = canonical[..., 0] * sina + canonical[..., 1] * cosa
        local = torch.stack([cx, cy, canonical[..., 2]], -1)
        half = size_targets / 2.0
        distance_targets = torch.cat([half - local, half + local], dim=-1)
        inside_mask = (distance_targets >= 0.).all(dim=-1)
        objectness_targets = ((euclidean_distance1 < pos_thr) & inside_mask).long()

        objectness_weights = objectness_masks / (torch.sum(objectness_masks) + 1e-6)
        box_loss_weights = objectness_targets.float() / (torch.sum(objectness_targets).float() + 1e-6)
        return (vote_targets, vote_target_masks, dir_class_targets, dir_res_targets, mask_targets,
                objectness_targets, objectness_weights, box_loss_weights, distance_targets,
                dir_targets, size_targets, center_targets)
