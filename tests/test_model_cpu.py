"""Host-side logic of the DeMF(VoteNet) model on CPU (the CUDA ops swapped for the oracle):
registry/config surface, state-dict names, shapes, batched projection and batched target
assignment against straightforward per-scene restatements of the reference code."""
import math

import numpy as np
import pytest
import torch

import demf_b200  # noqa: F401  (registers the model classes)
from demf_b200 import engine, synth
from demf_b200.mm import geometry
from demf_b200.mm.config import Config
from demf_b200.mm.registry import (ATTENTION, BACKBONES, BBOX_CODERS, DETECTORS, HEADS, LOSSES,
                                   SA_MODULES, TRANSFORMER_LAYER)
from oracle.cpu_backend import oracle_ops


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return engine.build_demf_votenet(num_points=4)


def test_registry_names_of_the_reference_config():
    # configs/demf/demf_votenet.py:26-182 type= strings
    assert "DeMFVoteNet" in DETECTORS and "PointNet2SASSG" in BACKBONES
    assert "DeMFVoteHead" in HEADS and "PointSAModule" in SA_MODULES
    for name in ("DeMFTransformerDecoderLayer", "DetrTransformerDecoderLayer", "BaseTransformerLayer"):
        assert name in TRANSFORMER_LAYER
    for name in ("MultiheadAttention", "MultiScaleDeformableAttention"):
        assert name in ATTENTION
    assert "DeMFClassAgnosticBBoxCoder" in BBOX_CODERS
    for name in ("CrossEntropyLoss", "SmoothL1Loss", "AxisAlignedIoULoss", "ChamferDistance"):
        assert name in LOSSES


def test_config_values_match_reference_contract():
    cfg = Config.fromfile(engine.CONFIG)
    bb = cfg.model.pts_backbone
    assert tuple(bb.num_points) == (2048, 1024, 512, 256)
    assert tuple(bb.radius) == (0.2, 0.4, 0.8, 1.2) and tuple(bb.num_samples) == (64, 32, 16, 16)
    head = cfg.model.pts_bbox_head
    assert head.decoder.num_layers == 1
    assert head.decoder.transformerlayers.attn_cfgs[1].num_points == 2  # reference value
    assert head.vote_aggregation_cfg.num_point == 256
    assert cfg.model.train_cfg.pts.sample_mod == "seed"


def test_state_dict_names_and_param_census(model):
    keys = set(model.state_dict())
    for k in ("pts_backbone.SA_modules.0.mlps.0.layer0.conv.weight",
              "pts_backbone.SA_modules.3.mlps.0.layer2.bn.running_var",
              "pts_backbone.FP_modules.1.mlps.layer1.conv.weight",
              "pts_bbox_head.vote_module.vote_conv.0.conv.weight",
              "pts_bbox_head.vote_module.conv_out.bias",
              "pts_bbox_head.vote_aggregation.mlps.0.layer0.conv.weight",
              "pts_bbox_head.conv_pred0.shared_convs.layer0.conv.weight",
              "pts_bbox_head.conv_pred1.conv_reg.weight",
              "pts_bbox_head.decoder.0.layer.attentions.1.sampling_offsets.weight",
              "pts_bbox_head.decoder.0.layer.attentions.1.value_proj.weight",
              "pts_bbox_head.decoder.0.layer.ffns.0.layers.0.0.weight",
              "pts_bbox_head.decoder.0.posembed.position_embedding_head.0.weight"):
        assert k in keys, k
    sd = model.state_dict()
    assert tuple(sd["pts_backbone.SA_modules.0.mlps.0.layer0.conv.weight"].shape) == (64, 4, 1, 1)
    assert tuple(sd["pts_backbone.SA_modules.1.mlps.0.layer0.conv.weight"].shape) == (128, 131, 1, 1)
    n = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert n(model.pts_backbone) == 641920          # SURVEY.md 8a census
    assert n(model.pts_bbox_head.vote_module) == 198659 + 512  # + the two conv biases (bias=True upstream)
    assert n(model.pts_bbox_head.vote_aggregation) == 198912
    assert n(model.pts_bbox_head.conv_pred0) + n(model.pts_bbox_head.conv_pred1) == 110676


def test_forward_shapes_and_loss_keys(model):
    batch = engine.synthetic_batch(2, num_points=4096, pyramid="S512", seed=3)
    with oracle_ops():
        model.eval()
        with torch.no_grad():
            res = model.forward_dummy(points=batch["points"], img=batch["img"],
                                      img_metas=batch["img_metas"])
            boxes, obj, sem = model.simple_test(points=batch["points"], img=batch["img"],
                                                img_metas=batch["img_metas"], nms=False)
        model.train()
        losses = model.forward_train(**batch)
    assert tuple(res["seed_points"].shape) == (2, 1024, 3)
    assert res["seed_indices"].dtype == torch.int64 and tuple(res["seed_indices"].shape) == (2, 1024)
    assert tuple(res["vote_points"].shape) == (2, 1024, 3)
    assert tuple(res["vote_features"].shape) == (2, 256, 1024)
    assert tuple(res["aggregated_points"].shape) == (2, 256, 3)
    assert len(res["decode_res_all"]) == 2
    for r in res["decode_res_all"]:
        assert tuple(r["center"].shape) == (2, 256, 3) and tuple(r["dir_class"].shape) == (2, 256, 12)
        assert tuple(r["obj_scores"].shape) == (2, 256, 2) and tuple(r["sem_scores"].shape) == (2, 256, 10)
    assert tuple(boxes.shape) == (2, 512, 7) and tuple(obj.shape) == (2, 512) and tuple(sem.shape) == (2, 512, 10)
    assert set(losses) == {"vote_loss", "objectness_loss", "dir_class_loss", "dir_res_loss",
                           "size_res_loss", "center_loss", "semantic_loss", "iou_loss"}
    assert all(torch.isfinite(v) for v in losses.values())
    # seed indices refer to the original cloud
    pts = batch["points"]
    seeds = torch.gather(pts[..., :3], 1, res["seed_indices"].unsqueeze(-1).expand(-1, -1, 3))
    assert torch.equal(seeds, res["seed_points"])


def test_rows_path_equals_channel_major_path():
    """PointSAModule on rows (one fused grouping launch + GEMMs) == the upstream-shaped path
    (QueryAndGroup -> Conv2d/BN2d -> max) on the same weights."""
    from demf_b200.mm.pointnet_modules import PointSAModule
    torch.manual_seed(1)
    sa = PointSAModule(mlp_channels=[5, 16, 32], num_point=64, radius=0.5, num_sample=8,
                       normalize_xyz=True).eval()
    for m in sa.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2.0)
    xyz = synth.make_points(2, 500, seed=5)[..., :3].contiguous()
    feats = torch.randn(2, 5, 500)
    with oracle_ops(), torch.no_grad():
        new_xyz, rows_out, idx = sa(xyz, feats)
        sa.groupers[0].return_grouped_idx = True  # forces the channel-major path
        new_xyz2, cm_out, idx2 = sa(xyz, feats)
    assert torch.equal(idx, idx2) and torch.equal(new_xyz, new_xyz2)
    torch.testing.assert_close(rows_out, cm_out, atol=1e-5, rtol=1e-5)


def test_batched_projection_matches_per_scene_reference_chain():
    metas = synth.make_img_metas(6, "REAL", seed=11)
    metas[2]["flip"] = True
    metas[3]["scale_factor"] = [0.8, 0.8, 0.8, 0.8]
    metas[4]["img_crop_offset"] = [3.0, -2.0]
    metas[5]["pcd_trans"] = [0.1, -0.2, 0.05]
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(6, 256, 3, generator=g) * torch.tensor([4.0, 3.0, 2.0]) + torch.tensor([-2.0, 1.0, -1.0])
    ref = []
    for p, meta in zip(xyz, metas):  # reference: class_agnostic_vote_head.py:524-547
        d = geometry.apply_3d_transformation(p, 'DEPTH', meta, reverse=True)
        uv = geometry.points_cam2img(d, d.new_tensor(meta['depth2img']))
        uv = geometry.coord_2d_transform(meta, uv, True)
        uv[:, 0] = uv[:, 0] / (meta['img_shape'][1] - 1)
        uv[:, 1] = uv[:, 1] / (meta['img_shape'][0] - 1)
        ref.append(torch.clamp(uv, 0, 1))
    ref = torch.stack(ref)
    mats, affs = geometry.fold_projection(metas)
    got = geometry.project_batched(xyz, mats, affs)
    torch.testing.assert_close(got, ref, atol=2e-5, rtol=0)
    assert 0.05 < (ref > 0).float().mean() and (ref < 1).float().mean() > 0.05


def _targets_single_reference(head, points, boxes, labels, agg):
    """Per-scene restatement of class_agnostic_vote_head.py:818-941 (python loops)."""
    N = points.shape[0]
    vt = points.new_zeros(N, 9)
    vm = points.new_zeros(N, dtype=torch.long)
    vidx = points.new_zeros(N, dtype=torch.long)
    inside = boxes.points_in_boxes(points)
    centre = boxes.gravity_center
    for i in range(labels.shape[0]):
        ind = torch.nonzero(inside[:, i], as_tuple=False).squeeze(-1)
        vm[ind] = 1
        tmp = vt[ind]
        votes = centre[i].unsqueeze(0) - points[ind, :3]
        for j in range(3):
            col = torch.nonzero(vidx[ind] == j, as_tuple=False).squeeze(-1)
            tmp[col, j * 3:j * 3 + 3] = votes[col]
            if j == 0:
                tmp[col] = votes[col].repeat(1, 3)
        vt[ind] = tmp
        vidx[ind] = torch.clamp(vidx[ind] + 1, max=2)
    c_t, s_t, dc_t, dr_t, d_t = head.bbox_coder.encode(boxes, labels, ret_dir_target=True)
    d2 = ((agg[:, None] - c_t[None]) ** 2).sum(-1)
    dist1, assign = d2.min(1)
    e = torch.sqrt(dist1 + 1e-6)
    om = points.new_zeros(agg.shape[0])
    om[e < 0.3] = 1.0
    om[e > 0.6] = 1.0
    c_t, dc_t, dr_t, s_t, d_t = c_t[assign], dc_t[assign], dr_t[assign] / (np.pi / 12), s_t[assign], d_t[assign]
    can = geometry.rotation_3d_in_axis((agg - c_t).unsqueeze(0).transpose(0, 1), -boxes.yaw[assign], 2).squeeze(1)
    half = s_t / 2
    dist = torch.cat([half - can, half + can], -1)
    ot = ((e < 0.3) & (dist >= 0).all(-1)).long()
    return vt, vm, s_t, c_t, dc_t, dr_t, labels[assign].long(), ot, om, dist, d_t


def test_batched_targets_match_per_scene_reference(model):
    head = model.pts_bbox_head
    B = 3
    pts = synth.make_points(B, 3000, seed=21, clustered=True)
    boxes, labels = engine.synthetic_gt(B, seed=21)
    boxes[1] = boxes[1].new_box(torch.zeros(0, 7))          # an empty scene
    labels[1] = labels[1].new_zeros(0)
    # boxes centred on cloud points, plus overlapping copies so that points fall into 2, 3, 4 boxes
    for b in (0, 2):
        t = boxes[b].tensor
        t[:, :3] = pts[b, torch.arange(len(t)) * 37, :3] - torch.tensor([0.0, 0.0, 0.4])
        t[:, 3:6] = t[:, 3:6].clamp(min=0.8)
        boxes[b] = boxes[b].new_box(t)
    t = boxes[0].tensor
    boxes[0] = boxes[0].new_box(torch.cat([t, t[:1] + 0.05, t[:1] - 0.05, t[:1] + 0.1]))
    labels[0] = torch.cat([labels[0], labels[0][:1].repeat(3)])
    g = torch.Generator().manual_seed(5)
    agg = torch.stack([b.gravity_center[torch.randint(0, max(len(b), 1), (256,), generator=g)]
                       if len(b) else torch.zeros(256, 3) for b in boxes])
    agg = agg + 0.25 * torch.randn(B, 256, 3, generator=g)
    got = head.get_targets(pts, boxes, labels, bbox_preds=dict(aggregated_points=agg))
    (vt, vm, dc, dr, mt, ot, ow, bw, dist, dt, st, ct) = got
    ref = []
    for b in range(B):
        bx, lb = boxes[b], labels[b]
        if len(lb) == 0:
            bx, lb = bx.new_box(torch.zeros(1, 7)), lb.new_zeros(1)
        ref.append(_targets_single_reference(head, pts[b], bx, lb, agg[b]))
    stack = lambda i: torch.stack([r[i] for r in ref])  # noqa: E731
    assert torch.equal(vm, stack(1)) and vm.sum() > 0
    torch.testing.assert_close(vt, stack(0), atol=1e-6, rtol=0)
    assert (vt[..., 0:3] != vt[..., 6:9]).any(), "no point fell into >= 3 boxes: test is vacuous"
    torch.testing.assert_close(st, stack(2))
    torch.testing.assert_close(ct, stack(3))
    assert torch.equal(dc, stack(4)) and torch.equal(mt, stack(6)) and torch.equal(ot, stack(7))
    torch.testing.assert_close(dr, stack(5), atol=1e-6, rtol=0)
    torch.testing.assert_close(dist, stack(9), atol=1e-5, rtol=0)
    om = stack(8)
    torch.testing.assert_close(ow, om / (om.sum() + 1e-6))
    torch.testing.assert_close(bw, ot.float() / (ot.sum().float() + 1e-6))
    assert ot.sum() > 0


def test_adamw_param_groups_follow_custom_keys(model):
    opt = engine.build_optimizer(model)
    lrs = {}
    for group in opt.param_groups:
        for name in group["names"]:
            lrs[name] = (group["lr"], group["weight_decay"])
    assert len(opt.param_groups) == 2      # base group + the `decoder` group (lr_mult 0.05)
    lr, wd = lrs["pts_bbox_head.decoder.0.layer.attentions.1.value_proj.weight"]
    assert math.isclose(lr, 0.008 * 0.05) and math.isclose(wd, 0.01)
    assert math.isclose(lrs["pts_backbone.SA_modules.0.mlps.0.layer0.conv.weight"][0], 0.008)
    n = sum(len(g["params"]) for g in opt.param_groups)
    assert n == sum(1 for p in model.parameters() if p.requires_grad)


def test_padded_gt_gives_the_same_targets(model):
    head = model.pts_bbox_head
    B = 3
    pts = synth.make_points(B, 2500, seed=2, clustered=True)
    boxes, labels = engine.synthetic_gt(B, seed=2)
    boxes[2] = boxes[2].new_box(torch.zeros(0, 7))
    labels[2] = labels[2].new_zeros(0)
    agg = torch.randn(B, 256, 3, generator=torch.Generator().manual_seed(1))
    a = head.get_targets(pts, boxes, labels, bbox_preds=dict(aggregated_points=agg))
    box, lab = engine.pad_gt(boxes, labels, max_gt=16)
    assert tuple(box.shape) == (B, 16, 7) and (lab[2] >= 0).sum() == 1
    b = head.get_targets(pts, box, lab, bbox_preds=dict(aggregated_points=agg))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_async_weight_grads_is_a_noop_on_cpu_and_restores_its_flag():
    from demf_b200.mm import bricks
    assert not bricks._ASYNC_WGRAD["on"]
    with bricks.async_weight_grads("cpu"):
        assert not bricks._ASYNC_WGRAD["on"]          # CUDA only: the CPU path keeps autograd's own accumulation
        cm = bricks.ConvModule(8, 16, 1, conv_cfg=dict(type="Conv1d"), norm_cfg=dict(type="BN1d"))
        x = torch.randn(32, 8, requires_grad=True)
        bricks.conv_module_rows(cm, x).sum().backward()
        assert cm.conv.weight.grad is not None and x.grad is not None
    assert not bricks._ASYNC_WGRAD["on"]


def test_sampling_tensor_list_of_the_graphed_step_follows_the_chain_structure():
    """GraphedTrainStep copies exactly these tensors from the sampler graph's outputs to the step graph's
    inputs: (idx, xyz) per level, the seed indices, then every non-empty grid."""
    lv = [(torch.zeros(2, 4, dtype=torch.int32), torch.zeros(2, 4, 3), None),
          (torch.zeros(2, 2, dtype=torch.int32), torch.zeros(2, 2, 3), None)]
    seed = (torch.zeros(2, 3, dtype=torch.int32), None)
    grids = [torch.zeros(10, dtype=torch.uint8), None]
    ts = engine.GraphedTrainStep._sampling_tensors((lv, seed, grids))
    assert [tuple(t.shape) for t in ts] == [(2, 4), (2, 4, 3), (2, 2), (2, 2, 3), (2, 3), (10,)]
    assert len(engine.GraphedTrainStep._sampling_tensors((lv, None, [None, None]))) == 4


def test_vote_loss_is_stage_independent():
    """DeMFVoteHead.loss computes the vote loss once on the GPU path because upstream's per-stage values are
    identical (same inputs): check that claim on the CPU path, where it is still computed per stage."""
    torch.manual_seed(3)
    model = engine.build_demf_votenet(num_points=4).train()
    batch = engine.synthetic_batch(1, 2048, "S512", seed=5)
    with oracle_ops():
        _, preds = model._forward_head(batch["points"], batch["img"], batch["img_metas"], "seed")
        head = model.pts_bbox_head
        common = {k: preds[k] for k in ("seed_points", "seed_indices", "aggregated_points", "vote_points")}
        pts = torch.stack(list(batch["points"])) if not torch.is_tensor(batch["points"]) else batch["points"]
        targets = head.get_targets(pts, batch["gt_bboxes_3d"], batch["gt_labels_3d"], bbox_preds=common)
        per_stage = [head._loss(dict(common, **d), pts, batch["gt_bboxes_3d"], batch["gt_labels_3d"],
                                targets=targets)["vote_loss"] for d in preds["decode_res_all"]]
        shared = head._vote_loss(common, targets)
    assert all(torch.equal(v, shared) for v in per_stage)


def test_simple_test_with_nms_on_cpu_matches_per_scene_oracle():
    """Detector-level post-processing through the oracle backend: simple_test(nms=True) returns the
    reference's per-scene dicts (demfnet.py:254-283) and they equal the literal per-scene NMS applied to the
    decoded boxes of simple_test(nms=False)."""
    from oracle import postprocess
    torch.manual_seed(7)
    model = engine.build_demf_votenet(num_points=4).eval()
    # random weights give tiny objectness: lower the score threshold so that the selection is not empty
    model.pts_bbox_head.test_cfg['score_thr'] = 0.0
    with torch.no_grad():   # ... and give the boxes a size of about a metre so that they hold points
        for pred in model.pts_bbox_head.conv_preds:
            pred.conv_reg.bias[3:6] = 1.0
    batch = engine.synthetic_batch(2, 2048, "S512", seed=9, with_gt=False)
    kw = dict(points=batch["points"], img_metas=batch["img_metas"], img=batch["img"])
    with torch.no_grad(), oracle_ops():
        box, obj, sem = model.simple_test(nms=False, **kw)
        out = model.simple_test(nms=True, **kw)
    pts = torch.stack(list(batch["points"])) if not torch.is_tensor(batch["points"]) else batch["points"]
    assert len(out) == 2
    total = 0
    for b, res in enumerate(out):
        wb, ws, wl = postprocess.multiclass_nms_single(obj[b], sem[b], box[b], pts[b, :, :3], 0.25, 0.0, True)
        assert set(res) == {"boxes_3d", "scores_3d", "labels_3d"}
        assert torch.equal(res["labels_3d"], wl) and torch.equal(res["boxes_3d"].tensor, wb)
        assert torch.equal(res["scores_3d"], ws)
        total += len(wl)
    assert total > 0
