"""Inference post-processing: DeMFVoteHead.get_bboxes / multiclass NMS.
CPU: the oracle (oracle/postprocess.py) against hand-derived answers and second formulations, and the
head's batched host logic against the per-scene oracle. GPU: the two kernels and the whole path against
the oracle -- masks, counts, labels and box selections must be identical."""
import math

import pytest
import torch

import demf_b200  # noqa: F401
from demf_b200 import engine, synth
from demf_b200.mm import geometry
from demf_b200.mm import point_ops as P
from oracle import postprocess as O
from oracle.cpu_backend import oracle_ops


# ------------------------------------------------------------------ generators ----
_scene_boxes, _scene_points = synth.make_box_predictions, synth.make_points_in_boxes


def _iou_matrix_nms(boxes, scores, classes, thresh):
    """Second formulation of aligned_3d_nms: full K x K IoU matrix, then the greedy sweep."""
    K = boxes.shape[0]
    lo = torch.max(boxes[:, None, :3], boxes[None, :, :3])
    hi = torch.min(boxes[:, None, 3:], boxes[None, :, 3:])
    inter = (hi - lo).clamp(min=0).prod(-1)
    vol = (boxes[:, 3:] - boxes[:, :3]).prod(-1)
    iou = inter / (vol[:, None] + vol[None] - inter) * (classes[:, None] == classes[None]).float()
    order = sorted(range(K), key=lambda i: (-float(scores[i]), -i))
    dead, pick = [False] * K, []
    for a, i in enumerate(order):
        if dead[i]:
            continue
        pick.append(i)
        for j in order[a + 1:]:
            if not dead[j] and not bool(iou[i, j] <= thresh):
                dead[j] = True
    return pick


# -------------------------------------------------------------------- CPU tests ----
def test_nms_known_answers():
    boxes = torch.tensor([[0, 0, 0, 2, 2, 2], [0.2, 0, 0, 2.2, 2, 2], [0.2, 0, 0, 2.2, 2, 2], [5, 5, 5, 6, 6, 6.0]])
    scores = torch.tensor([0.9, 0.8, 0.7, 0.1])
    # same class: box 1 and 2 overlap box 0 with IoU 7.2/8.8 = 0.82 -> dropped; box 3 is far away
    assert O.aligned_3d_nms(boxes, scores, torch.zeros(4, dtype=torch.long), 0.25).tolist() == [0, 3]
    # box 1 of another class survives box 0, then suppresses box 2 (same class, identical box)
    assert O.aligned_3d_nms(boxes, scores, torch.tensor([0, 1, 1, 0]), 0.25).tolist() == [0, 1, 3]
    # touching boxes: intersection volume 0 -> IoU 0 <= thr, both kept
    touch = torch.tensor([[0, 0, 0, 1, 1, 1], [1, 0, 0, 2, 1, 1.0]])
    assert O.aligned_3d_nms(touch, torch.tensor([0.5, 0.6]), torch.zeros(2, dtype=torch.long), 0.25).tolist() == [1, 0]
    # equal scores: the later index is picked first (last element of the stable ascending argsort)
    assert O.aligned_3d_nms(boxes[:3], torch.tensor([0.5, 0.5, 0.5]), torch.zeros(3, dtype=torch.long), 0.25).tolist() == [2]
    # two empty boxes of one class: 0/0 = NaN is not <= thr -> the second is dropped (upstream behaviour)
    empty = torch.tensor([[0, 0, 0, 0, 0, 0], [3, 3, 3, 3, 3, 3.0]])
    assert O.aligned_3d_nms(empty, torch.tensor([0.5, 0.6]), torch.zeros(2, dtype=torch.long), 0.25).tolist() == [1]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_nms_oracle_equals_iou_matrix_formulation(seed):
    box, obj, sem = _scene_boxes(1, 200, seed)
    bc = box[0].clone()
    bc[:, 2] -= bc[:, 5] / 2
    corners = O.box_corners_depth(bc)
    minmax = torch.cat([corners.min(1)[0], corners.max(1)[0]], -1)
    classes = sem[0].argmax(-1)
    want = _iou_matrix_nms(minmax, obj[0], classes, 0.25)
    assert O.aligned_3d_nms(minmax, obj[0], classes, 0.25).tolist() == want
    assert 10 < len(want) < 200


def test_points_in_boxes_known_answers_and_depth_frame_formulation():
    # unit cube on the floor at the origin, then the same cube turned by 45 degrees
    pts = torch.tensor([[0, 0, 0.5], [0.49, 0.49, 0.5], [0.6, 0, 0.5], [0, 0, 1.0], [0, 0, 1.01], [0.5, 0, 0.5],
                        [0.6, 0.0, 0.2], [0.45, 0.45, 0.2]])
    cube = torch.tensor([[0, 0, 0, 1, 1, 1, 0.0], [0, 0, 0, 1, 1, 1, math.pi / 4]])
    got = O.points_in_boxes_depth(pts, cube)
    # z is inclusive (|dz| <= h/2), x/y are strict; the turned cube reaches 0.707 along the axes but not (0.45,0.45)
    assert got[:, 0].tolist() == [1, 1, 0, 1, 0, 0, 0, 1]
    assert got[:, 1].tolist() == [1, 0, 1, 1, 0, 1, 1, 0]
    # product formulation (depth frame, used for target assignment too) on random data
    box, _, _ = _scene_boxes(2, 60, 5)
    bc = box.clone()
    bc[..., 2] -= bc[..., 5] / 2
    points = _scene_points(2, 4000, box, 5)
    prod = geometry.points_in_boxes_batch(points, bc)
    for b in range(2):
        want = O.points_in_boxes_depth(points[b], bc[b])
        differ = (prod[b] != want).sum().item()
        assert differ <= 2          # fp32 vs fp64 on a face
        assert want.sum().item() > 1000


def test_corner_hull_matches_oracle_corners():
    box, _, _ = _scene_boxes(2, 50, 7)
    bc = box.clone()
    bc[..., 2] -= bc[..., 5] / 2
    got = geometry.box_corner_minmax(bc)
    for b in range(2):
        c = O.box_corners_depth(bc[b])
        torch.testing.assert_close(got[b], torch.cat([c.min(1)[0], c.max(1)[0]], -1), atol=1e-6, rtol=0)
    # axis-aligned known answer
    one = torch.tensor([[1.0, 2.0, 0.5, 2.0, 4.0, 1.0, 0.0]])
    assert geometry.box_corner_minmax(one)[0].tolist() == [0.0, 0.0, 0.5, 2.0, 4.0, 1.5]


def _fake_preds(box, obj, sem, head):
    """Two prediction stages whose decode gives back `box`/`obj`/`sem` halves (ensemble_layers [0,1])."""
    B, R, C = sem.shape
    Q = R // 2
    nb = head.bbox_coder.num_dir_bins
    per = 2 * math.pi / nb
    stages = []
    for s in range(2):
        sl = slice(s * Q, (s + 1) * Q)
        ang = box[:, sl, 6]
        cls = torch.floor((ang + per / 2) % (2 * math.pi) / per).long()
        res = ang - (cls.float() * per)
        res = torch.where(res > math.pi, res - 2 * math.pi, res)
        dir_class = torch.nn.functional.one_hot(cls, nb).float() * 5
        dir_res = torch.zeros(B, Q, nb).scatter(2, cls[..., None], res[..., None])
        o = obj[:, sl].clamp(1e-4, 1 - 1e-4)
        stages.append(dict(center=box[:, sl, :3], size=box[:, sl, 3:6], dir_class=dir_class, dir_res=dir_res,
                           obj_scores=torch.stack([torch.zeros_like(o), torch.log(o / (1 - o))], -1),
                           sem_scores=torch.log(sem[:, sl])))
    return dict(decode_res_all=stages)


@pytest.fixture(scope="module")
def head():
    return engine.build_demf_votenet().pts_bbox_head


@pytest.mark.parametrize("per_class", [True, False])
def test_get_bboxes_host_logic_equals_per_scene_oracle(head, per_class):
    B, R, N = 3, 64, 3000
    box, obj, sem = _scene_boxes(B, R, 11)
    points = _scene_points(B, N, box, 11)
    preds = _fake_preds(box, obj, sem, head)
    old = head.test_cfg['per_class_proposal']
    head.test_cfg['per_class_proposal'] = per_class
    try:
        with oracle_ops():
            results = head.get_bboxes(points, preds, [dict() for _ in range(B)])
            raw = head.get_bboxes(points, preds, None, use_nms=False)
        dbox, dobj, dsem = head.decode_ensemble(preds)
    finally:
        head.test_cfg['per_class_proposal'] = old
    assert torch.equal(raw, dbox) and raw.shape == (B, R, 7)
    torch.testing.assert_close(dbox[..., :6], box[..., :6], atol=1e-6, rtol=0)
    torch.testing.assert_close(dobj, obj.clamp(1e-4, 1 - 1e-4), atol=1e-5, rtol=0)
    total = 0
    for b in range(B):
        wb, ws, wl = O.multiclass_nms_single(dobj[b], dsem[b], dbox[b], points[b], 0.25, 0.05, per_class)
        gb, gs, gl = results[b]
        assert isinstance(gb, geometry.DepthBoxes)
        assert torch.equal(gb.tensor, wb) and torch.equal(gs, ws) and torch.equal(gl, wl)
        total += len(wl)
    assert total > 0


def test_nms_batch_with_empty_scenes(head):
    """No box selected anywhere / in one scene: empty per-scene results of the right shapes, the other
    scenes unaffected (the batched gathers are built from per-scene counts)."""
    box, obj, sem = _scene_boxes(2, 32, 1)
    pts = _scene_points(2, 500, box, 1)
    with oracle_ops():
        none = head.multiclass_nms_batch(obj * 0.0, sem, box, pts)
        obj2 = obj.clone()
        obj2[0] = 0
        half = head.multiclass_nms_batch(obj2, sem, box, pts)
        full = head.multiclass_nms_batch(obj, sem, box, pts)
    for b, s, l in none:
        assert b.shape == (0, 7) and s.shape == (0,) and l.shape == (0,)
    assert half[0][0].shape == (0, 7) and len(half[1][2]) > 0
    for a, b in zip(half[1], full[1]):
        assert torch.equal(a, b)


# -------------------------------------------------------------------- GPU tests ----
@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K,stride", [(2, 20000, 512, 3), (1, 777, 5, 4), (3, 64, 33, 6), (1, 0, 4, 3)])
def test_box_point_count_matches_oracle(dev, B, N, K, stride):
    box, _, _ = _scene_boxes(B, K, K)
    bc = box.clone()
    bc[..., 2] -= bc[..., 5] / 2
    xyz = _scene_points(B, N, box, N) if N else torch.zeros(B, 0, 3)
    rows = torch.cat([xyz, torch.rand(B, N, stride - 3)], -1).contiguous()
    got = P.box_point_count(rows.to(dev), bc.to(dev)).cpu()
    assert got.shape == (B, K) and got.dtype == torch.int32
    # (a) same arithmetic as the torch formulation on the device: identical
    if N:
        same = geometry.points_in_boxes_batch(rows.to(dev)[..., :3], bc.to(dev)).sum(1).cpu().to(torch.int32)
        assert (got - same).abs().max().item() <= 1 and (got != same).sum().item() <= 2    # cos/sin last-ulp
    # (b) the fp64 oracle: at most a point on a face differs
    for b in range(B):
        want = O.points_in_boxes_depth(xyz[b], bc[b]).sum(0).to(torch.int32)
        assert (got[b] - want).abs().max().item() <= 1 if N else got[b].abs().sum().item() == 0
        assert (got[b] != want).sum().item() <= max(2, K // 50)


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,valid_frac", [(4, 512, 0.7), (1, 1, 1.0), (2, 37, 0.5), (2, 1000, 1.0), (1, 4096, 0.9),
                                            (2, 16, 0.0)])
def test_aligned_nms_matches_oracle(dev, B, K, valid_frac):
    box, obj, sem = _scene_boxes(B, K, K + 1)
    bc = box.clone()
    bc[..., 2] -= bc[..., 5] / 2
    minmax = geometry.box_corner_minmax(bc)
    classes = sem.argmax(-1)
    g = torch.Generator().manual_seed(K)
    valid = torch.rand(B, K, generator=g) < valid_frac
    obj[:, : K // 4] = obj[:, K // 4: 2 * (K // 4)]             # duplicated scores: exercises the tie rule
    got = P.aligned_3d_nms(minmax.to(dev), obj.to(dev), classes.to(dev), valid.to(dev), 0.25).cpu()
    assert got.dtype == torch.bool and got.shape == (B, K)
    for b in range(B):
        inds = valid[b].nonzero().flatten()
        picked = O.aligned_3d_nms(minmax[b][inds], obj[b][inds], classes[b][inds], 0.25)
        want = torch.zeros(K, dtype=torch.bool)
        want[inds[picked]] = True
        assert torch.equal(got[b], want)


@pytest.mark.gpu
def test_aligned_nms_edge_cases(dev):
    # empty boxes (NaN IoU), and more boxes than the kernel supports
    empty = torch.tensor([[[0, 0, 0, 0, 0, 0], [3, 3, 3, 3, 3, 3.0], [0, 0, 0, 1, 1, 1.0]]])
    keep = P.aligned_3d_nms(empty.to(dev), torch.tensor([[0.5, 0.6, 0.4]], device=dev),
                            torch.zeros(1, 3, dtype=torch.long, device=dev), torch.ones(1, 3, dtype=torch.bool, device=dev), 0.25)
    # box 1 (empty) against box 0 (empty): 0/0 = NaN is not <= thr -> dropped; against the unit cube: IoU 0
    assert keep.cpu().tolist() == [[False, True, True]]
    assert O.aligned_3d_nms(empty[0], torch.tensor([0.5, 0.6, 0.4]), torch.zeros(3, dtype=torch.long), 0.25).tolist() == [1, 2]
    with pytest.raises(RuntimeError):
        P.aligned_3d_nms(torch.zeros(1, 5000, 6, device=dev), torch.zeros(1, 5000, device=dev),
                         torch.zeros(1, 5000, dtype=torch.long, device=dev),
                         torch.ones(1, 5000, dtype=torch.bool, device=dev), 0.25)
    assert P.aligned_3d_nms(torch.zeros(0, 8, 6, device=dev), torch.zeros(0, 8, device=dev),
                            torch.zeros(0, 8, dtype=torch.long, device=dev),
                            torch.zeros(0, 8, dtype=torch.bool, device=dev), 0.25).shape == (0, 8)


@pytest.mark.gpu
@pytest.mark.parametrize("per_class", [True, False])
def test_get_bboxes_gpu_equals_oracle(dev, head, per_class):
    """BASELINE geometry: 512 ensembled proposals against 20 000 points per scene."""
    B, R, N = 4, 512, 20000
    box, obj, sem = _scene_boxes(B, R, 21)
    points = _scene_points(B, N, box, 21)
    preds = _fake_preds(box, obj, sem, head)
    preds_dev = dict(decode_res_all=[{k: v.to(dev) for k, v in st.items()} for st in preds['decode_res_all']])
    old = head.test_cfg['per_class_proposal']
    head.test_cfg['per_class_proposal'] = per_class
    try:
        results = head.get_bboxes(points.to(dev), preds_dev, [dict() for _ in range(B)])
        dbox, dobj, dsem = (t.cpu() for t in head.decode_ensemble(preds_dev))
    finally:
        head.test_cfg['per_class_proposal'] = old
    n_sel = 0
    for b in range(B):
        wb, ws, wl = O.multiclass_nms_single(dobj[b], dsem[b], dbox[b], points[b], 0.25, 0.05, per_class)
        gb, gs, gl = results[b]
        assert gb.tensor.is_cuda
        assert torch.equal(gl.cpu(), wl)
        torch.testing.assert_close(gb.tensor.cpu(), wb, atol=1e-6, rtol=0)
        torch.testing.assert_close(gs.cpu(), ws, atol=1e-6, rtol=0)
        n_sel += len(wl)
    assert n_sel > 20


@pytest.mark.gpu
def test_simple_test_with_nms_end_to_end(dev):
    torch.manual_seed(0)
    model = engine.build_demf_votenet().to(dev).eval()
    B = 2
    pts = synth.make_points(B, 20000, seed=3, clustered=True).to(dev)
    pyr = [f.to(dev) for f in synth.make_pyramid(B, "S512")]
    metas = synth.make_img_metas(B, "S512")
    with torch.no_grad():
        box, obj, sem = model.simple_test(points=pts, img_metas=metas, img=pyr, nms=False)
        out = model.simple_test(points=[p for p in pts], img_metas=metas, img=pyr, nms=True)
    assert len(out) == B
    for b in range(B):
        wb, ws, wl = O.multiclass_nms_single(obj[b].cpu(), sem[b].cpu(), box[b].cpu(), pts[b, :, :3].cpu(),
                                             0.25, 0.05, True)
        r = out[b]
        assert set(r) == {"boxes_3d", "scores_3d", "labels_3d"} and not r["scores_3d"].is_cuda
        assert torch.equal(r["labels_3d"], wl)
        torch.testing.assert_close(r["boxes_3d"].tensor, wb, atol=1e-6, rtol=0)
        torch.testing.assert_close(r["scores_3d"], ws, atol=1e-6, rtol=0)
