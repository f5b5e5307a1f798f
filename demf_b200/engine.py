"""Building, feeding and stepping DeMF(VoteNet) on B200s.

What the reference gets from mmdet3d's `build_model` + `train_model` (train.py:107-147) for this
path: build the detector from the config, make optimizer groups (AdamW, `decoder` params at
lr_mult 0.05, configs/demf/demf_votenet.py:16-24), wrap for data parallelism and step. Data
parallelism here is one process per GPU and ONE flat fp32 gradient buffer: every trainable
parameter's `.grad` is a view into it, so the only collective of a step is a single NCCL
all-reduce (mean) of ~8.8 MB over NVLink -- no bucketing, no hooks, no unused-parameter search
(the frozen image branch never enters the buffer, which is what `find_unused_parameters=True`
works around upstream, demf_votenet.py:279).
"""
import os

import torch
import torch.distributed as dist

from . import synth
from .mm.config import Config
from .mm.geometry import DepthBoxes
from .mm.registry import build_model

CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "demf_votenet.py")


def build_demf_votenet(num_points=4, cfg_options=None, init=True):
    """DeMFVoteNet from demf_b200/configs/demf_votenet.py. `num_points` = sampling points per
    level and head of the deformable cross attention (2 in the reference config, 4 in
    BASELINE.json's configs)."""
    cfg = Config.fromfile(CONFIG)
    model_cfg = cfg.model.to_dict()
    model_cfg["pts_bbox_head"]["decoder"]["transformerlayers"]["attn_cfgs"][1]["num_points"] = num_points
    if cfg_options:
        c = Config(dict(model=model_cfg))
        c.merge_from_dict({("model." + k): v for k, v in cfg_options.items()})
        model_cfg = c.model.to_dict()
    model = build_model(model_cfg)
    if init:
        model.init_weights()
    return model


def set_gemm_precision(mode):
    """Arithmetic of the dense projections (library GEMMs): 'fp32' = IEEE fp32 FMA,
    'tf32' = fp32 storage with TF32 tensor-core products (what the reference's PyTorch 1.8
    does by default on Ampere and later). The sampling / index kernels are always fp32."""
    assert mode in ("fp32", "tf32")
    on = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = on
    torch.backends.cudnn.allow_tf32 = on


# ----------------------------------------------------------------------------- data ---
def synthetic_gt(B, seed=0, min_boxes=3, max_boxes=8, mean_sizes=None):
    """Random SUN-RGB-D-like ground truth: per scene a DepthBoxes (G,7) and labels (G,)."""
    g = torch.Generator().manual_seed(seed + 32452843)
    if mean_sizes is None:
        mean_sizes = Config.fromfile(CONFIG).model.pts_bbox_head.bbox_coder.mean_sizes
    mean_sizes = torch.tensor(mean_sizes)
    boxes, labels = [], []
    for _ in range(B):
        n = int(torch.randint(min_boxes, max_boxes + 1, (1,), generator=g))
        lab = torch.randint(0, mean_sizes.shape[0], (n,), generator=g)
        size = mean_sizes[lab] * (0.8 + 0.4 * torch.rand(n, 3, generator=g))
        xy = torch.rand(n, 2, generator=g) * 5.0 - 2.5
        z = torch.rand(n, 1, generator=g) * 0.5
        yaw = (torch.rand(n, 1, generator=g) * 2 - 1) * 3.14159
        boxes.append(DepthBoxes(torch.cat([xy, z, size, yaw], -1)))
        labels.append(lab)
    return boxes, labels


def synthetic_batch(B, num_points=20000, pyramid="S512", seed=0, device=None, clustered=True,
                    with_gt=True, pin=False):
    """One batch in forward_train's keyword form: points (B,N,4), img = the 4-level pyramid the
    frozen image branch would emit, img_metas, gt boxes / labels."""
    points = synth.make_points(B, num_points, seed=seed, clustered=clustered)
    levels = synth.make_pyramid(B, pyramid, seed=seed)
    metas = synth.make_img_metas(B, pyramid, seed=seed)
    if pin:
        points = points.pin_memory()
        levels = [lv.pin_memory() for lv in levels]
    if device is not None:
        points = points.to(device, non_blocking=True)
        levels = [lv.to(device, non_blocking=True) for lv in levels]
    batch = dict(points=points, img=levels, img_metas=metas)
    if with_gt:
        boxes, labels = synthetic_gt(B, seed=seed)
        if device is not None:
            boxes = [b.to(device) for b in boxes]
            labels = [lab.to(device) for lab in labels]
        batch.update(gt_bboxes_3d=boxes, gt_labels_3d=labels)
    return batch


# ------------------------------------------------------------------------ training ---
class FlatGradients:
    """All trainable gradients in one contiguous fp32 buffer (params keep their own storage)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.buffer = torch.zeros(total, dtype=torch.float32, device=ref.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.buffer[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.buffer.zero_()
        for p, g in zip(self.params, self._views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g  # re-attach if something replaced / dropped the view

    def _views(self):
        off = 0
        for p in self.params:
            n = p.numel()
            yield self.buffer[off:off + n].view_as(p)
            off += n

    def all_reduce_mean(self, group=None):
        """The step's only collective. Returns the async work handle (None when not distributed)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        self.buffer.div_(dist.get_world_size(group))
        return dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def clip_norm_(self, max_norm, norm_type=2):
        assert norm_type == 2
        norm = torch.linalg.vector_norm(self.buffer)
        self.buffer.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


def build_optimizer(model, cfg=None):
    """AdamW with mmcv's paramwise `custom_keys` (lr_mult / decay_mult by parameter-name
    substring; configs/demf/demf_votenet.py:16-24)."""
    if cfg is None:
        cfg = Config.fromfile(CONFIG).optimizer.to_dict()
    cfg = dict(cfg)
    assert cfg.pop("type") == "AdamW"
    custom = (cfg.pop("paramwise_cfg", None) or {}).get("custom_keys", {})
    base_lr, base_wd = cfg["lr"], cfg.get("weight_decay", 0.0)
    groups = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        group = dict(params=[p], lr=base_lr, weight_decay=base_wd)
        for key in sorted(custom, key=len, reverse=True):
            if key in name:
                group["lr"] = base_lr * custom[key].get("lr_mult", 1.0)
                group["weight_decay"] = base_wd * custom[key].get("decay_mult", 1.0)
                break
        groups.append(group)
    return torch.optim.AdamW(groups, **cfg, foreach=True)


class Trainer:
    """forward_train -> sum of losses -> backward into the flat buffer -> one all-reduce ->
    clip (max_norm 10, schedule_3x.py:6) -> AdamW."""

    def __init__(self, model, grad_clip=10.0, group=None):
        self.model = model
        self.flat = FlatGradients(model.parameters())
        self.optimizer = build_optimizer(model)
        self.grad_clip = grad_clip
        self.group = group

    def step(self, batch):
        self.flat.zero()
        losses = self.model.forward_train(**batch)
        total = sum(losses.values())
        total.backward()
        work = self.flat.all_reduce_mean(self.group)
        if work is not None:
            work.wait()
        if self.grad_clip:
            self.flat.clip_norm_(self.grad_clip)
        self.optimizer.step()
        return total.detach(), losses


# ----------------------------------------------------------------------- inference ---
class GraphedForward:
    """`model.simple_test` for a fixed batch shape as ONE CUDA-graph launch.

    The eager forward is ~290 kernel launches (ours + library GEMMs + element-wise glue) for
    ~5 ms of device work: on a B200 the host cannot issue them fast enough, the GPU idles between
    launches. Capturing the whole forward -- including the side-stream FPS chain, which becomes a
    parallel branch of the graph -- removes the host from the loop: a step is the input copies
    into the static buffers plus one cudaGraphLaunch.

    Static inputs: points (B,N,4), the pyramid levels, and the folded projection (mats (B,3,4),
    affs (B,4)) that `geometry.fold_projection(img_metas)` computes on the host per batch.
    """

    def __init__(self, model, example, pool=None, warmup=3):
        assert not model.training, "GraphedForward captures the eval-mode forward"
        dev = example["points"].device
        assert dev.type == "cuda"
        from .mm import geometry
        self.model = model
        self.device = dev
        self.points = example["points"].clone()
        self.levels = [lv.clone() for lv in example["img"]]
        self.metas = example["img_metas"]
        mats, affs = geometry.fold_projection(self.metas)
        self._mats_host = mats.pin_memory()
        self._affs_host = affs.pin_memory()
        self.mats = mats.to(dev)
        self.affs = affs.to(dev)
        self._fold = geometry.fold_projection
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool):
            self.outputs = self._eager()
        self.pool = self.graph.pool()

    def _eager(self):
        return self.model.simple_test(points=self.points, img=self.levels, img_metas=self.metas,
                                      projection=(self.mats, self.affs))

    def load(self, points, levels, img_metas=None):
        """Copy a new batch (host-pinned or device tensors, same shapes) into the static buffers."""
        self.points.copy_(points, non_blocking=True)
        for dst, src in zip(self.levels, levels):
            dst.copy_(src, non_blocking=True)
        if img_metas is not None:
            mats, affs = self._fold(img_metas)
            self._mats_host.copy_(mats)
            self._affs_host.copy_(affs)
            self.mats.copy_(self._mats_host, non_blocking=True)
            self.affs.copy_(self._affs_host, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.outputs

    def __call__(self, points, levels, img_metas=None):
        self.load(points, levels, img_metas)
        return self.replay()
