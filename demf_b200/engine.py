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


def build_demf_votenet(num_points=4, cfg_options=None, init=True, img_encoder=False, img_branch=False):
    """DeMFVoteNet from demf_b200/configs/demf_votenet.py. `num_points` = sampling points per
    level and head of the deformable cross attention (2 in the reference config, 4 in
    BASELINE.json's configs). `img_encoder=True` also builds the frozen Deformable-DETR encoder of
    the image branch; the pyramids given to the model are then the neck's output. `img_branch=True`
    builds the whole frozen branch (ResNet-50 -> ChannelMapper -> encoder, demfnet.py:42-49): `img` is
    then the (B,3,H,W) normalised image batch."""
    cfg = Config.fromfile(CONFIG)
    model_cfg = cfg.model.to_dict()
    if img_encoder or img_branch:
        model_cfg["img_encoder"] = cfg.img_encoder_cfg.to_dict()
    if img_branch:
        model_cfg["img_backbone"] = cfg.img_backbone_cfg.to_dict()
        model_cfg["img_neck"] = cfg.img_neck_cfg.to_dict()
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
    does by default on Ampere and later). The sampling / index kernels are always fp32.
    The fused set-abstraction kernel (csrc/sa_fused.cu) computes TF32 products, so it is part of
    the 'tf32' mode only; 'fp32' runs the levels layer by layer with IEEE fp32 library GEMMs."""
    assert mode in ("fp32", "tf32")
    on = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = on
    torch.backends.cudnn.allow_tf32 = on
    from .mm.pointnet_modules import BasePointSAModule
    BasePointSAModule.fused_eval = on


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

    ALIGN = 4   # elements: every gradient starts on a 16-byte boundary (vector atomics of the wgrad kernel)

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        a = self.ALIGN
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + a - 1) // a * a
        ref = self.params[0]
        self.buffer = torch.zeros(off, dtype=torch.float32, device=ref.device)   # padding stays zero
        for p, g in zip(self.params, self._views()):
            p.grad = g

    def zero(self):
        self.buffer.zero_()
        for p, g in zip(self.params, self._views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g  # re-attach if something replaced / dropped the view

    def _views(self):
        for p, off in zip(self.params, self.offsets):
            yield self.buffer[off:off + p.numel()].view_as(p)

    # Autograd ADDS into a `.grad` that already exists: with every `.grad` a view of the buffer, backward ends in
    # one tiny `add_` launch per parameter (~200 per step). Parameters whose gradient our kernels accumulate in
    # place (bricks marks them `_demf_direct_grad`: the rows convolutions' weights) keep their views; all others
    # hand autograd an empty slot, keep the tensor it produces, and `collect()` moves those into the buffer with
    # one multi-tensor copy.
    def release(self):
        """Before backward: drop the views of the parameters autograd itself accumulates."""
        for p in self.params:
            if not getattr(p, "_demf_direct_grad", False):
                p.grad = None

    def collect(self):
        """After backward: every gradient back in the flat buffer, every `.grad` a view of it again."""
        srcs, dsts = [], []
        for p, v in zip(self.params, self._views()):
            g = p.grad
            if g is not None and g.data_ptr() != v.data_ptr():
                srcs.append(g if g.is_contiguous() else g.contiguous())
                dsts.append(v)
            p.grad = v
        if srcs:
            torch._foreach_copy_(dsts, srcs)

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


def build_optimizer(model, cfg=None, capturable=False):
    """AdamW with mmcv's paramwise `custom_keys` (lr_mult / decay_mult by parameter-name
    substring; configs/demf/demf_votenet.py:16-24). Parameters that end up with the same
    (lr, weight_decay) share one group, and the update is the fused multi-tensor kernel: one
    launch per group instead of ~10 small launches per parameter."""
    if cfg is None:
        cfg = Config.fromfile(CONFIG).optimizer.to_dict()
    cfg = dict(cfg)
    assert cfg.pop("type") == "AdamW"
    custom = (cfg.pop("paramwise_cfg", None) or {}).get("custom_keys", {})
    base_lr, base_wd = cfg["lr"], cfg.get("weight_decay", 0.0)
    groups = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        lr, wd = base_lr, base_wd
        for key in sorted(custom, key=len, reverse=True):
            if key in name:
                lr = base_lr * custom[key].get("lr_mult", 1.0)
                wd = base_wd * custom[key].get("decay_mult", 1.0)
                break
        groups.setdefault((lr, wd), dict(params=[], names=[], lr=lr, weight_decay=wd))
        groups[(lr, wd)]["params"].append(p)
        groups[(lr, wd)]["names"].append(name)
    on_cuda = all(p.is_cuda for g in groups.values() for p in g["params"])
    extra = dict(fused=True, capturable=capturable) if on_cuda else dict(foreach=False)
    groups = list(groups.values())
    if on_cuda and capturable:
        # a captured step reads the learning rate from DEVICE memory: `set_lr` / an LR schedule then
        # takes effect on the next replay (a Python float would be baked into the graph)
        dev = groups[0]["params"][0].device
        for g in groups:
            g["lr_mult"] = g["lr"] / base_lr if base_lr else 1.0
            g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=dev)
        cfg["lr"] = torch.tensor(float(base_lr), dtype=torch.float32, device=dev)
    else:
        for g in groups:
            g["lr_mult"] = g["lr"] / base_lr if base_lr else 1.0
    return torch.optim.AdamW(groups, **cfg, **extra)


def set_lr(optimizer, base_lr):
    """Set the base learning rate; every group keeps its paramwise `lr_mult` (configs/demf/
    demf_votenet.py:17-24). Works for float and device-tensor rates (captured steps), in place."""
    for g in optimizer.param_groups:
        lr = float(base_lr) * g.get("lr_mult", 1.0)
        if torch.is_tensor(g["lr"]):
            g["lr"].fill_(lr)
        else:
            g["lr"] = lr


def step_lr(base_lr, epoch, steps=(24, 32), gamma=0.1):
    """configs/_base_/schedules/schedule_3x.py:5-9: lr_config = dict(policy='step', step=[24, 32])."""
    return base_lr * gamma ** sum(epoch >= s for s in steps)


def broadcast_module_state(model, src=0, group=None):
    """Every parameter and buffer from rank `src` to all ranks (what MMDistributedDataParallel does at
    construction); replicas then start identical whatever each rank's RNG did before."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def module_state_checksum(model):
    """float64 sum of squares over parameters and floating buffers (debug aid: equal across ranks)."""
    acc = 0.0
    for t in list(model.parameters()) + [b for b in model.buffers() if b.is_floating_point()]:
        acc += float(t.detach().double().pow(2).sum())
    return acc


def pad_gt(gt_bboxes_3d, gt_labels_3d, max_gt=None, device=None):
    """Lists of per-scene boxes / labels -> fixed-shape (B,G,7) f32 and (B,G) i64 (label -1 =
    padding). An empty scene gets one all-zero box with label 0, as upstream
    (class_agnostic_vote_head.py:766-775). Fixed shapes are what a captured training step needs."""
    tensors = [b.tensor if hasattr(b, "tensor") else b for b in gt_bboxes_3d]
    G = max([t.shape[0] for t in tensors] + [1])
    if max_gt is not None:
        assert G <= max_gt, f"{G} boxes in a scene > max_gt={max_gt}"
        G = max_gt
    B = len(tensors)
    box = torch.zeros(B, G, 7)
    label = torch.full((B, G), -1, dtype=torch.long)
    for b, (t, lab) in enumerate(zip(tensors, gt_labels_3d)):
        g = t.shape[0]
        if g == 0:
            label[b, 0] = 0
            continue
        box[b, :g] = t.cpu()
        label[b, :g] = lab.cpu().long()
    if device is not None:
        box, label = box.to(device), label.to(device)
    return box, label


class Trainer:
    """forward_train -> sum of losses -> backward into the flat buffer -> one all-reduce ->
    clip (max_norm 10, schedule_3x.py:6) -> AdamW."""

    def __init__(self, model, grad_clip=10.0, group=None, capturable=False, sync_state=True):
        self.model = model
        if sync_state:
            broadcast_module_state(model, group=group)
        self.flat = FlatGradients(model.parameters())
        self.optimizer = build_optimizer(model, capturable=capturable)
        self.base_lr = float(Config.fromfile(CONFIG).optimizer.lr)
        self.grad_clip = grad_clip
        self.group = group

    def set_lr(self, base_lr):
        """New base learning rate (every group keeps its lr_mult); takes effect on the next step, captured
        or not."""
        self.base_lr = float(base_lr)
        set_lr(self.optimizer, base_lr)

    def load_state_dict(self, state_dict, strict=True):
        """Load model weights and re-synchronise the replicas (rank 0's copy wins)."""
        out = self.model.load_state_dict(state_dict, strict=strict)
        broadcast_module_state(self.model, group=self.group)
        return out

    def step(self, batch, sync_collective=False):
        from .mm.bricks import async_weight_grads, deferred_batch_counters
        self.flat.zero()
        self.flat.release()
        # dW GEMMs on a second stream, joined on exit; BatchNorm step counters bumped by one launch on exit
        with async_weight_grads(self.flat.buffer.device), deferred_batch_counters():
            losses = self.model.forward_train(**batch)
            total = getattr(losses, "total", None)     # the head's own reduction when it has one
            if total is None:
                total = sum(losses.values())
            total.backward()
        self.flat.collect()
        work = self.flat.all_reduce_mean(self.group)
        if work is not None:
            work.wait()
        if self.grad_clip:
            self.flat.clip_norm_(self.grad_clip)
        self.optimizer.step()
        return total.detach(), losses


class GraphedTrainStep:
    """One whole training step -- forward, loss, backward, gradient all-reduce, clip, AdamW -- as
    ONE CUDA-graph launch for a fixed batch shape.

    An eager step is ~1 500 kernel launches for ~10 ms of device work and is bound by the host;
    captured, a step is the input copies plus one cudaGraphLaunch. Ground truth is padded to
    `max_gt` boxes per scene (pad_gt) so that every shape is static; the NCCL all-reduce of the
    flat gradient buffer is captured with the rest. The warm-up steps needed before capture are
    undone (parameters, BN buffers and optimizer state are restored), so step 1 of the graph is
    step 1 of training.
    """

    def __init__(self, trainer, example, max_gt=16, warmup=3, pipeline_sampling=True):
        import copy
        from .mm import geometry
        model = trainer.model
        assert model.training
        assert trainer.optimizer.defaults.get("capturable"), \
            "GraphedTrainStep needs Trainer(model, capturable=True): the captured AdamW reads step and lr from device memory"
        dev = example["points"].device
        self.trainer = trainer
        self.max_gt = max_gt
        self.presampled = None
        self._prefetched = None
        self._fold = geometry.fold_projection
        self.points = example["points"].clone()
        self.levels = [lv.clone() for lv in example["img"]]
        self.metas = example["img_metas"]
        mats, affs = geometry.fold_projection(self.metas)
        self._mats_host, self._affs_host = mats.pin_memory(), affs.pin_memory()
        self.mats, self.affs = mats.to(dev), affs.to(dev)
        box, label = pad_gt(example["gt_bboxes_3d"], example["gt_labels_3d"], max_gt)
        self._box_host, self._label_host = box.pin_memory(), label.pin_memory()
        self._staged = None    # recorded behind the last copies out of the pinned staging buffers
        self.box, self.label = box.to(dev), label.to(dev)

        snapshot = copy.deepcopy(model.state_dict())
        self.stream = torch.cuda.Stream(device=dev)
        if pipeline_sampling:
            self._capture_sampler(model, dev, warmup)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss, self.losses = self._eager()
        # undo the warm-up / capture-time updates in place (captured pointers stay valid)
        with torch.no_grad():
            model.load_state_dict(snapshot)
            for state in trainer.optimizer.state.values():
                for v in state.values():
                    if torch.is_tensor(v):
                        v.zero_()
        torch.cuda.synchronize(dev)

    # ---- sampling one batch ahead ------------------------------------------------------------
    # Furthest point sampling is ~2 ms of strictly sequential iterations on a handful of SMs and depends
    # only on the coordinates. Inside the step graph it sits on the critical path in front of the first
    # set-abstraction level; as its own graph on a second stream it runs for batch i+1 underneath the
    # backward pass of batch i. The step graph reads the indices from static buffers.
    @staticmethod
    def _sampling_tensors(samp):
        levels, seed_fps, grids = samp
        ts = [t for idx, xyz, _ in levels for t in (idx, xyz)]
        if seed_fps is not None:
            ts.append(seed_fps[0])
        ts += [g for g in grids if g is not None]
        return ts

    def _capture_sampler(self, model, dev, warmup):
        mod = model.train_cfg['pts']['sample_mod']
        self.points_next = self.points.clone()
        self.sample_stream = torch.cuda.Stream(device=dev)
        self.sample_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.sample_stream), torch.no_grad():
            for _ in range(warmup):
                model.presample(self.points_next, mod)
        torch.cuda.current_stream(dev).wait_stream(self.sample_stream)
        torch.cuda.synchronize(dev)
        self.sample_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.sample_graph, stream=self.sample_stream), torch.no_grad():
            nxt = model.presample(self.points_next, mod)
        self._next_tensors = self._sampling_tensors(nxt)
        # the step graph's own copy: (idx, new_xyz, no event) per level, seed indices, grid
        levels, seed_fps, grids = nxt
        self.presampled = ([(idx.clone(), xyz.clone(), None) for idx, xyz, _ in levels],
                           None if seed_fps is None else (seed_fps[0].clone(), None),
                           [None if g is None else g.clone() for g in grids])
        self._cur_tensors = self._sampling_tensors(self.presampled)
        self._sampled = torch.cuda.Event()
        self._consumed = torch.cuda.Event()

    def _sample_into_next(self, points, stream):
        with torch.cuda.stream(stream):
            self.points_next.copy_(points, non_blocking=True)
            self.sample_graph.replay()

    def _eager(self):
        from .mm.bricks import async_weight_grads, deferred_batch_counters
        t = self.trainer
        t.flat.zero()
        t.flat.release()
        with async_weight_grads(self.points.device), deferred_batch_counters():
            losses = t.model.forward_train(points=self.points, img=self.levels, img_metas=self.metas,
                                           gt_bboxes_3d=self.box, gt_labels_3d=self.label,
                                           projection=(self.mats, self.affs), presampled=self.presampled)
            total = getattr(losses, "total", None)     # the head's own reduction when it has one
            if total is None:
                total = sum(losses.values())
            total.backward()
        t.flat.collect()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(t.group) > 1:
            t.flat.buffer.div_(dist.get_world_size(t.group))
            dist.all_reduce(t.flat.buffer, op=dist.ReduceOp.SUM, group=t.group)
        if t.grad_clip:
            t.flat.clip_norm_(t.grad_clip)
        t.optimizer.step()
        return total.detach(), {k: v.detach() for k, v in losses.items()}

    def load(self, batch):
        self.points.copy_(batch["points"], non_blocking=True)
        for dst, src in zip(self.levels, batch["img"]):
            dst.copy_(src, non_blocking=True)
        mats, affs = self._fold(batch["img_metas"])
        # The pinned staging buffers are rewritten by the HOST: the copies queued from them by the previous call
        # must have run first, or a host that is a step ahead of the device would hand THAT step this batch's
        # matrices / boxes (the host only ever waits here when it is a whole step ahead).
        if self._staged is not None:
            self._staged.synchronize()
        self._mats_host.copy_(mats)
        self._affs_host.copy_(affs)
        self.mats.copy_(self._mats_host, non_blocking=True)
        self.affs.copy_(self._affs_host, non_blocking=True)
        box, label = batch["gt_bboxes_3d"], batch["gt_labels_3d"]
        if not torch.is_tensor(box):
            box, label = pad_gt(box, label, self.max_gt)
        if box.is_cuda:
            self.box.copy_(box, non_blocking=True)
            self.label.copy_(label, non_blocking=True)
        else:
            self._box_host.copy_(box)
            self._label_host.copy_(label)
            self.box.copy_(self._box_host, non_blocking=True)
            self.label.copy_(self._label_host, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record(torch.cuda.current_stream(self.points.device))

    def __call__(self, batch=None, next_batch=None):
        """Run one step (on the current stream) and return the static (total loss, loss dict).
        `next_batch`: the batch of the following call; its sampling chain is started on the second
        stream right away so that it overlaps this step."""
        if batch is not None:
            self.load(batch)
        if self.presampled is not None:
            dev = self.points.device
            main = torch.cuda.current_stream(dev)
            if batch is not None:
                if self._prefetched is not None and self._prefetched is batch["points"]:
                    main.wait_event(self._sampled)
                else:   # nothing prefetched for this batch: sample in line
                    self._sample_into_next(batch["points"], main)
                for dst, src in zip(self._cur_tensors, self._next_tensors):
                    dst.copy_(src, non_blocking=True)
                self._consumed.record(main)
            self._prefetched = None
            if next_batch is not None:
                self.sample_stream.wait_event(self._consumed)
                self.sample_stream.wait_stream(main)     # next_batch's points may be produced on main
                self._sample_into_next(next_batch["points"], self.sample_stream)
                self._sampled.record(self.sample_stream)
                self._prefetched = next_batch["points"]
        self.graph.replay()
        return self.loss, self.losses


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

    def __init__(self, model, example, pool=None, warmup=3, stream=None, copy_stream=None):
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
        self._staged = None    # recorded behind the last copies out of the pinned staging buffers
        self.mats = mats.to(dev)
        self.affs = affs.to(dev)
        self._fold = geometry.fold_projection
        # warm up and capture on the stream this graph will be replayed on: library workspaces
        # (cuBLAS/cuBLASLt) are per stream, so graphs captured on different streams can be in
        # flight at the same time (ForwardPipeline) without sharing one
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        # With a `copy_stream` the forward is captured as TWO graphs, split in front of the first reader of the
        # image features (the decoder's cross attention; heads.forward_points / forward_images): the pyramid -- 95 %
        # of a batch's input bytes -- is copied in on `copy_stream` while the first graph (backbone, votes,
        # aggregation: most of a forward's latency) already runs on the points; the stream waits for the copy
        # between the two launches.
        self.copy_stream = copy_stream
        self.graph_b = None
        self.pts_ready = self.img_ready = self._free = None
        self.graph = torch.cuda.CUDAGraph()
        if copy_stream is None:
            with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool, stream=self.stream):
                self.outputs = self._eager()
            self.pool = self.graph.pool()
            return
        self.pts_ready = torch.cuda.Event()      # points + projection of the batch about to run have landed
        self.img_ready = torch.cuda.Event()      # ... and its pyramid
        self._free = torch.cuda.Event()          # the last replay no longer reads the static pyramid
        with torch.cuda.stream(self.stream), torch.no_grad():
            self._eager_split()
        torch.cuda.synchronize(dev)
        with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool, stream=self.stream):
            self._state = self.model.simple_test_points(self.points, self.metas, projection=(self.mats, self.affs))
        self.pool = self.graph.pool()
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph_b, pool=self.pool, stream=self.stream):
            self.outputs = self.model.simple_test_images(self._state, self.levels, self.metas)

    def _eager(self):
        return self.model.simple_test(points=self.points, img=self.levels, img_metas=self.metas,
                                      projection=(self.mats, self.affs), nms=False)

    def _eager_split(self):
        state = self.model.simple_test_points(self.points, self.metas, projection=(self.mats, self.affs))
        return self.model.simple_test_images(state, self.levels, self.metas)

    def load(self, points, levels, img_metas=None):
        """Copy a new batch (host-pinned or device tensors, same shapes) into the static buffers."""
        if self.copy_stream is None:
            self._copy_points(points, img_metas)
            for dst, src in zip(self.levels, levels):
                dst.copy_(src, non_blocking=True)
            return
        # every input copy of this batch on the copy stream, small ones first: points (+ projection), then the
        # pyramid. (The points must not be copied on the lane's stream: while the copy stream keeps the host link
        # busy with queued pyramids, a small copy issued on another stream waits until that queue has drained --
        # measured: the first kernel of every lane then starts after ALL pyramid copies, 31 ms for 20 steps.)
        self.copy_stream.wait_event(self._free)
        with torch.cuda.stream(self.copy_stream):
            self._copy_points(points, img_metas)
            self.pts_ready.record(self.copy_stream)
            for dst, src in zip(self.levels, levels):
                dst.copy_(src, non_blocking=True)
            self.img_ready.record(self.copy_stream)

    def _copy_points(self, points, img_metas):
        self.points.copy_(points, non_blocking=True)
        if img_metas is not None:
            mats, affs = self._fold(img_metas)
            # the host rewrites the pinned staging buffers: the copies queued from them by the previous load must
            # have run (else a host several submissions ahead would give THAT batch this batch's projection)
            if self._staged is not None:
                self._staged.synchronize()
            self._mats_host.copy_(mats)
            self._affs_host.copy_(affs)
            self.mats.copy_(self._mats_host, non_blocking=True)
            self.affs.copy_(self._affs_host, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(torch.cuda.current_stream(self.device))

    def replay(self):
        cur = torch.cuda.current_stream(self.device)
        if self.graph_b is not None:
            cur.wait_event(self.pts_ready)
        self.graph.replay()
        if self.graph_b is not None:
            cur.wait_event(self.img_ready)
            self.graph_b.replay()
            self._free.record(cur)
        return self.outputs

    def __call__(self, points, levels, img_metas=None):
        self.load(points, levels, img_metas)
        return self.replay()


class ForwardPipeline:
    """`lanes` forward graphs in flight at once, each on its own stream with its own static
    buffers and memory pool.

    One forward is a dependency chain (FPS -> group -> MLP -> ... -> decoder) in which the
    furthest-point sampling kernels are latency-bound and occupy a fraction of each SM: a single
    forward cannot fill a B200. Batches are independent, so consecutive batches alternate
    between lanes: while batch i runs its GEMMs, batch i+1 runs its sampling chain and its H2D
    copies. `submit` returns the lane's static output tensors; they are valid after
    `lane_done(slot).synchronize()` (or any later stream sync) and until that lane is submitted to
    again.

    `late_images=True`: every slot is captured as two graphs (point branch | decoder) and all input copies go
    through one copy stream in submission order, each batch's points first, then its pyramid: a lane starts on
    the points while the pyramid -- 95 % of the input bytes -- is still crossing the host link. Same results;
    a single batch end to end 3.6 -> 2.9 ms, 20 steps from an idle pipeline 20.4 -> 19.8 ms, steady state
    unchanged (the link is the bound). Wants at least two slots per lane (see __init__).
    """

    def __init__(self, model, examples, lanes=2, late_images=False):
        if isinstance(examples, dict):
            examples = [examples] * lanes
        dev = examples[0]["points"].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        # late_images: split graphs (see GraphedForward) and ONE stream for every slot's pyramid copies, so that
        # they cross the host link one after the other. (Copies issued on a stream per lane share the link instead:
        # all pyramids then land at about the same time, the lanes fall into step and the link idles while they all
        # compute -- measured 1.44 ms per step against 0.89.) A copy has to wait until its slot's previous replay
        # has released the static buffers; with at least twice as many slots as lanes that replay is long over and
        # the wait never holds up the copies queued behind it.
        self.copy_streams = [torch.cuda.Stream(device=dev)] * lanes if late_images else None
        self.slots = []          # one captured forward (own static buffers) per example
        pools = [None] * lanes   # slots of one lane never overlap in time: they share a pool
        for j, ex in enumerate(examples):
            lane = j % lanes
            g = GraphedForward(model, ex, pool=pools[lane], stream=self.streams[lane],
                               copy_stream=self.copy_streams[lane] if late_images else None)
            pools[lane] = g.pool
            self.slots.append(g)
        # the first launch of an instantiated graph also uploads it to the device: do that here, not inside the
        # caller's first (possibly timed) round over the slots
        for g in self.slots:
            with torch.cuda.stream(g.stream):
                g.replay()
        torch.cuda.synchronize(dev)
        self._events = [torch.cuda.Event() for _ in self.slots]
        self._next = 0

    def submit(self, points=None, levels=None, img_metas=None, outputs_to=None):
        """Enqueue one batch on the next slot (slots alternate between lanes). With
        points/levels given they are first copied into the slot's static buffers on its lane's
        stream (H2D when they are pinned host tensors); `outputs_to` = pinned host tensors that
        receive the results (D2H on the lane). Returns (slot index, static output tensors)."""
        k = self._next
        self._next = (k + 1) % len(self.slots)
        slot = self.slots[k]
        slot.stream.wait_stream(torch.cuda.current_stream(slot.device))
        with torch.cuda.stream(slot.stream):
            if points is not None:
                slot.load(points, levels, img_metas)
            outs = slot.replay()
            if outputs_to is not None:
                for dst, src in zip(outputs_to, outs):
                    dst.copy_(src, non_blocking=True)
            self._events[k].record(slot.stream)
        return k, outs

    def lane_done(self, k):
        return self._events[k]

    def join(self):
        """Make the current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.slots[0].device)
        for stream in self.streams:
            cur.wait_stream(stream)
        if self.copy_streams:
            cur.wait_stream(self.copy_streams[0])
