"""Ordered kernel list of ONE warm eval forward (torch.profiler / CUPTI; eager launches on one stream, so
the durations are warm-cache but serialised). Usage: python tools/forward_kernels.py [--batch 8]"""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demf_b200  # noqa: E402,F401
from demf_b200 import engine  # noqa: E402


def train_step(args, dev):
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile
    model = engine.build_demf_votenet(num_points=4).to(dev).train()
    trainer = engine.Trainer(model, capturable=True)
    batch = engine.synthetic_batch(args.batch, 20000, "S512", seed=1, device=dev)
    box, lab = engine.pad_gt(batch["gt_bboxes_3d"], batch["gt_labels_3d"], 16, dev)
    batch = dict(batch, gt_bboxes_3d=box, gt_labels_3d=lab)
    for _ in range(3):
        trainer.step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        trainer.step(batch)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")
        name = re.sub(r"\(.*", "", name)[:120]
        agg[name][0] += 1
        agg[name][1] += e.time_range.end - e.time_range.start
        n += 1
    total = sum(v[1] for v in agg.values())
    print(f"{n} kernels, {total:.1f} us summed (warm, serialised)")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{us:9.1f} us {100 * us / total:5.1f}%  x{cnt:4d}  {name}")


def train_sections(args, dev):
    """Kernel count and summed kernel time of the sections of one eager training step."""
    from torch.profiler import ProfilerActivity, profile
    model = engine.build_demf_votenet(num_points=4).to(dev).train()
    trainer = engine.Trainer(model, capturable=True)
    batch = engine.synthetic_batch(args.batch, 20000, "S512", seed=1, device=dev)
    box, lab = engine.pad_gt(batch["gt_bboxes_3d"], batch["gt_labels_3d"], 16, dev)
    batch = dict(batch, gt_bboxes_3d=box, gt_labels_3d=lab)
    for _ in range(3):
        trainer.step(batch)
    torch.cuda.synchronize()
    head = model.pts_bbox_head
    state = {}

    def section(name, fn):
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            state[name] = fn()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        us = sum(e.time_range.end - e.time_range.start for e in evs)
        print(f"{name:28s} {len(evs):5d} kernels {us:9.1f} us")

    trainer.flat.zero()
    section("forward (backbone + head)", lambda: model._forward_head(
        batch["points"], batch["img"], batch["img_metas"], "seed"))
    points, preds = state["forward (backbone + head)"]
    common = {k: preds[k] for k in ("seed_points", "seed_indices", "aggregated_points", "vote_points")}
    section("get_targets", lambda: head.get_targets(points, box, lab, bbox_preds=common))
    targets = state["get_targets"]
    section("vote loss", lambda: head._vote_loss(common, targets))
    section("stage losses (2 stages)", lambda: [head._loss(dict(common, **d), points, box, lab, targets=targets,
                                                          vote_loss=state["vote loss"])
                                              for d in preds["decode_res_all"]])
    total = sum(sum(v for v in st.values()) for st in state["stage losses (2 stages)"]) / 2
    section("backward", lambda: total.backward())
    section("clip + AdamW", lambda: (trainer.flat.clip_norm_(trainer.grad_clip), trainer.optimizer.step()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--train", action="store_true", help="one training step instead; prints per-kernel totals")
    ap.add_argument("--sections", action="store_true", help="kernel count / time per section of a training step")
    args = ap.parse_args()
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda:0")
    engine.set_gemm_precision("tf32")
    torch.manual_seed(0)
    if args.sections:
        return train_sections(args, dev)
    if args.train:
        return train_step(args, dev)
    model = engine.build_demf_votenet(num_points=4).to(dev).eval()
    batch = engine.synthetic_batch(args.batch, 20000, "S512", seed=1, device=dev, with_gt=False)
    kw = dict(points=batch["points"], img_metas=batch["img_metas"], img=batch["img"], nms=False)
    with torch.no_grad():
        for _ in range(3):
            model.simple_test(**kw)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model.simple_test(**kw)
            torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    total = 0.0
    for e in evs:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")
        name = re.sub(r"\(.*", "", name)
        dur = e.time_range.end - e.time_range.start
        total += dur
        print(f"{dur:9.1f} us  {name[:110]}")
    print(f"{len(evs)} kernels, {total:.1f} us summed")


if __name__ == "__main__":
    main()
