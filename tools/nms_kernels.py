"""Ordered kernel list of one DeMFVoteHead.multiclass_nms_batch call (batch 8, 512 boxes, 20 000 points)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demf_b200  # noqa: E402,F401
from demf_b200 import engine, synth  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda:0")
    head = engine.build_demf_votenet().pts_bbox_head
    B = 8
    box, obj, sem = (t.to(dev) for t in synth.make_box_predictions(B, 512, seed=5))
    pts = synth.make_points_in_boxes(B, 20000, box.cpu(), seed=5).to(dev)
    for _ in range(3):
        head.multiclass_nms_batch(obj, sem, box, pts)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        head.multiclass_nms_batch(obj, sem, box, pts)
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    for e in evs:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")[:90]
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {name}")
    print(len(evs), "kernels; span", evs[-1].time_range.end - t0, "us")


if __name__ == "__main__":
    main()
