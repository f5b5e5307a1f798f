"""Time the graphed training step (batch 4) under a torch BLAS preference: python tools/train_probe.py [cublas|cublaslt]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demf_b200  # noqa: E402,F401
from demf_b200 import engine  # noqa: E402


def main():
    pref = sys.argv[1] if len(sys.argv) > 1 else "default"
    if pref != "default":
        torch.backends.cuda.preferred_blas_library(pref)
    dev = torch.device("cuda:0")
    engine.set_gemm_precision("tf32")
    torch.manual_seed(0)
    model = engine.build_demf_votenet(num_points=4).to(dev).train()
    trainer = engine.Trainer(model, capturable=True)
    sets = [engine.synthetic_batch(4, 20000, "S512", seed=7 + i, device=dev) for i in range(4)]
    for ts in sets:
        ts["gt_bboxes_3d"], ts["gt_labels_3d"] = engine.pad_gt(ts["gt_bboxes_3d"], ts["gt_labels_3d"], 16, dev)
    step = engine.GraphedTrainStep(trainer, sets[0], max_gt=16)
    for i in range(4):
        step(sets[i % 4], next_batch=sets[(i + 1) % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 30
    for i in range(n):
        loss, _ = step(sets[i % 4], next_batch=sets[(i + 1) % 4])
    b.record()
    b.synchronize()
    print(f"blas={pref}: {a.elapsed_time(b) / n:.3f} ms/step, loss {float(loss):.3f}")


if __name__ == "__main__":
    main()
