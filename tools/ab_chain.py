"""A/B on one box: the sampling-chain shortcut on/off -- pipeline throughput (8 graphs in flight) and
single-batch latency of the eval forward at the bench shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import engine
from demf_b200.mm.pointnet_modules import PointNet2SASSG

dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
sets = [engine.synthetic_batch(8, 20000, "S512", seed=1234 + i, device=dev, with_gt=False) for i in range(8)]
for rep in range(2):
    for shortcut in (False, True):
        PointNet2SASSG.chain_shortcut = shortcut
        with torch.no_grad():
            for i in range(3):
                model.simple_test(points=sets[i]["points"], img=sets[i]["img"], img_metas=sets[i]["img_metas"], nms=False)
        pipe = engine.ForwardPipeline(model, sets, lanes=8)
        for _ in range(16):
            pipe.submit()
        pipe.join(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            pipe.submit()
        pipe.join(); b.record(); b.synchronize()
        thr = a.elapsed_time(b) / 200
        a.record()
        for i in range(100):
            pipe.slots[i % 8].replay()
        b.record(); b.synchronize()
        lat = a.elapsed_time(b) / 100
        print(f"chain_shortcut={shortcut}: {thr:.4f} ms/step ({8 / thr * 1e3:.0f} scenes/s), single batch {lat:.3f} ms", flush=True)
        del pipe
