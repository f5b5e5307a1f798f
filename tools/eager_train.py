"""N eager training steps (batch 4, bench shapes) -- ncu launch-list / host-profile target."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
engine.set_gemm_precision("tf32")
torch.manual_seed(99)
dev = torch.device("cuda:0")
model = engine.build_demf_votenet(num_points=4).to(dev).train()
trainer = engine.Trainer(model)
batch = engine.synthetic_batch(4, 20000, "S512", seed=777, device=dev)
for _ in range(2):
    trainer.step(batch)
torch.cuda.synchronize()
if os.environ.get("HOSTPROF"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(n):
        trainer.step(batch)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
else:
    t0 = time.perf_counter()
    for _ in range(n):
        trainer.step(batch)
    torch.cuda.synchronize()
    print("ms/step", 1e3 * (time.perf_counter() - t0) / n)
