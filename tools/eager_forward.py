"""N eager (no CUDA graph) eval forwards at the bench shapes -- the ncu launch-list target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
engine.set_gemm_precision("tf32")
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
batch = engine.synthetic_batch(B, 20000, "S512", seed=1234, device=dev, with_gt=False)
with torch.no_grad():
    for _ in range(n):
        model.simple_test(points=batch["points"], img=batch["img"], img_metas=batch["img_metas"], nms=False)
torch.cuda.synchronize()
print("done")
