import sys, os
sys.path.insert(0, os.getcwd())
import torch
from demf_b200 import engine
dev = torch.device("cuda:0")
torch.manual_seed(4)
model = engine.build_demf_votenet(num_points=4).to(dev).train()
model.pts_backbone.overlap_sampling = sys.argv[1] == "1"
trainer = engine.Trainer(model, capturable=True)
batch = engine.synthetic_batch(2, 20000, "S512", seed=20, device=dev)
step = engine.GraphedTrainStep(trainer, batch, max_gt=16)
for i in range(3):
    total, losses = step(batch)
torch.cuda.synchronize()
print("OK", total.item())
