"""Which Python lines issue the library (aten) ops of one training step, forward and backward? (development tool, GPU)
The autograd engine is put on the calling thread so that one TorchDispatchMode sees both directions.
usage: python tools/train_census.py"""
import os
import sys
import traceback
from collections import Counter

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import engine  # noqa: E402

SKIP = ("aten.view", "aten._unsafe_view", "aten.t.", "aten.transpose", "aten.detach", "aten.as_strided", "aten.slice",
        "aten.select", "aten.unsqueeze", "aten.squeeze", "aten.expand", "aten.permute", "aten.alias", "aten.reshape",
        "aten.split", "aten.unbind", "aten._reshape_alias", "aten.empty", "aten.new_empty", "aten.sym_", "aten.stride",
        "aten.is_", "aten.size", "aten.lift_fresh", "aten.unfold", "aten.chunk", "aten.narrow")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Census(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.sites = Counter()
        self.ops = Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not name.startswith(SKIP):
            where = "?"
            for fr in reversed(traceback.extract_stack()):
                if "demf_b200" in fr.filename and "tools/" not in fr.filename:
                    where = f"{fr.filename.replace(root + '/', '')}:{fr.lineno} {fr.name}"
                    break
            self.sites[(where, name)] += 1
            self.ops[name] += 1
        return func(*args, **(kwargs or {}))


dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(99)
model = engine.build_demf_votenet(num_points=4).to(dev).train()
trainer = engine.Trainer(model)
batch = engine.synthetic_batch(4, 20000, "S512", seed=777, device=dev)
for _ in range(3):
    trainer.step(batch)
torch.cuda.synchronize()
torch.autograd.set_multithreading_enabled(False)
with Census() as c:
    trainer.step(batch)
torch.cuda.synchronize()
print("ops:", sum(c.ops.values()))
for k, v in c.ops.most_common(40):
    print(f"  {v:4d}  {k}")
print("by call site:")
for (w, name), v in sorted(c.sites.items(), key=lambda kv: -kv[1])[:150]:
    print(f"  {v:3d}  {name:34s} {w}")
