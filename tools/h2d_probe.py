"""Pinned H2D bandwidth of this box at 1..N concurrent ranks (run under torchrun for N>1)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
numa = bench.pin_to_gpu_numa_node(lr) if os.environ.get("PIN", "1") == "1" else None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
def barrier():
    if world > 1:
        dist.barrier(device_ids=[lr])
    torch.cuda.synchronize()
for mb in (47, 256):
    g = bench.host_link_probe(dev, mb << 20, 20, barrier)
    t = torch.tensor([g], device=dev, dtype=torch.float64)
    if world > 1:
        gs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gs, t); per = [float(x) for x in gs]
    else:
        per = [g]
    if rank == 0:
        print(json.dumps({"h2d_probe_MB": mb, "ranks": world, "per_rank_gbs": [round(x, 1) for x in per], "aggregate_gbs": round(sum(per), 1), "numa": numa}), flush=True)
if world > 1:
    dist.destroy_process_group()
