"""e2e leg only: pinned host batches -> ForwardPipeline(late_images on/off) -> pinned host results; K steps timed from
an idle pipeline (as the driver's 20-step window does), results of both modes compared."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import engine
dev = torch.device("cuda:0")
engine.set_gemm_precision("tf32")
torch.manual_seed(0)
model = engine.build_demf_votenet(num_points=4).to(dev).eval()
L = int(os.environ.get("LANES", "10"))
S = int(os.environ.get("SLOTS", str(2 * L)))
sets = [engine.synthetic_batch(8, 20000, "S512", seed=1 + i, device=dev, with_gt=False) for i in range(S)]
host = [engine.synthetic_batch(8, 20000, "S512", seed=100 + i, device="cpu", with_gt=False, pin=True) for i in range(S)]
with torch.no_grad():
    for i in range(3):
        model.simple_test(points=sets[i]["points"], img=sets[i]["img"], img_metas=sets[i]["img_metas"], nms=False)
res = {}
for late in (False, True):
    pipe = engine.ForwardPipeline(model, sets, lanes=L, late_images=late)
    outs = [[torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory() for t in pipe.slots[0].outputs] for _ in range(S)]
    def run(K):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            hb = host[i % S]
            pipe.submit(hb["points"], hb["img"], hb["img_metas"], outputs_to=outs[i % S])
        pipe.join()
        b.record(); b.synchronize()
        return a.elapsed_time(b)
    def run_dev(K):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            pipe.submit()
        pipe.join()
        b.record(); b.synchronize()
        return a.elapsed_time(b)
    run_dev(20)
    print(f"late={late} device-resident K=200: {min(run_dev(200) for _ in range(3)) / 200:.3f} ms/step")
    run(20)
    for K in (1, 20, 200):
        ms = min(run(K) for _ in range(3))
        print(f"late={late} lanes={L} K={K}: {ms:.2f} ms total, {ms / K:.3f} ms/step, {8 * K / ms * 1e3:.0f} scenes/s")
    res[late] = [[t.clone() for t in o] for o in outs]
    del pipe
same = all(torch.equal(a, b) for oa, ob in zip(res[False], res[True]) for a, b in zip(oa, ob))
print("outputs identical:", same)

if os.environ.get("TRACE"):
    from torch.profiler import ProfilerActivity, profile
    pipe = engine.ForwardPipeline(model, sets, lanes=L, late_images=True)
    outs = [[torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory() for t in pipe.slots[0].outputs] for _ in range(S)]
    for i in range(20):
        hb = host[i % S]; pipe.submit(hb["points"], hb["img"], hb["img_metas"], outputs_to=outs[i % S])
    pipe.join(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(20):
            hb = host[i % S]; pipe.submit(hb["points"], hb["img"], hb["img_metas"], outputs_to=outs[i % S])
        pipe.join(); torch.cuda.synchronize()
    evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e[0] for e in evs)
    big = [(s - t0, e - t0, n) for s, e, n in evs if "Memcpy HtoD" in n and e - s > 100]
    big.sort()
    print("big H2D copies:", len(big))
    for s, e, n in big[:24]:
        print(f"  {s / 1e3:8.2f} ms -> {e / 1e3:8.2f} ms  ({(e - s) / 1e3:.2f} ms)")
    ks = sorted((s - t0, e - t0, n) for s, e, n in evs if "Memcpy" not in n and "Memset" not in n)
    print("first kernel at", ks[0][0] / 1e3, "last end", max(k[1] for k in ks) / 1e3)
