"""Small inference GEMMs inside a CUDA graph: the library (torch.addmm / _addmm_activation, TF32) against our tcgen05
rows kernel (P.gemm_rows_fwd), 24 dependent launches per replay, time per launch."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200.mm import point_ops as P
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
def graph_time(fn, n=24, reps=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(n): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / n)
    return statistics.median(ts)
for R, K, N, relu in [(128, 256, 256, False), (128, 64, 64, False), (2048, 64, 64, False), (65536, 64, 64, False), (524288, 64, 64, False), (2048, 256, 512, False), (2048, 256, 256, True), (2048, 256, 1024, True), (2048, 1024, 256, False),
                      (8192, 256, 256, True), (4096, 512, 256, True), (8192, 256, 260, False), (43520, 256, 256, False),
                      (2048, 128, 12, False), (2048, 8, 256, True)]:
    x = torch.randn(R, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(R, N, device=dev)
    lib = graph_time((lambda: torch._addmm_activation(b, x, w.t(), out=y)) if relu else (lambda: torch.addmm(b, x, w.t(), out=y)))
    try:
        ours = graph_time(lambda: P.gemm_rows_fwd(x, w, bias=b, relu=relu, out=y))
    except Exception as e:
        ours = float("nan")
    print(f"R={R} K={K} N={N} relu={relu}: library {lib:.2f} us  ours {ours:.2f} us")
