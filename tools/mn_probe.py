"""Which MN-major operand encoding does the tensor core take? (development probe for csrc/gemm_tf32.cu)"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demf_b200 import _lib
from demf_b200.mm import point_ops as P
dev = torch.device("cuda:0"); lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
R, K, N = 1024, 64, 128
dy = torch.randn(R, N, generator=g, device=dev); w = torch.randn(N, K, generator=g, device=dev) / N ** 0.5
x = torch.randn(R, K, generator=g, device=dev)
want_dx = dy.double() @ w.double(); want_dw = dy.double().t() @ x.double()
for tma, layout, sbo in itertools.product((4, 3), (1, 2), (512, 1024, 256)):
    lib.demf_gemm_debug_mn(sbo, layout, tma)
    dx = P.gemm_rows_dgrad(dy, w); dw = torch.zeros(N, K, device=dev); P.gemm_wgrad_(dw, dy, x)
    torch.cuda.synchronize()
    e1 = (dx.double() - want_dx).abs().max().item() / want_dx.abs().max().item()
    e2 = (dw.double() - want_dw).abs().max().item() / want_dw.abs().max().item()
    print(f"tma_swizzle={tma} layout={layout} sbo={sbo}: dgrad rel err {e1:.4f}  wgrad rel err {e2:.4f}  gemm_error={P.gemm_error()}", flush=True)
lib.demf_gemm_debug_mn(512, 1, 4)
