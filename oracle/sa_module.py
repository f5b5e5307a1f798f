"""CPU restatement of one PointSAModule forward in eval mode -- TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the
product path).

Follows upstream mmdet3d 0.18.1 `ops/pointnet_modules/point_sa_module.py` (BasePointSAModule
.forward: QueryAndGroup -> mlps[i] -> F.max_pool2d(kernel=[1, ns])) as the reference reaches it
from configs/demf/demf_votenet.py:48-62,155-162 and class_agnostic_vote_head.py:455, with the
Conv2d(1x1, bias=False)+BN2d(eval)+ReLU layers given as already-folded (W, b) pairs in
UPSTREAM channel order [xyz(3), feat(C)].

`tf32=True` rounds every GEMM operand to TF32 (nearest, ties away from zero = cvt.rna) before a
float64 product: the arithmetic class of the tensor-core kernel (csrc/sa_fused.cu) and of the
cuDNN TF32 convolutions PyTorch runs by default for the reference on Ampere and later.
"""
import numpy as np
import torch

from . import cref


def tf32_rna(x):
    """float32 tensor -> nearest TF32 value (10-bit mantissa), ties away from zero."""
    a = np.ascontiguousarray(x.detach().cpu().numpy().astype(np.float32))
    bits = a.view(np.uint32)
    out = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return torch.from_numpy(out.copy())


def sa_forward(xyz, new_xyz, features, min_radius, max_radius, ns, normalize_xyz, weights,
               biases, tf32=False):
    """xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or None; weights[i] (Cout,Cin) with
    weights[0] columns ordered [xyz(3), feat(C)]; biases[i] (Cout,) -> (idx (B,M,ns) i32,
    out (B,C3,M))."""
    idx, grouped = cref.query_and_group(xyz, new_xyz, features, min_radius, max_radius, ns,
                                        use_xyz=True, normalize_xyz=normalize_xyz)
    B, K, M, S = grouped.shape
    x = grouped.permute(0, 2, 3, 1).reshape(B * M * S, K)          # rows, upstream channel order
    for w, b in zip(weights, biases):
        w, b = w.detach().cpu().float(), b.detach().cpu().float()
        if tf32:
            y = (tf32_rna(x).double() @ tf32_rna(w).double().t()).float() + b
        else:
            y = x @ w.t() + b
        x = torch.relu(y)
    out = x.view(B, M, S, -1).amax(dim=2)                          # (B,M,C3)
    return idx, out.transpose(1, 2).contiguous()
