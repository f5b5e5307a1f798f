"""PointNet++ ops with the mmdet3d.ops names and signatures, backed by libdemf_b200.so.

Stands in for `from mmdet3d.ops import build_sa_module, furthest_point_sample`
(reference: demf/modeling/heads/class_agnostic_vote_head.py:13) and the ops those modules use
internally (mmdet3d 0.18.1 ops/{furthest_point_sample, ball_query, group_points,
gather_points, interpolate}). Semantics follow upstream: float32 / int32 contiguous CUDA
tensors, `assert`-style contiguity checks (no silent .contiguous()), index outputs
non-differentiable. CUDA only -- there is deliberately no CPU implementation.
"""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "demf_b200 ops run on CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor")


def _p(t):
    return None if t is None else t.data_ptr()


class FurthestPointSampling(Function):
    """xyz (B,N,3) f32 -> idx (B,m) i32, idx[:,0]=0 (upstream furthest_point_sample)."""

    @staticmethod
    def forward(ctx, points_xyz, num_points):
        assert points_xyz.is_contiguous()
        _need_cuda(points_xyz)
        assert points_xyz.dtype == torch.float32 and points_xyz.dim() == 3 and points_xyz.size(2) == 3
        B, N = points_xyz.shape[:2]
        lib = _lib.load()
        with torch.cuda.device_of(points_xyz):
            idx = torch.empty(B, num_points, dtype=torch.int32, device=points_xyz.device)
            ws_bytes = lib.demf_fps_workspace_bytes(B, N, num_points)
            ws = (torch.empty(ws_bytes // 4, dtype=torch.float32, device=points_xyz.device)
                  if ws_bytes else None)
            if idx.numel():
                _lib.check(lib.demf_fps(_p(points_xyz), B, N, num_points, _p(ws), _p(idx), None,
                                        _stream()), "demf_fps")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


def furthest_point_sample_xyz(points_xyz, num_points, grid=None, unique_prefix=None, return_prefix=False,
                              certify=None):
    """furthest_point_sample that also returns the picked points' coordinates (B,m,3): the kernel
    has them in registers when it writes an index, so the separate gather launch (and the int64
    index copy in front of it) disappears. `grid` = ball_grid workspace: grid-pruned kernel.
    Sampling is not differentiable; use gather on the indices when gradients must flow to xyz.

    The sampling chain: with `return_prefix` (grid kernel) a third result (B,) i32 certifies, per scene, how many
    leading picks were the unique arg-max of their iteration; handed back as `unique_prefix` when the cloud to
    sample IS that pick sequence, scenes whose certificate covers `num_points` get idx = 0..m-1 without
    iterating (identical to the ordinary result; see include/demf_b200.h). `certify`: how many leading picks the
    caller will ask about (default all): only those iterations pay for the bookkeeping."""
    assert points_xyz.is_contiguous()
    _need_cuda(points_xyz, grid, unique_prefix)
    B, N = points_xyz.shape[:2]
    lib = _lib.load()
    prefix = None
    with torch.cuda.device_of(points_xyz):
        idx = torch.empty(B, num_points, dtype=torch.int32, device=points_xyz.device)
        new_xyz = torch.empty(B, num_points, 3, dtype=torch.float32, device=points_xyz.device)
        if return_prefix:
            prefix = torch.zeros(B, dtype=torch.int32, device=points_xyz.device)
        if idx.numel():
            if grid is not None:
                _lib.check(lib.demf_fps_grid_prefix(
                    _p(points_xyz), _p(grid), B, N, int(num_points), _p(idx), _p(new_xyz), _p(prefix),
                    int(num_points if certify is None else min(certify, num_points)), _stream()), "demf_fps_grid")
            else:
                ws_bytes = lib.demf_fps_workspace_bytes(B, N, num_points)
                ws = (torch.empty(ws_bytes // 4, dtype=torch.float32, device=points_xyz.device)
                      if ws_bytes else None)
                if unique_prefix is not None:
                    assert unique_prefix.dtype == torch.int32 and unique_prefix.numel() == B
                _lib.check(lib.demf_fps_prefix(_p(points_xyz), B, N, int(num_points), _p(ws), _p(idx),
                                               _p(new_xyz), _p(unique_prefix), _stream()), "demf_fps")
    if return_prefix:
        return idx, new_xyz, prefix
    return idx, new_xyz


class BallQuery(Function):
    """First `sample_num` points (index order) with d2==0 or min_r^2 <= d2 < max_r^2."""

    @staticmethod
    def forward(ctx, min_radius, max_radius, sample_num, xyz, center_xyz):
        assert center_xyz.is_contiguous()
        assert xyz.is_contiguous()
        assert min_radius < max_radius
        _need_cuda(xyz, center_xyz)
        B, N, _ = xyz.shape
        M = center_xyz.size(1)
        with torch.cuda.device_of(xyz):
            idx = torch.empty(B, M, sample_num, dtype=torch.int32, device=xyz.device)
            if idx.numel():
                _lib.check(_lib.load().demf_ball_query(_p(xyz), _p(center_xyz), B, N, M, min_radius,
                                                       max_radius, sample_num, _p(idx), _stream()),
                           "demf_ball_query")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None, None


ball_query = BallQuery.apply


class GroupingOperation(Function):
    """features (B,C,N), indices (B,M,ns) -> (B,C,M,ns)."""

    @staticmethod
    def forward(ctx, features, indices):
        assert features.is_contiguous()
        assert indices.is_contiguous()
        _need_cuda(features, indices)
        B, C, N = features.shape
        _, M, ns = indices.shape
        with torch.cuda.device_of(features):
            out = torch.empty(B, C, M, ns, dtype=features.dtype, device=features.device)
            if out.numel():
                _lib.check(_lib.load().demf_group_fwd(_p(features), _p(indices), B, C, N, M, ns,
                                                      _p(out), _stream()), "demf_group_fwd")
        ctx.for_backwards = (indices, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indices, N = ctx.for_backwards
        B, C, M, ns = grad_out.shape
        grad_out = grad_out.contiguous()
        with torch.cuda.device_of(grad_out):
            grad = torch.zeros(B, C, N, dtype=grad_out.dtype, device=grad_out.device)
            if grad_out.numel():
                _lib.check(_lib.load().demf_group_bwd(_p(grad_out), _p(indices), B, C, N, M, ns,
                                                      _p(grad), _stream()), "demf_group_bwd")
        return grad, None


grouping_operation = GroupingOperation.apply


class GatherPoints(Function):
    """features (B,C,N), indices (B,M) -> (B,C,M)."""

    @staticmethod
    def forward(ctx, features, indices):
        assert features.is_contiguous()
        assert indices.is_contiguous()
        _need_cuda(features, indices)
        B, C, N = features.shape
        M = indices.size(1)
        with torch.cuda.device_of(features):
            out = torch.empty(B, C, M, dtype=features.dtype, device=features.device)
            if out.numel():
                _lib.check(_lib.load().demf_gather_fwd(_p(features), _p(indices), B, C, N, M, _p(out),
                                                       _stream()), "demf_gather_fwd")
        ctx.for_backwards = (indices, C, N)
        ctx.mark_non_differentiable(indices)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indices, C, N = ctx.for_backwards
        B, M = indices.shape
        grad_out = grad_out.contiguous()
        with torch.cuda.device_of(grad_out):
            grad = torch.zeros(B, C, N, dtype=grad_out.dtype, device=grad_out.device)
            if grad_out.numel():
                _lib.check(_lib.load().demf_gather_bwd(_p(grad_out), _p(indices), B, C, N, M,
                                                       _p(grad), _stream()), "demf_gather_bwd")
        return grad, None


gather_points = GatherPoints.apply


def three_nn_squared(target, source):
    """Raw kernel result: SQUARED distances (B,n,3) f32 and indices (B,n,3) i32."""
    assert target.is_contiguous()
    assert source.is_contiguous()
    _need_cuda(target, source)
    B, n, _ = target.shape
    m = source.size(1)
    with torch.cuda.device_of(target):
        dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=target.device)
        idx = torch.empty(B, n, 3, dtype=torch.int32, device=target.device)
        if idx.numel():
            _lib.check(_lib.load().demf_three_nn(_p(target), _p(source), B, n, m, _p(dist2),
                                                 _p(idx), _stream()), "demf_three_nn")
    return dist2, idx


class ThreeNN(Function):
    """target (B,n,3), source (B,m,3) -> (dist (B,n,3) = sqrt(d2), idx (B,n,3) i32).
    The square root is torch.sqrt on the device, as in the upstream wrapper."""

    @staticmethod
    def forward(ctx, target, source):
        dist2, idx = three_nn_squared(target, source)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,C,m), indices (B,n,3), weight (B,n,3) -> (B,C,n)."""

    @staticmethod
    def forward(ctx, features, indices, weight):
        assert features.is_contiguous()
        assert indices.is_contiguous()
        assert weight.is_contiguous()
        _need_cuda(features, indices, weight)
        B, C, m = features.shape
        n = indices.size(1)
        ctx.three_interpolate_for_backward = (indices, weight, m)
        with torch.cuda.device_of(features):
            out = torch.empty(B, C, n, dtype=features.dtype, device=features.device)
            if out.numel():
                _lib.check(_lib.load().demf_three_interpolate_fwd(
                    _p(features), _p(indices), _p(weight), B, C, m, n, _p(out), _stream()),
                    "demf_three_interpolate_fwd")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indices, weight, m = ctx.three_interpolate_for_backward
        B, C, n = grad_out.shape
        grad_out = grad_out.contiguous()
        with torch.cuda.device_of(grad_out):
            grad = torch.zeros(B, C, m, dtype=grad_out.dtype, device=grad_out.device)
            if grad_out.numel():
                _lib.check(_lib.load().demf_three_interpolate_bwd(
                    _p(grad_out), _p(indices), _p(weight), B, C, n, m, _p(grad), _stream()),
                    "demf_three_interpolate_bwd")
        return grad, None, None


three_interpolate = ThreeInterpolate.apply


class _FusedQueryAndGroup(Function):
    """ball query + group + centre subtraction + radius normalisation + concat, one launch.

    Returns (idx, grouped) with grouped = cat([ (xyz[idx]-centre)/r , features[idx] ], dim=1).
    Backward reuses the grouping scatter: d/dfeatures, d/dxyz (vote aggregation needs it: the
    grouped coordinates depend on the learned vote offsets) and d/dcentre.
    """

    @staticmethod
    def forward(ctx, xyz, center_xyz, features, min_radius, max_radius, sample_num, use_xyz,
                normalize_xyz):
        assert xyz.is_contiguous()
        assert center_xyz.is_contiguous()
        assert features is None or features.is_contiguous()
        _need_cuda(xyz, center_xyz, features)
        B, N, _ = xyz.shape
        M = center_xyz.size(1)
        C = 0 if features is None else features.size(1)
        Cx = 3 if use_xyz else 0
        with torch.cuda.device_of(xyz):
            idx = torch.empty(B, M, sample_num, dtype=torch.int32, device=xyz.device)
            out = torch.empty(B, Cx + C, M, sample_num, dtype=torch.float32, device=xyz.device)
            if idx.numel():
                _lib.check(_lib.load().demf_query_and_group_fwd(
                    _p(xyz), _p(features), _p(center_xyz), B, N, M, C, min_radius, max_radius,
                    sample_num, int(use_xyz), int(normalize_xyz), _p(idx), _p(out), _stream()),
                    "demf_query_and_group_fwd")
        ctx.saved = (idx, N, C, Cx, (1.0 / max_radius) if normalize_xyz else 1.0)
        ctx.mark_non_differentiable(idx)
        return idx, out

    @staticmethod
    def backward(ctx, _grad_idx, grad_out):
        idx, N, C, Cx, scale = ctx.saved
        need_xyz, need_center, need_feat = ctx.needs_input_grad[:3]
        g_xyz = g_center = g_feat = None
        if Cx and (need_xyz or need_center):
            gx = grad_out[:, :Cx] * scale  # (B,3,M,ns)
            if need_center:
                g_center = -gx.sum(-1).transpose(1, 2).contiguous()
            if need_xyz:
                g_xyz = GroupingOperation.backward(_Ctx(idx, N), gx.contiguous())[0]
                g_xyz = g_xyz.transpose(1, 2).contiguous()
        if C and need_feat:
            g_feat = GroupingOperation.backward(_Ctx(idx, N), grad_out[:, Cx:].contiguous())[0]
        return g_xyz, g_center, g_feat, None, None, None, None, None


class _Ctx:
    def __init__(self, indices, N):
        self.for_backwards = (indices, N)


class QueryAndGroup(torch.nn.Module):
    """mmdet3d.ops.QueryAndGroup (ops/group_points/group_points.py), same ctor and forward.

    The common configuration (use_xyz, features given, no extra return values) runs as ONE
    fused kernel; the rarely used options fall back to the individual ops (still CUDA).
    """

    def __init__(self, max_radius, sample_num, min_radius=0, use_xyz=True,
                 return_grouped_xyz=False, normalize_xyz=False, uniform_sample=False,
                 return_unique_cnt=False, return_grouped_idx=False):
        super().__init__()
        self.max_radius = max_radius
        self.min_radius = min_radius
        self.sample_num = sample_num
        self.use_xyz = use_xyz
        self.return_grouped_xyz = return_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.uniform_sample = uniform_sample
        self.return_unique_cnt = return_unique_cnt
        self.return_grouped_idx = return_grouped_idx
        if self.return_unique_cnt:
            assert self.uniform_sample, \
                'uniform_sample should be True when returning the count of unique samples'
        if self.max_radius is None:
            raise NotImplementedError("kNN grouping (max_radius=None) is not on the DeMF path")
        if self.uniform_sample:
            raise NotImplementedError("uniform_sample is not on the DeMF path")

    def forward(self, points_xyz, center_xyz, features=None):
        fused_ok = not (self.uniform_sample or self.return_grouped_xyz or self.return_unique_cnt
                        or self.return_grouped_idx) and (self.use_xyz or features is not None)
        if fused_ok:
            _, new_features = _FusedQueryAndGroup.apply(
                points_xyz, center_xyz, features,
                float(self.min_radius), float(self.max_radius), self.sample_num,
                bool(self.use_xyz), bool(self.normalize_xyz))
            return new_features
        return self._forward_unfused(points_xyz, center_xyz, features)

    def _forward_unfused(self, points_xyz, center_xyz, features):
        idx = ball_query(self.min_radius, self.max_radius, self.sample_num, points_xyz, center_xyz)
        xyz_trans = points_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz_diff = grouped_xyz - center_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz_diff = grouped_xyz_diff / self.max_radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = (torch.cat([grouped_xyz_diff, grouped_features], dim=1)
                            if self.use_xyz else grouped_features)
        else:
            assert self.use_xyz, 'Cannot have not features and not use xyz as a feature!'
            new_features = grouped_xyz_diff
        ret = [new_features]
        if self.return_grouped_xyz:
            ret.append(grouped_xyz)
        if self.return_grouped_idx:
            ret.append(idx)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(torch.nn.Module):
    """mmdet3d.ops.GroupAll: one group holding every point (SA module with num_point=None)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz \
                else grouped_features
        return grouped_xyz


# ----------------------------------------------------------------- point-major rows --
def group_rows_width(C):
    """Row width K of the grouped tensor for C feature channels: roundup(C,4)+4."""
    return ((C + 3) // 4) * 4 + 4


def group_rows_columns(C):
    """Column index, for each of the K row slots, into upstream's [xyz(3), feat(C)] channel
    order; -1 marks a zero-pad slot. Used to permute 1x1-conv weights to the row layout."""
    Cp = ((C + 3) // 4) * 4
    cols = [3 + c for c in range(C)] + [-1] * (Cp - C) + [0, 1, 2, -1]
    return cols


class QueryAndGroupRows(Function):
    """Ball query + grouping as GEMM-ready rows (csrc/rows.cu), one launch.

    xyz (B,N,3), center_xyz (B,M,3), feat_rows (B,N,C) or None ->
      idx (B,M,ns) i32, rows (B,M,ns,K) with K = group_rows_width(C):
      rows[b,m,s] = [feat_rows[b,idx] | 0.. | (xyz[idx]-centre)(/max_radius) | 0]
    Same neighbour rows as upstream QueryAndGroup (use_xyz=True), xyz columns last.
    """

    @staticmethod
    def forward(ctx, xyz, center_xyz, feat_rows, min_radius, max_radius, sample_num,
                normalize_xyz, grid=None):
        assert xyz.is_contiguous()
        assert center_xyz.is_contiguous()
        assert feat_rows is None or feat_rows.is_contiguous()
        _need_cuda(xyz, center_xyz, feat_rows)
        B, N, _ = xyz.shape
        M = center_xyz.size(1)
        C = 0 if feat_rows is None else feat_rows.size(2)
        K = group_rows_width(C)
        with torch.cuda.device_of(xyz):
            idx = torch.empty(B, M, sample_num, dtype=torch.int32, device=xyz.device)
            out = torch.empty(B, M, sample_num, K, dtype=torch.float32, device=xyz.device)
            if idx.numel():
                _lib.check(_lib.load().demf_query_and_group_rows_fwd(
                    _p(xyz), _p(feat_rows), _p(center_xyz), B, N, M, C, min_radius, max_radius,
                    sample_num, int(normalize_xyz), 1, _p(grid), _p(idx), _p(out), _stream()),
                    "demf_query_and_group_rows_fwd")
        ctx.saved = (idx, N, C, (1.0 / max_radius) if normalize_xyz else 1.0)
        ctx.mark_non_differentiable(idx)
        return idx, out

    @staticmethod
    def backward(ctx, _grad_idx, grad_out):
        idx, N, C, scale = ctx.saved
        need_xyz, need_center, need_feat = ctx.needs_input_grad[:3]
        B, M, ns, _ = grad_out.shape
        grad_out = grad_out.contiguous()
        dev = grad_out.device
        with torch.cuda.device_of(grad_out):
            g_feat = torch.zeros(B, N, C, dtype=torch.float32, device=dev) if (C and need_feat) else None
            g_xyz = torch.zeros(B, N, 3, dtype=torch.float32, device=dev) if need_xyz else None
            g_center = torch.empty(B, M, 3, dtype=torch.float32, device=dev) if need_center else None
            if grad_out.numel() and (g_feat is not None or g_xyz is not None or g_center is not None):
                _lib.check(_lib.load().demf_group_rows_bwd(
                    _p(grad_out), _p(idx), B, N, M, C, ns, scale,
                    _p(g_feat), _p(g_xyz), _p(g_center), _stream()), "demf_group_rows_bwd")
        return g_xyz, g_center, g_feat, None, None, None, None, None


def query_and_group_rows(xyz, center_xyz, feat_rows, min_radius, max_radius, sample_num,
                         normalize_xyz, grid=None):
    """`grid` = ball_grid(xyz, r >= max_radius): same rows, but each centre only tests its 3x3x3
    cell neighbourhood instead of the whole cloud."""
    return QueryAndGroupRows.apply(xyz, center_xyz, feat_rows, float(min_radius), float(max_radius),
                                   int(sample_num), bool(normalize_xyz), grid)


def ball_grid(xyz, radius):
    """Bin xyz (B,N,3) into a uniform grid with cell edge >= radius (csrc/ball_grid.cu): the opaque
    workspace tensor accepted as `grid=` by ball_query_grid / query_and_group_rows / sa_fused for
    queries on the SAME xyz with max_radius <= radius."""
    assert xyz.is_contiguous()
    _need_cuda(xyz)
    B, N, _ = xyz.shape
    lib = _lib.load()
    with torch.cuda.device_of(xyz):
        ws = torch.empty(int(lib.demf_ball_grid_workspace_bytes(B, N)), dtype=torch.uint8,
                         device=xyz.device)
        _lib.check(lib.demf_ball_grid_build(_p(xyz), B, N, float(radius), _p(ws), _stream()),
                   "demf_ball_grid_build")
    return ws


def furthest_point_sample_grid(xyz, num_points, grid):
    """furthest_point_sample through a ball_grid workspace of the same xyz (csrc/fps.cu,
    fps_grid_kernel): the cloud sits cell-ordered in the shared memory of a small cluster and a new
    sample only revisits the 32-point blocks whose bounding box it can reach. Identical indices."""
    assert xyz.is_contiguous()
    _need_cuda(xyz, grid)
    B, N, _ = xyz.shape
    with torch.cuda.device_of(xyz):
        idx = torch.empty(B, num_points, dtype=torch.int32, device=xyz.device)
        if idx.numel():
            _lib.check(_lib.load().demf_fps_grid(_p(xyz), _p(grid), B, N, int(num_points), _p(idx),
                                                 None, _stream()), "demf_fps_grid")
    return idx


def ball_query_grid(min_radius, max_radius, sample_num, xyz, center_xyz, grid):
    """ball_query through a ball_grid workspace: bit-identical (B,M,ns) int32 rows."""
    assert xyz.is_contiguous() and center_xyz.is_contiguous()
    _need_cuda(xyz, center_xyz, grid)
    B, N, _ = xyz.shape
    M = center_xyz.size(1)
    with torch.cuda.device_of(xyz):
        idx = torch.zeros(B, M, sample_num, dtype=torch.int32, device=xyz.device)
        if idx.numel():
            _lib.check(_lib.load().demf_ball_query_grid(
                _p(xyz), _p(center_xyz), _p(grid), B, N, M, float(min_radius), float(max_radius),
                int(sample_num), _p(idx), _stream()), "demf_ball_query_grid")
    return idx


class ThreeInterpolateRows(Function):
    """feat_rows (B,m,C), indices (B,n,3), weight (B,n,3) -> (B,n,C); upstream's fma order."""

    @staticmethod
    def forward(ctx, feat_rows, indices, weight):
        assert feat_rows.is_contiguous()
        assert indices.is_contiguous()
        assert weight.is_contiguous()
        _need_cuda(feat_rows, indices, weight)
        B, m, C = feat_rows.shape
        n = indices.size(1)
        ctx.saved = (indices, weight, m)
        with torch.cuda.device_of(feat_rows):
            out = torch.empty(B, n, C, dtype=feat_rows.dtype, device=feat_rows.device)
            if out.numel():
                _lib.check(_lib.load().demf_three_interpolate_rows_fwd(
                    _p(feat_rows), _p(indices), _p(weight), B, C, m, n, _p(out), _stream()),
                    "demf_three_interpolate_rows_fwd")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indices, weight, m = ctx.saved
        B, n, C = grad_out.shape
        grad_out = grad_out.contiguous()
        with torch.cuda.device_of(grad_out):
            grad = torch.zeros(B, m, C, dtype=grad_out.dtype, device=grad_out.device)
            if grad_out.numel():
                _lib.check(_lib.load().demf_three_interpolate_rows_bwd(
                    _p(grad_out), _p(indices), _p(weight), B, C, n, m, _p(grad), _stream()),
                    "demf_three_interpolate_rows_bwd")
        return grad, None, None


three_interpolate_rows = ThreeInterpolateRows.apply


def gather_rows(rows, indices):
    """rows (B,N,C), indices (B,M) i32 -> (B,M,C). Tiny (centre coordinates); torch.gather."""
    return torch.gather(rows, 1, indices.long().unsqueeze(-1).expand(-1, -1, rows.size(-1)))


# ------------------------------------------------- fused set abstraction (inference) --
def sa_fused_supported(C, sample_num, widths):
    """True when csrc/sa_fused.cu covers this (feature channels, nsample, 3 MLP widths)."""
    return len(widths) == 3 and bool(_lib.load().demf_sa_fused_supported(
        int(C), int(sample_num), int(widths[0]), int(widths[1]), int(widths[2])))


def sa_pack_mlp(weights, biases):
    """Three BN-folded (Cout,Cin) weights (first one with its columns in row order,
    group_rows_columns) + biases -> (wpack, bias, widths) for `sa_fused`: each matrix in the
    swizzled K-major chunk image the kernel bulk-copies into shared memory, TF32-rounded."""
    assert len(weights) == 3 and len(biases) == 3
    _need_cuda(*weights)
    lib = _lib.load()
    dev = weights[0].device
    sizes = [int(lib.demf_sa_pack_floats(w.size(0), w.size(1))) for w in weights]
    with torch.cuda.device_of(weights[0]):
        wpack = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        for w, n in zip(weights, sizes):
            w = w.detach().float().contiguous()
            _lib.check(lib.demf_sa_pack_weights(_p(w), w.size(0), w.size(1),
                                                wpack.data_ptr() + 4 * off, _stream()),
                       "demf_sa_pack_weights")
            off += n
    bias = torch.cat([b.detach().float().flatten() for b in biases]).contiguous()
    return wpack, bias, tuple(int(w.size(0)) for w in weights)


def sa_fused(xyz, center_xyz, feat_rows, min_radius, max_radius, sample_num, normalize_xyz, wpack,
             bias, widths, idx=None, return_idx=False, grid=None):
    """Ball query + grouping + 3-layer MLP (bias, ReLU) + max over the neighbourhood in one launch
    (csrc/sa_fused.cu): xyz (B,N,3), center_xyz (B,M,3), feat_rows (B,N,C) or None -> (B,M,c3)
    rows. `idx` (B,M,ns) i32 given: grouping uses it instead of running the ball query.
    Inference only (no autograd)."""
    assert xyz.is_contiguous() and center_xyz.is_contiguous()
    assert feat_rows is None or feat_rows.is_contiguous()
    _need_cuda(xyz, center_xyz, feat_rows, wpack, bias)
    B, N, _ = xyz.shape
    M = center_xyz.size(1)
    C = 0 if feat_rows is None else feat_rows.size(2)
    query = idx is None
    with torch.cuda.device_of(xyz):
        out = torch.empty(B, M, widths[2], dtype=torch.float32, device=xyz.device)
        if query and (return_idx or grid is not None):
            idx = torch.empty(B, M, sample_num, dtype=torch.int32, device=xyz.device)
        elif not query:
            assert idx.is_contiguous() and idx.dtype == torch.int32
        if out.numel():
            _lib.check(_lib.load().demf_sa_fused_fwd(
                _p(xyz), _p(feat_rows), _p(center_xyz), B, N, M, C, float(min_radius),
                float(max_radius), int(sample_num), int(bool(normalize_xyz)), int(query), _p(wpack),
                _p(bias), widths[0], widths[1], widths[2], _p(grid), _p(idx), _p(out), _stream()),
                "demf_sa_fused_fwd")
    return (out, idx) if return_idx else out


# --------------------------------------------------- fused glue kernels (inference) --
def chain_indices(level_indices):
    """[idx_0 (B,M0) i32, idx_1 (B,M1) i32, ...] (each indexing the previous level's points) ->
    [(B,M_l) i64 indices into the ORIGINAL cloud], one launch (csrc/glue.cu). Equals upstream's
    `sa_indices[i+1] = torch.gather(sa_indices[i], 1, idx.long())` chain started from arange."""
    assert 1 <= len(level_indices) <= 4
    _need_cuda(*level_indices)
    B = level_indices[0].size(0)
    outs = [torch.empty(t.shape, dtype=torch.int64, device=t.device) for t in level_indices]
    pad = 4 - len(level_indices)
    ins = [t.contiguous() for t in level_indices]
    args = [B, len(ins)]
    for t in ins:
        args += [_p(t), t.size(1)]
    args += [None, 0] * pad
    args += [_p(o) for o in outs] + [None] * pad
    with torch.cuda.device_of(ins[0]):
        _lib.check(_lib.load().demf_chain_indices(*args, _stream()), "demf_chain_indices")
    return outs


def interp_cat_rows(src_rows, skip_rows, indices, dist2):
    """three_nn squared distances + indices -> inverse-distance weights -> interpolation of
    src_rows (B,m,C1) -> concatenated with skip_rows (B,n,C2) (or None): (B,n,C1+C2), one launch
    (csrc/glue.cu; PointFPModule's sqrt/reciprocal/sum/div/three_interpolate/cat). Inference only."""
    assert src_rows.is_contiguous() and indices.is_contiguous() and dist2.is_contiguous()
    assert skip_rows is None or skip_rows.is_contiguous()
    _need_cuda(src_rows, skip_rows, indices, dist2)
    B, m, C1 = src_rows.shape
    n = indices.size(1)
    C2 = 0 if skip_rows is None else skip_rows.size(2)
    with torch.cuda.device_of(src_rows):
        out = torch.empty(B, n, C1 + C2, dtype=torch.float32, device=src_rows.device)
        if out.numel():
            _lib.check(_lib.load().demf_interp_cat_rows_fwd(
                _p(src_rows), _p(skip_rows), _p(indices), _p(dist2), B, C1, C2, m, n, _p(out),
                _stream()), "demf_interp_cat_rows_fwd")
    return out


def decode_boxes(res, num_dir_bins, box, obj_prob, sem_prob, row_offset):
    """One prediction stage `res` (split_pred dict of (B,Q,*) row tensors, views allowed) ->
    box (7), objectness probability and semantic probabilities written at rows
    [row_offset, row_offset+Q) of the (B,R,7)/(B,R)/(B,R,classes) outputs, one launch."""
    names = ("center", "size", "dir_class", "dir_res", "obj_scores", "sem_scores")
    ts = [res[k] for k in names]
    _need_cuda(*ts)
    B, Q = ts[0].shape[:2]
    for t in ts:  # row tensors: (B,Q,c) with unit channel stride and b-stride = Q * q-stride
        assert t.stride(2) == 1 and t.stride(0) == Q * t.stride(1), t.stride()
    args = []
    for t in ts:
        args += [_p(t), t.stride(1)]
    with torch.cuda.device_of(box):
        _lib.check(_lib.load().demf_decode_boxes(
            *args, B, Q, int(num_dir_bins), ts[5].size(2), box.size(1), int(row_offset), _p(box),
            _p(obj_prob), _p(sem_prob), _stream()), "demf_decode_boxes")


def bias_layer_norm_rows(x, gamma, beta, eps, bias=None, residual=None, out=None, post_add=None):
    """LayerNorm(x + bias + residual) over the last axis of contiguous rows x (R, C) in one pass
    (csrc/glue.cu); C a multiple of 128 up to 1024. `out` may be x itself. With `post_add` (R, C)
    also returns out + post_add (the next attention's query + positional embedding)."""
    _need_cuda(x, gamma, beta)
    R, C = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    assert residual is None or (residual.is_contiguous() and residual.shape == x.shape)
    assert post_add is None or (post_add.is_contiguous() and post_add.shape == x.shape)
    if out is None:
        out = torch.empty_like(x)
    out2 = torch.empty_like(x) if post_add is not None else None
    with torch.cuda.device_of(x):
        _lib.check(_lib.load().demf_bias_layer_norm_rows(
            _p(x), _p(bias) if bias is not None else None, _p(residual) if residual is not None else None,
            _p(gamma), _p(beta), R, C, float(eps), _p(out), _p(post_add) if post_add is not None else None,
            _p(out2) if out2 is not None else None, _stream()), "demf_bias_layer_norm_rows")
    return out if out2 is None else (out, out2)


def box_point_count(points, boxes, gravity_centre=False):
    """points (B,N,3+) rows, boxes (B,K,7) bottom-centre (or gravity-centre) -> (B,K) int32 number of
    points inside each box (column sums of mmdet3d `points_in_boxes`), one launch for the batch."""
    _need_cuda(points, boxes)
    B, N = points.shape[:2]
    K = boxes.shape[1]
    assert points.dtype == torch.float32
    assert N == 0 or (points.stride(2) == 1 and points.stride(0) == N * points.stride(1)), points.stride()
    boxes = boxes.contiguous().float()
    counts = torch.zeros(B, K, dtype=torch.int32, device=points.device)
    if B == 0 or K == 0 or N == 0:
        return counts
    with torch.cuda.device_of(points):
        _lib.check(_lib.load().demf_box_point_count(_p(points), points.stride(1), _p(boxes), B, N, K,
                                                    int(bool(gravity_centre)), _p(counts), _stream()),
                   "demf_box_point_count")
    return counts


def aligned_3d_nms(minmax, scores, classes, valid, thresh):
    """Batched mmdet3d `aligned_3d_nms`: minmax (B,K,6), scores (B,K), classes (B,K) int64, valid (B,K)
    bool -> (B,K) bool mask of the boxes the greedy class-aware sweep keeps (one CTA per scene)."""
    _need_cuda(minmax, scores, classes, valid)
    B, K = scores.shape
    minmax, scores = minmax.contiguous().float(), scores.contiguous().float()
    classes = classes.contiguous().long()
    valid8 = valid.contiguous().to(torch.uint8)
    keep = torch.zeros(B, K, dtype=torch.uint8, device=scores.device)
    if B == 0 or K == 0:
        return keep.bool()
    with torch.cuda.device_of(scores):
        _lib.check(_lib.load().demf_aligned_3d_nms(_p(minmax), _p(scores), _p(classes), _p(valid8), B, K,
                                                   float(thresh), _p(keep), _stream()), "demf_aligned_3d_nms")
    return keep.bool()


def levels_to_rows(levels, out=None):
    """Pyramid levels, each (B,C,H,W) contiguous -> token rows (B, sum H*W, C) in level order, one launch
    (the flatten(2).transpose(1,2) + cat of the reference's multi-level flatten)."""
    import ctypes
    _need_cuda(*levels)
    B, C = levels[0].shape[:2]
    assert 1 <= len(levels) <= 8
    for f in levels:
        assert f.is_contiguous() and f.dtype == torch.float32 and f.shape[:2] == (B, C)
    hw = [f.shape[2] * f.shape[3] for f in levels]
    if out is None:
        out = torch.empty(B, sum(hw), C, dtype=torch.float32, device=levels[0].device)
    assert out.is_contiguous() and out.shape == (B, sum(hw), C)
    ptrs = (ctypes.c_void_p * len(levels))(*[f.data_ptr() for f in levels])
    sizes = (ctypes.c_int * len(levels))(*hw)
    with torch.cuda.device_of(out):
        _lib.check(_lib.load().demf_levels_to_rows(ptrs, sizes, len(levels), B, C, _p(out), _stream()),
                   "demf_levels_to_rows")
    return out


def vote_tail(votes, seed_xyz, seed_rows, xyz_range=None, norm_feats=True):
    """VoteModule tail for one vote per seed: votes (R, >=3+C) rows of conv_out, seed_xyz (B,N,3),
    seed_rows (B,N,C) -> vote_xyz (B,N,3), offset (B,N,3), vote_rows (B,N,C); one launch."""
    import ctypes
    _need_cuda(votes, seed_xyz, seed_rows)
    B, N, C = seed_rows.shape
    assert votes.dim() == 2 and votes.stride(1) == 1 and votes.shape[0] == B * N and votes.shape[1] >= C + 3
    assert seed_xyz.is_contiguous() and seed_rows.is_contiguous() and seed_xyz.shape == (B, N, 3)
    vote_xyz = torch.empty_like(seed_xyz)
    offset = torch.empty_like(seed_xyz)
    vote_rows = torch.empty_like(seed_rows)
    rng = (ctypes.c_float * 3)(*[float(v) for v in xyz_range]) if xyz_range is not None else None
    with torch.cuda.device_of(votes):
        _lib.check(_lib.load().demf_vote_tail(_p(votes), votes.stride(0), _p(seed_xyz), _p(seed_rows), B * N, C,
                                              rng, int(bool(norm_feats)), _p(vote_xyz), _p(offset), _p(vote_rows),
                                              _stream()), "demf_vote_tail")
    return vote_xyz, offset, vote_rows


class _BatchNormReluRows(torch.autograd.Function):
    """Training-mode BatchNorm (+ ReLU) on rows, two launches forward and two backward (csrc/bn_rows.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu, state, prestats=False):
        R, C = x.shape
        y = torch.empty_like(x)
        if prestats:   # sums already in `state` (epilogue of the GEMM that produced x): finalise + one pass
            mean, invstd = bn_finalize(state, R, C, eps, momentum, running_mean, running_var)
            with torch.cuda.device_of(x):
                _lib.check(_lib.load().demf_bn_rows_apply(
                    _p(x), R, C, _p(gamma), _p(beta), _p(mean), _p(invstd), int(relu), _p(y), _stream()),
                    "demf_bn_rows_apply")
        else:
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            invstd = torch.empty(C, dtype=torch.float32, device=x.device)
            with torch.cuda.device_of(x):
                _lib.check(_lib.load().demf_bn_rows_fwd(
                    _p(x), R, C, _p(gamma), _p(beta), float(eps), float(momentum), int(relu),
                    _p(running_mean) if running_mean is not None else None,
                    _p(running_var) if running_var is not None else None, _p(state), _p(mean), _p(invstd), _p(y),
                    _stream()), "demf_bn_rows_fwd")
        ctx.save_for_backward(x, y, gamma, mean, invstd, state)
        ctx.relu = bool(relu)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        x, y, gamma, mean, invstd, state = ctx.saved_tensors
        R, C = x.shape
        grad_y = grad_y.contiguous()
        grad_x = torch.empty_like(x)
        grads = torch.empty(4, C, dtype=torch.float32, device=x.device)     # grad_gamma, grad_beta, coef(2)
        with torch.cuda.device_of(x):
            _lib.check(_lib.load().demf_bn_rows_bwd(
                _p(grad_y), _p(y), _p(x), R, C, _p(gamma), _p(mean), _p(invstd), int(ctx.relu), _p(state),
                _p(grads[2:]), _p(grad_x), _p(grads[0]), _p(grads[1]), _stream()), "demf_bn_rows_bwd")
        return grad_x, grads[0], grads[1], None, None, None, None, None, None, None


class _BatchNormReluMaxRows(torch.autograd.Function):
    """Training BatchNorm + ReLU + max over the ns rows of each centre (csrc/bn_rows.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, ns, state, prestats=False):
        R, C = x.shape
        M = R // ns
        pooled = torch.empty(M, C, dtype=torch.float32, device=x.device)
        arg = torch.empty(M, C, dtype=torch.uint8, device=x.device)
        if prestats:
            mean, invstd = bn_finalize(state, R, C, eps, momentum, running_mean, running_var)
            with torch.cuda.device_of(x):
                _lib.check(_lib.load().demf_bn_max_rows_apply(
                    _p(x), M, int(ns), C, _p(gamma), _p(beta), _p(mean), _p(invstd), _p(pooled), _p(arg),
                    _stream()), "demf_bn_max_rows_apply")
        else:
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            invstd = torch.empty(C, dtype=torch.float32, device=x.device)
            with torch.cuda.device_of(x):
                _lib.check(_lib.load().demf_bn_max_rows_fwd(
                    _p(x), M, int(ns), C, _p(gamma), _p(beta), float(eps), float(momentum),
                    _p(running_mean) if running_mean is not None else None,
                    _p(running_var) if running_var is not None else None, _p(state), _p(mean), _p(invstd),
                    _p(pooled), _p(arg), _stream()), "demf_bn_max_rows_fwd")
        ctx.save_for_backward(x, pooled, arg, gamma, mean, invstd, state)
        ctx.ns = int(ns)
        return pooled

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_pooled):
        x, pooled, arg, gamma, mean, invstd, state = ctx.saved_tensors
        R, C = x.shape
        M = R // ctx.ns
        grad_pooled = grad_pooled.contiguous()
        grad_x = torch.empty_like(x)
        grads = torch.empty(4, C, dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _lib.check(_lib.load().demf_bn_max_rows_bwd(
                _p(grad_pooled), _p(pooled), _p(arg), _p(x), M, ctx.ns, C, _p(gamma), _p(mean), _p(invstd),
                _p(state), _p(grads[2:]), _p(grad_x), _p(grads[0]), _p(grads[1]), _stream()),
                "demf_bn_max_rows_bwd")
        return grad_x, grads[0], grads[1], None, None, None, None, None, None, None


def batch_norm_relu_max_rows(x, gamma, beta, running_mean, running_var, momentum, eps, ns, state, prestats=False):
    """x (M*ns, C) rows -> (M, C) = max over each centre's ns rows of relu(batch_norm(x)); differentiable in
    x, gamma, beta. The normalised tensor is never materialised. `prestats`: the sums are already in `state`
    (gemm_rows_fwd(..., bn_state=state) produced x)."""
    _need_cuda(x, gamma, beta)
    assert x.dim() == 2 and x.is_contiguous() and x.dtype == torch.float32 and x.shape[0] % ns == 0 and ns <= 255
    return _BatchNormReluMaxRows.apply(x, gamma, beta, running_mean, running_var, momentum, eps, ns, state,
                                       bool(prestats))


def bn_rows_supported(channels):
    return bool(_lib.load().demf_bn_rows_supported(int(channels)))


def bn_rows_state(channels, device):
    """Zeroed persistent accumulator block a BatchNorm layer owns for batch_norm_relu_rows."""
    nbytes = int(_lib.load().demf_bn_rows_state_bytes(int(channels)))
    return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=device)


def batch_norm_relu_rows(x, gamma, beta, running_mean, running_var, momentum, eps, relu, state, prestats=False):
    """y = [relu](batch_norm(x)) with batch statistics over the rows of x (R, C), differentiable in x,
    gamma, beta; running statistics updated in place. `state` from bn_rows_state(C, device). `prestats`: the
    sums are already in `state` (gemm_rows_fwd(..., bn_state=state) produced x)."""
    _need_cuda(x, gamma, beta)
    assert x.dim() == 2 and x.is_contiguous() and x.dtype == torch.float32 and x.shape[0] > 0
    return _BatchNormReluRows.apply(x, gamma, beta, running_mean, running_var, momentum, eps, relu, state,
                                    bool(prestats))


def project_points(xyz, mats, affs):
    """xyz (B,Q,3), mats (B,3,4), affs (B,4) -> (B,Q,2) normalised, clamped image coordinates; one launch
    (geometry.project_batched: batched GEMM, divide, scale, shift, clamp). Inference only."""
    _need_cuda(xyz, mats, affs)
    xyz, mats, affs = xyz.contiguous(), mats.contiguous(), affs.contiguous()
    B, Q = xyz.shape[:2]
    assert mats.shape == (B, 3, 4) and affs.shape == (B, 4) and xyz.dtype == torch.float32
    out = torch.empty(B, Q, 2, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device_of(xyz):
        _lib.check(_lib.load().demf_project_points(_p(xyz), _p(mats), _p(affs), B, Q, _p(out), _stream()),
                   "demf_project_points")
    return out


def nms_select(boxes, obj_scores, sem_scores, counts, min_points, nms_thr, score_thr):
    """multiclass_nms_single up to the selection mask, for the batch: boxes (B,K,7) gravity-centre,
    obj_scores (B,K), sem_scores (B,K,C), counts (B,K) i32 -> selected (B,K) bool, classes (B,K) i64,
    num_selected (B,) i32 (device). Three launches (csrc/postprocess.cu)."""
    _need_cuda(boxes, obj_scores, sem_scores, counts)
    boxes, obj_scores, sem_scores = boxes.contiguous().float(), obj_scores.contiguous().float(), \
        sem_scores.contiguous().float()
    B, K, C = sem_scores.shape
    dev = boxes.device
    minmax = torch.empty(B, K, 6, dtype=torch.float32, device=dev)
    classes = torch.empty(B, K, dtype=torch.int64, device=dev)
    valid = torch.empty(B, K, dtype=torch.uint8, device=dev)
    selected = torch.zeros(B, K, dtype=torch.uint8, device=dev)
    nsel = torch.zeros(B, dtype=torch.int32, device=dev)
    if B and K:
        with torch.cuda.device_of(boxes):
            _lib.check(_lib.load().demf_nms_select(
                _p(boxes), _p(obj_scores), _p(sem_scores), _p(counts), B, K, C, int(min_points), float(nms_thr),
                float(score_thr), _p(minmax), _p(classes), _p(valid), _p(selected), _p(nsel), _stream()),
                "demf_nms_select")
    return selected.bool(), classes, nsel


# ------------------------------------------------------------ training GEMMs (csrc/gemm_tf32.cu) ---
def gemm_supported(K, N):
    """Shapes the TMA-fed tcgen05 GEMM kernels take: row strides multiples of 16 bytes, K <= 512."""
    return bool(_lib.load().demf_gemm_supported(int(K), int(N)))


def _rows2d(t):
    assert t.dim() == 2 and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) % 4 == 0 \
        and t.data_ptr() % 16 == 0, "fp32 rows with a 16-byte aligned base and row stride"
    return t


def gemm_rows_fwd(x, w, bias=None, relu=False, bn_state=None, out=None):
    """y (R,N) = x (R,K) @ w (N,K)^T [+ bias] [ReLU] on the tcgen05 tensor cores (TF32 products, fp32
    accumulation). `bn_state` (from bn_rows_state(N, device)): the per-channel sum / sum of squares of y are
    added to the layer's accumulators by the kernel's epilogue (see bn_finalize)."""
    _need_cuda(x, w, bias)
    x, w = _rows2d(x), _rows2d(w)
    R, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    y = torch.empty(R, N, dtype=torch.float32, device=x.device) if out is None else _rows2d(out)
    with torch.cuda.device_of(x):
        _lib.check(_lib.load().demf_gemm_rows_fwd(
            _p(x), x.stride(0), _p(w), w.stride(0), _p(bias), R, K, N, int(bool(relu)), _p(bn_state), _p(y),
            y.stride(0), _stream()), "demf_gemm_rows_fwd")
    return y


def gemm_rows_dgrad(dy, w, out=None):
    """dx (R,K) = dy (R,N) @ w (N,K): the same weight matrix read as an MN-major tensor-core operand."""
    _need_cuda(dy, w)
    dy, w = _rows2d(dy), _rows2d(w)
    R, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N
    dx = torch.empty(R, K, dtype=torch.float32, device=dy.device) if out is None else _rows2d(out)
    with torch.cuda.device_of(dy):
        _lib.check(_lib.load().demf_gemm_rows_dgrad(
            _p(dy), dy.stride(0), _p(w), w.stride(0), R, N, K, _p(dx), dx.stride(0), _stream()),
            "demf_gemm_rows_dgrad")
    return dx


def gemm_wgrad_(dw, dy, x):
    """dw (N,K) += dy (R,N)^T @ x (R,K), split over row slabs (red.global.add); in place, returns dw."""
    _need_cuda(dw, dy, x)
    dy, x = _rows2d(dy), _rows2d(x)
    R, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == R and dw.shape == (N, K) and dw.dtype == torch.float32 and dw.stride(1) == 1
    with torch.cuda.device_of(dy):
        for k0 in range(0, K, 512):     # the kernel's TMEM accumulator holds 512 input channels per launch
            kw = min(512, K - k0)
            xs, ds = x[:, k0:k0 + kw], dw[:, k0:k0 + kw]
            _lib.check(_lib.load().demf_gemm_wgrad(
                _p(dy), dy.stride(0), _p(xs), x.stride(0), R, N, kw, _p(ds), dw.stride(0), _stream()),
                "demf_gemm_wgrad")
    return dw


def gemm_error():
    return int(_lib.load().demf_gemm_error())


def bn_finalize(state, R, C, eps, momentum, running_mean, running_var):
    """(mean, invstd) of the batch from the sums a gemm_rows_fwd(..., bn_state=state) epilogue accumulated;
    updates the running statistics and re-zeroes the accumulators."""
    mean = torch.empty(C, dtype=torch.float32, device=state.device)
    invstd = torch.empty(C, dtype=torch.float32, device=state.device)
    with torch.cuda.device_of(state):
        _lib.check(_lib.load().demf_bn_finalize(
            _p(state), int(R), int(C), float(eps), float(momentum), _p(mean), _p(invstd), _p(running_mean),
            _p(running_var), _stream()), "demf_bn_finalize")
    return mean, invstd


def gemm_rows_dgrad_bn(dy, w, y_prev, mean, invstd, gamma, beta, bn_state):
    """g (R,K) = (dy (R,N) @ w (N,K)) zeroed where relu(bn(y_prev)) was inactive; the BatchNorm backward's two
    reductions (sum g, sum g*y_prev) land in `bn_state`. Returns g."""
    _need_cuda(dy, w, y_prev)
    dy, w, y_prev = _rows2d(dy), _rows2d(w), _rows2d(y_prev)
    R, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N and y_prev.shape == (R, K)
    g = torch.empty(R, K, dtype=torch.float32, device=dy.device)
    with torch.cuda.device_of(dy):
        _lib.check(_lib.load().demf_gemm_rows_dgrad_bn(
            _p(dy), dy.stride(0), _p(w), w.stride(0), R, N, K, _p(y_prev), y_prev.stride(0), _p(mean), _p(invstd),
            _p(gamma), _p(beta), _p(bn_state), _p(g), g.stride(0), _stream()), "demf_gemm_rows_dgrad_bn")
    return g


def bn_bwd_from_masked(g, y_prev, gamma, mean, invstd, bn_state):
    """Finish the BatchNorm backward after gemm_rows_dgrad_bn: -> (grad of y_prev, grad_gamma, grad_beta)."""
    R, C = g.shape
    grads = torch.empty(4, C, dtype=torch.float32, device=g.device)     # grad_gamma, grad_beta, coef(2)
    grad_x = torch.empty_like(g)
    lib = _lib.load()
    with torch.cuda.device_of(g):
        _lib.check(lib.demf_bn_bwd_finalize(_p(bn_state), R, C, _p(mean), _p(invstd), _p(grads[0]), _p(grads[1]),
                                            _p(grads[2:]), _stream()), "demf_bn_bwd_finalize")
        _lib.check(lib.demf_bn_rows_bwd_apply(_p(g), _p(y_prev), R, C, _p(gamma), _p(mean), _p(invstd), _p(grads[2:]),
                                              _p(grad_x), _stream()), "demf_bn_rows_bwd_apply")
    return grad_x, grads[0], grads[1]


def bn_rows_apply(y, gamma, beta, mean, invstd, relu=True):
    """z = [relu]((y - mean) * invstd * gamma + beta) with given statistics (one pass)."""
    R, C = y.shape
    z = torch.empty_like(y)
    with torch.cuda.device_of(y):
        _lib.check(_lib.load().demf_bn_rows_apply(_p(y), R, C, _p(gamma), _p(beta), _p(mean), _p(invstd), int(relu),
                                                  _p(z), _stream()), "demf_bn_rows_apply")
    return z


# ------------------------------------------------------------ per-stage detection loss (csrc/loss.cu) ---
class _StageLoss(torch.autograd.Function):
    """(7,) = weighted sums (objectness, dir_class, dir_res, size, center, semantic, iou) of one prediction stage;
    one launch forward, one backward."""

    @staticmethod
    def forward(ctx, center, size, dir_class, dir_res_norm, obj, sem, targets, cfg):
        obj_t, obj_w, box_w, size_t, center_t, dir_class_t, dir_res_t, sem_t = targets
        rows = center.shape[0] * center.shape[1]
        nb = dir_class.shape[-1]
        ns = 0 if sem is None else sem.shape[-1]
        cfg_c = (ctypes.c_float * 12)(*cfg)
        out = torch.zeros(7, dtype=torch.float32, device=center.device)
        with torch.cuda.device_of(center):
            _lib.check(_lib.load().demf_stage_loss_fwd(
                _p(center), _p(size), _p(dir_class), _p(dir_res_norm), _p(obj), _p(sem), _p(obj_t), _p(obj_w),
                _p(box_w), _p(size_t), _p(center_t), _p(dir_class_t), _p(dir_res_t), _p(sem_t), rows, nb, ns,
                ctypes.cast(cfg_c, ctypes.c_void_p), _p(out), _stream()), "demf_stage_loss_fwd")
        ctx.save_for_backward(center, size, dir_class, dir_res_norm, obj, *( [sem] if sem is not None else []),
                              obj_t, obj_w, box_w, size_t, center_t, dir_class_t, dir_res_t,
                              *([sem_t] if sem is not None else []))
        ctx.has_sem = sem is not None
        ctx.cfg = tuple(cfg)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, up):
        t = list(ctx.saved_tensors)
        center, size, dir_class, dir_res_norm, obj = t[:5]
        k = 5
        sem = None
        if ctx.has_sem:
            sem = t[k]
            k += 1
        obj_t, obj_w, box_w, size_t, center_t, dir_class_t, dir_res_t = t[k:k + 7]
        sem_t = t[k + 7] if ctx.has_sem else None
        rows = center.shape[0] * center.shape[1]
        nb = dir_class.shape[-1]
        ns = 0 if sem is None else sem.shape[-1]
        cfg_c = (ctypes.c_float * 12)(*ctx.cfg)
        up = up.contiguous().float()
        gs = [torch.empty_like(x) for x in (center, size, dir_class, dir_res_norm, obj)]
        g_sem = torch.empty_like(sem) if sem is not None else None
        with torch.cuda.device_of(center):
            _lib.check(_lib.load().demf_stage_loss_bwd(
                _p(center), _p(size), _p(dir_class), _p(dir_res_norm), _p(obj), _p(sem), _p(obj_t), _p(obj_w),
                _p(box_w), _p(size_t), _p(center_t), _p(dir_class_t), _p(dir_res_t), _p(sem_t), rows, nb, ns,
                ctypes.cast(cfg_c, ctypes.c_void_p), _p(up), _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(gs[4]),
                _p(g_sem), _stream()), "demf_stage_loss_bwd")
        return gs[0], gs[1], gs[2], gs[3], gs[4], g_sem, None, None


def mha_supported(head_dim):
    return bool(_lib.load().demf_mha_supported(int(head_dim)))


def mha_rows(q, k, v, B, H, scale=None, batch_first=False):
    """softmax(q k^T * scale) v per (scene, head) in exact fp32 (csrc/mha.cu), inference only. q (Lq*B, E) rows of a
    (Lq, B, E) tensor (a column slice of a wider projection is fine: unit column stride, 16-byte aligned rows), k / v
    (Lk*B, E); heads are the H column groups of width E / H. `batch_first`: the rows are those of (B, L, E) tensors
    instead. -> (Lq*B, E) in the same row order."""
    _need_cuda(q, k, v)
    E = q.shape[1]
    D = E // H
    Lq, Lk = q.shape[0] // B, k.shape[0] // B
    for t in (q, k, v):
        assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.float32 and t.shape[1] == E
    assert k.shape[0] == v.shape[0] and q.shape[0] == Lq * B and k.shape[0] == Lk * B and D * H == E
    out = torch.empty(Lq * B, E, dtype=torch.float32, device=q.device)
    with torch.cuda.device_of(q):
        _lib.check(_lib.load().demf_mha_fwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                                            out.stride(0), Lq, Lk, B, H, D,
                                            float(D ** -0.5 if scale is None else scale), int(bool(batch_first)),
                                            _stream()), "demf_mha_fwd")
    return out


def col_sum_add_(out, x):
    """out (N,) += x (R,N).sum(0) in one launch (csrc/bn_rows.cu): the bias gradient of a Linear layer."""
    _need_cuda(out, x)
    assert x.dim() == 2 and x.stride(1) == 1 and out.is_contiguous() and out.numel() == x.shape[1]
    assert x.dtype == torch.float32 and out.dtype == torch.float32
    with torch.cuda.device_of(x):
        _lib.check(_lib.load().demf_col_sum_add(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), _stream()),
                   "demf_col_sum_add")
    return out


def stage_loss(center, size, dir_class, dir_res_norm, obj, sem, targets, cfg):
    """Predictions (B,Q,C) fp32 contiguous CUDA tensors; targets = (objectness_targets i64, objectness_weights,
    box_loss_weights, size_targets, center_targets, dir_class_targets i64, dir_res_targets, mask_targets i64 | None);
    cfg = 12 floats (see include/demf_b200.h). -> (7,) losses, differentiable in the predictions."""
    _need_cuda(center, size, dir_class, dir_res_norm, obj, sem)
    for x in (center, size, dir_class, dir_res_norm, obj) + ((sem,) if sem is not None else ()):
        assert x.is_contiguous() and x.dtype == torch.float32
    targets = tuple(None if t is None else t.contiguous() for t in targets)
    assert targets[0].dtype == torch.int64 and targets[5].dtype == torch.int64
    return _StageLoss.apply(center, size, dir_class, dir_res_norm, obj, sem, targets, tuple(float(c) for c in cfg))


# ------------------------------------------------- pipelined first-level set abstraction (csrc/sa_pipe.cu) ---
def sa_pipe_supported(C, sample_num, widths, M):
    return bool(_lib.load().demf_sa_pipe_supported(int(C), int(sample_num), int(widths[0]), int(widths[1]),
                                                   int(widths[2]), int(M)))


def sa_pipe(xyz, center_xyz, feat_rows, max_radius, sample_num, normalize_xyz, wpack, bias, idx, packed=None):
    """sa_fused for the first backbone level (one feature channel, widths 64/64/128) as a warp-specialised
    pipeline over 128-row tiles; `idx` = the ball-query rows (B,M,ns) i32. -> (B,M,128) rows. Inference only.
    `packed`: the same cloud as one contiguous (B,N,4) tensor [xyz | feature] (the detector's raw input), when the
    caller has it: the gather then issues one 16-byte load per neighbour."""
    assert xyz.is_contiguous() and center_xyz.is_contiguous() and feat_rows.is_contiguous() and idx.is_contiguous()
    _need_cuda(xyz, center_xyz, feat_rows, wpack, bias, idx)
    B, N, _ = xyz.shape
    M = center_xyz.size(1)
    assert feat_rows.shape == (B, N, 1) and idx.shape == (B, M, sample_num) and idx.dtype == torch.int32
    if packed is not None:
        assert packed.shape == (B, N, 4) and packed.is_contiguous() and packed.dtype == torch.float32
        _need_cuda(packed)
    with torch.cuda.device_of(xyz):
        out = torch.empty(B, M, 128, dtype=torch.float32, device=xyz.device)
        _lib.check(_lib.load().demf_sa_pipe_fwd(_p(xyz), _p(feat_rows), _p(packed), _p(center_xyz), _p(idx), B, N, M,
                                                int(sample_num), float(max_radius), int(bool(normalize_xyz)),
                                                _p(wpack), _p(bias), _p(out), _stream()), "demf_sa_pipe_fwd")
    return out


def sa_pipe_error():
    return int(_lib.load().demf_sa_pipe_error())
