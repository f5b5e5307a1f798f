"""PointNet++ modules with the mmdet3d names the DeMF configs instantiate.

Stands in for mmdet3d 0.18.1 `ops/pointnet_modules/{point_sa_module,point_fp_module}.py`,
`models/backbones/pointnet2_sa_ssg.py`, `models/model_utils/vote_module.py` and
`models/dense_heads/base_conv_bbox_head.py` (reached from configs/demf/demf_votenet.py:48-62,
142-162 and demf/modeling/heads/class_agnostic_vote_head.py:382-403). Constructor arguments,
forward signatures, return shapes and parameter names (state-dict keys) are upstream's.

What is different is the data layout INSIDE a module: activations are point-major rows
(B,N,C) -- a neighbour is one contiguous row for the grouping kernel (csrc/rows.cu), and every
kernel-size-1 convolution is a plain GEMM over rows -- and a module hands its features to the
next one as a (B,C,N) *view* of those rows, so the upstream interface holds without a single
layout copy between modules.
"""
import torch
import torch.nn as nn

from . import point_ops as P
from .bricks import (BaseModule, ConvModule, _foldable, _folded_cached, as_rows, build_conv_layer,
                     conv_module_rows_max, conv_module_rows, fold_key, linear_rows, sa_mlp_train_rows)
from .registry import BACKBONES, SA_MODULES


# ------------------------------------------------------------------------- SA ---------
class BasePointSAModule(nn.Module):
    """mmdet3d BasePointSAModule: sample centres, group neighbours, shared MLP, pool."""

    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], dilated_group=False, use_xyz=True, pool_mod='max',
                 normalize_xyz=False, grouper_return_grouped_xyz=False,
                 grouper_return_grouped_idx=False):
        super().__init__()
        assert len(radii) == len(sample_nums) == len(mlp_channels)
        assert pool_mod in ['max', 'avg']
        assert isinstance(fps_mod, (list, tuple))
        assert isinstance(fps_sample_range_list, (list, tuple))
        assert len(fps_mod) == len(fps_sample_range_list)
        if isinstance(mlp_channels, tuple):
            mlp_channels = list(map(list, mlp_channels))
        self.mlp_channels = mlp_channels
        if isinstance(num_point, int):
            self.num_point = [num_point]
        elif isinstance(num_point, (list, tuple)):
            self.num_point = list(num_point)
        elif num_point is None:
            self.num_point = None
        else:
            raise NotImplementedError('Error type of num_point!')
        if list(fps_mod) != ['D-FPS'] or list(fps_sample_range_list) != [-1]:
            raise NotImplementedError("only D-FPS over the whole cloud is on the DeMF path")
        self.pool_mod = pool_mod
        self.use_xyz = use_xyz
        self.normalize_xyz = normalize_xyz
        self.dilated_group = dilated_group
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        self.fps_mod_list = fps_mod
        self.fps_sample_range_list = fps_sample_range_list
        for i in range(len(radii)):
            radius = radii[i]
            sample_num = sample_nums[i]
            if num_point is not None:
                if dilated_group and i != 0:
                    min_radius = radii[i - 1]
                else:
                    min_radius = 0
                grouper = P.QueryAndGroup(
                    radius, sample_num, min_radius=min_radius, use_xyz=use_xyz,
                    normalize_xyz=normalize_xyz,
                    return_grouped_xyz=grouper_return_grouped_xyz,
                    return_grouped_idx=grouper_return_grouped_idx)
            else:
                grouper = P.GroupAll(use_xyz)
            self.groupers.append(grouper)

    def _sample_points(self, points_xyz, features, indices, target_xyz):
        if indices is not None:
            assert indices.shape[1] == self.num_point[0]
            if target_xyz is not None:  # centres already gathered by the caller (sampling chain)
                new_xyz = target_xyz
            else:
                new_xyz = P.gather_rows(points_xyz, indices) if self.num_point is not None else None
        elif target_xyz is not None:
            new_xyz = target_xyz.contiguous()
        else:
            if self.num_point is not None:
                indices = P.furthest_point_sample(points_xyz.contiguous(), self.num_point[0])
                new_xyz = P.gather_rows(points_xyz, indices)
            else:
                new_xyz = None
        return new_xyz, indices

    def _pool_features(self, features):
        """(B,C,M,ns) -> (B,C,M) (channel-major path)."""
        if self.pool_mod == 'max':
            return features.max(dim=-1)[0]
        return features.mean(dim=-1)

    def _rows_ok(self, grouper):
        return (isinstance(grouper, P.QueryAndGroup) and grouper.use_xyz
                and not (grouper.return_grouped_xyz or grouper.return_grouped_idx
                         or grouper.uniform_sample))

    # Inference: the whole level -- ball query, grouping, the three BN-folded 1x1 convolutions
    # and the max over the neighbourhood -- is ONE tcgen05 kernel (csrc/sa_fused.cu); neither the
    # grouped tensor nor any activation reaches HBM. `fused_eval = False` keeps the layer-by-layer
    # path (group rows kernel + library GEMMs), which is also what training uses.
    fused_eval = True
    pipe_eval = True      # first level: warp-specialised tile pipeline (csrc/sa_pipe.cu) instead of sa_fused.cu

    def _fused_pack(self, mlp, C):
        """(wpack, bias, widths) of this MLP for P.sa_fused, rebuilt when a parameter changes."""
        layers = list(mlp)
        folded = [_folded_cached(cm, P.group_rows_columns(C) if j == 0 else None)
                  for j, cm in enumerate(layers)]
        key = tuple(fold_key(cm) for cm in layers)     # (storage, version) of every source parameter
        cache = mlp.__dict__.get("_sa_pack")
        if cache is None or cache[0] != key:
            cache = (key,) + P.sa_pack_mlp([w for w, _ in folded], [b for _, b in folded])
            mlp.__dict__["_sa_pack"] = cache
        return cache[1:]

    def _fused_ok(self, grouper, mlp, points_xyz, C):
        return (self.fused_eval and self.pool_mod == 'max' and points_xyz.is_cuda
                and not torch.is_grad_enabled() and len(mlp) == 3
                and all(isinstance(cm, ConvModule) and _foldable(cm) and cm.with_activation
                        for cm in mlp)
                and P.sa_fused_supported(C, grouper.sample_num,
                                         [cm.conv.out_channels for cm in mlp]))

    # Clouds at least this large are binned into a uniform grid first (P.ball_grid) and the ball
    # query only tests each centre's 3x3x3 cell neighbourhood -- identical index rows.
    grid_min_points = 2048

    def ball_grid(self, points_xyz):
        """Grid workspace for this module's queries on points_xyz, or None for small clouds."""
        if (not points_xyz.is_cuda or points_xyz.size(1) < self.grid_min_points
                or not all(self._rows_ok(g) for g in self.groupers)):
            return None
        return P.ball_grid(points_xyz.contiguous(), max(g.max_radius for g in self.groupers))

    def forward(self, points_xyz, features=None, indices=None, target_xyz=None, grid=None, packed=None):
        """points_xyz (B,N,3), features (B,C,N) -> new_xyz (B,M,3), new_features (B,sum C_k,M),
        indices (B,M). `grid`: a workspace from self.ball_grid(points_xyz) built ahead of time.
        `packed`: the contiguous (B,N,4) tensor points_xyz and a single feature channel were sliced
        from, when the caller still has it (the first-level kernel gathers 16-byte rows from it)."""
        new_xyz, indices = self._sample_points(points_xyz, features, indices, target_xyz)
        points_xyz = points_xyz.contiguous()
        if grid is None:
            grid = self.ball_grid(points_xyz)
        out = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            if self._rows_ok(grouper):
                feat_rows = None if features is None else as_rows(features)
                C = 0 if feat_rows is None else feat_rows.size(-1)
                if self._fused_ok(grouper, mlp, points_xyz, C) and self.pipe_eval and grouper.min_radius == 0 \
                        and P.sa_pipe_supported(C, grouper.sample_num, [cm.conv.out_channels for cm in mlp],
                                                new_xyz.size(1)):
                    # first backbone level: the pipelined kernel (csrc/sa_pipe.cu) behind the exact grid query
                    wpack, bias, widths = self._fused_pack(mlp, C)
                    centres = new_xyz.contiguous()
                    if grid is not None:
                        nbr = P.ball_query_grid(0.0, grouper.max_radius, grouper.sample_num, points_xyz, centres, grid)
                    else:
                        nbr = P.ball_query(0.0, grouper.max_radius, grouper.sample_num, points_xyz, centres)
                    out.append(P.sa_pipe(points_xyz, centres, feat_rows, grouper.max_radius, grouper.sample_num,
                                         grouper.normalize_xyz, wpack, bias, nbr,
                                         packed=packed if (packed is not None and packed.is_cuda and C == 1
                                                           and packed.is_contiguous()) else None))
                    continue
                if self._fused_ok(grouper, mlp, points_xyz, C):
                    wpack, bias, widths = self._fused_pack(mlp, C)
                    out.append(P.sa_fused(points_xyz, new_xyz.contiguous(), feat_rows,
                                          grouper.min_radius, grouper.max_radius,
                                          grouper.sample_num, grouper.normalize_xyz, wpack, bias,
                                          widths, grid=grid))
                    continue
                _, rows = P.query_and_group_rows(
                    points_xyz, new_xyz.contiguous(), feat_rows, grouper.min_radius,
                    grouper.max_radius, grouper.sample_num, grouper.normalize_xyz, grid)
                B, M, ns, K = rows.shape
                x = rows.view(B * M * ns, K)
                layers = list(mlp)
                pooled = None
                if self.pool_mod == 'max' and torch.is_grad_enabled():
                    # training: the whole MLP + max as a chain of fused units on the tensor-core kernels
                    pooled = sa_mlp_train_rows(mlp, x, ns, P.group_rows_columns(C))
                    if pooled is not None:
                        out.append(pooled.view(B, M, -1))
                        continue
                for j, layer in enumerate(layers):
                    cols = P.group_rows_columns(C) if j == 0 else None
                    if j == len(layers) - 1 and self.pool_mod == 'max':
                        # training: the last layer's BatchNorm + ReLU and the max over the neighbourhood in
                        # one pass (the (B*M*ns, C') activation is never written)
                        pooled = conv_module_rows_max(layer, x, ns, cols)
                        if pooled is not None:
                            break
                    x = conv_module_rows(layer, x, cols)
                if pooled is not None:
                    out.append(pooled.view(B, M, -1))
                    continue
                x = x.view(B, M, ns, -1)
                if self.pool_mod == 'max':  # amax: no index tensor when nothing back-propagates
                    x = x.max(dim=2)[0] if torch.is_grad_enabled() else x.amax(dim=2)
                else:
                    x = x.mean(dim=2)
                out.append(x)  # rows (B,M,C')
            else:
                grouped = grouper(points_xyz, new_xyz, features)
                if isinstance(grouped, tuple):
                    grouped = grouped[0]
                out.append(self._pool_features(mlp(grouped)).transpose(1, 2))
        rows = out[0] if len(out) == 1 else torch.cat(out, dim=-1)
        return new_xyz, rows.transpose(1, 2), indices


@SA_MODULES.register_module()
class PointSAModuleMSG(BasePointSAModule):
    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], dilated_group=False, norm_cfg=dict(type='BN2d'),
                 use_xyz=True, pool_mod='max', normalize_xyz=False, bias='auto'):
        super().__init__(num_point=num_point, radii=radii, sample_nums=sample_nums,
                         mlp_channels=mlp_channels, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list, dilated_group=dilated_group,
                         use_xyz=use_xyz, pool_mod=pool_mod, normalize_xyz=normalize_xyz)
        for i in range(len(self.mlp_channels)):
            mlp_channel = self.mlp_channels[i]
            if use_xyz:
                mlp_channel[0] += 3
            mlp = nn.Sequential()
            for j in range(len(mlp_channel) - 1):
                mlp.add_module(
                    f'layer{j}',
                    ConvModule(mlp_channel[j], mlp_channel[j + 1], kernel_size=(1, 1),
                               stride=(1, 1), conv_cfg=dict(type='Conv2d'), norm_cfg=norm_cfg,
                               bias=bias))
            self.mlps.append(mlp)


@SA_MODULES.register_module()
class PointSAModule(PointSAModuleMSG):
    def __init__(self, mlp_channels, num_point=None, radius=None, num_sample=None,
                 norm_cfg=dict(type='BN2d'), use_xyz=True, pool_mod='max', fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], normalize_xyz=False):
        super().__init__(mlp_channels=[list(mlp_channels)], num_point=num_point, radii=[radius],
                         sample_nums=[num_sample], norm_cfg=norm_cfg, use_xyz=use_xyz,
                         pool_mod=pool_mod, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list, normalize_xyz=normalize_xyz)


# ------------------------------------------------------------------------- FP ---------
class PointFPModule(BaseModule):
    """mmdet3d PointFPModule: 3-NN inverse-distance interpolation + skip concat + shared MLP."""

    def __init__(self, mlp_channels, norm_cfg=dict(type='BN2d'), init_cfg=None):
        super().__init__(init_cfg=init_cfg)
        self.fp16_enabled = False
        self.mlps = nn.Sequential()
        for i in range(len(mlp_channels) - 1):
            self.mlps.add_module(
                f'layer{i}',
                ConvModule(mlp_channels[i], mlp_channels[i + 1], kernel_size=(1, 1), stride=(1, 1),
                           conv_cfg=dict(type='Conv2d'), norm_cfg=norm_cfg))

    def forward(self, target, source, target_feats, source_feats):
        """target (B,n,3), source (B,m,3), target_feats (B,C1,n), source_feats (B,C2,m)
        -> (B,M,n)."""
        src_rows = as_rows(source_feats)
        if (source is not None and src_rows.is_cuda and not torch.is_grad_enabled()
                and src_rows.size(-1) % 4 == 0
                and (target_feats is None or target_feats.size(1) % 4 == 0)):
            # inference: weights, interpolation and the skip concatenation in one launch
            d2, idx = P.three_nn_squared(target.contiguous(), source.contiguous())
            x = P.interp_cat_rows(src_rows, None if target_feats is None else as_rows(target_feats),
                                  idx, d2)
            B, n, C = x.shape
            x = x.view(B * n, C)
            for layer in self.mlps:
                x = conv_module_rows(layer, x)
            return x.view(B, n, -1).transpose(1, 2)
        if source is not None:
            dist, idx = P.three_nn(target.contiguous(), source.contiguous())
            dist_reciprocal = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_reciprocal, dim=2, keepdim=True)
            weight = dist_reciprocal / norm
            if src_rows.size(-1) % 4 == 0:
                interpolated = P.three_interpolate_rows(src_rows, idx, weight)
            else:
                interpolated = P.three_interpolate(source_feats.contiguous(), idx,
                                                   weight).transpose(1, 2)
        else:
            interpolated = src_rows.expand(src_rows.size(0), target.size(1), src_rows.size(2))
        if target_feats is not None:
            x = torch.cat([interpolated, as_rows(target_feats)], dim=-1)
        else:
            x = interpolated
        B, n, C = x.shape
        x = x.reshape(B * n, C)
        for layer in self.mlps:
            x = conv_module_rows(layer, x)
        return x.view(B, n, -1).transpose(1, 2)


# ------------------------------------------------------------------- backbone ---------
@BACKBONES.register_module()
class PointNet2SASSG(BaseModule):
    """PointNet++ single-scale-grouping backbone (configs/demf/demf_votenet.py:48-62)."""

    def __init__(self, in_channels, num_points=(2048, 1024, 512, 256), radius=(0.2, 0.4, 0.8, 1.2),
                 num_samples=(64, 32, 16, 16),
                 sa_channels=((64, 64, 128), (128, 128, 256), (128, 128, 256), (128, 128, 256)),
                 fp_channels=((256, 256), (256, 256)), norm_cfg=dict(type='BN2d'),
                 sa_cfg=dict(type='PointSAModule', pool_mod='max', use_xyz=True,
                             normalize_xyz=True),
                 init_cfg=None):
        super().__init__(init_cfg=init_cfg)
        from .registry import build_sa_module
        self.num_sa = len(sa_channels)
        self.num_fp = len(fp_channels)
        assert len(num_points) == len(radius) == len(num_samples) == len(sa_channels)
        assert len(sa_channels) >= len(fp_channels)
        self.SA_modules = nn.ModuleList()
        sa_in_channel = in_channels - 3  # number of channels without xyz
        skip_channel_list = [sa_in_channel]
        for sa_index in range(self.num_sa):
            cur_sa_mlps = list(sa_channels[sa_index])
            cur_sa_mlps = [sa_in_channel] + cur_sa_mlps
            sa_out_channel = cur_sa_mlps[-1]
            self.SA_modules.append(
                build_sa_module(num_point=num_points[sa_index], radius=radius[sa_index],
                                num_sample=num_samples[sa_index], mlp_channels=cur_sa_mlps,
                                norm_cfg=norm_cfg, cfg=sa_cfg))
            skip_channel_list.append(sa_out_channel)
            sa_in_channel = sa_out_channel
        self.FP_modules = nn.ModuleList()
        fp_source_channel = skip_channel_list.pop()
        fp_target_channel = skip_channel_list.pop()
        for fp_index in range(len(fp_channels)):
            cur_fp_mlps = list(fp_channels[fp_index])
            cur_fp_mlps = [fp_source_channel + fp_target_channel] + cur_fp_mlps
            self.FP_modules.append(PointFPModule(mlp_channels=cur_fp_mlps))
            if fp_index != len(fp_channels) - 1:
                fp_source_channel = cur_fp_mlps[-1]
                fp_target_channel = skip_channel_list.pop()

    def _identity_indices(self, batch, num_points, device):
        """(B,N) int64 arange rows (sa_indices[0]); constant, so built once per shape and device."""
        cache = self.__dict__.setdefault("_arange_cache", {})
        key = (batch, num_points, str(device))
        if key not in cache:
            cache[key] = torch.arange(num_points, device=device).unsqueeze(0).repeat(batch, 1).long()
        return cache[key]

    @staticmethod
    def _split_point_feats(points):
        xyz = points[..., 0:3].contiguous()
        features = points[..., 3:].transpose(1, 2) if points.size(-1) > 3 else None
        return xyz, features

    # Furthest point sampling only ever sees coordinates: level i+1 samples the centres level i
    # picked, never its features. The whole sampling chain (2048 -> 1024 -> 512 -> 256, plus the
    # head's 1024 -> 256 proposal sampling when asked for) is therefore issued up front on a
    # side stream, and the grouping / MLP work of level i on the main stream overlaps the serial
    # FPS iterations of levels > i. Each level publishes an event the main stream waits on.
    prefetch_seed_fps = None   # set by the detector: (fp level whose xyz are the seeds, m)
    grid_fps = True            # first-level FPS through the ball-query grid (identical indices)
    chain_shortcut = True      # levels > 0: certified pick sequences are sampled without iterating
    overlap_sampling = True    # False: run the sampling chain on the current stream

    def _side_stream(self, device):
        streams = self.__dict__.setdefault("_streams", {})
        if device not in streams:
            streams[device] = torch.cuda.Stream(device=device)
        return streams[device]

    presampled = None   # result of sample() on the cloud the next forward() will see (see below)

    def sample(self, points):
        """The sampling chain of forward(points) on the current stream: it depends on the coordinates
        only, not on any weight, so a caller may run it AHEAD of the forward (e.g. for the next batch
        while this batch's backward runs) and hand the result over through `self.presampled`."""
        xyz, _ = self._split_point_feats(points)
        return self._sampling_chain(xyz, overlap=False)

    def _sampling_chain(self, xyz, overlap=None):
        """-> per-level (indices (B,M) i32, new_xyz (B,M,3), ready event), seed fps or None, and the
        ball-query grid of the input cloud (or None)."""
        if overlap is None:
            overlap = xyz.is_cuda and self.overlap_sampling
        levels, seed_fps, grids = [], None, []
        if overlap:
            main = torch.cuda.current_stream(xyz.device)
            side = self._side_stream(xyz.device)
            side.wait_stream(main)
            ctx = torch.cuda.stream(side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            cur = xyz
            # a large input cloud is binned once into the uniform grid that serves both the first
            # level's ball query and its (grid-pruned) furthest point sampling
            grid0 = self.SA_modules[0].ball_grid(xyz)
            grids.append(grid0)
            prefix = None     # uniqueness certificate of the first level's picks (chain shortcut)
            for i, sa in enumerate(self.SA_modules):
                if cur.is_cuda:   # the kernel writes the picked coordinates next to the indices
                    if i == 0 and grid0 is not None and self.grid_fps and self.chain_shortcut:
                        later = [s.num_point[0] for s in self.SA_modules[1:]]
                        if self.prefetch_seed_fps is not None:
                            later.append(self.prefetch_seed_fps[1])
                        idx, new_xyz, prefix = P.furthest_point_sample_xyz(
                            cur, sa.num_point[0], grid0, return_prefix=True, certify=max(later) if later else 1)
                    else:
                        # levels > 0 sample the previous level's pick sequence: scenes whose certificate
                        # covers this level's count get idx = 0..m-1 without iterating (same result)
                        idx, new_xyz = P.furthest_point_sample_xyz(
                            cur, sa.num_point[0], grid0 if (i == 0 and self.grid_fps) else None,
                            unique_prefix=prefix if i > 0 else None)
                else:
                    idx = P.furthest_point_sample(cur, sa.num_point[0])
                    new_xyz = P.gather_rows(cur, idx).contiguous()
                ev = None
                if overlap:
                    ev = torch.cuda.Event()
                    ev.record(side)
                cur = new_xyz
                # the next level's ball-query grid over the centres just picked (None for small clouds):
                # built here, off the main stream; the NEXT level's event (recorded later on this
                # stream) covers it
                if i + 1 < len(self.SA_modules):
                    grids.append(self.SA_modules[i + 1].ball_grid(cur))
                levels.append((idx, new_xyz, ev))
                if self.prefetch_seed_fps is not None and self.prefetch_seed_fps[0] == i + 1:
                    if cur.is_cuda and prefix is not None:   # the seeds are a pick sequence too
                        seed_idx = P.furthest_point_sample_xyz(cur, self.prefetch_seed_fps[1],
                                                               unique_prefix=prefix)[0]
                    else:
                        seed_idx = P.furthest_point_sample(cur, self.prefetch_seed_fps[1])
                    seed_fps = (seed_idx, torch.cuda.Event() if overlap else None)
                    if overlap:
                        seed_fps[1].record(side)
        return levels, seed_fps, grids

    def forward(self, points):
        """points (B,N,3+C) -> dict of fp_xyz / fp_features / fp_indices / sa_*."""
        xyz, features = self._split_point_feats(points)
        batch, num_points = xyz.shape[:2]
        chained = all(getattr(sa, "num_point", None) is not None and len(sa.num_point) == 1
                      for sa in self.SA_modules)
        if self.presampled is not None:
            levels, seed_fps, grids = self.presampled
        else:
            levels, seed_fps, grids = self._sampling_chain(xyz) if chained else (None, None, None)
        indices = self._identity_indices(batch, num_points, xyz.device)
        sa_xyz, sa_features, sa_indices = [xyz], [features], [indices]
        # indices of every level into the ORIGINAL cloud: on CUDA one launch for the whole chain,
        # issued after the loop (the main stream has then waited for every level's sampling)
        chain_on_device = levels is not None and xyz.is_cuda and len(levels) <= 4
        for i in range(self.num_sa):
            if levels is not None:
                idx, new_xyz, ev = levels[i]
                if ev is not None:
                    torch.cuda.current_stream(xyz.device).wait_event(ev)
                cur_xyz, cur_features, cur_indices = self.SA_modules[i](
                    sa_xyz[i], sa_features[i], indices=idx, target_xyz=new_xyz,
                    grid=grids[i] if i < len(grids) else None,
                    packed=points if (i == 0 and points.size(-1) == 4 and not torch.is_grad_enabled()) else None)
            else:
                cur_xyz, cur_features, cur_indices = self.SA_modules[i](sa_xyz[i], sa_features[i])
            sa_xyz.append(cur_xyz)
            sa_features.append(cur_features)
            if not chain_on_device:
                sa_indices.append(torch.gather(sa_indices[-1], 1, cur_indices.long()))
        if chain_on_device:
            sa_indices += P.chain_indices([lv[0] for lv in levels])
        fp_xyz, fp_features, fp_indices = [sa_xyz[-1]], [sa_features[-1]], [sa_indices[-1]]
        for i in range(self.num_fp):
            fp_features.append(self.FP_modules[i](sa_xyz[self.num_sa - i - 1],
                                                  sa_xyz[self.num_sa - i],
                                                  sa_features[self.num_sa - i - 1],
                                                  fp_features[-1]))
            fp_xyz.append(sa_xyz[self.num_sa - i - 1])
            fp_indices.append(sa_indices[self.num_sa - i - 1])
        ret = dict(fp_xyz=fp_xyz, fp_features=fp_features, fp_indices=fp_indices, sa_xyz=sa_xyz,
                   sa_features=sa_features, sa_indices=sa_indices)
        if seed_fps is not None:
            if seed_fps[1] is not None:
                torch.cuda.current_stream(xyz.device).wait_event(seed_fps[1])
            ret['seed_fps_indices'] = seed_fps[0]
        return ret


# ---------------------------------------------------------------- vote module ---------
class VoteModule(nn.Module):
    """mmdet3d VoteModule: per-seed MLP -> (xyz offset, feature residual) -> votes."""

    def __init__(self, in_channels, vote_per_seed=1, gt_per_seed=3, num_points=-1,
                 conv_channels=(16, 16), conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'),
                 act_cfg=dict(type='ReLU'), norm_feats=True, with_res_feat=True,
                 vote_xyz_range=None, vote_loss=None):
        super().__init__()
        from .registry import build_loss
        self.in_channels = in_channels
        self.vote_per_seed = vote_per_seed
        self.gt_per_seed = gt_per_seed
        self.num_points = num_points
        self.norm_feats = norm_feats
        self.with_res_feat = with_res_feat
        assert vote_xyz_range is None or isinstance(vote_xyz_range, (list, tuple))
        self.vote_xyz_range = vote_xyz_range
        if vote_loss is not None:
            self.vote_loss = build_loss(vote_loss)
        prev_channels = in_channels
        vote_conv_list = []
        for k in range(len(conv_channels)):
            vote_conv_list.append(
                ConvModule(prev_channels, conv_channels[k], 1, padding=0, conv_cfg=conv_cfg,
                           norm_cfg=norm_cfg, act_cfg=act_cfg, bias=True, inplace=True))
            prev_channels = conv_channels[k]
        self.vote_conv = nn.Sequential(*vote_conv_list)
        if with_res_feat:
            out_channel = (3 + in_channels) * self.vote_per_seed
        else:
            out_channel = 3 * self.vote_per_seed
        self.conv_out = nn.Conv1d(prev_channels, out_channel, 1)

    fused_eval = True

    def _padded_conv_out(self):
        """conv_out's (3+C, C) weight and bias zero-padded to a multiple of 4 output rows, rebuilt when a
        parameter changes."""
        w, b = self.conv_out.weight, self.conv_out.bias
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version)
        cache = self.__dict__.get('_conv_out_pad')
        if cache is None or cache[0] != key:
            pad = (-w.shape[0]) % 4
            cache = (key, torch.nn.functional.pad(w.detach().flatten(1), (0, 0, 0, pad)).contiguous(),
                     torch.nn.functional.pad(b.detach(), (0, pad)).contiguous())
            self.__dict__['_conv_out_pad'] = cache
        return cache[1], cache[2]

    def forward(self, seed_points, seed_feats):
        """seed_points (B,N,3), seed_feats (B,C,N) -> vote_points (B,N*vps,3),
        vote_feats (B,C,N*vps), offset (B,3,N*vps)."""
        if self.num_points != -1:
            assert self.num_points < seed_points.shape[1], \
                f'Number of vote points ({self.num_points}) should be smaller than seed ' \
                f'points size ({seed_points.shape[1]})'
            seed_points = seed_points[:, :self.num_points]
            seed_feats = seed_feats[..., :self.num_points]
        batch_size, feat_channels, num_seed = seed_feats.shape
        num_vote = num_seed * self.vote_per_seed
        seed_rows = as_rows(seed_feats)                       # (B,N,C)
        x = seed_rows.reshape(batch_size * num_seed, feat_channels)
        for layer in self.vote_conv:
            x = conv_module_rows(layer, x)
        if self.fused_eval and not torch.is_grad_enabled() and x.is_cuda and self.vote_per_seed == 1 \
                and self.with_res_feat and feat_channels % 128 == 0 and feat_channels <= 512 \
                and x.dtype == torch.float32:
            # inference: conv_out as ONE aligned GEMM (3+C outputs padded to a multiple of 4 so that the
            # library picks its tcgen05 kernel instead of the unaligned sm80 fallback), then offset clamp,
            # seed + offset, seed_feat + residual and the L2 normalisation in one launch
            w, b = self._padded_conv_out()
            votes = torch.addmm(b, x, w.t())
            seed_xyz = seed_points.contiguous()
            seed_c = seed_rows.contiguous()
            vote_points, offset, vote_rows = P.vote_tail(votes, seed_xyz, seed_c, self.vote_xyz_range,
                                                         self.norm_feats)
            return vote_points, vote_rows.transpose(2, 1), offset.transpose(2, 1)
        votes = linear_rows(x, self.conv_out.weight.flatten(1), self.conv_out.bias)
        votes = votes.reshape(batch_size, num_seed, self.vote_per_seed, -1)
        offset = votes[:, :, :, 0:3]
        if self.vote_xyz_range is not None:
            limited = []
            for axis in range(len(self.vote_xyz_range)):
                limited.append(offset[..., axis].clamp(min=-self.vote_xyz_range[axis],
                                                       max=self.vote_xyz_range[axis]))
            offset = torch.stack(limited, -1)
        vote_points = (seed_points.unsqueeze(2) + offset).contiguous()
        vote_points = vote_points.view(batch_size, num_vote, 3)
        offset = offset.reshape(batch_size, num_vote, 3).transpose(2, 1)
        if self.with_res_feat:
            res_feats = votes[:, :, :, 3:]
            vote_rows = (seed_rows.unsqueeze(2) + res_feats).contiguous()
            vote_rows = vote_rows.view(batch_size, num_vote, feat_channels)
            if self.norm_feats:
                features_norm = torch.norm(vote_rows, p=2, dim=2)
                vote_rows = vote_rows.div(features_norm.unsqueeze(2))
            vote_feats = vote_rows.transpose(2, 1)
        else:
            vote_feats = seed_feats
        return vote_points, vote_feats, offset

    def get_loss(self, seed_points, vote_points, seed_indices, vote_targets_mask, vote_targets):
        """Chamfer-style vote loss: min over the gt_per_seed candidate votes, masked."""
        batch_size, num_seed = seed_points.shape[:2]
        seed_gt_votes_mask = torch.gather(vote_targets_mask, 1, seed_indices).float()
        seed_indices_expand = seed_indices.unsqueeze(-1).repeat(1, 1, 3 * self.gt_per_seed)
        seed_gt_votes = torch.gather(vote_targets, 1, seed_indices_expand)
        seed_gt_votes = seed_gt_votes + seed_points.repeat(1, 1, self.gt_per_seed)
        weight = seed_gt_votes_mask / (torch.sum(seed_gt_votes_mask) + 1e-6)
        distance = self.vote_loss(
            vote_points.view(batch_size * num_seed, -1, 3),
            seed_gt_votes.view(batch_size * num_seed, -1, 3),
            dst_weight=weight.view(batch_size * num_seed, 1))[1]
        return torch.sum(torch.min(distance, dim=1)[0])


# ------------------------------------------------------------- prediction head ---------
class BaseConvBboxHead(BaseModule):
    """mmdet3d BaseConvBboxHead: shared k=1 convs, then a classification and a regression conv."""

    def __init__(self, in_channels=0, shared_conv_channels=(), cls_conv_channels=(),
                 num_cls_out_channels=0, reg_conv_channels=(), num_reg_out_channels=0,
                 conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'),
                 act_cfg=dict(type='ReLU'), bias='auto', init_cfg=None, *args, **kwargs):
        super().__init__(init_cfg=init_cfg)
        assert in_channels > 0
        assert num_cls_out_channels > 0
        assert num_reg_out_channels > 0
        self.in_channels = in_channels
        self.shared_conv_channels = shared_conv_channels
        self.cls_conv_channels = cls_conv_channels
        self.num_cls_out_channels = num_cls_out_channels
        self.reg_conv_channels = reg_conv_channels
        self.num_reg_out_channels = num_reg_out_channels
        self.conv_cfg, self.norm_cfg, self.act_cfg, self.bias = conv_cfg, norm_cfg, act_cfg, bias
        if len(self.shared_conv_channels) > 0:
            self.shared_convs = self._add_conv_branch(self.in_channels, self.shared_conv_channels)
            out_channels = self.shared_conv_channels[-1]
        else:
            out_channels = self.in_channels
        prev_channel = out_channels
        if len(self.cls_conv_channels) > 0:
            self.cls_convs = self._add_conv_branch(prev_channel, self.cls_conv_channels)
            prev_channel = self.cls_conv_channels[-1]
        self.conv_cls = build_conv_layer(conv_cfg, in_channels=prev_channel,
                                         out_channels=num_cls_out_channels, kernel_size=1)
        prev_channel = out_channels
        if len(self.reg_conv_channels) > 0:
            self.reg_convs = self._add_conv_branch(prev_channel, self.reg_conv_channels)
            prev_channel = self.reg_conv_channels[-1]
        self.conv_reg = build_conv_layer(conv_cfg, in_channels=prev_channel,
                                         out_channels=num_reg_out_channels, kernel_size=1)

    def _add_conv_branch(self, in_channels, conv_channels):
        conv_spec = [in_channels] + list(conv_channels)
        conv_layers = nn.Sequential()
        for i in range(len(conv_spec) - 1):
            conv_layers.add_module(
                f'layer{i}',
                ConvModule(conv_spec[i], conv_spec[i + 1], kernel_size=1, padding=0,
                           conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg, act_cfg=self.act_cfg,
                           bias=self.bias, inplace=True))
        return conv_layers

    def _merged_out(self):
        """[conv_cls; conv_reg] as one zero-padded (W, b), rebuilt when a parameter changes."""
        ts = (self.conv_cls.weight, self.conv_cls.bias, self.conv_reg.weight, self.conv_reg.bias)
        key = tuple((t.data_ptr(), t._version) for t in ts)
        cache = self.__dict__.get("_merged_out_cache")
        if cache is None or cache[0] != key:
            w = torch.cat([ts[0].detach().flatten(1), ts[2].detach().flatten(1)], 0)
            b = torch.cat([ts[1].detach(), ts[3].detach()], 0)
            pad = (-w.shape[0]) % 4
            cache = (key, torch.nn.functional.pad(w, (0, 0, 0, pad)).contiguous(),
                     torch.nn.functional.pad(b, (0, pad)).contiguous(), ts[0].shape[0], ts[2].shape[0])
            self.__dict__["_merged_out_cache"] = cache
        return cache[1:]

    def forward(self, feats):
        """feats (B,C,N) -> cls_score (B,num_cls,N), bbox_pred (B,num_reg,N)."""
        rows = as_rows(feats)
        B, N, C = rows.shape
        x = rows.reshape(B * N, C)
        if len(self.shared_conv_channels) > 0:
            for layer in self.shared_convs:
                x = conv_module_rows(layer, x)
        x_cls = x_reg = x
        if not torch.is_grad_enabled() and x.is_cuda and not self.cls_conv_channels and not self.reg_conv_channels:
            # inference: both output convs read the same rows -> ONE GEMM over the concatenated weights,
            # padded to a multiple of 4 outputs so that the library takes an aligned tcgen05 kernel
            w, b, n_cls, n_reg = self._merged_out()
            out = torch.addmm(b, x, w.t()).view(B, N, -1)
            return out[..., :n_cls].transpose(1, 2), out[..., n_cls:n_cls + n_reg].transpose(1, 2)
        if len(self.cls_conv_channels) > 0:
            for layer in self.cls_convs:
                x_cls = conv_module_rows(layer, x_cls)
        cls_score = linear_rows(x_cls, self.conv_cls.weight.flatten(1), self.conv_cls.bias)
        if len(self.reg_conv_channels) > 0:
            for layer in self.reg_convs:
                x_reg = conv_module_rows(layer, x_reg)
        bbox_pred = linear_rows(x_reg, self.conv_reg.weight.flatten(1), self.conv_reg.bias)
        return cls_score.reshape(B, N, -1).transpose(1, 2), bbox_pred.reshape(B, N, -1).transpose(1, 2)
