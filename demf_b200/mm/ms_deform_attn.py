"""`mmcv.ops.multi_scale_deform_attn` surface over the sm_100a kernels.

Stands in for `from mmcv.ops.multi_scale_deform_attn import MultiScaleDeformableAttention`
(reference: demf/modeling/layers/transformer.py:8-15; cfg configs/demf/demf_votenet.py:79-85).
`MultiScaleDeformableAttnFunction.apply(value, spatial_shapes, level_start_index,
sampling_locations, attention_weights, im2col_step)` has upstream's signature; the module keeps
upstream's parameter names (`sampling_offsets`, `attention_weights`, `value_proj`,
`output_proj`) and initialisation. The four Linear layers are library GEMMs; the sampling core
is demf_msda_fwd / demf_msda_bwd. CUDA only.
"""
import math

import torch
import torch.nn as nn
from torch.autograd.function import Function, once_differentiable

from .. import _lib
from .bricks import BaseModule, linear_nd
from .registry import ATTENTION


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _KernelTimer:
    """Optional CUDA-event bracket around each forward sampling-kernel launch (bench.py's
    roofline figure is measured with it, live, on the launch stream)."""

    def __init__(self):
        self.enabled = False
        self.pairs = []

    def enable(self):
        self.enabled, self.pairs = True, []

    def disable(self):
        self.enabled = False

    def drain(self):
        torch.cuda.synchronize()
        out = [a.elapsed_time(b) for a, b in self.pairs]
        self.pairs = []
        return out

    def read_last(self):
        """Elapsed ms of the most recent pair (re-recorded by every replay of a captured graph);
        the caller synchronises first."""
        if not self.pairs:
            return None
        a, b = self.pairs[-1]
        return a.elapsed_time(b)


KERNEL_TIMER = _KernelTimer()


class MultiScaleDeformableAttnFunction(Function):

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        tensors = (value, value_spatial_shapes, value_level_start_index, sampling_locations,
                   attention_weights)
        names = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
        for t, n in zip(tensors, names):
            if not t.is_contiguous():
                raise RuntimeError(f"{n} tensor has to be contiguous")
            if not t.is_cuda:
                raise RuntimeError(f"{n} must be a CUDA tensor (no CPU fallback in demf_b200)")
        if value.dtype != torch.float32:
            raise RuntimeError("ms_deform_attn: only float32 sampling is implemented")
        B, S, H, D = value.shape
        _, Q, _, L, P, _ = sampling_locations.shape
        step = min(B, im2col_step)
        if step <= 0 or B % step != 0:
            raise RuntimeError(f"batch({B}) must divide im2col_step({step})")
        shapes = value_spatial_shapes.to(torch.int64)
        lsi = value_level_start_index.to(torch.int64)
        ctx.im2col_step = im2col_step
        with torch.cuda.device_of(value):
            out = torch.empty(B, Q, H * D, dtype=value.dtype, device=value.device)
            if KERNEL_TIMER.enabled:
                ext = dict(external=True) if torch.cuda.is_current_stream_capturing() else {}
                ev = (torch.cuda.Event(enable_timing=True, **ext),
                      torch.cuda.Event(enable_timing=True, **ext))
                ev[0].record()
            _lib.check(_lib.load().demf_msda_fwd(
                value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), sampling_locations.data_ptr(),
                attention_weights.data_ptr(), B, S, H, D, Q, L, P, out.data_ptr(), _stream()),
                "demf_msda_fwd")
            if KERNEL_TIMER.enabled:
                ev[1].record()
                KERNEL_TIMER.pairs.append(ev)
        ctx.save_for_backward(value, shapes, lsi, sampling_locations, attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        B, S, H, D = value.shape
        _, Q, _, L, P, _ = loc.shape
        grad_output = grad_output.contiguous()
        with torch.cuda.device_of(value):
            grad_value = torch.zeros_like(value)
            grad_loc = torch.empty_like(loc)
            grad_attn = torch.empty_like(attn)
            _lib.check(_lib.load().demf_msda_bwd(
                value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                grad_output.data_ptr(), B, S, H, D, Q, L, P, grad_value.data_ptr(),
                grad_loc.data_ptr(), grad_attn.data_ptr(), _stream()), "demf_msda_bwd")
        return grad_value, None, None, grad_loc, grad_attn, None


def msda_proj_supported(head_dim, num_levels, num_points):
    return bool(_lib.load().demf_msda_proj_fwd_supported(head_dim, num_levels, num_points))


def msda_from_projections(value, spatial_shapes, level_start_index, proj, reference_points,
                          num_levels, num_points, proj_add=None):
    """Inference MSDA straight from the query projections: softmax of the logits, sampling
    locations and sampling in ONE launch (csrc/msda.cu, kProj).
    value (B,S,H,D); proj (B*Q, H*L*P*3) = [offsets | logits]; reference_points (B,Q,L,2|4);
    proj_add: optional tensor of proj's shape added to it inside the kernel."""
    B, S, H, D = value.shape
    Q = reference_points.shape[1]
    assert value.is_cuda and value.dtype == torch.float32 and value.is_contiguous()
    assert proj.is_contiguous() and proj.shape == (B * Q, H * num_levels * num_points * 3)
    assert reference_points.is_contiguous() and reference_points.dtype == torch.float32
    assert reference_points.shape[2] == num_levels
    assert proj_add is None or (proj_add.is_contiguous() and proj_add.shape == proj.shape
                                and proj_add.dtype == proj.dtype)
    out = torch.empty(B, Q, H * D, dtype=value.dtype, device=value.device)
    with torch.cuda.device_of(value):
        if KERNEL_TIMER.enabled:
            ext = dict(external=True) if torch.cuda.is_current_stream_capturing() else {}
            ev = (torch.cuda.Event(enable_timing=True, **ext), torch.cuda.Event(enable_timing=True, **ext))
            ev[0].record()
        _lib.check(_lib.load().demf_msda_proj_fwd(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), proj.data_ptr(),
            proj_add.data_ptr() if proj_add is not None else None, reference_points.data_ptr(), reference_points.shape[-1], B, S, H, D, Q, num_levels,
            num_points, out.data_ptr(), _stream()), "demf_msda_proj_fwd")
        if KERNEL_TIMER.enabled:
            ev[1].record()
            KERNEL_TIMER.pairs.append(ev)
    return out


_validated_shapes = {}


def _check_num_value(spatial_shapes, num_value):
    """`(h*w).sum() == num_value` (upstream asserts it every call, which is a device sync; here a
    given shapes tensor is validated once)."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, int(num_value), str(spatial_shapes.device))
    if key not in _validated_shapes:
        assert (spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum() == num_value
        if len(_validated_shapes) > 256:
            _validated_shapes.clear()
        _validated_shapes[key] = True


@ATTENTION.register_module()
class MultiScaleDeformableAttention(BaseModule):
    """mmcv MultiScaleDeformableAttention (Deformable-DETR attention), same ctor and forward."""

    fused_eval = True   # inference: projections -> output in one launch (msda_from_projections)

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} and "
                             f"{num_heads}")
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)
        self._is_init = True

    def _fused_query_proj(self):
        """[sampling_offsets; attention_weights] as one (W, b), rebuilt when a parameter changes."""
        ts = (self.sampling_offsets.weight, self.sampling_offsets.bias, self.attention_weights.weight,
              self.attention_weights.bias)
        key = tuple((t.data_ptr(), t._version) for t in ts)
        cache = self.__dict__.get("_qproj_cache")
        if cache is None or cache[0] != key:
            cache = (key, torch.cat([ts[0], ts[2]], 0).detach().contiguous(),
                     torch.cat([ts[1], ts[3]], 0).detach().contiguous())
            self.__dict__["_qproj_cache"] = cache
        return cache[1], cache[2]

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        _check_num_value(spatial_shapes, num_value)

        value = linear_nd(value, self.value_proj)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        if not torch.is_grad_enabled():
            # inference: the offset and weight projections read the same rows -- one GEMM over the
            # concatenated (cached) weights on a contiguous copy of the permuted query
            w, b = self._fused_query_proj()
            proj = torch.addmm(b, query.reshape(bs * num_query, -1), w.t())
            if self.fused_eval and value.is_cuda and value.dtype == torch.float32 \
                    and reference_points.shape[-1] in (2, 4) \
                    and msda_proj_supported(value.shape[-1], self.num_levels, self.num_points):
                # softmax, sampling locations and sampling in one launch
                output = msda_from_projections(
                    value.contiguous(), spatial_shapes, level_start_index, proj,
                    reference_points.contiguous().float(), self.num_levels, self.num_points)
                output = self.output_proj(output.to(query.dtype))
                if not self.batch_first:
                    output = output.permute(1, 0, 2)
                return self.dropout(output) + identity
            n_off = self.sampling_offsets.out_features
            sampling_offsets = proj[:, :n_off].view(
                bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
            attention_weights = proj[:, n_off:].view(
                bs, num_query, self.num_heads, self.num_levels * self.num_points)
        else:
            sampling_offsets = linear_nd(query, self.sampling_offsets).view(
                bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
            attention_weights = linear_nd(query, self.attention_weights).view(
                bs, num_query, self.num_heads, self.num_levels * self.num_points)
        attention_weights = attention_weights.softmax(-1).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.num_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get "
                             f"{reference_points.shape[-1]} instead.")
        output = MultiScaleDeformableAttnFunction.apply(
            value.contiguous().float(), spatial_shapes, level_start_index,
            sampling_locations.contiguous().float(), attention_weights.contiguous().float(),
            self.im2col_step)
        output = linear_nd(output.to(query.dtype), self.output_proj)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output) + identity
