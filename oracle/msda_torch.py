"""Pure-torch formulation of multi-scale deformable attention.

TEST INFRASTRUCTURE ONLY (see oracle/demf_oracle.c). PARITY UNPINNED.

This restates mmcv 1.3.18 `mmcv.ops.multi_scale_deform_attn.
multi_scale_deformable_attn_pytorch` -- the CPU path of the stack the reference
runs on (requirements.txt:2, imported at demf/modeling/layers/transformer.py:8-15):
per level, reshape the value slab to an image batch, `F.grid_sample(bilinear,
zeros, align_corners=False)` at `2*loc-1`, then a weighted sum over levels*points.
It serves two purposes: (1) an independent second formulation the C oracle is
cross-checked against; (2) the "reference CPU path" timed by bench.py --impl
reference / cpu_baseline for the MSDA step.
"""
import torch
import torch.nn.functional as F


def multi_scale_deformable_attn_pytorch(value, value_spatial_shapes, sampling_locations,
                                        attention_weights):
    """value (B,S,H,D); shapes (L,2) [h,w]; loc (B,Q,H,L,P,2) in [0,1]; w (B,Q,H,L,P) -> (B,Q,H*D)."""
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    sizes = [(int(h), int(w)) for h, w in value_spatial_shapes.tolist()]
    assert sum(h * w for h, w in sizes) == S
    grids = sampling_locations * 2.0 - 1.0  # grid_sample convention, x first
    per_level = []
    start = 0
    for lvl, (h, w) in enumerate(sizes):
        slab = value[:, start:start + h * w]  # (B, hw, H, D)
        start += h * w
        img = slab.permute(0, 2, 3, 1).reshape(B * H, D, h, w)
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * H, Q, P, 2)
        per_level.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))  # (B*H, D, Q, P)
    sampled = torch.stack(per_level, dim=3).reshape(B * H, D, Q, L * P)
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(B * H, 1, Q, L * P)
    out = (sampled * wts).sum(-1)  # (B*H, D, Q)
    return out.reshape(B, H * D, Q).transpose(1, 2).contiguous()
