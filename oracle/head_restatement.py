"""Statement-by-statement CPU restatement of DeMFVoteHead.forward / transformer_decoder and of the
decoder layer it calls.  TEST INFRASTRUCTURE ONLY (imported by tests/ alone). PARITY UNPINNED: the
reference ships no fixtures for this path; the upstream module semantics restated here are those of
mmdet3d 0.18.1 / mmcv-full 1.3.18 / mmdet 2.14.0.

Follows, in the reference's own (B,C,N) channel-major layout and statement order, with plain
torch.nn.functional calls on the weights of a state dict -- none of demf_b200's nn.Modules:
  * demf/modeling/heads/class_agnostic_vote_head.py:405-466  forward (sample_mod='seed')
  * :468-512  transformer_decoder, :514-522 get_valid_ratio, :524-547 get_reference_points,
    :549-594 prepare_decoder_inputs
  * demf/modeling/layers/transformer.py:18-36 PositionEmbeddingLearned, :55-80 decoder layer forward
  * demf/core/bbox/coders/class_agnostic_bbox_coder.py:196-240 split_pred
  * [UPSTREAM] VoteModule.forward, PointSAModule.forward (QueryAndGroup + Conv2d/BN2d/ReLU + max_pool2d),
    BaseConvBboxHead.forward, BaseTransformerLayer.forward with operation_order
    (self_attn, norm, cross_attn, norm, ffn, norm), MultiheadAttention, MultiScaleDeformableAttention,
    FFN
so a wiring error in demf_b200/modeling/{heads,layers}.py or demf_b200/mm/{bricks,pointnet_modules,
ms_deform_attn}.py cannot hide behind "the same module on CPU".
Index ops (FPS, ball query) come from the C oracle; the attention core is mmcv's own CPU formulation
(oracle/msda_torch.py). Eval mode: BatchNorm uses running statistics, dropout is the identity.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cref
from .msda_torch import multi_scale_deformable_attn_pytorch


def _conv_bn_relu(x, sd, prefix, bn='bn', eps=1e-5):
    """mmcv ConvModule with kernel size 1 in eval mode: conv (1d or 2d by the weight's rank) -> BN -> ReLU."""
    w = sd[prefix + '.conv.weight']
    conv = F.conv1d if w.dim() == 3 else F.conv2d
    y = conv(x, w, sd.get(prefix + '.conv.bias'))
    y = F.batch_norm(y, sd[f'{prefix}.{bn}.running_mean'], sd[f'{prefix}.{bn}.running_var'],
                     sd[f'{prefix}.{bn}.weight'], sd[f'{prefix}.{bn}.bias'], False, 0.1, eps)
    return F.relu(y)


def vote_module(sd, seed_points, seed_feats, vote_per_seed=1):
    """[UPSTREAM] mmdet3d VoteModule.forward (norm_feats=True, with_res_feat=True, no xyz range)."""
    batch_size, feat_channels, num_seed = seed_feats.shape
    num_vote = num_seed * vote_per_seed
    x = _conv_bn_relu(seed_feats, sd, 'vote_module.vote_conv.0')
    x = _conv_bn_relu(x, sd, 'vote_module.vote_conv.1')
    votes = F.conv1d(x, sd['vote_module.conv_out.weight'], sd['vote_module.conv_out.bias'])
    votes = votes.transpose(2, 1).view(batch_size, num_seed, vote_per_seed, -1)
    offset = votes[:, :, :, 0:3]
    vote_points = (seed_points.unsqueeze(2) + offset).contiguous()
    vote_points = vote_points.view(batch_size, num_vote, 3)
    offset = offset.reshape(batch_size, num_vote, 3).transpose(2, 1)
    res_feats = votes[:, :, :, 3:]
    vote_feats = (seed_feats.transpose(2, 1).unsqueeze(2) + res_feats).contiguous()
    vote_feats = vote_feats.view(batch_size, num_vote, feat_channels).transpose(2, 1).contiguous()
    features_norm = torch.norm(vote_feats, p=2, dim=1)
    vote_feats = vote_feats.div(features_norm.unsqueeze(1))
    return vote_points, vote_feats, offset


def _group(features, idx):
    """[UPSTREAM] grouping_operation: features (B,C,N), idx (B,M,ns) -> (B,C,M,ns)."""
    B, C, _ = features.shape
    _, M, ns = idx.shape
    flat = idx.reshape(B, 1, M * ns).long().expand(-1, C, -1)
    return torch.gather(features, 2, flat).view(B, C, M, ns)


def sa_module(sd, prefix, points_xyz, features, indices, radius, nsample):
    """[UPSTREAM] PointSAModule.forward with given `indices` (use_xyz, normalize_xyz, max pool):
    new_xyz = gather(xyz, indices); QueryAndGroup; 3x (Conv2d 1x1 no bias + BN2d + ReLU); max_pool2d."""
    xyz_flipped = points_xyz.transpose(1, 2).contiguous()
    new_xyz = torch.gather(xyz_flipped, 2, indices.long().unsqueeze(1).expand(-1, 3, -1)).transpose(1, 2).contiguous()
    idx = cref.ball_query(0.0, radius, nsample, points_xyz.contiguous(), new_xyz)
    grouped_xyz = _group(xyz_flipped, idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
    grouped_xyz = grouped_xyz / radius
    new_features = torch.cat([grouped_xyz, _group(features, idx)], dim=1)
    for j in range(3):
        new_features = _conv_bn_relu(new_features, sd, f'{prefix}.mlps.0.layer{j}')
    new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(-1)
    return new_xyz, new_features, indices


def conv_pred(sd, prefix, feats):
    """[UPSTREAM] BaseConvBboxHead.forward: shared convs -> conv_cls / conv_reg."""
    x = _conv_bn_relu(feats, sd, f'{prefix}.shared_convs.layer0')
    x = _conv_bn_relu(x, sd, f'{prefix}.shared_convs.layer1')
    cls_score = F.conv1d(x, sd[f'{prefix}.conv_cls.weight'], sd[f'{prefix}.conv_cls.bias'])
    bbox_pred = F.conv1d(x, sd[f'{prefix}.conv_reg.weight'], sd[f'{prefix}.conv_reg.bias'])
    return cls_score, bbox_pred


def split_pred(cls_preds, reg_preds, base_xyz, num_dir_bins=12):
    """class_agnostic_bbox_coder.py:196-240."""
    results = {}
    cls_preds_trans = cls_preds.transpose(2, 1)
    reg_preds_trans = reg_preds.transpose(2, 1)
    results['center'] = base_xyz + reg_preds_trans[..., 0:3].contiguous()
    results['size'] = reg_preds_trans[..., 3:6].contiguous()
    results['dir_class'] = reg_preds_trans[..., 6:6 + num_dir_bins].contiguous()
    dir_res_norm = reg_preds_trans[..., 6 + num_dir_bins:6 + 2 * num_dir_bins].contiguous()
    results['dir_res_norm'] = dir_res_norm
    results['dir_res'] = dir_res_norm * (np.pi / num_dir_bins)
    results['obj_scores'] = cls_preds_trans[..., 0:2].contiguous()
    if cls_preds_trans.shape[-1] > 2:
        results['sem_scores'] = cls_preds_trans[..., 2:].contiguous()
    return results


# --- [UPSTREAM] coordinate helpers, per scene ------------------------------------
def _apply_3d_transformation_reverse(pcd, meta):
    rot = torch.as_tensor(np.asarray(meta.get('pcd_rotation', np.eye(3))), dtype=pcd.dtype)
    scale = meta.get('pcd_scale_factor', 1.0)
    trans = torch.as_tensor(np.asarray(meta.get('pcd_trans', np.zeros(3))), dtype=pcd.dtype)
    pcd = pcd.clone()
    for op in list(meta.get('transformation_3d_flow', []))[::-1]:
        if op == 'T':
            pcd = pcd - trans
        elif op == 'S':
            pcd = pcd * (1.0 / scale)
        elif op == 'R':
            pcd = pcd @ rot.inverse()
        elif op == 'HF':
            if meta.get('pcd_horizontal_flip', False):
                pcd[:, 0] = -pcd[:, 0]
        elif op == 'VF':
            if meta.get('pcd_vertical_flip', False):
                pcd[:, 1] = -pcd[:, 1]
        else:
            raise AssertionError(op)
    return pcd


def get_reference_points(seeds_3d_batch, img_metas):
    """class_agnostic_vote_head.py:524-547."""
    uv_all = []
    for seeds_3d, meta in zip(seeds_3d_batch, img_metas):
        img_shape = meta['img_shape']
        xyz_depth = _apply_3d_transformation_reverse(seeds_3d, meta)
        depth2img = torch.as_tensor(np.asarray(meta['depth2img']), dtype=xyz_depth.dtype)
        proj = torch.eye(4, dtype=xyz_depth.dtype)
        proj[:depth2img.shape[0], :depth2img.shape[1]] = depth2img
        p4 = torch.cat([xyz_depth, xyz_depth.new_ones(xyz_depth.shape[0], 1)], -1) @ proj.T
        uv = p4[:, :2] / p4[:, 2:3]
        sf = meta['scale_factor'][:2] if 'scale_factor' in meta else [1.0, 1.0]
        crop = meta.get('img_crop_offset', [0.0, 0.0])
        uv = uv.clone()
        uv[:, 0] = uv[:, 0] * float(sf[0]) + float(crop[0])
        uv[:, 1] = uv[:, 1] * float(sf[1]) + float(crop[1])
        if meta.get('flip', False):
            uv[:, 0] = img_shape[1] - uv[:, 0]
        uv[:, 0] = uv[:, 0] / (img_shape[1] - 1)
        uv[:, 1] = uv[:, 1] / (img_shape[0] - 1)
        uv_all.append(torch.clamp(uv, 0, 1))
    return torch.stack(uv_all, dim=0)


def get_valid_ratio(mask):
    _, H, W = mask.shape
    valid_H = torch.sum(~mask[:, :, 0], 1)
    valid_W = torch.sum(~mask[:, 0, :], 1)
    return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)


def prepare_decoder_inputs(seeds_3d, mlvl_feats, img_metas):
    """class_agnostic_vote_head.py:549-594."""
    reference_points = get_reference_points(seeds_3d, img_metas)
    batch_size = mlvl_feats[0].size(0)
    input_img_h, input_img_w = img_metas[0]['batch_input_shape']
    img_masks = mlvl_feats[0].new_ones((batch_size, input_img_h, input_img_w))
    for img_id in range(batch_size):
        img_h, img_w = img_metas[img_id]['img_shape'][:2]
        img_masks[img_id, :img_h, :img_w] = 0
    mlvl_masks = [F.interpolate(img_masks[None], size=feat.shape[-2:]).to(torch.bool).squeeze(0)
                  for feat in mlvl_feats]
    feat_flatten, mask_flatten, spatial_shapes = [], [], []
    for feat, mask in zip(mlvl_feats, mlvl_masks):
        spatial_shapes.append(tuple(feat.shape[-2:]))
        feat_flatten.append(feat.flatten(2).transpose(1, 2))
        mask_flatten.append(mask.flatten(1))
    feat_flatten = torch.cat(feat_flatten, 1)
    mask_flatten = torch.cat(mask_flatten, 1)
    spatial_shapes = torch.as_tensor(spatial_shapes, dtype=torch.long)
    level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
    valid_ratios = torch.stack([get_valid_ratio(m) for m in mlvl_masks], 1)
    return feat_flatten.permute(1, 0, 2), mask_flatten, reference_points, spatial_shapes, \
        level_start_index, valid_ratios


# --- [UPSTREAM] transformer bricks -----------------------------------------------------------------
def multihead_self_attention(sd, prefix, query, query_pos, num_heads=8):
    """mmcv MultiheadAttention as BaseTransformerLayer calls it for 'self_attn': q = k = query + query_pos,
    v = query, nn.MultiheadAttention math written out, identity + dropout(out) (eval)."""
    Lq, B, E = query.shape
    hd = E // num_heads
    qk = query + query_pos
    w, b = sd[prefix + '.attn.in_proj_weight'], sd[prefix + '.attn.in_proj_bias']
    q = F.linear(qk, w[:E], b[:E])
    k = F.linear(qk, w[E:2 * E], b[E:2 * E])
    v = F.linear(query, w[2 * E:], b[2 * E:])
    q = q.contiguous().view(Lq, B * num_heads, hd).transpose(0, 1) * (float(hd) ** -0.5)
    k = k.contiguous().view(Lq, B * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(Lq, B * num_heads, hd).transpose(0, 1)
    attn = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).contiguous().view(Lq, B, E)
    out = F.linear(out, sd[prefix + '.attn.out_proj.weight'], sd[prefix + '.attn.out_proj.bias'])
    return query + out


def ms_deformable_cross_attention(sd, prefix, query, value, query_pos, key_padding_mask, reference_points,
                                  spatial_shapes, level_start_index, num_heads=8, num_levels=4, num_points=4):
    """mmcv MultiScaleDeformableAttention.forward (batch_first=False), multi_scale_deform_attn.py:290-360."""
    identity = query
    query = query + query_pos
    query = query.permute(1, 0, 2)
    value = value.permute(1, 0, 2)
    bs, num_query, _ = query.shape
    bs, num_value, _ = value.shape
    assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum()) == num_value
    value = F.linear(value, sd[prefix + '.value_proj.weight'], sd[prefix + '.value_proj.bias'])
    if key_padding_mask is not None:
        value = value.masked_fill(key_padding_mask[..., None], 0.0)
    value = value.view(bs, num_value, num_heads, -1)
    sampling_offsets = F.linear(query, sd[prefix + '.sampling_offsets.weight'],
                                sd[prefix + '.sampling_offsets.bias']).view(
        bs, num_query, num_heads, num_levels, num_points, 2)
    attention_weights = F.linear(query, sd[prefix + '.attention_weights.weight'],
                                 sd[prefix + '.attention_weights.bias']).view(
        bs, num_query, num_heads, num_levels * num_points)
    attention_weights = attention_weights.softmax(-1).view(bs, num_query, num_heads, num_levels, num_points)
    assert reference_points.shape[-1] == 2
    offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
    sampling_locations = reference_points[:, :, None, :, None, :] \
        + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
    output = multi_scale_deformable_attn_pytorch(value, spatial_shapes, sampling_locations, attention_weights)
    output = F.linear(output, sd[prefix + '.output_proj.weight'], sd[prefix + '.output_proj.bias'])
    return output.permute(1, 0, 2) + identity


def decoder_layer(sd, prefix, query, query_pos, value, key_padding_mask, reference_points, spatial_shapes,
                  level_start_index, valid_ratios, num_points):
    """demf/modeling/layers/transformer.py:55-80 + [UPSTREAM] BaseTransformerLayer.forward."""
    assert reference_points.shape[-1] == 2
    reference_points_input = reference_points[:, :, None] * valid_ratios[:, None]
    pe = prefix + '.posembed.position_embedding_head'
    xyz = query_pos.transpose(1, 2).contiguous()
    x = F.conv1d(xyz, sd[pe + '.0.weight'], sd[pe + '.0.bias'])
    x = F.batch_norm(x, sd[pe + '.1.running_mean'], sd[pe + '.1.running_var'], sd[pe + '.1.weight'],
                     sd[pe + '.1.bias'], False, 0.1, 1e-5)
    x = F.conv1d(F.relu(x), sd[pe + '.3.weight'], sd[pe + '.3.bias'])
    query_pos_embed = x.permute(2, 0, 1)

    lp = prefix + '.layer'
    E = query.shape[-1]

    def norm(x, i):
        return F.layer_norm(x, (E,), sd[f'{lp}.norms.{i}.weight'], sd[f'{lp}.norms.{i}.bias'], 1e-5)

    query = multihead_self_attention(sd, lp + '.attentions.0', query, query_pos_embed)
    query = norm(query, 0)
    query = ms_deformable_cross_attention(sd, lp + '.attentions.1', query, value, query_pos_embed,
                                          key_padding_mask, reference_points_input, spatial_shapes,
                                          level_start_index, num_points=num_points)
    query = norm(query, 1)
    ffn = F.linear(F.relu(F.linear(query, sd[f'{lp}.ffns.0.layers.0.0.weight'], sd[f'{lp}.ffns.0.layers.0.0.bias'])),
                   sd[f'{lp}.ffns.0.layers.1.weight'], sd[f'{lp}.ffns.0.layers.1.bias'])
    query = query + ffn
    return norm(query, 2)


def head_forward(sd, seed_points, seed_features, seed_indices, img_features, img_metas, num_proposal=256,
                 num_points=4, num_decoder_layers=1, radius=0.3, nsample=16):
    """class_agnostic_vote_head.py:405-512 with sample_mod='seed'. `sd` = state dict of the head
    (CPU tensors)."""
    vote_points, vote_features, vote_offset = vote_module(sd, seed_points, seed_features)
    results = dict(seed_points=seed_points, seed_indices=seed_indices, vote_points=vote_points,
                   vote_features=vote_features, vote_offset=vote_offset)
    sample_indices = cref.furthest_point_sample(seed_points.contiguous(), num_proposal)
    aggregated_points, features, aggregated_indices = sa_module(
        sd, 'vote_aggregation', vote_points, vote_features, sample_indices, radius, nsample)
    results['aggregated_points'] = aggregated_points
    results['aggregated_indices'] = aggregated_indices

    decode_res_all = []
    cls_predictions, reg_predictions = conv_pred(sd, 'conv_pred0', features)
    decode_res = split_pred(cls_predictions, reg_predictions, aggregated_points)
    decode_res_all.append(decode_res)
    feat_flatten, mask_flatten, reference_points, spatial_shapes, level_start_index, valid_ratios = \
        prepare_decoder_inputs(aggregated_points, img_features, img_metas)
    query = features.permute(2, 0, 1)
    for i in range(num_decoder_layers):
        query_pos = torch.cat([decode_res['center'], decode_res['size']], dim=-1).detach().clone()
        query = decoder_layer(sd, f'decoder.{i}', query, query_pos, feat_flatten, mask_flatten, reference_points,
                              spatial_shapes, level_start_index, valid_ratios, num_points)
        cls_predictions, reg_predictions = conv_pred(sd, f'conv_pred{i + 1}', query.permute(1, 2, 0))
        decode_res = split_pred(cls_predictions, reg_predictions, aggregated_points)
        decode_res_all.append(decode_res)
    results['decode_res_all'] = decode_res_all
    results['_reference_points'] = reference_points
    results['_query'] = query
    return results
