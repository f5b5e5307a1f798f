"""Decoder layer of the DeMF fusion stage (reference: demf/modeling/layers/transformer.py:18-80):
a learned 6-d (centre, size) positional embedding plus a DETR decoder layer whose cross
attention is multi-scale deformable attention over the image pyramid."""
import torch
import torch.nn as nn

from ..mm.ms_deform_attn import MultiScaleDeformableAttention
from ..mm.registry import TRANSFORMER_LAYER, build_transformer_layer


class PositionEmbeddingLearned(nn.Module):
    """(B,Q,input_channel) -> (B,num_pos_feats,Q): Conv1d-BN-ReLU-Conv1d (transformer.py:18-36)."""

    def __init__(self, cfg):
        super().__init__()
        input_channel = cfg['input_channel']
        num_pos_feats = cfg['num_pos_feats']
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
            nn.BatchNorm1d(num_pos_feats),
            nn.ReLU(inplace=True),
            nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def _folded(self, conv, bn):
        """(W', b') of conv followed by eval-mode BatchNorm, rebuilt only when a tensor changes."""
        tensors = (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        cache = self.__dict__.get("_fold_cache")
        if cache is None or cache[0] != key:
            scale = torch.rsqrt(bn.running_var + bn.eps) * bn.weight
            w = (conv.weight.flatten(1) * scale.unsqueeze(1)).contiguous()
            b = ((conv.bias - bn.running_mean) * scale + bn.bias).contiguous()
            cache = (key, w, b)
            self.__dict__["_fold_cache"] = cache
        return cache[1], cache[2]

    def forward(self, xyz):
        conv1, bn, _, conv2 = self.position_embedding_head
        if not torch.is_grad_enabled() and not bn.training and bn.track_running_stats:
            # inference: both kernel-size-1 convolutions are GEMMs over the (B*Q, 6) rows, the
            # BatchNorm folds into the first one -> two launches instead of seven
            w1, b1 = self._folded(conv1, bn)
            B, Q, C = xyz.shape
            h = torch._addmm_activation(b1, xyz.reshape(B * Q, C), w1.t())      # + ReLU
            out = torch.addmm(conv2.bias, h, conv2.weight.flatten(1).t())
            return out.view(B, Q, -1).transpose(1, 2)
        xyz = xyz.transpose(1, 2).contiguous()
        return self.position_embedding_head(xyz)


@TRANSFORMER_LAYER.register_module()
class DeMFTransformerDecoderLayer(nn.Module):

    def __init__(self, *args, transformerlayers=None, posembed=None, **kwargs):
        super().__init__()
        self.layer = build_transformer_layer(transformerlayers)
        self.posembed = PositionEmbeddingLearned(posembed)

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MultiScaleDeformableAttention):
                m.init_weights()

    def forward(self, query, query_pos, *args, reference_points=None, valid_ratios=None, **kwargs):
        """query (Q,B,C); query_pos (B,Q,6); reference_points (B,Q,2|4) in [0,1];
        valid_ratios (B,L,2). Returns (Q,B,C)."""
        if reference_points.shape[-1] == 4:
            reference_points_input = reference_points[:, :, None] * \
                torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        else:
            assert reference_points.shape[-1] == 2
            reference_points_input = reference_points[:, :, None] * valid_ratios[:, None]
        query_pos_embed = self.posembed(query_pos).permute(2, 0, 1)
        return self.layer(query, *args, query_pos=query_pos_embed,
                          reference_points=reference_points_input, **kwargs)
