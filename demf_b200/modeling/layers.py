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

    def rows(self, xyz):
        """Inference: (B,Q,6) -> (B*Q, num_pos_feats) rows, two GEMMs (BatchNorm folded into the first)."""
        conv1, bn, _, conv2 = self.position_embedding_head
        w1, b1 = self._folded(conv1, bn)
        B, Q, C = xyz.shape
        h = torch._addmm_activation(b1, xyz.reshape(B * Q, C), w1.t())      # + ReLU
        return torch.addmm(conv2.bias, h, conv2.weight.flatten(1).t())

    def forward(self, xyz):
        conv1, bn, _, conv2 = self.position_embedding_head
        if not torch.is_grad_enabled() and not bn.training and bn.track_running_stats:
            # inference: both kernel-size-1 convolutions are GEMMs over the (B*Q, 6) rows, the
            # BatchNorm folds into the first one -> two launches instead of seven
            B, Q, _ = xyz.shape
            return self.rows(xyz).view(B, Q, -1).transpose(1, 2)
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

    # ------------------------------------------------------- inference on query rows ----
    fused_eval = True

    def _rows_path_ok(self, query, value, kwargs):
        from ..mm.bricks import FFN, MultiheadAttention
        from ..mm.ms_deform_attn import msda_proj_supported
        if torch.is_grad_enabled() or self.training or not query.is_cuda or query.dtype != torch.float32 \
                or value is None or kwargs.get('key') is not None or kwargs.get('attn_masks') is not None:
            return False
        ok = self.__dict__.get('_rows_ok')
        if ok is None:
            layer = self.layer
            ok = tuple(layer.operation_order) == ('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm') \
                and not layer.batch_first and len(layer.attentions) == 2 and len(layer.ffns) == 1
            if ok:
                sa, ca, ffn = layer.attentions[0], layer.attentions[1], layer.ffns[0]
                bn = self.posembed.position_embedding_head[1]
                ok = type(sa) is MultiheadAttention and not sa.batch_first \
                    and sa.attn._qkv_same_embed_dim and sa.attn.in_proj_bias is not None \
                    and sa.attn.bias_k is None and not sa.attn.add_zero_attn \
                    and isinstance(ca, MultiScaleDeformableAttention) and not ca.batch_first \
                    and msda_proj_supported(ca.embed_dims // ca.num_heads, ca.num_levels, ca.num_points) \
                    and type(ffn) is FFN and ffn.num_fcs == 2 and ffn.add_identity \
                    and type(ffn.layers[0][1]) is nn.ReLU and isinstance(ffn.dropout_layer, nn.Identity) \
                    and all(type(n) is nn.LayerNorm and n.elementwise_affine for n in layer.norms) \
                    and ca.embed_dims % 128 == 0 and ca.embed_dims <= 1024 and bn.track_running_stats
            self.__dict__['_rows_ok'] = bool(ok)
        return ok

    def _forward_rows(self, query, query_pos, value, reference_points_input, spatial_shapes,
                      level_start_index, key_padding_mask):
        """The whole layer on (B*Q, C) rows: 9 GEMMs (bias / ReLU in their epilogues), the fused
        attention kernel, the projection-fed MSDA launch and three bias + residual + LayerNorm
        passes (the second also emits the cross attention's query + pos) -- no permute copies, no
        stand-alone adds. Same arithmetic as the module-by-module path."""
        import torch.nn.functional as F

        from ..mm import point_ops as P
        from ..mm.ms_deform_attn import msda_from_projections
        layer = self.layer
        sa, ca, ffn = layer.attentions[0], layer.attentions[1], layer.ffns[0]
        n1, n2, n3 = layer.norms
        Q, B, C = query.shape
        x = query.permute(1, 0, 2).reshape(B * Q, C)               # rows; a view when (B,C,Q)^T came in
        pos = self.posembed.rows(query_pos)                         # (B*Q, C)
        # --- self attention: q = k = x + pos, v = x
        mha = sa.attn
        H = mha.num_heads
        xp = x + pos
        w_in, b_in = mha.in_proj_weight, mha.in_proj_bias
        qk = torch.addmm(b_in[:2 * C], xp, w_in[:2 * C].t())       # (B*Q, 2C): q | k
        v = torch.addmm(b_in[2 * C:], x, w_in[2 * C:].t())
        if P.mha_supported(C // H):
            # exact-fp32 attention among the proposals (csrc/mha.cu) instead of torch's sm80 memory-efficient kernel
            att = P.mha_rows(qk[:, :C], qk[:, C:], v, B, H, batch_first=True)
        else:
            qk5 = qk.view(B, Q, 2, H, C // H)
            att = F.scaled_dot_product_attention(qk5[:, :, 0].transpose(1, 2), qk5[:, :, 1].transpose(1, 2),
                                                 v.view(B, Q, H, C // H).transpose(1, 2))     # (B,H,Q,d)
            att = att.transpose(1, 2).reshape(B * Q, C)
        t = torch.mm(att, mha.out_proj.weight.t())
        x, xp = P.bias_layer_norm_rows(t, n1.weight, n1.bias, n1.eps, bias=mha.out_proj.bias, residual=x,
                                       out=t, post_add=pos)
        # --- cross attention over the pyramid
        val = value.permute(1, 0, 2)                                # (B,S,C), contiguous by construction
        S = val.shape[1]
        val = torch.addmm(ca.value_proj.bias, val.reshape(B * S, C), ca.value_proj.weight.t())
        if key_padding_mask is not None:
            val.masked_fill_(key_padding_mask.reshape(-1, 1), 0.0)
        wq, bq = ca._fused_query_proj()
        proj = torch.addmm(bq, xp, wq.t())
        o = msda_from_projections(val.view(B, S, ca.num_heads, -1), spatial_shapes, level_start_index, proj,
                                  reference_points_input.contiguous(), ca.num_levels, ca.num_points)
        t = torch.mm(o.view(B * Q, C), ca.output_proj.weight.t())
        x = P.bias_layer_norm_rows(t, n2.weight, n2.bias, n2.eps, bias=ca.output_proj.bias, residual=x, out=t)
        # --- FFN
        fc1, fc2 = ffn.layers[0][0], ffn.layers[1]
        h = torch._addmm_activation(fc1.bias, x, fc1.weight.t())
        t = torch.mm(h, fc2.weight.t())
        x = P.bias_layer_norm_rows(t, n3.weight, n3.bias, n3.eps, bias=fc2.bias, residual=x, out=t)
        return x.view(B, Q, C).transpose(0, 1)                      # (Q,B,C) view

    def forward(self, query, query_pos, *args, reference_points=None, valid_ratios=None, **kwargs):
        """query (Q,B,C); query_pos (B,Q,6); reference_points (B,Q,2|4) in [0,1];
        valid_ratios (B,L,2). Returns (Q,B,C)."""
        if self.fused_eval and not args and set(kwargs) <= {'key', 'value', 'key_padding_mask', 'spatial_shapes',
                                                            'level_start_index'} \
                and self._rows_path_ok(query, kwargs.get('value'), kwargs):
            if reference_points.shape[-1] == 4:
                ref = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
            else:
                ref = reference_points[:, :, None] * valid_ratios[:, None]
            return self._forward_rows(query, query_pos, kwargs['value'], ref, kwargs['spatial_shapes'],
                                      kwargs['level_start_index'], kwargs.get('key_padding_mask'))
        if reference_points.shape[-1] == 4:
            reference_points_input = reference_points[:, :, None] * \
                torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        else:
            assert reference_points.shape[-1] == 2
            reference_points_input = reference_points[:, :, None] * valid_ratios[:, None]
        query_pos_embed = self.posembed(query_pos).permute(2, 0, 1)
        return self.layer(query, *args, query_pos=query_pos_embed,
                          reference_points=reference_points_input, **kwargs)
