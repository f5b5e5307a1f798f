"""Image-branch encoder of DeMF (reference: demf/modeling/layers/deform_detr_encoder.py:12-157):
the Deformable-DETR encoder cut out of the 2D detector -- six post-norm layers of multi-scale
deformable SELF attention (every pixel of the 4-level pyramid is a query, Q = S) plus FFN --
run frozen on the neck's pyramid; its output pyramid is what the DeMF head samples.

Same constructor, parameter names (`encoder.layers.N.*`, `level_embeds`) and forward contract
as the reference, so `img_bbox_head.transformer.encoder.*` / `...level_embeds` keys of a
Deformable-DETR checkpoint land here after the detector's key rewrite (demfnet.py:85-101).

Host-side differences (results identical): spatial shapes stay Python tuples, so building the
reference points and splitting the memory costs no device->host sync (the reference iterates
over the device tensor); the padding masks, sine encodings, valid ratios and reference points
depend only on the image shapes and are cached per (pyramid shape, image shapes).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..mm.bricks import BaseModule
from ..mm.registry import HEADS, build_positional_encoding, build_transformer_layer_sequence


@HEADS.register_module()
class DeformableDetrEncoder(BaseModule):

    def __init__(self, encoder=None,
                 positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True),
                 num_feature_levels=4, embed_dims=256, init_cfg=None):
        super().__init__(init_cfg=init_cfg)
        self.encoder = build_transformer_layer_sequence(encoder)
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.num_feature_levels = num_feature_levels
        self.embed_dims = embed_dims
        # torch.Tensor(L, C) upstream, i.e. uninitialised until a checkpoint is loaded; start
        # from N(0,1) like Deformable-DETR's own init so a randomly built model is finite
        self.level_embeds = nn.Parameter(torch.empty(num_feature_levels, embed_dims).normal_())
        self._geometry_cache = {}

    def init_weights(self):
        """Xavier-uniform on every >1-d `weight`, as the reference does (deform_detr_encoder.py:31-36);
        the deformable attention keeps its own offset/weight initialisation."""
        from ..mm.ms_deform_attn import MultiScaleDeformableAttention
        for m in self.modules():
            if hasattr(m, 'weight') and isinstance(m.weight, torch.Tensor) and m.weight.dim() > 1:
                nn.init.xavier_uniform_(m.weight)
                if getattr(m, 'bias', None) is not None:
                    nn.init.constant_(m.bias, 0.)
        for m in self.modules():
            if isinstance(m, MultiScaleDeformableAttention):
                m.init_weights()
        self._is_init = True

    # ------------------------------------------------------------------ geometry ----
    @staticmethod
    def get_valid_ratio(mask):
        """mask (B,H,W) bool, True = padding -> (B,2) = (valid width, valid height) fractions."""
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel centres of every level, normalised by the valid extent and re-scaled per sampled
        level: (B, sum(H*W), L, 2) (deform_detr_encoder.py:48-67)."""
        reference_points_list = []
        for lvl, (H, W) in enumerate(spatial_shapes):
            H, W = int(H), int(W)
            ref_y, ref_x = torch.meshgrid(
                torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device),
                torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device), indexing='ij')
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
            reference_points_list.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(reference_points_list, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def _geometry(self, shapes, img_metas, device):
        """Everything that depends only on the shapes: per-level padding masks, flattened
        positional encodings (without the level embedding), valid ratios, reference points and
        the shape / start-index tensors the attention kernel reads."""
        input_h, input_w = img_metas[0]['batch_input_shape']
        img_shapes = tuple(tuple(int(v) for v in m['img_shape'][:2]) for m in img_metas)
        key = (shapes, (int(input_h), int(input_w)), img_shapes, str(device))
        hit = self._geometry_cache.get(key)
        if hit is not None:
            return hit
        img_masks = torch.ones((len(img_metas), int(input_h), int(input_w)), dtype=torch.float32)
        for i, (h, w) in enumerate(img_shapes):
            img_masks[i, :h, :w] = 0
        img_masks = img_masks.to(device)
        masks, pos = [], []
        for hw in shapes:
            masks.append(F.interpolate(img_masks[None], size=hw).to(torch.bool).squeeze(0))
            pos.append(self.positional_encoding(masks[-1]).flatten(2).transpose(1, 2))
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        spatial_shapes = torch.as_tensor(shapes, dtype=torch.long).to(device)
        starts = [0]
        for h, w in shapes[:-1]:
            starts.append(starts[-1] + h * w)
        level_start_index = torch.as_tensor(starts, dtype=torch.long).to(device)
        mask_flatten = torch.cat([m.flatten(1) for m in masks], 1)
        geo = dict(masks=masks, pos=pos, valid_ratios=valid_ratios, spatial_shapes=spatial_shapes,
                   level_start_index=level_start_index, mask_flatten=mask_flatten,
                   any_padding=bool(mask_flatten.any()),
                   reference_points=self.get_reference_points(shapes, valid_ratios, device))
        if len(self._geometry_cache) > 16:
            self._geometry_cache.clear()
        self._geometry_cache[key] = geo
        return geo

    # ------------------------------------------------------- inference on token rows ----
    fused_eval = True

    def _rows_path_ok(self, feat):
        """The frozen branch as configured upstream: post-norm (self_attn, norm, ffn, norm) layers of
        deformable attention + 2-layer ReLU FFN + LayerNorm, no gradients, fp32 on the GPU."""
        from ..mm.ms_deform_attn import MultiScaleDeformableAttention, msda_proj_supported
        if torch.is_grad_enabled() or not feat.is_cuda or feat.dtype != torch.float32 or self.training:
            return False
        ok = self.__dict__.get('_rows_ok')
        if ok is None:
            ok = getattr(self.encoder, 'post_norm', None) is None
            for layer in self.encoder.layers:
                att = layer.attentions[0] if len(layer.attentions) == 1 else None
                ffn = layer.ffns[0] if len(layer.ffns) == 1 else None
                ok = ok and layer.operation_order == ('self_attn', 'norm', 'ffn', 'norm') \
                    and isinstance(att, MultiScaleDeformableAttention) and not att.batch_first \
                    and msda_proj_supported(att.embed_dims // att.num_heads, att.num_levels, att.num_points) \
                    and ffn is not None and ffn.num_fcs == 2 and ffn.add_identity \
                    and type(ffn.layers[0][1]) is nn.ReLU and isinstance(ffn.dropout_layer, nn.Identity) \
                    and all(type(n) is nn.LayerNorm and n.elementwise_affine for n in layer.norms) \
                    and self.embed_dims % 128 == 0 and self.embed_dims <= 1024
            self.__dict__['_rows_ok'] = bool(ok)
        return ok

    def _pos_projection(self, li, att, pos_rows, geo):
        """(pos + level_embed) @ [W_off; W_attn]^T + b for layer `li`: the positional half of the
        query projection is the same for every image of this geometry -- computed once and added
        to x @ W^T inside the sampling kernel."""
        w, b = att._fused_query_proj()
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, self.level_embeds._version,
               self.level_embeds.data_ptr())
        cache = geo.setdefault('pos_proj', {})
        hit = cache.get(li)
        if hit is None or hit[0] != key:
            hit = (key, torch.addmm(b, pos_rows, w.t()))
            cache[li] = hit
        return hit[1], w

    def _encode_rows(self, mlvl_feats, pos, geo):
        """All layers on contiguous token rows (B*S, C): per layer 5 GEMMs (bias and ReLU in their
        epilogues), one projection-fed MSDA launch and two bias + residual + LayerNorm passes;
        nothing else touches the activations."""
        from ..mm import point_ops as P
        from ..mm.ms_deform_attn import msda_from_projections
        bs, c = mlvl_feats[0].shape[:2]
        if len(mlvl_feats) <= 8 and all(f.is_contiguous() for f in mlvl_feats):
            x = P.levels_to_rows(list(mlvl_feats)).view(-1, c)                                # (B*S,C)
        else:
            x = torch.cat([f.flatten(2) for f in mlvl_feats], 2).transpose(1, 2).reshape(-1, c)
        S = x.shape[0] // bs
        pos_rows = geo.get('pos_rows')
        key = (self.level_embeds._version, self.level_embeds.data_ptr())
        if pos_rows is None or pos_rows[0] != key:
            rows = torch.cat([p + self.level_embeds[lvl].view(1, 1, -1) for lvl, p in enumerate(pos)], 1)
            pos_rows = (key, rows.reshape(-1, c).contiguous())
            geo['pos_rows'] = pos_rows
        pos_rows = pos_rows[1]
        mask = geo['mask_flatten'].reshape(-1, 1) if geo['any_padding'] else None
        ref = geo['reference_points']
        for li, layer in enumerate(self.encoder.layers):
            att, ffn, (n1, n2) = layer.attentions[0], layer.ffns[0], layer.norms
            pos_proj, wq = self._pos_projection(li, att, pos_rows, geo)
            proj = torch.mm(x, wq.t())               # + pos_proj inside the kernel = (x + pos) @ Wq^T + bq
            value = torch.addmm(att.value_proj.bias, x, att.value_proj.weight.t())
            if mask is not None:
                value.masked_fill_(mask, 0.0)
            o = msda_from_projections(value.view(bs, S, att.num_heads, -1), geo['spatial_shapes'],
                                      geo['level_start_index'], proj, ref, att.num_levels, att.num_points,
                                      proj_add=pos_proj)
            t = torch.mm(o.view(-1, c), att.output_proj.weight.t())
            x = P.bias_layer_norm_rows(t, n1.weight, n1.bias, n1.eps, bias=att.output_proj.bias,
                                       residual=x, out=t)              # LN(identity + out_proj(o))
            fc1, fc2 = ffn.layers[0][0], ffn.layers[1]
            h = torch._addmm_activation(fc1.bias, x, fc1.weight.t())        # ReLU in the epilogue
            t = torch.mm(h, fc2.weight.t())
            x = P.bias_layer_norm_rows(t, n2.weight, n2.bias, n2.eps, bias=fc2.bias, residual=x, out=t)
        return x

    # ------------------------------------------------------------------- forward ----
    def forward(self, mlvl_feats, img_metas):
        """mlvl_feats: L tensors (B,C,H_l,W_l); img_metas[i]: 'batch_input_shape' (H,W) of the padded
        batch and 'img_shape' (h,w,c) of image i. Returns L tensors of the same shapes."""
        shapes = tuple((int(f.shape[-2]), int(f.shape[-1])) for f in mlvl_feats)
        geo = self._geometry(shapes, img_metas, mlvl_feats[0].device)
        return self.transformer(mlvl_feats, geo['masks'], None, _geometry=geo)

    def transformer(self, mlvl_feats, mlvl_masks, mlvl_pos_embeds, _geometry=None, **kwargs):
        bs, c = mlvl_feats[0].shape[:2]
        shapes = tuple((int(f.shape[-2]), int(f.shape[-1])) for f in mlvl_feats)
        device = mlvl_feats[0].device
        if _geometry is not None:
            geo = _geometry
            pos = geo['pos']
        else:  # reference call convention: masks and (B,C,H,W) encodings given by the caller
            pos = [p.flatten(2).transpose(1, 2) for p in mlvl_pos_embeds]
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in mlvl_masks], 1)
            starts = [0]
            for h, w in shapes[:-1]:
                starts.append(starts[-1] + h * w)
            geo = dict(valid_ratios=valid_ratios,
                       spatial_shapes=torch.as_tensor(shapes, dtype=torch.long).to(device),
                       level_start_index=torch.as_tensor(starts, dtype=torch.long).to(device),
                       mask_flatten=torch.cat([m.flatten(1) for m in mlvl_masks], 1),
                       any_padding=True,
                       reference_points=self.get_reference_points(shapes, valid_ratios, device))
        if self.fused_eval and self._rows_path_ok(mlvl_feats[0]) and not kwargs:
            memory = self._encode_rows(mlvl_feats, pos, geo).view(bs, -1, c).permute(0, 2, 1)   # (B,C,S)
            outs, start = [], 0
            for h, w in shapes:
                outs.append(memory[:, :, start:start + h * w].reshape(bs, c, h, w))
                start += h * w
            return outs
        feat_flatten = torch.cat([f.flatten(2) for f in mlvl_feats], 2).permute(2, 0, 1)   # (S,B,C)
        lvl_pos = torch.cat([p + self.level_embeds[lvl].view(1, 1, -1) for lvl, p in enumerate(pos)], 1)
        lvl_pos = lvl_pos.permute(1, 0, 2)                                                # (S,B,C)
        # an all-False padding mask leaves the values untouched: skip the masked_fill passes
        mask = geo['mask_flatten'] if geo['any_padding'] else None
        memory = self.encoder(
            query=feat_flatten, key=None, value=None, query_pos=lvl_pos, query_key_padding_mask=mask,
            spatial_shapes=geo['spatial_shapes'], reference_points=geo['reference_points'],
            level_start_index=geo['level_start_index'], valid_ratios=geo['valid_ratios'], **kwargs)
        memory = memory.permute(1, 2, 0)                                                  # (B,C,S)
        outs, start = [], 0
        for h, w in shapes:
            outs.append(memory[:, :, start:start + h * w].reshape(bs, c, h, w))
            start += h * w
        return outs
