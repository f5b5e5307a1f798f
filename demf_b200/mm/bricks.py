"""nn building blocks with the mmcv.cnn / mmcv.runner names the DeMF path instantiates.

BaseModule, ConvModule (conv -> norm -> act with upstream's attribute names `conv`, `bn`,
`activate`, so state-dict keys match released checkpoints), norm / conv
builders, and the transformer bricks of mmcv.cnn.bricks.transformer 1.3.18 that
configs/demf/demf_votenet.py:68-91 builds: MultiheadAttention, FFN, BaseTransformerLayer and
mmdet's DetrTransformerDecoderLayer. Dense projections are plain library GEMMs (cuBLAS via
torch); the sm_100a kernels live behind point_ops.py and ms_deform_attn.py.
"""
import copy
import math
import warnings

import torch
import torch.nn as nn

from .config import ConfigDict
from .registry import (ATTENTION, FEEDFORWARD_NETWORK, POSITIONAL_ENCODING, TRANSFORMER_LAYER,
                       TRANSFORMER_LAYER_SEQUENCE, build_transformer_layer, build_attention,
                       build_feedforward_network)


class BaseModule(nn.Module):
    """mmcv.runner.BaseModule: carries init_cfg; init_weights() recurses into children."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):
        if self._is_init:
            return
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


_NORMS = {
    "BN": ("bn", nn.BatchNorm2d), "BN1d": ("bn", nn.BatchNorm1d), "BN2d": ("bn", nn.BatchNorm2d),
    "BN3d": ("bn", nn.BatchNorm3d), "LN": ("ln", nn.LayerNorm), "GN": ("gn", nn.GroupNorm),
}
_CONVS = {"Conv1d": nn.Conv1d, "Conv2d": nn.Conv2d, "Conv": nn.Conv2d, None: nn.Conv2d}
_ACTS = {"ReLU": nn.ReLU, "LeakyReLU": nn.LeakyReLU, "GELU": nn.GELU, "Sigmoid": nn.Sigmoid,
         "Tanh": nn.Tanh}


def build_norm_layer(cfg, num_features, postfix=""):
    """-> (name, layer) like mmcv.cnn.build_norm_layer."""
    cfg = dict(cfg)
    layer_type = cfg.pop("type")
    if layer_type not in _NORMS:
        raise KeyError(f"Unrecognized norm type {layer_type}")
    abbr, cls = _NORMS[layer_type]
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if layer_type == "GN":
        layer = cls(num_channels=num_features, **cfg)
    else:
        layer = cls(num_features, **cfg)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(type="Conv2d") if cfg is None else dict(cfg)
    layer_type = cfg.pop("type")
    if layer_type not in _CONVS:
        raise KeyError(f"Unrecognized conv type {layer_type}")
    return _CONVS[layer_type](*args, **kwargs, **cfg)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    return _ACTS[cfg.pop("type")](**cfg)


def build_dropout(cfg):
    if cfg is None:
        return nn.Identity()
    cfg = dict(cfg)
    kind = cfg.pop("type", "Dropout")
    if kind != "Dropout":
        raise KeyError(f"Unrecognized dropout type {kind}")
    return nn.Dropout(p=cfg.get("drop_prob", cfg.get("p", 0.5)))


class ConvModule(nn.Module):
    """conv -> norm -> activation. bias='auto' means "no bias when a norm follows"."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True, order=("conv", "norm", "act")):
        super().__init__()
        assert tuple(order) == ("conv", "norm", "act"), "only the default order is on the DeMF path"
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        else:
            self.norm_name = None
        if self.with_activation:
            act_cfg = dict(act_cfg)
            if act_cfg["type"] in ("ReLU", "LeakyReLU"):
                act_cfg.setdefault("inplace", inplace)
            self.activate = build_activation_layer(act_cfg)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def init_weights(self):
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)
        if self.with_norm and getattr(self.norm, "weight", None) is not None:
            nn.init.constant_(self.norm.weight, 1)
            nn.init.constant_(self.norm.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.norm(x)
        if self.with_activation:
            x = self.activate(x)
        return x


# ------------------------------------------------------------------------ transformer --
@ATTENTION.register_module()
class MultiheadAttention(BaseModule):
    """mmcv MultiheadAttention: nn.MultiheadAttention + positional encodings + residual."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type="Dropout", drop_prob=0.), init_cfg=None, batch_first=False,
                 **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer) if dropout_layer else None
        if "dropout" in kwargs:  # deprecated alias still used by configs/demf/demf_votenet.py:78
            attn_drop = kwargs["dropout"]
            dropout_layer["drop_prob"] = kwargs.pop("dropout")
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        shared_qk = key is query and key_pos is query_pos   # self attention: q and k are one tensor
        if query_pos is not None:
            if torch.is_grad_enabled():
                query = query + query_pos
            else:  # a dense result: the input projections then are single addmm launches
                query = torch.add(query, query_pos, out=torch.empty(
                    query.shape, dtype=query.dtype, device=query.device))
        if shared_qk:
            key = query
        elif key_pos is not None:
            key = key + key_pos
        if not torch.is_grad_enabled() and not value.is_contiguous():
            value = value.contiguous()
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self._attn_train_rows(query, key, value, attn_mask, key_padding_mask, shared_qk)
        if out is None:
            out = self._attn_eval_rows(query, key, value, attn_mask, key_padding_mask, shared_qk)
        if out is None:
            # the attention map itself is never used: without it PyTorch takes its fused
            # scaled-dot-product path instead of mul + bmm + softmax + bmm + mean
            out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                            key_padding_mask=key_padding_mask, need_weights=False)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class _ParamRows:
    """Rows [lo, hi) of a packed parameter, with what `linear_rows(owner=...)` reads of a parameter."""
    __slots__ = ("par", "lo", "hi", "requires_grad")

    def __init__(self, par, lo, hi):
        self.par, self.lo, self.hi, self.requires_grad = par, lo, hi, par.requires_grad

    @property
    def grad(self):
        return None if self.par.grad is None else self.par.grad[self.lo:self.hi]


class _SliceOwner:
    __slots__ = ("weight", "bias", "params")


def _slice_owner(weight, bias, lo, hi):
    """What `linear_rows(owner=...)` needs of a module, for rows [lo, hi) of a packed projection: `.weight` /
    `.bias` objects whose `.grad` is the matching view of the packed parameter's gradient (None until it exists)."""
    o = _SliceOwner()
    o.weight = _ParamRows(weight, lo, hi)
    o.bias = None if bias is None else _ParamRows(bias, lo, hi)
    o.params = [weight] + ([] if bias is None else [bias])
    return o


def _mha_train_rows(self, query, key, value, attn_mask, key_padding_mask, shared_qk):
    """nn.MultiheadAttention.forward (need_weights=False) for TRAINING on CUDA in the TF32 mode with the four
    projections on our tensor-core GEMMs (`linear_rows`; the library picks sm80 kernels for these 1024-row
    products) and the attention itself through F.scaled_dot_product_attention, exactly what the module's own
    fast path composes. (L,B,E) in, (L,B,E) out; None when the case is not this plain one."""
    attn = self.attn
    if not (torch.is_grad_enabled() and query.is_cuda and NATIVE_TRAIN_GEMM and torch.backends.cuda.matmul.allow_tf32
            and attn_mask is None and key_padding_mask is None and attn._qkv_same_embed_dim
            and attn.in_proj_weight is not None and attn.bias_k is None and not attn.add_zero_attn
            and query.dim() == 3 and attn.embed_dim % 4 == 0):
        return None
    E, H = attn.embed_dim, attn.num_heads
    Lq, B, _ = query.shape
    Lk = key.shape[0]
    w, b = attn.in_proj_weight, attn.in_proj_bias

    def proj(x, lo, hi):
        rows = x.reshape(-1, E)
        return linear_rows(rows, w[lo:hi], None if b is None else b[lo:hi], owner=_slice_owner(w, b, lo, hi))
    if shared_qk:
        qk = proj(query, 0, 2 * E)                      # one GEMM for both: (L*B, 2E)
        q, k = qk[:, :E], qk[:, E:]
    else:
        q, k = proj(query, 0, E), proj(key, E, 2 * E)
    v = proj(value, 2 * E, 3 * E)
    # (L*B, E) -> (B, H, L, E/H), as nn.functional.multi_head_attention_forward lays heads out
    q = q.reshape(Lq, B * H, E // H).transpose(0, 1).reshape(B, H, Lq, E // H)
    k = k.reshape(Lk, B * H, E // H).transpose(0, 1).reshape(B, H, Lk, E // H)
    v = v.reshape(Lk, B * H, E // H).transpose(0, 1).reshape(B, H, Lk, E // H)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, None, attn.dropout if attn.training else 0.0)
    o = o.permute(2, 0, 1, 3).reshape(Lq * B, E)
    o = linear_rows(o, attn.out_proj.weight, attn.out_proj.bias, owner=attn.out_proj)
    return o.view(Lq, B, E)


def _mha_eval_rows(self, query, key, value, attn_mask, key_padding_mask, shared_qk):
    """nn.MultiheadAttention.forward (need_weights=False) in INFERENCE on CUDA: the projections are the library's
    bias GEMMs (one for q and k together when they are the same tensor), the attention itself our exact-fp32 kernel
    (csrc/mha.cu) instead of torch's sm80 memory-efficient attention. (L,B,E) in, (L,B,E) out; None when the case is
    not this plain one."""
    from . import point_ops as P
    attn = self.attn
    if not (not torch.is_grad_enabled() and not attn.training and query.is_cuda and query.dtype == torch.float32
            and attn_mask is None and key_padding_mask is None and attn._qkv_same_embed_dim
            and attn.in_proj_weight is not None and attn.in_proj_bias is not None and attn.bias_k is None
            and not attn.add_zero_attn and query.dim() == 3 and attn.embed_dim % attn.num_heads == 0
            and attn.out_proj.bias is not None and self.fused_eval_attention
            and P.mha_supported(attn.embed_dim // attn.num_heads)):
        return None
    E, H = attn.embed_dim, attn.num_heads
    Lq, B, _ = query.shape
    w, b = attn.in_proj_weight, attn.in_proj_bias
    q2, k2, v2 = query.reshape(-1, E), key.reshape(-1, E), value.reshape(-1, E)
    if shared_qk:
        qk = torch.addmm(b[:2 * E], q2, w[:2 * E].t())          # (L*B, 2E): q | k
        q, k = qk[:, :E], qk[:, E:]
    else:
        q, k = torch.addmm(b[:E], q2, w[:E].t()), torch.addmm(b[E:2 * E], k2, w[E:2 * E].t())
    v = torch.addmm(b[2 * E:], v2, w[2 * E:].t())
    o = P.mha_rows(q, k, v, B, H)
    o = torch.addmm(attn.out_proj.bias, o, attn.out_proj.weight.t())
    return o.view(Lq, B, E)


MultiheadAttention._attn_train_rows = _mha_train_rows
MultiheadAttention._attn_eval_rows = _mha_eval_rows
MultiheadAttention.fused_eval_attention = True   # False: nn.MultiheadAttention's own scaled_dot_product_attention


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type="ReLU", inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2, f"num_fcs should be no less than 2. got {num_fcs}."
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.num_fcs = num_fcs
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels),
                                        build_activation_layer(act_cfg), nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def _layers(self, x):
        if not (x.is_cuda and torch.is_grad_enabled()):
            return self.layers(x)
        for layer in self.layers:     # training: the Linear layers on our tensor-core GEMM kernels
            if isinstance(layer, nn.Sequential):
                for sub in layer:
                    x = linear_nd(x, sub) if isinstance(sub, nn.Linear) else sub(x)
            else:
                x = linear_nd(x, layer) if isinstance(layer, nn.Linear) else layer(x)
        return x

    def forward(self, x, identity=None):
        out = self._layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


@TRANSFORMER_LAYER.register_module()
class BaseTransformerLayer(BaseModule):
    """Generic (self_attn | cross_attn | norm | ffn)* layer, post-norm unless order starts with norm."""

    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024, num_fcs=2,
                               ffn_drop=0., act_cfg=dict(type="ReLU", inplace=True)),
                 operation_order=None, norm_cfg=dict(type="LN"), init_cfg=None, batch_first=False,
                 **kwargs):
        ffn_cfgs = copy.deepcopy(dict(ffn_cfgs)) if isinstance(ffn_cfgs, dict) else copy.deepcopy(ffn_cfgs)
        deprecated = dict(feedforward_channels="feedforward_channels", ffn_dropout="ffn_drop",
                          ffn_num_fcs="num_fcs")
        for ori, new in deprecated.items():
            if ori in kwargs:
                ffn_cfgs[new] = kwargs[ori]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        assert set(operation_order) & {"self_attn", "norm", "ffn", "cross_attn"} == set(operation_order)
        num_attn = operation_order.count("self_attn") + operation_order.count("cross_attn")
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        else:
            assert num_attn == len(attn_cfgs), (
                f"The length of attn_cfg {len(attn_cfgs)} is not consistent with the number of "
                f"attention in operation_order {operation_order}.")
        self.num_attn = num_attn
        self.operation_order = tuple(operation_order)
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == "norm"
        self.attentions = ModuleList()
        index = 0
        for name in operation_order:
            if name in ("self_attn", "cross_attn"):
                cfg = dict(attn_cfgs[index])
                if "batch_first" in cfg:
                    assert self.batch_first == cfg["batch_first"]
                else:
                    cfg["batch_first"] = self.batch_first
                attention = build_attention(cfg)
                attention.operation_name = name
                self.attentions.append(attention)
                index += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = ModuleList()
        num_ffns = operation_order.count("ffn")
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
        assert len(ffn_cfgs) == num_ffns
        for i in range(num_ffns):
            cfg = dict(ffn_cfgs[i])
            if "embed_dims" not in cfg:
                cfg["embed_dims"] = self.embed_dims
            else:
                assert cfg["embed_dims"] == self.embed_dims
            cfg.setdefault("type", "FFN")
            self.ffns.append(build_feedforward_network(cfg))
        self.norms = ModuleList()
        for _ in range(operation_order.count("norm")):
            self.norms.append(build_norm_layer(norm_cfg, self.embed_dims)[1])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
            warnings.warn(f"Use same attn_mask in all attentions in {self.__class__.__name__} ")
        else:
            assert len(attn_masks) == self.num_attn
        for layer in self.operation_order:
            if layer == "self_attn":
                temp_key = temp_value = query
                query = self.attentions[attn_index](
                    query, temp_key, temp_value, identity if self.pre_norm else None,
                    query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == "norm":
                query = self.norms[norm_index](query)
                norm_index += 1
            elif layer == "cross_attn":
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == "ffn":
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


@TRANSFORMER_LAYER.register_module()
class DetrTransformerDecoderLayer(BaseTransformerLayer):
    """mmdet DetrTransformerDecoderLayer: self_attn, norm, cross_attn, norm, ffn, norm."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type="ReLU", inplace=True), norm_cfg=dict(type="LN"), ffn_num_fcs=2,
                 **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                         ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                         norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
        assert len(operation_order) == 6
        assert set(operation_order) == {"self_attn", "norm", "cross_attn", "ffn"}


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class TransformerLayerSequence(BaseModule):
    """mmcv TransformerLayerSequence: `num_layers` layers built from one config (or a list of
    configs), applied in order with the same keyword arguments."""

    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        for layer in self.layers:
            query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos,
                          attn_masks=attn_masks, query_key_padding_mask=query_key_padding_mask,
                          key_padding_mask=key_padding_mask, **kwargs)
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerEncoder(TransformerLayerSequence):
    """mmdet DetrTransformerEncoder: the layer sequence plus a final LayerNorm that exists only
    for pre-norm layers (the reference's post-norm encoder has none)."""

    def __init__(self, *args, post_norm_cfg=dict(type="LN"), **kwargs):
        super().__init__(*args, **kwargs)
        if post_norm_cfg is not None:
            self.post_norm = build_norm_layer(post_norm_cfg, self.embed_dims)[1] \
                if self.pre_norm else None
        else:
            assert not self.pre_norm, f"Use prenorm in {self.__class__.__name__}, Please specify post_norm_cfg"
            self.post_norm = None

    def forward(self, *args, **kwargs):
        x = super().forward(*args, **kwargs)
        if self.post_norm is not None:
            x = self.post_norm(x)
        return x


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(BaseModule):
    """mmdet SinePositionalEncoding: mask (B,H,W), nonzero = padding -> (B, 2*num_feats, H, W);
    first half encodes the row coordinate, second half the column coordinate, each coordinate the
    running count of valid pixels, optionally normalised to [0, scale]."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6,
                 offset=0., init_cfg=None):
        super().__init__(init_cfg)
        if normalize:
            assert isinstance(scale, (float, int)), \
                f"when normalize is set, scale should be provided and in float or int type, found {type(scale)}"
        self.num_feats = num_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = scale
        self.eps = eps
        self.offset = offset

    def forward(self, mask):
        mask = mask.to(torch.int)
        not_mask = 1 - mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)

    def __repr__(self):
        return (f"{self.__class__.__name__}(num_feats={self.num_feats}, temperature={self.temperature}, "
                f"normalize={self.normalize}, scale={self.scale}, eps={self.eps})")


def to_config_dict(cfg):
    return cfg if isinstance(cfg, ConfigDict) or cfg is None else ConfigDict(cfg)


# ------------------------------------------------------- kernel-size-1 convs on rows --
def batch_norm_rows(bn, x):
    """Apply a BatchNorm{1,2}d module to point-major rows x (R, C): identical statistics to
    running the module on the (B,C,...) tensor those rows came from (BN reduces over every
    axis but C), with the module's own buffers and momentum handling."""
    if bn.momentum is None:
        factor = 0.0
    else:
        factor = bn.momentum
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            factor = 1.0 / float(bn.num_batches_tracked)
    use_batch_stats = bn.training or (bn.running_mean is None and bn.running_var is None)
    return torch.nn.functional.batch_norm(
        x, bn.running_mean if not bn.training or bn.track_running_stats else None,
        bn.running_var if not bn.training or bn.track_running_stats else None,
        bn.weight, bn.bias, use_batch_stats, factor, bn.eps)


_column_index_cache = {}


# BatchNorm's num_batches_tracked += 1 is one tiny kernel per layer and step (25 in the DeMF training step). Inside a
# `deferred_batch_counters()` block (engine.Trainer.step) the counters are collected and bumped by ONE foreach launch
# when the block exits; outside it they are bumped on the spot, as nn.BatchNorm does.
_NBT = {"pending": None}


def bump_batches_tracked(bn):
    t = bn.num_batches_tracked
    if t is None:
        return
    if _NBT["pending"] is not None:
        _NBT["pending"].append(t)
    else:
        t.add_(1)


class deferred_batch_counters:
    def __enter__(self):
        self.prev = _NBT["pending"]
        _NBT["pending"] = []
        return self

    def __exit__(self, *exc):
        pending, _NBT["pending"] = _NBT["pending"], self.prev
        if pending and exc[0] is None:
            uniq = list({id(t): t for t in pending}.values())
            counts = [sum(1 for q in pending if q is t) for t in uniq]
            if all(c == 1 for c in counts):
                torch._foreach_add_(uniq, 1)
            else:
                for t, c in zip(uniq, counts):
                    t.add_(c)
        return False


def permute_weight_columns(w, cols, zero_pad=True):
    """w (Cout, Cin) -> (Cout, len(cols)); cols[j] = source column or -1 for a zero column."""
    if not zero_pad:
        # training: the padding columns of the ROWS are exact zeros, so a padding column of the weight may hold
        # anything finite -- it is given a copy of column 0 (one index_select, no pad kernel; the gradient that
        # flows back into column 0 through the copy is a sum of products with those zeros: exactly 0)
        key = (tuple(cols), w.size(1), str(w.device), "dup")
        idx = _column_index_cache.get(key)
        if idx is None:
            idx = torch.as_tensor([c if c >= 0 else 0 for c in cols]).to(w.device)
            _column_index_cache[key] = idx
        return w.index_select(1, idx)
    key = (tuple(cols), w.size(1), str(w.device))
    idx = _column_index_cache.get(key)
    if idx is None:  # built once per layout and device: no host->device copy in the step
        idx = torch.as_tensor([c if c >= 0 else w.size(1) for c in cols]).to(w.device)
        _column_index_cache[key] = idx
    return torch.nn.functional.pad(w, (0, 1)).index_select(1, idx)


def _fold_conv_bn(cm, cols):
    """(W', b') with  act(bn(conv(x))) == act(x @ W'^T + b')  for a ConvModule in eval mode:
    W' = W * gamma / sqrt(var + eps) (rows), b' = beta + (bias - mean) * gamma / sqrt(var + eps)."""
    w = cm.conv.weight.flatten(1)
    if cols is not None:
        w = permute_weight_columns(w, cols)
    b = cm.conv.bias
    if cm.with_norm:
        bn = cm.norm
        scale = torch.rsqrt(bn.running_var + bn.eps)
        if bn.weight is not None:
            scale = scale * bn.weight
        w = w * scale.unsqueeze(1)
        shift = -bn.running_mean * scale if b is None else (b - bn.running_mean) * scale
        b = shift if bn.bias is None else shift + bn.bias
    elif b is None:
        b = w.new_zeros(w.size(0))
    return w.contiguous(), b.contiguous()


def _folded_cached(cm, cols):
    tensors = [cm.conv.weight, cm.conv.bias]
    if cm.with_norm:
        bn = cm.norm
        tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    key = (tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors),
           None if cols is None else tuple(cols))
    cache = cm.__dict__.get("_rows_fold")
    if cache is None or cache[0] != key:
        cache = (key,) + _fold_conv_bn(cm, cols)
        cm.__dict__["_rows_fold"] = cache
    return cache[1], cache[2]


def fold_key(cm):
    """Identity of the folded (W', b') `_folded_cached` last built for `cm`: (storage, version) of every
    source tensor. Callers that cache something derived from the folded weights key on this, not on the
    folded tensors' addresses (a refold allocates new tensors and the allocator reuses freed addresses)."""
    cache = cm.__dict__.get("_rows_fold")
    return None if cache is None else cache[0]


def _foldable(cm):
    if cm.with_norm:
        bn = cm.norm
        if bn.training or not bn.track_running_stats or not isinstance(
                bn, nn.modules.batchnorm._BatchNorm):
            return False
    return (not cm.with_activation) or type(cm.activate) is nn.ReLU


def conv_module_rows(cm, x, cols=None):
    """A kernel-size-1 ConvModule (Conv1d/Conv2d -> BN -> act) applied to rows x (R, Cin'): the
    1x1 convolution IS a GEMM over rows, so no im2col / NCHW round trip is needed.
    `cols` permutes the weight's input columns to the row layout (point_ops.group_rows_columns).

    Inference (BN in eval mode, no autograd): BN is folded into the weights once and the layer
    is ONE library GEMM with a bias+ReLU epilogue -- the activation tensor is written exactly
    once and never re-read by a normalisation or clamp pass.
    Training: GEMM, then batch-statistics BN (F.batch_norm on rows), then the activation."""
    if not torch.is_grad_enabled() and _foldable(cm):
        w, b = _folded_cached(cm, cols)
        if cm.with_activation:
            return torch._addmm_activation(b, x, w.t())
        return torch.addmm(b, x, w.t())
    w = cm.conv.weight.flatten(1)
    if cols is not None:
        w = permute_weight_columns(w, cols, zero_pad=False)   # x's padding columns are exact zeros (csrc/rows.cu)
    y, prestats = _linear_rows(cm, x, w, direct_wgrad=cols is None)
    if cm.with_norm and _fused_bn_ok(cm, y):
        # training: batch statistics (from the GEMM's epilogue when it ran on our tensor-core kernel) +
        # normalise + ReLU
        from . import point_ops as P
        bn = cm.norm
        bump_batches_tracked(bn)
        return P.batch_norm_relu_rows(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum,
                                      bn.eps, cm.with_activation, _bn_state(bn, y.device), prestats=prestats)
    assert not prestats
    if cm.with_norm:
        y = batch_norm_rows(cm.norm, y)
    if cm.with_activation:
        y = cm.activate(y)
    return y


FUSED_BN_TRAIN = True
FUSED_BN_MAX_TRAIN = True


def conv_module_rows_max(cm, x, ns, cols=None):
    """Training only: conv -> BN (batch statistics) -> ReLU -> max over every `ns` consecutive rows, for the
    last layer of a set-abstraction MLP: (M*ns, Cin) -> (M, Cout) without materialising the normalised
    (M*ns, Cout) tensor (csrc/bn_rows.cu). Returns None when the fused path does not apply."""
    if not (FUSED_BN_MAX_TRAIN and torch.is_grad_enabled() and cm.with_norm and cm.with_activation
            and x.is_cuda and ns <= 255 and x.shape[0] % ns == 0):
        return None
    if not _fused_bn_shape_ok(cm, x):
        return None
    w = cm.conv.weight.flatten(1)
    if cols is not None:
        w = permute_weight_columns(w, cols, zero_pad=False)   # x's padding columns are exact zeros (csrc/rows.cu)
    y, prestats = _linear_rows(cm, x, w, direct_wgrad=cols is None)
    from . import point_ops as P
    bn = cm.norm
    bump_batches_tracked(bn)
    return P.batch_norm_relu_max_rows(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum,
                                      bn.eps, ns, _bn_state(bn, y.device), prestats=prestats)


# ------------------------------------------------ weight gradients off the critical path --
# In backward, a 1x1 convolution on rows needs dX = dY W (the next link of the chain) and dW = dY^T X,
# which nothing downstream waits for. Autograd runs both on one stream, so the bandwidth-heavy dW GEMMs
# (K = all grouped rows of the level) sit between the many small latency-bound kernels of the chain.
# A trainer that owns the gradient buffers can ask for dW on a second stream instead: it is accumulated
# straight into `param.grad` there (no autograd accumulation node, no extra add) and the trainer joins the
# stream before it reads any gradient. Off unless a `with async_weight_grads(device):` block is active.
_ASYNC_WGRAD = {"on": False, "streams": {}}


def _wgrad_stream(device):
    key = str(device)
    if key not in _ASYNC_WGRAD["streams"]:
        _ASYNC_WGRAD["streams"][key] = torch.cuda.Stream(device=device)
    return _ASYNC_WGRAD["streams"][key]


def _mark_direct(conv):
    """These parameters' gradients are accumulated in place by our kernels (see engine.FlatGradients.release)."""
    if hasattr(conv, "params"):          # a view of a packed parameter (bricks._slice_owner)
        for par in conv.params:
            par._demf_direct_grad = True
        return
    conv.weight._demf_direct_grad = True
    if conv.bias is not None:
        conv.bias._demf_direct_grad = True


class async_weight_grads:
    """Context for forward + backward of one step: weight gradients of the rows convolutions are computed
    on a side stream into the pre-allocated `param.grad` buffers; leaving the block joins that stream."""

    def __init__(self, device):
        self.device = torch.device(device)

    def __enter__(self):
        self.prev = _ASYNC_WGRAD["on"]
        _ASYNC_WGRAD["on"] = self.device.type == "cuda"
        return self

    def __exit__(self, *exc):
        _ASYNC_WGRAD["on"] = self.prev
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).wait_stream(_wgrad_stream(self.device))
        return False


class _LinearRowsAsyncWgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight2d, bias, conv):
        ctx.save_for_backward(x, weight2d)
        ctx.conv = conv
        return torch.nn.functional.linear(x, weight2d, bias)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        conv = ctx.conv
        gy = gy.contiguous()
        gx = gy.mm(w) if ctx.needs_input_grad[0] else None
        cur = torch.cuda.current_stream(gy.device)
        side = _wgrad_stream(gy.device)
        side.wait_stream(cur)
        gy.record_stream(side)
        x.record_stream(side)
        with torch.cuda.stream(side):
            conv.weight.grad.view(w.shape).addmm_(gy.t(), x)
            if conv.bias is not None and conv.bias.grad is not None:
                conv.bias.grad.add_(gy.sum(0))
        return gx, None, None, None


# Training GEMMs on the hand-written tcgen05 + TMA kernels (csrc/gemm_tf32.cu) instead of the library: forward
# with the BatchNorm statistics in its epilogue, data gradient, split-K weight gradient. TF32 products, so only
# in the 'tf32' arithmetic mode (engine.set_gemm_precision); the strict-fp32 mode keeps IEEE library GEMMs.
NATIVE_TRAIN_GEMM = True


def _tc_ok(x, w):
    return (NATIVE_TRAIN_GEMM and x.is_cuda and torch.is_grad_enabled() and torch.backends.cuda.matmul.allow_tf32
            and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0 and x.stride(1) == 1
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and w.dtype == torch.float32
            and w.stride(1) == 1 and w.stride(0) % 4 == 0 and w.data_ptr() % 16 == 0
            and x.shape[1] % 4 == 0 and w.shape[0] % 4 == 0 and x.shape[1] <= 2048 and w.shape[0] <= 2048)


class _LinearRowsTC(torch.autograd.Function):
    """y = x @ w^T (+ bias) with all three GEMMs of the layer on csrc/gemm_tf32.cu. `bn_state`: the forward
    epilogue also accumulates the per-channel sums of y for the BatchNorm behind the convolution. `conv`: the
    weight gradient is accumulated straight into conv.weight.grad on the weight-gradient stream (inside an
    `async_weight_grads` block); otherwise it is returned through autograd."""

    @staticmethod
    def forward(ctx, x, w, bias, conv, bn_state):
        from . import point_ops as P
        ctx.save_for_backward(x, w)
        ctx.conv = conv
        ctx.has_bias = bias is not None
        return P.gemm_rows_fwd(x, w, bias=bias, bn_state=bn_state)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        from . import point_ops as P
        x, w = ctx.saved_tensors
        conv = ctx.conv
        gy = gy.contiguous()
        gx = P.gemm_rows_dgrad(gy, w) if ctx.needs_input_grad[0] else None
        direct = (conv is not None and _ASYNC_WGRAD["on"] and conv.weight.grad is not None
                  and (conv.bias is None or conv.bias.grad is not None))
        if direct:
            cur = torch.cuda.current_stream(gy.device)
            side = _wgrad_stream(gy.device)
            side.wait_stream(cur)
            gy.record_stream(side)
            x.record_stream(side)
            with torch.cuda.stream(side):
                P.gemm_wgrad_(conv.weight.grad.view(w.shape), gy, x)
                if conv.bias is not None:
                    P.col_sum_add_(conv.bias.grad, gy)
            return gx, None, None, None, None
        gw = gb = None
        if ctx.needs_input_grad[1]:
            gw = torch.zeros_like(w)
            P.gemm_wgrad_(gw, gy, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None, None


class _BnReluConvRows(torch.autograd.Function):
    """[BatchNorm (batch statistics) + ReLU of layer l] -> [1x1 convolution of layer l+1] on rows as ONE autograd
    node, so that the backward of the pair can be fused across the layer boundary (csrc/gemm_tf32.cu):
      forward   mean/invstd of y_l from the sums the previous GEMM's epilogue left in `state_prev`; z_l = relu(bn(y_l))
                (one pass); y_{l+1} = z_l W^T with ITS BatchNorm's sums accumulated into `state_next` by the epilogue;
      backward  g = (dy_{l+1} W) masked by the ReLU and the two BatchNorm-backward reductions come out of ONE GEMM
                launch (gemm_rows_dgrad_bn) -- no pass over (dz, z, y) -- then one pass writes dL/dy_l; the weight
                gradient dy_{l+1}^T z_l goes to the split-K tensor-core kernel (weight-gradient stream when the
                trainer owns the gradient buffers).
    Mirrors, for mmcv ConvModule stacks (conv -> BN -> ReLU -> conv ...), torch's batch_norm / threshold /
    conv backward chain."""

    @staticmethod
    def forward(ctx, y_prev, gamma, beta, w, bn_prev, conv, state_prev, state_next, direct):
        from . import point_ops as P
        R, C = y_prev.shape
        mean, invstd = P.bn_finalize(state_prev, R, C, bn_prev.eps, bn_prev.momentum, bn_prev.running_mean,
                                     bn_prev.running_var)
        z = P.bn_rows_apply(y_prev, gamma, beta, mean, invstd, True)
        y = P.gemm_rows_fwd(z, w, bn_state=state_next)
        ctx.save_for_backward(y_prev, z, w, gamma, beta, mean, invstd, state_prev)
        ctx.conv = conv if direct else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        from . import point_ops as P
        y_prev, z, w, gamma, beta, mean, invstd, state_prev = ctx.saved_tensors
        conv = ctx.conv
        gy = gy.contiguous()
        g = P.gemm_rows_dgrad_bn(gy, w, y_prev, mean, invstd, gamma, beta, state_prev)
        gyp, ggamma, gbeta = P.bn_bwd_from_masked(g, y_prev, gamma, mean, invstd, state_prev)
        gw = None
        if conv is not None and _ASYNC_WGRAD["on"] and conv.weight.grad is not None:
            cur = torch.cuda.current_stream(gy.device)
            side = _wgrad_stream(gy.device)
            side.wait_stream(cur)
            gy.record_stream(side)
            z.record_stream(side)
            with torch.cuda.stream(side):
                P.gemm_wgrad_(conv.weight.grad.view(w.shape), gy, z)
        elif ctx.needs_input_grad[3]:
            gw = torch.zeros_like(w)
            P.gemm_wgrad_(gw, gy, z)
        return gyp, ggamma, gbeta, gw, None, None, None, None, None


def sa_mlp_train_rows(mlp, x, ns, cols0=None):
    """Training path of a set-abstraction shared MLP + max on grouped rows x (M*ns, K0): conv -> [BN+ReLU -> conv]*
    -> BN+ReLU+max over every `ns` rows, every GEMM on the tcgen05 kernels, BatchNorm statistics out of the GEMM
    epilogues, BatchNorm-backward reductions out of the data-gradient GEMMs. Returns pooled (M, C_last), or None
    when the chain does not apply (then the caller runs layer by layer)."""
    layers = list(mlp)
    if not (NATIVE_TRAIN_GEMM and FUSED_BN_TRAIN and FUSED_BN_MAX_TRAIN and torch.is_grad_enabled() and len(layers) >= 2
            and x.is_cuda and ns <= 255 and x.shape[0] % ns == 0):
        return None
    for cm in layers:
        if not (isinstance(cm, ConvModule) and cm.with_norm and cm.with_activation and cm.conv.bias is None
                and _fused_bn_shape_ok(cm, x) and cm.conv.out_channels <= 256 and cm.conv.out_channels % 4 == 0):
            return None
    w0 = layers[0].conv.weight.flatten(1)
    if cols0 is not None:
        w0 = permute_weight_columns(w0, cols0, zero_pad=False)   # x's padding columns are exact zeros (csrc/rows.cu)
    if not _tc_ok(x, w0):
        return None
    from . import point_ops as P
    for cm in layers:
        bump_batches_tracked(cm.norm)
    y, prestats = _linear_rows(layers[0], x, w0, direct_wgrad=cols0 is None)
    assert prestats
    for prev, cm in zip(layers[:-1], layers[1:]):
        bn, conv = prev.norm, cm.conv
        eligible = _ASYNC_WGRAD["on"] and conv.weight.requires_grad
        if eligible:
            _mark_direct(conv)      # from the next step on the trainer leaves this gradient's view in place
        direct = eligible and conv.weight.grad is not None
        y = _BnReluConvRows.apply(y, bn.weight, bn.bias, conv.weight.flatten(1), bn, conv, _bn_state(bn, y.device),
                                  _bn_state(cm.norm, y.device), direct)
    bn = layers[-1].norm
    return P.batch_norm_relu_max_rows(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps,
                                      ns, _bn_state(bn, y.device), prestats=True)


def linear_rows(x, w, bias=None, owner=None):
    """x (R,K) @ w (N,K)^T + bias for a plain Linear / kernel-size-1 convolution WITHOUT a BatchNorm behind it.
    `owner`: the nn.Linear / conv module whose `.weight` / `.bias` ARE (views of) w and bias -- inside an
    `async_weight_grads` block its gradients are then accumulated in place on the weight-gradient stream.
    Training on CUDA in the TF32 mode: all three GEMMs on csrc/gemm_tf32.cu (an output width that is not a
    multiple of 4 is zero-padded so that row strides stay 16-byte multiples, e.g. VoteModule.conv_out's 259 and
    conv_reg's 30 channels -- upstream's unaligned shapes fall to the library's sm80 kernels); otherwise the
    library."""
    N = w.shape[0]
    pad = (-N) % 4
    if x.is_cuda and torch.is_grad_enabled() and x.dim() == 2 and x.shape[1] % 4 == 0:
        wp, bp = w, bias
        if pad:
            wp = torch.nn.functional.pad(w, (0, 0, 0, pad))
            bp = None if bias is None else torch.nn.functional.pad(bias, (0, pad))
        xc = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) else x.contiguous()
        wp = wp if (wp.stride(1) == 1 and wp.stride(0) % 4 == 0 and wp.data_ptr() % 16 == 0) else wp.contiguous()
        if _tc_ok(xc, wp):
            direct = None
            if owner is not None and not pad and _ASYNC_WGRAD["on"] and owner.weight.requires_grad \
                    and (owner.bias is None or owner.bias.requires_grad):
                _mark_direct(owner)
                if owner.weight.grad is not None and (owner.bias is None or owner.bias.grad is not None):
                    direct = owner
            y = _LinearRowsTC.apply(xc, wp, bp, direct, None)
            return y[:, :N] if pad else y
    return torch.nn.functional.linear(x, w, bias)


def linear_nd(x, lin):
    """nn.Linear `lin` applied to x (..., K) through linear_rows."""
    lead = x.shape[:-1]
    y = linear_rows(x.reshape(-1, x.shape[-1]), lin.weight, lin.bias, owner=lin)
    return y.reshape(*lead, y.shape[-1])


def _bn_state(bn, device):
    from . import point_ops as P
    state = bn.__dict__.get("_rows_state")
    if state is None or state.device != device:
        state = P.bn_rows_state(bn.num_features, device)
        bn.__dict__["_rows_state"] = state
    return state


def _linear_rows(cm, x, w, direct_wgrad=True):
    """x @ w^T + bias for a ConvModule's 1x1 convolution; `w` is the flattened (possibly column-permuted) view
    of cm.conv.weight. Returns (y, prestats): prestats = the BatchNorm statistics of y were accumulated into the
    layer's state block by the GEMM epilogue. `direct_wgrad`: w IS the parameter (not a permuted copy), so its
    gradient may be accumulated in place on the weight-gradient stream."""
    conv = cm.conv
    eligible = (direct_wgrad and _ASYNC_WGRAD["on"] and x.is_cuda and torch.is_grad_enabled()
                and conv.weight.requires_grad)
    if eligible:
        _mark_direct(conv)          # from the next step on the trainer leaves these gradients' views in place
    direct = eligible and conv.weight.grad is not None and (conv.bias is None or conv.bias.grad is not None)
    if _tc_ok(x, w):
        fuse = cm.with_norm and w.shape[0] <= 256 and _fused_bn_shape_ok(cm, x)
        state = _bn_state(cm.norm, x.device) if fuse else None
        return _LinearRowsTC.apply(x, w, conv.bias, conv if direct else None, state), fuse
    if direct:
        return _LinearRowsAsyncWgrad.apply(x, w, conv.bias, conv), False
    return torch.nn.functional.linear(x, w, conv.bias), False


def _fused_bn_shape_ok(cm, x):
    """_fused_bn_ok for the output of this module's convolution applied to rows x, before it exists."""
    if not cm.with_norm:
        return False
    bn = cm.norm
    if not (FUSED_BN_TRAIN and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 1
            and isinstance(bn, nn.modules.batchnorm._BatchNorm) and bn.training and bn.affine
            and bn.track_running_stats and bn.momentum is not None):
        return False
    if cm.with_activation and type(cm.activate) is not nn.ReLU:
        return False
    from . import point_ops as P
    return P.bn_rows_supported(bn.num_features)


def _fused_bn_ok(cm, y):
    """Batch-statistics BatchNorm (+ ReLU or nothing) on contiguous fp32 CUDA rows of a supported width."""
    return y.is_contiguous() and _fused_bn_shape_ok(cm, y)


def as_rows(features):
    """(B,C,N) -> point-major (B,N,C) without a copy when the tensor already is a transposed
    view of rows (which is what the modules of this package hand to each other)."""
    rows = features.transpose(1, 2)
    return rows if rows.is_contiguous() else rows.contiguous()
