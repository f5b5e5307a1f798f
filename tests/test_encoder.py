"""Image-branch encoder: DeformableDetrEncoder against a straight-line
restatement of the reference's forward (demf/modeling/layers/deform_detr_encoder.py:69-157 +
mmcv BaseTransformerLayer / MultiScaleDeformableAttention / FFN arithmetic) written with plain torch
functions and the oracle's grid_sample MSDA; on the GPU the same module runs the CUDA MSDA kernel
and is compared with the CPU result."""
import math

import pytest
import torch
import torch.nn.functional as F

import demf_b200  # noqa: F401
from demf_b200 import engine
from demf_b200.mm.config import Config
from demf_b200.mm.registry import (HEADS, POSITIONAL_ENCODING, TRANSFORMER_LAYER_SEQUENCE, build_head,
                                   build_positional_encoding)
from oracle import msda_torch
from oracle.cpu_backend import oracle_ops

SHAPES = ((12, 16), (6, 8), (3, 4), (2, 2))


def _encoder(num_layers=2, seed=0):
    torch.manual_seed(seed)
    cfg = Config.fromfile(engine.CONFIG).img_encoder_cfg.to_dict()
    cfg["encoder"]["num_layers"] = num_layers
    enc = build_head(cfg)
    enc.init_weights()
    # init_weights leaves offsets/attention weights input-independent (zero weight): perturb every
    # parameter so the test exercises the data-dependent paths
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in enc.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return enc.eval()


def _inputs(B=2, seed=0, padded=True):
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(B, 256, h, w, generator=g) for h, w in SHAPES]
    H, W = SHAPES[0][0] * 8, SHAPES[0][1] * 8
    metas = [dict(batch_input_shape=(H, W), img_shape=(H, W, 3)) for _ in range(B)]
    if padded:
        metas[0]["img_shape"] = (H - 24, W - 40, 3)
        if B > 1:
            metas[1]["img_shape"] = (H, W - 16, 3)
    return feats, metas


# ------------------------------------------------------------------ restatement ----
def _sine_encoding(mask, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6, offset=-0.5):
    """Per-pixel loops over mmdet's SinePositionalEncoding formula (normalize=True)."""
    B, H, W = mask.shape
    out = torch.zeros(B, 2 * num_feats, H, W)
    for b in range(B):
        valid = ~mask[b]
        for y in range(H):
            for x in range(W):
                ye = float(valid[: y + 1, x].sum())
                xe = float(valid[y, : x + 1].sum())
                yn = torch.tensor((ye + offset) / (float(valid[:, x].sum()) + eps) * scale, dtype=torch.float32)
                xn = torch.tensor((xe + offset) / (float(valid[y, :].sum()) + eps) * scale, dtype=torch.float32)
                for i in range(num_feats):
                    d = torch.tensor(float(temperature), dtype=torch.float32) ** (2 * (i // 2) / num_feats)
                    fn = torch.sin if i % 2 == 0 else torch.cos
                    out[b, i, y, x] = fn(yn / d)
                    out[b, num_feats + i, y, x] = fn(xn / d)
    return out


def _restated_forward(enc, feats, metas):
    """The reference's forward, one statement at a time, no caching, no fused projections."""
    B = feats[0].shape[0]
    H, W = metas[0]["batch_input_shape"]
    img_masks = torch.ones(B, H, W)
    for i, m in enumerate(metas):
        h, w, _ = m["img_shape"]
        img_masks[i, :h, :w] = 0
    masks = [F.interpolate(img_masks[None], size=f.shape[-2:]).to(torch.bool).squeeze(0) for f in feats]
    pos = [enc.positional_encoding(m) for m in masks]
    feat_flat, mask_flat, pos_flat, shapes = [], [], [], []
    for lvl, (f, m, p) in enumerate(zip(feats, masks, pos)):
        shapes.append(f.shape[-2:])
        feat_flat.append(f.flatten(2).transpose(1, 2))
        mask_flat.append(m.flatten(1))
        pos_flat.append(p.flatten(2).transpose(1, 2) + enc.level_embeds[lvl].view(1, 1, -1))
    x = torch.cat(feat_flat, 1)            # (B,S,C) batch-first here
    mask = torch.cat(mask_flat, 1)
    qpos = torch.cat(pos_flat, 1)
    valid = []
    for m in masks:
        _, h, w = m.shape
        valid.append(torch.stack([(~m[:, 0, :]).sum(1).float() / w, (~m[:, :, 0]).sum(1).float() / h], -1))
    valid = torch.stack(valid, 1)          # (B,L,2)
    refs = []
    for lvl, (h, w) in enumerate(shapes):
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        ry = ys.reshape(-1)[None] / (valid[:, None, lvl, 1] * h)
        rx = xs.reshape(-1)[None] / (valid[:, None, lvl, 0] * w)
        refs.append(torch.stack((rx, ry), -1))
    ref = torch.cat(refs, 1)[:, :, None] * valid[:, None]          # (B,S,L,2)
    shapes_t = torch.tensor([list(s) for s in shapes])
    normalizer = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()
    for layer in enc.encoder.layers:
        att, ffn, (n1, n2) = layer.attentions[0], layer.ffns[0], layer.norms
        q = x + qpos
        value = F.linear(x, att.value_proj.weight, att.value_proj.bias)
        value = value.masked_fill(mask[..., None], 0.0).view(B, -1, att.num_heads, 256 // att.num_heads)
        off = F.linear(q, att.sampling_offsets.weight, att.sampling_offsets.bias).view(
            B, -1, att.num_heads, att.num_levels, att.num_points, 2)
        w = F.linear(q, att.attention_weights.weight, att.attention_weights.bias).view(
            B, -1, att.num_heads, att.num_levels * att.num_points).softmax(-1).view(
            B, -1, att.num_heads, att.num_levels, att.num_points)
        loc = ref[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
        o = msda_torch.multi_scale_deformable_attn_pytorch(value, shapes_t, loc, w)
        x = F.layer_norm(x + F.linear(o, att.output_proj.weight, att.output_proj.bias), (256,), n1.weight, n1.bias)
        fc1, fc2 = ffn.layers[0][0], ffn.layers[1]
        x = F.layer_norm(x + F.linear(F.relu(F.linear(x, fc1.weight, fc1.bias)), fc2.weight, fc2.bias),
                         (256,), n2.weight, n2.bias)
    mem = x.permute(0, 2, 1)
    outs, start = [], 0
    for h, w in shapes:
        outs.append(mem[:, :, start:start + h * w].reshape(B, 256, h, w))
        start += h * w
    return outs, dict(masks=masks, valid=valid, ref=ref)


# --------------------------------------------------------------------- CPU tests ----
def test_registered_names_and_reference_config():
    assert "DeformableDetrEncoder" in HEADS
    assert "DetrTransformerEncoder" in TRANSFORMER_LAYER_SEQUENCE
    assert "TransformerLayerSequence" in TRANSFORMER_LAYER_SEQUENCE
    assert "SinePositionalEncoding" in POSITIONAL_ENCODING
    cfg = Config.fromfile(engine.CONFIG).img_encoder_cfg          # configs/demf/demf_votenet.py:28-47
    assert cfg.encoder.type == "DetrTransformerEncoder" and cfg.encoder.num_layers == 6
    tl = cfg.encoder.transformerlayers
    assert tuple(tl.operation_order) == ("self_attn", "norm", "ffn", "norm")
    assert tl.feedforward_channels == 1024 and tl.attn_cfgs.type == "MultiScaleDeformableAttention"
    pe = cfg.positional_encoding
    assert (pe.num_feats, pe.normalize, pe.offset) == (128, True, -0.5)


def test_parameter_names_match_a_deformable_detr_checkpoint():
    enc = build_head(Config.fromfile(engine.CONFIG).img_encoder_cfg.to_dict())
    keys = set(enc.state_dict())
    assert "level_embeds" in keys and enc.level_embeds.shape == (4, 256)
    for i in (0, 5):
        for k in ("attentions.0.sampling_offsets.weight", "attentions.0.attention_weights.bias",
                  "attentions.0.value_proj.weight", "attentions.0.output_proj.bias",
                  "ffns.0.layers.0.0.weight", "ffns.0.layers.1.bias", "norms.0.weight", "norms.1.bias"):
            assert f"encoder.layers.{i}.{k}" in keys
    assert len(keys) == 1 + 6 * 16
    assert enc.encoder.post_norm is None          # post-norm layers: no final LayerNorm
    assert sum(p.numel() for p in enc.parameters()) == 4 * 256 + 6 * (
        256 * 256 * 2 + 256 * 2 + 256 * 256 + 256 + 128 * 256 + 128 + 2 * 256 * 1024 + 1024 + 256 + 4 * 256)


def test_sine_positional_encoding_known_answer():
    pe = build_positional_encoding(dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5))
    mask = torch.zeros(2, 3, 4, dtype=torch.bool)
    mask[0, :, 3] = True        # one padded column
    mask[1, 2, :] = True        # one padded row
    got = pe(mask)
    assert got.shape == (2, 256, 3, 4)
    torch.testing.assert_close(got, _sine_encoding(mask), atol=2e-6, rtol=0)
    # first valid pixel, channel 0: sin((1 - 0.5) / (n_valid + eps) * 2 pi)
    assert got[0, 0, 0, 0].item() == pytest.approx(math.sin(0.5 / (3 + 1e-6) * 2 * math.pi), abs=1e-6)


def test_geometry_follows_the_reference_statements():
    enc = _encoder(num_layers=1)
    feats, metas = _inputs()
    _, ref = _restated_forward(enc, feats, metas)
    geo = enc._geometry(SHAPES, metas, torch.device("cpu"))
    for a, b in zip(geo["masks"], ref["masks"]):
        assert torch.equal(a, b)
    assert torch.equal(geo["valid_ratios"], ref["valid"])
    assert torch.equal(geo["reference_points"], ref["ref"])
    assert geo["spatial_shapes"].tolist() == [list(s) for s in SHAPES]
    assert geo["level_start_index"].tolist() == [0, 192, 240, 252]
    assert geo["any_padding"]
    # padded image 0: 24 of 96 rows and 40 of 128 columns masked at stride 8
    assert geo["valid_ratios"][0, 0].tolist() == pytest.approx([11 / 16, 9 / 12])
    assert enc._geometry(SHAPES, metas, torch.device("cpu")) is geo          # cached


@pytest.mark.parametrize("padded", [True, False])
def test_encoder_equals_restated_reference_forward(padded):
    enc = _encoder(num_layers=2)
    feats, metas = _inputs(padded=padded)
    with torch.no_grad(), oracle_ops():
        got = enc(feats, metas)
        want, _ = _restated_forward(enc, feats, metas)
    assert [tuple(o.shape) for o in got] == [tuple(f.shape) for f in feats]
    for g, w in zip(got, want):
        torch.testing.assert_close(g, w, atol=2e-5, rtol=0)
    assert not torch.allclose(got[0], feats[0], atol=1e-2)


def test_reference_call_convention_of_transformer():
    """`transformer(mlvl_feats, mlvl_masks, mlvl_pos_embeds)` with caller-built masks and encodings
    (deform_detr_encoder.py:95-99) gives the same pyramid as forward()."""
    enc = _encoder(num_layers=1)
    feats, metas = _inputs()
    with torch.no_grad(), oracle_ops():
        got = enc(feats, metas)
        geo = enc._geometry(SHAPES, metas, torch.device("cpu"))
        pos = [enc.positional_encoding(m) for m in geo["masks"]]
        again = enc.transformer(feats, geo["masks"], pos)
    for a, b in zip(got, again):
        torch.testing.assert_close(a, b, atol=1e-6, rtol=0)


def test_detector_builds_and_loads_stage1_checkpoint_keys():
    model = engine.build_demf_votenet(img_encoder=True)
    assert model.with_img_encoder
    assert not any(p.requires_grad for p in model.img_encoder.parameters())     # frozen branch
    model.train()
    assert not model.img_encoder.training                                       # demfnet.py:103-112
    # a stage-1 (Deformable-DETR) checkpoint keeps the encoder under img_bbox_head.transformer.*
    sd = model.state_dict()
    stage1 = {}
    for k, v in sd.items():
        if k.startswith("img_encoder."):
            stage1[k.replace("img_encoder", "img_bbox_head.transformer")] = torch.full_like(v, 0.25)
        else:
            stage1[k] = v
    stage1["img_bbox_head.cls_branches.0.weight"] = torch.zeros(3, 3)         # dropped by the rewrite
    model.load_state_dict(stage1, strict=True)
    assert float(model.img_encoder.level_embeds[0, 0]) == 0.25
    # the default model keeps the name-only placeholder and passes pyramids through
    assert not engine.build_demf_votenet().with_img_encoder


def test_detector_runs_pyramids_through_the_encoder():
    model = engine.build_demf_votenet(img_encoder=True).eval()
    feats, metas = _inputs()
    with torch.no_grad(), oracle_ops():
        out = model.extract_img_feat(feats, metas)
        want = model.img_encoder(feats, metas)
    for a, b in zip(out, want):
        assert torch.equal(a, b)


# --------------------------------------------------------------------- GPU tests ----
@pytest.mark.gpu
@pytest.mark.parametrize("padded", [True, False])
def test_encoder_gpu_matches_cpu_oracle(padded):
    """CUDA MSDA kernel in the self-attention regime (Q = S) through two encoder layers; fp32 GEMMs.
    Tolerance 1e-4 absolute on LayerNorm-scaled outputs (north_star's MSDA tolerance)."""
    engine.set_gemm_precision("fp32")
    try:
        enc = _encoder(num_layers=2)
        feats, metas = _inputs(B=2, padded=padded)
        with torch.no_grad(), oracle_ops():
            want = enc(feats, metas)
        enc_gpu = _encoder(num_layers=2).cuda()
        with torch.no_grad():
            got = enc_gpu([f.cuda() for f in feats], metas)
        masks = enc._geometry(SHAPES, metas, torch.device("cpu"))["masks"]
        for g, w, m in zip(got, want, masks):
            assert g.is_cuda and torch.isfinite(g).all()
            # Compared on the valid pixels. Under the padding mask the sine encoding divides by
            # (0 + eps) wherever a whole row or column is padding: arguments of ~3e6 rad, where one
            # ulp of `temperature ** k` (pow differs between libm and CUDA) moves sin() by O(1e-2).
            # Those pixels are zeroed as values in every layer and never reach a valid output.
            keep = ~m[:, None].expand_as(w)
            torch.testing.assert_close(g.cpu()[keep], w[keep], atol=1e-4, rtol=0)
    finally:
        engine.set_gemm_precision("tf32")


@pytest.mark.gpu
@pytest.mark.parametrize("padded", [True, False])
def test_encoder_rows_path_equals_layer_sequence(padded):
    """The inference path on token rows (fused projections, MSDA from projections, bias + residual +
    LayerNorm kernel) against the module-by-module layer sequence on the same device."""
    from demf_b200.modeling.encoder import DeformableDetrEncoder
    engine.set_gemm_precision("fp32")
    try:
        enc = _encoder(num_layers=3).cuda()
        feats, metas = _inputs(B=2, padded=padded)
        feats = [f.cuda() for f in feats]
        with torch.no_grad():
            assert enc._rows_path_ok(feats[0])
            fast = enc(feats, metas)
            DeformableDetrEncoder.fused_eval = False
            slow = enc(feats, metas)
    finally:
        DeformableDetrEncoder.fused_eval = True
        engine.set_gemm_precision("tf32")
    masks = enc._geometry(SHAPES, metas, feats[0].device)["masks"]
    for a, b, m in zip(fast, slow, masks):
        keep = ~m[:, None].expand_as(a)
        torch.testing.assert_close(a[keep], b[keep], atol=2e-5, rtol=0)


@pytest.mark.gpu
def test_encoder_gpu_full_size_properties():
    """Six layers at BASELINE's pyramid (S = 5440 tokens, batch 8): finite, deterministic, and
    independent of what lies under the padding mask (values there are zeroed before sampling)."""
    from demf_b200 import synth
    torch.manual_seed(0)
    enc = build_head(Config.fromfile(engine.CONFIG).img_encoder_cfg.to_dict())
    enc.init_weights()
    enc = enc.cuda().eval()
    feats = [f.cuda() for f in synth.make_pyramid(8, "S512")]
    metas = synth.make_img_metas(8, "S512")
    for m in metas[:4]:
        m["img_shape"] = (448, 384, 3)
    with torch.no_grad():
        a = enc(feats, metas)
        b = enc(feats, metas)
        dirty = [f.clone() for f in feats]
        geo = enc._geometry(tuple(synth.PYRAMIDS["S512"]), metas, feats[0].device)
        for f, m in zip(dirty, geo["masks"]):
            f.masked_fill_(m[:, None], 1e3)
        c = enc(dirty, metas)
    for x, y, z, m in zip(a, b, c, geo["masks"]):
        assert torch.isfinite(x).all() and torch.equal(x, y)
        keep = ~m[:, None].expand_as(x)
        torch.testing.assert_close(x[keep], z[keep], atol=1e-4, rtol=0)
