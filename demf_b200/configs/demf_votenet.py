# Model section of the reference's configs/demf/demf_votenet.py:26-182 (DeMF on VoteNet, SUN RGB-D),
# restated value for value so that the same `model = dict(type='DeMFVoteNet', ...)` builds here.
# BASELINE.json's configs feed synthetic 4-level pyramids in place of the frozen image branch, so the
# `model` below leaves img_backbone / img_neck / img_encoder out (the detector then takes `img` as the
# pyramid itself). The branch IS implemented: `img_backbone_cfg` / `img_neck_cfg` restate
# configs/deformdetr/imvotenet_image.py:3-20 and `img_encoder_cfg` configs/demf/demf_votenet.py:28-47;
# demf_b200.engine.build_demf_votenet(img_encoder=True) adds the encoder (pyramids handed in are then
# the neck's output), build_demf_votenet(img_branch=True) the whole branch (`img` = (B,3,H,W) images).
# The reference's own config file builds unchanged (tests/test_model_cpu.py).
# `num_points` of the deformable cross attention is 2 in the reference file (line 83); the
# BASELINE.json configs quote 4 (the mmcv default) -- override with
# Config.merge_from_dict({'model.pts_bbox_head.decoder.transformerlayers.attn_cfgs.1.num_points': 4})
# or demf_b200.engine.build_demf_votenet(num_points=4).

class_names = ('bed', 'table', 'sofa', 'chair', 'toilet', 'desk', 'dresser', 'night_stand',
               'bookshelf', 'bathtub')

lr = 0.008
optimizer = dict(
    type='AdamW', lr=lr, weight_decay=0.01,
    paramwise_cfg=dict(custom_keys={'decoder': dict(lr_mult=0.05, decay_mult=1.0)}))
optimizer_config = dict(grad_clip=dict(max_norm=10, norm_type=2))

_mean_sizes = [[2.114256, 1.620300, 0.927272], [0.791118, 1.279516, 0.718182],
               [0.923508, 1.867419, 0.845495], [0.591958, 0.552978, 0.827272],
               [0.699104, 0.454178, 0.75625], [0.69519, 1.346299, 0.736364],
               [0.528526, 1.002642, 1.172878], [0.500618, 0.632163, 0.683424],
               [0.404671, 1.071108, 1.688889], [0.76584, 1.398258, 0.472728]]

img_encoder_cfg = dict(
    type='DeformableDetrEncoder',
    encoder=dict(
        type='DetrTransformerEncoder',
        num_layers=6,
        transformerlayers=dict(
            type='BaseTransformerLayer',
            attn_cfgs=dict(type='MultiScaleDeformableAttention', embed_dims=256),
            feedforward_channels=1024,
            ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'ffn', 'norm'))),
    positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True, offset=-0.5),
    num_feature_levels=4,
    embed_dims=256)

img_backbone_cfg = dict(
    type='ResNet', depth=50, num_stages=4, out_indices=(1, 2, 3), frozen_stages=1,
    norm_cfg=dict(type='BN', requires_grad=False), norm_eval=True, style='pytorch')
img_neck_cfg = dict(
    type='ChannelMapper', in_channels=[512, 1024, 2048], kernel_size=1, out_channels=256, act_cfg=None,
    norm_cfg=dict(type='GN', num_groups=32), num_outs=4)

model = dict(
    type='DeMFVoteNet',
    pts_backbone=dict(
        type='PointNet2SASSG',
        in_channels=4,
        num_points=(2048, 1024, 512, 256),
        radius=(0.2, 0.4, 0.8, 1.2),
        num_samples=(64, 32, 16, 16),
        sa_channels=((64, 64, 128), (128, 128, 256), (128, 128, 256), (128, 128, 256)),
        fp_channels=((256, 256), (256, 256)),
        norm_cfg=dict(type='BN2d'),
        sa_cfg=dict(type='PointSAModule', pool_mod='max', use_xyz=True, normalize_xyz=True)),
    pts_bbox_head=dict(
        type='DeMFVoteHead',
        pred_layer_cfg=dict(in_channels=256, shared_conv_channels=(128, 128), bias=True,
                            conv_pred_layers=2),
        decoder=dict(
            type='DeMFTransformerDecoderLayer',
            num_layers=1,
            transformerlayers=dict(
                type='DetrTransformerDecoderLayer',
                attn_cfgs=[
                    dict(type='MultiheadAttention', embed_dims=256, num_heads=8, dropout=0.4),
                    dict(type='MultiScaleDeformableAttention', num_heads=8, num_levels=4,
                         num_points=2, dropout=0.4, embed_dims=256),
                ],
                feedforward_channels=1024,
                ffn_dropout=0.1,
                operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')),
            posembed=dict(input_channel=6, num_pos_feats=256)),
        num_classes=10,
        bbox_coder=dict(type='DeMFClassAgnosticBBoxCoder', num_dir_bins=12, with_rot=True,
                        num_sizes=10, mean_sizes=_mean_sizes),
        conv_cfg=dict(type='Conv1d'),
        norm_cfg=dict(type='BN1d'),
        objectness_loss=dict(type='CrossEntropyLoss', class_weight=[0.2, 0.8], reduction='sum',
                             loss_weight=5.0),
        dir_class_loss=dict(type='CrossEntropyLoss', reduction='sum', loss_weight=1.0),
        dir_res_loss=dict(type='SmoothL1Loss', reduction='sum', loss_weight=10.0),
        size_class_loss=dict(type='CrossEntropyLoss', reduction='sum', loss_weight=1.0),
        size_res_loss=dict(type='SmoothL1Loss', reduction='sum', loss_weight=10.0, beta=0.0625),
        center_loss=dict(type='SmoothL1Loss', beta=1.0 / 9.0, reduction='sum', loss_weight=10.0),
        iou_loss=dict(type='AxisAlignedIoULoss', reduction='sum', loss_weight=12.0 / 3.0),
        semantic_loss=dict(type='CrossEntropyLoss', reduction='sum', loss_weight=1.0),
        vote_module_cfg=dict(
            in_channels=256, vote_per_seed=1, gt_per_seed=3, conv_channels=(256, 256),
            conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'), norm_feats=True,
            vote_loss=dict(type='ChamferDistance', mode='l1', reduction='none',
                           loss_dst_weight=10.0)),
        vote_aggregation_cfg=dict(type='PointSAModule', num_point=256, radius=0.3, num_sample=16,
                                  mlp_channels=[256, 256, 256, 256], use_xyz=True,
                                  normalize_xyz=True)),
    num_sampled_seed=1024,
    freeze_img_branch=True,
    train_cfg=dict(pts=dict(pos_distance_thr=0.3, neg_distance_thr=0.6, sample_mod='seed')),
    test_cfg=dict(
        img_rcnn=dict(score_thr=0.1),
        pts=dict(ensemble_layers=[0, 1], sample_mod='seed', nms_thr=0.25, score_thr=0.05,
                 per_class_proposal=True)))

find_unused_parameters = True
