"""Seeded synthetic SUN-RGB-D-shaped inputs.

There is no dataset offline; every test and benchmark runs on these tensors.
All generators are CPU + torch.Generator so that the same seed gives the same
bytes here and on the GPU box.
"""
import math

import torch

# 4-level feature pyramids, (h, w) per level, 256 channels.
PYRAMIDS = {
    # 512x512 image, strides 8/16/32/64 -- headline reading of "4-level 512x512 img feats"
    "S512": ((64, 64), (32, 32), (16, 16), (8, 8)),
    # SUN RGB-D at 800x1088 after Pad(32) (configs/demf/demf_votenet.py:194-197), strides 8..64
    "REAL": ((100, 136), (50, 68), (25, 34), (13, 17)),
    # level-0 = 512x512: the HBM-resident reading
    "XL": ((512, 512), (256, 256), (128, 128), (64, 64)),
}
PYRAMID_IMAGE = {"S512": (512, 512), "REAL": (800, 1088), "XL": (4096, 4096)}


def pyramid_tokens(name):
    return sum(h * w for h, w in PYRAMIDS[name])


def make_points(B, N=20000, seed=0, clustered=False):
    """(B,N,4) float32: x,y~U(-3,3), z~U(0,2.5); 4th channel = height above the 1st percentile
    (shift_height=True, demf_votenet.py:188). `clustered` puts the points on a few random
    box surfaces + noise so that ball queries saturate nsample as real scans do."""
    g = torch.Generator().manual_seed(seed)
    if not clustered:
        xyz = torch.rand(B, N, 3, generator=g) * torch.tensor([6.0, 6.0, 2.5]) - torch.tensor(
            [3.0, 3.0, 0.0])
    else:
        xyz = torch.empty(B, N, 3)
        for b in range(B):
            nbox = 8
            centre = torch.rand(nbox, 3, generator=g) * torch.tensor([5.0, 5.0, 1.5]) - torch.tensor(
                [2.5, 2.5, 0.0])
            size = torch.rand(nbox, 3, generator=g) * 1.2 + 0.3
            which = torch.randint(0, nbox, (N,), generator=g)
            u = torch.rand(N, 3, generator=g) - 0.5
            face = torch.randint(0, 3, (N,), generator=g)
            sign = torch.randint(0, 2, (N,), generator=g).float() - 0.5
            u[torch.arange(N), face] = sign  # snap one coordinate to a face
            xyz[b] = centre[which] + u * size[which] + 0.01 * torch.randn(N, 3, generator=g)
    floor = torch.quantile(xyz[..., 2], 0.01, dim=1, keepdim=True)
    height = xyz[..., 2] - floor
    return torch.cat([xyz, height.unsqueeze(-1)], -1).contiguous()


def make_pyramid(B, name="S512", channels=256, seed=0):
    """List of 4 (B,channels,h,w) float32 ~N(0,1): stands in for the frozen image branch output
    (demf/modeling/detectors/demfnet.py:124-132)."""
    g = torch.Generator().manual_seed(seed + 7919)
    return [torch.randn(B, channels, h, w, generator=g) for h, w in PYRAMIDS[name]]


def make_img_metas(B, name="S512", seed=0, augment=True):
    """img_metas with the keys DeMFVoteHead.get_reference_points / prepare_decoder_inputs read
    (class_agnostic_vote_head.py:524-568): SUN RGB-D-like depth2img, 3D aug flow, 2D scale."""
    g = torch.Generator().manual_seed(seed + 104729)
    H, W = PYRAMID_IMAGE[name]
    metas = []
    for _ in range(B):
        # SUN RGB-D Kinect intrinsics are ~529.5 px focal at 730x530; scale to this image.
        scale = W / 730.0
        fx = 529.5 * scale
        K = torch.tensor([[fx, 0.0, W / 2.0], [0.0, fx, H / 2.0], [0.0, 0.0, 1.0]])
        # depth (x right, y forward, z up) -> camera (x right, y down, z forward), small tilt
        tilt = (torch.rand(1, generator=g).item() - 0.5) * 0.2
        Rt = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
        c, s = math.cos(tilt), math.sin(tilt)
        Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
        depth2img = K @ Rx @ Rt
        meta = dict(
            img_shape=(H, W, 3), batch_input_shape=(H, W), ori_shape=(H, W, 3),
            depth2img=depth2img.tolist(),
            scale_factor=[1.0, 1.0, 1.0, 1.0], flip=False, img_crop_offset=[0.0, 0.0],
            transformation_3d_flow=[], pcd_trans=[0.0, 0.0, 0.0], pcd_scale_factor=1.0,
            pcd_rotation=torch.eye(3).tolist(), pcd_horizontal_flip=False,
            pcd_vertical_flip=False)
        if augment:
            ang = (torch.rand(1, generator=g).item() * 2 - 1) * 0.523599
            ca, sa = math.cos(ang), math.sin(ang)
            meta["pcd_rotation"] = [[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]]
            meta["pcd_scale_factor"] = 0.85 + 0.3 * torch.rand(1, generator=g).item()
            meta["pcd_horizontal_flip"] = bool(torch.rand(1, generator=g).item() < 0.5)
            meta["transformation_3d_flow"] = ["HF", "R", "S", "T"]
        metas.append(meta)
    return metas


def make_msda_inputs(B=8, Q=256, H=8, D=32, name="S512", P=4, seed=0, shapes=None):
    """Kernel-level MSDA inputs: value ~N(0,1); loc = ref~U(0,1) + off~N(0,0.05) (a few samples
    fall outside [0,1] and exercise the zero padding); weights = softmax of N(0,1) logits."""
    g = torch.Generator().manual_seed(seed + 15485863)
    shapes = tuple(shapes) if shapes is not None else PYRAMIDS[name]
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    spatial_shapes = torch.tensor(shapes, dtype=torch.int64)
    level_start_index = torch.cat(
        [spatial_shapes.new_zeros(1), spatial_shapes.prod(1).cumsum(0)[:-1]])
    value = torch.randn(B, S, H, D, generator=g)
    ref = torch.rand(B, Q, 1, 1, 1, 2, generator=g)
    loc = (ref + 0.05 * torch.randn(B, Q, H, L, P, 2, generator=g)).contiguous()
    attn = torch.softmax(torch.randn(B, Q, H, L * P, generator=g), -1).view(B, Q, H, L, P)
    return value, spatial_shapes, level_start_index, loc, attn.contiguous()


def make_box_predictions(B, K, seed, classes=10):
    """Decoded head outputs for post-processing tests and timings: (B,K,7) gravity-centre boxes clustered
    around a dozen objects per scene (so that NMS has work to do), objectness (B,K) ~U(0,1) and semantic
    probabilities (B,K,classes) peaked at the object's class."""
    g = torch.Generator().manual_seed(seed)
    nobj = 12
    centre = torch.rand(B, nobj, 3, generator=g) * torch.tensor([5.0, 5.0, 1.5]) - torch.tensor([2.5, 2.5, 0.0])
    size = torch.rand(B, nobj, 3, generator=g) * 1.2 + 0.3
    yaw = torch.rand(B, nobj, generator=g) * 2 * math.pi
    which = torch.randint(0, nobj, (B, K), generator=g)
    take = lambda t: torch.gather(t, 1, which[..., None].expand(-1, -1, t.shape[-1]))  # noqa: E731
    box = torch.cat([take(centre) + 0.08 * torch.randn(B, K, 3, generator=g),
                     take(size) * (1 + 0.1 * torch.randn(B, K, 3, generator=g)).clamp(0.5, 1.5),
                     (torch.gather(yaw, 1, which) + 0.1 * torch.randn(B, K, generator=g))[..., None] % (2 * math.pi)], -1)
    obj = torch.rand(B, K, generator=g)
    sem = torch.softmax(torch.randn(B, K, classes, generator=g) + 3 * torch.nn.functional.one_hot(
        which % classes, classes), -1)
    return box, obj, sem


def make_points_in_boxes(B, N, box, seed):
    """(B,N,3) points: half uniform in the room, half inside boxes drawn from the first third of `box`
    (the rest of the boxes stay nearly empty and are filtered by the > 5 points rule)."""
    g = torch.Generator().manual_seed(seed + 99)
    K = box.shape[1]
    uni = torch.rand(B, N // 2, 3, generator=g) * torch.tensor([6.0, 6.0, 2.5]) - torch.tensor([3.0, 3.0, 0.0])
    which = torch.randint(0, K // 3, (B, N - N // 2), generator=g)       # only the first third of the boxes
    bsel = torch.gather(box, 1, which[..., None].expand(-1, -1, 7))
    u = (torch.rand(B, N - N // 2, 3, generator=g) - 0.5) * bsel[..., 3:6]
    c, s = torch.cos(bsel[..., 6]), torch.sin(bsel[..., 6])
    local = torch.stack([u[..., 0] * c + u[..., 1] * s, -u[..., 0] * s + u[..., 1] * c, u[..., 2]], -1)
    return torch.cat([uni, local + bsel[..., :3]], 1).contiguous()
