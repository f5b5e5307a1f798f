"""ctypes front end of the CPU oracle (oracle/demf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of demf_oracle.c. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. PARITY UNPINNED (the reference ships no golden vectors).

All functions take and return CPU torch tensors with the shapes of the upstream
mmdet3d / mmcv op wrappers.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdemf_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/libdemf_oracle.so with gcc (no GPU, no torch needed)."""
    src = os.path.join(_HERE, "demf_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a CC without libgomp; the Makefile pins gcc
    subprocess.check_call(["make", "-C", _HERE, "-B", "libdemf_oracle.so"], env=env,
                          stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()  # no-op when libdemf_oracle.so is newer than demf_oracle.c
        _lib = ctypes.CDLL(_SO)
        _lib.demf_ref_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().demf_ref_num_threads())


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    assert t.dtype == torch.float32 and t.device.type == "cpu", (t.dtype, t.device)
    return t.contiguous()


def _i32(t):
    assert t.dtype == torch.int32 and t.device.type == "cpu"
    return t.contiguous()


def _check(rc, name):
    if rc != 0:
        raise RuntimeError(f"oracle {name} failed with code {rc}")


def furthest_point_sample(xyz, m):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = torch.zeros(B, m, dtype=torch.int32)
    _check(lib().demf_ref_fps(_p(xyz), B, N, m, _p(idx)), "fps")
    return idx


def ball_query(min_radius, max_radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    _check(lib().demf_ref_ball_query(_p(xyz), _p(new_xyz), B, N, M, ctypes.c_float(min_radius),
                                     ctypes.c_float(max_radius), nsample, _p(idx)), "ball_query")
    return idx


def grouping_operation(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    _, M, ns = idx.shape
    out = torch.empty(B, C, M, ns)
    _check(lib().demf_ref_group_fwd(_p(features), _p(idx), B, C, N, M, ns, _p(out)), "group_fwd")
    return out


def grouping_operation_backward(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M, ns = grad_out.shape
    g = torch.zeros(B, C, N)
    _check(lib().demf_ref_group_bwd(_p(grad_out), _p(idx), B, C, N, M, ns, _p(g)), "group_bwd")
    return g


def gather_points(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M)
    _check(lib().demf_ref_gather_fwd(_p(features), _p(idx), B, C, N, M, _p(out)), "gather_fwd")
    return out


def gather_points_backward(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M = grad_out.shape
    g = torch.zeros(B, C, N)
    _check(lib().demf_ref_gather_bwd(_p(grad_out), _p(idx), B, C, N, M, _p(g)), "gather_bwd")
    return g


def query_and_group(xyz, new_xyz, features, min_radius, max_radius, ns, use_xyz=True,
                    normalize_xyz=False):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0 if features is None else features.shape[1]
    if features is not None:
        features = _f32(features)
    idx = torch.zeros(B, M, ns, dtype=torch.int32)
    out = torch.empty(B, (3 if use_xyz else 0) + C, M, ns)
    _check(lib().demf_ref_query_and_group_fwd(
        _p(xyz), _p(features) if features is not None else None, _p(new_xyz), B, N, M, C,
        ctypes.c_float(min_radius), ctypes.c_float(max_radius), ns, int(use_xyz),
        int(normalize_xyz), _p(idx), _p(out)), "query_and_group")
    return idx, out


def three_nn(unknown, known):
    """Returns (dist, idx) with dist = sqrt(dist2), exactly like the upstream wrapper."""
    dist2, idx = three_nn_squared(unknown, known)
    return torch.sqrt(dist2), idx


def three_nn_squared(unknown, known):
    """The kernel's own outputs: squared distances and indices."""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(B, n, 3)
    idx = torch.empty(B, n, 3, dtype=torch.int32)
    _check(lib().demf_ref_three_nn(_p(unknown), _p(known), B, n, m, _p(dist2), _p(idx)),
           "three_nn")
    return dist2, idx


def three_interpolate(features, idx, weight):
    features, idx, weight = _f32(features), _i32(idx), _f32(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = torch.empty(B, C, n)
    _check(lib().demf_ref_three_interpolate_fwd(_p(features), _p(idx), _p(weight), B, C, m, n,
                                                _p(out)), "three_interpolate_fwd")
    return out


def three_interpolate_backward(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    g = torch.zeros(B, C, m)
    _check(lib().demf_ref_three_interpolate_bwd(_p(grad_out), _p(idx), _p(weight), B, C, n, m,
                                                _p(g)), "three_interpolate_bwd")
    return g


def _msda_dims(value, sampling_locations):
    B, S, H, D = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    return B, S, H, D, Q, L, P


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations,
                           attention_weights):
    value, loc, aw = _f32(value), _f32(sampling_locations), _f32(attention_weights)
    shapes = spatial_shapes.to(torch.int64).contiguous()
    lsi = level_start_index.to(torch.int64).contiguous()
    B, S, H, D, Q, L, P = _msda_dims(value, loc)
    out = torch.empty(B, Q, H * D)
    _check(lib().demf_ref_msda_fwd(_p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), B, S, H, D, Q,
                                   L, P, _p(out)), "msda_fwd")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_locations,
                            attention_weights, grad_output):
    value, loc, aw = _f32(value), _f32(sampling_locations), _f32(attention_weights)
    go = _f32(grad_output)
    shapes = spatial_shapes.to(torch.int64).contiguous()
    lsi = level_start_index.to(torch.int64).contiguous()
    B, S, H, D, Q, L, P = _msda_dims(value, loc)
    gv = torch.zeros_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(aw)
    _check(lib().demf_ref_msda_bwd(_p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), _p(go), B, S,
                                   H, D, Q, L, P, _p(gv), _p(gl), _p(ga)), "msda_bwd")
    return gv, gl, ga
