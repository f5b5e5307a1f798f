"""Brute-force numpy second formulations of the point ops.

TEST INFRASTRUCTURE ONLY (see oracle/demf_oracle.c). PARITY UNPINNED.

Independent of the C oracle's loop structure: they build the full distance matrix
and use vectorised selection (cumsum-of-mask first-k, stable argsort top-3, argmax).
float32 fused multiply-add is emulated through float64 (a float32 product is exact
in float64; the one remaining double rounding is a ~2^-29 event per operation), in
the nvcc contraction order documented in demf_oracle.c.
"""
import numpy as np


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def sqdist_matrix(a, b):
    """a (n,3), b (m,3) float32 -> (n,m) float32 of |a_i - b_j|^2 in the CUDA rounding order."""
    d = a[:, None, :].astype(np.float32) - b[None, :, :].astype(np.float32)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return _fma32(dz, dz, _fma32(dx, dx, (dy * dy).astype(np.float32)))


def ball_query(min_radius, max_radius, nsample, xyz, new_xyz):
    """xyz (B,N,3), new_xyz (B,M,3) -> (B,M,nsample) int32; first nsample hits in index order."""
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = np.zeros((B, M, nsample), np.int32)
    r0 = np.float32(min_radius) * np.float32(min_radius)
    r1 = np.float32(max_radius) * np.float32(max_radius)
    for b in range(B):
        d2 = sqdist_matrix(new_xyz[b], xyz[b])
        hit = (d2 == 0) | ((d2 >= r0) & (d2 < r1))
        rank = np.cumsum(hit, axis=1) - 1  # rank of each hit within its row
        cnt = hit.sum(1)
        for m in range(M):
            if cnt[m] == 0:
                continue
            ks = np.nonzero(hit[m] & (rank[m] < nsample))[0]
            out[b, m, :] = ks[0]
            out[b, m, :len(ks)] = ks
    return out


def three_nn(unknown, known):
    """-> (dist (B,n,3) float32 = sqrt(d2), idx (B,n,3) int32); ties resolved by lower index."""
    B, n, _ = unknown.shape
    dist = np.zeros((B, n, 3), np.float32)
    idx = np.zeros((B, n, 3), np.int32)
    for b in range(B):
        d2 = sqdist_matrix(unknown[b], known[b])
        order = np.argsort(d2, axis=1, kind="stable")[:, :3]
        k = order.shape[1]
        idx[b, :, :k] = order
        dist[b, :, :k] = np.sqrt(np.take_along_axis(d2, order, 1))
        if k < 3:  # fewer than three known points: upstream leaves best=1e40, idx=0
            dist[b, :, k:] = np.float32(np.inf)
    return dist, idx


def furthest_point_sample(xyz, m):
    """Index-order tie rule (argmax = first maximum). Equals the upstream block-tree rule
    whenever the running-min distances have a unique maximum, i.e. without duplicate points."""
    B, N, _ = xyz.shape
    out = np.zeros((B, m), np.int32)
    for b in range(B):
        temp = np.full(N, 1e10, np.float32)
        old = 0
        for j in range(1, m):
            d = sqdist_matrix(xyz[b], xyz[b, old:old + 1])[:, 0]
            temp = np.minimum(temp, d)
            old = int(np.argmax(temp))
            out[b, j] = old
    return out
