/*
 * demf_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or as the
 * reported CPU baseline. Nothing under demf_b200/ imports, links or executes it.
 *
 * PARITY UNPINNED: haoy945/DeMF ships no native code, no tests and no golden
 * vectors. The arithmetic of its hot path lives in pinned,
 * un-vendored dependencies (requirements.txt:2-4):
 *     mmdet3d == 0.18.1   ops/{furthest_point_sample,ball_query,group_points,
 *                              gather_points,interpolate}/src/<op>_cuda.cu
 *     mmcv_full == 1.3.18 ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh
 * neither of which is present under /root/reference or installable offline.
 * This file restates the published algorithm of those kernels; each function
 * cites the upstream kernel it follows and the reference call site that reaches
 * it. Floating-point contraction follows what nvcc emits for the upstream source
 * expressions (probed with nvcc 12.9, sm_80 and sm_100a):
 *     (a-x)*(a-x)+(b-y)*(b-y)+(c-z)*(c-z)  ->  fmaf(dz,dz, fmaf(dx,dx, dy*dy))
 *     w1*v1+w2*v2+w3*v3+w4*v4              ->  fmaf(w4,v4, fmaf(w3,v3, fmaf(w1,v1, w2*v2)))
 *     w0*p0+w1*p1+w2*p2                    ->  fmaf(w2,p2, fmaf(w0,p0, w1*p1))
 * Build with -ffp-contract=off so that the compiler adds no contraction of its own.
 *
 * Host-pointer twins of include/demf_b200.h (same argument order, no stream).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define DEMF_API __attribute__((visibility("default")))

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  /* a - b, in the operand order the caller's upstream kernel uses */
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

DEMF_API int demf_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* -------------------------------------------------------------------------
 * FPS: mmdet3d furthest_point_sample_cuda.cu, furthest_point_sampling_kernel
 * <block_size>; reference call sites class_agnostic_vote_head.py:429-430 and
 * the SA modules of configs/demf/demf_votenet.py:48-62.
 * One block per scene with T = min(1024, 2^floor(log2 N)) threads. Thread t
 * scans k = t, t+T, ... keeping the first strict maximum of
 * temp[k] = min(temp[k], d(k, old)); a shared-memory tree (stride T/2 .. 1)
 * keeps the lower slot unless the upper one is strictly greater. The emulation
 * below reproduces exactly that order, hence the tie behaviour for duplicate
 * points. temp starts at 1e10 (upstream Python wrapper).
 * ------------------------------------------------------------------------- */
static int fps_block_threads(int N) {
  int p = 1;
  while (p * 2 <= N && p * 2 <= 1024) p *= 2;
  return p;
}

DEMF_API int demf_ref_fps(const float* xyz, int B, int N, int m, int32_t* idx) {
  if (!xyz || !idx) return -1;
  if (B < 0 || N <= 0 || m < 0) return -2;
  if (m == 0 || B == 0) return 0;
  const int T = fps_block_threads(N);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float* p = xyz + (size_t)b * N * 3;
    int32_t* out = idx + (size_t)b * m;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    float* dists = (float*)malloc(sizeof(float) * (size_t)T);
    int* dists_i = (int*)malloc(sizeof(int) * (size_t)T);
    for (int k = 0; k < N; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int t = 0; t < T; ++t) {
        dists[t] = -1.0f;
        dists_i[t] = 0;
      }
      for (int k = 0; k < N; ++k) { /* k in increasing order == each thread's own order */
        const int t = k & (T - 1);
        const float d = sqdist(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
        const float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > dists[t]) {
          dists[t] = d2;
          dists_i[t] = k;
        }
      }
      for (int s = T / 2; s >= 1; s /= 2) {
        for (int t = 0; t < s; ++t) {
          const float v1 = dists[t], v2 = dists[t + s];
          if (v2 > v1) {
            dists[t] = v2;
            dists_i[t] = dists_i[t + s];
          }
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
    free(temp);
    free(dists);
    free(dists_i);
  }
  return 0;
}

/* -------------------------------------------------------------------------
 * ball_query: mmdet3d ball_query_cuda.cu, ball_query_kernel (0.18.1 form with
 * min_radius/max_radius). Reached through QueryAndGroup in every PointSAModule
 * (demf_votenet.py:52-53,58-62,155-162). Output rows start zeroed.
 * ------------------------------------------------------------------------- */
DEMF_API int demf_ref_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M,
                                 float min_radius, float max_radius, int nsample, int32_t* idx) {
  if (!xyz || !new_xyz || !idx) return -1;
  if (B < 0 || N <= 0 || M < 0 || nsample <= 0) return -2;
  const float min_r2 = min_radius * min_radius;
  const float max_r2 = max_radius * max_radius;
  const long total = (long)B * M;
#pragma omp parallel for schedule(static)
  for (long bm = 0; bm < total; ++bm) {
    const int b = (int)(bm / M);
    const float* p = xyz + (size_t)b * N * 3;
    const float* c = new_xyz + (size_t)bm * 3;
    int32_t* o = idx + (size_t)bm * nsample;
    for (int l = 0; l < nsample; ++l) o[l] = 0;
    const float nx = c[0], ny = c[1], nz = c[2];
    int cnt = 0;
    for (int k = 0; k < N; ++k) {
      const float d2 = sqdist(nx, ny, nz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
      if (d2 == 0 || (d2 >= min_r2 && d2 < max_r2)) {
        if (cnt == 0)
          for (int l = 0; l < nsample; ++l) o[l] = k;
        o[cnt] = k;
        ++cnt;
        if (cnt >= nsample) break;
      }
    }
  }
  return 0;
}

/* -------------------------------------------------------------------------
 * group_points: mmdet3d group_points_cuda.cu (forward gather, backward atomicAdd
 * scatter). out[b,c,m,s] = features[b,c,idx[b,m,s]].
 * ------------------------------------------------------------------------- */
DEMF_API int demf_ref_group_fwd(const float* features, const int32_t* idx, int B, int C, int N,
                                int M, int ns, float* out) {
  if (!features || !idx || !out) return -1;
  if (B < 0 || C < 0 || N <= 0 || M < 0 || ns < 0) return -2;
  const long rows = (long)B * C;
#pragma omp parallel for schedule(static)
  for (long bc = 0; bc < rows; ++bc) {
    const int b = (int)(bc / C);
    const float* f = features + (size_t)bc * N;
    const int32_t* id = idx + (size_t)b * M * ns;
    float* o = out + (size_t)bc * M * ns;
    for (long i = 0; i < (long)M * ns; ++i) o[i] = f[id[i]];
  }
  return 0;
}

DEMF_API int demf_ref_group_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N,
                                int M, int ns, float* grad_features) {
  if (!grad_out || !idx || !grad_features) return -1;
  if (B < 0 || C < 0 || N <= 0 || M < 0 || ns < 0) return -2;
  const long rows = (long)B * C;
#pragma omp parallel for schedule(static)
  for (long bc = 0; bc < rows; ++bc) {
    const int b = (int)(bc / C);
    float* g = grad_features + (size_t)bc * N;
    const int32_t* id = idx + (size_t)b * M * ns;
    const float* go = grad_out + (size_t)bc * M * ns;
    for (long i = 0; i < (long)M * ns; ++i) g[id[i]] += go[i];
  }
  return 0;
}

/* gather_points: mmdet3d gather_points_cuda.cu. out[b,c,m] = features[b,c,idx[b,m]]. */
DEMF_API int demf_ref_gather_fwd(const float* features, const int32_t* idx, int B, int C, int N,
                                 int M, float* out) {
  return demf_ref_group_fwd(features, idx, B, C, N, M, 1, out);
}
DEMF_API int demf_ref_gather_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N,
                                 int M, float* grad_features) {
  return demf_ref_group_bwd(grad_out, idx, B, C, N, M, 1, grad_features);
}

/* -------------------------------------------------------------------------
 * QueryAndGroup.forward (mmdet3d ops/group_points/group_points.py): ball query,
 * grouped_xyz - centre, optional / max_radius, cat([grouped_xyz, grouped_feat]).
 * The subtraction and the division are separate float32 torch ops upstream, so
 * they are separate roundings here.
 * ------------------------------------------------------------------------- */
DEMF_API int demf_ref_query_and_group_fwd(const float* xyz, const float* features,
                                          const float* new_xyz, int B, int N, int M, int C,
                                          float min_radius, float max_radius, int ns, int use_xyz,
                                          int normalize_xyz, int32_t* idx, float* out) {
  if (!xyz || !new_xyz || !idx || !out) return -1;
  if (C > 0 && !features) return -1;
  if (!use_xyz && C == 0) return -3;
  int rc = demf_ref_ball_query(xyz, new_xyz, B, N, M, min_radius, max_radius, ns, idx);
  if (rc) return rc;
  const int Cx = use_xyz ? 3 : 0;
  const int Co = Cx + C;
  const long total = (long)B * M;
  /* `grouped_xyz /= max_radius` runs on the GPU upstream, where ATen evaluates
   * tensor / python_float as tensor * (1.0f / float) (BinaryDivTrueKernel.cu). */
  const float inv_radius = 1.0f / max_radius;
#pragma omp parallel for schedule(static)
  for (long bm = 0; bm < total; ++bm) {
    const int b = (int)(bm / M), m = (int)(bm % M);
    const int32_t* id = idx + (size_t)bm * ns;
    const float* p = xyz + (size_t)b * N * 3;
    const float* c = new_xyz + (size_t)bm * 3;
    for (int s = 0; s < ns; ++s) {
      const int k = id[s];
      if (use_xyz) {
        for (int a = 0; a < 3; ++a) {
          float v = p[k * 3 + a] - c[a];
          if (normalize_xyz) v = v * inv_radius;
          out[(((size_t)b * Co + a) * M + m) * ns + s] = v;
        }
      }
      for (int ch = 0; ch < C; ++ch)
        out[(((size_t)b * Co + Cx + ch) * M + m) * ns + s] = features[((size_t)b * C + ch) * N + k];
    }
  }
  return 0;
}

/* -------------------------------------------------------------------------
 * three_nn: mmdet3d three_nn_cuda.cu, three_nn_kernel. best1..3 start at 1e40
 * held in double; candidates in index order; strict '<' cascade. Output is the
 * SQUARED distance (the upstream Python wrapper applies torch.sqrt afterwards).
 * ------------------------------------------------------------------------- */
DEMF_API int demf_ref_three_nn(const float* unknown, const float* known, int B, int n, int m,
                               float* dist2, int32_t* idx) {
  if (!unknown || !known || !dist2 || !idx) return -1;
  if (B < 0 || n < 0 || m <= 0) return -2;
  const long total = (long)B * n;
#pragma omp parallel for schedule(static)
  for (long bn = 0; bn < total; ++bn) {
    const int b = (int)(bn / n);
    const float* u = unknown + (size_t)bn * 3;
    const float* kn = known + (size_t)b * m * 3;
    const float ux = u[0], uy = u[1], uz = u[2];
    double best1 = 1e40, best2 = 1e40, best3 = 1e40;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < m; ++k) {
      const float d = sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
      if (d < best1) {
        best3 = best2; i3 = i2;
        best2 = best1; i2 = i1;
        best1 = d; i1 = k;
      } else if (d < best2) {
        best3 = best2; i3 = i2;
        best2 = d; i2 = k;
      } else if (d < best3) {
        best3 = d; i3 = k;
      }
    }
    dist2[bn * 3 + 0] = (float)best1;
    dist2[bn * 3 + 1] = (float)best2;
    dist2[bn * 3 + 2] = (float)best3;
    idx[bn * 3 + 0] = i1;
    idx[bn * 3 + 1] = i2;
    idx[bn * 3 + 2] = i3;
  }
  return 0;
}

/* three_interpolate: mmdet3d three_interpolate_cuda.cu (forward + atomicAdd backward). */
DEMF_API int demf_ref_three_interpolate_fwd(const float* features, const int32_t* idx,
                                            const float* weight, int B, int C, int m, int n,
                                            float* out) {
  if (!features || !idx || !weight || !out) return -1;
  if (B < 0 || C < 0 || m <= 0 || n < 0) return -2;
  const long rows = (long)B * C;
#pragma omp parallel for schedule(static)
  for (long bc = 0; bc < rows; ++bc) {
    const int b = (int)(bc / C);
    const float* f = features + (size_t)bc * m;
    const int32_t* id = idx + (size_t)b * n * 3;
    const float* w = weight + (size_t)b * n * 3;
    float* o = out + (size_t)bc * n;
    for (int i = 0; i < n; ++i)
      o[i] = fmaf(w[i * 3 + 2], f[id[i * 3 + 2]],
                  fmaf(w[i * 3 + 0], f[id[i * 3 + 0]], w[i * 3 + 1] * f[id[i * 3 + 1]]));
  }
  return 0;
}

DEMF_API int demf_ref_three_interpolate_bwd(const float* grad_out, const int32_t* idx,
                                            const float* weight, int B, int C, int n, int m,
                                            float* grad_features) {
  if (!grad_out || !idx || !weight || !grad_features) return -1;
  if (B < 0 || C < 0 || m <= 0 || n < 0) return -2;
  const long rows = (long)B * C;
#pragma omp parallel for schedule(static)
  for (long bc = 0; bc < rows; ++bc) {
    const int b = (int)(bc / C);
    float* g = grad_features + (size_t)bc * m;
    const int32_t* id = idx + (size_t)b * n * 3;
    const float* w = weight + (size_t)b * n * 3;
    const float* go = grad_out + (size_t)bc * n;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) g[id[i * 3 + k]] += go[i] * w[i * 3 + k];
  }
  return 0;
}

/* -------------------------------------------------------------------------
 * MSDA forward: mmcv ms_deform_attn_cuda_kernel.cuh, ms_deformable_im2col_gpu_kernel
 * + ms_deform_attn_im2col_bilinear. Reached from transformer.py:73-78 through
 * mmcv MultiScaleDeformableAttention (cfg demf_votenet.py:79-85).
 * align_corners=False pixel mapping (loc*size - 0.5), zero padding, sample
 * skipped unless -1 < h < H and -1 < w < W.
 * ------------------------------------------------------------------------- */
static inline float msda_bilinear(const float* v, int height, int width, int nheads, int channels,
                                  float h, float w, int m, int c) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low;
  const float hh = 1 - lh, hw = 1 - lw;
  const long w_stride = (long)nheads * channels;
  const long h_stride = (long)width * w_stride;
  const long base = (long)m * channels + c;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = v[h_low * h_stride + w_low * w_stride + base];
  if (h_low >= 0 && w_high <= width - 1) v2 = v[h_low * h_stride + w_high * w_stride + base];
  if (h_high <= height - 1 && w_low >= 0) v3 = v[h_high * h_stride + w_low * w_stride + base];
  if (h_high <= height - 1 && w_high <= width - 1)
    v4 = v[h_high * h_stride + w_high * w_stride + base];
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return fmaf(w4, v4, fmaf(w3, v3, fmaf(w1, v1, w2 * v2)));
}

DEMF_API int demf_ref_msda_fwd(const float* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const float* sampling_loc,
                               const float* attn_weight, int B, int S, int H, int D, int Q, int L,
                               int P, float* out) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !out)
    return -1;
  if (B < 0 || S <= 0 || H <= 0 || D <= 0 || Q < 0 || L <= 0 || P <= 0) return -2;
  const long total = (long)B * Q * H;
#pragma omp parallel for schedule(static)
  for (long bqh = 0; bqh < total; ++bqh) {
    const int m = (int)(bqh % H);
    const int b = (int)(bqh / ((long)Q * H));
    const float* loc = sampling_loc + (size_t)bqh * L * P * 2;
    const float* aw = attn_weight + (size_t)bqh * L * P;
    const size_t qid_stride = (size_t)H * D;
    for (int c = 0; c < D; ++c) {
      float col = 0;
      int wi = 0;
      for (int l = 0; l < L; ++l) {
        const int sh = (int)spatial_shapes[2 * l], sw = (int)spatial_shapes[2 * l + 1];
        const float* v = value + ((size_t)b * S + (size_t)level_start_index[l]) * qid_stride;
        for (int p = 0; p < P; ++p, ++wi) {
          const float loc_w = loc[2 * wi], loc_h = loc[2 * wi + 1];
          const float weight = aw[wi];
          const float h_im = loc_h * sh - 0.5f;
          const float w_im = loc_w * sw - 0.5f;
          if (h_im > -1 && w_im > -1 && h_im < sh && w_im < sw)
            col = fmaf(msda_bilinear(v, sh, sw, H, D, h_im, w_im, m, c), weight, col);
        }
      }
      out[(size_t)bqh * D + c] = col;
    }
  }
  return 0;
}

/* -------------------------------------------------------------------------
 * MSDA backward: mmcv ms_deformable_col2im_gpu_kernel_* + ms_deform_attn_col2im_
 * bilinear. grad_value accumulates (atomicAdd upstream; the accumulation order is
 * unspecified there, so comparisons on grad_value use a float tolerance);
 * grad_sampling_loc / grad_attn_weight are sums over the D channels of one head.
 * grad_value must be zero on entry. Single-threaded over b to keep the scatter
 * race-free; parallel over b.
 * ------------------------------------------------------------------------- */
DEMF_API int demf_ref_msda_bwd(const float* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const float* sampling_loc,
                               const float* attn_weight, const float* grad_out, int B, int S, int H,
                               int D, int Q, int L, int P, float* grad_value,
                               float* grad_sampling_loc, float* grad_attn_weight) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight ||
      !grad_out || !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return -1;
  if (B < 0 || S <= 0 || H <= 0 || D <= 0 || Q < 0 || L <= 0 || P <= 0) return -2;
  const size_t qid_stride = (size_t)H * D;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    for (long qh = 0; qh < (long)Q * H; ++qh) {
      const long bqh = (long)b * Q * H + qh;
      const int m = (int)(qh % H);
      const float* loc = sampling_loc + (size_t)bqh * L * P * 2;
      const float* aw = attn_weight + (size_t)bqh * L * P;
      float* gloc = grad_sampling_loc + (size_t)bqh * L * P * 2;
      float* gaw = grad_attn_weight + (size_t)bqh * L * P;
      int wi = 0;
      for (int l = 0; l < L; ++l) {
        const int sh = (int)spatial_shapes[2 * l], sw = (int)spatial_shapes[2 * l + 1];
        const size_t lvl_off = ((size_t)b * S + (size_t)level_start_index[l]) * qid_stride;
        const float* v = value + lvl_off;
        float* gv = grad_value + lvl_off;
        for (int p = 0; p < P; ++p, ++wi) {
          const float loc_w = loc[2 * wi], loc_h = loc[2 * wi + 1];
          const float weight = aw[wi];
          const float h = loc_h * sh - 0.5f;
          const float w = loc_w * sw - 0.5f;
          float g_w = 0, g_h = 0, g_a = 0;
          if (h > -1 && w > -1 && h < sh && w < sw) {
            const int h_low = (int)floorf(h), w_low = (int)floorf(w);
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h - h_low, lw = w - w_low;
            const float hh = 1 - lh, hw = 1 - lw;
            const long w_stride = (long)qid_stride;
            const long h_stride = (long)sw * w_stride;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            for (int c = 0; c < D; ++c) {
              const long base = (long)m * D + c;
              const float top_grad = grad_out[(size_t)bqh * D + c];
              const float top_grad_value = top_grad * weight;
              float grad_h_weight = 0, grad_w_weight = 0;
              float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
              if (h_low >= 0 && w_low >= 0) {
                const long o = h_low * h_stride + w_low * w_stride + base;
                v1 = v[o];
                grad_h_weight -= hw * v1;
                grad_w_weight -= hh * v1;
                gv[o] += w1 * top_grad_value;
              }
              if (h_low >= 0 && w_high <= sw - 1) {
                const long o = h_low * h_stride + w_high * w_stride + base;
                v2 = v[o];
                grad_h_weight -= lw * v2;
                grad_w_weight += hh * v2;
                gv[o] += w2 * top_grad_value;
              }
              if (h_high <= sh - 1 && w_low >= 0) {
                const long o = h_high * h_stride + w_low * w_stride + base;
                v3 = v[o];
                grad_h_weight += hw * v3;
                grad_w_weight -= lh * v3;
                gv[o] += w3 * top_grad_value;
              }
              if (h_high <= sh - 1 && w_high <= sw - 1) {
                const long o = h_high * h_stride + w_high * w_stride + base;
                v4 = v[o];
                grad_h_weight += lw * v4;
                grad_w_weight += lh * v4;
                gv[o] += w4 * top_grad_value;
              }
              const float val = fmaf(w4, v4, fmaf(w3, v3, fmaf(w1, v1, w2 * v2)));
              g_a += top_grad * val;
              g_w += sw * grad_w_weight * top_grad_value;
              g_h += sh * grad_h_weight * top_grad_value;
            }
          }
          gaw[wi] = g_a;
          gloc[2 * wi] = g_w;
          gloc[2 * wi + 1] = g_h;
        }
      }
    }
  }
  return 0;
}
