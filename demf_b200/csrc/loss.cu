// The per-stage detection loss of DeMFVoteHead in two launches (one forward, one backward).
//
// Reference: demf/modeling/heads/class_agnostic_vote_head.py:622-712 (`_loss`) evaluates, per prediction stage,
// seven weighted sums over the B*Q proposals -- objectness CE (class weights 0.2/0.8), direction-class CE,
// direction-residual SmoothL1, size SmoothL1, centre SmoothL1, semantic CE and the axis-aligned IoU loss
// (configs/demf/demf_votenet.py:113-141; mmdet CrossEntropyLoss / SmoothL1Loss, mmdet3d AxisAlignedIoULoss,
// all reduction='sum' with per-proposal weights) -- as ~85 element-wise / reduction launches forward and ~100
// backward, each a few microseconds of a (B,Q,<=12) tensor: pure launch latency on the training step's critical
// path. Here one thread owns one proposal: the forward kernel evaluates all seven terms in registers and
// block-reduces them into seven accumulators; the backward kernel recomputes the same quantities and writes the
// gradient of the weighted sum with respect to every prediction tensor, scaled by the seven upstream gradients.
// Same formulas, same operation order per element as the torch modules (demf_b200/mm/losses.py); sums differ only
// by the order of the fp32 reduction.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kLossThreads = 128;
constexpr int kMaxClasses = 16;

struct StageLossArgs {
  // predictions, (rows, C) contiguous
  const float* center;        // 3
  const float* size;          // 3
  const float* dir_class;     // num_dir_bins
  const float* dir_res_norm;  // num_dir_bins
  const float* obj;           // 2
  const float* sem;           // num_sem
  // targets
  const long long* obj_t;
  const float* obj_w;
  const float* box_w;
  const float* size_t_;
  const float* center_t;
  const long long* dir_class_t;
  const float* dir_res_t;
  const long long* sem_t;
  long rows;
  int num_dir_bins, num_sem;
  float cw0, cw1;                       // objectness class weights
  float w_obj, w_dircls, w_dirres, w_size, w_center, w_sem, w_iou;   // loss_weight of each term
  float beta_dirres, beta_size, beta_center;
};

__device__ __forceinline__ float smooth_l1(float d, float beta) {
  const float a = fabsf(d);
  return a < beta ? 0.5f * a * a / beta : a - 0.5f * beta;
}
__device__ __forceinline__ float smooth_l1_grad(float d, float beta) {
  const float a = fabsf(d);
  return a < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}

// log-sum-exp of n logits (n <= kMaxClasses), torch's log_softmax formulation: max + log(sum exp(x - max))
__device__ __forceinline__ float lse(const float* x, int n) {
  float m = x[0];
  for (int i = 1; i < n; ++i) m = fmaxf(m, x[i]);
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += expf(x[i] - m);
  return m + logf(s);
}

struct Iou {
  float iou;
  float d_center[3], d_size[3];   // d iou / d center, d iou / d size
};

// axis_aligned_iou_aligned of (center -+ size/2) and (center_t -+ size_t/2), with its gradient w.r.t. the prediction
__device__ __forceinline__ Iou iou_and_grad(const float* c, const float* s, const float* ct, const float* st,
                                            bool want_grad) {
  float c1[3], c2[3], t1[3], t2[3], e1[3], wh[3], raw[3];
  float area1 = 1.f, area2 = 1.f, ov = 1.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float h = s[k] / 2.0f, ht = st[k] / 2.0f;
    c1[k] = c[k] - h;
    c2[k] = c[k] + h;
    t1[k] = ct[k] - ht;
    t2[k] = ct[k] + ht;
    e1[k] = c2[k] - c1[k];
    area1 *= e1[k];
    area2 *= t2[k] - t1[k];
    raw[k] = fminf(c2[k], t2[k]) - fmaxf(c1[k], t1[k]);
    wh[k] = fmaxf(raw[k], 0.f);
    ov *= wh[k];
  }
  const float un_raw = area1 + area2 - ov;
  const float un = fmaxf(un_raw, 1e-6f);
  Iou r;
  r.iou = ov / un;
#pragma unroll
  for (int k = 0; k < 3; ++k) r.d_center[k] = r.d_size[k] = 0.f;
  if (want_grad) {
    const float d_ov = 1.f / un, d_un = -ov / (un * un);
    const bool un_live = un_raw >= 1e-6f;     // clamp(min=eps) passes the gradient where the input is >= eps
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int a = (k + 1) % 3, b = (k + 2) % 3;
      // overlap = prod wh; union = area1 + area2 - overlap
      const float ov_wh = wh[a] * wh[b];
      const float g_wh = (d_ov - (un_live ? d_un : 0.f)) * ov_wh;     // through overlap (and -overlap in the union)
      const float g_raw = raw[k] >= 0.f ? g_wh : 0.f;
      // raw = min(c2, t2) - max(c1, t1)
      float g_c2 = c2[k] < t2[k] ? g_raw : (c2[k] == t2[k] ? 0.5f * g_raw : 0.f);
      float g_c1 = c1[k] > t1[k] ? -g_raw : (c1[k] == t1[k] ? -0.5f * g_raw : 0.f);
      // area1 = prod (c2 - c1)
      const float g_e1 = (un_live ? d_un : 0.f) * e1[a] * e1[b];
      g_c2 += g_e1;
      g_c1 -= g_e1;
      r.d_center[k] = g_c1 + g_c2;
      r.d_size[k] = (g_c2 - g_c1) * 0.5f;
    }
  }
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[7] += (objectness, dir_class, dir_res, size, center, semantic, iou)
__global__ void __launch_bounds__(kLossThreads) stage_loss_fwd_kernel(const StageLossArgs a, float* __restrict__ out) {
  const long r = (long)blockIdx.x * kLossThreads + threadIdx.x;
  float l[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (r < a.rows) {
    const float bw = a.box_w[r];
    {  // objectness: F.cross_entropy(weight=class_weight, reduction='none') * objectness_weights
      const float* o = a.obj + r * 2;
      const long long t = a.obj_t[r];
      const float cw = t ? a.cw1 : a.cw0;
      l[0] = cw * (lse(o, 2) - o[t]) * a.obj_w[r];
    }
    {
      const float* x = a.dir_class + r * a.num_dir_bins;
      const long long t = a.dir_class_t[r];
      l[1] = (lse(x, a.num_dir_bins) - x[t]) * bw;
      const float d = a.dir_res_norm[r * a.num_dir_bins + t] - a.dir_res_t[r];
      l[2] = smooth_l1(d, a.beta_dirres) * bw;
    }
    float c[3], s[3], ct[3], st[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c[k] = a.center[r * 3 + k];
      s[k] = a.size[r * 3 + k];
      ct[k] = a.center_t[r * 3 + k];
      st[k] = a.size_t_[r * 3 + k];
      l[3] += smooth_l1(s[k] - st[k], a.beta_size) * bw;
      l[4] += smooth_l1(c[k] - ct[k], a.beta_center) * bw;
    }
    if (a.sem != nullptr) {
      const float* x = a.sem + r * a.num_sem;
      l[5] = (lse(x, a.num_sem) - x[a.sem_t[r]]) * bw;
    }
    if (a.w_iou != 0.f) l[6] = (1.f - iou_and_grad(c, s, ct, st, false).iou) * bw;
  }
  __shared__ float part[7][kLossThreads / 32];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float v = warp_sum(l[i]);
    if ((threadIdx.x & 31) == 0) part[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float v = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) v += part[threadIdx.x][w];
    const float lw[7] = {a.w_obj, a.w_dircls, a.w_dirres, a.w_size, a.w_center, a.w_sem, a.w_iou};
    atomicAdd(out + threadIdx.x, v * lw[threadIdx.x]);
  }
}

// gradients of sum_i up[i] * loss_i w.r.t. every prediction tensor
__global__ void __launch_bounds__(kLossThreads) stage_loss_bwd_kernel(const StageLossArgs a, const float* __restrict__ up,
                                                                      float* __restrict__ g_center,
                                                                      float* __restrict__ g_size,
                                                                      float* __restrict__ g_dir_class,
                                                                      float* __restrict__ g_dir_res_norm,
                                                                      float* __restrict__ g_obj,
                                                                      float* __restrict__ g_sem) {
  const long r = (long)blockIdx.x * kLossThreads + threadIdx.x;
  if (r >= a.rows) return;
  const float bw = a.box_w[r];
  const float u_obj = up[0] * a.w_obj, u_dc = up[1] * a.w_dircls, u_dr = up[2] * a.w_dirres, u_sz = up[3] * a.w_size,
              u_ct = up[4] * a.w_center, u_sem = up[5] * a.w_sem, u_iou = up[6] * a.w_iou;
  {
    const float* o = a.obj + r * 2;
    const long long t = a.obj_t[r];
    const float k = (t ? a.cw1 : a.cw0) * a.obj_w[r] * u_obj;
    const float z = lse(o, 2);
    g_obj[r * 2 + 0] = k * (expf(o[0] - z) - (t == 0 ? 1.f : 0.f));
    g_obj[r * 2 + 1] = k * (expf(o[1] - z) - (t == 1 ? 1.f : 0.f));
  }
  {
    const int n = a.num_dir_bins;
    const float* x = a.dir_class + r * n;
    const long long t = a.dir_class_t[r];
    const float z = lse(x, n);
    const float d = a.dir_res_norm[r * n + t] - a.dir_res_t[r];
    const float gd = smooth_l1_grad(d, a.beta_dirres) * bw * u_dr;
    for (int i = 0; i < n; ++i) {
      g_dir_class[r * n + i] = bw * u_dc * (expf(x[i] - z) - (i == t ? 1.f : 0.f));
      g_dir_res_norm[r * n + i] = i == t ? gd : 0.f;
    }
  }
  float c[3], s[3], ct[3], st[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    c[k] = a.center[r * 3 + k];
    s[k] = a.size[r * 3 + k];
    ct[k] = a.center_t[r * 3 + k];
    st[k] = a.size_t_[r * 3 + k];
  }
  Iou io;
  if (a.w_iou != 0.f) io = iou_and_grad(c, s, ct, st, true);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float gc = smooth_l1_grad(c[k] - ct[k], a.beta_center) * bw * u_ct;
    float gs = smooth_l1_grad(s[k] - st[k], a.beta_size) * bw * u_sz;
    if (a.w_iou != 0.f) {   // loss = (1 - iou) * bw
      gc -= io.d_center[k] * bw * u_iou;
      gs -= io.d_size[k] * bw * u_iou;
    }
    g_center[r * 3 + k] = gc;
    g_size[r * 3 + k] = gs;
  }
  if (a.sem != nullptr) {
    const int n = a.num_sem;
    const float* x = a.sem + r * n;
    const long long t = a.sem_t[r];
    const float z = lse(x, n);
    for (int i = 0; i < n; ++i) g_sem[r * n + i] = bw * u_sem * (expf(x[i] - z) - (i == t ? 1.f : 0.f));
  }
}

int fill_args(StageLossArgs* a, const float* center, const float* size, const float* dir_class,
              const float* dir_res_norm, const float* obj, const float* sem, const int64_t* obj_t, const float* obj_w,
              const float* box_w, const float* size_t_, const float* center_t, const int64_t* dir_class_t,
              const float* dir_res_t, const int64_t* sem_t, long rows, int num_dir_bins, int num_sem,
              const float* cfg) {
  a->center = center;
  a->size = size;
  a->dir_class = dir_class;
  a->dir_res_norm = dir_res_norm;
  a->obj = obj;
  a->sem = sem;
  a->obj_t = reinterpret_cast<const long long*>(obj_t);
  a->obj_w = obj_w;
  a->box_w = box_w;
  a->size_t_ = size_t_;
  a->center_t = center_t;
  a->dir_class_t = reinterpret_cast<const long long*>(dir_class_t);
  a->dir_res_t = dir_res_t;
  a->sem_t = reinterpret_cast<const long long*>(sem_t);
  a->rows = rows;
  a->num_dir_bins = num_dir_bins;
  a->num_sem = num_sem;
  a->cw0 = cfg[0];
  a->cw1 = cfg[1];
  a->w_obj = cfg[2];
  a->w_dircls = cfg[3];
  a->w_dirres = cfg[4];
  a->w_size = cfg[5];
  a->w_center = cfg[6];
  a->w_sem = cfg[7];
  a->w_iou = cfg[8];
  a->beta_dirres = cfg[9];
  a->beta_size = cfg[10];
  a->beta_center = cfg[11];
  return 0;
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

/* cfg (host, 12 floats): objectness class weights (2); loss weights objectness, dir_class, dir_res, size, center,
 * semantic, iou (7); SmoothL1 beta of dir_res, size, center (3). losses: 7 device floats, ZERO on entry. */
int demf_stage_loss_fwd(const float* center, const float* size, const float* dir_class, const float* dir_res_norm,
                        const float* obj, const float* sem, const int64_t* obj_t, const float* obj_w,
                        const float* box_w, const float* size_t_, const float* center_t, const int64_t* dir_class_t,
                        const float* dir_res_t, const int64_t* sem_t, long rows, int num_dir_bins, int num_sem,
                        const float* cfg, float* losses, void* stream) {
  DEMF_REQUIRE_PTR(center);
  DEMF_REQUIRE_PTR(size);
  DEMF_REQUIRE_PTR(dir_class);
  DEMF_REQUIRE_PTR(dir_res_norm);
  DEMF_REQUIRE_PTR(obj);
  DEMF_REQUIRE_PTR(obj_t);
  DEMF_REQUIRE_PTR(obj_w);
  DEMF_REQUIRE_PTR(box_w);
  DEMF_REQUIRE_PTR(size_t_);
  DEMF_REQUIRE_PTR(center_t);
  DEMF_REQUIRE_PTR(dir_class_t);
  DEMF_REQUIRE_PTR(dir_res_t);
  DEMF_REQUIRE_PTR(cfg);
  DEMF_REQUIRE_PTR(losses);
  DEMF_REQUIRE(rows > 0 && num_dir_bins > 0 && num_dir_bins <= kMaxClasses && num_sem >= 0 && num_sem <= kMaxClasses,
               DEMF_E_SIZE);
  DEMF_REQUIRE((sem == nullptr) == (sem_t == nullptr), DEMF_E_SIZE);
  StageLossArgs a;
  fill_args(&a, center, size, dir_class, dir_res_norm, obj, sem, obj_t, obj_w, box_w, size_t_, center_t, dir_class_t,
            dir_res_t, sem_t, rows, num_dir_bins, num_sem, cfg);
  const unsigned blocks = (unsigned)((rows + kLossThreads - 1) / kLossThreads);
  stage_loss_fwd_kernel<<<blocks, kLossThreads, 0, as_stream(stream)>>>(a, losses);
  return after_launch("stage_loss_fwd_kernel");
}

int demf_stage_loss_bwd(const float* center, const float* size, const float* dir_class, const float* dir_res_norm,
                        const float* obj, const float* sem, const int64_t* obj_t, const float* obj_w,
                        const float* box_w, const float* size_t_, const float* center_t, const int64_t* dir_class_t,
                        const float* dir_res_t, const int64_t* sem_t, long rows, int num_dir_bins, int num_sem,
                        const float* cfg, const float* upstream, float* g_center, float* g_size, float* g_dir_class,
                        float* g_dir_res_norm, float* g_obj, float* g_sem, void* stream) {
  DEMF_REQUIRE_PTR(center);
  DEMF_REQUIRE_PTR(cfg);
  DEMF_REQUIRE_PTR(upstream);
  DEMF_REQUIRE_PTR(g_center);
  DEMF_REQUIRE_PTR(g_size);
  DEMF_REQUIRE_PTR(g_dir_class);
  DEMF_REQUIRE_PTR(g_dir_res_norm);
  DEMF_REQUIRE_PTR(g_obj);
  DEMF_REQUIRE(rows > 0 && num_dir_bins > 0 && num_dir_bins <= kMaxClasses && num_sem >= 0 && num_sem <= kMaxClasses,
               DEMF_E_SIZE);
  DEMF_REQUIRE((sem == nullptr) == (g_sem == nullptr), DEMF_E_SIZE);
  StageLossArgs a;
  fill_args(&a, center, size, dir_class, dir_res_norm, obj, sem, obj_t, obj_w, box_w, size_t_, center_t, dir_class_t,
            dir_res_t, sem_t, rows, num_dir_bins, num_sem, cfg);
  const unsigned blocks = (unsigned)((rows + kLossThreads - 1) / kLossThreads);
  stage_loss_bwd_kernel<<<blocks, kLossThreads, 0, as_stream(stream)>>>(a, upstream, g_center, g_size, g_dir_class,
                                                                      g_dir_res_norm, g_obj, g_sem);
  return after_launch("stage_loss_bwd_kernel");
}

}  // extern "C"
