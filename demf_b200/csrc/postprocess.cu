// Inference post-processing of the vote head, the two loops that upstream runs per scene on the host
// side of `multiclass_nms_single` (mmdet3d 0.18.1 models/dense_heads/vote_head.py, reached from
// demf/modeling/heads/class_agnostic_vote_head.py:739-743):
//
//   box_point_count_kernel : how many points of the scene lie inside each decoded box -- the column sums
//                            of `bbox.points_in_boxes(points)` (a (N,K) int tensor upstream, 10 M entries
//                            per scene, of which only `sum > 5` is used). Same local-frame test as the
//                            roiaware_pool3d kernel (z within the box, |x_local| < dx/2, |y_local| < dy/2).
//   aligned_nms_kernel     : `aligned_3d_nms` -- greedy class-aware NMS on axis-aligned (min,max) boxes.
//                            Upstream is a Python while-loop with ~20 tiny launches and a host sync per
//                            kept box; here one CTA per scene sorts the scores in shared memory and runs
//                            the same greedy sweep with the block's threads testing all remaining boxes.
//
// Both are latency-, not bandwidth-bound (a scene is 240 KB of points and 512 boxes); they exist to remove
// ~10^4 launches and ~500 host syncs per batch from the evaluation loop.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kBoxesPerWarp = 4;

// boxes (B,K,7) = (x, y, z_bottom, dx, dy, dz, yaw); points rows of `stride` floats (xyz first).
// gravity != 0: boxes carry the GRAVITY centre (origin (0.5,0.5,0.5), what the coder decodes); the bottom
// centre is formed first exactly as the host conversion would (z - dz*0.5), so counts do not depend on
// which form the caller hands over.
__global__ void __launch_bounds__(256) box_point_count_kernel(const float* __restrict__ points, int stride,
                                                              const float* __restrict__ boxes, int N, int K,
                                                              int gravity, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int k0 = (blockIdx.x * (blockDim.x >> 5) + warp) * kBoxesPerWarp;
  if (k0 >= K) return;
  float cx[kBoxesPerWarp], cy[kBoxesPerWarp], cz[kBoxesPerWarp], hx[kBoxesPerWarp], hy[kBoxesPerWarp],
      hz[kBoxesPerWarp], ca[kBoxesPerWarp], sa[kBoxesPerWarp];
  int cnt[kBoxesPerWarp];
#pragma unroll
  for (int i = 0; i < kBoxesPerWarp; ++i) {
    const int k = min(k0 + i, K - 1);
    const float* bx = boxes + ((long)b * K + k) * 7;
    const float dz = __ldg(bx + 5);
    cx[i] = __ldg(bx);
    cy[i] = __ldg(bx + 1);
    const float zb = gravity ? __fsub_rn(__ldg(bx + 2), __fmul_rn(dz, 0.5f)) : __ldg(bx + 2);
    cz[i] = __fadd_rn(zb, __fmul_rn(dz, 0.5f));
    hx[i] = __fmul_rn(__ldg(bx + 3), 0.5f);
    hy[i] = __fmul_rn(__ldg(bx + 4), 0.5f);
    hz[i] = __fmul_rn(dz, 0.5f);
    ca[i] = cosf(__ldg(bx + 6));
    sa[i] = sinf(__ldg(bx + 6));
    cnt[i] = 0;
  }
  const float* p = points + (long)b * N * stride;
  for (int n = lane; n < N; n += 32) {
    const float x = __ldg(p + (long)n * stride), y = __ldg(p + (long)n * stride + 1),
                z = __ldg(p + (long)n * stride + 2);
#pragma unroll
    for (int i = 0; i < kBoxesPerWarp; ++i) {
      const float dx = __fsub_rn(x, cx[i]), dy = __fsub_rn(y, cy[i]), dz = __fsub_rn(z, cz[i]);
      // world -> box frame, products rounded separately like the torch expression of the oracle
      const float lx = __fsub_rn(__fmul_rn(dx, ca[i]), __fmul_rn(dy, sa[i]));
      const float ly = __fadd_rn(__fmul_rn(dx, sa[i]), __fmul_rn(dy, ca[i]));
      cnt[i] += (fabsf(dz) <= hz[i]) && (fabsf(lx) < hx[i]) && (fabsf(ly) < hy[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kBoxesPerWarp; ++i) {
    int c = cnt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && k0 + i < K) counts[(long)b * K + k0 + i] = c;
  }
}

// ---------------------------------------------------------------------------------- NMS --
constexpr int kNmsThreads = 1024;
constexpr int kNmsMaxBoxes = 4096;

// order: score descending, ties -> larger index first (= taking the last element of a stable
// ascending argsort, which is what the upstream loop does)
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia > ib);
}

// One CTA per scene. minmax (B,K,6) = (x1,y1,z1,x2,y2,z2); valid (B,K) u8 selects the boxes that take
// part; keep (B,K) u8 receives 1 for the boxes the greedy sweep picks.
__global__ void __launch_bounds__(kNmsThreads) aligned_nms_kernel(const float* __restrict__ minmax,
                                                                  const float* __restrict__ scores,
                                                                  const int64_t* __restrict__ classes,
                                                                  const uint8_t* __restrict__ valid, int K,
                                                                  int Kpad, float thresh,
                                                                  uint8_t* __restrict__ keep) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_score = reinterpret_cast<float*>(smem_raw);          // Kpad
  int* s_idx = reinterpret_cast<int*>(s_score + Kpad);          // Kpad
  float* s_box = reinterpret_cast<float*>(s_idx + Kpad);        // Kpad * 6 (sorted order)
  float* s_area = s_box + (size_t)Kpad * 6;                     // Kpad
  int* s_cls = reinterpret_cast<int*>(s_area + Kpad);           // Kpad
  uint8_t* s_dead = reinterpret_cast<uint8_t*>(s_cls + Kpad);   // Kpad
  __shared__ int s_n;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  minmax += (long)b * K * 6;
  scores += (long)b * K;
  classes += (long)b * K;
  valid += (long)b * K;
  keep += (long)b * K;

  for (int i = tid; i < Kpad; i += kNmsThreads) {
    const bool v = i < K && valid[i];
    s_score[i] = v ? scores[i] : -INFINITY;
    s_idx[i] = v ? i : -1;   // padding / invalid entries sort to the end (index -1 loses every tie)
    if (i < K) keep[i] = 0;
  }
  if (tid == 0) s_n = 0;
  __syncthreads();
  // bitonic sort, `before` order
  for (int size = 2; size <= Kpad; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      for (int i = tid; i < Kpad; i += kNmsThreads) {
        const int j = i ^ strd;
        if (j > i) {
          const bool up = (i & size) == 0;
          const float si = s_score[i], sj = s_score[j];
          const int ii = s_idx[i], ij = s_idx[j];
          const bool swap = up ? before(sj, ij, si, ii) : before(si, ii, sj, ij);
          if (swap) {
            s_score[i] = sj; s_score[j] = si;
            s_idx[i] = ij; s_idx[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < Kpad; i += kNmsThreads) {
    const int k = s_idx[i];
    s_dead[i] = k < 0;
    if (k >= 0) {
      atomicAdd(&s_n, 1);
      float v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = s_box[(size_t)i * 6 + c] = minmax[(long)k * 6 + c];
      s_area[i] = __fmul_rn(__fmul_rn(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1])), __fsub_rn(v[5], v[2]));
      s_cls[i] = (int)classes[k];
    }
  }
  __syncthreads();
  const int n = s_n;  // valid boxes occupy sorted positions [0, n)
  for (int i = 0; i < n; ++i) {
    if (s_dead[i]) continue;  // uniform: read after the barrier of the previous iteration
    if (tid == 0) keep[s_idx[i]] = 1;
    const float x1 = s_box[i * 6], y1 = s_box[i * 6 + 1], z1 = s_box[i * 6 + 2], x2 = s_box[i * 6 + 3],
                y2 = s_box[i * 6 + 4], z2 = s_box[i * 6 + 5];
    const float ai = s_area[i];
    const int ci = s_cls[i];
    for (int j = i + 1 + tid; j < n; j += kNmsThreads) {
      if (s_dead[j]) continue;
      const float il = fmaxf(0.f, __fsub_rn(fminf(x2, s_box[j * 6 + 3]), fmaxf(x1, s_box[j * 6])));
      const float iw = fmaxf(0.f, __fsub_rn(fminf(y2, s_box[j * 6 + 4]), fmaxf(y1, s_box[j * 6 + 1])));
      const float ih = fmaxf(0.f, __fsub_rn(fminf(z2, s_box[j * 6 + 5]), fmaxf(z1, s_box[j * 6 + 2])));
      const float inter = __fmul_rn(__fmul_rn(il, iw), ih);
      float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, s_area[j]), inter));
      iou = __fmul_rn(iou, s_cls[j] == ci ? 1.f : 0.f);
      if (!(iou <= thresh)) s_dead[j] = 1;  // NaN (two empty boxes) is dropped, as `iou <= thresh` is upstream
    }
    __syncthreads();
  }
}

// Everything multiclass_nms_single derives per box before the NMS, one thread per box: the axis-aligned hull
// of the rotated box (DepthInstance3DBoxes.corners -> min / max, same fp32 operations in the same order as
// geometry.box_corner_minmax), the semantic class (first maximum), and whether it holds more than
// `min_points` points. boxes: gravity-centre (B*K,7).
__global__ void __launch_bounds__(256) nms_prepare_kernel(const float* __restrict__ boxes,
                                                          const float* __restrict__ sem, int C,
                                                          const int* __restrict__ counts, int min_points, long total,
                                                          float* __restrict__ minmax, int64_t* __restrict__ classes,
                                                          uint8_t* __restrict__ valid) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const float* bx = boxes + e * 7;
    const float x = __ldg(bx), y = __ldg(bx + 1), dx = __ldg(bx + 3), dy = __ldg(bx + 4), dz = __ldg(bx + 5);
    const float zb = __fsub_rn(__ldg(bx + 2), __fmul_rn(dz, 0.5f));
    const float ca = cosf(__ldg(bx + 6)), sa = sinf(__ldg(bx + 6));
    const float fx[4] = {-0.5f, -0.5f, 0.5f, 0.5f}, fy[4] = {-0.5f, 0.5f, 0.5f, -0.5f};
    float x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sx = __fmul_rn(fx[k], dx), sy = __fmul_rn(fy[k], dy);
      const float rx = __fadd_rn(__fadd_rn(__fmul_rn(sx, ca), __fmul_rn(sy, sa)), x);
      const float ry = __fadd_rn(__fadd_rn(__fmul_rn(-sx, sa), __fmul_rn(sy, ca)), y);
      x0 = fminf(x0, rx); x1 = fmaxf(x1, rx);
      y0 = fminf(y0, ry); y1 = fmaxf(y1, ry);
    }
    const float z0 = __fadd_rn(zb, __fmul_rn(dz, 0.0f)), z1 = __fadd_rn(zb, __fmul_rn(dz, 1.0f));
    float* o = minmax + e * 6;
    o[0] = x0; o[1] = y0; o[2] = fminf(z0, z1);
    o[3] = x1; o[4] = y1; o[5] = fmaxf(z0, z1);
    const float* sp = sem + e * C;
    int best = 0;
    float bv = __ldg(sp);
    for (int k = 1; k < C; ++k) {
      const float v = __ldg(sp + k);
      if (v > bv) { bv = v; best = k; }
    }
    classes[e] = best;
    valid[e] = counts[e] > min_points;
  }
}

// keep &= score > score_thr; per-scene number of selected boxes (for the single host read)
__global__ void __launch_bounds__(256) nms_finish_kernel(const float* __restrict__ scores, float score_thr, int K,
                                                         uint8_t* __restrict__ keep, int* __restrict__ nsel) {
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const long e = (long)b * K + i;
    const uint8_t k = keep[e] && scores[e] > score_thr;
    keep[e] = k;
    c += k;
  }
  if (c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) nsel[b] = s_cnt;
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_box_point_count(const float* points, int point_stride, const float* boxes, int B, int N, int K,
                         int gravity_centre, int32_t* counts, void* stream) {
  DEMF_REQUIRE_PTR(points);
  DEMF_REQUIRE_PTR(boxes);
  DEMF_REQUIRE_PTR(counts);
  DEMF_REQUIRE(B >= 0 && N >= 0 && K >= 0 && point_stride >= 3, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  if (B == 0 || K == 0) return 0;
  const int boxes_per_block = 8 * kBoxesPerWarp;
  dim3 grid((K + boxes_per_block - 1) / boxes_per_block, B);
  box_point_count_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, point_stride, boxes, N, K, gravity_centre,
                                                             counts);
  return after_launch("box_point_count_kernel");
}

int demf_aligned_3d_nms(const float* minmax, const float* scores, const int64_t* classes, const uint8_t* valid,
                        int B, int K, float thresh, uint8_t* keep, void* stream) {
  DEMF_REQUIRE_PTR(minmax);
  DEMF_REQUIRE_PTR(scores);
  DEMF_REQUIRE_PTR(classes);
  DEMF_REQUIRE_PTR(valid);
  DEMF_REQUIRE_PTR(keep);
  DEMF_REQUIRE(B >= 0 && K >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(K <= kNmsMaxBoxes, DEMF_E_UNSUPPORTED);
  if (B == 0 || K == 0) return 0;
  int Kpad = 2;
  while (Kpad < K) Kpad <<= 1;
  const size_t smem = (size_t)Kpad * (4 + 4 + 24 + 4 + 4 + 1);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(aligned_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kNmsMaxBoxes * (4 + 4 + 24 + 4 + 4 + 1));
    configured = true;
  }
  aligned_nms_kernel<<<B, kNmsThreads, smem, as_stream(stream)>>>(minmax, scores, classes, valid, K, Kpad,
                                                                thresh, keep);
  return after_launch("aligned_nms_kernel");
}

int demf_nms_select(const float* boxes, const float* obj_scores, const float* sem_scores, const int32_t* counts,
                    int B, int K, int C, int min_points, float nms_thresh, float score_thresh, float* minmax,
                    int64_t* classes, uint8_t* valid, uint8_t* selected, int32_t* num_selected, void* stream) {
  DEMF_REQUIRE_PTR(boxes);
  DEMF_REQUIRE_PTR(obj_scores);
  DEMF_REQUIRE_PTR(sem_scores);
  DEMF_REQUIRE_PTR(counts);
  DEMF_REQUIRE_PTR(minmax);
  DEMF_REQUIRE_PTR(classes);
  DEMF_REQUIRE_PTR(valid);
  DEMF_REQUIRE_PTR(selected);
  DEMF_REQUIRE_PTR(num_selected);
  DEMF_REQUIRE(B >= 0 && K >= 0 && C >= 1, DEMF_E_SIZE);
  DEMF_REQUIRE(K <= kNmsMaxBoxes, DEMF_E_UNSUPPORTED);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const long total = (long)B * K;
  if (total > 0) {
    long blocks = (total + 255) / 256;
    if (blocks > (long)kNumSMs * 8) blocks = (long)kNumSMs * 8;
    nms_prepare_kernel<<<(unsigned)blocks, 256, 0, st>>>(boxes, sem_scores, C, counts, min_points, total, minmax,
                                                        classes, valid);
    if (int rc = after_launch("nms_prepare_kernel")) return rc;
    if (int rc = demf_aligned_3d_nms(minmax, obj_scores, classes, valid, B, K, nms_thresh, selected, stream)) return rc;
  }
  nms_finish_kernel<<<B, 256, 0, st>>>(obj_scores, score_thresh, K, selected, num_selected);
  return after_launch("nms_finish_kernel");
}

}  // extern "C"
