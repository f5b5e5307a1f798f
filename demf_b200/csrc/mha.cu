// Scaled-dot-product attention among the proposals (inference): softmax(Q K^T / sqrt(d)) V in exact fp32.
//
// Reference: the self-attention of DeMFTransformerDecoderLayer (demf/modeling/layers/transformer.py:55-80, built from
// configs/demf/demf_votenet.py:76-78 as mmcv `MultiheadAttention` = torch `nn.MultiheadAttention`, embed 288, 8 heads
// -> head width 36) over the 256 proposals of a scene. PyTorch runs it through its memory-efficient attention, an
// sm80 kernel (`fmha_cutlassF_f32_aligned_64x64_rf_sm80`, 32-54 us per batch of 8 here). The problem is tiny
// (256 x 256 x 36 per head), so it is done in fp32 FMAs -- more accurate than the TF32 tensor-core path and with no
// padding of the 36-wide heads:
//   CTA = 64 lane groups x QT queries of one (scene, head); K and V of that head are staged once in shared memory
//   (row-major, an ODD number of 16-byte units per row); FOUR lanes form a group, lane s taking keys j = s (mod 4),
//   so a warp reads four different K rows per instruction -- conflict-free 16-byte accesses, every row broadcast to
//   the eight groups of the warp -- and every value read feeds the QT queries of the group (the kernel is bound by
//   these loads: 4 QT FMAs per 16-byte load); online softmax per chunk of 8 keys; the four partial (max, sum,
//   accumulator) triples of a query are merged with two shuffle steps.
// Rows are the (L, B, E) tensors nn.MultiheadAttention works on, flattened: token l of scene b is row l*B + b (or
// b*L + l with `batch_first`, the layout of the decoder layer's rows path), head h occupies columns [h*D, h*D + D).
#include <cfloat>

#include "common.cuh"

namespace demf {
namespace {

constexpr int kMhaThreads = 256;
constexpr int kMhaGroups = kMhaThreads / 4;    // four lanes per query group

// D4 = head width / 4; QT = queries per lane group (register tile: every K / V value read from shared memory feeds
// QT queries -- the kernel is bound by shared-memory loads, one 16-byte load per 4 QT FMAs).
template <int D4, int QT>
__global__ void __launch_bounds__(kMhaThreads) mha_fwd_kernel(const float* __restrict__ q, long ldq,
                                                             const float* __restrict__ k, long ldk,
                                                             const float* __restrict__ v, long ldv,
                                                             float* __restrict__ o, long ldo, int Lq, int Lk, int B,
                                                             int H, float scale, int batch_first) {
  constexpr int D = 4 * D4;
  constexpr int RS = D4 | 1;             // row stride in float4: odd, so the four rows a warp reads hit distinct banks
  extern __shared__ __align__(16) float4 mha_smem[];
  float4* Ks = mha_smem;                 // [Lk][RS]
  float4* Vs = mha_smem + (size_t)Lk * RS;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int tid = threadIdx.x;
  for (int e = tid; e < Lk * D4; e += kMhaThreads) {
    const int j = e / D4, c = e - j * D4;
    const long row = batch_first ? (long)b * Lk + j : (long)j * B + b;
    Ks[j * RS + c] = __ldg(reinterpret_cast<const float4*>(k + row * ldk + h * D) + c);
    Vs[j * RS + c] = __ldg(reinterpret_cast<const float4*>(v + row * ldv + h * D) + c);
  }
  const int q0 = (blockIdx.y * kMhaGroups + (tid >> 2)) * QT;   // first query of this lane group
  const int sub = tid & 3;
  float4 qr[QT][D4];
#pragma unroll
  for (int t = 0; t < QT; ++t) {
    const int qi = min(q0 + t, Lq - 1);
    const long qrow = batch_first ? (long)b * Lq + qi : (long)qi * B + b;
    const float4* qp = reinterpret_cast<const float4*>(q + qrow * ldq + h * D);
#pragma unroll
    for (int c = 0; c < D4; ++c) {
      const float4 x = __ldg(qp + c);
      qr[t][c] = make_float4(x.x * scale, x.y * scale, x.z * scale, x.w * scale);   // as SDPA: q is scaled first
    }
  }
  __syncthreads();
  float m[QT], l[QT];
  float4 acc[QT][D4];
#pragma unroll
  for (int t = 0; t < QT; ++t) {
    m[t] = -FLT_MAX;
    l[t] = 0.f;
#pragma unroll
    for (int c = 0; c < D4; ++c) acc[t][c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j0 = 0; j0 < Lk; j0 += 8) {         // two keys per lane and chunk: j0 + sub, j0 + sub + 4
    const int ja = j0 + sub, jb = j0 + sub + 4;
    const float4* ka = Ks + (size_t)min(ja, Lk - 1) * RS;
    const float4* kb = Ks + (size_t)min(jb, Lk - 1) * RS;
    float sa[QT], sb[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) sa[t] = sb[t] = 0.f;
#pragma unroll
    for (int c = 0; c < D4; ++c) {
      const float4 xa = ka[c], xb = kb[c];
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        sa[t] = fmaf(qr[t][c].x, xa.x, sa[t]);
        sb[t] = fmaf(qr[t][c].x, xb.x, sb[t]);
        sa[t] = fmaf(qr[t][c].y, xa.y, sa[t]);
        sb[t] = fmaf(qr[t][c].y, xb.y, sb[t]);
        sa[t] = fmaf(qr[t][c].z, xa.z, sa[t]);
        sb[t] = fmaf(qr[t][c].z, xb.z, sb[t]);
        sa[t] = fmaf(qr[t][c].w, xa.w, sa[t]);
        sb[t] = fmaf(qr[t][c].w, xb.w, sb[t]);
      }
    }
    float pa[QT], pb[QT], corr[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      if (ja >= Lk) sa[t] = -FLT_MAX;
      if (jb >= Lk) sb[t] = -FLT_MAX;
      const float mx = fmaxf(m[t], fmaxf(sa[t], sb[t]));
      corr[t] = __expf(m[t] - mx);           // m = -FLT_MAX on the first chunk: exp(-huge) = 0
      pa[t] = ja < Lk ? __expf(sa[t] - mx) : 0.f;
      pb[t] = jb < Lk ? __expf(sb[t] - mx) : 0.f;
      l[t] = fmaf(l[t], corr[t], pa[t] + pb[t]);
      m[t] = mx;
    }
    const float4* va = Vs + (size_t)min(ja, Lk - 1) * RS;
    const float4* vb = Vs + (size_t)min(jb, Lk - 1) * RS;
#pragma unroll
    for (int c = 0; c < D4; ++c) {
      const float4 xa = va[c], xb = vb[c];
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        float4 a = acc[t][c];
        a.x = fmaf(pb[t], xb.x, fmaf(pa[t], xa.x, a.x * corr[t]));
        a.y = fmaf(pb[t], xb.y, fmaf(pa[t], xa.y, a.y * corr[t]));
        a.z = fmaf(pb[t], xb.z, fmaf(pa[t], xa.z, a.z * corr[t]));
        a.w = fmaf(pb[t], xb.w, fmaf(pa[t], xa.w, a.w * corr[t]));
        acc[t][c] = a;
      }
    }
  }
  // merge the four lanes of a query group: (m, l, acc) triples combine like two softmax chunks
#pragma unroll
  for (int t = 0; t < QT; ++t) {
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[t], off);
      const float lo = __shfl_xor_sync(0xffffffffu, l[t], off);
      const float mx = fmaxf(m[t], mo);
      const float ca = __expf(m[t] - mx), cb = __expf(mo - mx);
      l[t] = l[t] * ca + lo * cb;
#pragma unroll
      for (int c = 0; c < D4; ++c) {
        float4 a = acc[t][c];
        const float ox = __shfl_xor_sync(0xffffffffu, a.x, off), oy = __shfl_xor_sync(0xffffffffu, a.y, off);
        const float oz = __shfl_xor_sync(0xffffffffu, a.z, off), ow = __shfl_xor_sync(0xffffffffu, a.w, off);
        a.x = a.x * ca + ox * cb;
        a.y = a.y * ca + oy * cb;
        a.z = a.z * ca + oz * cb;
        a.w = a.w * ca + ow * cb;
        acc[t][c] = a;
      }
      m[t] = mx;
    }
    if (q0 + t < Lq && sub == t % 4) {       // the four lanes hold the same result: lane t writes query t
      const float inv = 1.0f / l[t];
      const long orow = batch_first ? (long)b * Lq + (q0 + t) : (long)(q0 + t) * B + b;
      float4* op = reinterpret_cast<float4*>(o + orow * ldo + h * D);
#pragma unroll
      for (int c = 0; c < D4; ++c)
        op[c] = make_float4(acc[t][c].x * inv, acc[t][c].y * inv, acc[t][c].z * inv, acc[t][c].w * inv);
    }
  }
}

template <int D4, int QT>
int launch_mha(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, float* o, long ldo,
               int Lq, int Lk, int B, int H, float scale, int batch_first, cudaStream_t st) {
  const size_t smem = (size_t)Lk * (D4 | 1) * 16 * 2;
  if (smem > 200 * 1024) {
    set_error("demf_mha_fwd: %d keys of width %d need %zu bytes of shared memory", Lk, 4 * D4, smem);
    return DEMF_E_UNSUPPORTED;
  }
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    const cudaError_t e = cudaFuncSetAttribute(mha_fwd_kernel<D4, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("demf_mha_fwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured = smem;
  }
  dim3 grid(B * H, (Lq + kMhaGroups * QT - 1) / (kMhaGroups * QT));
  mha_fwd_kernel<D4, QT><<<grid, kMhaThreads, smem, st>>>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, B, H, scale, batch_first);
  return after_launch("mha_fwd_kernel");
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_mha_supported(int D) { return (D == 32 || D == 36 || D == 64) ? 1 : 0; }

/* softmax(Q K^T * scale) V per (scene, head), inference. q (Lq*B, >= H*D) rows with stride ldq floats (token l of
 * scene b = row l*B + b, or b*L + l when batch_first; head h = columns [h*D, h*D+D)), k / v (Lk*B, ..) likewise,
 * o (Lq*B, ..) written. All row
 * starts and strides 16-byte aligned. Replaces torch's scaled_dot_product_attention under nn.MultiheadAttention in the
 * decoder layer's self-attention (demf/modeling/layers/transformer.py:55-80). */
int demf_mha_fwd(const float* q, long ldq, const float* k, long ldk, const float* v, long ldv, float* o, long ldo,
                 int Lq, int Lk, int B, int H, int D, float scale, int batch_first, void* stream) {
  DEMF_REQUIRE_PTR(q);
  DEMF_REQUIRE_PTR(k);
  DEMF_REQUIRE_PTR(v);
  DEMF_REQUIRE_PTR(o);
  DEMF_REQUIRE(Lq > 0 && Lk > 0 && B > 0 && H > 0 && (long)B * H <= 2147483647L, DEMF_E_SIZE);
  DEMF_REQUIRE(demf_mha_supported(D), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(ldq >= (long)H * D && ldk >= (long)H * D && ldv >= (long)H * D && ldo >= (long)H * D, DEMF_E_SIZE);
  auto al = [](const void* p, long ld) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ld % 4 == 0; };
  DEMF_REQUIRE(al(q, ldq) && al(k, ldk) && al(v, ldv) && al(o, ldo), DEMF_E_UNSUPPORTED);
  cudaStream_t st = as_stream(stream);
  switch (D) {
    case 32: return launch_mha<8, 2>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, B, H, scale, batch_first, st);
    case 36: return launch_mha<9, 2>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, B, H, scale, batch_first, st);
    default: return launch_mha<16, 1>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, B, H, scale, batch_first, st);
  }
}

}  // extern "C"
