// Multi-scale deformable attention (forward + backward) for sm_100a.
//
// Data layout in HBM (upstream's, kept): value (B,S,H,D) f32 -- one (b,s,h) row is D floats,
// 128 bytes at D=32, i.e. exactly one cache line per bilinear corner.
//
// Upstream runs one thread per output scalar: the 32 threads of a warp (= the 32 channels of
// one head) each re-read the same location/weight scalars and each recompute the same
// bilinear set-up. Here the work is split in two phases inside one launch:
//   phase 1  one thread per SAMPLE (b,q,h,l,p): coalesced float2/float loads of the location
//            and weight, one bilinear set-up, result = a 36-byte record in shared memory
//            {4 element offsets, 4 corner weights (0 for padded corners), attention weight};
//   phase 2  D/4 lanes per (b,q,h) tuple, each lane owning 4 channels: per sample one
//            broadcast read of the record and four 16-byte loads (LDG.128) -- the D/4 lanes of
//            a tuple cover one full corner row, so every corner is one coalesced line request
//            -- with the sample loop unrolled so 16 line requests per lane are in flight.
// The accumulation order per channel is upstream's
//   col = fma(fma(w4,v4,fma(w3,v3,fma(w1,v1,w2*v2))), attn, col)
// so results agree with the oracle to rounding of zero-weight terms only.
//
// Roofline: HBM/L2 bandwidth. Algorithmic bytes per launch:
//   B*Q*H*L*P*(4*D*4 + 12) + B*Q*H*D*4.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kWarps = 4;  // 128-thread CTAs: at Q=256, B=8 the grid is 1024 CTAs = 6.9 per SM (8 warps: 3.46 -> a 4-vs-3 imbalance)
constexpr int kThreads = kWarps * 32;
constexpr int kMaxLevels = 16;

struct LevelInfo {
  int h[kMaxLevels];
  int w[kMaxLevels];
  int start[kMaxLevels];
};

// Reads the (L,2) int64 shapes and (L) int64 level starts (device memory) into shared memory.
__device__ __forceinline__ void load_levels(LevelInfo* s, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ lsi, int L) {
  if (threadIdx.x < L) {
    s->h[threadIdx.x] = (int)__ldg(shapes + 2 * threadIdx.x);
    s->w[threadIdx.x] = (int)__ldg(shapes + 2 * threadIdx.x + 1);
    s->start[threadIdx.x] = (int)__ldg(lsi + threadIdx.x);
  }
}

struct Bilinear {
  int off[4];   // element offsets of the 4 corner rows relative to value[b] (head offset included)
  float w[4];   // hh*hw, hh*lw, lh*hw, lh*lw ; 0 where the corner is padding
  float lh, lw; // fractional parts (backward)
  unsigned valid;  // bit k: corner k inside the map; bit 4: sample inside (-1,H)x(-1,W)
};

// The bilinear set-up of mmcv ms_deform_attn_im2col_bilinear for one sample.
__device__ __forceinline__ Bilinear setup_sample(float loc_w, float loc_h, int sh, int sw,
                                                 int level_start, int H, int D, int head) {
  Bilinear r;
  // upstream: loc*size - 0.5 with the product rounded to float first (the 0.5 literal is a
  // double there, which forbids an FMA); __fmul_rn keeps nvcc from contracting it here.
  const float h_im = __fmul_rn(loc_h, (float)sh) - 0.5f;
  const float w_im = __fmul_rn(loc_w, (float)sw) - 0.5f;
  r.valid = 0;
  r.lh = r.lw = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.off[k] = (level_start * H + head) * D;  // always a legal address
    r.w[k] = 0.f;
  }
  if (h_im > -1.f && w_im > -1.f && h_im < (float)sh && w_im < (float)sw) {
    const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
    const float hh = 1.f - lh, hw = 1.f - lw;
    r.lh = lh;
    r.lw = lw;
    const bool t = h_low >= 0, bt = h_high <= sh - 1, l = w_low >= 0, rt = w_high <= sw - 1;
    r.valid = 16u | (t && l ? 1u : 0u) | (t && rt ? 2u : 0u) | (bt && l ? 4u : 0u) |
              (bt && rt ? 8u : 0u);
    if (t && l) { r.off[0] = ((level_start + h_low * sw + w_low) * H + head) * D; r.w[0] = hh * hw; }
    if (t && rt) { r.off[1] = ((level_start + h_low * sw + w_high) * H + head) * D; r.w[1] = hh * lw; }
    if (bt && l) { r.off[2] = ((level_start + h_high * sw + w_low) * H + head) * D; r.w[2] = lh * hw; }
    if (bt && rt) { r.off[3] = ((level_start + h_high * sw + w_high) * H + head) * D; r.w[3] = lh * lw; }
  }
  return r;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float bil(float w1, float v1, float w2, float v2, float w3, float v3,
                                     float w4, float v4) {
  return __fmaf_rn(w4, v4, __fmaf_rn(w3, v3, __fmaf_rn(w1, v1, __fmul_rn(w2, v2))));
}

// ------------------------------------------------------------------ forward, fast path --
// kLanes = D/4 lanes per (b,q,h) tuple; 32/kLanes tuples per warp.
//
// kProj = true is the same kernel fed by the attention module's raw projections instead of finished
// locations and weights (what mmcv computes with five elementwise launches between the Linear
// layers and the sampling kernel, multi_scale_deform_attn.py:322-349): `loc` then points at rows
// (B*Q, H*L*P*3) = [offsets (H,L,P,2) | logits (H,L,P)] and `ref` at reference points (B,Q,L,refdim);
// phase 1 forms  loc = ref + off / (W_l, H_l)   (refdim 2)  or  ref.xy + off / P * ref.wh * 0.5
// (refdim 4) with the same IEEE operations in the same order as the torch expressions, and the
// softmax over the L*P logits of a tuple with torch's own reduction shape (xor butterfly over
// L*P lanes: max, exp(x - max), sum, divide), so the result equals the unfused composition.
template <int kLanes, bool kProj>
__global__ void __launch_bounds__(kThreads) msda_fwd_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
    const float* __restrict__ attn, const float* __restrict__ ref, int refdim, int S, int H, int Q,
    int L, int P, long tuples, float* __restrict__ out) {
  // kProj: `attn`, when given, is a second projection tensor of the same shape added to `loc`'s rows
  // (the positional half of the query projection, constant per image geometry)
  constexpr int D = kLanes * 4;
  constexpr int kTuplesPerWarp = 32 / kLanes;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelInfo lv;
  load_levels(&lv, shapes, lsi, L);
  __syncthreads();

  const int LP = L * P;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int rec_per_warp = kTuplesPerWarp * LP;
  int4* s_off = reinterpret_cast<int4*>(smem_raw) + (size_t)warp * rec_per_warp;
  float4* s_w = reinterpret_cast<float4*>(smem_raw + (size_t)kWarps * rec_per_warp * 16) +
                (size_t)warp * rec_per_warp;
  float* s_a = reinterpret_cast<float*>(smem_raw + (size_t)kWarps * rec_per_warp * 32) +
               (size_t)warp * rec_per_warp;

  const long tuple0 = ((long)blockIdx.x * kWarps + warp) * kTuplesPerWarp;  // first tuple of warp
  if (tuple0 >= tuples) return;
  const int ntup = (int)min((long)kTuplesPerWarp, tuples - tuple0);

  // ---- phase 1: one lane per sample
  const int nrec = ntup * LP;
  if constexpr (!kProj) {
    for (int r = lane; r < nrec; r += 32) {
      const int t = r / LP;
      const int j = r - t * LP;
      const int l = j / P;
      const long tuple = tuple0 + t;
      const int head = (int)(tuple % H);
      const long s = tuple * LP + j;
      const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + s);
      const float a = __ldg(attn + s);
      const Bilinear bl = setup_sample(xy.x, xy.y, lv.h[l], lv.w[l], lv.start[l], H, D, head);
      s_off[r] = make_int4(bl.off[0], bl.off[1], bl.off[2], bl.off[3]);
      s_w[r] = make_float4(bl.w[0], bl.w[1], bl.w[2], bl.w[3]);
      s_a[r] = a;
    }
  } else {
    const int HLP = H * LP;
    const float inv_p = 1.f / (float)P;
    // LP is a power of two <= 32 here (host check): the LP samples of a tuple sit in LP adjacent lanes
    for (int r0 = 0; r0 < nrec; r0 += 32) {
      const int r = r0 + lane;
      const bool active = r < nrec;
      float logit = -INFINITY;
      float lx = 0.f, ly = 0.f;
      int l = 0, head = 0;
      if (active) {
        const int t = r / LP;
        const int j = r - t * LP;
        l = j / P;
        const long tuple = tuple0 + t;
        head = (int)(tuple % H);
        const long row = tuple / H;  // b*Q + q
        const float* pr = loc + row * (long)HLP * 3;
        float ox = __ldg(pr + ((long)head * LP + j) * 2);
        float oy = __ldg(pr + ((long)head * LP + j) * 2 + 1);
        logit = __ldg(pr + (long)HLP * 2 + head * LP + j);
        if (attn) {
          const float* pa = attn + row * (long)HLP * 3;
          ox = __fadd_rn(ox, __ldg(pa + ((long)head * LP + j) * 2));
          oy = __fadd_rn(oy, __ldg(pa + ((long)head * LP + j) * 2 + 1));
          logit = __fadd_rn(logit, __ldg(pa + (long)HLP * 2 + head * LP + j));
        }
        const float* rp = ref + (row * L + l) * refdim;
        if (refdim == 2) {
          lx = __fadd_rn(__ldg(rp), __fdiv_rn(ox, (float)lv.w[l]));
          ly = __fadd_rn(__ldg(rp + 1), __fdiv_rn(oy, (float)lv.h[l]));
        } else {
          lx = __fadd_rn(__ldg(rp), __fmul_rn(__fmul_rn(__fmul_rn(ox, inv_p), __ldg(rp + 2)), 0.5f));
          ly = __fadd_rn(__ldg(rp + 1), __fmul_rn(__fmul_rn(__fmul_rn(oy, inv_p), __ldg(rp + 3)), 0.5f));
        }
      }
      float mx = logit;
      for (int o = LP >> 1; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = active ? expf(logit - mx) : 0.f;
      float sum = e;
      for (int o = LP >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (active) {
        const Bilinear bl = setup_sample(lx, ly, lv.h[l], lv.w[l], lv.start[l], H, D, head);
        s_off[r] = make_int4(bl.off[0], bl.off[1], bl.off[2], bl.off[3]);
        s_w[r] = make_float4(bl.w[0], bl.w[1], bl.w[2], bl.w[3]);
        s_a[r] = __fdiv_rn(e, sum);
      }
    }
  }
  __syncwarp();

  // ---- phase 2: kLanes lanes per tuple, 4 channels per lane
  const int g = lane / kLanes;
  const int c4 = (lane % kLanes) * 4;
  if (g >= ntup) return;
  const long tuple = tuple0 + g;
  const long b = tuple / ((long)Q * H);
  const float* vb = value + b * (long)S * H * D + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int r0 = g * LP;
#pragma unroll 4
  for (int j = 0; j < LP; ++j) {
    const int4 o = s_off[r0 + j];
    const float4 w = s_w[r0 + j];
    const float a = s_a[r0 + j];
    const float4 v1 = ldg4(vb + o.x);
    const float4 v2 = ldg4(vb + o.y);
    const float4 v3 = ldg4(vb + o.z);
    const float4 v4 = ldg4(vb + o.w);
    acc.x = __fmaf_rn(bil(w.x, v1.x, w.y, v2.x, w.z, v3.x, w.w, v4.x), a, acc.x);
    acc.y = __fmaf_rn(bil(w.x, v1.y, w.y, v2.y, w.z, v3.y, w.w, v4.y), a, acc.y);
    acc.z = __fmaf_rn(bil(w.x, v1.z, w.y, v2.z, w.z, v3.z, w.w, v4.z), a, acc.z);
    acc.w = __fmaf_rn(bil(w.x, v1.w, w.y, v2.w, w.z, v3.w, w.w, v4.w), a, acc.w);
  }
  *reinterpret_cast<float4*>(out + tuple * D + c4) = acc;
}

// ------------------------------------------------------- forward, any D (scalar path) --
// One thread per output scalar (upstream's mapping); used when D is not 4*2^k.
__global__ void __launch_bounds__(kThreads) msda_fwd_generic_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
    const float* __restrict__ attn, int S, int H, int D, int Q, int L, int P, long total,
    float* __restrict__ out) {
  __shared__ LevelInfo lv;
  load_levels(&lv, shapes, lsi, L);
  __syncthreads();
  const int LP = L * P;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long tuple = e / D;
    const int head = (int)(tuple % H);
    const long b = tuple / ((long)Q * H);
    const float* vb = value + b * (long)S * H * D + c;
    float col = 0.f;
    for (int j = 0; j < LP; ++j) {
      const long s = tuple * LP + j;
      const int l = j / P;
      const Bilinear bl = setup_sample(__ldg(loc + 2 * s), __ldg(loc + 2 * s + 1), lv.h[l], lv.w[l],
                                       lv.start[l], H, D, head);
      if (bl.valid & 16u) {
        const float v = bil(bl.w[0], __ldg(vb + bl.off[0]), bl.w[1], __ldg(vb + bl.off[1]), bl.w[2],
                            __ldg(vb + bl.off[2]), bl.w[3], __ldg(vb + bl.off[3]));
        col = __fmaf_rn(v, __ldg(attn + s), col);
      }
    }
    out[e] = col;
  }
}

// ----------------------------------------------------------------- backward, fast path --
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int kLanes>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = kLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int kLanes>
__global__ void __launch_bounds__(kThreads) msda_bwd_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
    const float* __restrict__ attn, const float* __restrict__ grad_out, int S, int H, int Q, int L,
    int P, long tuples, float* __restrict__ grad_value, float* __restrict__ grad_loc,
    float* __restrict__ grad_attn) {
  constexpr int D = kLanes * 4;
  constexpr int kTuplesPerWarp = 32 / kLanes;
  __shared__ LevelInfo lv;
  load_levels(&lv, shapes, lsi, L);
  __syncthreads();

  const int LP = L * P;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const long tuple0 = ((long)blockIdx.x * kWarps + warp) * kTuplesPerWarp;
  if (tuple0 >= tuples) return;
  const int g = lane / kLanes;
  const int sub = lane % kLanes;
  const int c4 = sub * 4;
  // Lanes of a tuple beyond the end still take part in the shuffles; they work on a clamped
  // tuple and write nothing.
  const bool live = tuple0 + g < tuples;
  const long tuple = live ? tuple0 + g : tuples - 1;
  const int head = (int)(tuple % H);
  const long b = tuple / ((long)Q * H);
  const float* vb = value + b * (long)S * H * D + c4;
  float* gvb = grad_value + b * (long)S * H * D + c4;
  const float4 go = ldg4(grad_out + tuple * D + c4);

  for (int j = 0; j < LP; ++j) {
    const long s = tuple * LP + j;
    const int l = j / P;
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + s);
    const float a = __ldg(attn + s);
    const int sh = lv.h[l], sw = lv.w[l];
    const Bilinear bl = setup_sample(xy.x, xy.y, sh, sw, lv.start[l], H, D, head);
    float g_a = 0.f, g_w = 0.f, g_h = 0.f;
    if (bl.valid & 16u) {
      const float lh = bl.lh, lw = bl.lw, hh = 1.f - lh, hw = 1.f - lw;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (bl.valid >> k) & 1u ? ldg4(vb + bl.off[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float tg[4] = {go.x, go.y, go.z, go.w};
      const float vv[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x},
                              {v[0].y, v[1].y, v[2].y, v[3].y},
                              {v[0].z, v[1].z, v[2].z, v[3].z},
                              {v[0].w, v[1].w, v[2].w, v[3].w}};
      float tgv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float top = tg[c];
        const float top_v = top * a;
        tgv[c] = top_v;
        // upstream col2im_bilinear, corner by corner
        const float gh = -hw * vv[c][0] - lw * vv[c][1] + hw * vv[c][2] + lw * vv[c][3];
        const float gw = -hh * vv[c][0] + hh * vv[c][1] - lh * vv[c][2] + lh * vv[c][3];
        const float val = bil(bl.w[0], vv[c][0], bl.w[1], vv[c][1], bl.w[2], vv[c][2], bl.w[3],
                              vv[c][3]);
        g_a += top * val;
        g_w += (float)sw * gw * top_v;
        g_h += (float)sh * gh * top_v;
      }
      if (live) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((bl.valid >> k) & 1u)
            red_add_v4(gvb + bl.off[k], bl.w[k] * tgv[0], bl.w[k] * tgv[1], bl.w[k] * tgv[2],
                       bl.w[k] * tgv[3]);
      }
    }
    // sum over the D channels of the head (kLanes lanes x 4 channels)
    g_a = group_sum<kLanes>(g_a);
    g_w = group_sum<kLanes>(g_w);
    g_h = group_sum<kLanes>(g_h);
    if (live && sub == 0) {
      grad_attn[s] = g_a;
      reinterpret_cast<float2*>(grad_loc)[s] = make_float2(g_w, g_h);
    }
  }
}

// ------------------------------------------------------ backward, any D (scalar path) --
// One warp per (b,q,h) tuple, lanes stride the channels; scalar atomics.
__global__ void __launch_bounds__(kThreads) msda_bwd_generic_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
    const float* __restrict__ attn, const float* __restrict__ grad_out, int S, int H, int D, int Q,
    int L, int P, long tuples, float* __restrict__ grad_value, float* __restrict__ grad_loc,
    float* __restrict__ grad_attn) {
  __shared__ LevelInfo lv;
  load_levels(&lv, shapes, lsi, L);
  __syncthreads();
  const int LP = L * P;
  const unsigned lane = lane_id();
  const long tuple = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  if (tuple >= tuples) return;
  const int head = (int)(tuple % H);
  const long b = tuple / ((long)Q * H);
  const float* vb = value + b * (long)S * H * D;
  float* gvb = grad_value + b * (long)S * H * D;
  for (int j = 0; j < LP; ++j) {
    const long s = tuple * LP + j;
    const int l = j / P;
    const int sh = lv.h[l], sw = lv.w[l];
    const float a = __ldg(attn + s);
    const Bilinear bl = setup_sample(__ldg(loc + 2 * s), __ldg(loc + 2 * s + 1), sh, sw,
                                     lv.start[l], H, D, head);
    float g_a = 0.f, g_w = 0.f, g_h = 0.f;
    if (bl.valid & 16u) {
      const float lh = bl.lh, lw = bl.lw, hh = 1.f - lh, hw = 1.f - lw;
      for (int c = lane; c < D; c += 32) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (bl.valid >> k) & 1u ? __ldg(vb + bl.off[k] + c) : 0.f;
        const float top = __ldg(grad_out + tuple * D + c);
        const float top_v = top * a;
        const float gh = -hw * v[0] - lw * v[1] + hw * v[2] + lw * v[3];
        const float gw = -hh * v[0] + hh * v[1] - lh * v[2] + lh * v[3];
        g_a += top * bil(bl.w[0], v[0], bl.w[1], v[1], bl.w[2], v[2], bl.w[3], v[3]);
        g_w += (float)sw * gw * top_v;
        g_h += (float)sh * gh * top_v;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((bl.valid >> k) & 1u) atomicAdd(gvb + bl.off[k] + c, bl.w[k] * top_v);
      }
    }
    g_a = group_sum<32>(g_a);
    g_w = group_sum<32>(g_w);
    g_h = group_sum<32>(g_h);
    if (lane == 0) {
      grad_attn[s] = g_a;
      grad_loc[2 * s] = g_w;
      grad_loc[2 * s + 1] = g_h;
    }
  }
}

int check_dims(int B, int S, int H, int D, int Q, int L, int P) {
  if (B < 0 || S <= 0 || H <= 0 || D <= 0 || Q < 0 || L <= 0 || P <= 0) {
    set_error("msda: bad sizes B=%d S=%d H=%d D=%d Q=%d L=%d P=%d", B, S, H, D, Q, L, P);
    return DEMF_E_SIZE;
  }
  if (L > kMaxLevels) {
    set_error("msda: at most %d levels are supported (got %d)", kMaxLevels, L);
    return DEMF_E_UNSUPPORTED;
  }
  // offsets inside one batch element are 32-bit
  if ((long)S * H * D >= (1L << 31)) {
    set_error("msda: S*H*D = %ld exceeds 32-bit offsets", (long)S * H * D);
    return DEMF_E_SIZE;
  }
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline bool fast_path_lanes(int D, int* lanes) {
  if (D % 4) return false;
  const int l = D / 4;
  if (l < 1 || l > 32 || (l & (l - 1))) return false;
  *lanes = l;
  return true;
}

template <int kLanes, bool kProj = false>
int launch_fwd(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
               const float* attn, int S, int H, int Q, int L, int P, long tuples, float* out,
               cudaStream_t st, const float* ref = nullptr, int refdim = 0) {
  constexpr int tpw = 32 / kLanes;
  const size_t smem = (size_t)kWarps * tpw * L * P * 36;
  auto k = msda_fwd_kernel<kLanes, kProj>;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return -1;  // caller falls back to the scalar path
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const long blocks = (tuples + (long)kWarps * tpw - 1) / ((long)kWarps * tpw);
  k<<<(unsigned)blocks, kThreads, smem, st>>>(value, shapes, lsi, loc, attn, ref, refdim, S, H, Q, L, P,
                                             tuples, out);
  return after_launch(kProj ? "msda_fwd_kernel<proj>" : "msda_fwd_kernel");
}

template <int kLanes>
int launch_bwd(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
               const float* attn, const float* grad_out, int S, int H, int Q, int L, int P,
               long tuples, float* gv, float* gl, float* ga, cudaStream_t st) {
  constexpr int tpw = 32 / kLanes;
  const long blocks = (tuples + (long)kWarps * tpw - 1) / ((long)kWarps * tpw);
  msda_bwd_kernel<kLanes><<<(unsigned)blocks, kThreads, 0, st>>>(value, shapes, lsi, loc, attn,
                                                                grad_out, S, H, Q, L, P, tuples, gv,
                                                                gl, ga);
  return after_launch("msda_bwd_kernel");
}

}  // namespace
DEMF_DEFINE_TRACE_SETTER(trace_set_msda)
}  // namespace demf

using namespace demf;

extern "C" {

int demf_msda_fwd(const float* value, const int64_t* spatial_shapes,
                  const int64_t* level_start_index, const float* sampling_loc,
                  const float* attn_weight, int B, int S, int H, int D, int Q, int L, int P,
                  float* out, void* stream) {
  DEMF_REQUIRE_PTR(value);
  DEMF_REQUIRE_PTR(spatial_shapes);
  DEMF_REQUIRE_PTR(level_start_index);
  DEMF_REQUIRE_PTR(sampling_loc);
  DEMF_REQUIRE_PTR(attn_weight);
  DEMF_REQUIRE_PTR(out);
  if (int rc = check_dims(B, S, H, D, Q, L, P)) return rc;
  const long tuples = (long)B * Q * H;
  if (tuples == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int lanes = 0;
  int rc = -1;
  if (fast_path_lanes(D, &lanes) && aligned16(value) && aligned16(out) && aligned16(sampling_loc)) {
    switch (lanes) {
#define DEMF_CASE(n)                                                                             \
  case n:                                                                                        \
    rc = launch_fwd<n>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, S, H, \
                       Q, L, P, tuples, out, st);                                                \
    break;
      DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(4) DEMF_CASE(8) DEMF_CASE(16) DEMF_CASE(32)
#undef DEMF_CASE
    }
  }
  if (rc != -1) return rc;
  const long total = tuples * D;
  long blocks = (total + kThreads - 1) / kThreads;
  if (blocks > (long)kNumSMs * 32) blocks = (long)kNumSMs * 32;
  msda_fwd_generic_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(
      value, spatial_shapes, level_start_index, sampling_loc, attn_weight, S, H, D, Q, L, P, total,
      out);
  return after_launch("msda_fwd_generic_kernel");
}

int demf_msda_proj_fwd_supported(int D, int L, int P) {
  int lanes = 0;
  const long lp = (long)L * P;
  // one tuple's L*P samples must fill whole shuffle groups and a warp's records whole passes
  return fast_path_lanes(D, &lanes) && lp >= 1 && lp <= 32 && (lp & (lp - 1)) == 0 &&
         ((32 / lanes) * lp) % 32 == 0 && L <= kMaxLevels;
}

int demf_msda_proj_fwd(const float* value, const int64_t* spatial_shapes,
                       const int64_t* level_start_index, const float* proj, const float* proj_add,
                       const float* ref_points, int ref_dim, int B, int S, int H, int D, int Q, int L,
                       int P, float* out, void* stream) {
  DEMF_REQUIRE_PTR(value);
  DEMF_REQUIRE_PTR(spatial_shapes);
  DEMF_REQUIRE_PTR(level_start_index);
  DEMF_REQUIRE_PTR(proj);
  DEMF_REQUIRE_PTR(ref_points);
  DEMF_REQUIRE_PTR(out);
  if (int rc = check_dims(B, S, H, D, Q, L, P)) return rc;
  if (ref_dim != 2 && ref_dim != 4) {
    set_error("demf_msda_proj_fwd: ref_dim must be 2 or 4 (got %d)", ref_dim);
    return DEMF_E_SIZE;
  }
  if (!demf_msda_proj_fwd_supported(D, L, P) || !aligned16(value) || !aligned16(out)) {
    set_error("demf_msda_proj_fwd: unsupported D=%d, L*P=%d or unaligned buffers "
              "(see demf_msda_proj_fwd_supported)", D, L * P);
    return DEMF_E_UNSUPPORTED;
  }
  const long tuples = (long)B * Q * H;
  if (tuples == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int lanes = 0;
  fast_path_lanes(D, &lanes);
  int rc = -1;
  switch (lanes) {
#define DEMF_CASE(n)                                                                                 \
  case n:                                                                                            \
    rc = launch_fwd<n, true>(value, spatial_shapes, level_start_index, proj, proj_add, S, H, Q, L, P, \
                             tuples, out, st, ref_points, ref_dim);                                  \
    break;
    DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(4) DEMF_CASE(8) DEMF_CASE(16) DEMF_CASE(32)
#undef DEMF_CASE
  }
  if (rc == -1) {
    set_error("demf_msda_proj_fwd: shared memory request too large for L*P=%d", L * P);
    return DEMF_E_UNSUPPORTED;
  }
  return rc;
}

int demf_msda_bwd(const float* value, const int64_t* spatial_shapes,
                  const int64_t* level_start_index, const float* sampling_loc,
                  const float* attn_weight, const float* grad_out, int B, int S, int H, int D, int Q,
                  int L, int P, float* grad_value, float* grad_sampling_loc, float* grad_attn_weight,
                  void* stream) {
  DEMF_REQUIRE_PTR(value);
  DEMF_REQUIRE_PTR(spatial_shapes);
  DEMF_REQUIRE_PTR(level_start_index);
  DEMF_REQUIRE_PTR(sampling_loc);
  DEMF_REQUIRE_PTR(attn_weight);
  DEMF_REQUIRE_PTR(grad_out);
  DEMF_REQUIRE_PTR(grad_value);
  DEMF_REQUIRE_PTR(grad_sampling_loc);
  DEMF_REQUIRE_PTR(grad_attn_weight);
  if (int rc = check_dims(B, S, H, D, Q, L, P)) return rc;
  const long tuples = (long)B * Q * H;
  if (tuples == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int lanes = 0;
  if (fast_path_lanes(D, &lanes) && aligned16(value) && aligned16(grad_value) &&
      aligned16(grad_out) && aligned16(sampling_loc) && aligned16(grad_sampling_loc)) {
    switch (lanes) {
#define DEMF_CASE(n)                                                                             \
  case n:                                                                                        \
    return launch_bwd<n>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,    \
                         grad_out, S, H, Q, L, P, tuples, grad_value, grad_sampling_loc,         \
                         grad_attn_weight, st);
      DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(4) DEMF_CASE(8) DEMF_CASE(16) DEMF_CASE(32)
#undef DEMF_CASE
    }
  }
  const long blocks = (tuples * 32 + kThreads - 1) / kThreads;
  DEMF_REQUIRE(blocks < (1L << 31), DEMF_E_SIZE);
  msda_bwd_generic_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(
      value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_out, S, H, D, Q, L,
      P, tuples, grad_value, grad_sampling_loc, grad_attn_weight);
  return after_launch("msda_bwd_generic_kernel");
}

}  // extern "C"
