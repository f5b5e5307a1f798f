// Set abstraction, first level, in inference -- a warp-specialised PIPELINE over 128-row tiles.
//
// Same contract and arithmetic as sa_fused.cu (ball-query rows -> grouping -> 3 x [1x1 conv, folded BN, bias,
// ReLU] -> max over the neighbourhood; TF32 products, fp32 accumulation; replaces, for mmdet3d
// `PointSAModule.forward` in eval mode -- configs/demf/demf_votenet.py:48-62 -- group_points + cuDNN convs +
// max_pool2d), for the geometry whose three weight matrices fit shared memory together: C = 1 input feature
// (+ xyz: 8-wide rows), widths 64 / 64 / 128 -- the backbone's first level, the largest of the five (131 072
// grouped rows per scene, 0.20 of the 0.57 ms the fused kernels took per step in round 1).
//
// sa_fused.cu gives each 128-row tile to a "lane" of four warps that walks gather -> MMA 0 -> epilogue 0 ->
// MMA 1 -> epilogue 1 -> MMA 2 -> max strictly in sequence; four lanes per SM overlap four such chains, and a
// tile costs the SUM of its phases' latencies (16 k cycles per lane, 4 k per SM). Here every phase has its own
// warps and the tiles flow through them:
//
//   warps 4-7    gather     neighbour indices (coalesced) -> point + feature (one 16-byte load from the packed
//                           (B,N,4) cloud when the caller has it) -> (xyz - centre) / r -> the layer-0 operand A0
//                           (K-major SWIZZLE_128B rows, 8 floats used); batches of four tiles, eight-slot ring
//   warps 0-2    MMA        tcgen05.mma kind::tf32, ONE issuing thread per layer, each in its own loop over tiles
//                           (tcgen05.commit tracks the committing thread's instructions only):
//                           L0: D0 = A0 W0^T (K 8)   L1: D1 = A1 W1^T (K 64)   L2: D2^T = W2 A2^T (K 64)
//   warps 8-15   epilogue 0 D0 (TMEM, 32 columns per warp) -> + b0, ReLU, tf32 -> A1 (shared memory)
//   warps 16-23  epilogue 1 D1 -> + b1, ReLU, tf32 -> A2
//   warps 24-31  max        D2^T: TMEM lanes = the 128 output channels, columns = the tile's rows, so the max over
//                           a centre's ns rows is an in-thread max chain; + b2, ReLU, coalesced store
//
// Every buffer (A1, A2, the three accumulators) exists twice -- A0 eight times -- and is handed over through mbarrier
// pairs; the weights (56 KB, host-packed operand images of sa_pack_weights) are resident. A CTA takes a contiguous
// range of tiles. A tile then costs the SLOWEST phase instead of the sum: measured (DEMF_SAP_PROF build, ncu source
// counters) that is the issue slots of the epilogue / max warps, which is why their instruction streams are pared
// down (FADD2 bias, VIADDMNMX ReLU + rounding, FMNMX3) and the roles are eight warps wide.
#include "common.cuh"
#include "umma.cuh"

namespace demf {
namespace {

using namespace umma;

__device__ int g_sap_error = 0;  // sticky: first mbarrier time-out (never expected)

constexpr int kRows = 128;
constexpr int kC1 = 64, kC2 = 64, kC3 = 128;
constexpr int kThreads = 32 * 32;   // warps 0-2: one MMA issuer per layer; 4-7 gather; 8-15, 16-23 epilogues; 24-31 max
constexpr int kW0 = 0, kW1 = 8192, kW2 = kW1 + 16384, kBias = kW2 + 32768;     // byte offsets in shared memory
constexpr int kA0 = kBias + 1024, kA1 = kA0 + 2 * 16384, kA2 = kA1 + 2 * 32768, kBars = kA2 + 2 * 32768;
constexpr int kSmemBytes = kBars + 512 + 1024;
constexpr uint32_t kD0 = 0, kD1 = 128, kD2 = 256;   // TMEM columns: D0[b] = b*64, D1[b] = 128 + b*64, D2T[b] = 256 + b*128

struct SaPipeParams {
  const float* xyz;       // (B,N,3)
  const float* feat;      // (B,N,1)
  const float4* pts4;     // optional (B,N,4) = [xyz | feat] packed (the raw input cloud): one 16-byte load per row
  const float* centres;   // (B,M,3)
  const int32_t* nbr;     // (B,M,ns) ball-query rows
  const float* wpack;     // operand images of W0 (64x8 -> one chunk), W1 (64x64), W2 (128x64)
  const float* bias;      // b0[64], b1[64], b2[128]
  float* out;             // (B,M,128)
  int N, M, ns;
  float scale;            // 1/r when normalize_xyz else 1
  int tiles;              // B * M / (128 / ns)
};

#ifdef DEMF_SAP_PROF
__device__ long long g_sap_prof[16];   // [code] = cycles block 0's lane 0 of each role spent in each wait; [0] = total
#endif
#ifdef DEMF_SAP_PROF
#define SAP_T0() const long long sap_t0 = clock64()
#define SAP_T(code, cond) if (blockIdx.x == 0 && (cond)) atomicAdd((unsigned long long*)&g_sap_prof2[code], (unsigned long long)(clock64() - sap_t0))
__device__ long long g_sap_prof2[16];
#else
#define SAP_T0()
#define SAP_T(code, cond)
#endif
__device__ __forceinline__ void wait_flag(uint32_t bar, uint32_t parity, int code) {
#ifdef DEMF_SAP_PROF
  const long long t0 = clock64();
#endif
  // try_wait with a suspend-time hint: the warp sleeps in hardware instead of polling (the polling loops of 20+
  // waiting warps were ~40 % of all issued instructions); bounded, a protocol bug must not hang the GPU
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok && spin < (1u << 20); ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000u) : "memory");
  if (!ok) atomicCAS(&g_sap_error, 0, code);
#ifdef DEMF_SAP_PROF
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ((1u << (threadIdx.x >> 5)) & 0x01010117u)) atomicAdd((unsigned long long*)&g_sap_prof[code], (unsigned long long)(clock64() - t0));
#endif
}
__device__ __forceinline__ void arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bar.sync over one role's warps. The canonical hand-off of operands written with ordinary stores to the tensor
// core (async proxy): every WRITER executes fence.proxy.async, the role meets at this barrier, ONE thread arrives on
// the mbarrier the MMA issuer waits for (one arrive per tile instead of one per warp).
__device__ __forceinline__ void role_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float tf32_op(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__global__ void __launch_bounds__(kThreads, 1) sa_pipe_kernel(const SaPipeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
#ifdef DEMF_SAP_PROF
  const long long t_begin = clock64();
#endif
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBars);
  // barriers: a0_full[8], a0_free[8] (the layer-0 operand ring has 8 slots), then kind * 2 + buffer for the kinds
  //           2 d0_full 3 d0_free 4 a1_full 5 a1_free 6 d1_full 7 d1_free 8 a2_full 9 a2_free 10 d2_full 11 d2_free
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int kind, int b) { return bar0 + 8u * (kind < 2 ? kind * 8 + b : 16 + (kind - 2) * 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  if (threadIdx.x == 0) {
    for (int sl = 0; sl < 8; ++sl) {
      mbar_init(bar(0, sl), 1);  // one gather thread, after the role's bar.sync
      mbar_init(bar(1, sl), 1);  // tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(2, b), 1);
      mbar_init(bar(3, b), 8);   // eight epilogue-0 warps
      mbar_init(bar(4, b), 1);   // one epilogue-0 thread, after the role's bar.sync
      mbar_init(bar(5, b), 1);
      mbar_init(bar(6, b), 1);
      mbar_init(bar(7, b), 8);   // eight epilogue-1 warps
      mbar_init(bar(8, b), 1);
      mbar_init(bar(9, b), 1);
      mbar_init(bar(10, b), 1);
      mbar_init(bar(11, b), 8);  // eight max warps
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  // resident weights + biases: plain copies, then made visible to the tensor core's (async) proxy
  {
    const float4* src = reinterpret_cast<const float4*>(p.wpack);
    float4* dst = reinterpret_cast<float4*>(smem + kW0);
    for (int i = threadIdx.x; i < (8192 + 16384 + 32768) / 16; i += kThreads) dst[i] = __ldg(src + i);
    float* bs = reinterpret_cast<float*>(smem + kBias);
    for (int i = threadIdx.x; i < kC1 + kC2 + kC3; i += kThreads) bs[i] = __ldg(p.bias + i);
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const float* bias_s = reinterpret_cast<const float*>(smem + kBias);
  // this CTA's tiles: the contiguous range [t_first, t_first + nt)
  const int t_base = p.tiles / (int)gridDim.x, t_rem = p.tiles % (int)gridDim.x;
  const int nt = t_base + ((int)blockIdx.x < t_rem ? 1 : 0);
  const int t_first = (int)blockIdx.x * t_base + min((int)blockIdx.x, t_rem);
  const int cpt = kRows / p.ns;
  const int tps = p.M / cpt;                    // tiles per scene

  if (warp < 3) {
    // ------------------------------------------------------------------------------ MMA issuers, one thread per layer
    // (a single thread issuing all 17 instructions of a tile, with its six barrier waits in between, was the
    // slowest stage of the pipeline; tcgen05.commit tracks the instructions of the committing thread only)
    if (lane == 0) {
      const uint32_t id64 = instr_desc_tf32(128, 64), id128 = instr_desc_tf32(128, 128);
      const uint32_t w0 = smem_u32(smem + kW0), w1 = smem_u32(smem + kW1), w2 = smem_u32(smem + kW2);
      if (warp == 0) {
        for (int i = 0; i < nt; ++i) {                         // layer 0: D0 = A0 W0^T, K = 8
          const int b = (int)(i & 1);
          const uint32_t ph = (uint32_t)((i >> 1) & 1);
          // A0 ring: slot = i % 8 = (16 KB chunk, 32-byte k slice): four tiles share one SWIZZLE_128B chunk, each
          // in its own 8-float k step (the descriptor start advances by 32 bytes, as for any k step)
          const int sl = (int)(i & 7);
          wait_flag(bar(0, sl), (uint32_t)((i >> 3) & 1), 1);
          wait_flag(bar(3, b), ph ^ 1u, 2);
          tc_fence_after_sync();
          mma_tf32(tmem + kD0 + b * 64u, smem_desc_sw128(smem_u32(smem + kA0 + (sl >> 2) * 16384) + (sl & 3) * 32),
                   smem_desc_sw128(w0), id64, 0u);
          mma_commit(bar(1, sl));
          mma_commit(bar(2, b));
        }
      } else if (warp == 1) {
        for (int j = 0; j < nt; ++j) {                         // layer 1: D1 = A1 W1^T, K = 64
          const int b = (int)(j & 1);
          const uint32_t ph = (uint32_t)((j >> 1) & 1);
          wait_flag(bar(4, b), ph, 3);
          wait_flag(bar(7, b), ph ^ 1u, 4);
          tc_fence_after_sync();
          SAP_T0();
          const uint32_t a = smem_u32(smem + kA1 + b * 32768);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_tf32(tmem + kD1 + b * 64u, smem_desc_sw128(a + kc * 16384 + ks * 32),
                       smem_desc_sw128(w1 + kc * 8192 + ks * 32), id64, (kc | ks) ? 1u : 0u);
          mma_commit(bar(5, b));
          mma_commit(bar(6, b));
          SAP_T(3, true);
        }
      } else {
        for (int j = 0; j < nt; ++j) {                         // layer 2, transposed: D2^T = W2 A2^T, K = 64
          const int b = (int)(j & 1);
          const uint32_t ph = (uint32_t)((j >> 1) & 1);
          wait_flag(bar(8, b), ph, 5);
          wait_flag(bar(11, b), ph ^ 1u, 6);
          tc_fence_after_sync();
          const uint32_t a = smem_u32(smem + kA2 + b * 32768);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_tf32(tmem + kD2 + b * 128u, smem_desc_sw128(w2 + kc * 16384 + ks * 32),
                       smem_desc_sw128(a + kc * 16384 + ks * 32), id128, (kc | ks) ? 1u : 0u);
          mma_commit(bar(9, b));
          mma_commit(bar(10, b));
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------------------ gather
    // Each thread owns one grouped row of every tile and runs ahead of itself (see the register ring below), so
    // neither dependent global-memory latency is exposed; the eight-slot operand ring lets the warps run ahead
    // of the MMA.
    const int row = (warp - 4) * 32 + (int)lane;
    const int rc = row / p.ns;
    struct Pt { float x, y, z, f, cx, cy, cz; };
    // tile T = grouped rows [128 T, 128 T + 128) = centres [T cpt, (T + 1) cpt) of scene T / tps
    auto nbr_of = [&](int j) -> int {
      return __ldg(p.nbr + (long)(t_first + j) * kRows + row);
    };
    int pt_scene = t_first / tps, pt_t = t_first - pt_scene * tps;   // the tile whose points are loaded next
    auto load_pt = [&](int j, int k) -> Pt {
      const float* c = p.centres + ((long)(t_first + j) * cpt + rc) * 3;
      const float* pt = p.xyz + ((long)pt_scene * p.N + k) * 3;
      Pt r;
      if (p.pts4 != nullptr) {   // one sector per row instead of four scalar requests (the L1 tag stage is what
        const float4 v = __ldg(p.pts4 + (long)pt_scene * p.N + k);   // a fully scattered gather is bound by)
        r.x = v.x;
        r.y = v.y;
        r.z = v.z;
        r.f = v.w;
      } else {
        r.x = __ldg(pt);
        r.y = __ldg(pt + 1);
        r.z = __ldg(pt + 2);
        r.f = __ldg(p.feat + (long)pt_scene * p.N + k);
      }
      r.cx = __ldg(c);
      r.cy = __ldg(c + 1);
      r.cz = __ldg(c + 2);
      if (++pt_t == tps) {
        pt_t = 0;
        ++pt_scene;
      }
      return r;
    };
    // Batches of kBatch tiles. fence.proxy.async is a full memory barrier for the thread (MEMBAR.ALL.CTA): it
    // also waits for global loads in flight, so the loads of the next batch are issued AFTER the fence of this one,
    // and one exposed load latency is shared by kBatch tiles (the eight-slot operand ring holds two batches, so
    // the MMA never waits for it).
    constexpr int kBatch = 4;
    Pt ring[kBatch];
    int knext[kBatch];
#pragma unroll
    for (int d = 0; d < kBatch; ++d) {
      ring[d] = Pt{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      knext[d] = 0;
      if (d < nt) knext[d] = nbr_of(d);
    }
#pragma unroll
    for (int d = 0; d < kBatch; ++d) {
      if (d < nt) ring[d] = load_pt(d, knext[d]);
      if (d + kBatch < nt) knext[d] = nbr_of(d + kBatch);
    }
    for (int j0 = 0; j0 < nt; j0 += kBatch) {
#pragma unroll
      for (int d = 0; d < kBatch; ++d) {
        const int j = j0 + d;
        if (j < nt) {
          const Pt cur = ring[d];
          const float dx = __fmul_rn(__fsub_rn(cur.x, cur.cx), p.scale);
          const float dy = __fmul_rn(__fsub_rn(cur.y, cur.cy), p.scale);
          const float dz = __fmul_rn(__fsub_rn(cur.z, cur.cz), p.scale);
          const int sl = j & 7;
          wait_flag(bar(1, sl), (uint32_t)(((j >> 3) & 1) ^ 1), 7);   // layer-0 MMA of the tile that last used the slot
          unsigned char* a0 = smem + kA0 + (sl >> 2) * 16384;
          const uint32_t u = (uint32_t)(sl & 3) * 2u;
          *reinterpret_cast<float4*>(a0 + sw128_offset((uint32_t)row, u)) =
              make_float4(tf32_op(cur.f), tf32_op(0.f), tf32_op(0.f), tf32_op(0.f));
          *reinterpret_cast<float4*>(a0 + sw128_offset((uint32_t)row, u + 1u)) =
              make_float4(tf32_op(dx), tf32_op(dy), tf32_op(dz), tf32_op(0.f));
        }
      }
      fence_proxy_async();
      role_sync(1, 128);
      if (warp == 4 && lane == 0) {
#pragma unroll
        for (int d = 0; d < kBatch; ++d)
          if (j0 + d < nt) arrive(bar(0, (j0 + d) & 7));
      }
#pragma unroll
      for (int d = 0; d < kBatch; ++d)
        if (j0 + d + kBatch < nt) ring[d] = load_pt(j0 + d + kBatch, knext[d]);
#pragma unroll
      for (int d = 0; d < kBatch; ++d)
        if (j0 + d + 2 * kBatch < nt) knext[d] = nbr_of(j0 + d + 2 * kBatch);
    }
  } else if (warp >= 8 && warp < 24) {
    // ------------------------------------------------------------------------------ epilogues 0 and 1
    // Eight warps each: warp (q, h) turns TMEM lanes 32q.. (its quarter) x columns 32h.. of the accumulator into
    // one 32-float half row of the next layer's operand (chunk h of the SWIZZLE_128B image).
    const int e = (warp - 8) >> 3;                    // 0: D0 -> A1 (bias b0), 1: D1 -> A2 (bias b1)
    const int q = warp & 3, h = ((warp - 8) >> 2) & 1;
    const int row = q * 32 + (int)lane;
    const float* bs = bias_s + e * kC1 + h * 32;
    const int k_full = e ? 6 : 2, k_dfree = e ? 7 : 3, k_afull = e ? 8 : 4, k_afree = e ? 9 : 5;
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      wait_flag(bar(k_full, b), ph, 8 + e);
      tc_fence_after_sync();
      SAP_T0();
      uint32_t u0[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (e ? kD1 : kD0) + b * 64u + h * 32u, u0);
      const float4* bs4 = reinterpret_cast<const float4*>(bs);
      float4 bb[4];                                   // bias of the first 16 columns; the rest follows in the loop
#pragma unroll
      for (int s = 0; s < 4; ++s) bb[s] = bs4[s];
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(bar(k_dfree, b));         // the accumulator may be overwritten
      SAP_T(0, threadIdx.x == 256);
      wait_flag(bar(k_afree, b), ph ^ 1u, 10 + e);    // the MMA that last read this operand buffer is done
      unsigned char* dst = smem + (e ? kA2 : kA1) + b * 32768 + h * 16384;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 v[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int c = 16 * half + 4 * s;
          add_pair(u0[c + 0], u0[c + 1], bb[s].x, bb[s].y, v[s].x, v[s].y);
          add_pair(u0[c + 2], u0[c + 3], bb[s].z, bb[s].w, v[s].z, v[s].w);
          v[s].x = relu_tf32_op(v[s].x);
          v[s].y = relu_tf32_op(v[s].y);
          v[s].z = relu_tf32_op(v[s].z);
          v[s].w = relu_tf32_op(v[s].w);
        }
        if (half == 0) {
#pragma unroll
          for (int s = 0; s < 4; ++s) bb[s] = bs4[4 + s];
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
          *reinterpret_cast<float4*>(dst + sw128_offset((uint32_t)row, (uint32_t)(4 * half + s))) = v[s];
      }
      SAP_T(1, threadIdx.x == 256);
      fence_proxy_async();
      role_sync(2 + e, 256);
      if (((warp - 8) & 7) == 0 && lane == 0) arrive(bar(k_afull, b));
      SAP_T(2, threadIdx.x == 256);
    }
  } else if (warp >= 24) {
    // ------------------------------------------------------------------------------ max over the neighbourhood
    // Eight warps: warp (q, h) reduces channels 32q.. (TMEM lanes) over grouped rows 64h..64h+63 (columns of D2^T).
    const int q = warp & 3, h = (warp - 24) >> 2;
    const int ch = q * 32 + (int)lane;                // output channel = TMEM lane
    const float b2 = bias_s[kC1 + kC2 + ch];
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      float* o = p.out + ((long)(t_first + j) * cpt) * kC3 + ch;   // centre (scene, m) is row T cpt + .. of (B*M, 128)
      wait_flag(bar(10, b), ph, 12);
      tc_fence_after_sync();
      float run = -3.0e38f;
      SAP_T0();
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        const int blk = 2 * h + i2;
        uint32_t u[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + kD2 + b * 128u + blk * 32u, u);
        tmem_ld_wait();
        if (p.ns == 16) {
          float ma = __uint_as_float(u[0]), mb = __uint_as_float(u[16]);
#pragma unroll
          for (int i = 1; i < 15; i += 2) {
            ma = max3(ma, __uint_as_float(u[i]), __uint_as_float(u[i + 1]));
            mb = max3(mb, __uint_as_float(u[16 + i]), __uint_as_float(u[17 + i]));
          }
          ma = fmaxf(ma, __uint_as_float(u[15]));
          mb = fmaxf(mb, __uint_as_float(u[31]));
          o[(long)(2 * blk) * kC3] = fmaxf(ma + b2, 0.f);
          o[(long)(2 * blk + 1) * kC3] = fmaxf(mb + b2, 0.f);
        } else {
          float mx = __uint_as_float(u[0]);
#pragma unroll
          for (int i = 1; i < 31; i += 2) mx = max3(mx, __uint_as_float(u[i]), __uint_as_float(u[i + 1]));
          mx = fmaxf(mx, __uint_as_float(u[31]));
          run = fmaxf(run, mx);
          const int per = p.ns >> 5;                   // 32-row blocks per centre (1 or 2)
          if ((blk + 1) % per == 0) {
            o[(long)(blk / per) * kC3] = fmaxf(run + b2, 0.f);
            run = -3.0e38f;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(bar(11, b));
      SAP_T(4, threadIdx.x == 768);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
#ifdef DEMF_SAP_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd((unsigned long long*)&g_sap_prof[0], (unsigned long long)(clock64() - t_begin));
    atomicAdd((unsigned long long*)&g_sap_prof[15], (unsigned long long)nt);
  }
#endif
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_free(tmem, 512);
  }
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_sa_pipe_supported(int C, int ns, int c1, int c2, int c3, int M) {
  return (C == 1 && (ns == 16 || ns == 32 || ns == 64) && c1 == kC1 && c2 == kC2 && c3 == kC3 && M > 0 &&
          M % (kRows / ns) == 0)
             ? 1
             : 0;
}

#ifdef DEMF_SAP_PROF
int demf_sa_pipe_profile(long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_sap_prof, sizeof(long long) * 16);
  cudaMemcpyFromSymbol(out16 + 16, g_sap_prof2, sizeof(long long) * 16);
  long long z[16] = {0};
  cudaMemcpyToSymbol(g_sap_prof, z, sizeof(z));
  cudaMemcpyToSymbol(g_sap_prof2, z, sizeof(z));
  return 0;
}
#endif

int demf_sa_pipe_error(void) {
  int e = 0;
  cudaMemcpyFromSymbol(&e, g_sap_error, sizeof(int));
  return e;
}

/* The pipelined first-level set-abstraction kernel: idx = ball-query rows (B,M,ns) of the same query (e.g.
 * demf_ball_query_grid); wpack / bias as demf_sa_pack_weights lays them out for widths (64, 64, 128) over
 * 8-wide rows [feat | 0 0 0 | (xyz - centre)/r | 0]; out (B,M,128). Same results as demf_sa_fused_fwd.
 * points4 (optional, may be NULL): the same cloud packed as (B,N,4) rows [x y z feat], 16-byte aligned -- the
 * gather then needs one load per neighbour instead of four. */
int demf_sa_pipe_fwd(const float* xyz, const float* feat_rows, const float* points4, const float* new_xyz,
                     const int32_t* idx, int B, int N, int M, int ns, float max_radius, int normalize_xyz,
                     const float* wpack, const float* bias, float* out, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(feat_rows);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(wpack);
  DEMF_REQUIRE_PTR(bias);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(B > 0 && N > 0 && M > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(demf_sa_pipe_supported(1, ns, kC1, kC2, kC3, M), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15u) == 0, DEMF_E_UNSUPPORTED);
  SaPipeParams p;
  p.xyz = xyz;
  p.feat = feat_rows;
  DEMF_REQUIRE((reinterpret_cast<uintptr_t>(points4) & 15u) == 0, DEMF_E_UNSUPPORTED);
  p.pts4 = reinterpret_cast<const float4*>(points4);
  p.centres = new_xyz;
  p.nbr = idx;
  p.wpack = wpack;
  p.bias = bias;
  p.out = out;
  p.N = N;
  p.M = M;
  p.ns = ns;
  p.scale = normalize_xyz ? 1.0f / max_radius : 1.0f;
  p.tiles = B * (M / (kRows / ns));
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(sa_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  const int grid = (int)(p.tiles < kNumSMs ? p.tiles : kNumSMs);
  sa_pipe_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(p);
  return after_launch("sa_pipe_kernel");
}

}  // extern "C"
