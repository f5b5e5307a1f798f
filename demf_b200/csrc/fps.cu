// Furthest point sampling for sm_100a: one thread-block CLUSTER per scene.
//
// Upstream (mmdet3d furthest_point_sampling_kernel) runs ONE block per scene: m-1 dependent
// iterations, each a strided pass over all N points through global memory plus a 10-level
// shared-memory tree with a __syncthreads per level. At batch 8 that is 8 of 148 SMs busy
// for ~2047 x (20k-point pass) at SA1.
//
// Here a scene is owned by a cluster of C CTAs (C = 1..16, chosen so that all scenes are
// co-resident). Every point lives in REGISTERS for the whole kernel (x, y, z, running min
// distance, tie-break priority): an iteration is
//   1. kPPT fused distance updates per thread, no memory traffic at all;
//   2. warp arg-max with two REDUX.MAX (distance bits, then priority among equals);
//   3. one __syncthreads; every warp re-reduces the per-warp candidates (no second barrier);
//   4. C packets of 20 bytes pushed into every CTA of the cluster through distributed shared
//      memory (st.async + mbarrier complete_tx), so each CTA learns the winner and ITS
//      COORDINATES without touching global memory; two mbarriers alternate by iteration parity
//      and are re-armed two iterations ahead, so no cluster-wide barrier sits in the loop.
//
// Result parity: distances use the exact rounding of upstream (common.cuh: sqdist) and the
// arg-max key is (distance bits, priority) where priority encodes upstream's reduction order:
// thread t = k mod T scans k = t, t+T, .. and keeps the FIRST strict maximum; the block tree
// keeps the lower slot on ties, which orders threads by the bit-reversal of t. The winner is
// therefore max distance, then min bitrev(k mod T), then min k div T -- exactly what the oracle
// emulates (oracle/demf_oracle.c: demf_ref_fps). T = min(1024, 2^floor(log2 N)).
//
// A single-CTA global-memory kernel (fps_generic_kernel) covers N too large for registers.
#include <cooperative_groups.h>

#include <cstdlib>

#include "ball_grid.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace demf {
namespace {

constexpr int kMaxCluster = 16;

// ---- PTX helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_LOOP;\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// 16-byte and 4-byte remote stores that complete `bytes` on the destination CTA's mbarrier.
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
          "r"(remote_addr),
      "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t a, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                   remote_addr),
               "r"(a), "r"(remote_bar)
               : "memory");
}

// priority of point k under upstream's reduction order; larger inv-priority wins ties.
__device__ __forceinline__ uint32_t inv_priority(int k, int log2T) {
  const uint32_t t = (uint32_t)k & ((1u << log2T) - 1u);
  const uint32_t br = log2T ? (__brev(t) >> (32 - log2T)) : 0u;
  const uint32_t prio = (br << 20) | ((uint32_t)k >> log2T);  // < 2^30
  return 0xffffffffu - prio;                                  // >= 0xC0000000 > 0
}
__device__ __forceinline__ int index_from_inv_priority(uint32_t inv, int log2T) {
  const uint32_t prio = 0xffffffffu - inv;
  const uint32_t br = prio >> 20;
  const uint32_t t = log2T ? (__brev(br) >> (32 - log2T)) : 0u;
  return (int)(((prio & 0xfffffu) << log2T) | t);
}

struct Packet {  // what a CTA tells the cluster about its best point (20 bytes used)
  uint32_t dist_bits, inv_prio;
  float x, y;
  float z;
  uint32_t pad[3];
};

template <int kThreads>
struct FpsSmem {
  static constexpr int kWarps = kThreads / 32;
  alignas(16) Packet inbox[2][kMaxCluster];  // [iteration parity][source CTA]
  alignas(16) float4 warp_xyz[2][kWarps];
  uint2 warp_key[2][kWarps];
  alignas(8) uint64_t bar[2];
};

// -------------------------------------------------------------------------------------
template <int kThreads, int kPPT>
__global__ void __launch_bounds__(kThreads, 1) fps_cluster_kernel(const float* __restrict__ xyz,
                                                                  int N, int m, int log2T,
                                                                  int32_t* __restrict__ idx,
                                                                  float* __restrict__ new_xyz,
                                                                  const int32_t* __restrict__ unique_prefix) {
  constexpr int kWarps = kThreads / 32;
  __shared__ FpsSmem<kThreads> sm;
  const unsigned long long trace_t0 = trace_begin();

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned lane = lane_id();
  const int warp = tid >> 5;
  const float* cloud = xyz + (long)b * N * 3;
  int32_t* out = idx + (long)b * m;

  // ---- sampling a cloud that IS the pick sequence of an earlier furthest point sampling ----------------
  // If xyz[b] = (p_0, p_1, ...) are the picks, in order, of FPS on a superset, and each of the first m picks
  // was the UNIQUE arg-max of its iteration (no second point at the same running distance: certified by the
  // kernel that produced them, see fps_grid_kernel), then p_k is also the unique arg-max of iteration k of
  // FPS on this cloud -- the running distances are the same fp32 values, taken over a subset that contains
  // the maximum -- whatever the tie rule. The result is idx = 0..m-1 without a single iteration.
  if (unique_prefix != nullptr && __ldg(unique_prefix + b) >= m && m <= N) {
    if (rank == 0) {
      for (int i = tid; i < m; i += kThreads) out[i] = i;
      if (new_xyz) {
        float* o = new_xyz + (long)b * m * 3;
        for (int i = tid; i < m * 3; i += kThreads) o[i] = __ldg(cloud + i);
      }
    }
    trace_end(1, trace_t0);
    return;   // every CTA of the cluster takes this branch: nobody waits for anybody
  }

  // ---- load this thread's points into registers (coalesced: consecutive tid = consecutive k)
  float px[kPPT], py[kPPT], pz[kPPT], pd[kPPT];
  uint32_t pp[kPPT];
#pragma unroll
  for (int j = 0; j < kPPT; ++j) {
    const int k = (j * (int)C + (int)rank) * kThreads + tid;
    if (k < N) {
      px[j] = __ldg(cloud + (long)k * 3 + 0);
      py[j] = __ldg(cloud + (long)k * 3 + 1);
      pz[j] = __ldg(cloud + (long)k * 3 + 2);
      pd[j] = 1e10f;
      pp[j] = inv_priority(k, log2T);
    } else {  // padding slot: distance pinned to 0 and priority 0 -> loses to every real point
      px[j] = py[j] = pz[j] = 0.f;
      pd[j] = 0.f;
      pp[j] = 0u;
    }
  }
  float ox = __ldg(cloud + 0), oy = __ldg(cloud + 1), oz = __ldg(cloud + 2);  // idx[0] = 0
  float* oxyz = new_xyz ? new_xyz + (long)b * m * 3 : nullptr;  // optional: coordinates of the picks
  if (rank == 0 && tid == 0) {
    out[0] = 0;
    if (oxyz) {
      oxyz[0] = ox;
      oxyz[1] = oy;
      oxyz[2] = oz;
    }
  }

  const unsigned tx_bytes = 20u * C;
  if (C > 1) {
    if (tid == 0) {
      mbar_init(&sm.bar[0], 1);
      mbar_init(&sm.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // arm both parities before anyone can send
      mbar_arrive_expect_tx(&sm.bar[0], tx_bytes);
      mbar_arrive_expect_tx(&sm.bar[1], tx_bytes);
    }
    cluster.sync();
  }

  for (int it = 1; it < m; ++it) {
    const int par = it & 1;
    // 1. distance update + thread-local arg-max
    uint32_t bd = 0u, bp = 0u;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int j = 0; j < kPPT; ++j) {
      const float d = sqdist(px[j], py[j], pz[j], ox, oy, oz);
      const float t = fminf(d, pd[j]);
      pd[j] = t;
      const uint32_t tb = __float_as_uint(t);
      const bool better = (tb > bd) || (tb == bd && pp[j] > bp);
      bd = better ? tb : bd;
      bp = better ? pp[j] : bp;
      bx = better ? px[j] : bx;
      by = better ? py[j] : by;
      bz = better ? pz[j] : bz;
    }
    // 2. warp arg-max
    const uint32_t wd = __reduce_max_sync(0xffffffffu, bd);
    const uint32_t wp = __reduce_max_sync(0xffffffffu, bd == wd ? bp : 0u);
    if (bd == wd && bp == wp) {  // unique unless the whole warp is padding (then any lane will do)
      sm.warp_xyz[par][warp] = make_float4(bx, by, bz, 0.f);
      sm.warp_key[par][warp] = make_uint2(wd, wp);
    }
    __syncthreads();
    // 3. every warp reduces the kWarps candidates
    uint2 key = (lane < (unsigned)kWarps) ? sm.warp_key[par][lane] : make_uint2(0u, 0u);
    const uint32_t cd = __reduce_max_sync(0xffffffffu, key.x);
    const uint32_t cp = __reduce_max_sync(0xffffffffu, key.x == cd ? key.y : 0u);
    const unsigned who = __ffs(__ballot_sync(0xffffffffu, lane < (unsigned)kWarps && key.x == cd &&
                                                              key.y == cp)) - 1;
    const float4 cxyz = sm.warp_xyz[par][who];

    uint32_t win_p;
    if (C == 1) {
      ox = cxyz.x;
      oy = cxyz.y;
      oz = cxyz.z;
      win_p = cp;
    } else {
      // 4. tell every CTA of the cluster (including this one)
      if (warp == 0 && lane < C) {
        const uint32_t dst = map_to_cta(smem_u32(&sm.inbox[par][rank]), lane);
        const uint32_t dbar = map_to_cta(smem_u32(&sm.bar[par]), lane);
        st_async_v4(dst, cd, cp, __float_as_uint(cxyz.x), __float_as_uint(cxyz.y), dbar);
        st_async_b32(dst + 16, __float_as_uint(cxyz.z), dbar);
      }
      mbar_wait(&sm.bar[par], (unsigned)(((it - 1) >> 1) & 1));  // k-th use of this barrier
      // Re-arm this parity for iteration it+2. No packet of it+2 can be in flight yet: its
      // senders first need this CTA's packet of it+1, which is sent after this point.
      if (tid == 0) mbar_arrive_expect_tx(&sm.bar[par], tx_bytes);
      Packet pk;
      if (lane < C) {
        const uint4 q = *reinterpret_cast<const uint4*>(&sm.inbox[par][lane]);
        pk.dist_bits = q.x;
        pk.inv_prio = q.y;
      } else {
        pk.dist_bits = 0u;
        pk.inv_prio = 0u;
      }
      const uint32_t gd = __reduce_max_sync(0xffffffffu, pk.dist_bits);
      const uint32_t gp = __reduce_max_sync(0xffffffffu, pk.dist_bits == gd ? pk.inv_prio : 0u);
      const unsigned src = __ffs(__ballot_sync(0xffffffffu, lane < C && pk.dist_bits == gd &&
                                                                pk.inv_prio == gp)) - 1;
      const Packet* w = &sm.inbox[par][src];
      ox = w->x;
      oy = w->y;
      oz = w->z;
      win_p = gp;
    }
    if (rank == 0 && tid == 0) {
      out[it] = index_from_inv_priority(win_p, log2T);
      if (oxyz) {
        oxyz[it * 3 + 0] = ox;
        oxyz[it * 3 + 1] = oy;
        oxyz[it * 3 + 2] = oz;
      }
    }
  }
  if (C > 1) cluster.sync();  // nobody may exit while peers can still write into its inbox
  trace_end(1, trace_t0);
}

// ---- grid-pruned FPS: the cloud lives in the SHARED memory of a small cluster ----------------
// The register-resident kernel above updates every point in every iteration and needs 16 CTAs per
// 20k-point scene; its cost is the per-iteration latency chain times the 128 SMs it holds. Here the
// scene is taken in the cell order of the ball-query grid (ball_grid.cuh: points sorted by uniform
// grid cell), cut into C slabs (one per CTA of a cluster of 2..8) that sit in shared memory
// (x, y, z, tie-break priority, running min distance), and further into blocks of 32 consecutive
// points. Each block keeps, in the registers of one lane, its bounding box, the largest running
// distance of its points and its best (distance, priority) key. A new sample s can only lower a
// running distance of block b if  lb2(s, box_b) < max_b,  where lb2 is the SAME fma-ordered
// expression as the point distance evaluated on the box: every rounding step is monotone in the
// coordinate differences, so lb2 <= d(p, s) for every p in the box and a skipped block provably
// changes nothing. After a few dozen samples only ~10 blocks of 625 are touched per iteration.
// The arg-max key and the cluster exchange are those of fps_cluster_kernel: identical indices.
#ifdef DEMF_FPS_STATS
__device__ unsigned long long g_fps_stats[4];
#endif
constexpr int kGridFpsThreads = 512;
constexpr int kGridFpsWarps = kGridFpsThreads / 32;

__global__ void __launch_bounds__(kGridFpsThreads, 1) fps_grid_kernel(const float* __restrict__ xyz,
                                                                      const void* __restrict__ grid, int N, int m,
                                                                      int log2T, int cap, int32_t* __restrict__ idx,
                                                                      float* __restrict__ new_xyz,
                                                                      int32_t* __restrict__ unique_prefix, int cert) {
  extern __shared__ __align__(16) unsigned char gsm[];
  __shared__ FpsSmem<kGridFpsThreads> sm;
  __shared__ __align__(16) Packet flat_inbox[2][32];  // [iteration parity][source CTA * 16 + warp]
  const unsigned long long trace_t0 = trace_begin();
  float4* pts = reinterpret_cast<float4*>(gsm);                 // cap entries: x, y, z, inv-priority bits
  float* tmp = reinterpret_cast<float*>(gsm + (size_t)cap * 16);  // cap running min distances

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned lane = lane_id();
  const int warp = tid >> 5;
  int32_t* out = idx + (long)b * m;
  const BallGridView g = ball_grid_view(grid, b, N);

  // ---- this CTA's slab of the cell-ordered cloud
  const int begin = (int)rank * cap;
  for (int i = tid; i < cap; i += kGridFpsThreads) {
    const int q = begin + i;
    if (q < N) {
      const float4 p = __ldg(g.sorted + q);
      pts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(inv_priority(__float_as_int(p.w), log2T)));
      tmp[i] = 1e10f;
    } else {  // padding: distance pinned to 0, priority 0 -> never wins, never changes
      pts[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
      tmp[i] = 0.f;
    }
  }
  __syncthreads();
  // ---- block metadata in registers: lane j of warp w owns block w + 16*j
  const int nblk = cap >> 5;
  const int my_blk = warp + kGridFpsWarps * (int)lane;
  const bool has_blk = my_blk < nblk && begin + my_blk * 32 < N;
  float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;
  uint32_t bmax = 0u, bkd = 0u, bkp = 0u;
  int bli = 0;
  // Uniqueness certificate for the sampling chain (see fps_cluster_kernel): `btie` = this block's best running
  // distance is attained by more than one of its points; first_tie = first iteration whose global arg-max was
  // not provably unique (m if none). Conservative: any doubt counts as a tie.
  bool btie = true;     // all distances start equal (1e10)
  int first_tie = m;
  if (has_blk) {
    lox = loy = loz = 3.0e38f;
    hix = hiy = hiz = -3.0e38f;
    for (int i = 0; i < 32; ++i) {
      const int q = my_blk * 32 + i;
      if (begin + q < N) {
        const float4 p = pts[q];
        lox = fminf(lox, p.x); hix = fmaxf(hix, p.x);
        loy = fminf(loy, p.y); hiy = fmaxf(hiy, p.y);
        loz = fminf(loz, p.z); hiz = fmaxf(hiz, p.z);
      }
    }
    bmax = __float_as_uint(1e10f);
  }
  float ox = __ldg(xyz + (long)b * N * 3 + 0), oy = __ldg(xyz + (long)b * N * 3 + 1),
        oz = __ldg(xyz + (long)b * N * 3 + 2);  // idx[0] = 0
  float* oxyz = new_xyz ? new_xyz + (long)b * m * 3 : nullptr;  // optional: coordinates of the picks
  if (rank == 0 && tid == 0) {
    out[0] = 0;
    if (oxyz) {
      oxyz[0] = ox;
      oxyz[1] = oy;
      oxyz[2] = oz;
    }
  }

  const bool flat = C * kGridFpsWarps <= 32;
  const unsigned tx_bytes = flat ? 20u * C * kGridFpsWarps : 20u * C;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(&sm.bar[0], tx_bytes);
    mbar_arrive_expect_tx(&sm.bar[1], tx_bytes);
  }
  cluster.sync();

  for (int it = 1; it < m; ++it) {
    const int par = it & 1;
    // a. which of this warp's blocks can the new sample change?
    bool dirty = false;
    if (has_blk) {
      const float dx = fmaxf(0.f, fmaxf(__fsub_rn(lox, ox), __fsub_rn(ox, hix)));
      const float dy = fmaxf(0.f, fmaxf(__fsub_rn(loy, oy), __fsub_rn(oy, hiy)));
      const float dz = fmaxf(0.f, fmaxf(__fsub_rn(loz, oz), __fsub_rn(oz, hiz)));
      const float lb2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      dirty = lb2 < __uint_as_float(bmax);
    }
    unsigned todo = __ballot_sync(0xffffffffu, dirty);
#ifdef DEMF_FPS_STATS
    if (lane == 0) {
      atomicAdd(&g_fps_stats[0], (unsigned long long)__popc(todo));
      atomicMax(&g_fps_stats[1], (unsigned long long)__popc(todo));
      if (it >= 64) atomicAdd(&g_fps_stats[2], (unsigned long long)__popc(todo));
    }
#endif
    // b. update those blocks, 32 points at a time
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const int q = (warp + kGridFpsWarps * j) * 32 + (int)lane;
      const float4 p = pts[q];
      const float d = sqdist(p.x, p.y, p.z, ox, oy, oz);
      const float t = fminf(d, tmp[q]);
      tmp[q] = t;
      const uint32_t tb = __float_as_uint(t), pr = __float_as_uint(p.w);
      const uint32_t wd = __reduce_max_sync(0xffffffffu, tb);
      const uint32_t wp = __reduce_max_sync(0xffffffffu, tb == wd ? pr : 0u);
      const unsigned at = __ballot_sync(0xffffffffu, tb == wd);
      const bool multi = (at & (at - 1u)) != 0u;            // more than one point at the block's best distance
      const int who = (multi ? __ffs(__ballot_sync(0xffffffffu, tb == wd && pr == wp)) : __ffs(at)) - 1;
      if ((int)lane == j) {
        bmax = wd;
        bkd = wd;
        bkp = wp;
        bli = who;
        btie = multi;
      }
    }
    // c. this warp's candidate = best key over its blocks
    const uint32_t kd = has_blk ? bkd : 0u, kp = has_blk ? bkp : 0u;
    const uint32_t wd = __reduce_max_sync(0xffffffffu, kd);
    const uint32_t wp = __reduce_max_sync(0xffffffffu, kd == wd ? kp : 0u);
    // blocks of this warp at its best distance; the tie flag of each rides along in bit 0 of a second mask
    const unsigned same = __ballot_sync(0xffffffffu, has_blk && kd == wd);
    const bool several = (same & (same - 1u)) != 0u;
    const unsigned owners = several ? __ballot_sync(0xffffffffu, has_blk && kd == wd && kp == wp) : same;
    const int src = owners ? __ffs(owners) - 1 : 0;
    // tie inside this warp: two of its blocks at the best distance, or several points inside the best block
    // (only the first `cert` iterations are certified: that is all the later sampling levels ask for)
    const bool track = it < cert;
    bool wtie = false;
    if (track) wtie = several || __ballot_sync(0xffffffffu, has_blk && kd == wd && btie) != 0u || wd == 0u;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int)lane == src && owners) c = pts[my_blk * 32 + bli];
    c.x = __shfl_sync(0xffffffffu, c.x, src);
    c.y = __shfl_sync(0xffffffffu, c.y, src);
    c.z = __shfl_sync(0xffffffffu, c.z, src);
    uint32_t gp;
    if (flat) {
      // d'. flat exchange (C * 16 warps <= 32): every warp pushes its candidate straight into the
      // inbox of every CTA of the cluster (its own included) and every warp reduces the <= 32
      // candidates itself: no CTA barrier and no second exchange stage in the loop
      if (lane < C) {
        const uint32_t dst = map_to_cta(smem_u32(&flat_inbox[par][rank * kGridFpsWarps + warp]), lane);
        const uint32_t dbar = map_to_cta(smem_u32(&sm.bar[par]), lane);
        // the tie flag rides in bit 30 of the priority word, which is 1 in every real inverse priority
        // (inv_priority >= 0xC0000000): cleared = "this warp's best distance is not unique"; the receiver restores it
        st_async_v4(dst, owners ? wd : 0u, owners ? (wtie ? (wp & ~0x40000000u) : wp) : 0u, __float_as_uint(c.x),
                    __float_as_uint(c.y), dbar);
        st_async_b32(dst + 16, __float_as_uint(c.z), dbar);
      }
      mbar_wait(&sm.bar[par], (unsigned)(((it - 1) >> 1) & 1));
      if (tid == 0) mbar_arrive_expect_tx(&sm.bar[par], tx_bytes);
      uint32_t pd = 0u, pp = 0u, ptie = 0u;
      if (lane < C * kGridFpsWarps) {
        const uint2 q2 = *reinterpret_cast<const uint2*>(&flat_inbox[par][lane]);
        pd = q2.x;
        pp = q2.y;
        ptie = (pp != 0u && (pp & 0x40000000u) == 0u) ? 1u : 0u;
        pp = pp != 0u ? (pp | 0x40000000u) : 0u;
      }
      const uint32_t gd = __reduce_max_sync(0xffffffffu, pd);
      gp = __reduce_max_sync(0xffffffffu, pd == gd ? pp : 0u);
      if (track) {
        const unsigned at_max = __ballot_sync(0xffffffffu, lane < C * kGridFpsWarps && pd == gd);
        const bool gtie = __popc(at_max) > 1 || __ballot_sync(0xffffffffu, (at_max >> lane) & 1u && ptie != 0u) != 0u ||
                          gd == 0u;
        if (gtie && first_tie == m) first_tie = it;
      }
      const unsigned srcc = __ffs(__ballot_sync(0xffffffffu, lane < C * kGridFpsWarps && pd == gd && pp == gp)) - 1;
      const Packet* w = &flat_inbox[par][srcc];
      ox = w->x;
      oy = w->y;
      oz = w->z;
    } else {
      if (track && first_tie == m) first_tie = it;   // the two-stage exchange carries no uniqueness information
      if (lane == 0) {
        sm.warp_xyz[par][warp] = make_float4(c.x, c.y, c.z, 0.f);
        sm.warp_key[par][warp] = make_uint2(owners ? wd : 0u, owners ? wp : 0u);
      }
      __syncthreads();
      // d. CTA arg-max over the warps, then the cluster exchange (as in fps_cluster_kernel)
      uint2 key = (lane < (unsigned)kGridFpsWarps) ? sm.warp_key[par][lane] : make_uint2(0u, 0u);
      const uint32_t cd = __reduce_max_sync(0xffffffffu, key.x);
      const uint32_t cp = __reduce_max_sync(0xffffffffu, key.x == cd ? key.y : 0u);
      const unsigned who = __ffs(__ballot_sync(0xffffffffu, lane < (unsigned)kGridFpsWarps && key.x == cd &&
                                                                key.y == cp)) - 1;
      const float4 cxyz = sm.warp_xyz[par][who];
      if (warp == 0 && lane < C) {
        const uint32_t dst = map_to_cta(smem_u32(&sm.inbox[par][rank]), lane);
        const uint32_t dbar = map_to_cta(smem_u32(&sm.bar[par]), lane);
        st_async_v4(dst, cd, cp, __float_as_uint(cxyz.x), __float_as_uint(cxyz.y), dbar);
        st_async_b32(dst + 16, __float_as_uint(cxyz.z), dbar);
      }
      mbar_wait(&sm.bar[par], (unsigned)(((it - 1) >> 1) & 1));
      if (tid == 0) mbar_arrive_expect_tx(&sm.bar[par], tx_bytes);
      Packet pk;
      if (lane < C) {
        const uint4 qd = *reinterpret_cast<const uint4*>(&sm.inbox[par][lane]);
        pk.dist_bits = qd.x;
        pk.inv_prio = qd.y;
      } else {
        pk.dist_bits = 0u;
        pk.inv_prio = 0u;
      }
      const uint32_t gd = __reduce_max_sync(0xffffffffu, pk.dist_bits);
      gp = __reduce_max_sync(0xffffffffu, pk.dist_bits == gd ? pk.inv_prio : 0u);
      const unsigned srcc = __ffs(__ballot_sync(0xffffffffu, lane < C && pk.dist_bits == gd && pk.inv_prio == gp)) - 1;
      const Packet* w = &sm.inbox[par][srcc];
      ox = w->x;
      oy = w->y;
      oz = w->z;
    }
    if (rank == 0 && tid == 0) {
      out[it] = index_from_inv_priority(gp, log2T);
      if (oxyz) {
        oxyz[it * 3 + 0] = ox;
        oxyz[it * 3 + 1] = oy;
        oxyz[it * 3 + 2] = oz;
      }
    }
  }
  if (unique_prefix != nullptr && rank == 0 && tid == 0) unique_prefix[b] = min(first_tie, max(cert, 1));
  cluster.sync();  // nobody may exit while peers can still write into its inbox
  trace_end(1, trace_t0);
}

// ---- fallback: one CTA per scene, points and running distances in global memory ----------
// (N beyond what the register-resident kernel covers). Same key, same result.
__global__ void __launch_bounds__(1024, 1) fps_generic_kernel(const float* __restrict__ xyz, int N,
                                                              int m, int log2T,
                                                              float* __restrict__ temp,
                                                              int32_t* __restrict__ idx,
                                                              float* __restrict__ new_xyz) {
  __shared__ uint2 warp_key[2][32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const unsigned lane = lane_id();
  const int warp = tid >> 5;
  const float* cloud = xyz + (long)b * N * 3;
  float* tmp = temp + (long)b * N;
  int32_t* out = idx + (long)b * m;
  for (int k = tid; k < N; k += 1024) tmp[k] = 1e10f;
  int old = 0;
  float* oxyz = new_xyz ? new_xyz + (long)b * m * 3 : nullptr;
  if (tid == 0) {
    out[0] = 0;
    if (oxyz) {
      oxyz[0] = __ldg(cloud + 0);
      oxyz[1] = __ldg(cloud + 1);
      oxyz[2] = __ldg(cloud + 2);
    }
  }
  for (int it = 1; it < m; ++it) {
    const int par = it & 1;
    const float ox = __ldg(cloud + (long)old * 3), oy = __ldg(cloud + (long)old * 3 + 1),
                oz = __ldg(cloud + (long)old * 3 + 2);
    uint32_t bd = 0u, bp = 0u;
    for (int k = tid; k < N; k += 1024) {
      const float d = sqdist(__ldg(cloud + (long)k * 3), __ldg(cloud + (long)k * 3 + 1),
                             __ldg(cloud + (long)k * 3 + 2), ox, oy, oz);
      const float t = fminf(d, tmp[k]);
      tmp[k] = t;
      const uint32_t tb = __float_as_uint(t), p = inv_priority(k, log2T);
      const bool better = (tb > bd) || (tb == bd && p > bp);
      bd = better ? tb : bd;
      bp = better ? p : bp;
    }
    const uint32_t wd = __reduce_max_sync(0xffffffffu, bd);
    const uint32_t wp = __reduce_max_sync(0xffffffffu, bd == wd ? bp : 0u);
    if (lane == 0) warp_key[par][warp] = make_uint2(wd, wp);
    __syncthreads();
    const uint2 key = warp_key[par][lane];
    const uint32_t cd = __reduce_max_sync(0xffffffffu, key.x);
    const uint32_t cp = __reduce_max_sync(0xffffffffu, key.x == cd ? key.y : 0u);
    old = index_from_inv_priority(cp, log2T);
    if (tid == 0) {
      out[it] = old;
      if (oxyz) {
        oxyz[it * 3 + 0] = __ldg(cloud + (long)old * 3);
        oxyz[it * 3 + 1] = __ldg(cloud + (long)old * 3 + 1);
        oxyz[it * 3 + 2] = __ldg(cloud + (long)old * 3 + 2);
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
int floor_log2(int n) {
  int l = 0;
  while ((2 << l) <= n) ++l;
  return l;
}

constexpr int kPPTMenu[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 40};
constexpr int kMaxPPT = 40;

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// set by demf_fps_prefix for the launch it is about to make (host-side call state of this thread)
thread_local const int32_t* t_unique_prefix = nullptr;

template <int kThreads, int kPPT>
int launch_cluster(const float* xyz, int B, int N, int m, int C, int log2T, int32_t* idx, float* new_xyz,
                   cudaStream_t st, bool probe_only, int* max_clusters) {
  auto kernel = fps_cluster_kernel<kThreads, kPPT>;
  if (C > 8) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return -1;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (probe_only) {
    static int cached[kMaxCluster + 1] = {};  // per instantiation: 0 = unknown, else result + 1
    if (cached[C] == 0) {
      int n = 0;
      cfg.gridDim = dim3(C, 1, 1);  // the answer does not depend on the batch
      const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
      }
      cached[C] = n + 1;
    }
    *max_clusters = cached[C] - 1;
    return 0;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, xyz, N, m, log2T, idx, new_xyz, t_unique_prefix);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("fps_cluster_kernel<%d,%d> C=%d: %s", kThreads, kPPT, C, cudaGetErrorString(e));
    return (int)e;
  }
  return after_launch("fps_cluster_kernel");
}

template <int kThreads>
int dispatch_ppt(int ppt, const float* xyz, int B, int N, int m, int C, int log2T, int32_t* idx, float* new_xyz,
                 cudaStream_t st, bool probe_only, int* max_clusters) {
  switch (ppt) {
#define DEMF_CASE(p) \
  case p:            \
    return launch_cluster<kThreads, p>(xyz, B, N, m, C, log2T, idx, new_xyz, st, probe_only, max_clusters);
    DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(3) DEMF_CASE(4) DEMF_CASE(5) DEMF_CASE(6) DEMF_CASE(8)
    DEMF_CASE(10) DEMF_CASE(12) DEMF_CASE(16) DEMF_CASE(20) DEMF_CASE(24) DEMF_CASE(32) DEMF_CASE(40)
#undef DEMF_CASE
  }
  return -1;
}

int dispatch(int threads, int ppt, const float* xyz, int B, int N, int m, int C, int log2T,
             int32_t* idx, float* new_xyz, cudaStream_t st, bool probe_only, int* max_clusters) {
  switch (threads) {
    case 128:
      return dispatch_ppt<128>(ppt, xyz, B, N, m, C, log2T, idx, new_xyz, st, probe_only, max_clusters);
    case 256:
      return dispatch_ppt<256>(ppt, xyz, B, N, m, C, log2T, idx, new_xyz, st, probe_only, max_clusters);
    case 512:
      return dispatch_ppt<512>(ppt, xyz, B, N, m, C, log2T, idx, new_xyz, st, probe_only, max_clusters);
  }
  return -1;
}

int round_up_ppt(int need) {
  for (int p : kPPTMenu)
    if (p >= need) return p;
  return -1;
}

// Picks (threads, cluster size, points per thread). Returns false if the register-resident
// kernel cannot cover N at this batch size.
//   - a cluster exchange costs about as much as ~25 distance updates per thread, so clouds
//     that fit one CTA at <= kSoloPPT points per thread stay on one CTA (no exchange at all);
//   - larger clouds take the smallest cluster that brings the per-thread work down to
//     `want_ppt`, bounded by what keeps every scene's cluster resident at once.
constexpr int kSoloPPT = 24;

bool plan(int B, int N, int* threads, int* C, int* ppt) {
  const int T = env_int("DEMF_FPS_THREADS", 256);
  if (T != 128 && T != 256 && T != 512) return false;
  const int want_ppt = env_int("DEMF_FPS_PPT", 5);
  const int forced = env_int("DEMF_FPS_CLUSTER", 0);
  int cmax = kMaxCluster;
  while (cmax > 1 && (long)cmax * B > kNumSMs) cmax >>= 1;  // all scenes co-resident
  int c = 1;
  if (forced > 0) {
    c = forced < cmax ? forced : cmax;
  } else if ((long)T * kSoloPPT < N) {
    while (c < cmax && (long)c * T * want_ppt < N) c <<= 1;
  }
  const int p = round_up_ppt((N + c * T - 1) / (c * T));
  if (p < 0) return false;
  if (T == 512 && p > 16) return false;  // 512 threads x 1 CTA/SM: 128 registers per thread
  *threads = T;
  *C = c;
  *ppt = p;
  return true;
}

}  // namespace
DEMF_DEFINE_TRACE_SETTER(trace_set_fps)
}  // namespace demf

using namespace demf;

extern "C" {

size_t demf_fps_workspace_bytes(int B, int N, int m) {
  (void)m;
  int t, c, p;
  if (B <= 0 || N <= 0) return 0;
  if (plan(B, N, &t, &c, &p)) return 0;
  return (size_t)B * N * sizeof(float);
}

int demf_fps_prefix(const float* xyz, int B, int N, int m, void* workspace, int32_t* idx, float* new_xyz,
                    const int32_t* unique_prefix, void* stream) {
  t_unique_prefix = unique_prefix;
  const int rc = demf_fps(xyz, B, N, m, workspace, idx, new_xyz, stream);
  t_unique_prefix = nullptr;
  return rc;
}

int demf_fps(const float* xyz, int B, int N, int m, void* workspace, int32_t* idx, float* new_xyz,
             void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && m >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535 && N < (1 << 30), DEMF_E_SIZE);
  if (B == 0 || m == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int log2T = floor_log2(N);
  if (log2T > 10) log2T = 10;  // upstream block: min(1024, 2^floor(log2 N)) threads
  int threads, C, ppt;
  if (plan(B, N, &threads, &C, &ppt)) {
    // shrink the cluster until every scene's cluster is co-resident on this device
    while (C > 1) {
      int fit = 0;
      if (dispatch(threads, ppt, xyz, B, N, m, C, log2T, idx, new_xyz, st, true, &fit) == 0 && fit >= B) break;
      C >>= 1;
      const int need = (N + C * threads - 1) / (C * threads);
      ppt = round_up_ppt(need);
      if (ppt < 0) break;
    }
    if (ppt > 0 && !(threads == 512 && ppt > 16)) {
      const int rc = dispatch(threads, ppt, xyz, B, N, m, C, log2T, idx, new_xyz, st, false, nullptr);
      if (rc != -1) return rc;
    }
  }
  if (workspace == nullptr) {
    set_error("demf_fps: N=%d at B=%d needs the global-memory kernel; pass demf_fps_workspace_bytes() "
              "bytes of workspace", N, B);
    return DEMF_E_WORKSPACE;
  }
  fps_generic_kernel<<<B, 1024, 0, st>>>(xyz, N, m, log2T, static_cast<float*>(workspace), idx, new_xyz);
  return after_launch("fps_generic_kernel");
}

#ifdef DEMF_FPS_STATS
unsigned long long demf_fps_stat(int i) {
  unsigned long long v[4];
  cudaMemcpyFromSymbol(v, g_fps_stats, sizeof(v));
  return v[i];
}
#endif

/* Grid-pruned FPS: same indices as demf_fps; `grid` = demf_ball_grid_build workspace of the SAME xyz
 * (any radius). The cloud sits in the shared memory of a 2-, 4- or 8-CTA cluster per scene. */
int demf_fps_grid(const float* xyz, const void* grid, int B, int N, int m, int32_t* idx, float* new_xyz,
                  void* stream) {
  return demf_fps_grid_prefix(xyz, grid, B, N, m, idx, new_xyz, nullptr, 0, stream);
}

int demf_fps_grid_prefix(const float* xyz, const void* grid, int B, int N, int m, int32_t* idx, float* new_xyz,
                         int32_t* unique_prefix, int certify, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(grid);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && m >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535 && N < (1 << 30), DEMF_E_SIZE);
  if (B == 0 || m == 0) return 0;
  int log2T = floor_log2(N);
  if (log2T > 10) log2T = 10;
  // smallest cluster whose slab (20 bytes per point) fits one SM's shared memory: the fewer CTAs,
  // the cheaper the per-iteration exchange and the fewer SMs a scene holds (measured at N=20000,
  // 8 forwards in flight: C=2 1.06 ms/step, C=4 1.09, C=8 1.29). A lane owns one block of 32 points:
  // at most 16 warps x 32 lanes = 512 blocks per CTA.
  int C = env_int("DEMF_FPS_GRID_CLUSTER", 0);
  if (C != 2 && C != 4 && C != 8) {
    C = 2;
    while (C < 8 && ((((N + C - 1) / C) + 31) / 32 * 32) * 20 > 200 * 1024) C <<= 1;
  }
  int cap = (((N + C - 1) / C) + 31) / 32 * 32;
  DEMF_REQUIRE(cap / 32 <= kGridFpsWarps * 32 && (size_t)cap * 20 <= 200 * 1024, DEMF_E_UNSUPPORTED);
  const size_t smem = (size_t)cap * 20;
  static size_t configured = 0;
  if (smem > configured) {
    const cudaError_t e = cudaFuncSetAttribute(fps_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("demf_fps_grid: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(kGridFpsThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, fps_grid_kernel, xyz, grid, N, m, log2T, cap, idx, new_xyz, unique_prefix,
                                           unique_prefix ? (certify < m ? certify : m) : 0);
  if (e != cudaSuccess) {
    set_error("demf_fps_grid: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return after_launch("fps_grid_kernel");
}

}  // extern "C"
