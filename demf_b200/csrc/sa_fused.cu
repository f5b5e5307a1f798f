// One launch per set-abstraction level in inference: ball query -> grouping -> 3-layer shared MLP
// (BN folded, bias + ReLU) -> max over the neighbourhood. Nothing but the (B,M,C3) result leaves
// the SM: the grouped tensor (a5) and the three activation tensors (a6) of SURVEY.md section 8 --
// ~275 MB per scene upstream -- live in shared memory and TMEM only.
//
// Replaces, for `PointSAModule.forward` in eval mode (mmdet3d ops/pointnet_modules/point_sa_module.py,
// reached from configs/demf/demf_votenet.py:48-62,155-162): ball_query_kernel + 2x group_points_kernel
// + transpose/sub/div/cat + 3x (cuDNN 1x1 conv + BN + ReLU) + max_pool2d.
//
// CTA = G <= 16 centres of one scene (G*ns = 128*tiles grouped rows), 18 warps:
//   warp 0      : tcgen05.mma issuer (one lane)          D[tmem] += A[smem] * W[smem]^T, kind::tf32
//   warp 1      : weight producer (one lane)             1-D bulk copies of host-packed, pre-swizzled
//                                                         weight chunks from L2 into a ring (or once,
//                                                         when all three matrices fit: "resident")
//   warps 2..17 : workers                                ball query (uniform-grid neighbourhood, or the
//                                                         cloud staged through smem), neighbour-row
//                                                         gather into the swizzled A operand, TMEM
//                                                         epilogues
// The workers form 1, 2 or 4 LANES; a lane owns a 128-row tile at a time with its own operand
// region, TMEM columns and mbarrier pair, so that while one lane runs an epilogue or a gather the
// tensor pipe works on another lane's tile (the phases of one tile are strictly dependent and
// each costs a few hundred ns of latency). Per tile and lane:
//   gather A0 (128 x K0, in passes of `cpp` 32-column chunks when K0 is wide) -> MMA layer 0 ->
//   epilogue (bias, ReLU, tf32 round) writes act1 over A0 -> MMA layer 1 -> epilogue -> act2 ->
//   MMA layer 2 -> epilogue: column max over the ns rows of each centre (register butterfly across
//   the 32 TMEM lanes a warp owns), bias, ReLU, store.
//
// Arithmetic: TF32 products (operands rounded to nearest, ties away: cvt.rna), fp32 accumulation --
// the same class as the cuDNN/cuBLAS TF32 convolutions PyTorch >= 1.7 runs for the reference on
// Ampere and later. Indices are exact (the ball query is the fp32 fma-ordered one of ball_query.cu).
//
// Roofline: tensor pipe / L2 (weights re-streamed per tile unless resident). HBM bytes per scene =
// N*12 + N*C*4 + M*12 + M*C3*4 (+ weights once).
#include <cstdlib>
#include "ball_grid.cuh"
#include "common.cuh"
#include "umma.cuh"

namespace demf {
namespace {

using namespace umma;

constexpr int kWorkerWarps = 16;
constexpr int kWorkers = kWorkerWarps * 32;  // 512
constexpr int kThreads = kWorkers + 64;      // + issuer warp + producer warp
constexpr int kTileRows = 128;
constexpr int kChunkBytes = kTileRows * 128;  // one 32-float K chunk of a 128-row operand
constexpr int kCloudTile = 2048;              // cloud points per ball-query tile (24 KB)
constexpr int kMaxSlots = 8;
constexpr int kMaxLanes = 4;

__device__ int g_sa_error = 0;  // sticky: first protocol time-out (never expected)
long long* g_sa_prof = nullptr;
int g_sa_max_lanes = kMaxLanes;  // tuning knobs (demf_sa_fused_tune)
int g_sa_sleep_ns = 0;

struct SaParams {
  const float* xyz;
  const float* feat;
  const float* new_xyz;
  const float* wpack;
  const float* bias;
  float* out;
  int32_t* idx;
  const void* grid;  // ball_grid workspace of this cloud, or NULL: scan the whole cloud
  long long* prof;   // debug: clock64 stamps of the first worker thread of CTA (0,0), or NULL
  int N, M, C, ns, G, tiles, K0;
  int c[3];
  int query, normalize_xyz;
  float min_r2, max_r2, inv_radius;
  int lanes, cpp, npass, lane_act_bytes;  // tile pipelines; A0 chunks per layer-0 pass; passes
  int slots, slot_bytes, resident;
  int tmem_cols, lane_cols, acc_col[3];
  int sleep_ns;  // back-off between mbarrier polls of the worker warps (0 = spin)
  int t2;        // layer 2 runs transposed (D^T = W3 * act2^T): lanes = channels, columns = rows
};

struct SmemLayout {
  int ring, rows, centres, bias, part, bars, total;
};

__host__ __device__ inline SmemLayout smem_layout(const SaParams& p) {
  SmemLayout L;
  L.ring = p.lanes * p.lane_act_bytes;
  L.rows = L.ring + p.slots * p.slot_bytes;
  L.centres = L.rows + p.G * p.ns * 4;
  L.bias = L.centres + p.G * 16;
  L.part = L.bias + (p.c[0] + p.c[1] + p.c[2]) * 4;
  L.bars = (L.part + (p.ns == 64 ? p.lanes * 2 * p.c[2] * 4 : 0) + 15) & ~15;
  L.total = L.bars + (3 * kMaxSlots + 2 * kMaxLanes) * 8 + 16 + 1024;  // + alignment slack
  return L;
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Round-to-nearest (ties away from zero) to TF32 for an MMA OPERAND in one integer add: the tensor
// core reads the top 19 bits of the word, so adding half a TF32 ulp to the magnitude is all
// cvt.rna.tf32 does for finite values (the low 13 bits left behind are ignored by the hardware).
__device__ __forceinline__ float tf32_operand(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

__device__ __forceinline__ float4 tf32_operand4(float4 v) {
  return make_float4(tf32_operand(v.x), tf32_operand(v.y), tf32_operand(v.z), tf32_operand(v.w));
}

__device__ __forceinline__ void named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ bool worker_sync_and(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.and.pred p, 1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)pred), "n"(kWorkers)
      : "memory");
  return r != 0;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Bounded wait that also gives up as soon as any thread of the CTA has flagged a failure.
__device__ __forceinline__ bool wait_or_fail(uint32_t bar, uint32_t parity, volatile int* failed, int code,
                                             int sleep_ns = 0) {
  if (mbar_try_wait(bar, parity)) return true;
  unsigned long long t0 = 0;
  for (uint32_t spin = 1;; ++spin) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if (mbar_try_wait(bar, parity)) return true;
    if ((spin & 255u) == 0) {
      if (*failed) return false;
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 400ull * 1000 * 1000) {  // 0.4 s: a protocol bug, not a slow kernel
        *failed = code;
        atomicCAS(&g_sa_error, 0, code);
        return false;
      }
    }
  }
}

// 32 -> 1 (FULL) or 32 -> 2 (half warps) column maxima held across the lanes of a warp.
// After the call: FULL: v[0] = max over the 32 lanes of column `lane`;
//                 half: v[0], v[1] = max over the 16 lanes of this half of columns 2*(lane&15)+{0,1}.
template <bool FULL>
__device__ __forceinline__ void lane_max_transpose(float (&v)[32], unsigned lane) {
  if (FULL) {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = up ? v[i] : v[i + 16];
      const float keep = up ? v[i + 16] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
  }
  constexpr int n0 = FULL ? 16 : 32;
#pragma unroll
  for (int n = n0, o = 8; o >= 1; n >>= 1, o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}

// The (layer, first chunk, chunk count) of MMA group g of a tile: groups 0..npass-1 are the passes of
// layer 0, then layers 1 and 2.
__device__ __forceinline__ void group_span(const SaParams& p, int nch0, int g, int& layer, int& ch0, int& nch) {
  if (g < p.npass) {
    layer = 0;
    ch0 = g * p.cpp;
    nch = min(p.cpp, nch0 - ch0);
  } else {
    layer = g - p.npass + 1;
    ch0 = 0;
    nch = p.c[layer - 1] >> 5;
  }
}

struct IssueCtx {
  uint32_t tmem, act_u32, ring_u32, wfull0, wempty0, acc_full0;
  int nch0;
  volatile int* failed;
};

// Issue the tcgen05.mma instructions of MMA group g of lane `ln` (one thread) and commit the lane's
// acc_full barrier. Resident weights: chunk `cid` sits in ring slot `cid` for the whole kernel.
__device__ __forceinline__ bool issue_group(const SaParams& p, const IssueCtx& c, int ln, int g, uint32_t& slot,
                                            uint32_t& wph, bool check_w = true) {
  int layer, ch0, nch;
  group_span(p, c.nch0, g, layer, ch0, nch);
  const uint32_t idesc = instr_desc_tf32(kTileRows, p.c[layer]);
  const int cid = layer == 0 ? ch0 : (layer == 1 ? c.nch0 : c.nch0 + (p.c[0] >> 5));
  const uint32_t d = c.tmem + ln * p.lane_cols + p.acc_col[layer];
  const uint32_t a0 = c.act_u32 + ln * p.lane_act_bytes;
  tc_fence_after_sync();
  for (int ch = 0; ch < nch; ++ch) {
    const uint32_t s = p.resident ? (uint32_t)(cid + ch) : slot;
    // resident weights arrive once: after a lane's first tile their barriers need no second look
    if (check_w && !wait_or_fail(c.wfull0 + 8 * s, p.resident ? 0u : wph, c.failed, 2)) return false;
    const int ksteps = layer == 0 ? (min(32, p.K0 - 32 * (ch0 + ch)) >> 3) : 4;
    // a K step of 8 tf32 = +32 bytes = +2 in the descriptor's (address >> 4) field
    const uint64_t da = smem_desc_sw128(a0 + ch * kChunkBytes), dw = smem_desc_sw128(c.ring_u32 + s * p.slot_bytes);
    if (layer == 2 && p.t2) {
      // D^T[channel, row] += W3[channel, k] * act2[row, k]: the weight chunk is the M operand (128
      // channels per MMA; the second half of a 256-wide layer sits 128 rows = 16 KB further), the
      // activation tile the N operand. 128 accumulator columns per 128 channels.
      const uint32_t it = instr_desc_tf32(128, kTileRows);
      for (int h = 0; h < (p.c[2] >> 7); ++h) {
        const uint64_t dwh = dw + (uint64_t)(h * (128 * 128 >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(d + h * 128, dwh + 2 * k, da + 2 * k, it, (ch | k) != 0);
      }
    } else if (ksteps == 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32(d, da + 2 * k, dw + 2 * k, idesc, ch0 | ch | k);
    } else {
      for (int k = 0; k < ksteps; ++k) mma_tf32(d, da + 2 * k, dw + 2 * k, idesc, ch0 | ch | k);
    }
    if (!p.resident) {
      mma_commit(c.wempty0 + 8 * slot);
      if (++slot == (uint32_t)p.slots) {
        slot = 0;
        wph ^= 1;
      }
    }
  }
  mma_commit(c.acc_full0 + 8 * ln);
  return true;
}

__global__ void __launch_bounds__(kThreads, 1) sa_fused_fwd_kernel(const SaParams p) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned chunks: align the carve-up by hand
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = smem_layout(p);
  const unsigned long long trace_t0 = trace_begin();

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = tid & 31u;
  const int b = blockIdx.y;
  const int m_base = blockIdx.x * p.G;
  const int cpt = kTileRows / p.ns;  // centres per tile
  const int ntiles = min(p.tiles, (min(p.G, p.M - m_base) + cpt - 1) / cpt);
  const int nrounds = (ntiles + p.lanes - 1) / p.lanes;
  const int ngroups = p.npass + 2;

  unsigned char* ring = smem + L.ring;
  int32_t* rows = reinterpret_cast<int32_t*>(smem + L.rows);
  float4* centres = reinterpret_cast<float4*>(smem + L.centres);
  float* bias_s = reinterpret_cast<float*>(smem + L.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxSlots + 2 * kMaxLanes);
  volatile int* failed = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const uint32_t wfull0 = smem_u32(&bars[0]), wempty0 = smem_u32(&bars[kMaxSlots]);
  const uint32_t op_ready0 = smem_u32(&bars[2 * kMaxSlots]), acc_full0 = smem_u32(&bars[2 * kMaxSlots + kMaxLanes]);
  const uint32_t act_u32 = smem_u32(smem), ring_u32 = smem_u32(ring);
  const int wpl = kWorkerWarps / p.lanes;  // worker warps per lane

  if (tid == 0) {
    *failed = 0;
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(wfull0 + 8 * s, 1);
      mbar_init(wempty0 + 8 * s, 1);
    }
    for (int l = 0; l < kMaxLanes; ++l) {
      mbar_init(op_ready0 + 8 * l, wpl);
      mbar_init(acc_full0 + 8 * l, 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  }
  for (int i = tid; i < p.c[0] + p.c[1] + p.c[2]; i += kThreads)
    bias_s[i] = __ldg(p.bias + i);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int nch0 = (p.K0 + 31) >> 5;
  const IssueCtx ictx{tmem, act_u32, ring_u32, wfull0, wempty0, acc_full0, nch0, failed};

  if (warp == 0) {
    // =============================== MMA issuer ===============================
    // Streamed weights only: MMA groups are issued in the fixed order the producer warp mirrors.
    // (With resident weights every lane issues its own MMAs, see publish() in the worker code.)
    // The k-th group of a lane waits for the k-th completion of that lane's op_ready barrier.
    if (lane == 0 && !p.resident) {
      uint32_t slot = 0, wph = 0;
      bool ok = true;
      for (int r = 0; r < nrounds && ok; ++r) {
        for (int g = 0; g < ngroups && ok; ++g) {
          for (int ln = 0; ln < p.lanes && ok; ++ln) {
            if (r * p.lanes + ln >= ntiles) break;
            ok = wait_or_fail(op_ready0 + 8 * ln, (uint32_t)(r * ngroups + g) & 1u, failed, 1);
            ok = ok && issue_group(p, ictx, ln, g, slot, wph);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================= weight producer ============================
    if (lane == 0) {
      if (p.resident) {
        const float* src = p.wpack;
        int s = 0;
        for (int l = 0; l < 3; ++l) {
          const int nch = l == 0 ? nch0 : (p.c[l - 1] >> 5);
          const uint32_t bytes = p.c[l] * 128;
          for (int ch = 0; ch < nch; ++ch, ++s) {
            mbar_expect_tx(wfull0 + 8 * s, bytes);
            bulk_g2s(ring_u32 + s * p.slot_bytes, src, bytes, wfull0 + 8 * s);
            src += bytes >> 2;
          }
        }
      } else {
        uint32_t slot = 0, ph = 0;
        bool ok = true;
        for (int r = 0; r < nrounds && ok; ++r) {
          for (int g = 0; g < ngroups && ok; ++g) {
            int layer, ch0, nch;
            group_span(p, nch0, g, layer, ch0, nch);
            const uint32_t bytes = p.c[layer] * 128;
            // chunk offsets in wpack: layer 0 at 0, layer 1 after nch0 chunks of c0 rows, ...
            const float* base = p.wpack;
            if (layer >= 1) base += (size_t)nch0 * p.c[0] * 32;
            if (layer >= 2) base += (size_t)(p.c[0] >> 5) * p.c[1] * 32;
            for (int ln = 0; ln < p.lanes && ok; ++ln) {
              if (r * p.lanes + ln >= ntiles) break;
              for (int ch = 0; ch < nch; ++ch) {
                ok = wait_or_fail(wempty0 + 8 * slot, ph ^ 1, failed, 3);
                if (!ok) break;
                mbar_expect_tx(wfull0 + 8 * slot, bytes);
                bulk_g2s(ring_u32 + slot * p.slot_bytes, base + (size_t)(ch0 + ch) * (bytes >> 2), bytes,
                         wfull0 + 8 * slot);
                if (++slot == (uint32_t)p.slots) {
                  slot = 0;
                  ph ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else {
    // ================================= workers ================================
    const int wt = tid - 64;
    const int ww = wt >> 5;
    int nstamp = 0;
    const bool prof = p.prof != nullptr && wt == 0 && blockIdx.x == 0 && blockIdx.y == 0;
#define SA_STAMP()                                             \
  do {                                                         \
    if (prof && nstamp < 62) p.prof[1 + nstamp++] = clock64(); \
  } while (0)
    SA_STAMP();
    const float* cloud = p.xyz + (long)b * p.N * 3;
    const int ns = p.ns;

    // ---- phase Q: neighbour index rows of the G centres into shared memory (all 16 warps)
    {
      const int m = m_base + ww;
      const bool active = ww < p.G && m < p.M;
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (active) {
        const float* c = p.new_xyz + ((long)b * p.M + m) * 3;
        cx = __ldg(c + 0);
        cy = __ldg(c + 1);
        cz = __ldg(c + 2);
      }
      if (ww < p.G && lane == 0) centres[ww] = make_float4(cx, cy, cz, 0.f);
      int32_t* row = rows + ww * ns;  // only dereferenced when ww < G
      if (p.query && p.grid) {
        // exact grid query: each warp tests only the 3x3x3 cell neighbourhood of its centre
        if (ww < p.G) {
          int* scratch = reinterpret_cast<int*>(smem) + ww * (kGridCap + kGridHist);  // A0 regions are free
          if (active) {
            const BallGridView g = ball_grid_view(p.grid, b, p.N);
            ball_query_warp(g, cloud, p.N, cx, cy, cz, p.min_r2, p.max_r2, ns, row, scratch,
                            scratch + kGridCap, lane);
            if (p.idx) {
              int32_t* grow = p.idx + ((long)b * p.M + m) * ns;
              for (int l = lane; l < ns; l += 32) grow[l] = row[l];
            }
          } else {
            for (int l = lane; l < ns; l += 32) row[l] = 0;
          }
        }
      } else if (p.query) {
        float* tile = reinterpret_cast<float*>(smem);  // the A0 regions are free until the first gather
        int cnt = 0, first = 0;
        bool done = !active;
        for (int base = 0; base < p.N; base += kCloudTile) {
          const int npts = min(kCloudTile, p.N - base);
          named_sync(1, kWorkers);
          for (int i = wt; i < npts * 3; i += kWorkers) tile[i] = __ldg(cloud + (long)base * 3 + i);
          named_sync(1, kWorkers);
          if (!done) {
            for (int j = 0; j < npts; j += 32) {
              const int q = j + lane;
              bool hit = false;
              if (q < npts) {
                const float d2 = sqdist(cx, cy, cz, tile[q * 3 + 0], tile[q * 3 + 1], tile[q * 3 + 2]);
                hit = (d2 == 0.f) || (d2 >= p.min_r2 && d2 < p.max_r2);
              }
              const unsigned ballot = __ballot_sync(0xffffffffu, hit);
              if (ballot) {
                if (cnt == 0) first = base + j + (__ffs(ballot) - 1);
                const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
                if (hit && pos < ns) row[pos] = base + q;
                cnt += __popc(ballot);
                if (cnt >= ns) {
                  done = true;
                  break;
                }
              }
            }
          }
          if (worker_sync_and(done)) break;
        }
        if (ww < p.G) {
          if (cnt > ns) cnt = ns;
          __syncwarp();
          for (int l = cnt + lane; l < ns; l += 32) row[l] = first;  // inactive / empty ball: 0
          __syncwarp();
          if (active && p.idx) {
            int32_t* grow = p.idx + ((long)b * p.M + m) * ns;
            for (int l = lane; l < ns; l += 32) grow[l] = row[l];
          }
        }
      } else if (ww < p.G) {
        const int32_t* grow = p.idx + ((long)b * p.M + (active ? m : 0)) * ns;
        for (int l = lane; l < ns; l += 32) row[l] = active ? __ldg(grow + l) : 0;
      }
      named_sync(1, kWorkers);  // rows + centres visible; query scratch (aliasing A0) no longer used
    }
    SA_STAMP();

    // ---- phase M: this warp's lane works through tiles ln, ln + lanes, ...
    const int ln = ww / wpl;         // lane (tile pipeline) of this warp
    const int wl = ww - ln * wpl;    // warp within the lane
    const int lt = wl * 32 + lane;   // thread within the lane
    const int lthreads = wpl * 32;
    const int ncg = wpl >> 2;        // column groups per TMEM lane quarter
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int cg = wl >> 2;
    const uint32_t lane_base = ((uint32_t)(q * 32) << 16) + ln * p.lane_cols;
    const int r_epi = q * 32 + lane;
    unsigned char* act = smem + ln * p.lane_act_bytes;
    float* part = reinterpret_cast<float*>(smem + L.part) + ln * 2 * p.c[2];
    const uint32_t op_ready = op_ready0 + 8 * ln, acc_full = acc_full0 + 8 * ln;
    const int C4 = (p.C + 3) >> 2;  // feature slots per row; slot C4 = xyz; beyond = zero
    const int S4 = p.K0 >> 2;       // 16-byte slots per A0 row
    const bool vec = (p.C & 3) == 0;
    const float* fb = p.feat ? p.feat + (long)b * p.N * p.C : nullptr;
    const float scale = p.normalize_xyz ? p.inv_radius : 1.f;
    uint32_t accp = 0;
    bool ok = true;
    // The operand of MMA group g of this lane's tile is in shared memory: hand it to the tensor pipe.
    // Resident weights: the lane's first thread issues the MMAs itself (no hand-off latency, lanes
    // issue in parallel); streamed weights: signal the issuer warp, which keeps the ring order.
    auto publish = [&](int g, bool first_tile) {
      fence_proxy_async();
      tc_fence_before_sync();
      if (p.resident) {
        named_sync(2 + ln, lthreads);
        if (lt == 0) {
          uint32_t unused_slot = 0, unused_ph = 0;
          issue_group(p, ictx, ln, g, unused_slot, unused_ph, first_tile);
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(op_ready);
      }
    };

    // two tail elements (unaligned feature slot / xyz slot / zero slot) of tile t for this thread
    auto tail_load = [&](int t, int s_lo, int t_lo, int w, int e0, float4 (&v)[2], int (&off)[2]) {
      const int32_t* trow = rows + t * kTileRows;
      const int total = kTileRows * w;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = e0 + u * lthreads;
        off[u] = -1;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < total) {
          const int r = e / w;
          const int j = t_lo + (e - r * w);
          const int k = trow[r];
          off[u] = ((j - s_lo) >> 3) * kChunkBytes + sw128_offset(r, j & 7);
          if (j < C4) {
            const float* f = fb + (long)k * p.C + j * 4;
            const int left = p.C - j * 4;
            v[u].x = __ldg(f);
            v[u].y = left > 1 ? __ldg(f + 1) : 0.f;
            v[u].z = left > 2 ? __ldg(f + 2) : 0.f;
            v[u].w = left > 3 ? __ldg(f + 3) : 0.f;
          } else if (j == C4) {
            const float4 c = centres[t * cpt + r / ns];
            const float* pt = cloud + (long)k * 3;
            v[u].x = __fsub_rn(__ldg(pt + 0), c.x);
            v[u].y = __fsub_rn(__ldg(pt + 1), c.y);
            v[u].z = __fsub_rn(__ldg(pt + 2), c.z);
            if (p.normalize_xyz) {
              v[u].x = __fmul_rn(v[u].x, scale);
              v[u].y = __fmul_rn(v[u].y, scale);
              v[u].z = __fmul_rn(v[u].z, scale);
            }
          }
        }
      }
    };
    for (int t = ln; t < ntiles && ok; t += p.lanes) {
      const int32_t* trow = rows + t * kTileRows;  // rows of centre g are contiguous: g*ns
      // ---- layer 0: gather the 128 grouped rows (one pass = `cpp` chunks = 8*cpp slots per row)
      for (int pass = 0; pass < p.npass && ok; ++pass) {
        const int s_lo = pass * p.cpp * 8, s_hi = min(S4, s_lo + p.cpp * 8);
        // (a) feature slots: batches of independent 16-byte loads, then the swizzled stores
        const int f_hi = min(s_hi, C4);
        if (vec && f_hi > s_lo) {
          const int w = f_hi - s_lo, total = kTileRows * w;
          const int w_shift = (w & (w - 1)) == 0 ? 31 - __clz(w) : -1;
          constexpr int U = 4;
          for (int e0 = lt; e0 < total; e0 += lthreads * U) {
            float4 v[U];
            int off[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int e = e0 + u * lthreads;
              off[u] = -1;
              if (e < total) {
                const int r = w_shift >= 0 ? (e >> w_shift) : (e / w);
                const int j = s_lo + (e - r * w);
                v[u] = __ldg(reinterpret_cast<const float4*>(fb + (long)trow[r] * p.C) + j);
                off[u] = ((j - s_lo) >> 3) * kChunkBytes + sw128_offset(r, j & 7);
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (off[u] >= 0) *reinterpret_cast<float4*>(act + off[u]) = tf32_operand4(v[u]);
          }
        }
        // (b) the remaining slots of the pass: unaligned features, the xyz slot, zero padding
        const int t_lo = vec ? max(s_lo, C4) : s_lo;
        if (s_hi > t_lo) {
          const int w = s_hi - t_lo, total = kTileRows * w;
          for (int e0 = lt; e0 < total; e0 += 2 * lthreads) {
            float4 v[2];
            int off[2];
            tail_load(t, s_lo, t_lo, w, e0, v, off);  // both elements' loads are issued before either store
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (off[u] >= 0) *reinterpret_cast<float4*>(act + off[u]) = tf32_operand4(v[u]);
          }
        }
        publish(pass, t == ln);
        SA_STAMP();
        if (pass + 1 < p.npass) {  // the next pass overwrites the region: its MMAs must be done
          ok = wait_or_fail(acc_full, accp, failed, 6, p.sleep_ns);
          accp ^= 1;
        }
      }

      // ---- layers 0 and 1: accumulator -> bias, ReLU, tf32 -> next layer's A operand
      for (int l = 0; l < 2 && ok; ++l) {
        ok = wait_or_fail(acc_full, accp, failed, 4, p.sleep_ns);
        accp ^= 1;
        tc_fence_after_sync();
        SA_STAMP();
        const float* bl = bias_s + (l == 0 ? 0 : p.c[0]);
        for (int blk = cg; blk < (p.c[l] >> 5); blk += ncg) {
          uint32_t u[32];
          tmem_ld32(tmem + lane_base + p.acc_col[l] + blk * 32, u);
          tmem_ld_wait();
          unsigned char* dst = act + blk * kChunkBytes;
          const float4* bl4 = reinterpret_cast<const float4*>(bl + blk * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bj = bl4[j];
            float4 v;
            // FADD2 + VIADDMNMX: 1.5 issue slots per element instead of 3 (same bits as fmaxf + rounding add)
            add_pair(u[4 * j + 0], u[4 * j + 1], bj.x, bj.y, v.x, v.y);
            add_pair(u[4 * j + 2], u[4 * j + 3], bj.z, bj.w, v.z, v.w);
            v.x = relu_tf32_op(v.x);
            v.y = relu_tf32_op(v.y);
            v.z = relu_tf32_op(v.z);
            v.w = relu_tf32_op(v.w);
            *reinterpret_cast<float4*>(dst + sw128_offset(r_epi, j)) = v;
          }
        }
        publish(p.npass + l, t == ln);
        SA_STAMP();
      }
      if (!ok) break;

      // ---- layer 2: max over the ns rows of each centre, then bias + ReLU (both commute with max)
      ok = wait_or_fail(acc_full, accp, failed, 5, p.sleep_ns);
      accp ^= 1;
      tc_fence_after_sync();
      SA_STAMP();
      const float* b2 = bias_s + p.c[0] + p.c[1];
      const int c3 = p.c[2];
      const int m_tile = m_base + t * cpt;
      if (p.t2) {
        // transposed accumulator: this thread owns channel ch of every 128-channel half, the 128
        // columns are the tile's grouped rows -> the max over a centre's ns rows is in-thread
        const int upb = ns == 64 ? 2 : 1;            // 32-column blocks per unit (a centre or a pair)
        const int nunits = (c3 >> 7) * (4 / upb);
        for (int uidx = cg; uidx < nunits; uidx += ncg) {
          const int h = uidx / (4 / upb), blk0 = (uidx - h * (4 / upb)) * upb;
          const int ch = h * 128 + q * 32 + (int)lane;
          const int mt = m_tile;
          const float bias_c = b2[ch];
          uint32_t u[32];
          tmem_ld32(tmem + lane_base + p.acc_col[2] + h * 128 + blk0 * 32, u);
          tmem_ld_wait();
          float m0 = __uint_as_float(u[0]), m1 = __uint_as_float(u[16]);
#pragma unroll
          for (int i = 1; i < 15; i += 2) {            // FMNMX3: two elements per issue slot
            m0 = max3(m0, __uint_as_float(u[i]), __uint_as_float(u[i + 1]));
            m1 = max3(m1, __uint_as_float(u[16 + i]), __uint_as_float(u[17 + i]));
          }
          m0 = fmaxf(m0, __uint_as_float(u[15]));
          m1 = fmaxf(m1, __uint_as_float(u[31]));
          if (ns == 16) {
            const int m = mt + blk0 * 2;
            if (m < p.M) p.out[((long)b * p.M + m) * c3 + ch] = fmaxf(m0 + bias_c, 0.f);
            if (m + 1 < p.M) p.out[((long)b * p.M + m + 1) * c3 + ch] = fmaxf(m1 + bias_c, 0.f);
          } else {
            m0 = fmaxf(m0, m1);
            if (ns == 64) {
              tmem_ld32(tmem + lane_base + p.acc_col[2] + h * 128 + (blk0 + 1) * 32, u);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i += 2) m0 = max3(m0, __uint_as_float(u[i]), __uint_as_float(u[i + 1]));
            }
            const int m = mt + blk0 / upb;
            if (m < p.M) p.out[((long)b * p.M + m) * c3 + ch] = fmaxf(m0 + bias_c, 0.f);
          }
        }
        tc_fence_before_sync();
        SA_STAMP();
        continue;
      }
      float keep[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int blk = cg + ncg * it;
        if (blk >= (c3 >> 5)) break;
        uint32_t u[32];
        tmem_ld32(tmem + lane_base + p.acc_col[2] + blk * 32, u);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
        if (ns == 16) {
          lane_max_transpose<false>(v, lane);
          const int m = m_tile + q * 2 + (lane >> 4);
          if (m < p.M) {
            const int col = blk * 32 + 2 * (lane & 15);
            float2 o;
            o.x = fmaxf(v[0] + b2[col], 0.f);
            o.y = fmaxf(v[1] + b2[col + 1], 0.f);
            *reinterpret_cast<float2*>(p.out + ((long)b * p.M + m) * c3 + col) = o;
          }
        } else {
          lane_max_transpose<true>(v, lane);
          const int col = blk * 32 + lane;
          if (ns == 32) {
            const int m = m_tile + q;
            if (m < p.M) p.out[((long)b * p.M + m) * c3 + col] = fmaxf(v[0] + b2[col], 0.f);
          } else {  // ns == 64: a centre spans two lane quarters -> combine through smem
            if (q & 1) {
              part[(q >> 1) * c3 + col] = v[0];
            } else {
              keep[it] = v[0];
            }
          }
        }
      }
      if (ns == 64) {
        named_sync(2 + ln, lthreads);
        if (!(q & 1)) {
          const int m = m_tile + (q >> 1);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int blk = cg + ncg * it;
            if (blk >= (c3 >> 5)) break;
            const int col = blk * 32 + lane;
            const float mx = fmaxf(keep[it], part[(q >> 1) * c3 + col]);
            if (m < p.M) p.out[((long)b * p.M + m) * c3 + col] = fmaxf(mx + b2[col], 0.f);
          }
        }
      }
      tc_fence_before_sync();
      SA_STAMP();
    }
    if (prof) p.prof[0] = nstamp;
  }

  // ---- teardown
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_free(tmem, p.tmem_cols);
  }
  trace_end(2, trace_t0);
}

// (Cout, Cin) row-major fp32 -> ceil(Cin/32) chunks of Cout x 32 floats in the swizzled K-major
// operand image the kernel bulk-copies, zero padded, rounded to TF32 (nearest, ties away).
__global__ void sa_pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int chunks,
                                       float* __restrict__ packed) {
  const long total = (long)chunks * Cout * 32;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e / ((long)Cout * 32));
    const int rem = (int)(e - (long)c * Cout * 32);
    const int n = rem >> 5, kk = rem & 31;
    const int k = c * 32 + kk;
    const float v = k < Cin ? tf32_rna(__ldg(w + (long)n * Cin + k)) : 0.f;
    const long byte = (long)c * Cout * 128 + sw128_offset(n, kk >> 2) + (kk & 3) * 4;
    packed[byte >> 2] = v;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Pick lanes / passes / ring depth for one level. Returns false when nothing fits.
bool configure(SaParams& p, int B) {
  const int c1 = p.c[0], c2 = p.c[1], c3 = p.c[2];
  const int cpt = kTileRows / p.ns;
  const int nch0 = (p.K0 + 31) / 32, nch1 = c1 / 32, nch2 = c2 / 32;
  const int total_chunks = nch0 + nch1 + nch2;
  const int cmax = c1 > c2 ? (c1 > c3 ? c1 : c3) : (c2 > c3 ? c2 : c3);
  const int act_chunks = nch1 > nch2 ? nch1 : nch2;
  const int budget = 227 * 1024;
  p.slot_bytes = cmax * 128;
  const int max_tiles = 16 / cpt > 0 ? 16 / cpt : 1;  // <= 16 centres per CTA (one per warp in phase Q)
  const int scratch = p.query ? kCloudTile * 12 : 0;  // phase Q borrows the operand regions

  for (int lanes = kMaxLanes; lanes >= 1; lanes >>= 1) {
    if (lanes > max_tiles || lanes > g_sa_max_lanes) continue;
    const int lane_cols = 512 / lanes;
    if (c1 + c2 > lane_cols || c3 > lane_cols) continue;
    // layer-0 passes of `cpp` chunks: fewer passes first; a pass shorter than the activation
    // region would not shrink the lane's operand region any further
    const int cpp_min = nch0 < act_chunks ? nch0 : act_chunks;
    for (int cpp = nch0; cpp >= cpp_min; --cpp) {
      p.lanes = lanes;
      p.cpp = cpp;
      p.npass = (nch0 + cpp - 1) / cpp;
      int region = (cpp > act_chunks ? cpp : act_chunks) * kChunkBytes;
      while (lanes * region < scratch) region += kChunkBytes;
      p.lane_act_bytes = region;
      p.tiles = max_tiles;
      p.G = max_tiles * cpt;
      // resident weights if they fit, else the deepest ring (>= 2 slots) that does
      p.resident = 1;
      p.slots = total_chunks;
      p.slot_bytes = cmax * 128;
      if (total_chunks > kMaxSlots || smem_layout(p).total > budget) {
        p.resident = 0;
        p.slots = 0;
        for (int sl = kMaxSlots; sl >= 2; --sl) {
          p.slots = sl;
          if (smem_layout(p).total <= budget) break;
          p.slots = 0;
        }
      }
      if (p.slots == 0) continue;
      p.lane_cols = lane_cols;
      const int a1 = c1;
      p.acc_col[0] = 0;
      p.acc_col[1] = a1;
      p.acc_col[2] = (a1 + c2 + c3 <= lane_cols) ? a1 + c2 : 0;
      const int top = p.acc_col[2] + c3 > a1 + c2 ? p.acc_col[2] + c3 : a1 + c2;
      const int used = (lanes - 1) * lane_cols + top;
      p.tmem_cols = used <= 32 ? 32 : used <= 64 ? 64 : used <= 128 ? 128 : used <= 256 ? 256 : 512;
      // fewer tiles per CTA when the grid would not fill the GPU (but keep every lane busy) -- unless halving
      // costs a whole extra wave of CTAs for the same number of tile rounds per CTA (the vote-aggregation level:
      // 128 CTAs x 2 overlapped tiles beat 256 CTAs x 1 tile, 58 vs 66 us)
      int tiles = max_tiles;
      auto cost = [&](int t) {
        const long ctas = (long)B * ((p.M + t * cpt - 1) / (t * cpt));
        return ((ctas + kNumSMs - 1) / kNumSMs) * ((t + lanes - 1) / lanes);
      };
      while (tiles > 1 && tiles / 2 >= lanes &&
             (long)B * ((p.M + tiles * cpt - 1) / (tiles * cpt)) < kNumSMs && cost(tiles / 2) < cost(tiles))
        tiles >>= 1;
      p.tiles = tiles;
      p.G = tiles * cpt;
      p.t2 = (c3 % 128 == 0) ? 1 : 0;
      return true;
    }
  }
  return false;
}

}  // namespace
DEMF_DEFINE_TRACE_SETTER(trace_set_sa)
}  // namespace demf

using namespace demf;

extern "C" {

long demf_sa_pack_floats(int Cout, int Cin) { return (long)((Cin + 31) / 32) * Cout * 32; }

int demf_sa_pack_weights(const float* w, int Cout, int Cin, float* packed, void* stream) {
  DEMF_REQUIRE_PTR(w);
  DEMF_REQUIRE_PTR(packed);
  DEMF_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cin > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(aligned16(packed), DEMF_E_UNSUPPORTED);
  const int chunks = (Cin + 31) / 32;
  const long total = (long)chunks * Cout * 32;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  sa_pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, Cout, Cin, chunks, packed);
  return after_launch("sa_pack_weights_kernel");
}

int demf_sa_fused_supported(int C, int ns, int c1, int c2, int c3) {
  if (!(ns == 16 || ns == 32 || ns == 64)) return 0;
  if (C < 0 || C > 512) return 0;
  const int cs[3] = {c1, c2, c3};
  for (int c : cs)
    if (c < 32 || c > 256 || c % 32 != 0) return 0;
  return 1;
}

/* debug: device buffer of 64 int64 that receives [count, clock64 stamps...] of one worker thread */
int demf_sa_fused_set_profile(long long* device_buffer) {
  g_sa_prof = device_buffer;
  return 0;
}

/* tuning knobs (development): most tile pipelines per CTA (1, 2 or 4) and the worker warps'
 * back-off between mbarrier polls in ns (0 = spin) */
int demf_sa_fused_tune(int max_lanes, int sleep_ns) {
  g_sa_max_lanes = max_lanes < 1 ? 1 : max_lanes;
  g_sa_sleep_ns = sleep_ns < 0 ? 0 : sleep_ns;
  return 0;
}

int demf_sa_fused_error(void) {
  int e = 0;
  cudaMemcpyFromSymbol(&e, g_sa_error, sizeof(int));
  return e;
}

static int sa_fused_launch(const float* xyz, const float* feat_rows, const float* new_xyz, int B, int N, int M,
                           int C, float min_radius, float max_radius, int ns, int normalize_xyz, int query,
                           const float* wpack, const float* bias, int c1, int c2, int c3, const void* grid,
                           int32_t* idx, float* out, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(wpack);
  DEMF_REQUIRE_PTR(bias);
  DEMF_REQUIRE_PTR(out);
  if (C > 0) DEMF_REQUIRE_PTR(feat_rows);
  if (!query) DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  DEMF_REQUIRE(demf_sa_fused_supported(C, ns, c1, c2, c3), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(aligned16(wpack) && aligned16(out) && (C % 4 != 0 || aligned16(feat_rows)), DEMF_E_UNSUPPORTED);
  if (B == 0 || M == 0) return 0;

  SaParams p{};
  p.xyz = xyz;
  p.feat = C > 0 ? feat_rows : nullptr;
  p.new_xyz = new_xyz;
  p.wpack = wpack;
  p.bias = bias;
  p.out = out;
  p.idx = idx;
  p.N = N;
  p.M = M;
  p.C = C;
  p.ns = ns;
  p.K0 = ((((C + 3) / 4) * 4 + 4) + 7) / 8 * 8;
  p.c[0] = c1;
  p.c[1] = c2;
  p.c[2] = c3;
  p.query = query;
  p.normalize_xyz = normalize_xyz;
  p.min_r2 = min_radius * min_radius;
  p.max_r2 = max_radius * max_radius;
  p.inv_radius = 1.0f / max_radius;
  p.prof = g_sa_prof;
  p.sleep_ns = g_sa_sleep_ns;
  p.grid = (query && ns <= kGridCap) ? grid : nullptr;
  // With a grid and an index buffer from the caller, the neighbour search runs as its own launch:
  // it is a latency-bound index chase that wants 64 warps per SM, while this kernel -- one CTA per
  // SM because of its operand regions -- could only give it 16.
  if (p.grid && idx) {
    const int rc = launch_ball_query_grid(xyz, new_xyz, grid, B, N, M, min_radius, max_radius, ns, idx,
                                          as_stream(stream));
    if (rc != 0) return rc;
    p.grid = nullptr;
    p.query = query = 0;
  }
  static_assert(kWorkerWarps * (kGridCap + kGridHist) * 4 <= kCloudTile * 12, "grid scratch fits the tile");
  DEMF_REQUIRE(configure(p, B), DEMF_E_UNSUPPORTED);
  const SmemLayout L = smem_layout(p);

  auto kernel = sa_fused_fwd_kernel;
  static int configured_smem = 0;
  if (L.total > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) {
      set_error("demf_sa_fused_fwd: cannot reserve %d bytes of shared memory: %s", L.total,
                cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_smem = L.total;
  }
  dim3 grid_dim((M + p.G - 1) / p.G, B);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid_dim;
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = as_stream(stream);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) {
    set_error("demf_sa_fused_fwd: launch failed: %s (grid %u x %u, %d B smem, lanes %d, slots %d, resident %d)",
              cudaGetErrorString(e), grid_dim.x, grid_dim.y, L.total, p.lanes, p.slots, p.resident);
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return after_launch("sa_fused_fwd_kernel");
}

int demf_sa_fused_fwd(const float* xyz, const float* feat_rows, const float* new_xyz, int B, int N, int M,
                      int C, float min_radius, float max_radius, int ns, int normalize_xyz, int query,
                      const float* wpack, const float* bias, int c1, int c2, int c3, const void* grid,
                      int32_t* idx, float* out, void* stream) {
  return sa_fused_launch(xyz, feat_rows, new_xyz, B, N, M, C, min_radius, max_radius, ns, normalize_xyz, query,
                         wpack, bias, c1, c2, c3, grid, idx, out, stream);
}

}  // extern "C"
