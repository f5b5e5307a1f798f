// TF32 tensor-core GEMMs for the TRAINING step of the shared MLPs (sm_100a): tcgen05.mma with TMEM
// accumulators, every operand tile moved by the TMA engine through tensor maps (cp.async.bulk.tensor,
// SWIZZLE_128B), results stored by TMA as well. Three kernels replace what upstream runs through
// cuDNN / cuBLAS for the 1x1 convolutions of mmcv ConvModule (conv -> BN -> ReLU) inside mmdet3d
// PointSAModule / PointFPModule / VoteModule / BaseConvBboxHead
// (configs/demf/demf_votenet.py:48-62,142-162; demf/modeling/heads/class_agnostic_vote_head.py:382-403):
//
//   rows_gemm_kernel<false>  forward      Y (R,N)   = X (R,K) W^T   W (N,K) row-major (K-major B operand)
//                            + optional bias, ReLU, and the per-channel sum / sum of squares of Y that the
//                            BatchNorm behind the convolution needs (fp32 per CTA, double atomics into the
//                            layer's accumulator block): the separate statistics pass over Y disappears.
//   rows_gemm_kernel<true>   data grad    dX (R,K)  = dY (R,N) W    same W, read as an MN-major B operand
//   wgrad_kernel             weight grad  dW (N,K) += dY^T X        both operands MN-major straight from the
//                            row-major activations (no transposes), split over row slabs, partial tiles
//                            accumulated with red.global.add.v4.f32. Replaces the library's sm80 split-K
//                            kernel (cutlass_80_tensorop_s1688gemm_*_nt) that ran at ~55 % of HBM speed.
//
// All three are HBM-bound streaming kernels: R is 10^4..10^6 rows, K and N at most 512. Algorithmic bytes:
// forward / data grad R*(K+N)*4, weight grad R*(K+N)*4 (+ the (N,K) tile once per slab).
//
// Operand convention (umma.cuh): a tile is a stack of 128-byte rows, SWIZZLE_128B (16-byte unit j of row r
// at ((j ^ (r & 7)) << 4)), 1024-byte aligned -- exactly what a TMA box {32 floats, rows} with
// CU_TENSOR_MAP_SWIZZLE_128B writes. K-major operand: rows = M/N index, the 32 floats = K (k step of 8 =
// +32 bytes). MN-major operand: rows = K index, the 32 floats = M/N index; a k step of 8 = +1024 bytes, the
// next 32 M/N indices are the next box (leading-dimension byte offset = box bytes).
#include <cuda.h>  // CUtensorMap and its enums only: cuTensorMapEncodeTiled is resolved at run time

#include "common.cuh"
#include "umma.cuh"

namespace demf {
namespace {

using namespace umma;

__device__ int g_gemm_error = 0;  // sticky: first mbarrier time-out (never expected)

__device__ __forceinline__ void wait_or_flag(uint32_t bar, uint32_t parity, int code) {
  if (!mbar_wait(bar, parity)) atomicCAS(&g_gemm_error, 0, code);
}

// ---- TMA through tensor maps --------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
// bring one box into L2 ahead of ordinary loads of it (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// MN-major TF32 operand. The only shared-memory layout the tensor core takes for 32-bit MN-major operands is the
// 128-byte swizzle with 32-BYTE atoms (UMMA layout type 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a tile is
// a stack of 128-byte rows (row = one k, the 32 floats = 32 consecutive M/N indices) whose four 32-byte chunks
// are permuted by (k & 3); 4 k rows = one 512-byte atom. `lbo` = bytes between consecutive groups of 32 M/N
// indices (= one TMA box), stride-byte-offset = 512 between consecutive groups of 4 k.
struct MnTune { int sbo; int layout; };
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t addr, uint32_t lbo, MnTune t) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(((uint32_t)t.sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>((uint32_t)t.layout & 7u) << 61;
  return d;
}

__host__ __device__ constexpr uint32_t instr_desc_tf32_major(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

constexpr int kTileRows = 128;                 // UMMA M
constexpr int kChunk = 32;                     // floats per 128-byte operand row
constexpr int kABytes = kTileRows * 128;       // one A chunk: 128 rows x 32 k
constexpr int kStageOut = kTileRows * 128;     // one 32-column slab of the output tile
constexpr int kGemmThreads = 320;              // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 two epilogue groups
constexpr int kWgradThreads = 192;             // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kSmemBudget = 227 * 1024 - 1024; // dynamic shared memory we allow ourselves (1 KB alignment slack on top)

struct RowsGemmParams {
  long R;          // rows
  int K;           // reduction length (valid columns of A)
  int N;           // valid output columns in total
  int n_block;     // output columns per blockIdx.y (multiple of 16, <= 256)
  int stages;
  int relu;
  const float* bias;   // (N) or null
  double* stats;       // (2, N) double accumulators of a BatchNorm layer's state block, or null
  MnTune mn;
  int epi_groups;      // 1 or 2 epilogue warp groups in use
  int col_mode;        // 0 none; 1 forward: sum y, sum y^2; 2 data gradient through BN+ReLU: mask, sum g, sum g*y
  const float* bn_y;   // mode 2: pre-BN activations of the layer whose output gradient this kernel produces (R,N)
  long bn_ldy;
  const float* bn_mean;
  const float* bn_invstd;
  const float* bn_gamma;
  const float* bn_beta;
};

// ================================================================ forward / data gradient ==========
// grid = (CTAs over row tiles [persistent], N blocks). Shared memory: `stages` x (A chunk 16 KB + B chunk
// n_block x 128 B), two 16 KB output slabs, barriers, per-CTA statistics.
template <bool kBMN>
__global__ void __launch_bounds__(kGemmThreads, 1)
rows_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_p,
                 const RowsGemmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 1 KB aligned
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int nmma = p.n_block;                              // MMA N (multiple of 16)
  const int b_bytes = kBMN ? ((nmma + 31) / 32) * 4096 : nmma * 128;
  const int stage_bytes = kABytes + ((b_bytes + 1023) & ~1023);
  unsigned char* out_stage = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + 4 * kStageOut);
  // bars: full[stages], empty[stages], tmem_full[2], tmem_empty[2]
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * p.stages;
  const uint32_t tfull0 = empty0 + 8u * p.stages, tempty0 = tfull0 + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);  // column-pass partials: [2 groups][4 warps][8][2][32]
  const int n0 = blockIdx.y * p.n_block;                   // first output column of this CTA
  const int kc_total = (p.K + kChunk - 1) / kChunk;
  const long tiles = (p.R + kTileRows - 1) / kTileRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8u * a, 1);
      mbar_init(tempty0 + 8u * a, 4);   // one arrival per epilogue warp
    }
    mbar_fence_init();
    tma_prefetch_map(&map_a);
    tma_prefetch_map(&map_b);
    tma_prefetch_map(&map_y);
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (p.col_mode == 2) {
          // the epilogue of this tile will read the previous layer's pre-BN activations (map_p): start pulling
          // them into L2 now, a whole operand pipeline ahead of their use
          for (int g = 0; n0 + g * 32 < min(p.N, n0 + p.n_block); ++g)
            tma_prefetch_l2_2d(&map_p, n0 + g * 32, (int)(tile * kTileRows));
        }
        for (int kc = 0; kc < kc_total; ++kc, ++it) {
          const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
          wait_or_flag(empty0 + 8u * s, ph ^ 1u, 1);
          const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes), b_dst = a_dst + kABytes;
          mbar_expect_tx(full0 + 8u * s, kABytes + b_bytes);
          tma_load_2d(a_dst, &map_a, kc * kChunk, (int)(tile * kTileRows), full0 + 8u * s);
          if (kBMN) {   // W (K_red, N_out) row-major: boxes of 32 output columns x 32 reduction rows
            for (int g = 0; g * 32 < nmma; ++g)
              tma_load_2d(b_dst + g * 4096, &map_b, n0 + g * 32, kc * kChunk, full0 + 8u * s);
          } else {      // W (N_out, K_red) row-major: one box of 32 reduction columns x n_block rows
            tma_load_2d(b_dst, &map_b, kc * kChunk, n0, full0 + 8u * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = instr_desc_tf32_major(kTileRows, (uint32_t)nmma, 0u, kBMN ? 1u : 0u);
      uint32_t it = 0, t_local = 0;
      for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t_local) {
        const uint32_t acc = t_local & 1u, aph = (t_local >> 1) & 1u;
        wait_or_flag(tempty0 + 8u * acc, aph ^ 1u, 2);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem + acc * 256u;
        for (int kc = 0; kc < kc_total; ++kc, ++it) {
          const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
          wait_or_flag(full0 + 8u * s, ph, 3);
          tc_fence_after_sync();
          const uint32_t a_src = smem_u32(smem + (size_t)s * stage_bytes), b_src = a_src + kABytes;
          const int kleft = p.K - kc * kChunk;
          const int ksteps = kleft >= kChunk ? 4 : (kleft + 7) / 8;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = smem_desc_sw128(a_src + ks * 32);
            const uint64_t db = kBMN ? smem_desc_sw128_mn(b_src + ks * 1024, 4096u, p.mn) : smem_desc_sw128(b_src + ks * 32);
            mma_tf32(d_tmem, da, db, idesc, (kc | ks) ? 1u : 0u);
          }
          mma_commit(empty0 + 8u * s);          // slot reusable once these MMAs have read it
        }
        mma_commit(tfull0 + 8u * acc);          // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups of 4 warps
    // Group e drains accumulator buffer e (every other tile of this CTA), so one group's TMEM reads, staging
    // and column pass overlap the other's; each group owns two 16 KB output slabs, its named barriers and its
    // bulk-store groups.
    const int egrp = (warp - 2) >> 2;                      // 0 or 1
    const int q = warp & 3;                                // TMEM lane quarter this warp may touch
    const int row_in_tile = q * 32 + (int)lane;
    const bool leader = (warp - 2) % 4 == 0 && lane == 0;  // first thread of the group: owns its bulk stores
    const int bar_a = 1 + 2 * egrp, bar_b = 2 + 2 * egrp;
    unsigned char* my_stage = out_stage + (size_t)egrp * 2 * kStageOut;
    float* my_part = s_sum + (size_t)egrp * 2048 + (size_t)(warp & 3) * 512;    // [8 groups][2][32 lanes] per warp
    const int groups = (min(p.n_block, p.N - n0) + 31) / 32;
    if (p.col_mode != 0) {
      for (int i = lane; i < 512; i += 32) my_part[i] = 0.f;
      __syncwarp();
    }
    uint32_t slab = 0;
    const uint32_t t_step = (uint32_t)p.epi_groups;
    for (uint32_t t_local = egrp; egrp < p.epi_groups && blockIdx.x + (long)t_local * gridDim.x < tiles;
         t_local += t_step) {
      const long tile = blockIdx.x + (long)t_local * gridDim.x;
      const uint32_t acc = t_local & 1u, aph = (t_local >> 1) & 1u;
      wait_or_flag(tfull0 + 8u * acc, aph, 4);
      tc_fence_after_sync();
      const long row_base = tile * kTileRows + q * 32;
      const int nvalid = (int)max(0L, min(32L, p.R - row_base));
      for (int g = 0; g < groups; ++g, ++slab) {
        const int col0 = n0 + g * 32;
        // mode 2: this lane's column of the previous layer's pre-BN activations for the warp's 32 rows -- 32
        // independent coalesced loads (one 128-byte row segment per instruction) issued before anything else
        float yv[32];
        if (p.col_mode == 2) {
          const bool col_ok = col0 + (int)lane < p.N;
          const float* ycol = p.bn_y + row_base * p.bn_ldy + col0 + (int)lane;
          if (col_ok && nvalid == 32) {       // full tile (all but the last one): no per-row predicate
#pragma unroll
            for (int i = 0; i < 32; ++i) yv[i] = __ldg(ycol + (long)i * p.bn_ldy);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) yv[i] = (col_ok && i < nvalid) ? __ldg(ycol + (long)i * p.bn_ldy) : 0.f;
          }
        }
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)g * 32u, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) f[i] += __ldg(p.bias + col0 + i);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        // ---- stage the 128 x 32 slab (swizzled rows) and hand it to the TMA engine
        unsigned char* dst = my_stage + (slab & 1u) * kStageOut;
        if (leader) bulk_wait_read<1>();                 // the store that last read this buffer has drained
        named_bar_sync(bar_a, 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float4*>(dst + sw128_offset((uint32_t)row_in_tile, (uint32_t)j)) =
              make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        if (p.col_mode != 0) {
          // Column pass over the staged slab: lane l owns column l of the group and walks the 32 rows its own
          // warp just wrote (the 32 lanes read the 32 floats of one swizzled row: conflict-free).
          //   mode 1 (forward): sum and sum of squares of y for the BatchNorm behind this convolution;
          //   mode 2 (data gradient): the slab is dL/dz of the PREVIOUS layer's BatchNorm+ReLU output: zero it
          //          where that ReLU was inactive (recomputed from the layer's pre-BN activations `bn_y` with
          //          the expression of bn_apply_kernel) and accumulate sum g and sum g*y -- the two reductions
          //          of the BatchNorm backward -- so that no separate pass over (dz, z, y) is needed.
          __syncwarp();
          const int col = col0 + (int)lane;
          const bool col_ok = col < p.N;
          float sa = 0.f, sb = 0.f;
          float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
          if (p.col_mode == 2 && col_ok) {
            mu = __ldg(p.bn_mean + col);
            is = __ldg(p.bn_invstd + col);
            ga = __ldg(p.bn_gamma + col);
            be = __ldg(p.bn_beta + col);
          }
          if (p.col_mode == 2) {
            // (row & 7) == (i & 7): the swizzled position of this lane's cell in row i is a compile-time
            // function of i once the lane's slot is known
            unsigned char* base = dst + (uint32_t)(q * 32) * 128u + ((lane & 3u) << 2);
            const uint32_t slot = lane >> 2;
            if (nvalid == 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float* cell = reinterpret_cast<float*>(base + i * 128 + ((slot ^ (uint32_t)(i & 7)) << 4));
                const float z = (yv[i] - mu) * is * ga + be;
                const float val = z > 0.f ? *cell : 0.f;
                *cell = val;
                sa += val;
                sb = fmaf(val, yv[i], sb);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (i < nvalid) {
                  float* cell = reinterpret_cast<float*>(base + i * 128 + ((slot ^ (uint32_t)(i & 7)) << 4));
                  const float z = (yv[i] - mu) * is * ga + be;
                  const float val = z > 0.f ? *cell : 0.f;
                  *cell = val;
                  sa += val;
                  sb = fmaf(val, yv[i], sb);
                }
              }
            }
          } else {
            const unsigned char* base = dst + (uint32_t)(q * 32) * 128u + ((lane & 3u) << 2);
            const uint32_t slot = lane >> 2;
            if (nvalid == 32) {
              float sa2 = 0.f, sb2 = 0.f;      // two chains: the 32 dependent adds are the pass's latency
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float v0 = *reinterpret_cast<const float*>(base + i * 128 + ((slot ^ (uint32_t)(i & 7)) << 4));
                const float v1 =
                    *reinterpret_cast<const float*>(base + (i + 1) * 128 + ((slot ^ (uint32_t)((i + 1) & 7)) << 4));
                sa += v0;
                sb = fmaf(v0, v0, sb);
                sa2 += v1;
                sb2 = fmaf(v1, v1, sb2);
              }
              sa += sa2;
              sb += sb2;
            } else {
              for (int i = 0; i < nvalid; ++i) {
                const float val = *reinterpret_cast<const float*>(base + i * 128 + ((slot ^ (uint32_t)(i & 7)) << 4));
                sa += val;
                sb = fmaf(val, val, sb);
              }
            }
          }
          if (g < 8) {
            my_part[g * 64 + lane] += sa;
            my_part[g * 64 + 32 + lane] += sb;
          }
        }
        fence_proxy_async();
        named_bar_sync(bar_b, 128);
        if (leader) {
          tma_store_2d(&map_y, col0, (int)(tile * kTileRows), smem_u32(dst));
          bulk_commit();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8u * acc) : "memory");
      }
    }
    if (leader) bulk_wait_all();
    if (p.col_mode != 0 && egrp < p.epi_groups) {
      named_bar_sync(bar_a, 128);
      const float* gp = s_sum + (size_t)egrp * 2048;
      const int ncols = min(p.n_block, p.N - n0);
      const int tid = (warp - 2) % 4 * 32 + (int)lane;
      for (int c = tid; c < ncols; c += 128) {
        const int g = c >> 5, l = c & 31;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          a += gp[w * 512 + g * 64 + l];
          b += gp[w * 512 + g * 64 + 32 + l];
        }
        atomicAdd(p.stats + n0 + c, (double)a);
        atomicAdd(p.stats + p.N + n0 + c, (double)b);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_free(tmem, 512);
  }
}

// ========================================================================= weight gradient ==========
// dW (N,K) += dY(R,N)^T X(R,K) over the rows [slab*rows_per_slab, ...). grid = (slabs, ceil(N/128)).
// A = dY^T: M index = output channel (a 128-wide block per blockIdx.y), MN-major. B = X^T: N index = input
// channel (all of them, <= 512), MN-major. One stage = 32 rows: 4 A boxes + ceil(K/32) B boxes of 4 KB.
struct WgradParams {
  long R;
  long rows_per_slab;   // multiple of rows_per_stage
  int rows_per_stage;   // 32 or 64
  MnTune mn;
  int N;                // output channels (rows of dW)
  int K;                // input channels (columns of dW)
  int ldw;              // row stride of dW in floats
  int stages;
  float* dw;
};

// rows (reduction) per stage: 64 when at least three such stages fit (fewer, larger TMA boxes), else 32

__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
             const WgradParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int kgroups = (p.K + 31) / 32;                 // B boxes per stage
  const int box_bytes = p.rows_per_stage * 128;        // one box: rows_per_stage x 32 floats
  const int stage_bytes = (4 + kgroups) * box_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * p.stages, done0 = empty0 + 8u * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  const int m0 = blockIdx.y * 128;
  const long r0 = (long)blockIdx.x * p.rows_per_slab;
  const long r1 = min(p.R, r0 + p.rows_per_slab);
  const int steps = (int)((r1 - r0 + p.rows_per_stage - 1) / p.rows_per_stage);
  const int ncols = kgroups * 32;                      // accumulator columns (<= 512)
  const uint32_t tmem_cols = ncols <= 32 ? 32u : ncols <= 64 ? 64u : ncols <= 128 ? 128u : ncols <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(done0, 1);
    mbar_fence_init();
    tma_prefetch_map(&map_dy);
    tma_prefetch_map(&map_x);
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (steps <= 0) {   // empty slab (can only be the last one): nothing to add
    __syncthreads();
    if (warp == 2) tmem_free(tmem, tmem_cols);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < steps; ++it) {
        const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
        wait_or_flag(empty0 + 8u * s, ph ^ 1u, 11);
        const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
        mbar_expect_tx(full0 + 8u * s, (uint32_t)stage_bytes);
        const int row = (int)(r0 + (long)it * p.rows_per_stage);
        // slabs are multiples of the stage height, so only the global tail (row >= R) is partial, and the
        // tensor map zero-fills it: no row is counted twice
        for (int g = 0; g < 4; ++g) tma_load_2d(dst + g * box_bytes, &map_dy, m0 + g * 32, row, full0 + 8u * s);
        for (int g = 0; g < kgroups; ++g)
          tma_load_2d(dst + (4 + g) * box_bytes, &map_x, g * 32, row, full0 + 8u * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int it = 0; it < steps; ++it) {
        const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
        wait_or_flag(full0 + 8u * s, ph, 12);
        tc_fence_after_sync();
        const uint32_t a_src = smem_u32(smem + (size_t)s * stage_bytes), b_src = a_src + 4 * box_bytes;
        for (int ks = 0; ks < p.rows_per_stage / 8; ++ks) {
          const uint64_t da = smem_desc_sw128_mn(a_src + ks * 1024, (uint32_t)box_bytes, p.mn);
          for (int nb = 0; nb * 256 < ncols; ++nb) {
            const int nw = min(256, ncols - nb * 256);
            const uint64_t db = smem_desc_sw128_mn(b_src + nb * 8 * box_bytes + ks * 1024, (uint32_t)box_bytes, p.mn);
            mma_tf32(tmem + nb * 256u, da, db, instr_desc_tf32_major(128, (uint32_t)nw, 1u, 1u),
                     (it | ks) ? 1u : 0u);
          }
        }
        mma_commit(empty0 + 8u * s);
      }
      mma_commit(done0);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + (int)lane;            // row of dW owned by this thread
    wait_or_flag(done0, 0u, 13);
    tc_fence_after_sync();
    const bool vec = (p.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15u) == 0;
    for (int g = 0; g < kgroups; ++g) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 32u, v);
      tmem_ld_wait();
      if (m < p.N) {
        float* dst = p.dw + (long)m * p.ldw + g * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = g * 32 + 4 * j;
          if (vec && c + 3 < p.K) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                         "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                         "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                         : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (c + i < p.K) atomicAdd(dst + 4 * j + i, __uint_as_float(v[4 * j + i]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_free(tmem, tmem_cols);
  }
}

// ------------------------------------------------- BatchNorm statistics handed over by the GEMM epilogue --
// accum (2,C) doubles filled by rows_gemm_kernel; one block: mean / invstd / running statistics, accum re-zeroed.
__global__ void bn_finalize_kernel(double* accum, long R, int C, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double ds = accum[c], dq = accum[C + c];
    accum[c] = 0.0;
    accum[C + c] = 0.0;
    const double m = ds / (double)R;
    double var = dq / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = 1.0f / sqrtf((float)var + eps);
    if (running_mean) {
      const double unbiased = R > 1 ? var * ((double)R / (double)(R - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// accum = (sum g, sum g*y) from the mode-2 column pass: grad_beta, grad_gamma and the two coefficients of
// bn_bwd_apply_kernel (mean(g), invstd^2 * mean(g (y - mean))); accum re-zeroed.
__global__ void bn_bwd_finalize_kernel(double* accum, long R, int C, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* __restrict__ grad_gamma,
                                       float* __restrict__ grad_beta, float* __restrict__ coef) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double ds = accum[c], dgy = accum[C + c];
    accum[c] = 0.0;
    accum[C + c] = 0.0;
    const double dq = dgy - (double)mean[c] * ds;
    const float is = invstd[c];
    grad_beta[c] = (float)ds;
    grad_gamma[c] = (float)(dq * (double)is);
    coef[c] = (float)(ds / (double)R);
    coef[C + c] = (float)(dq / (double)R * (double)is * (double)is);
  }
}

// ------------------------------------------------------------------------------- host side -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// (rows, cols) fp32 matrix with row stride `ld` floats; boxes of 32 columns x box_rows rows, SWIZZLE_128B,
// out-of-bounds elements read as zero / are not written.
int g_epi_groups = 2;
int g_mn_sbo = 512, g_mn_layout = 1, g_mn_tma = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;   // see demf_gemm_debug_mn

int make_map(CUtensorMap* map, const float* base, long rows, int cols, long ld, int box_rows, bool mn_major = false) {
  EncodeTiledFn fn = encode_tiled();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return DEMF_E_UNSUPPORTED;
  }
  // The driver entry point needs a current context in THIS thread. A thread whose first CUDA call is this one
  // (an autograd worker running our backward before any runtime call) has none yet: binding the device's
  // primary context is what the runtime would do on its first call.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess) cudaSetDevice(d);
    cudaFree(nullptr);
    ctx_bound = true;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? (CUtensorMapSwizzle)g_mn_tma : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a %ld x %d matrix, ld %ld, box 32 x %d", (int)r, rows, cols, ld,
              box_rows);
    return DEMF_E_UNSUPPORTED;
  }
  return 0;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct BnMask {   // mode-2 column pass (data gradient through the previous layer's BatchNorm + ReLU)
  const float* y = nullptr;
  long ldy = 0;
  const float* mean = nullptr;
  const float* invstd = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
};

template <bool kBMN>
int launch_rows_gemm(const float* a, long lda, const float* w, long ldw, const float* bias, double* stats, float* y,
                     long ldy, long R, int K, int N, int relu, cudaStream_t st, const BnMask* bn = nullptr) {
  // output columns per CTA: balanced blocks of at most 256, multiples of 16
  // (several blocks: multiples of 32 so that a block's last 32-column output slab never reaches into the next one)
  int nb = (N + 255) / 256;
  // Few row tiles (the small layers of the step: 1 024 - 16 384 rows): a CTA per 128 rows x 256 columns leaves most
  // SMs idle and walks its K chunks and its eight 32-column output slabs one after the other (15 us for one
  // 128 x 256 x 256 tile). Narrower column blocks (>= 64 columns) on more CTAs: shorter stages -> deeper ring,
  // two slabs per epilogue, the A tile re-read from L2.
  {
    const long tiles_ = (R + kTileRows - 1) / kTileRows;
    while (tiles_ * nb * 2 <= kNumSMs && (N + 2 * nb - 1) / (2 * nb) >= 64) nb *= 2;
  }
  // (MN-major B operand: whole 32-column swizzle atoms)
  const int n_block = (nb == 1 && !kBMN) ? ((N + 15) & ~15) : ((((N + nb - 1) / nb) + 31) & ~31);
  RowsGemmParams p;
  p.R = R;
  p.K = K;
  p.N = N;
  p.n_block = n_block;
  p.relu = relu;
  p.bias = bias;
  p.stats = stats;
  p.mn.sbo = g_mn_sbo;
  p.mn.layout = g_mn_layout;
  p.epi_groups = g_epi_groups;
  p.col_mode = stats == nullptr ? 0 : (bn != nullptr ? 2 : 1);
  p.bn_y = bn ? bn->y : nullptr;
  p.bn_ldy = bn ? bn->ldy : 0;
  p.bn_mean = bn ? bn->mean : nullptr;
  p.bn_invstd = bn ? bn->invstd : nullptr;
  p.bn_gamma = bn ? bn->gamma : nullptr;
  p.bn_beta = bn ? bn->beta : nullptr;
  const int b_bytes = kBMN ? ((n_block + 31) / 32) * 4096 : ((n_block * 128 + 1023) & ~1023);
  const int stage_bytes = kABytes + b_bytes;
  const int fixed = 4 * kStageOut + 256 /* barriers, tmem slot */ + 16384 /* column-pass partials */;
  int stages = (kSmemBudget - fixed) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) {
    set_error("rows_gemm: tile does not fit shared memory (N block %d)", n_block);
    return DEMF_E_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed + 1024;
  CUtensorMap ma, mb, my;
  if (int rc = make_map(&ma, a, R, K, lda, kTileRows)) return rc;
  if (kBMN) {
    if (int rc = make_map(&mb, w, K, N, ldw, 32, true)) return rc;      // (K_red rows, N_out cols)
  } else {
    if (int rc = make_map(&mb, w, N, K, ldw, n_block)) return rc;       // (N_out rows, K_red cols)
  }
  if (int rc = make_map(&my, y, R, N, ldy, kTileRows)) return rc;
  const long tiles = (R + kTileRows - 1) / kTileRows;
  int gx = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  if (nb > 1) gx = (int)(tiles < kNumSMs / nb ? tiles : kNumSMs / nb);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(rows_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(rows_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  CUtensorMap mp = my;
  if (bn != nullptr) {
    if (int rc = make_map(&mp, bn->y, R, N, bn->ldy, kTileRows)) return rc;
  }
  rows_gemm_kernel<kBMN><<<dim3(gx, nb), kGemmThreads, smem, st>>>(ma, mb, my, mp, p);
  return after_launch(kBMN ? "rows_gemm_kernel<dgrad>" : "rows_gemm_kernel<fwd>");
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_gemm_supported(int K, int N) {
  // TMA row strides must be multiples of 16 bytes (the weight gradient takes 512 input channels per launch)
  return (K > 0 && N > 0 && K % 4 == 0 && N % 4 == 0 && K <= 2048 && N <= 2048) ? 1 : 0;
}

/* development: the MN-major operand encoding (stride byte offset, UMMA layout type, CUtensorMapSwizzle) */
int demf_gemm_debug_mn(int sbo, int layout, int tma_swizzle) {
  g_mn_sbo = sbo;
  g_mn_layout = layout;
  g_mn_tma = tma_swizzle;
  return 0;
}

int demf_gemm_tune(int epilogue_groups) {
  g_epi_groups = epilogue_groups == 1 ? 1 : 2;
  return 0;
}

int demf_gemm_error(void) {
  int e = 0;
  cudaMemcpyFromSymbol(&e, g_gemm_error, sizeof(int));
  return e;
}

int demf_gemm_rows_fwd(const float* x, long ldx, const float* w, long ldw, const float* bias, long R, int K, int N,
                       int relu, void* bn_state, float* y, long ldy, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(w);
  DEMF_REQUIRE_PTR(y);
  DEMF_REQUIRE(R > 0 && K > 0 && N > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0 && ldx >= K && ldw >= K && ldy >= N, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(x) && al16(w) && al16(y), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(bn_state == nullptr || N <= 256, DEMF_E_UNSUPPORTED);
  return launch_rows_gemm<false>(x, ldx, w, ldw, bias, static_cast<double*>(bn_state), y, ldy, R, K, N, relu,
                                 as_stream(stream));
}

int demf_gemm_rows_dgrad(const float* dy, long lddy, const float* w, long ldw, long R, int N, int K, float* dx,
                         long lddx, void* stream) {
  DEMF_REQUIRE_PTR(dy);
  DEMF_REQUIRE_PTR(w);
  DEMF_REQUIRE_PTR(dx);
  DEMF_REQUIRE(R > 0 && K > 0 && N > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(lddy % 4 == 0 && ldw % 4 == 0 && lddx % 4 == 0 && lddy >= N && ldw >= K && lddx >= K,
               DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(dy) && al16(w) && al16(dx), DEMF_E_UNSUPPORTED);
  // dX (R,K) = dY (R,N) W (N,K): reduction over N, output columns K; W read as (N_red rows, K_out cols)
  return launch_rows_gemm<true>(dy, lddy, w, ldw, nullptr, nullptr, dx, lddx, R, N, K, 0, as_stream(stream));
}

int demf_gemm_wgrad(const float* dy, long lddy, const float* x, long ldx, long R, int N, int K, float* dw, int ldw,
                    void* stream) {
  DEMF_REQUIRE_PTR(dy);
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(dw);
  DEMF_REQUIRE(R > 0 && K > 0 && N > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(K <= 512, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && lddy >= N && ldx >= K && ldw >= K, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(dy) && al16(x) && (reinterpret_cast<uintptr_t>(dw) & 3u) == 0, DEMF_E_UNSUPPORTED);
  WgradParams p;
  p.R = R;
  p.N = N;
  p.K = K;
  p.ldw = ldw;
  p.dw = dw;
  const int kgroups = (K + 31) / 32;
  int rows_per_stage = 64;
  if ((kSmemBudget - 1024) / ((4 + kgroups) * rows_per_stage * 128) < 3) rows_per_stage = 32;
  const int stage_bytes = (4 + kgroups) * rows_per_stage * 128;
  int stages = (kSmemBudget - 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  DEMF_REQUIRE(stages >= 2, DEMF_E_UNSUPPORTED);
  p.stages = stages;
  p.rows_per_stage = rows_per_stage;
  const int mblocks = (N + 127) / 128;
  // slabs: enough CTAs to fill the GPU, at least 256 rows each
  long slabs = (2L * kNumSMs) / mblocks;
  if (slabs < 1) slabs = 1;
  long rows_per_slab = (R + slabs - 1) / slabs;
  if (rows_per_slab < 256) rows_per_slab = 256;
  rows_per_slab = (rows_per_slab + rows_per_stage - 1) / rows_per_stage * rows_per_stage;
  slabs = (R + rows_per_slab - 1) / rows_per_slab;
  p.rows_per_slab = rows_per_slab;
  CUtensorMap mdy, mx;
  if (int rc = make_map(&mdy, dy, R, N, lddy, rows_per_stage, true)) return rc;
  if (int rc = make_map(&mx, x, R, K, ldx, rows_per_stage, true)) return rc;
  p.mn.sbo = g_mn_sbo;
  p.mn.layout = g_mn_layout;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 1024;
  wgrad_kernel<<<dim3((unsigned)slabs, mblocks), kWgradThreads, smem, as_stream(stream)>>>(mdy, mx, p);
  return after_launch("wgrad_kernel");
}

int demf_gemm_rows_dgrad_bn(const float* dy, long lddy, const float* w, long ldw, long R, int N, int K,
                            const float* y_prev, long ldy_prev, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, void* bn_state, float* g, long ldg, void* stream) {
  DEMF_REQUIRE_PTR(dy);
  DEMF_REQUIRE_PTR(w);
  DEMF_REQUIRE_PTR(y_prev);
  DEMF_REQUIRE_PTR(mean);
  DEMF_REQUIRE_PTR(invstd);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(bn_state);
  DEMF_REQUIRE_PTR(g);
  DEMF_REQUIRE(R > 0 && K > 0 && N > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(K <= 256, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(lddy % 4 == 0 && ldw % 4 == 0 && ldg % 4 == 0 && ldy_prev % 4 == 0 && lddy >= N && ldw >= K &&
                   ldg >= K && ldy_prev >= K,
               DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(dy) && al16(w) && al16(g) && al16(y_prev), DEMF_E_UNSUPPORTED);
  BnMask bn;
  bn.y = y_prev;
  bn.ldy = ldy_prev;
  bn.mean = mean;
  bn.invstd = invstd;
  bn.gamma = gamma;
  bn.beta = beta;
  return launch_rows_gemm<true>(dy, lddy, w, ldw, nullptr, static_cast<double*>(bn_state), g, ldg, R, N, K, 0,
                                as_stream(stream), &bn);
}

int demf_bn_bwd_finalize(void* state, long R, int C, const float* mean, const float* invstd, float* grad_gamma,
                         float* grad_beta, float* coef, void* stream) {
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(mean);
  DEMF_REQUIRE_PTR(invstd);
  DEMF_REQUIRE_PTR(grad_gamma);
  DEMF_REQUIRE_PTR(grad_beta);
  DEMF_REQUIRE_PTR(coef);
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  bn_bwd_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(static_cast<double*>(state), R, C, mean, invstd, grad_gamma,
                                                          grad_beta, coef);
  return after_launch("bn_bwd_finalize_kernel");
}

int demf_bn_finalize(void* state, long R, int C, float eps, float momentum, float* save_mean, float* save_invstd,
                     float* running_mean, float* running_var, void* stream) {
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(save_mean);
  DEMF_REQUIRE_PTR(save_invstd);
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  bn_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(static_cast<double*>(state), R, C, eps, momentum, save_mean,
                                                      save_invstd, running_mean, running_var);
  return after_launch("bn_finalize_kernel");
}

}  // extern "C"
