// Exact ball query through a uniform grid (the same index rows as the brute-force scan).
//
// Upstream's ball query (mmdet3d ball_query_kernel) tests every centre against every point:
// 2048 x 20000 distance tests per scene at the first set-abstraction level, of which ~100 hit.
// Here the cloud of one scene is binned once into a <= 32x32x16 grid with cell edge >= radius
// (demf_ball_grid_build: counting sort, one CTA per scene), and a centre only tests the points of
// its 3x3x3 cell neighbourhood -- ~50-300 candidates instead of 20000. The neighbourhood is a
// superset of the ball, the membership test is the same fp32 expression, and the result row is
// rebuilt from the hit SET: "the first nsample hits in index order, padded with the first hit"
// = the nsample smallest hit indices in ascending order. Hence bit-identical rows.
//
// Selection without a sort: hits are appended to a per-warp buffer (cap kGridCap) while a 64-bucket
// histogram over the index range is kept; if the ball holds more hits than the buffer, the
// histogram gives an index bound T below which at least nsample (and at most cap) hits lie and the
// candidates are re-tested keeping only indices < T. Ranks come from counting smaller elements.
// If even that overflows (more than cap hits below T) the caller falls back to the full scan.
#pragma once
#include "common.cuh"

namespace demf {

constexpr int kGridX = 32, kGridY = 32, kGridZ = 16;
constexpr int kGridCells = kGridX * kGridY * kGridZ;  // 16384
constexpr int kGridCap = 256;                         // hit buffer entries per warp
constexpr int kGridHist = 64;

struct BallGridHeader {  // 32 bytes, one per scene
  float x0, y0, z0, inv_h;
  int nx, ny, nz;
  float radius;  // the grid answers queries with max_radius <= radius
};

// workspace layout per scene (all 16-byte aligned):
//   [BallGridHeader][cell_start: kGridCells+4 ints][sorted: N float4 = (x, y, z, index bits)]
__host__ __device__ inline size_t ball_grid_scene_bytes(int N) {
  return sizeof(BallGridHeader) + (size_t)(kGridCells + 4) * 4 + (size_t)N * 16;
}

#ifdef __CUDACC__
struct BallGridView {
  BallGridHeader h;
  const int* cell_start;
  const float4* sorted;
};

__device__ __forceinline__ BallGridView ball_grid_view(const void* workspace, int b, int N) {
  const unsigned char* base = static_cast<const unsigned char*>(workspace) + (size_t)b * ball_grid_scene_bytes(N);
  BallGridView v;
  const int4* hp = reinterpret_cast<const int4*>(base);
  const int4 a = __ldg(hp), c = __ldg(hp + 1);
  v.h.x0 = __int_as_float(a.x);
  v.h.y0 = __int_as_float(a.y);
  v.h.z0 = __int_as_float(a.z);
  v.h.inv_h = __int_as_float(a.w);
  v.h.nx = c.x;
  v.h.ny = c.y;
  v.h.nz = c.z;
  v.h.radius = __int_as_float(c.w);
  v.cell_start = reinterpret_cast<const int*>(base + sizeof(BallGridHeader));
  v.sorted = reinterpret_cast<const float4*>(base + sizeof(BallGridHeader) + (size_t)(kGridCells + 4) * 4);
  return v;
}

// Monotone in v: the same expression bins the points (build) and locates the centres (query).
__device__ __forceinline__ int ball_grid_coord(float v, float v0, float inv_h) {
  return __float2int_rd(__fmul_rn(__fsub_rn(v, v0), inv_h));
}

// Smallest index bound (a multiple of the histogram bucket width) below which at least `ns` of the
// histogrammed hits lie; *below = how many exactly. Returns 0 when fewer than ns hits were counted.
__device__ __forceinline__ int hist_bound(const int* hist, int ns, int shift, unsigned lane, int* below = nullptr) {
  const int h0 = hist[2 * lane], h1 = hist[2 * lane + 1];
  int incl = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  const int excl = incl - h0 - h1;
  int cand = kGridHist, cnt = 0;  // first bucket whose inclusive prefix reaches ns
  if (excl + h0 >= ns && excl < ns) {
    cand = 2 * lane;
    cnt = excl + h0;
  } else if (incl >= ns && excl + h0 < ns) {
    cand = 2 * lane + 1;
    cnt = incl;
  }
  const unsigned who = __ballot_sync(0xffffffffu, cand < kGridHist);
  if (!who) return 0;
  const int src = __ffs(who) - 1;
  cand = __shfl_sync(0xffffffffu, cand, src);
  cnt = __shfl_sync(0xffffffffu, cnt, src);
  if (below) *below = cnt;
  return (cand + 1) << shift;
}

// One warp, one centre. Writes the full index row (ns entries) to `row` and returns the number of
// distinct hits kept (min(hits, ns)), or -1 when the caller must fall back to the full scan.
// buf: kGridCap ints, hist: kGridHist ints -- per-warp scratch (shared memory).
__device__ __forceinline__ int ball_grid_query_warp(const BallGridView& g, int N, float cx, float cy, float cz,
                                                    float min_r2, float max_r2, int ns, int32_t* row,
                                                    int* buf, int* hist, unsigned lane) {
  const int ix = ball_grid_coord(cx, g.h.x0, g.h.inv_h);
  const int iy = ball_grid_coord(cy, g.h.y0, g.h.inv_h);
  const int iz = ball_grid_coord(cz, g.h.z0, g.h.inv_h);
  const int xlo = max(ix - 1, 0), xhi = min(ix + 1, g.h.nx - 1);
  const int ylo = max(iy - 1, 0), yhi = min(iy + 1, g.h.ny - 1);
  const int zlo = max(iz - 1, 0), zhi = min(iz + 1, g.h.nz - 1);
  int shift = 0;
  while (((N - 1) >> shift) >= kGridHist) ++shift;

  // The 3x3 cell columns (contiguous in z) of the neighbourhood as one flat candidate range: lane
  // l < 9 fetches column l's [begin, end) -- one round trip instead of nine dependent ones -- and
  // every lane keeps the nine (offset, begin) pairs, so that candidate t of the flat range maps to
  // sorted[begin_c + t - offset_c] and consecutive lanes test consecutive candidates.
  int cbeg = 0, clen = 0;
  if (lane < 9 && zlo <= zhi) {
    const int x = ix - 1 + (int)lane / 3, y = iy - 1 + (int)lane % 3;
    if (x >= 0 && x < g.h.nx && y >= 0 && y < g.h.ny) {
      const int c0 = (x * kGridY + y) * kGridZ;
      cbeg = __ldg(g.cell_start + c0 + zlo);
      clen = __ldg(g.cell_start + c0 + zhi + 1) - cbeg;
    }
  }
  int incl = clen;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 8);
  int coff[9], cb[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    coff[c] = __shfl_sync(0xffffffffu, incl - clen, c);
    cb[c] = __shfl_sync(0xffffffffu, cbeg, c);
  }
  (void)xlo; (void)xhi; (void)ylo; (void)yhi;

  int bound = 0x7fffffff;  // keep hits with index < bound
  for (int pass = 0; pass < 2; ++pass) {
    for (int l = lane; l < kGridHist; l += 32) hist[l] = 0;
    __syncwarp();
    int cnt = 0;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + (int)lane;
      bool hit = false;
      int k = 0;
      if (t < total) {
        int q = cb[0] + t;
#pragma unroll
        for (int c = 1; c < 9; ++c)
          if (t >= coff[c]) q = cb[c] + (t - coff[c]);
        const float4 p = __ldg(g.sorted + q);
        k = __float_as_int(p.w);
        const float d2 = sqdist(cx, cy, cz, p.x, p.y, p.z);
        hit = ((d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2)) && k < bound;
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, hit);
      if (ballot) {
        const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
        if (hit) {
          if (pos < kGridCap) buf[pos] = k;
          atomicAdd(&hist[k >> shift], 1);
        }
        cnt += __popc(ballot);
      }
    }
    __syncwarp();
    if (cnt <= kGridCap) {
      int n = cnt;
      if (cnt > ns + 32) {
        // many more hits than wanted: the histogram gives an index bound below which >= ns of them
        // lie; only those compete for the ns smallest (ranking is quadratic in the candidates)
        const int lim = hist_bound(hist, ns, shift, lane);
        if (lim > 0) {
          n = 0;
          for (int e0 = 0; e0 < cnt; e0 += 32) {
            const int e = e0 + (int)lane;
            const int v = e < cnt ? buf[e] : 0x7fffffff;
            const unsigned keep = __ballot_sync(0xffffffffu, v < lim);
            __syncwarp();
            if (v < lim) buf[n + __popc(keep & ((1u << lane) - 1u))] = v;   // n <= e0: never ahead of the reads
            n += __popc(keep);
            __syncwarp();
          }
        }
      }
      // rank = number of kept hits with a smaller index (indices are distinct)
      for (int e = lane; e < n; e += 32) {
        const int v = buf[e];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += buf[j] < v;
        if (rank < ns) row[rank] = v;
      }
      __syncwarp();
      const int kept = min(cnt, ns);
      const int first = cnt > 0 ? row[0] : 0;
      for (int l = kept + lane; l < ns; l += 32) row[l] = first;
      __syncwarp();
      return kept;
    }
    if (pass == 1) return -1;
    // more hits than the buffer holds: smallest index bound with >= ns hits below it
    int below = 0;
    const int lim = hist_bound(hist, ns, shift, lane, &below);
    if (lim <= 0 || below > kGridCap) return -1;
    bound = lim;
    __syncwarp();
    continue;
  }
  return -1;
}

// The same query with the selection done through a BITMAP over the index range (the stand-alone launch, where a
// warp can afford N/8 bytes of shared memory): every hit sets bit k of `bm`; the row -- the ns smallest hit indices in
// ascending order -- is then read off the bitmap with a popcount prefix over the lanes' word ranges. One pass over
// the candidates, no hit buffer, no histogram, no O(hits^2) ranking (31 % of the instructions of the buffer form at
// the first level), and no overflow case: any number of hits is handled. `bm`: `ball_bitmap_words(N)` words.
__host__ __device__ inline int ball_bitmap_lane_words(int N) { return ((((N + 31) >> 5) + 31) / 32 + 3) & ~3; }
__host__ __device__ inline int ball_bitmap_words(int N) { return 32 * ball_bitmap_lane_words(N); }

__device__ __forceinline__ int ball_grid_query_bitmap_warp(const BallGridView& g, int N, float cx, float cy, float cz,
                                                           float min_r2, float max_r2, int ns, int32_t* row,
                                                           unsigned* bm, unsigned lane) {
  const int ix = ball_grid_coord(cx, g.h.x0, g.h.inv_h);
  const int iy = ball_grid_coord(cy, g.h.y0, g.h.inv_h);
  const int iz = ball_grid_coord(cz, g.h.z0, g.h.inv_h);
  const int zlo = max(iz - 1, 0), zhi = min(iz + 1, g.h.nz - 1);
  const int per = ball_bitmap_lane_words(N);          // words per lane, a multiple of 4
  uint4* mine = reinterpret_cast<uint4*>(bm + lane * per);
  for (int i = 0; i < per / 4; ++i) mine[i] = make_uint4(0u, 0u, 0u, 0u);
  // the 3x3 cell columns as one flat candidate range (see ball_grid_query_warp)
  int cbeg = 0, clen = 0;
  if (lane < 9 && zlo <= zhi) {
    const int x = ix - 1 + (int)lane / 3, y = iy - 1 + (int)lane % 3;
    if (x >= 0 && x < g.h.nx && y >= 0 && y < g.h.ny) {
      const int c0 = (x * kGridY + y) * kGridZ;
      cbeg = __ldg(g.cell_start + c0 + zlo);
      clen = __ldg(g.cell_start + c0 + zhi + 1) - cbeg;
    }
  }
  int incl = clen;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 8);
  int coff[9], cb[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    coff[c] = __shfl_sync(0xffffffffu, incl - clen, c);
    cb[c] = __shfl_sync(0xffffffffu, cbeg, c) - coff[c];   // candidate t of column c is sorted[cb[c] + t]
  }
  __syncwarp();
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + (int)lane;
    if (t < total) {
      int base = cb[0];
#pragma unroll
      for (int c = 1; c < 9; ++c)
        if (t >= coff[c]) base = cb[c];
      const float4 p = __ldg(g.sorted + base + t);
      const int k = __float_as_int(p.w);
      const float d2 = sqdist(cx, cy, cz, p.x, p.y, p.z);
      if ((d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2)) atomicOr(bm + (k >> 5), 1u << (k & 31));
    }
  }
  __syncwarp();
  // popcount prefix over the lanes' word ranges, then every lane writes the indices of its set bits
  int mycount = 0;
  for (int i = 0; i < per / 4; ++i) {
    const uint4 v = mine[i];
    mycount += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
  }
  int upto = mycount;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, upto, o);
    if ((int)lane >= o) upto += t;
  }
  const int cnt = __shfl_sync(0xffffffffu, upto, 31);
  int pos = upto - mycount;
  if (mycount > 0 && pos < ns) {
    const unsigned* w = bm + lane * per;
    for (int i = 0; i < per && pos < ns; ++i) {
      unsigned v = w[i];
      while (v != 0u && pos < ns) {
        const int bit = __ffs(v) - 1;
        v &= v - 1u;
        row[pos++] = ((int)lane * per + i) * 32 + bit;
      }
    }
  }
  __syncwarp();
  const int kept = min(cnt, ns);
  const int first = cnt > 0 ? row[0] : 0;
  for (int l = kept + lane; l < ns; l += 32) row[l] = first;
  __syncwarp();
  return kept;
}

// Full scan of the cloud by one warp straight from global memory (rare fallback of the grid path).
__device__ __forceinline__ int ball_scan_warp(const float* __restrict__ cloud, int N, float cx, float cy,
                                              float cz, float min_r2, float max_r2, int ns, int32_t* row,
                                              unsigned lane) {
  int cnt = 0, first = 0;
  for (int j = 0; j < N; j += 32) {
    const int q = j + lane;
    bool hit = false;
    if (q < N) {
      const float d2 = sqdist(cx, cy, cz, __ldg(cloud + q * 3L), __ldg(cloud + q * 3L + 1), __ldg(cloud + q * 3L + 2));
      hit = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (ballot) {
      if (cnt == 0) first = j + (__ffs(ballot) - 1);
      const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
      if (hit && pos < ns) row[pos] = q;
      cnt += __popc(ballot);
      if (cnt >= ns) break;
    }
  }
  __syncwarp();
  const int kept = min(cnt, ns);
  for (int l = kept + lane; l < ns; l += 32) row[l] = first;
  __syncwarp();
  return kept;
}

// Grid query with the fallback folded in.
__device__ __forceinline__ int ball_query_warp(const BallGridView& g, const float* __restrict__ cloud, int N,
                                               float cx, float cy, float cz, float min_r2, float max_r2,
                                               int ns, int32_t* row, int* buf, int* hist, unsigned lane) {
  int kept = -1;
  if (max_r2 <= g.h.radius * g.h.radius)
    kept = ball_grid_query_warp(g, N, cx, cy, cz, min_r2, max_r2, ns, row, buf, hist, lane);
  if (kept < 0) kept = ball_scan_warp(cloud, N, cx, cy, cz, min_r2, max_r2, ns, row, lane);
  return kept;
}
#endif  // __CUDACC__

// ball_grid.cu: the stand-alone grid query launch (also used by the fused set-abstraction entry point)
int launch_ball_query_grid(const float* xyz, const float* new_xyz, const void* grid, int B, int N, int M,
                           float min_radius, float max_radius, int ns, int32_t* idx, cudaStream_t stream);

}  // namespace demf
