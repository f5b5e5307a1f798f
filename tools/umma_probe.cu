// Stand-alone check of the tcgen05 plumbing in demf_b200/csrc/umma.cuh on a real B200:
//   D(128 x N) = A(128 x K) * W(N x K)^T   with kind::tf32, SWIZZLE_128B K-major operands,
// A staged by generic stores, W by 1-D bulk copies of a host-packed swizzled image, accumulator
// read back with tcgen05.ld 32x32b. Compared against a double-precision host product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe tools/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../demf_b200/csrc/umma.cuh"

using namespace demf::umma;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Wp,
                                                    int N, int K, float* __restrict__ D, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [A chunk 16 KB][W chunk N*128][barriers]
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_s = base;
  unsigned char* w_s = base + 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + N * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t w_bar = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[1]);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(w_bar, 1);
    mbar_init(mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = instr_desc_tf32(128, N);

  const int chunks = (K + 31) / 32;
  uint32_t phase = 0;
  bool ok = true;
  for (int c = 0; c < chunks; ++c) {
    const int kc = min(32, K - c * 32);  // floats in this chunk (multiple of 8)
    if (tid == 0) {
      mbar_expect_tx(w_bar, N * 128);
      bulk_g2s(smem_u32(w_s), Wp + (size_t)c * N * 32, N * 128, w_bar);
    }
    // A chunk: 128 rows x kc floats; thread -> (row, slot) with slots of one row on adjacent threads
    for (int e = tid; e < 128 * 8; e += 128) {
      const int r = e >> 3, j = e & 7;
      if (j * 4 < kc) {
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + c * 32 + j * 4);
        *reinterpret_cast<float4*>(a_s + sw128_offset(r, j)) = v;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      ok = mbar_wait(w_bar, phase);
      tc_fence_after_sync();
      for (int k = 0; k < kc / 8; ++k) {
        mma_tf32(tmem, smem_desc_sw128(smem_u32(a_s) + k * 32), smem_desc_sw128(smem_u32(w_s) + k * 32),
                 idesc, (c | k) != 0);
      }
      mma_commit(mma_bar);
    }
    ok = mbar_wait(mma_bar, phase) && ok;  // single stage: wait before overwriting the operands
    phase ^= 1;
    if (!ok) break;
  }
  if (!ok) atomicExch(err, 1);
  tc_fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
    float* out = D + (size_t)tid * N + n0;
    for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 256);
}


// ---- timing: cycles from the first tcgen05.mma issue to the commit mbarrier completing, for `reps`
// back-to-back (128 x N x 8) tf32 MMAs on resident operands (measures issue cost + pipe latency)
__global__ void __launch_bounds__(128) mma_time_kernel(int N, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_s = base;
  unsigned char* w_s = base + 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + 256 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t mma_bar = smem_u32(&bars[1]);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 128 + 256 * 128) / 4; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f;
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = instr_desc_tf32(128, N);
  if (tid == 0) {
    uint32_t phase = 0;
    for (int trial = 0; trial < 4; ++trial) {
      const long long t0 = clock64();
      for (int k = 0; k < reps; ++k)
        mma_tf32(tmem, smem_desc_sw128(smem_u32(a_s) + (k & 3) * 32), smem_desc_sw128(smem_u32(w_s) + (k & 3) * 32),
                 idesc, k != 0);
      const long long t1 = clock64();
      mma_commit(mma_bar);
      mbar_wait(mma_bar, phase);
      const long long t2 = clock64();
      phase ^= 1;
      out[trial * 2] = t1 - t0;
      out[trial * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 256);
}

static void time_mma() {
  long long* d;
  cudaMalloc(&d, 64);
  const size_t smem = 1024 + 128 * 128 + 256 * 128 + 64;
  cudaFuncSetAttribute(mma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int Ns[] = {64, 128, 256};
  const int Rs[] = {1, 4, 8, 16, 32, 64};
  for (int N : Ns)
    for (int R : Rs) {
      mma_time_kernel<<<1, 128, smem>>>(N, R, d);
      cudaDeviceSynchronize();
      long long h[8];
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      printf("mma 128x%dx8 tf32 x%2d: issue %lld clk, issue->done %lld clk (last trial; first trial %lld)\n", N, R, h[6],
             h[7], h[1]);
    }
  cudaFree(d);
}

static bool run(int N, int K) {
  std::vector<float> A(128 * (size_t)K), W((size_t)N * K);
  srand(N * 1000 + K);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : W) x = (rand() % 2001 - 1000) / 1000.f;
  const int chunks = (K + 31) / 32;
  std::vector<float> Wp((size_t)chunks * N * 32, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const int c = k / 32, kk = k % 32;
      const size_t byte = (size_t)c * N * 128 + sw128_offset(n, kk / 4) + (kk % 4) * 4;
      Wp[byte / 4] = W[(size_t)n * K + k];
    }
  float *dA, *dW, *dD;
  int* dErr;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dW, Wp.size() * 4);
  cudaMalloc(&dD, 128 * (size_t)N * 4);
  cudaMalloc(&dErr, 4);
  cudaMemset(dErr, 0, 4);
  cudaMemset(dD, 0xff, 128 * (size_t)N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, Wp.data(), Wp.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + 128 * 128 + (size_t)N * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dW, N, K, dD, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * (size_t)N);
  int err = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)r * K + k] * W[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(s - D[(size_t)r * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  const bool pass = e == cudaSuccess && err == 0 && maxerr <= 2e-3 * maxref + 1e-3;
  printf("N=%3d K=%3d  cuda=%s timeout=%d  max|err|=%.3e (max|ref|=%.2f)  %s\n", N, K,
         cudaGetErrorName(e), err, maxerr, maxref, pass ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dErr);
  return pass;
}

int main() {
  bool ok = true;
  const int cases[][2] = {{64, 8}, {64, 32}, {128, 64}, {128, 136}, {256, 128}, {256, 264}, {32, 40}};
  for (auto& c : cases) ok = run(c[0], c[1]) && ok;
  printf(ok ? "umma probe: ALL PASS\n" : "umma probe: FAILURES\n");
  time_mma();
  return ok ? 0 : 1;
}
