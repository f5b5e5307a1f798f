// Stand-alone check of 2-CTA (cta_group::2) tcgen05 MMA on a real B200, in the operand convention of
// demf_b200/csrc/umma.cuh: a CTA pair computes
//   D(256 x N) = A(256 x K) * W(N x K)^T,   kind::tf32,
// each CTA staging ITS 128 rows of A and ITS HALF (N/2 rows) of W in its own shared memory
// (SWIZZLE_128B, K-major); the leader CTA issues the MMAs, the commit is multicast to both CTAs, and
// each CTA reads its 128 accumulator rows from its own TMEM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build_tmp/umma2_probe tools/umma2_probe.cu
#include <cooperative_groups.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../demf_b200/csrc/umma.cuh"

namespace cg = cooperative_groups;
using namespace demf::umma;

__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t columns) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(columns)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free2(uint32_t taddr, uint32_t columns) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(columns) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
    probe2_kernel(const float* __restrict__ A, const float* __restrict__ W, int N, int K, float* __restrict__ D,
                  int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = K / 32;
  unsigned char* a_s = base;                          // chunks x (128 rows x 128 B)
  unsigned char* w_s = base + chunks * 128 * 128;     // chunks x (N/2 rows x 128 B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + chunks * (N / 2) * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t mma_bar = smem_u32(&bars[0]);
  const int tid = threadIdx.x, warp = tid >> 5;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc2(smem_u32(tmem_slot), 256);
  // this CTA's rows of A and its half of W
  for (int c = 0; c < chunks; ++c) {
    for (int e = tid; e < 128 * 8; e += 128) {
      const int r = e >> 3, j = e & 7;
      *reinterpret_cast<float4*>(a_s + c * 128 * 128 + sw128_offset(r, j)) =
          *reinterpret_cast<const float4*>(A + (size_t)(rank * 128 + r) * K + c * 32 + j * 4);
    }
    for (int e = tid; e < (N / 2) * 8; e += 128) {
      const int n = e >> 3, j = e & 7;
      *reinterpret_cast<float4*>(w_s + c * (N / 2) * 128 + sw128_offset(n, j)) =
          *reinterpret_cast<const float4*>(W + (size_t)(rank * (N / 2) + n) * K + c * 32 + j * 4);
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  cluster.sync();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = instr_desc_tf32(256, N);
    for (int c = 0; c < chunks; ++c)
      for (int k = 0; k < 4; ++k)
        mma2_tf32(tmem, smem_desc_sw128(smem_u32(a_s) + c * 128 * 128) + 2 * k,
                  smem_desc_sw128(smem_u32(w_s) + c * (N / 2) * 128) + 2 * k, idesc, (c | k) != 0);
    mma2_commit_multicast(mma_bar, 3);
  }
  const bool ok = mbar_wait(mma_bar, 0);
  if (!ok) atomicExch(err, 1);
  tc_fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
    float* out = D + (size_t)(rank * 128 + tid) * N + n0;
    for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  cluster.sync();
  if (warp == 0) tmem_free2(tmem, 256);
}

static bool run(int N, int K) {
  std::vector<float> A(256 * (size_t)K), W((size_t)N * K);
  srand(N * 1000 + K);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : W) x = (rand() % 2001 - 1000) / 1000.f;
  float *dA, *dW, *dD;
  int* dErr;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dD, 256 * (size_t)N * 4);
  cudaMalloc(&dErr, 4);
  cudaMemset(dErr, 0, 4);
  cudaMemset(dD, 0xff, 256 * (size_t)N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const int chunks = K / 32;
  const size_t smem = 1024 + (size_t)chunks * (128 * 128 + (N / 2) * 128) + 64;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe2_kernel<<<2, 128, smem>>>(dA, dW, N, K, dD, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(256 * (size_t)N);
  int err = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)r * K + k] * W[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(s - D[(size_t)r * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  const bool pass = e == cudaSuccess && err == 0 && maxerr <= 2e-3 * maxref + 1e-3;
  printf("2-CTA N=%3d K=%3d  cuda=%s timeout=%d  max|err|=%.3e (max|ref|=%.2f)  %s\n", N, K, cudaGetErrorName(e),
         err, maxerr, maxref, pass ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dErr);
  return pass;
}

int main() {
  bool ok = true;
  const int cases[][2] = {{128, 32}, {128, 128}, {256, 64}, {64, 32}};
  for (auto& c : cases) ok = run(c[0], c[1]) && ok;
  printf(ok ? "umma2 probe: ALL PASS\n" : "umma2 probe: FAILURES\n");
  return ok ? 0 : 1;
}
