// Stand-alone check of tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form), in the conventions of
// demf_b200/csrc/umma.cuh:   D(128 x N) = A(128 x K) * W(N x K)^T,  kind::tf32,
// A written to TMEM columns by the threads that own its rows (tcgen05.st 32x32b: thread t of warp w owns
// row 32*w + t, column k holds A[row][k]) -- exactly what an epilogue that keeps the activation on chip
// would do (tcgen05.ld accumulator -> bias/ReLU -> tcgen05.st into the next layer's A columns) -- and
// W staged in shared memory as SWIZZLE_128B K-major chunks as everywhere else in the library.
// this removes the activation store to and the activation read from shared
// memory, the largest share of the fused set-abstraction kernel's smem traffic.
//
// Verified on B200 at the end of round 1 (ALL PASS for N in {64,128,256}, K in {32,64,128}; errors are those of
// TF32 operand truncation): A lives in TMEM as lanes = rows, ONE COLUMN PER k, a K step of 8 tf32 = 8 columns.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build_tmp/umma_ts_probe tools/umma_ts_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../demf_b200/csrc/umma.cuh"

using namespace demf::umma;

__global__ void __launch_bounds__(128) probe_ts_kernel(const float* __restrict__ A, const float* __restrict__ W, int N,
                                                       int K, float* __restrict__ D, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int chunks = K / 32;
  unsigned char* w_s = base;                                        // chunks x (N rows x 128 B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + chunks * N * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t mma_bar = smem_u32(&bars[0]);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int c = 0; c < chunks; ++c)
    for (int e = tid; e < N * 8; e += 128) {
      const int n = e >> 3, j = e & 7;
      *reinterpret_cast<float4*>(w_s + c * N * 128 + sw128_offset(n, j)) =
          *reinterpret_cast<const float4*>(W + (size_t)n * K + c * 32 + j * 4);
    }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc = tmem;             // columns [0, N): the accumulator
  const uint32_t act = tmem + 256;       // columns [256, 256 + K): A, one column per k
  // every thread stores ITS row of A (row = tid) into the TMEM lanes of its warp's quarter
  for (int k0 = 0; k0 < K; k0 += 32) {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(A[(size_t)tid * K + k0 + i]);
    tmem_st32(act + ((uint32_t)(warp * 32) << 16) + k0, v);
  }
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = instr_desc_tf32(128, N);
    for (int c = 0; c < chunks; ++c)
      for (int k = 0; k < 4; ++k)   // a K step of 8 tf32 = 8 TMEM columns of A, +32 bytes of the W chunk
        mma_tf32_ts(acc, act + c * 32 + k * 8, smem_desc_sw128(smem_u32(w_s) + c * N * 128) + 2 * k, idesc,
                    (c | k) != 0);
    mma_commit(mma_bar);
  }
  const bool ok = mbar_wait(mma_bar, 0);
  if (!ok) atomicExch(err, 1);
  tc_fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 32) {
    uint32_t v[32];
    tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
    float* out = D + (size_t)tid * N + n0;
    for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 512);
}

static bool run(int N, int K) {
  std::vector<float> A(128 * (size_t)K), W((size_t)N * K);
  srand(N * 1000 + K);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : W) x = (rand() % 2001 - 1000) / 1000.f;
  float *dA, *dW, *dD;
  int* dErr;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dD, 128 * (size_t)N * 4);
  cudaMalloc(&dErr, 4);
  cudaMemset(dErr, 0, 4);
  cudaMemset(dD, 0xff, 128 * (size_t)N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + (size_t)(K / 32) * N * 128 + 64;
  cudaFuncSetAttribute(probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_ts_kernel<<<1, 128, smem>>>(dA, dW, N, K, dD, dErr);
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
  printf("A-in-TMEM N=%3d K=%3d  cuda=%s timeout=%d  max|err|=%.3e (max|ref|=%.2f)  %s\n", N, K, cudaGetErrorName(e),
         err, maxerr, maxref, pass ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dErr);
  return pass;
}

int main() {
  bool ok = true;
  const int cases[][2] = {{128, 32}, {128, 128}, {256, 128}, {64, 64}};
  for (auto& c : cases) ok = run(c[0], c[1]) && ok;
  printf(ok ? "umma .ts probe: ALL PASS\n" : "umma .ts probe: FAILURES\n");
  return ok ? 0 : 1;
}
