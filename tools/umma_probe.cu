// Bring-up probe for the tcgen05 primitives in clair_b200/csrc/tc_common.cuh.
// D[128 x N] = A[128 x K] * B[N x K]^T, fp16 operands (K-major, no-swizzle core-matrix layout), fp32 accumulate in TMEM.
// mode 0: LBO = k-chunk stride, SBO = 8-row-group stride (as documented);  mode 1: the two swapped.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../clair_b200/csrc/tc_common.cuh"
using namespace clairb::tc;

constexpr int M = 128, N = 256, K = 64;

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ Ag, const __half* __restrict__ Bg, float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* As = (__half*)smem;                       // [K/8][128][8]
  __half* Bs = (__half*)(smem + M * K * 2);         // [K/8][256][8]
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, (M + N) * K * 2);
    bulk_g2s(As, Ag, M * K * 2, &bar_load);
    bulk_g2s(Bs, Bg, N * K * 2, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(M, N);
    for (int j = 0; j < K / 16; ++j) {
      uint32_t a_addr = smem_u32(As) + j * 2 * (M * 16);
      uint32_t b_addr = smem_u32(Bs) + j * 2 * (N * 16);
      uint64_t ad = mode == 0 ? make_smem_desc(a_addr, M * 16, 128) : make_smem_desc(a_addr, 128, M * 16);
      uint64_t bd = mode == 0 ? make_smem_desc(b_addr, N * 16, 128) : make_smem_desc(b_addr, 128, N * 16);
      umma_f16(tmem, ad, bd, idesc, j > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int main() {
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  std::vector<__half> Ah(M * K), Bh(N * K);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { __half h = __float2half(A[r * K + k]); A[r * K + k] = __half2float(h); Ah[(k / 8) * M * 8 + r * 8 + k % 8] = h; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { __half h = __float2half(B[n * K + k]); B[n * K + k] = __half2float(h); Bh[(k / 8) * N * 8 + n * 8 + k % 8] = h; }
  std::vector<float> ref(M * N);
  for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * B[n * K + k]; ref[r * N + n] = (float)s; }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bh.size() * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  size_t smem = (M + N) * K * 2 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t st = cudaDeviceSynchronize();
    std::vector<float> D(M * N);
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs(D[i] - ref[i]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad; }
    printf("mode %d: %s  max|err| %.3e  mismatches %d / %d   D[0]=%f ref[0]=%f D[5*N+7]=%f ref=%f\n", mode, cudaGetErrorString(st), maxerr, bad, M * N, D[0], ref[0], D[5 * N + 7], ref[5 * N + 7]);
    if (st != cudaSuccess) return 1;
  }
  return 0;
}
