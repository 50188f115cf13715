// Probe: cta_group::2 MMA with the A operand in tensor memory (written by tcgen05.st), M=256, N=128, K=64.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../clair_b200/csrc/tc_common.cuh"
using namespace clairb::tc;
constexpr int PM = 256, PN = 128, PK = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe_ts(const __half* __restrict__ Ag /*[256][64] row-major*/, const __half* __restrict__ Bg, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* Bs = (__half*)smem;                        // this CTA's 64 rows (N) of B: [K/8][64][8]
  __shared__ uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_peer, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_pair<256>(&tmem_base);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  // every thread stores its A row (64 halves = 32 words) into TMEM columns 128..159
  {
    const int row = rank * 128 + warp * 32 + lane;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(Ag + (size_t)row * PK);
    uint32_t w[32];
    for (int i = 0; i < 32; ++i) w[i] = src[i];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 128;
    tmem_st16(ta, w);
    tmem_st16(ta + 16, w + 16);
    tmem_st_wait();
  }
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, 64 * PK * 2);
    bulk_g2s(Bs, Bg + (size_t)rank * 64 * PK, 64 * PK * 2, &bar_load);
    mbar_wait(&bar_load, 0);
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs: A in TMEM, B in smem
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(PM, PN);
    for (int j = 0; j < PK / 16; ++j) {
      uint64_t bd = make_smem_desc(smem_u32(Bs) + j * 2 * (64 * 16), 64 * 16, 128);
      umma_f16_pair_ts(tmem, tmem + 128 + j * 8, bd, idesc, j > 0);
    }
    umma_commit_pair(&bar_mma, 0b11);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = rank * 128 + warp * 32 + lane;
  for (int c = 0; c < PN; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * PN + c + i] = v[i];
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<256>(tmem);
}

int main() {
  std::vector<float> A(PM * PK), B(PN * PK);
  srand(4);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  std::vector<__half> Ah(PM * PK), Bh(PN * PK);
  for (int i = 0; i < PM * PK; ++i) { Ah[i] = __float2half(A[i]); A[i] = __half2float(Ah[i]); }
  for (int n = 0; n < PN; ++n)
    for (int k = 0; k < PK; ++k) {
      __half h = __float2half(B[n * PK + k]);
      B[n * PK + k] = __half2float(h);
      Bh[(size_t)(n / 64) * 64 * PK + (k / 8) * 64 * 8 + (n % 64) * 8 + k % 8] = h;
    }
  std::vector<float> ref(PM * PN);
  for (int r = 0; r < PM; ++r)
    for (int n = 0; n < PN; ++n) {
      double s = 0;
      for (int k = 0; k < PK; ++k) s += (double)A[r * PK + k] * B[n * PK + k];
      ref[r * PN + n] = (float)s;
    }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bh.size() * 2); cudaMalloc(&dD, PM * PN * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, PM * PN * 4);
  size_t smem = 64 * PK * 2 + 1024;
  probe_ts<<<2, 128, smem>>>(dA, dB, dD);
  cudaError_t st = cudaDeviceSynchronize();
  std::vector<float> Dh(PM * PN);
  cudaMemcpy(Dh.data(), dD, PM * PN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < PM * PN; ++i) { double e = fabs(Dh[i] - ref[i]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad; }
  printf("ts: %s  max|err| %.3e  mismatches %d / %d  D[0]=%f ref=%f D[200*N+70]=%f ref=%f\n", cudaGetErrorString(st), maxerr, bad,
         PM * PN, Dh[0], ref[0], Dh[200 * PN + 70], ref[200 * PN + 70]);
  return st != cudaSuccess || bad;
}
