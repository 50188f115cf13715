// Bring-up probes for the remaining tcgen05 mechanics the forward path uses (tc_common.cuh):
//   pair        : cta_group::2 MMA (M=256 over a 2-CTA cluster, B split by N), multicast commit, remote mbarrier arrive
//   amn0 / amn1 : A operand supplied MN-major (no swizzle), N=32, K=48; amn0: LBO=k-group stride, SBO=m-group stride
// usage: umma2_probe <test>     (one test per process: a faulting mode poisons the CUDA context)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../clair_b200/csrc/tc_common.cuh"
using namespace clairb::tc;

// ------------------------------------------------------------------------------------------------
constexpr int PM = 256, PN = 256, PK = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe_pair(const __half* __restrict__ Ag, const __half* __restrict__ Bg, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* As = (__half*)smem;                        // this CTA's 128 rows of A: [K/8][128][8]
  __half* Bs = (__half*)(smem + 128 * PK * 2);       // this CTA's 128 rows (N) of B: [K/8][128][8]
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
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, 2 * 128 * PK * 2);
    bulk_g2s(As, Ag + (size_t)rank * 128 * PK, 128 * PK * 2, &bar_load);
    bulk_g2s(Bs, Bg + (size_t)rank * 128 * PK, 128 * PK * 2, &bar_load);
    mbar_wait(&bar_load, 0);
    if (rank == 1) {
      mbar_arrive_cluster(map_to_cta(smem_u32(&bar_peer), 0));     // tell the leader our operands landed
    } else {
      mbar_wait_cluster(&bar_peer, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(PM, PN);
      for (int j = 0; j < PK / 16; ++j) {
        uint64_t ad = make_smem_desc(smem_u32(As) + j * 2 * (128 * 16), 128 * 16, 128);
        uint64_t bd = make_smem_desc(smem_u32(Bs) + j * 2 * (128 * 16), 128 * 16, 128);
        umma_f16_pair(tmem, ad, bd, idesc, j > 0);
      }
      umma_commit_pair(&bar_mma, 0b11);
    }
  }
  __syncwarp();
  mbar_wait_cluster(&bar_mma, 0);
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

int run_pair() {
  std::vector<float> A(PM * PK), B(PN * PK);
  srand(2);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  std::vector<__half> Ah(PM * PK), Bh(PN * PK);
  // per CTA q: rows [128q,128q+128) in [K/8][128][8] order
  for (int r = 0; r < PM; ++r)
    for (int k = 0; k < PK; ++k) {
      __half h = __float2half(A[r * PK + k]);
      A[r * PK + k] = __half2float(h);
      Ah[(size_t)(r / 128) * 128 * PK + (k / 8) * 128 * 8 + (r % 128) * 8 + k % 8] = h;
    }
  for (int n = 0; n < PN; ++n)
    for (int k = 0; k < PK; ++k) {
      __half h = __float2half(B[n * PK + k]);
      B[n * PK + k] = __half2float(h);
      Bh[(size_t)(n / 128) * 128 * PK + (k / 8) * 128 * 8 + (n % 128) * 8 + k % 8] = h;
    }
  std::vector<float> ref(PM * PN);
  for (int r = 0; r < PM; ++r)
    for (int n = 0; n < PN; ++n) {
      double s = 0;
      for (int k = 0; k < PK; ++k) s += (double)A[r * PK + k] * B[n * PK + k];
      ref[r * PN + n] = (float)s;
    }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, Ah.size() * 2);
  cudaMalloc(&dB, Bh.size() * 2);
  cudaMalloc(&dD, PM * PN * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, PM * PN * 4);
  size_t smem = 2 * 128 * PK * 2 + 1024;
  cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_pair<<<2, 128, smem>>>(dA, dB, dD);
  cudaError_t st = cudaDeviceSynchronize();
  std::vector<float> D(PM * PN);
  cudaMemcpy(D.data(), dD, PM * PN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0, badq[4] = {0, 0, 0, 0};
  for (int i = 0; i < PM * PN; ++i) {
    double e = fabs(D[i] - ref[i]);
    if (e > maxerr) maxerr = e;
    if (e > 1e-3) {
      ++bad;
      ++badq[((i / PN) / 128) * 2 + ((i % PN) / 128)];
    }
  }
  printf("pair: %s  max|err| %.3e  mismatches %d / %d  (by quadrant rows/cols: %d %d %d %d)  D[0]=%f ref=%f  D[200*N+130]=%f ref=%f\n",
         cudaGetErrorString(st), maxerr, bad, PM * PN, badq[0], badq[1], badq[2], badq[3], D[0], ref[0], D[200 * PN + 130],
         ref[200 * PN + 130]);
  return st != cudaSuccess || bad;
}

// ------------------------------------------------------------------------------------------------
constexpr int QM = 128, QN = 32, QK = 48;

__global__ void __launch_bounds__(128) probe_amn(const __half* __restrict__ Ag, const __half* __restrict__ Bg,
                                                 float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* As = (__half*)smem;                         // MN-major: [K/8][M/8][8 k][8 m]
  __half* Bs = (__half*)(smem + QM * QK * 2);         // K-major:  [K/8][N][8]
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<32>(&tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, (QM + QN) * QK * 2);
    bulk_g2s(As, Ag, QM * QK * 2, &bar_load);
    bulk_g2s(Bs, Bg, QN * QK * 2, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16_amn(QM, QN);
    const uint32_t kgrp = (QM / 8) * 128;            // bytes between k-groups of A
    for (int j = 0; j < QK / 16; ++j) {
      uint32_t a_addr = smem_u32(As) + j * 2 * kgrp;
      uint64_t ad = mode == 0 ? make_smem_desc(a_addr, kgrp, 128) : make_smem_desc(a_addr, 128, kgrp);
      uint64_t bd = make_smem_desc(smem_u32(Bs) + j * 2 * (QN * 16), QN * 16, 128);
      umma_f16(tmem, ad, bd, idesc, j > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < QN; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * QN + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int run_amn(int mode) {
  std::vector<float> A(QM * QK), B(QN * QK);
  srand(3);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  std::vector<__half> Ah(QM * QK), Bh(QN * QK);
  for (int m = 0; m < QM; ++m)
    for (int k = 0; k < QK; ++k) {
      __half h = __float2half(A[m * QK + k]);
      A[m * QK + k] = __half2float(h);
      Ah[(size_t)(k / 8) * (QM / 8) * 64 + (m / 8) * 64 + (k % 8) * 8 + (m % 8)] = h;
    }
  for (int n = 0; n < QN; ++n)
    for (int k = 0; k < QK; ++k) {
      __half h = __float2half(B[n * QK + k]);
      B[n * QK + k] = __half2float(h);
      Bh[(size_t)(k / 8) * QN * 8 + n * 8 + k % 8] = h;
    }
  std::vector<float> ref(QM * QN);
  for (int m = 0; m < QM; ++m)
    for (int n = 0; n < QN; ++n) {
      double s = 0;
      for (int k = 0; k < QK; ++k) s += (double)A[m * QK + k] * B[n * QK + k];
      ref[m * QN + n] = (float)s;
    }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, Ah.size() * 2);
  cudaMalloc(&dB, Bh.size() * 2);
  cudaMalloc(&dD, QM * QN * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, QM * QN * 4);
  size_t smem = (QM + QN) * QK * 2 + 1024;
  cudaFuncSetAttribute(probe_amn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_amn<<<1, 128, smem>>>(dA, dB, dD, mode);
  cudaError_t st = cudaDeviceSynchronize();
  std::vector<float> D(QM * QN);
  cudaMemcpy(D.data(), dD, QM * QN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < QM * QN; ++i) {
    double e = fabs(D[i] - ref[i]);
    if (e > maxerr) maxerr = e;
    if (e > 1e-3) ++bad;
  }
  printf("amn%d: %s  max|err| %.3e  mismatches %d / %d   D[0]=%f ref=%f D[77*N+5]=%f ref=%f\n", mode, cudaGetErrorString(st),
         maxerr, bad, QM * QN, D[0], ref[0], D[77 * QN + 5], ref[77 * QN + 5]);
  return st != cudaSuccess || bad;
}

// ------------------------------------------------------------------------------------------------
// amnsw: A operand MN-major with SWIZZLE_128B.  Atom = 8 k-rows x 64 m (128 B per row), 16-byte chunk index XORed with the
// k-row index.  mode 0: LBO = m-atom stride, SBO = k-atom stride; mode 1: swapped.
__global__ void __launch_bounds__(128) probe_amnsw(const __half* __restrict__ Ag, const __half* __restrict__ Bg,
                                                   float* __restrict__ D, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __half* As = (__half*)smem;                         // [k-atom 6][m-atom 2][8 k][64 m] swizzled
  __half* Bs = (__half*)(smem + QM * QK * 2);         // K-major no swizzle: [K/8][N][8]
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<32>(&tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, (QM + QN) * QK * 2);
    bulk_g2s(As, Ag, QM * QK * 2, &bar_load);
    bulk_g2s(Bs, Bg, QN * QK * 2, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16_amn(QM, QN);
    const uint32_t matom = 1024, katom = 2 * 1024;    // bytes: next 64 sites, next 8 k
    for (int j = 0; j < QK / 16; ++j) {
      uint32_t a_addr = smem_u32(As) + j * 2 * katom;
      uint64_t ad = mode == 0 ? make_smem_desc(a_addr, matom, katom) : make_smem_desc(a_addr, katom, matom);
      ad |= (uint64_t)2 << 61;                        // layout_type = SWIZZLE_128B
      uint64_t bd = make_smem_desc(smem_u32(Bs) + j * 2 * (QN * 16), QN * 16, 128);
      umma_f16(tmem, ad, bd, idesc, j > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < QN; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[row * QN + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int run_amnsw(int mode) {
  std::vector<float> A(QM * QK), B(QN * QK);
  srand(5);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  std::vector<__half> Ah(QM * QK), Bh(QN * QK);
  for (int m = 0; m < QM; ++m)
    for (int k = 0; k < QK; ++k) {
      __half h = __float2half(A[m * QK + k]);
      A[m * QK + k] = __half2float(h);
      const int ka = k / 8, kr = k % 8, ma = m / 64, chunk = (m % 64) / 8, e = m % 8;
      Ah[(size_t)(ka * 2 + ma) * 512 + kr * 64 + ((chunk ^ kr) * 8) + e] = h;
    }
  for (int n = 0; n < QN; ++n)
    for (int k = 0; k < QK; ++k) {
      __half h = __float2half(B[n * QK + k]);
      B[n * QK + k] = __half2float(h);
      Bh[(size_t)(k / 8) * QN * 8 + n * 8 + k % 8] = h;
    }
  std::vector<float> ref(QM * QN);
  for (int m = 0; m < QM; ++m)
    for (int n = 0; n < QN; ++n) {
      double s = 0;
      for (int k = 0; k < QK; ++k) s += (double)A[m * QK + k] * B[n * QK + k];
      ref[m * QN + n] = (float)s;
    }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, Ah.size() * 2);
  cudaMalloc(&dB, Bh.size() * 2);
  cudaMalloc(&dD, QM * QN * 4);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, QM * QN * 4);
  size_t smem = (QM + QN) * QK * 2 + 2048;
  cudaFuncSetAttribute(probe_amnsw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_amnsw<<<1, 128, smem>>>(dA, dB, dD, mode);
  cudaError_t st = cudaDeviceSynchronize();
  std::vector<float> D(QM * QN);
  cudaMemcpy(D.data(), dD, QM * QN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < QM * QN; ++i) {
    double e = fabs(D[i] - ref[i]);
    if (e > maxerr) maxerr = e;
    if (e > 1e-3) ++bad;
  }
  printf("amnsw%d: %s  max|err| %.3e  mismatches %d / %d   D[0]=%f ref=%f D[77*N+5]=%f ref=%f\n", mode, cudaGetErrorString(st),
         maxerr, bad, QM * QN, D[0], ref[0], D[77 * QN + 5], ref[77 * QN + 5]);
  return st != cudaSuccess || bad;
}

int main(int argc, char** argv) {
  const char* t = argc > 1 ? argv[1] : "pair";
  if (!strcmp(t, "pair")) return run_pair();
  if (!strcmp(t, "amn0")) return run_amn(0);
  if (!strcmp(t, "amn1")) return run_amn(1);
  if (!strcmp(t, "amnsw0")) return run_amnsw(0);
  if (!strcmp(t, "amnsw1")) return run_amnsw(1);
  fprintf(stderr, "unknown test %s\n", t);
  return 2;
}
