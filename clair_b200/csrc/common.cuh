// Shared constants and small device helpers for the clair_b200 forward path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace clairb {

// Graph constants (reference shared/param.py:9-12, clair/model.py:80-93, clair/task/main.py:10-29)
constexpr int T_STEPS = 33;
constexpr int F_IN = 32;          // 8 rows * 4 channels
constexpr int H = 128;            // LSTM units per direction
constexpr int G4 = 4 * H;         // gate columns per direction
constexpr int L3_UNITS = 30;
constexpr int L3_K = L3_UNITS * 2 * H;   // 7680
constexpr int L4_UNITS = 192;
constexpr int L5_UNITS = 96;
constexpr int L5_ALL = 4 * L5_UNITS;     // 384
constexpr int N_OUT = 90;
constexpr int SITE_ELEMS = T_STEPS * F_IN;   // 1056
constexpr int TILE = 128;         // sites per tile; every predict-batch is padded to a multiple of it

__device__ __constant__ const int kHeadOff[5] = {0, 21, 24, 57, 90};

// clair/selu.py:28-29
constexpr float SELU_ALPHA = 1.6732632423543772848170429916717f;
constexpr float SELU_SCALE = 1.0507009873554804934193349852946f;

__device__ __forceinline__ float selu_f(float x) {
  // scale * where(x>=0, x, alpha*(exp(x)-1))   (clair/selu.py:30)
  return x >= 0.f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * expm1f(x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Padded-site bookkeeping: a call of n sites is ceil(n/batch) predict-batches, each padded to
// bp = roundup(batch, TILE) rows so that tiles never straddle a batch.
struct SiteMap {
  int64_t n;        // real sites
  int batch;        // sites per predict-batch
  int bp;           // padded sites per predict-batch
  int64_t np;       // padded total = n_batches * bp
  __host__ __device__ int64_t real_row(int64_t p) const {   // -1 for padding rows
    int64_t b = p / bp;
    int j = (int)(p - b * bp);
    int64_t r = b * batch + j;
    return (j < batch && r < n) ? r : -1;
  }
  __host__ __device__ int64_t padded_row(int64_t r) const {
    int64_t b = r / batch;
    return b * bp + (r - b * batch);
  }
};

}  // namespace clairb
