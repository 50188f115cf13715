// MUFU / LSTM-cell throughput probe for the recurrent epilogue (B200): nvcc -arch=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float lstm_cell(float pi, float pg, float pf, float po, float& c) {
  const float a = ex2f(fminf(pi, 40.f)), b = ex2f(fminf(pg, 40.f)), d = ex2f(fminf(pf, 40.f));
  const float A = 1.f + a, B = 1.f + b, D = 1.f + d;
  const float AB = A * B;
  const float cn = fmaf(c, AB, (1.f - b) * D) * rcpf(AB * D);
  c = cn;
  const float e = ex2f(fminf(cn * (-2.f * LOG2E), 40.f)), f = ex2f(fminf(po, 40.f));
  return (1.f - e) * rcpf((1.f + e) * (1.f + f));
}
// polynomial 2^x on the FMA pipe (|rel err| ~ 2e-7): round-to-nearest split + degree-6 minimax on [-0.5, 0.5]
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: integer part in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;
  float p = 1.530177e-4f;
  p = fmaf(p, f, 1.339887e-3f);
  p = fmaf(p, f, 9.618437e-3f);
  p = fmaf(p, f, 5.550357e-2f);
  p = fmaf(p, f, 2.402265e-1f);
  p = fmaf(p, f, 6.931472e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float lstm_cell_poly(float pi, float pg, float pf, float po, float& c) {
  const float a = ex2f(fminf(pi, 40.f)), b = ex2_poly(fminf(pg, 40.f)), d = ex2f(fminf(pf, 40.f));
  const float A = 1.f + a, B = 1.f + b, D = 1.f + d;
  const float AB = A * B;
  const float cn = fmaf(c, AB, (1.f - b) * D) * rcpf(AB * D);
  c = cn;
  const float e = ex2f(fminf(cn * (-2.f * LOG2E), 40.f)), f = ex2_poly(fminf(po, 40.f));
  return (1.f - e) * rcpf((1.f + e) * (1.f + f));
}
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
  float v[32], c[8], acc = 0.f;
  for (int i = 0; i < 32; ++i) v[i] = (threadIdx.x * 37 + i * 11) % 64 * 0.05f - 1.6f;
  for (int i = 0; i < 8; ++i) c[i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2f(v[i]) - 1.5f;
    } else if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { float h = lstm_cell(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3], c[k]); acc += h; v[4 * k] += 1e-3f * h; }
    } else if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { float h = lstm_cell_poly(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3], c[k]); acc += h; v[4 * k] += 1e-3f * h; }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2_poly(v[i]) - 1.5f;
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[4] = {"ex2 (32 independent / thread)", "lstm_cell x8 (7 MUFU/cell)", "lstm_cell x8, 2 of 5 ex2 as polynomial", "ex2_poly x32"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) probe<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) probe<2><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 3) probe<3><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double per_iter = (double)h[0] / iters;
      if (mode == 0 || mode == 3) printf("%-40s warps/SM %2d: %.1f cyc/iter -> %.2f ex2 lanes/clk/SM\n", names[mode], warps, per_iter, warps * 32 * 32 / per_iter);
      else printf("%-40s warps/SM %2d: %.1f cyc/iter -> %.2f cells/clk/SM (%.0f cyc per 4096-cell gate block)\n", names[mode], warps, per_iter, warps * 32 * 8 / per_iter, 4096 / (warps * 32 * 8 / per_iter));
    }
  // accuracy of the polynomial
  return 0;
}
