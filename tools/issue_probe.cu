// Do MUFU pipe time and the issue slots of other instructions overlap on sm_100a?  nvcc -arch=sm_100a -O3 -o issue_probe issue_probe.cu
// Per thread: 16 independent ex2 chains + NF independent FFMA chains per ex2; cycles per (ex2 + NF ffma) group per warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int NF, int KIND>
__global__ void probe(float* out, long long* cyc, int iters) {
  float m[16], f[16];
  for (int i = 0; i < 16; ++i) { m[i] = (threadIdx.x * 37 + i * 11) % 64 * 0.01f - 0.3f; f[i] = m[i] * 0.5f; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (KIND == 0) m[i] = ex2f(m[i]) - 1.2f; else m[i] = rcpf(m[i] + 2.f);
#pragma unroll
      for (int j = 0; j < NF; ++j) { const int q = (i * NF + j) & 15; asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[q]) : "f"(0.999f), "f"(0.001f)); }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < 16; ++i) acc += m[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NF, int KIND>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) { probe<NF, KIND><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double per_group = (double)h[0] / iters / 16;     // cycles per (mufu + NF ffma) per warp
    printf("%s + %d ffma  warps/SM %2d (%d per scheduler): %.2f cycles per group per warp -> %.2f per scheduler per group\n", name, NF, warps, warps / 4,
           per_group, per_group / (warps / 4));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0, 0>("ex2", out, cyc); run<2, 0>("ex2", out, cyc); run<4, 0>("ex2", out, cyc); run<6, 0>("ex2", out, cyc); run<8, 0>("ex2", out, cyc); run<12, 0>("ex2", out, cyc);
  run<0, 1>("rcp", out, cyc); run<4, 1>("rcp", out, cyc);
  return 0;
}
