// Issue-rate probe: legacy mma.sync (TF32 m16n8k8, FP16 m16n8k16) and FFMA / FFMA2 on sm_100a.  nvcc -arch=sm_100a -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void tf32(float* out, int iters) {
  float d[8][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  float s = 0; for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void f16(float* out, int iters) {
  float d[8][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  float s = 0; for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma(float* out, int iters, float x, float y) {
  float d[16]; for (int j = 0; j < 16; ++j) d[j] = threadIdx.x + j;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = fmaf(d[j], x, y);
  float s = 0; for (int j = 0; j < 16; ++j) s += d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma3(float* out, int iters, const float* p) {      // three distinct register operands: a[i] * b[j] + c[i][j]
  float c[4][4] = {}, a[4], b[4];
  for (int j = 0; j < 4; ++j) { a[j] = p[threadIdx.x + j]; b[j] = p[threadIdx.x + 4 + j]; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[r][q] = fmaf(a[r], b[q], c[r][q]);
    a[i & 3] += 1.f;
  }
  float s = 0; for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) s += c[r][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma2k(float* out, int iters, const float* p) {
  float2 c[4][2] = {}; float a[4]; float2 b[2];
  for (int j = 0; j < 4; ++j) a[j] = p[threadIdx.x + j];
  b[0] = make_float2(p[threadIdx.x + 4], p[threadIdx.x + 5]); b[1] = make_float2(p[threadIdx.x + 6], p[threadIdx.x + 7]);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q) c[r][q] = __ffma2_rn(make_float2(a[r], a[r]), b[q], c[r][q]);
    a[i & 3] += 1.f;
  }
  float s = 0; for (int r = 0; r < 4; ++r) for (int q = 0; q < 2; ++q) s += c[r][q].x + c[r][q].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float run(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float *out, *p; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&p, 4096 * 4); cudaMemset(p, 0, 4096 * 4);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int thr = warps * 32, blocks = 148;
    float ms = run([&] { tf32<<<blocks, thr>>>(out, iters); });
    printf("warps/SM %2d  mma.sync tf32 m16n8k8 : %8.1f TFLOP/s\n", warps, 2.0 * 16 * 8 * 8 * 8 * iters * warps * blocks / ms / 1e9);
    ms = run([&] { f16<<<blocks, thr>>>(out, iters); });
    printf("warps/SM %2d  mma.sync f16 m16n8k16 : %8.1f TFLOP/s\n", warps, 2.0 * 16 * 8 * 16 * 8 * iters * warps * blocks / ms / 1e9);
    ms = run([&] { ffma<<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); });
    printf("warps/SM %2d  FFMA (2 reg + 1)      : %8.1f TFLOP/s\n", warps, 2.0 * 16 * 32 * iters * warps * blocks / ms / 1e9);
    ms = run([&] { ffma3<<<blocks, thr>>>(out, iters, p); });
    printf("warps/SM %2d  FFMA outer product    : %8.1f TFLOP/s\n", warps, 2.0 * 16 * 32 * iters * warps * blocks / ms / 1e9);
    ms = run([&] { ffma2k<<<blocks, thr>>>(out, iters, p); });
    printf("warps/SM %2d  FFMA2 outer product   : %8.1f TFLOP/s\n", warps, 2.0 * 16 * 32 * iters * warps * blocks / ms / 1e9);
  }
  return 0;
}
