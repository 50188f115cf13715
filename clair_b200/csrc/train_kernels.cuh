// Training step of the Clair network on the device: forward in training mode, focal loss, backward
// through the dense trunk, the slice-dense layer and both BiLSTMs (BPTT over 33 steps), global-norm clip, Adam.
//
// Reference: clair/model.py
//   forward in training phase   :400-622, tf.layers.dropout :434-459, selu.dropout_selu clair/selu.py:43-74
//   focal loss                  :783-805        L2 :689-694        total :696-709
//   clip_by_global_norm(5.0) + AdamOptimizer    :717-728
// Everything here is fp32 on the CUDA cores - the reference trains in fp32 (float_type is forced to tf.float32, :165-170) and
// the parity tests compare gradients with a float64 autograd restatement (oracle/train_oracle.py).  This is the first correct
// device path of the row: the per-step recurrences are small fused kernels (one CTA per 8 sites), every large contraction
// (input projections, weight gradients) is one tiled SGEMM over all 33 steps.  The tensor-core kernels of the inference
// path do not save the gate activations BPTT needs and are not used here.
//
// Layouts (row-major fp32): activations of a layer in time-major order [33][n][...]; per direction the recurrent kernels work
// in PROCESSING order s = 0..32 (s = t for fw, s = 32 - t for bw); gate columns in TF order i, c(candidate), f, o.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace clairb {
namespace train {

constexpr int ROWS = 8;                  // batches are padded to a multiple of 8 sites (zero rows)
constexpr float ALPHA_DROPOUT = -1.7580993408473766f;     // clair/selu.py:43

// Two independent fp32 FMAs in one instruction (FFMA2, scalar a broadcast over the pair): c + a * b, each half rounded like fmaf.
// A three-register FFMA issues every second cycle per scheduler on this part; the paired form is what reaches 128 FMA/clk/SM.
__device__ __forceinline__ float2 fma2(float a, float2 b, float2 c) { return __ffma2_rn(make_float2(a, a), b, c); }

// ---- C[M,N] = alpha * op(A)[M,K] . op(B)[K,N] + beta * C     (row-major; TA: A is stored [K][M]; TB: B is stored [N][K]) ----
constexpr int GM = 64, GN = 64, GK = 16;
// gridDim.z > 1: split-K - slice z of the contraction is added to C with atomics (C holds beta * C already; the weight-gradient
// GEMMs contract over 33 * n rows into a few hundred outputs and would otherwise run on a handful of CTAs)
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                                             const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc, int k_per_slice) {
  __shared__ float As[GK][GM + 4], Bs[GK][GN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int k_begin = blockIdx.z * k_per_slice;
  const int k_end = min(K, k_begin + k_per_slice);
  float2 acc[4][2] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
    for (int i = threadIdx.x; i < GM * GK; i += 256) {
      int m, k;
      if (TA) { m = i % GM; k = i / GM; } else { k = i % GK; m = i / GK; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < k_end) ? (TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
    }
    for (int i = threadIdx.x; i < GN * GK; i += 256) {
      int nn, k;
      if (TB) { k = i % GK; nn = i / GK; } else { nn = i % GN; k = i / GN; }
      const int gn = n0 + nn, gk = k0 + k;
      Bs[k][nn] = (gn < N && gk < k_end) ? (TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fma2(a[i], make_float2(b[0], b[1]), acc[i][0]);
        acc[i][1] = fma2(a[i], make_float2(b[2], b[3]), acc[i][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float* c = C + (size_t)gm * ldc + gn;
        const float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        if (gridDim.z > 1) atomicAdd(c, alpha * v);
        else *c = alpha * v + (beta != 0.f ? beta * *c : 0.f);
      }
    }
}

// The same contract on 128 x 128 x 16 tiles for the large contractions (input projections, weight gradients, L4), on the tensor
// cores with fp32 accuracy: every operand is split into two TF32 values (x = hi + lo, 11 + 11 significant bits) and a product is
// a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, accumulated in fp32 by mma.sync.m16n8k8 (the term dropped, a_lo.b_lo, is below 2^-22 of the
// product).  Measured on B200: 275 TFLOP/s of TF32 mma.sync = 92 TFLOP/s of such products, against 35 for the FFMA version of this
// tile (tools/probes/mma_rate.cu).  8 warps, a warp owns 64 x 32 of the tile (4 x 4 MMA tiles).  An operand whose global layout has
// k outermost is staged [k][136], the other kind [row][20]: both are stored with plain 16-byte stores and both give conflict-free
// fragment loads.  Two shared-memory stages; the next tile's global loads are in flight during the current tile's MMAs.
// Requires M, N, K, lda, ldb, ldc and k_per_slice to be multiples of 4 and 16-byte aligned bases (gemm() checks).
constexpr int BM = 128, BN = 128, BK = 16, LD_K = BM + 8, LD_R = BK + 4, STAGE_FLOATS = BM * LD_R;      // 2560 >= 16 * 136
// hi = x rounded to TF32's 10 mantissa bits in integer arithmetic (half away from zero); lo = x - hi is exact in fp32, has either
// sign, and is handed over as it is - the tensor core ignores its low 13 bits, an unbiased 2^-21 |x| at most.  cvt.rna.tf32 does the
// same rounding on the quarter-rate conversion pipe (twice per element it cost as much as the MMAs); cutting hi instead of rounding
// it biases every product the same way, which the near-cancelling sums of a weight gradient amplify (Adam test, 6 % of a step).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) sgemm_big(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float beta, float* __restrict__ C, int ldc, int k_per_slice) {
  constexpr bool A_K = TA, B_K = !TB;                    // operand stored with k outermost in global memory
  __shared__ __align__(16) float As[2][STAGE_FLOATS], Bs[2][STAGE_FLOATS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_slice;
  const int k_end = min(K, k_begin + k_per_slice);
  float4 ra[2], rb[2];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int k0) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int f = tid + 256 * p;
      if (A_K) { const int k = f >> 5, m = (f & 31) * 4; ra[p] = (m0 + m < M && k0 + k < k_end) ? *reinterpret_cast<const float4*>(A + (size_t)(k0 + k) * lda + m0 + m) : zero4; }
      else { const int m = f >> 2, k = (f & 3) * 4; ra[p] = (m0 + m < M && k0 + k < k_end) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + m) * lda + k0 + k) : zero4; }
      if (B_K) { const int k = f >> 5, nn = (f & 31) * 4; rb[p] = (n0 + nn < N && k0 + k < k_end) ? *reinterpret_cast<const float4*>(B + (size_t)(k0 + k) * ldb + n0 + nn) : zero4; }
      else { const int nn = f >> 2, k = (f & 3) * 4; rb[p] = (n0 + nn < N && k0 + k < k_end) ? *reinterpret_cast<const float4*>(B + (size_t)(n0 + nn) * ldb + k0 + k) : zero4; }
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int f = tid + 256 * p;
      *reinterpret_cast<float4*>(&As[buf][A_K ? (f >> 5) * LD_K + (f & 31) * 4 : (f >> 2) * LD_R + (f & 3) * 4]) = ra[p];
      *reinterpret_cast<float4*>(&Bs[buf][B_K ? (f >> 5) * LD_K + (f & 31) * 4 : (f >> 2) * LD_R + (f & 3) * 4]) = rb[p];
    }
  };
  float acc[4][4][4] = {};                               // [m tile][n tile][c0..c3]
  load(k_begin);
  stage(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) load(k0 + BK);
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nb = wn + 8 * j + g;
        split_tf32(B_K ? bs[(kk + tig) * LD_K + nb] : bs[nb * LD_R + kk + tig], bh[j][0], bl[j][0]);
        split_tf32(B_K ? bs[(kk + tig + 4) * LD_K + nb] : bs[nb * LD_R + kk + tig + 4], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mb = wm + 16 * i + g;
        uint32_t ah[4], al[4];
        split_tf32(A_K ? as[(kk + tig) * LD_K + mb] : as[mb * LD_R + kk + tig], ah[0], al[0]);
        split_tf32(A_K ? as[(kk + tig) * LD_K + mb + 8] : as[(mb + 8) * LD_R + kk + tig], ah[1], al[1]);
        split_tf32(A_K ? as[(kk + tig + 4) * LD_K + mb] : as[mb * LD_R + kk + tig + 4], ah[2], al[2]);
        split_tf32(A_K ? as[(kk + tig + 4) * LD_K + mb + 8] : as[(mb + 8) * LD_R + kk + tig + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], al, bh[j]);      // the small terms first; term-major so that four
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah, bl[j]);      // independent MMAs lie between two on one accumulator
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah, bh[j]);
      }
    }
    if (more) stage(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int gm = m0 + wm + 16 * i + g + 8 * half;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + wn + 8 * j + 2 * tig;
        if (gn >= N) continue;
        float* c = C + (size_t)gm * ldc + gn;
        float2 v = make_float2(acc[i][j][2 * half], acc[i][j][2 * half + 1]);
        if (gridDim.z > 1) {
          atomicAdd(c, v.x); atomicAdd(c + 1, v.y);
        } else {
          if (beta != 0.f) {
            const float2 o = *reinterpret_cast<const float2*>(c);
            v.x += beta * o.x; v.y += beta * o.y;
          }
          *reinterpret_cast<float2*>(c) = v;
        }
      }
    }
}

// rows of C (M x N, leading dimension ldc) <- beta * C, before a split-K GEMM adds its slices
__global__ void scale_matrix(float* __restrict__ C, int M, int N, int ldc, float beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)M * N) {
    float* c = C + (i / N) * ldc + i % N;
    *c = beta != 0.f ? beta * *c : 0.f;
  }
}

inline void gemm(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc,
                 cudaStream_t st, int64_t* launches) {
  const bool big = M >= 96 && N >= 96 && !((M | N | K | lda | ldb | ldc) & 3) && !(((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15);
  const int tm = big ? BM : GM, tn = big ? BN : GN, tk = big ? BK : GK;
  dim3 grid((N + tn - 1) / tn, (M + tm - 1) / tm);
  // split the contraction when the output tiles alone cannot fill the machine (148 SMs, 2 resident CTAs of the large kernel, 4 of
  // the small one); the slices are added to C with atomics
  int slices = 1;
  const int tiles = (int)(grid.x * grid.y);
  const int want = big ? 296 : 592;
  if (tiles < want / 2 && K >= 256) slices = min(64, max(1, min(want / tiles, K / 64)));
  int k_per_slice = ((K + slices - 1) / slices + tk - 1) / tk * tk;
  slices = (K + k_per_slice - 1) / k_per_slice;
  grid.z = slices;
  if (slices > 1) {
    scale_matrix<<<(unsigned)(((int64_t)M * N + 255) / 256), 256, 0, st>>>(C, M, N, ldc, beta);
    ++*launches;
  }
  if (big) {
    if (!ta && !tb) sgemm_big<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else if (ta && !tb) sgemm_big<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else if (!ta && tb) sgemm_big<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else sgemm_big<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
  } else if (!ta && !tb) sgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, k_per_slice);
  else if (ta && !tb) sgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, k_per_slice);
  else if (!ta && tb) sgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, k_per_slice);
  else sgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, k_per_slice);
  ++*launches;
}

// ---- small elementwise / layout kernels --------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// x [n][33][32] (float or int16) -> x_tm [33][n][32]
template <typename TIn>
__global__ void input_time_major(const TIn* __restrict__ x, float* __restrict__ x_tm, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * SITE_ELEMS) return;
  const int f = (int)(i % F_IN), t = (int)((i / F_IN) % T_STEPS);
  const int64_t b = i / SITE_ELEMS;
  x_tm[((size_t)t * n + b) * F_IN + f] = (float)x[i];
}
// a = selu(z + bias) in place (clair/selu.py:26-30)
__global__ void bias_selu(float* __restrict__ Z, const float* __restrict__ bias, int64_t rows, int N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * N) Z[i] = selu_f(Z[i] + bias[i % N]);
}
// dZ = dA * selu'(z), from the activation: selu'(z) = scale (z >= 0), a + scale * alpha (z < 0)
__global__ void selu_backward(float* __restrict__ dA, const float* __restrict__ A, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const float a = A[i];
    dA[i] *= a >= 0.f ? SELU_SCALE : a + SELU_SCALE * SELU_ALPHA;     // a >= 0 <=> z >= 0 (selu is monotone, selu(0) = 0)
  }
}
// column sums of D [rows][N] -> out[N] (bias gradients)
__global__ void column_sums(const float* __restrict__ D, int64_t rows, int N, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double s = 0.0;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) s += D[r * N + j];
  atomicAdd(&out[j], (float)s);
}
// tf.layers.dropout (inverted): y = x * mask / keep; used forwards on the activations and backwards on their gradients
__global__ void dropout_scale(float* __restrict__ X, const uint8_t* __restrict__ mask, float inv_keep, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) X[i] = mask[i] ? X[i] * inv_keep : 0.f;
}
// selu.dropout_selu forward: y = a * (x * m + alpha' * (1 - m)) + b; backward: dx = dy * a * m
__global__ void alpha_dropout_forward(float* __restrict__ X, const uint8_t* __restrict__ mask, float a, float b, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) X[i] = a * (mask[i] ? X[i] : ALPHA_DROPOUT) + b;
}
__global__ void alpha_dropout_backward(float* __restrict__ dX, const uint8_t* __restrict__ mask, float a, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dX[i] = mask[i] ? dX[i] * a : 0.f;
}
// keep-masks from a counter-based hash (murmur3 finaliser of seed, stream, index): 1 = kept with probability 1 - rate
__global__ void make_mask(uint8_t* __restrict__ mask, int64_t count, float rate, uint64_t seed, uint64_t stream) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint64_t h = seed * 0x9E3779B97F4A7C15ull + stream * 0xBF58476D1CE4E5B9ull + (uint64_t)i;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  const float u = (float)(h >> 40) * (1.f / 16777216.f);
  mask[i] = u >= rate;
}
// layer output [33][n][256] <- the two directions' h in processing order ([33][n][128] each; hbuf has a leading zero block)
__global__ void assemble_bidirectional(const float* __restrict__ h_fw, const float* __restrict__ h_bw, float* __restrict__ out, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T_STEPS * n * 2 * H) return;
  const int f = (int)(i % (2 * H));
  const int64_t tb = i / (2 * H);
  const int t = (int)(tb / n);
  const int64_t b = tb % n;
  out[i] = f < H ? h_fw[((size_t)t * n + b) * H + f] : h_bw[((size_t)(T_STEPS - 1 - t) * n + b) * H + (f - H)];
}
// the reverse: d(layer output) [33][n][256] -> per-direction dh in processing order
__global__ void split_bidirectional(const float* __restrict__ dout, float* __restrict__ dh_fw, float* __restrict__ dh_bw, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T_STEPS * n * 2 * H) return;
  const int f = (int)(i % (2 * H));
  const int64_t tb = i / (2 * H);
  const int t = (int)(tb / n);
  const int64_t b = tb % n;
  if (f < H) dh_fw[((size_t)t * n + b) * H + f] = dout[i];
  else dh_bw[((size_t)(T_STEPS - 1 - t) * n + b) * H + (f - H)] = dout[i];
}
// x [33][n][K] -> the same rows in reversed time order (input of a backward direction in processing order), or accumulate back
__global__ void reverse_time(const float* __restrict__ in, float* __restrict__ out, int n, int K, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T_STEPS * n * K) return;
  const int64_t per = (int64_t)n * K;
  const int t = (int)(i / per);
  const int64_t o = (int64_t)(T_STEPS - 1 - t) * per + i % per;
  if (accumulate) out[o] += in[i]; else out[o] = in[i];
}

// ---- the 33 steps of one LSTM direction, forward: z_s = pre[s] + h_{s-1} . W_h ; gates ; c_s, h_s ------------------------------
// (LSTMBlockCell, forget_bias 0; clair/model.py:299-305).  One thread-block CLUSTER of 8 CTAs carries 64 sites through all 33
// steps: CTA j owns hidden units 16j..16j+15, keeps their 64 gate columns of W_h (32 KB) and the whole h_{s-1} of the 64 sites
// (transposed, [unit][site]) in shared memory, computes the [64 x 64] gate tile on the FP32 pipe, and writes its 16 new h columns
// into the NEXT-step buffer of all 8 CTAs through distributed shared memory; one cluster barrier per step.  Thread (ty, tx) owns
// sites 4 ty..4 ty+3 of unit tx, i.e. the 4 x 4 accumulators (site, gate): the local gate columns are ordered unit-major.
constexpr int SEQ_ROWS = 64, SEQ_CTAS = 8, SEQ_UNITS = H / SEQ_CTAS;      // 64 sites per cluster, 16 units per CTA
constexpr int SEQ_FWD_SMEM = (H * 4 * SEQ_UNITS + 2 * H * SEQ_ROWS) * (int)sizeof(float);                                    // 96 KB
constexpr int SEQ_BWD_SMEM = (4 * SEQ_UNITS * H + 4 * SEQ_UNITS * SEQ_ROWS + 2 * SEQ_CTAS * SEQ_UNITS * SEQ_ROWS) * (int)sizeof(float);   // 112 KB

__global__ void __cluster_dims__(SEQ_CTAS, 1, 1) __launch_bounds__(256)
lstm_seq_forward(const float* __restrict__ pre, const float* __restrict__ Wh, const float* __restrict__ bias, float* __restrict__ gates,
                 float* __restrict__ cbuf, float* __restrict__ hbuf, int n) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float seq_smem[];
  float* Ws = seq_smem;                                  // [128 k][64 local columns = unit * 4 + gate]
  float* hT = seq_smem + H * 4 * SEQ_UNITS;              // [2][128 unit][64 site]
  const int j = (int)cluster.block_rank();
  const int r0 = (int)(blockIdx.x / SEQ_CTAS) * SEQ_ROWS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int unit = j * SEQ_UNITS + tx;
  const int row = r0 + 4 * ty;
  const bool live = row < n;                             // n is a multiple of 8: a group of 4 sites is in or out as a whole
  for (int i = threadIdx.x; i < H * 4 * SEQ_UNITS; i += 256) {
    const int k = i >> 6, lc = i & 63;
    Ws[i] = Wh[(size_t)k * G4 + (lc & 3) * H + j * SEQ_UNITS + (lc >> 2)];
  }
  for (int i = threadIdx.x; i < H * SEQ_ROWS; i += 256) hT[i] = 0.f;      // h_0 = 0
  float* remote[SEQ_CTAS];
#pragma unroll
  for (int d = 0; d < SEQ_CTAS; ++d) remote[d] = cluster.map_shared_rank(hT, d);
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  const float2 b01 = make_float2(bias[unit], bias[H + unit]), b23 = make_float2(bias[2 * H + unit], bias[3 * H + unit]);
  float2 z[4][2];                                        // site i: (i, g) and (f, o) pre-activations = x_s W_x (pre) + b + h_{s-1} W_h
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* q = pre + (size_t)(row + i) * G4 + unit;
    z[i][0] = live ? make_float2(q[0] + b01.x, q[H] + b01.y) : make_float2(0.f, 0.f);
    z[i][1] = live ? make_float2(q[2 * H] + b23.x, q[3 * H] + b23.y) : make_float2(0.f, 0.f);
  }
  cluster.sync();                                        // every CTA of the cluster runs and has its buffers initialised
  for (int s = 0; s < T_STEPS; ++s) {
    const float* hc = hT + (s & 1) * H * SEQ_ROWS;
    float2 zn[4][2];                                     // the next step's input projection, requested before the contraction
    if (s + 1 < T_STEPS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* q = pre + ((size_t)(s + 1) * n + row + i) * G4 + unit;
        zn[i][0] = live ? make_float2(q[0] + b01.x, q[H] + b01.y) : make_float2(0.f, 0.f);
        zn[i][1] = live ? make_float2(q[2 * H] + b23.x, q[3 * H] + b23.y) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(hc + k * SEQ_ROWS + 4 * ty);
      const float4 b = *reinterpret_cast<const float4*>(Ws + k * 4 * SEQ_UNITS + 4 * tx);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z[i][0] = fma2(av[i], make_float2(b.x, b.y), z[i][0]);
        z[i][1] = fma2(av[i], make_float2(b.z, b.w), z[i][1]);
      }
    }
    float hv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float ig = sigmoidf_(z[i][0].x), gg = tanhf(z[i][0].y), fg = sigmoidf_(z[i][1].x), og = sigmoidf_(z[i][1].y);
      c[i] = gg * ig + c[i] * fg;
      hv[i] = tanhf(c[i]) * og;
      if (live) {
        const size_t r = (size_t)s * n + row + i;
        gates[r * G4 + unit] = ig; gates[r * G4 + H + unit] = gg; gates[r * G4 + 2 * H + unit] = fg; gates[r * G4 + 3 * H + unit] = og;
        cbuf[(r + n) * H + unit] = c[i];
        hbuf[(r + n) * H + unit] = hv[i];
      }
    }
    const int off = ((s + 1) & 1) * H * SEQ_ROWS + unit * SEQ_ROWS + 4 * ty;
#pragma unroll
    for (int d = 0; d < SEQ_CTAS; ++d) *reinterpret_cast<float4*>(remote[d] + off) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    if (s + 1 < T_STEPS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { z[i][0] = zn[i][0]; z[i][1] = zn[i][1]; }
    }
    cluster.sync();                                      // h_s is complete in every CTA; nobody still reads h_{s-1}'s other buffer
  }
}

// ---- the 33 steps of one LSTM direction, backward: dh = dh_out[s] + dh_rec ; gate gradients dZ[s] ; dc_{s-1} ; dh_rec <- dZ[s] . W_h^T
// Same cluster shape.  CTA j produces the 64 gate-gradient columns of its 16 units, multiplies them by its 64 rows of W_h^T
// ([64 x 64] . [64 x 128], a partial sum of dh_rec for ALL 128 units), and scatters the partial columns to the CTA that owns each
// unit (slot j of that CTA, double-buffered); after the barrier every CTA adds its 8 slots in a fixed order.
__global__ void __cluster_dims__(SEQ_CTAS, 1, 1) __launch_bounds__(256)
lstm_seq_backward(const float* __restrict__ dh_out, const float* __restrict__ gates, const float* __restrict__ cbuf, const float* __restrict__ Wh,
                  float* __restrict__ dZ, float* __restrict__ dbias, int n) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float seq_smem[];
  float* Wt = seq_smem;                                  // [64 local gate columns][128 units of h_{s-1}]
  float* dzT = Wt + 4 * SEQ_UNITS * H;                   // [64 local gate columns][64 sites]
  float* slots = dzT + 4 * SEQ_UNITS * SEQ_ROWS;         // [2][8 source CTAs][16 units][64 sites]
  constexpr int SLOT = SEQ_UNITS * SEQ_ROWS;
  const int j = (int)cluster.block_rank();
  const int r0 = (int)(blockIdx.x / SEQ_CTAS) * SEQ_ROWS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int unit = j * SEQ_UNITS + tx;
  const int row = r0 + 4 * ty;
  const bool live = row < n;
  for (int i = threadIdx.x; i < 4 * SEQ_UNITS * H; i += 256) {
    const int lc = i >> 7, k = i & 127;
    Wt[i] = Wh[(size_t)k * G4 + (lc & 3) * H + j * SEQ_UNITS + (lc >> 2)];
  }
  float* remote[2];                                      // the two CTAs that own the units of this thread's partial columns
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) remote[hh] = cluster.map_shared_rank(slots, hh * 4 + (tx >> 2));
  float dc[4] = {0.f, 0.f, 0.f, 0.f}, dh_rec[4] = {0.f, 0.f, 0.f, 0.f};
  float db[4] = {0.f, 0.f, 0.f, 0.f};                   // bias gradient of this unit's four gates over the thread's sites and all steps
  float gi[4][4], cs[4], cp[4], dho[4];
  auto fetch = [&](int s) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t r = (size_t)s * n + row + i;
#pragma unroll
      for (int g = 0; g < 4; ++g) gi[i][g] = live ? gates[r * G4 + g * H + unit] : 0.f;
      cs[i] = live ? cbuf[(r + n) * H + unit] : 0.f;
      cp[i] = live ? cbuf[r * H + unit] : 0.f;
      dho[i] = live ? dh_out[r * H + unit] : 0.f;
    }
  };
  fetch(T_STEPS - 1);
  cluster.sync();
  for (int s = T_STEPS - 1; s >= 0; --s) {
    const int p = s & 1;
    float dz[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float ig = gi[i][0], gg = gi[i][1], fg = gi[i][2], og = gi[i][3];
      const float tc = tanhf(cs[i]);
      const float dh = dho[i] + dh_rec[i];
      const float dcs = dc[i] + dh * og * (1.f - tc * tc);
      dz[i][0] = dcs * gg * ig * (1.f - ig);
      dz[i][1] = dcs * ig * (1.f - gg * gg);
      dz[i][2] = dcs * cp[i] * fg * (1.f - fg);
      dz[i][3] = dh * tc * og * (1.f - og);
      dc[i] = dcs * fg;
      if (live) {
        const size_t r = (size_t)s * n + row + i;
#pragma unroll
        for (int g = 0; g < 4; ++g) { dZ[r * G4 + g * H + unit] = dz[i][g]; db[g] += dz[i][g]; }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<float4*>(dzT + (4 * tx + g) * SEQ_ROWS + 4 * ty) = make_float4(dz[0][g], dz[1][g], dz[2][g], dz[3][g]);
    __syncthreads();
    if (s > 0) fetch(s - 1);                             // independent of the recurrence: in flight during the contraction
    float2 acc[4][4] = {};                               // sites 4 ty..+3  x  units 4 tx..+3 and 64 + 4 tx..+3
#pragma unroll 8
    for (int k = 0; k < 4 * SEQ_UNITS; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(dzT + k * SEQ_ROWS + 4 * ty);
      const float4 b0 = *reinterpret_cast<const float4*>(Wt + k * H + 4 * tx);
      const float4 b1 = *reinterpret_cast<const float4*>(Wt + k * H + 64 + 4 * tx);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = fma2(av[i], bv[q], acc[i][q]);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(remote[hh] + (p * SEQ_CTAS + j) * SLOT + ((tx & 3) * 4 + q) * SEQ_ROWS + 4 * ty) =
            (q & 1) ? make_float4(acc[0][hh * 2 + (q >> 1)].y, acc[1][hh * 2 + (q >> 1)].y, acc[2][hh * 2 + (q >> 1)].y, acc[3][hh * 2 + (q >> 1)].y)
                    : make_float4(acc[0][hh * 2 + (q >> 1)].x, acc[1][hh * 2 + (q >> 1)].x, acc[2][hh * 2 + (q >> 1)].x, acc[3][hh * 2 + (q >> 1)].x);
    cluster.sync();                                      // all partial sums of this step have landed; dzT may be overwritten
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int src = 0; src < SEQ_CTAS; ++src) {
      const float4 v = *reinterpret_cast<const float4*>(slots + (p * SEQ_CTAS + src) * SLOT + tx * SEQ_ROWS + 4 * ty);
      sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) dh_rec[i] = sum[i];
  }
  cluster.sync();                                        // no CTA leaves while its slots may still be written
  // bias gradient: the 16 site groups of the CTA are added up through shared memory, then one atomic per (gate, unit) and cluster
#pragma unroll
  for (int g = 0; g < 4; ++g) dzT[(4 * tx + g) * SEQ_ROWS + ty] = db[g];
  __syncthreads();
  if (threadIdx.x < 4 * SEQ_UNITS) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) sum += dzT[threadIdx.x * SEQ_ROWS + q];
    atomicAdd(dbias + (threadIdx.x & 3) * H + j * SEQ_UNITS + (threadIdx.x >> 2), sum);
  }
}

// ---- slice-dense L3 (clair/model.py:225-244): per channel c, z3[b][o][c] = sum_t in[t][b][c] W3_c[t][o] + b3_c[o]; a3 = selu(z3) ----
// params of channel c: 33*30 kernel floats then 30 bias floats (the flat parameter order).  A CTA owns 32 adjacent channels - one
// per lane, so every global access is a coalesced 128-byte row of channels - with their 32 x 1020 parameters in shared memory
// ([parameter][lane], padded to 33), and a run of sites; grid = (8 channel groups, site chunks).
constexpr int L3_STRIDE = T_STEPS * L3_UNITS + L3_UNITS;      // 1020
constexpr int L3_LANES = 32, L3_PAD = 33;
constexpr int L3_SMEM = L3_STRIDE * L3_PAD * (int)sizeof(float);                  // 134,640 B
__device__ __forceinline__ void l3_stage_params(float* w_s, const float* __restrict__ P3, int c0) {
  for (int i = threadIdx.x; i < L3_STRIDE * L3_LANES; i += blockDim.x) {
    const int lane = i / L3_STRIDE, k = i - lane * L3_STRIDE;
    w_s[k * L3_PAD + lane] = P3[(size_t)(c0 + lane) * L3_STRIDE + k];
  }
  __syncthreads();
}
// 8 warps; a warp takes two sites per pass (sites_per_cta is a multiple of 16, n of 8)
__global__ void __launch_bounds__(256) l3_forward(const float* __restrict__ in, const float* __restrict__ P3, float* __restrict__ a3, int n, int sites_per_cta) {
  extern __shared__ __align__(16) float l3_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x * L3_LANES + lane;
  l3_stage_params(l3_smem, P3, blockIdx.x * L3_LANES);
  const int b_end = min(n, (int)(blockIdx.y + 1) * sites_per_cta);
  for (int b = blockIdx.y * sites_per_cta + 2 * warp; b < b_end; b += 16) {
    float2 x[T_STEPS];                                   // (site b, site b + 1)
#pragma unroll
    for (int t = 0; t < T_STEPS; ++t) x[t] = make_float2(in[((size_t)t * n + b) * 2 * H + c], in[((size_t)t * n + b + 1) * 2 * H + c]);
    for (int o = 0; o < L3_UNITS; ++o) {
      const float b3 = l3_smem[(T_STEPS * L3_UNITS + o) * L3_PAD + lane];
      float2 z = make_float2(b3, b3);
#pragma unroll
      for (int t = 0; t < T_STEPS; ++t) z = fma2(l3_smem[(t * L3_UNITS + o) * L3_PAD + lane], x[t], z);
      a3[(size_t)b * L3_K + o * 2 * H + c] = selu_f(z.x);
      a3[(size_t)(b + 1) * L3_K + o * 2 * H + c] = selu_f(z.y);
    }
  }
}
// d in[t][b][c] = sum_o dz3[b][o][c] W3_c[t][o]   (dz3 = da3 * selu'(a3) computed on the fly; da3 is overwritten with dz3)
__global__ void __launch_bounds__(256) l3_backward_input(float* __restrict__ da3, const float* __restrict__ a3, const float* __restrict__ P3,
                                                         float* __restrict__ din, int n, int sites_per_cta) {
  extern __shared__ __align__(16) float l3_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x * L3_LANES + lane;
  l3_stage_params(l3_smem, P3, blockIdx.x * L3_LANES);
  const int b_end = min(n, (int)(blockIdx.y + 1) * sites_per_cta);
  for (int b = blockIdx.y * sites_per_cta + 2 * warp; b < b_end; b += 16) {
    float2 d[L3_UNITS];                                  // (site b, site b + 1)
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) {
      const size_t at0 = (size_t)b * L3_K + o * 2 * H + c, at1 = at0 + L3_K;
      const float a0 = a3[at0], a1 = a3[at1];
      d[o] = make_float2(da3[at0] * (a0 >= 0.f ? SELU_SCALE : a0 + SELU_SCALE * SELU_ALPHA),
                         da3[at1] * (a1 >= 0.f ? SELU_SCALE : a1 + SELU_SCALE * SELU_ALPHA));
      da3[at0] = d[o].x;
      da3[at1] = d[o].y;
    }
    for (int t = 0; t < T_STEPS; ++t) {
      float2 sum = make_float2(0.f, 0.f);
#pragma unroll
      for (int o = 0; o < L3_UNITS; ++o) sum = fma2(l3_smem[(t * L3_UNITS + o) * L3_PAD + lane], d[o], sum);
      din[((size_t)t * n + b) * 2 * H + c] = sum.x;
      din[((size_t)t * n + b + 1) * 2 * H + c] = sum.y;
    }
  }
}
// dW3_c[t][o] += sum_b in[t][b][c] dz3[b][o][c],  db3_c[o] += sum_b dz3[b][o][c]  over the CTA's run of sites (G3 zeroed before).
// 11 warps: warp w holds the 3 x 30 sums of t = 3w..3w+2 (warp 0 also the 30 bias sums); the sums leave through shared memory so
// that the atomics run along a channel's 1020 contiguous parameters.  grid = (8 channel groups, site chunks)
__global__ void __launch_bounds__(352) l3_backward_weights(const float* __restrict__ in, const float* __restrict__ dz3, float* __restrict__ G3, int n,
                                                           int sites_per_cta) {
  extern __shared__ __align__(16) float l3_smem[];                   // [lane][1021]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x * L3_LANES + lane;
  const int b_end = min(n, (int)(blockIdx.y + 1) * sites_per_cta);
  float acc[3][L3_UNITS] = {}, bias[L3_UNITS] = {};
  for (int b = blockIdx.y * sites_per_cta; b < b_end; ++b) {
    float x[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) x[q] = in[((size_t)(3 * warp + q) * n + b) * 2 * H + c];
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) {
      const float d = dz3[(size_t)b * L3_K + o * 2 * H + c];
#pragma unroll
      for (int q = 0; q < 3; ++q) acc[q][o] = fmaf(x[q], d, acc[q][o]);
      if (warp == 0) bias[o] += d;
    }
  }
  constexpr int OUT_LD = L3_STRIDE + 1;
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) l3_smem[lane * OUT_LD + (3 * warp + q) * L3_UNITS + o] = acc[q][o];
  if (warp == 0) {
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) l3_smem[lane * OUT_LD + T_STEPS * L3_UNITS + o] = bias[o];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L3_STRIDE * L3_LANES; i += blockDim.x) {
    const int l = i / L3_STRIDE, k = i - l * L3_STRIDE;
    atomicAdd(G3 + (size_t)(blockIdx.x * L3_LANES + l) * L3_STRIDE + k, l3_smem[l * OUT_LD + k]);
  }
}

// ---- softmax + focal loss of one head, and its gradient w.r.t. the post-SELU logits z (clair/model.py:783-805) ----
// one thread per site; z [n][90] (all heads side by side), target [n][90]; dz written in place of z's gradient buffer; the four
// loss sums are accumulated in double.
__global__ void focal_loss_heads(const float* __restrict__ z, const float* __restrict__ target, float* __restrict__ probs, float* __restrict__ dz,
                                 double* __restrict__ loss, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  for (int k = 0; k < 4; ++k) {
    const int lo = kHeadOff[k], cnt = kHeadOff[k + 1] - lo;
    const float* zz = z + (size_t)b * N_OUT + lo;
    const float* tt = target + (size_t)b * N_OUT + lo;
    float mx = zz[0];
    for (int j = 1; j < cnt; ++j) mx = fmaxf(mx, zz[j]);
    float sum = 0.f, p[33];
    for (int j = 0; j < cnt; ++j) { p[j] = expf(zz[j] - mx); sum += p[j]; }
    double l = 0.0;
    float dp[33], dot = 0.f;
    for (int j = 0; j < cnt; ++j) {
      p[j] /= sum;
      probs[(size_t)b * N_OUT + lo + j] = p[j];
      const float t = tt[j];
      if (t > 0.f) {
        const float a = t - p[j], pc = fminf(fmaxf(p[j], 1e-8f), 1.f), lg = logf(pc);
        l -= (double)(a * a * lg);
        dp[j] = 2.f * a * lg - ((p[j] >= 1e-8f && p[j] <= 1.f) ? a * a / pc : 0.f);
      } else {
        const float q = 1.f - p[j], qc = fminf(fmaxf(q, 1e-8f), 1.f), lg = logf(qc);
        l -= (double)(p[j] * p[j] * lg);
        dp[j] = -2.f * p[j] * lg + ((q >= 1e-8f && q <= 1.f) ? p[j] * p[j] / qc : 0.f);
      }
      dot = fmaf(dp[j], p[j], dot);
    }
    for (int j = 0; j < cnt; ++j) dz[(size_t)b * N_OUT + lo + j] = p[j] * (dp[j] - dot);
    atomicAdd(&loss[k], l);
  }
}

// ---- optimiser (clair/model.py:689-694, 717-728) ---------------------------------------------------------------------------
// sum of squares of the kernels (L2 term without lambda) / of the gradients (global norm), in double
__global__ void sum_squares(const float* __restrict__ v, const uint8_t* __restrict__ is_kernel, int64_t count, double* __restrict__ out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    if (!is_kernel || is_kernel[i]) s += (double)v[i] * v[i];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
// g += lambda * w on the kernels (the gradient of lambda * ||w||^2 / 2)
__global__ void add_l2_gradient(float* __restrict__ g, const float* __restrict__ w, const uint8_t* __restrict__ is_kernel, float lambda, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count && is_kernel[i]) g[i] = fmaf(lambda, w[i], g[i]);
}
// clip by global norm, then Adam (TF defaults beta1 0.9, beta2 0.999, epsilon 1e-8; lr_t carries the bias corrections)
__global__ void adam_update(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, const double* __restrict__ sumsq,
                            float clip_norm, float lr_t, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float norm = (float)sqrt(*sumsq);
  const float scale = clip_norm / fmaxf(norm, clip_norm);
  const float gi = g[i] * scale;
  const float mi = 0.9f * m[i] + 0.1f * gi, vi = 0.999f * v[i] + 0.001f * gi * gi;
  m[i] = mi;
  v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + 1e-8f);
}

}  // namespace train
}  // namespace clairb
