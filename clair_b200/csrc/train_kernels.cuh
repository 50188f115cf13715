// Training step of the Clair network on the device: forward in training mode, focal loss, backward
// through the dense trunk, the slice-dense layer and both BiLSTMs (BPTT over 33 steps), global-norm clip, Adam.
//
// Reference: clair/model.py
//   forward in training phase   :400-622, tf.layers.dropout :434-459, selu.dropout_selu clair/selu.py:43-74
//   focal loss                  :783-805        L2 :689-694        total :696-709
//   clip_by_global_norm(5.0) + AdamOptimizer    :717-728
// Every result is fp32-grade - the reference trains in fp32 (float_type is forced to tf.float32, :165-170) and the parity tests
// compare gradients with a float64 autograd restatement (oracle/train_oracle.py).  The 33 steps of an LSTM direction are ONE
// launch (thread-block clusters, h exchanged through distributed shared memory); every large contraction (recurrences, input
// projections, weight gradients, L4) runs on the tensor cores as 3xTF32 mma.sync with the slabs accumulated in fp32 outside
// the MMA; the small ones (L5, heads) and the elementwise work on the CUDA cores.  The tcgen05 kernels of the inference path
// do not save the gate activations BPTT needs and are not used here.
//
// Layouts (row-major fp32): everything in global memory is time-major [33][n][...] in TIME order for both directions; the
// sequence kernels walk processing steps s = 0..32 and address time t = s (fw) or 32 - s (bw), so no array is ever reversed or
// copied per direction.  h and c of a direction live in 35 slabs, time t at slab t + 1, slabs 0 and 34 zero: the state before
// a step is slab t (fw) / t + 2 (bw).  Gate columns in TF order i, c(candidate), f, o.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace clairb {
namespace train {

constexpr int ROWS = 8;                  // batches are padded to a multiple of 8 sites (zero rows)
constexpr float ALPHA_DROPOUT = -1.7580993408473766f;     // clair/selu.py:43

// Two independent fp32 FMAs in one instruction (FFMA2, scalar a broadcast over the pair): c + a * b, each half rounded like fmaf.
// Halves the instruction count of an outer-product inner loop; measured here it is throughput-neutral against plain FFMA
// (tools/probes/mma_rate.cu: FFMA 72 TFLOP/s with two register operands, 55 as an outer product; the large contractions moved
// to the tensor cores instead, sgemm_big).
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
// fragment loads.  Three shared-memory stages filled by cp.async: two tiles are in flight while one is multiplied.
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
// One k = 8 slab of NT products that share the A fragment, with fp32-grade accuracy, added to acc.  The tensor core adds into its
// accumulator with truncation (round toward zero): accumulating a long contraction inside the MMA biases every sum the same way,
// and the 33-step recurrence compounds it - 16 % on one LSTM weight gradient of the Adam parity test.  So each slab is accumulated
// from zero (correction terms first, three MMAs per tile, term-major so that NT independent MMAs lie between two dependent ones)
// and joins the running sum through a rounded FADD.
template <int NT>
__device__ __forceinline__ void mma_3xtf32(float (&acc)[NT][4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[NT][2],
                                           const uint32_t (&bl)[NT][2]) {
  float v[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(v[j][0]), "=f"(v[j][1]), "=f"(v[j][2]), "=f"(v[j][3])
                 : "r"(al[0]), "r"(al[1]), "r"(al[2]), "r"(al[3]), "r"(bh[j][0]), "r"(bh[j][1]), "f"(0.f));
#pragma unroll
  for (int j = 0; j < NT; ++j) mma_tf32(v[j], ah, bl[j]);
#pragma unroll
  for (int j = 0; j < NT; ++j) mma_tf32(v[j], ah, bh[j]);
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] += v[j][q];
}
constexpr int BIG_STAGES = 3;
constexpr int BIG_SMEM = BIG_STAGES * 2 * STAGE_FLOATS * (int)sizeof(float);      // 61,440 B
// 16 bytes global -> shared without a register hop; `ok` false writes zeros (no bytes are read)
__device__ __forceinline__ void cp_async16_or_zero(float* smem_dst, const float* gmem_src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(ok ? 16 : 0));
}
template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) sgemm_big(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float beta, float* __restrict__ C, int ldc, int k_per_slice) {
  constexpr bool A_K = TA, B_K = !TB;                    // operand stored with k outermost in global memory
  extern __shared__ __align__(16) float big_smem[];      // [stage][A tile | B tile]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_slice;
  const int k_end = min(K, k_begin + k_per_slice);
  const int iters = (k_end - k_begin + BK - 1) / BK;
  // three stages of cp.async: tile it + 2 is requested while tile it is multiplied; one barrier per tile
  auto request = [&](int it) {
    if (it < iters) {
      const int k0 = k_begin + it * BK;
      float* as = big_smem + (it % BIG_STAGES) * 2 * STAGE_FLOATS;
      float* bs = as + STAGE_FLOATS;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int f = tid + 256 * p;
        if (A_K) {
          const int k = f >> 5, m = (f & 31) * 4;
          const bool ok = m0 + m < M && k0 + k < k_end;
          cp_async16_or_zero(as + k * LD_K + m, ok ? A + (size_t)(k0 + k) * lda + m0 + m : A, ok);
        } else {
          const int m = f >> 2, k = (f & 3) * 4;
          const bool ok = m0 + m < M && k0 + k < k_end;
          cp_async16_or_zero(as + m * LD_R + k, ok ? A + (size_t)(m0 + m) * lda + k0 + k : A, ok);
        }
        if (B_K) {
          const int k = f >> 5, nn = (f & 31) * 4;
          const bool ok = n0 + nn < N && k0 + k < k_end;
          cp_async16_or_zero(bs + k * LD_K + nn, ok ? B + (size_t)(k0 + k) * ldb + n0 + nn : B, ok);
        } else {
          const int nn = f >> 2, k = (f & 3) * 4;
          const bool ok = n0 + nn < N && k0 + k < k_end;
          cp_async16_or_zero(bs + nn * LD_R + k, ok ? B + (size_t)(n0 + nn) * ldb + k0 + k : B, ok);
        }
      }
    }
    cp_async_commit();                                   // an empty group when there is nothing left keeps the counting uniform
  };
  float acc[4][4][4] = {};                               // [m tile][n tile][c0..c3]
  request(0);
  request(1);
  for (int it = 0; it < iters; ++it) {
    cp_async_wait<1>();                                  // tile it has landed (this thread's part)
    __syncthreads();                                     // ... everybody's part; and everybody is done with tile it - 1
    request(it + 2);                                     // into the buffer tile it - 1 occupied
    const float* as = big_smem + (it % BIG_STAGES) * 2 * STAGE_FLOATS;
    const float* bs = as + STAGE_FLOATS;
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
        mma_3xtf32<4>(acc[i], ah, al, bh, bl);
      }
    }
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

// c_zeroed: C is known to hold zeros (a gradient block behind the step's memset): a split contraction with beta = 0 then needs no
// scaling pass in front of its atomics.
inline void gemm(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc,
                 cudaStream_t st, int64_t* launches, bool c_zeroed = false) {
  // the tensor-core tile for everything large; a long contraction is worth it even when only a quarter of the tile's rows exist
  // (layer 1's dW_x: 32 x 512 over 33 n rows)
  const bool big = (M >= 96 || (M >= 32 && K >= 4096)) && N >= 96 && !((M | N | K | lda | ldb | ldc) & 3) &&
                   !(((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15);
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
  if (slices > 1 && !(c_zeroed && beta == 0.f)) {
    scale_matrix<<<(unsigned)(((int64_t)M * N + 255) / 256), 256, 0, st>>>(C, M, N, ldc, beta);
    ++*launches;
  }
  if (big) {
    if (!ta && !tb) sgemm_big<false, false><<<grid, 256, BIG_SMEM, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else if (ta && !tb) sgemm_big<true, false><<<grid, 256, BIG_SMEM, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else if (!ta && tb) sgemm_big<false, true><<<grid, 256, BIG_SMEM, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
    else sgemm_big<true, true><<<grid, 256, BIG_SMEM, st>>>(M, N, K, A, lda, B, ldb, beta, C, ldc, k_per_slice);
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

// ---- the 33 steps of one LSTM direction, forward: z_s = x_s W_x (pre) + b + h_{s-1} . W_h ; gates ; c_s, h_s ---------------------
// (LSTMBlockCell, forget_bias 0; clair/model.py:299-305).  One thread-block CLUSTER of 8 CTAs carries 64 sites through all 33
// steps.  CTA j owns hidden units 16j..16j+15: it keeps their 64 gate columns of W_h and the whole h_{s-1} of the 64 sites
// ([unit][site], row stride 72) in shared memory and computes the [64 x 128] . [128 x 64] gate tile as 3xTF32 mma.sync (see
// sgemm_big).  Its 16 new h columns are one contiguous 4.6 KB run of the next step's buffer: one thread sends that run to the same
// place in the 7 other CTAs with cp.async.bulk over distributed shared memory, each copy completing bytes on the RECEIVER's
// mbarrier - the receivers wait on their own barrier, there is no cluster-wide barrier in the loop.  (The first version pushed h
// with st.shared::cluster and ran cluster.sync every step: 1.0 + 0.8 us of a 7.8 us step, the FFMA contraction another 4.3.)
// Why one buffer pair is enough: a CTA sends h_s only after its own step-s contraction, i.e. after it has every piece of h_{s-1};
// so when h_s has arrived from everybody, everybody is done reading h_{s-1}'s buffer and has received what was sent from it.
// Warp (mp, ug) owns sites 32mp..+31 x units 4ug..+3: 2 x 2 MMA tiles whose 8 columns are ordered (unit, gate pair), so that lane
// (g, tig) ends up with all four gates of unit 4ug+tig for sites g, g+8, g+16, g+24 of its half.
constexpr int SEQ_ROWS = 64, SEQ_CTAS = 8, SEQ_UNITS = H / SEQ_CTAS;      // 64 sites per cluster, 16 units per CTA
constexpr int SEQ_LD = SEQ_ROWS + 8;                                      // 72 = 8 (mod 32): conflict-free fragment loads
constexpr int SEQ_SLICE_BYTES = SEQ_UNITS * SEQ_LD * (int)sizeof(float);  // 4608: the h columns (or dh partial sums) of one CTA
constexpr int seq_fwd_smem(int rb) { return (H * SEQ_LD + 2 * H * (rb + 8)) * (int)sizeof(float) + 16; }      // 110,608 B (64 sites), 77,840 B (32)
constexpr int SEQ_BWD_LDW = H + 8;                                        // 136
constexpr int seq_bwd_smem(int rb) {                                      // 163,864 B (64 sites), 106,520 B (32)
  return (4 * SEQ_UNITS * SEQ_BWD_LDW + 4 * SEQ_UNITS * (rb + 8) + H * (rb + 8) + 2 * SEQ_CTAS * SEQ_UNITS * (rb + 8)) * (int)sizeof(float) + 24;
}

// local shared memory -> the same offset in CTA `rank` of the cluster, completing `bytes` on that CTA's mbarrier
__device__ __forceinline__ void bulk_to_cta(const void* src, void* dst_same_offset, uint64_t* bar_same_offset, uint32_t rank, uint32_t bytes) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::map_to_cta(tc::smem_u32(dst_same_offset), rank)),
               "r"(tc::smem_u32(src)), "r"(bytes), "r"(tc::map_to_cta(tc::smem_u32(bar_same_offset), rank))
               : "memory");
}

// RB = sites per cluster: 64 (a warp owns 2 x 2 MMA tiles) or 32 (1 x 2; half the work and half the exchange per CTA and step, so
// that two CTAs - of different clusters - share an SM and one computes while the other waits for its h to arrive).
template <int RB>
__global__ void __cluster_dims__(SEQ_CTAS, 1, 1) __launch_bounds__(256)
lstm_seq_forward(const float* __restrict__ pre, const float* __restrict__ Wh, const float* __restrict__ bias, float* __restrict__ gates,
                 float* __restrict__ cbuf, float* __restrict__ hbuf, float* __restrict__ lout, int col0, int n, int reverse) {
  constexpr int MI = RB / 32, LD = RB + 8, SLICE_BYTES = SEQ_UNITS * LD * (int)sizeof(float);      // LD = 8 (mod 32) for both
  extern __shared__ __align__(16) float seq_smem[];
  float* Ws = seq_smem;                                  // [128 units of h_{s-1}][72: 64 local gate columns]
  float* hT = Ws + H * SEQ_LD;                           // [2][128 units][LD: RB sites]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hT + 2 * H * LD);      // [2]: "h of this buffer has arrived from the 7 other CTAs"
  const int j = (int)tc::cluster_ctarank();
  const int r0 = (int)(blockIdx.x / SEQ_CTAS) * RB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const int mp = warp >> 2, ug = warp & 3;
  const int unit = j * SEQ_UNITS + 4 * ug + tig;         // this lane's hidden unit
  // local gate column lc = 16 ug' + 8 nt + nin  <->  unit 4 ug' + (nin >> 1), gate 2 nt + (nin & 1)
  for (int i = threadIdx.x; i < H * 4 * SEQ_UNITS; i += 256) {
    const int k = i >> 6, lc = i & 63;
    const int u = 4 * (lc >> 4) + ((lc & 7) >> 1), gate = 2 * ((lc >> 3) & 1) + (lc & 1);
    Ws[k * SEQ_LD + lc] = Wh[(size_t)k * G4 + gate * H + j * SEQ_UNITS + u];
  }
  for (int i = threadIdx.x; i < H * LD; i += 256) hT[i] = 0.f;        // h_0 = 0
  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::fence_barrier_init();
  }
  float bz[2][2];                                        // bias of (gate 2 nt + e) of this lane's unit
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) bz[nt][e] = bias[(2 * nt + e) * H + unit];
  // accumulators acc[mi][nt][2 half + e]: site 32 mp + 16 mi + 8 half + g, gate 2 nt + e; they start as pre + bias
  float acc[MI][2][4], c[MI][2] = {};
  auto fetch = [&](int s, float (&z)[MI][2][4]) {
    const int t = reverse ? T_STEPS - 1 - s : s;         // everything in global memory is in TIME order
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = r0 + 16 * MI * mp + 16 * mi + 8 * half + g;
        const float* q = pre + ((size_t)t * n + row) * G4 + unit;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) z[mi][nt][2 * half + e] = row < n ? q[(2 * nt + e) * H] + bz[nt][e] : 0.f;
      }
  };
  fetch(0, acc);
  tc::cluster_sync_all();                                // every CTA of the cluster runs and has its barriers initialised
  for (int s = 0; s < T_STEPS; ++s) {
    const float* hc = hT + (s & 1) * H * LD;
    float* hn = hT + ((s + 1) & 1) * H * LD;
    float nxt[MI][2][4];                                  // the next step's input projection, requested before the contraction
    if (s + 1 < T_STEPS) fetch(s + 1, nxt);
    if (s > 0) tc::mbar_wait(&bars[s & 1], ((s - 1) >> 1) & 1);
#pragma unroll 4
    for (int kk = 0; kk < H; kk += 8) {
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        split_tf32(Ws[(kk + tig) * SEQ_LD + 16 * ug + 8 * nt + g], bh[nt][0], bl[nt][0]);
        split_tf32(Ws[(kk + tig + 4) * SEQ_LD + 16 * ug + 8 * nt + g], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int mb = 16 * MI * mp + 16 * mi + g;
        uint32_t ah[4], al[4];
        split_tf32(hc[(kk + tig) * LD + mb], ah[0], al[0]);
        split_tf32(hc[(kk + tig) * LD + mb + 8], ah[1], al[1]);
        split_tf32(hc[(kk + tig + 4) * LD + mb], ah[2], al[2]);
        split_tf32(hc[(kk + tig + 4) * LD + mb + 8], ah[3], al[3]);
        mma_3xtf32<2>(acc[mi], ah, al, bh, bl);
      }
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int rl = 16 * MI * mp + 16 * mi + 8 * half + g;
        const float ig = sigmoidf_(acc[mi][0][2 * half]), gg = tanhf(acc[mi][0][2 * half + 1]);
        const float fg = sigmoidf_(acc[mi][1][2 * half]), og = sigmoidf_(acc[mi][1][2 * half + 1]);
        const float cv = gg * ig + c[mi][half] * fg;
        const float hv = tanhf(cv) * og;
        c[mi][half] = cv;
        hn[unit * LD + rl] = hv;
        if (r0 + rl < n) {
          const size_t r = (size_t)(reverse ? T_STEPS - 1 - s : s) * n + r0 + rl;
          gates[r * G4 + unit] = ig; gates[r * G4 + H + unit] = gg; gates[r * G4 + 2 * H + unit] = fg; gates[r * G4 + 3 * H + unit] = og;
          cbuf[(r + n) * H + unit] = cv;                 // slab t + 1 of 35: slabs 0 and 34 stay zero (the state before the first step)
          hbuf[(r + n) * H + unit] = hv;
          lout[r * 2 * H + col0 + unit] = hv;            // the layer's output [t][site][fw | bw]
        }
      }
    if (s + 1 < T_STEPS) {
      tc::fence_proxy_async();                           // the h columns just written are read by the bulk copies
      __syncthreads();
      if (threadIdx.x == 0) {
        tc::mbar_expect_tx(&bars[(s + 1) & 1], (SEQ_CTAS - 1) * SLICE_BYTES);
        float* mine = hn + j * SEQ_UNITS * LD;
#pragma unroll 1
        for (int d = 1; d < SEQ_CTAS; ++d) bulk_to_cta(mine, mine, &bars[(s + 1) & 1], (uint32_t)((j + d) & (SEQ_CTAS - 1)), SLICE_BYTES);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[mi][nt][q] = nxt[mi][nt][q];
    }
  }
  tc::cluster_sync_all();                                // nobody leaves while a copy from or into its shared memory may be in flight
}

// ---- the 33 steps of one LSTM direction, backward: dh = dh_out[s] + dh_rec ; gate gradients dZ[s] ; dc_{s-1} ; dh_rec <- dZ[s] . W_h^T
// Same cluster shape and the same exchange.  CTA j produces the 64 gate-gradient columns of its 16 units (thread (ty, tx): sites
// RPT ty..+RPT-1 of unit tx), multiplies them by its 64 rows of W_h^T - [RB x 64] . [64 x 128] on the tensor cores, a partial sum
// of dh_rec for ALL 128 units - and sends the 16 columns each other CTA owns into slot j of that CTA; when its own 7 slots have
// arrived it adds the 8 partial sums in a fixed order.  A slot copy reads the sender's partial sums asynchronously: every
// receiver acknowledges on the sender's `ack` barrier once its slots are complete, and the sender waits for the 7
// acknowledgements of a step before it overwrites the partial sums in the next (one step later: the wait is never felt).
// With RB = 32 sites per cluster the kernel needs 104 KB of shared memory and two CTAs share an SM, as in the forward kernel.
// Also accumulates the bias gradient.
template <int RB>
__global__ void __cluster_dims__(SEQ_CTAS, 1, 1) __launch_bounds__(256)
lstm_seq_backward(const float* __restrict__ dlout, int col0, const float* __restrict__ gates, const float* __restrict__ cbuf,
                  const float* __restrict__ Wh, float* __restrict__ dZ, float* __restrict__ dbias, int n, int reverse) {
  constexpr int MI = RB / 32, LD = RB + 8, RPT = RB / 16, SLOT = SEQ_UNITS * LD, SLICE_BYTES = SLOT * (int)sizeof(float);
  extern __shared__ __align__(16) float seq_smem[];
  float* Wt = seq_smem;                                  // [64 local gate columns][136: 128 units of h_{s-1}]
  float* dzs = Wt + 4 * SEQ_UNITS * SEQ_BWD_LDW;         // [64 local gate columns][LD: RB sites]
  float* part = dzs + 4 * SEQ_UNITS * LD;                // [128 units][LD]   this CTA's partial sums of dh_rec
  float* slots = part + H * LD;                          // [2][8 source CTAs][16 units][LD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 2 * SEQ_CTAS * SLOT);     // [0], [1]: slots of that parity complete; [2]: ack
  const int j = (int)tc::cluster_ctarank();
  const int r0 = (int)(blockIdx.x / SEQ_CTAS) * RB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const int mp = warp >> 2, nq = warp & 3;               // MMA role: sites 16 MI mp..  x  units 32 nq..+31 of the partial sums
  const int unit = j * SEQ_UNITS + tx;
  const int row = r0 + RPT * ty;
  const bool live = row < n;                             // n is a multiple of 8 >= RPT: a thread's sites are in or out together
  for (int i = threadIdx.x; i < 4 * SEQ_UNITS * H; i += 256) {
    const int lc = i >> 7, k = i & 127;                  // local gate column lc = 4 unit + gate
    Wt[lc * SEQ_BWD_LDW + k] = Wh[(size_t)k * G4 + (lc & 3) * H + j * SEQ_UNITS + (lc >> 2)];
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_init(&bars[2], SEQ_CTAS - 1);
    tc::fence_barrier_init();
  }
  float dc[RPT] = {}, dh_rec[RPT] = {};
  float db[4] = {0.f, 0.f, 0.f, 0.f};                   // bias gradient of this unit's four gates over the thread's sites and all steps
  float gi[RPT][4], cs[RPT], cp[RPT], dho[RPT];
  auto fetch = [&](int s) {
    const int t = reverse ? T_STEPS - 1 - s : s;         // time of processing step s; the step before it is time t -+ 1
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const size_t r = (size_t)t * n + row + i;
#pragma unroll
      for (int q = 0; q < 4; ++q) gi[i][q] = live ? gates[r * G4 + q * H + unit] : 0.f;
      cs[i] = live ? cbuf[(r + n) * H + unit] : 0.f;                                   // slab t + 1
      cp[i] = live ? cbuf[(r + (reverse ? 2 * (size_t)n : 0)) * H + unit] : 0.f;       // slab t + 2 (bw) or t (fw)
      dho[i] = live ? dlout[r * 2 * H + col0 + unit] : 0.f;
    }
  };
  fetch(T_STEPS - 1);
  tc::cluster_sync_all();
  for (int s = T_STEPS - 1; s >= 0; --s) {
    const int p = s & 1;
    float dz[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const float ig = gi[i][0], gg = gi[i][1], fg = gi[i][2], og = gi[i][3];
      const float tch = tanhf(cs[i]);
      const float dh = dho[i] + dh_rec[i];
      const float dcs = dc[i] + dh * og * (1.f - tch * tch);
      dz[i][0] = dcs * gg * ig * (1.f - ig);
      dz[i][1] = dcs * ig * (1.f - gg * gg);
      dz[i][2] = dcs * cp[i] * fg * (1.f - fg);
      dz[i][3] = dh * tch * og * (1.f - og);
      dc[i] = dcs * fg;
      if (live) {
        const size_t r = (size_t)(reverse ? T_STEPS - 1 - s : s) * n + row + i;
#pragma unroll
        for (int q = 0; q < 4; ++q) { dZ[r * G4 + q * H + unit] = dz[i][q]; db[q] += dz[i][q]; }
      }
    }
    if (s == 0) break;                                   // dh_rec of step -1 is not needed
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < RPT; ++i) dzs[(4 * tx + q) * LD + RPT * ty + i] = dz[i][q];
    __syncthreads();
    fetch(s - 1);                                        // independent of the recurrence: in flight during the contraction
    float acc[MI][4][4] = {};                            // [mi][nt][2 half + e]: site 16 MI mp + 16 mi + 8 half + g, unit 32 nq + 8 nt + 2 tig + e
#pragma unroll 2
    for (int kk = 0; kk < 4 * SEQ_UNITS; kk += 8) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        split_tf32(Wt[(kk + tig) * SEQ_BWD_LDW + 32 * nq + 8 * nt + g], bh[nt][0], bl[nt][0]);
        split_tf32(Wt[(kk + tig + 4) * SEQ_BWD_LDW + 32 * nq + 8 * nt + g], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int mb = 16 * MI * mp + 16 * mi + g;
        uint32_t ah[4], al[4];
        split_tf32(dzs[(kk + tig) * LD + mb], ah[0], al[0]);
        split_tf32(dzs[(kk + tig) * LD + mb + 8], ah[1], al[1]);
        split_tf32(dzs[(kk + tig + 4) * LD + mb], ah[2], al[2]);
        split_tf32(dzs[(kk + tig + 4) * LD + mb + 8], ah[3], al[3]);
        mma_3xtf32<4>(acc[mi], ah, al, bh, bl);
      }
    }
    // the slot copies of the previous exchange have read `part`: all 7 receivers said so
    if (s < T_STEPS - 1) tc::mbar_wait_cluster(&bars[2], (T_STEPS - 2 - s) & 1);
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) part[(32 * nq + 8 * nt + 2 * tig + (q & 1)) * LD + 16 * MI * mp + 16 * mi + 8 * (q >> 1) + g] = acc[mi][nt][q];
    tc::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc::mbar_expect_tx(&bars[p], (SEQ_CTAS - 1) * SLICE_BYTES);
      float* slot = slots + (p * SEQ_CTAS + j) * SLOT;   // slot j of the receiver
#pragma unroll 1
      for (int d = 1; d < SEQ_CTAS; ++d) {
        const int to = (j + d) & (SEQ_CTAS - 1);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tc::map_to_cta(tc::smem_u32(slot), (uint32_t)to)),
                     "r"(tc::smem_u32(part + to * SLOT)), "r"((uint32_t)SLICE_BYTES), "r"(tc::map_to_cta(tc::smem_u32(&bars[p]), (uint32_t)to))
                     : "memory");
      }
    }
    tc::mbar_wait(&bars[p], ((T_STEPS - 1 - s) >> 1) & 1);
    if (threadIdx.x == 0) {                              // my 7 slots are complete: their senders may reuse their partial sums
#pragma unroll 1
      for (int d = 1; d < SEQ_CTAS; ++d) tc::mbar_arrive_cluster(tc::map_to_cta(tc::smem_u32(&bars[2]), (uint32_t)((j + d) & (SEQ_CTAS - 1))));
    }
    float sum[RPT] = {};
#pragma unroll
    for (int src = 0; src < SEQ_CTAS; ++src) {
      const float* from = (src == j ? part + j * SLOT : slots + (p * SEQ_CTAS + src) * SLOT) + tx * LD + RPT * ty;
#pragma unroll
      for (int i = 0; i < RPT; ++i) sum[i] += from[i];
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) dh_rec[i] = sum[i];
  }
  tc::cluster_sync_all();                                // nobody leaves while a copy from or into its shared memory may be in flight
  // bias gradient: the 16 site groups of the CTA are added up through shared memory, then one atomic per (gate, unit) and cluster
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) dzs[(4 * tx + q) * LD + ty] = db[q];
  __syncthreads();
  if (threadIdx.x < 4 * SEQ_UNITS) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) sum += dzs[threadIdx.x * LD + q];
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

// out = in[0] + in[1] + in[2] + in[3] (four blocks of `count` floats back to back): the heads' shares of d(a4)
__global__ void sum_four(const float* __restrict__ in, int64_t count, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (in[i] + in[count + i]) + (in[2 * count + i] + in[3 * count + i]);
}

// the four focal-loss sums as floats behind a data-parallel caller's gradient buffer: they travel with the gradient all-reduce
__global__ void store_loss_sums(const double* __restrict__ d_loss, float* __restrict__ out) {
  if (threadIdx.x < 4) out[threadIdx.x] = (float)d_loss[threadIdx.x];
}

}  // namespace train
}  // namespace clairb
