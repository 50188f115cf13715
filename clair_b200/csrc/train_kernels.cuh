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
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace clairb {
namespace train {

constexpr int ROWS = 8;                  // sites per CTA of the per-step recurrent kernels
constexpr int WT = 16;                   // rows of the recurrent kernel staged in shared memory at a time
constexpr float ALPHA_DROPOUT = -1.7580993408473766f;     // clair/selu.py:43

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
  float acc[4][4] = {};
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
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
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
        if (gridDim.z > 1) atomicAdd(c, alpha * acc[i][j]);
        else *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
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
  dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
  // split the contraction when the output tiles alone cannot fill the machine (592 = 4 CTAs per SM on 148 SMs)
  int slices = 1;
  const int tiles = (int)(grid.x * grid.y);
  if (tiles < 296 && K >= 2048) slices = min(64, max(1, min(592 / tiles, K / 512)));
  int k_per_slice = ((K + slices - 1) / slices + GK - 1) / GK * GK;
  slices = (K + k_per_slice - 1) / k_per_slice;
  grid.z = slices;
  if (slices > 1) {
    scale_matrix<<<(unsigned)(((int64_t)M * N + 255) / 256), 256, 0, st>>>(C, M, N, ldc, beta);
    ++*launches;
  }
  if (!ta && !tb) sgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, k_per_slice);
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
// rows of C <- bias (before an accumulating GEMM)
__global__ void fill_rows(float* __restrict__ C, const float* __restrict__ bias, int64_t rows, int N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * N) C[i] = bias[i % N];
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

// ---- one LSTM step forward: z = pre[s] + h_{s-1} . W_h ; gates ; c_s, h_s  (LSTMBlockCell, forget_bias 0; clair/model.py:299-305) ----
// grid = n / ROWS CTAs of 128 threads (thread = hidden unit); W_h [128][512] is read once per CTA and step (L2 resident).
__global__ void __launch_bounds__(H) lstm_step_forward(const float* __restrict__ pre, const float* __restrict__ Wh, const float* __restrict__ h_prev,
                                                       const float* __restrict__ c_prev, float* __restrict__ gates, float* __restrict__ c_out,
                                                       float* __restrict__ h_out, int n) {
  __shared__ float hs[ROWS][H];
  __shared__ float ws[WT][G4];                           // WT rows of W_h at a time, loaded by the whole CTA (coalesced)
  const int u = threadIdx.x, r0 = blockIdx.x * ROWS;
  for (int r = 0; r < ROWS; ++r) hs[r][u] = h_prev[(size_t)(r0 + r) * H + u];
  float z[ROWS][4];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) z[r][g] = pre[(size_t)(r0 + r) * G4 + g * H + u];
  for (int k0 = 0; k0 < H; k0 += WT) {
    __syncthreads();
    for (int i = u; i < WT * G4 / 4; i += H)
      reinterpret_cast<float4*>(&ws[0][0])[i] = reinterpret_cast<const float4*>(Wh + (size_t)k0 * G4)[i];
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WT; ++kk) {
      float w[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) w[g] = ws[kk][g * H + u];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float hv = hs[r][k0 + kk];
#pragma unroll
        for (int g = 0; g < 4; ++g) z[r][g] = fmaf(hv, w[g], z[r][g]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const size_t row = (size_t)(r0 + r);
    const float i = sigmoidf_(z[r][0]), g = tanhf(z[r][1]), f = sigmoidf_(z[r][2]), o = sigmoidf_(z[r][3]);
    const float c = g * i + c_prev[row * H + u] * f;
    gates[row * G4 + u] = i; gates[row * G4 + H + u] = g; gates[row * G4 + 2 * H + u] = f; gates[row * G4 + 3 * H + u] = o;
    c_out[row * H + u] = c;
    h_out[row * H + u] = tanhf(c) * o;
  }
}

// ---- one LSTM step backward: dh = dh_out[s] + dh_rec ; gate gradients dZ[s] ; dc_{s-1} ; dh_rec <- dZ[s] . W_h^T ----
// WhT [512][128] is W_h transposed (rows = gate columns), so the second half reads it coalesced.
__global__ void __launch_bounds__(H) lstm_step_backward(const float* __restrict__ dh_out, float* __restrict__ dh_rec, float* __restrict__ dc,
                                                        const float* __restrict__ gates, const float* __restrict__ c, const float* __restrict__ c_prev,
                                                        const float* __restrict__ WhT, float* __restrict__ dZ, int n) {
  __shared__ float dz_s[ROWS][G4];
  const int u = threadIdx.x, r0 = blockIdx.x * ROWS;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const size_t row = (size_t)(r0 + r);
    const float i = gates[row * G4 + u], g = gates[row * G4 + H + u], f = gates[row * G4 + 2 * H + u], o = gates[row * G4 + 3 * H + u];
    const float tc = tanhf(c[row * H + u]);
    const float dh = dh_out[row * H + u] + dh_rec[row * H + u];
    const float dcs = dc[row * H + u] + dh * o * (1.f - tc * tc);
    const float dzi = dcs * g * i * (1.f - i), dzg = dcs * i * (1.f - g * g), dzf = dcs * c_prev[row * H + u] * f * (1.f - f), dzo = dh * tc * o * (1.f - o);
    dc[row * H + u] = dcs * f;
    dz_s[r][u] = dzi; dz_s[r][H + u] = dzg; dz_s[r][2 * H + u] = dzf; dz_s[r][3 * H + u] = dzo;
    dZ[row * G4 + u] = dzi; dZ[row * G4 + H + u] = dzg; dZ[row * G4 + 2 * H + u] = dzf; dZ[row * G4 + 3 * H + u] = dzo;
  }
  __shared__ float wts[4 * WT][H];                        // 64 rows of W_h^T at a time
  float acc[ROWS] = {};
  for (int k0 = 0; k0 < G4; k0 += 4 * WT) {
    __syncthreads();                                     // also orders the dz_s writes above before the first reads
    for (int i = u; i < 4 * WT * H / 4; i += H)
      reinterpret_cast<float4*>(&wts[0][0])[i] = reinterpret_cast<const float4*>(WhT + (size_t)k0 * H)[i];
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 4 * WT; ++kk) {
      const float w = wts[kk][u];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(dz_s[r][k0 + kk], w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) dh_rec[(size_t)(r0 + r) * H + u] = acc[r];
}

// ---- slice-dense L3 (clair/model.py:225-244): per channel c, z3[b][o][c] = sum_t in[t][b][c] W3_c[t][o] + b3_c[o]; a3 = selu(z3) ----
// params of channel c: 33*30 kernel floats then 30 bias floats (the flat parameter order).  grid = (256 channels, ceil(n / 128)).
constexpr int L3_STRIDE = T_STEPS * L3_UNITS + L3_UNITS;      // 1020
__global__ void __launch_bounds__(128) l3_forward(const float* __restrict__ in, const float* __restrict__ P3, float* __restrict__ a3, int n) {
  __shared__ float w[L3_STRIDE];
  const int c = blockIdx.x, b = blockIdx.y * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < L3_STRIDE; i += 128) w[i] = P3[(size_t)c * L3_STRIDE + i];
  __syncthreads();
  if (b >= n) return;
  float x[T_STEPS];
#pragma unroll
  for (int t = 0; t < T_STEPS; ++t) x[t] = in[((size_t)t * n + b) * 2 * H + c];
  for (int o = 0; o < L3_UNITS; ++o) {
    float z = w[T_STEPS * L3_UNITS + o];
#pragma unroll
    for (int t = 0; t < T_STEPS; ++t) z = fmaf(x[t], w[t * L3_UNITS + o], z);
    a3[(size_t)b * L3_K + o * 2 * H + c] = selu_f(z);
  }
}
// d in[t][b][c] = sum_o dz3[b][o][c] W3_c[t][o]   (dz3 = da3 * selu'(a3) computed on the fly; da3 is overwritten with dz3)
__global__ void __launch_bounds__(128) l3_backward_input(float* __restrict__ da3, const float* __restrict__ a3, const float* __restrict__ P3,
                                                         float* __restrict__ din, int n) {
  __shared__ float w[L3_STRIDE];
  const int c = blockIdx.x, b = blockIdx.y * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < L3_STRIDE; i += 128) w[i] = P3[(size_t)c * L3_STRIDE + i];
  __syncthreads();
  if (b >= n) return;
  float dz[L3_UNITS];
#pragma unroll
  for (int o = 0; o < L3_UNITS; ++o) {
    const size_t at = (size_t)b * L3_K + o * 2 * H + c;
    const float a = a3[at];
    dz[o] = da3[at] * (a >= 0.f ? SELU_SCALE : a + SELU_SCALE * SELU_ALPHA);
    da3[at] = dz[o];
  }
  for (int t = 0; t < T_STEPS; ++t) {
    float s = 0.f;
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) s = fmaf(dz[o], w[t * L3_UNITS + o], s);
    din[((size_t)t * n + b) * 2 * H + c] = s;
  }
}
// dW3_c[t][o] = sum_b in[t][b][c] dz3[b][o][c],  db3_c[o] = sum_b dz3[b][o][c]      grid = 256 channels, 1024 threads (>= 990 + 30)
__global__ void __launch_bounds__(1024) l3_backward_weights(const float* __restrict__ in, const float* __restrict__ dz3, float* __restrict__ G3, int n) {
  const int c = blockIdx.x, i = threadIdx.x;
  if (i >= L3_STRIDE) return;
  const bool bias = i >= T_STEPS * L3_UNITS;
  const int t = bias ? 0 : i / L3_UNITS, o = bias ? i - T_STEPS * L3_UNITS : i % L3_UNITS;
  float s = 0.f;
  for (int b = 0; b < n; ++b) {
    const float d = dz3[(size_t)b * L3_K + o * 2 * H + c];
    s = bias ? s + d : fmaf(in[((size_t)t * n + b) * 2 * H + c], d, s);
  }
  G3[(size_t)c * L3_STRIDE + i] = s;
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
// [K][N] -> [N][K]
__global__ void transpose_matrix(const float* __restrict__ in, float* __restrict__ out, int K, int N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)K * N) out[(i % N) * K + i / N] = in[i];
}

}  // namespace train
}  // namespace clairb
