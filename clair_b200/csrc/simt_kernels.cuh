// fp32 CUDA-core (SIMT) kernels for the whole forward graph.
//
// These are the first correct device path and stay as the on-device cross-check for the
// tensor-core kernels (tc_kernels.cuh): plain fp32 FMA accumulation, libm-accurate expf/tanhf.
// Activations live in "feature-major, site-minor" planes: plane[feature][padded site], so every
// kernel reads and writes 128-byte-coalesced rows of consecutive sites.
//
// Reference semantics: clair/model.py:400-622, clair/selu.py:26-30 (see oracle/clair_oracle.py).
#pragma once
#include "common.cuh"

namespace clairb {
namespace simt {

// ---------------------------------------------------------------------------------------------
// prep: x[n][33*32] (f32 or i16, site-major as clair/utils.py:95 reshapes it) -> xT[33*32][np]
// (model.py:403-418: reshape to [B,33,32] and go time-major).  Padding rows are zero-filled.
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256)
prep_input(const TIn* __restrict__ x, float* __restrict__ xT, SiteMap sm) {
  __shared__ float tile[32][SITE_ELEMS / 4 + 1];   // 32 sites x 264 elements (+1 pad)
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int e0 = blockIdx.y * (SITE_ELEMS / 4);    // quarter of the 1056 elements
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < 32; s += 8) {
    int64_t r = sm.real_row(p0 + s);
    for (int e = lane; e < SITE_ELEMS / 4; e += 32)
      tile[s][e] = r >= 0 ? (float)x[r * SITE_ELEMS + e0 + e] : 0.f;
  }
  __syncthreads();
  for (int e = warp; e < SITE_ELEMS / 4; e += 8)
    xT[(int64_t)(e0 + e) * sm.np + p0 + lane] = tile[lane][e];
}

// ---------------------------------------------------------------------------------------------
// One bidirectional LSTM layer (model.py:265-312): grid = (np/64, 2 directions).  A CTA owns 64
// sites of one direction for all 33 steps: [x_t ; h_{t-1}] sits in shared memory, the
// [(FIN+128) x 512] kernel streams through a double-buffered 16-row slab ring from L2 each step,
// c stays in registers.  Gate columns are pre-permuted on the host to
// colp = (u/64)*256 + (u%64)*4 + gate  (gate order i, c, f, o as TF's LSTMBlockCell).
// ---------------------------------------------------------------------------------------------
constexpr int LSTM_TM = 64;
constexpr int LSTM_TMP = 68;   // padded row pitch of the A tile (floats)
constexpr int LSTM_KS = 16;
constexpr int LSTM_THREADS = 512;

template <int FIN>
constexpr size_t lstm_smem_bytes() {
  return ((size_t)(FIN + H) * LSTM_TMP + 2 * LSTM_KS * G4) * sizeof(float);
}

template <int FIN>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_layer(const float* __restrict__ in,     // [33][FIN][np]
           const float* __restrict__ Wp,     // [2][FIN+128][512] permuted columns
           const float* __restrict__ bp,     // [2][512] permuted
           float* __restrict__ out,          // [33][256][np]  (fw -> features 0..127, bw -> 128..255)
           int64_t np) {
  constexpr int K = FIN + H;
  constexpr int NSLAB = K / LSTM_KS;
  extern __shared__ __align__(16) float smem[];
  float* A = smem;                       // [K][LSTM_TMP]
  float* Wb = smem + K * LSTM_TMP;       // [2][KS][512]
  const int dir = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * LSTM_TM;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const float* Wd = Wp + (size_t)dir * K * G4;

  float bv[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bv[j] = bp[dir * G4 + tx * 4 + j];
    bv[4 + j] = bp[dir * G4 + 256 + tx * 4 + j];
  }
  float c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.f;
  for (int i = tid; i < H * LSTM_TMP; i += LSTM_THREADS) A[FIN * LSTM_TMP + i] = 0.f;

  auto load_slab = [&](int s, int buf) {
    const float* src = Wd + (size_t)s * LSTM_KS * G4;
    float* dst = Wb + buf * LSTM_KS * G4;
#pragma unroll
    for (int i = 0; i < LSTM_KS * G4 / 4 / LSTM_THREADS; ++i) {
      int ch = tid + i * LSTM_THREADS;
      cp_async16(dst + ch * 4, src + ch * 4);
    }
  };

  for (int step = 0; step < T_STEPS; ++step) {
    const int t = dir ? (T_STEPS - 1 - step) : step;   // bw consumes t=32..0 (model.py:306-312)
    const float* src = in + (size_t)t * FIN * np + p0;
    for (int i = tid; i < FIN * (LSTM_TM / 4); i += LSTM_THREADS) {
      int f = i >> 4, ch = i & 15;
      cp_async16(&A[f * LSTM_TMP + ch * 4], src + (size_t)f * np + ch * 4);
    }
    load_slab(0, 0);
    cp_async_commit();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int s = 0; s < NSLAB; ++s) {
      if (s + 1 < NSLAB) {
        load_slab(s + 1, (s + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* wb = Wb + (s & 1) * LSTM_KS * G4;
      const float* ab = A + (s * LSTM_KS) * LSTM_TMP + ty * 8;
#pragma unroll
      for (int kk = 0; kk < LSTM_KS; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(ab + kk * LSTM_TMP);
        float4 a1 = *reinterpret_cast<const float4*>(ab + kk * LSTM_TMP + 4);
        float4 w0 = *reinterpret_cast<const float4*>(wb + kk * G4 + tx * 4);
        float4 w1 = *reinterpret_cast<const float4*>(wb + kk * G4 + 256 + tx * 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }

    // gates -> (c, h); units tx and 64+tx for rows ty*8..ty*8+7
    float hv[2][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float zi = acc[i][u * 4 + 0] + bv[u * 4 + 0];
        float zg = acc[i][u * 4 + 1] + bv[u * 4 + 1];
        float zf = acc[i][u * 4 + 2] + bv[u * 4 + 2];
        float zo = acc[i][u * 4 + 3] + bv[u * 4 + 3];
        float cn = tanhf(zg) * sigmoid_f(zi) + c[i][u] * sigmoid_f(zf);
        c[i][u] = cn;
        hv[u][i] = tanhf(cn) * sigmoid_f(zo);
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int unit = u * 64 + tx;
      float4 h0 = make_float4(hv[u][0], hv[u][1], hv[u][2], hv[u][3]);
      float4 h1 = make_float4(hv[u][4], hv[u][5], hv[u][6], hv[u][7]);
      float* as = A + (FIN + unit) * LSTM_TMP + ty * 8;
      *reinterpret_cast<float4*>(as) = h0;
      *reinterpret_cast<float4*>(as + 4) = h1;
      float* og = out + ((size_t)t * 2 * H + dir * H + unit) * np + p0 + ty * 8;
      *reinterpret_cast<float4*>(og) = h0;
      *reinterpret_cast<float4*>(og + 4) = h1;
    }
    // the next step's first __syncthreads (after its cp.async wait) orders these A writes
    // before any thread reads them; the x_t rows being overwritten were last read before the
    // final __syncthreads of the slab loop.
  }
}

// ---------------------------------------------------------------------------------------------
// L3 slice-dense (model.py:225-244, 464-471): per site and channel c, 33 -> 30 dense + SELU.
// in h2[33][256][np]; weights w3p[256][33][32] (o padded to 32), b3p[256][32];
// out l3T[o*256+c][np] (the row-major flatten index of model.py:474-478).
// grid = (np/128, 256/8), block 128: thread = one site, loop over 8 channels.
// ---------------------------------------------------------------------------------------------
constexpr int L3_CPB = 8;
__global__ void __launch_bounds__(128)
l3_slice_dense(const float* __restrict__ h2, const float* __restrict__ w3p,
               const float* __restrict__ b3p, float* __restrict__ l3T, int64_t np) {
  __shared__ __align__(16) float ws[L3_CPB][T_STEPS][32];
  __shared__ float bs[L3_CPB][32];
  const int c0 = blockIdx.y * L3_CPB;
  const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const float4* wsrc = reinterpret_cast<const float4*>(w3p + (size_t)c0 * T_STEPS * 32);
  float4* wdst = reinterpret_cast<float4*>(&ws[0][0][0]);
  for (int i = threadIdx.x; i < L3_CPB * T_STEPS * 8; i += 128) wdst[i] = wsrc[i];
  for (int i = threadIdx.x; i < L3_CPB * 32; i += 128) bs[i / 32][i % 32] = b3p[c0 * 32 + i];
  __syncthreads();
  for (int cc = 0; cc < L3_CPB; ++cc) {
    const int c = c0 + cc;
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = bs[cc][o];
#pragma unroll 3
    for (int t = 0; t < T_STEPS; ++t) {
      float v = h2[((size_t)t * 2 * H + c) * np + p];
#pragma unroll
      for (int o4 = 0; o4 < 8; ++o4) {
        float4 w = *reinterpret_cast<const float4*>(&ws[cc][t][o4 * 4]);
        acc[o4 * 4 + 0] = fmaf(v, w.x, acc[o4 * 4 + 0]);
        acc[o4 * 4 + 1] = fmaf(v, w.y, acc[o4 * 4 + 1]);
        acc[o4 * 4 + 2] = fmaf(v, w.z, acc[o4 * 4 + 2]);
        acc[o4 * 4 + 3] = fmaf(v, w.w, acc[o4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int o = 0; o < L3_UNITS; ++o) l3T[((size_t)o * 2 * H + c) * np + p] = selu_f(acc[o]);
  }
}

// ---------------------------------------------------------------------------------------------
// L4 dense 7680 -> 192 + SELU (model.py:482-488).  A^T = l3T[7680][np], W4[7680][192],
// out l4T[192][np].  grid = np/64, block 256: 8 rows x 6 cols per thread, K slabs of 16.
// ---------------------------------------------------------------------------------------------
constexpr int L4_TM = 64, L4_TMP = 68, L4_KS = 16;
constexpr size_t l4_smem_bytes() { return (size_t)2 * L4_KS * (L4_TMP + L4_UNITS) * sizeof(float); }

__global__ void __launch_bounds__(256)
l4_dense(const float* __restrict__ l3T, const float* __restrict__ W4, const float* __restrict__ b4,
         float* __restrict__ l4T, int64_t np) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                            // [2][KS][L4_TMP]
  float* Ws = smem + 2 * L4_KS * L4_TMP;       // [2][KS][192]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * L4_TM;
  auto load = [&](int s, int buf) {
    {  // A slab: 16 rows x 64 floats = 256 chunks
      int f = tid >> 4, ch = tid & 15;
      cp_async16(As + (buf * L4_KS + f) * L4_TMP + ch * 4,
                 l3T + ((size_t)s * L4_KS + f) * np + p0 + ch * 4);
    }
    const float* wsrc = W4 + (size_t)s * L4_KS * L4_UNITS;
    float* wdst = Ws + buf * L4_KS * L4_UNITS;
#pragma unroll
    for (int i = 0; i < 3; ++i) cp_async16(wdst + (tid + i * 256) * 4, wsrc + (tid + i * 256) * 4);
  };
  float acc[8][6];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
  constexpr int NS = L3_K / L4_KS;
  load(0, 0);
  cp_async_commit();
  for (int s = 0; s < NS; ++s) {
    if (s + 1 < NS) {
      load(s + 1, (s + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* ab = As + (s & 1) * L4_KS * L4_TMP + ty * 8;
    const float* wb = Ws + (s & 1) * L4_KS * L4_UNITS;
#pragma unroll
    for (int kk = 0; kk < L4_KS; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(ab + kk * L4_TMP);
      float4 a1 = *reinterpret_cast<const float4*>(ab + kk * L4_TMP + 4);
      float2 w0 = *reinterpret_cast<const float2*>(wb + kk * L4_UNITS + tx * 2);
      float2 w1 = *reinterpret_cast<const float2*>(wb + kk * L4_UNITS + 64 + tx * 2);
      float2 w2 = *reinterpret_cast<const float2*>(wb + kk * L4_UNITS + 128 + tx * 2);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int col = (j >> 1) * 64 + tx * 2 + (j & 1);
    const float b = b4[col];
    float4 v0 = make_float4(selu_f(acc[0][j] + b), selu_f(acc[1][j] + b), selu_f(acc[2][j] + b), selu_f(acc[3][j] + b));
    float4 v1 = make_float4(selu_f(acc[4][j] + b), selu_f(acc[5][j] + b), selu_f(acc[6][j] + b), selu_f(acc[7][j] + b));
    float* o = l4T + (size_t)col * np + p0 + ty * 8;
    *reinterpret_cast<float4*>(o) = v0;
    *reinterpret_cast<float4*>(o + 4) = v1;
  }
}

// ---------------------------------------------------------------------------------------------
// Tail: L5_1..4 (192->96, SELU) -> 4 heads (96->21/3/33/33, SELU) -> softmax  (model.py:507-622).
// in l4T[192][np]; W5[192][384] (four L5 kernels side by side), b5[384];
// Whd[96][90] (four head kernels side by side: column off_k+o reads L5 branch k), bhd[90].
// out probs[n][90] (real rows), logits[np][90] (padded rows; post-SELU, pre-softmax).
// grid = np/64, block 256.
// ---------------------------------------------------------------------------------------------
constexpr int TL_TM = 64, TL_TMP = 68, TL_KS = 16;
constexpr size_t tail_smem_bytes() {
  return (size_t)(L4_UNITS * TL_TMP + 2 * TL_KS * L5_ALL + L5_ALL * TL_TMP) * sizeof(float);
}

__global__ void __launch_bounds__(256)
tail_heads(const float* __restrict__ l4T, const float* __restrict__ W5, const float* __restrict__ b5,
           const float* __restrict__ Whd, const float* __restrict__ bhd,
           float* __restrict__ probs, float* __restrict__ logits, SiteMap sm) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                                  // [192][TL_TMP]   (later: logits [64][92])
  float* Ws = As + L4_UNITS * TL_TMP;                // [2][KS][384]    (later: Whd [96][90] + bhd)
  float* L5s = Ws + 2 * TL_KS * L5_ALL;              // [384][TL_TMP]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t np = sm.np;
  const int64_t p0 = (int64_t)blockIdx.x * TL_TM;

  for (int i = tid; i < L4_UNITS * 16; i += 256) {
    int f = i >> 4, ch = i & 15;
    cp_async16(As + f * TL_TMP + ch * 4, l4T + (size_t)f * np + p0 + ch * 4);
  }
  auto load = [&](int s, int buf) {
    const float* wsrc = W5 + (size_t)s * TL_KS * L5_ALL;
    float* wdst = Ws + buf * TL_KS * L5_ALL;
#pragma unroll
    for (int i = 0; i < 6; ++i) cp_async16(wdst + (tid + i * 256) * 4, wsrc + (tid + i * 256) * 4);
  };
  load(0, 0);
  cp_async_commit();
  float acc[8][12];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[i][j] = 0.f;
  constexpr int NS = L4_UNITS / TL_KS;   // 12
  for (int s = 0; s < NS; ++s) {
    if (s + 1 < NS) {
      load(s + 1, (s + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* ab = As + (s * TL_KS) * TL_TMP + ty * 8;
    const float* wb = Ws + (s & 1) * TL_KS * L5_ALL;
#pragma unroll
    for (int kk = 0; kk < TL_KS; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(ab + kk * TL_TMP);
      float4 a1 = *reinterpret_cast<const float4*>(ab + kk * TL_TMP + 4);
      float4 w0 = *reinterpret_cast<const float4*>(wb + kk * L5_ALL + tx * 4);
      float4 w1 = *reinterpret_cast<const float4*>(wb + kk * L5_ALL + 128 + tx * 4);
      float4 w2 = *reinterpret_cast<const float4*>(wb + kk * L5_ALL + 256 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const int col = (j >> 2) * 128 + tx * 4 + (j & 3);
    const float b = b5[col];
    float* o = L5s + col * TL_TMP + ty * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = selu_f(acc[i][j] + b);
  }
  // head weights into the (now idle) slab ring
  float* Whs = Ws;                 // [96][90]
  float* bhs = Ws + L5_UNITS * N_OUT;
  for (int i = tid; i < L5_UNITS * N_OUT; i += 256) Whs[i] = Whd[i];
  if (tid < N_OUT) bhs[tid] = bhd[tid];
  __syncthreads();

  float* Ls = As;                  // [64][92]
  {
    const int site = tid & 63, og = tid >> 6;
    for (int o = og; o < N_OUT; o += 4) {
      const int k = o < 21 ? 0 : (o < 24 ? 1 : (o < 57 ? 2 : 3));
      const float* l5 = L5s + (k * L5_UNITS) * TL_TMP + site;
      float s = bhs[o];
#pragma unroll 8
      for (int j = 0; j < L5_UNITS; ++j) s = fmaf(l5[j * TL_TMP], Whs[j * N_OUT + o], s);
      Ls[site * 92 + o] = selu_f(s);
    }
  }
  __syncthreads();
  {
    const int site = tid & 63, k = tid >> 6;
    const int off = kHeadOff[k], cnt = kHeadOff[k + 1] - off;
    const float* z = Ls + site * 92 + off;
    const int64_t p = p0 + site;
    const int64_t r = sm.real_row(p);
    float m = z[0];
    for (int o = 1; o < cnt; ++o) m = fmaxf(m, z[o]);
    float sum = 0.f;
    for (int o = 0; o < cnt; ++o) sum += expf(z[o] - m);
    const float inv = 1.f / sum;
    for (int o = 0; o < cnt; ++o) {
      logits[p * N_OUT + off + o] = z[o];
      if (r >= 0) probs[r * N_OUT + off + o] = expf(z[o] - m) * inv;
    }
  }
}

}  // namespace simt
}  // namespace clairb
