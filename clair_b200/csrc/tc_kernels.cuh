// Tensor-core (tcgen05 / TMEM / bulk-async-copy) kernels of the forward path for sm_100a.
//
// Numerics: every contraction runs on fp16 operands split hi/lo (x = hi + lo, both fp16) with the three products
// hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM (kind::f16).  That keeps ~22 mantissa bits per operand, which the
// <=1e-4 logit gate needs (single fp16/bf16/tf32 operands fail it: SURVEY.md section 7).
//
// A bidirectional LSTM layer (reference clair/model.py:265-312) is
//   layer 1: lstm_seq<FUSE_X>     gates = x_t . W_x + b + h_{t-1} . W_h, all inside the recurrent kernel (K_x = 32)
//   layer 2: xproj_pair           Gx[t,site,dir,512] = x_t . W_x + b   one large GEMM over all 33 steps (K_x = 256)
//            lstm_seq             gates = Gx[t] + h_{t-1} . W_h ; (c,h) update, 33 dependent steps
// Both run on CTA pairs (cta_group::2): a pair owns 256 sites, each CTA holds its 128 rows of the A operand and one
// half of the gate columns of the B operand, so the 256 KB fp16 hi/lo recurrent kernel of one direction is split
// 128 KB + 128 KB over the two SMs, stays resident in shared memory and is never re-read from L2.
//
// Operand tiles live in global memory already in the UMMA canonical K-major no-swizzle order
// [k/8][row][8] (see tc_common.cuh), so they move with plain 1-D bulk copies in both directions.
//
// Gate column order inside a direction: j = unit*4 + gate, gate in TF LSTMBlockCell order (i, c, f, o), so the four
// pre-activations of one hidden unit are one float4 of Gx and four adjacent TMEM columns.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace clairb {
namespace tc {

// ---- sizes --------------------------------------------------------------------------------------
constexpr int KCH = 1024;                      // halves per k-chunk of a 128-row tile: 128 rows x 8
constexpr int KCH_BYTES = 2048;
constexpr int GX_TILE_FLOATS = 2 * 128 * 128 * 4;   // per (t, tile): [dir][unit 128][row 128][gate 4]
constexpr int L3_TG = 5;                                  // time groups of 8 kept per channel (t = 0..39, 33..39 zero)
constexpr int L3A_HALVES = 2 * L3_TG * 1024;              // per (tile, channel): [hl][t/8][site/64][t%8][64 sites, 128B-swizzled]
constexpr int L3A_BYTES = L3A_HALVES * 2;                 // 20480
// per-channel weight blob of l3l4_fused: W3_c ([6 kc][hi 32 o | lo 32 o][8 t]) , W4_c hi|lo ([4 kc][192 n][8 o]) , b3_c[32]
constexpr int L3W_BYTES = 6 * 32 * 16;                    // 3072
constexpr int L4W_BYTES = 4 * 192 * 16;                   // 12288
constexpr int L3L4_BLOB_BYTES = 31 * 1024;                            // 30848 used, padded so ring stages stay 1 KB aligned

constexpr float LOG2E = 1.4426950408889634f;
// Gate pre-activations reach the epilogue pre-scaled (the scale is folded into W and b on the host):
//   pi = -log2e*zi, pg = -2*log2e*zg, pf = -log2e*zf, po = -log2e*zo      so that 2^p = e^-z (resp. e^-2z)
constexpr float GATE_SCALE[4] = {-LOG2E, -2.f * LOG2E, -LOG2E, -LOG2E};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// TF LSTMBlockCell (forget_bias 0, no clip, no peephole; model.py:299-305):
//   c' = tanh(zg)*sigmoid(zi) + c*sigmoid(zf),  h = tanh(c')*sigmoid(zo)
// written over common denominators: with a=e^-zi, b=e^-2zg, d=e^-zf:
//   c' = [(1-b)(1+d) + c(1+a)(1+b)] / [(1+a)(1+b)(1+d)]          3 ex2 + 1 rcp
//   h  = (1-e) / [(1+e)(1+f)],  e=e^-2c', f=e^-zo                 2 ex2 + 1 rcp
// Exponent arguments are clamped at 2^40 so the denominators stay finite (sigmoid(-27.7) ~ 1e-12).
__device__ __forceinline__ float lstm_cell(float pi, float pg, float pf, float po, float& c) {
  // The epilogue that calls this sits on its issue-slot bound, so the products are folded into FFMAs:
  //   (1+a)(1+b) = B + a*B,   (1+a)(1+b)(1+d) = AB + d*AB,   (1-b)(1+d) = nb + d*nb,   (1-e)*r = r - e*r
  const float a = ex2f(fminf(pi, 40.f)), b = ex2f(fminf(pg, 40.f)), d = ex2f(fminf(pf, 40.f));
  const float B = 1.f + b, nb = 1.f - b;
  const float AB = fmaf(a, B, B);
  const float cn = fmaf(c, AB, fmaf(d, nb, nb)) * rcpf(fmaf(AB, d, AB));
  c = cn;
  const float e = ex2f(fminf(cn * (-2.f * LOG2E), 40.f)), f = ex2f(fminf(po, 40.f));
  const float E = 1.f + e;
  const float r = rcpf(fmaf(E, f, E));
  return fmaf(-e, r, r);
}
// 8x8 transpose of fp16 values across the 8 lanes of a row group: on entry lane i (= lane & 7) holds 8 consecutive hidden
// units of row i as 4 words (two units per word); on exit it holds unit i for the 8 rows (two rows per word).
__device__ __forceinline__ void transpose8x8_h(uint32_t* w, int lane) {
  // 1) 2x2 blocks of halves between lanes i, i^1
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t p = __shfl_xor_sync(0xffffffffu, w[k], 1);
    w[k] = (lane & 1) ? __byte_perm(w[k], p, 0x3276) : __byte_perm(w[k], p, 0x5410);
  }
  // 2) 4x4 transpose of words between lanes (bits 1,2) and registers
#pragma unroll
  for (int k1 = 0; k1 < 2; ++k1) {
    const bool up = lane & 2;
    const uint32_t send = up ? w[2 * k1] : w[2 * k1 + 1];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
    if (up) w[2 * k1] = recv; else w[2 * k1 + 1] = recv;
  }
#pragma unroll
  for (int k0 = 0; k0 < 2; ++k0) {
    const bool up = lane & 4;
    const uint32_t send = up ? w[k0] : w[2 + k0];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 4);
    if (up) w[k0] = recv; else w[2 + k0] = recv;
  }
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 lds4(uint32_t smem_addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_addr));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// split 8 floats into fp16 hi / lo chunks (16 bytes each)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    float2 back = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                  *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                  *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}

// the three stages of transpose8x8_h as separate calls, so that the epilogue can spread them between its cell updates
__device__ __forceinline__ void t8_stage1(uint32_t* w, int lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t p = __shfl_xor_sync(0xffffffffu, w[k], 1);
    w[k] = (lane & 1) ? __byte_perm(w[k], p, 0x3276) : __byte_perm(w[k], p, 0x5410);
  }
}
__device__ __forceinline__ void t8_stage2(uint32_t* w, int lane) {
#pragma unroll
  for (int k1 = 0; k1 < 2; ++k1) {
    const bool up = lane & 2;
    const uint32_t send = up ? w[2 * k1] : w[2 * k1 + 1];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
    if (up) w[2 * k1] = recv; else w[2 * k1 + 1] = recv;
  }
}
__device__ __forceinline__ void t8_stage3(uint32_t* w, int lane) {
#pragma unroll
  for (int k0 = 0; k0 < 2; ++k0) {
    const bool up = lane & 4;
    const uint32_t send = up ? w[k0] : w[2 + k0];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 4);
    if (up) w[k0] = recv; else w[2 + k0] = recv;
  }
}

// ---------------------------------------------------------------------------------------------
// Epilogue of the recurrent kernels (lstm_seq<FUSE_X> of layer 1, lstm_seq_x2 of layer 2): all 33 steps of one epilogue
// warp.  G warps per TMEM lane quarter; warp `sub` of a quarter owns units sub*UW..+UW (UW = 32/G) of every gate block and
// walks them in slices of 8 units.  Per block: wait for the accumulator, pull the warp's whole share into registers, hand
// the accumulator back, then per slice the cell update (7 MUFU per cell: the pipe that bounds this code) and its TAIL:
//   A: h_t -> fp16 hi/lo words -> tcgen05.st into the h buffer (next step's A operand) [-> publish the block on hq[b]]
//   B: the same words -> the next layer's operand tile in global memory (OUT 0: K-major; OUT 2: MN-major, 8x8 transposes)
// The tail of a slice is not run behind its cells (all warps of a scheduler then convert / shuffle / store in lockstep while
// the MUFU pipe idles: measured 2 x 300 of 3.2 k cycles per block), it is issued piecewise BETWEEN the cell updates of the
// NEXT slice.  Only the last slice of a step cannot wait - the next step's first gate block contracts over it - so its part
// A runs at once and its part B rides with the first slice of the next step.
// ---------------------------------------------------------------------------------------------
struct EpiArgs {
  uint32_t tmem;
  uint64_t* acc_full;                    // [2] this CTA's "gate block complete" barriers
  uint32_t leader_hq, leader_acc_empty0, leader_acc_empty1;    // cluster addresses in the leader CTA
  uint32_t bias_base;                    // shared-memory address of [128 units][4 gates] floats (BIAS)
  void* Hout;
  int NT;
  int64_t np;
  int dir, tile;
  long long* trace;                      // CLAIRB_TRACE builds: [33 steps][64] clock64 stamps of epilogue warp 1 (events 16 + 8b + k)
};

// PIPE = false keeps the tail of a slice right behind its own cells (layer 2: measured 5 % faster that way - the chip runs
// this kernel at its 1000 W cap, and the deferred publish / extra live registers cost more than the overlap returns).
template <int OUT, int G, bool BIAS, bool PIPE>
__device__ __forceinline__ void lstm_epilogue(const EpiArgs& a, int warp, int lane) {
#ifdef CLAIRB_TRACE
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 32;
  auto stamp = [&](int s, int ev) { if (tr_on) a.trace[s * 64 + ev] = clock64(); };
#else
  auto stamp = [&](int, int) {};
#endif
  static_assert(OUT == 0 || OUT == 1 || OUT == 2, "unknown output layout");
  constexpr int UW = 32 / G, NSL = UW / 8;
  const int ew = warp - 1;
  const int quarter = warp & 3;                        // TMEM lane quarter this warp may touch
  const int sub = ew >> 2;                             // 0..G-1
  const int r = quarter * 32 + lane;
  const uint32_t lane_base = a.tmem + ((uint32_t)(quarter * 32) << 16);
  float c[4][UW];
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int i = 0; i < UW; ++i) c[b][i] = 0.f;
  uint32_t use0 = 0, use1 = 0;
  uint32_t dhi[4] = {0, 0, 0, 0}, dlo[4] = {0, 0, 0, 0};      // part B owed for the last slice of the previous step
  int t_prev = 0;

  auto out_ptr = [&](int t, int u0) -> uint8_t* {      // where the 16-byte hi chunk of this lane goes (lo: see lo_off)
    if (OUT == 0) {
      return (uint8_t*)((__half*)a.Hout + ((size_t)t * a.NT + a.tile) * (2 * 32 * KCH) + (size_t)(a.dir * 16 + (u0 >> 3)) * KCH + r * 8);
    } else {
      // OUT == 2: MN-major SWIZZLE_128B tiles for the slice-dense MMA (K = time, M = site), per channel
      //   H2t[tile][c][hl][t/8 (5)][site/64 (2)][t%8][16-byte chunk ((site%64)/8) ^ (t%8)][site%8]
      // after the 8x8 transposes every lane holds one channel x 8 sites = one 16-byte chunk
      const int ch = a.dir * H + u0 + (lane & 7);
      const int rg = r >> 3;                                    // site group of 8 within the tile (0..15)
      return (uint8_t*)a.Hout + ((size_t)a.tile * 2 * H + ch) * L3A_BYTES + (size_t)(t >> 3) * 2048 + (rg >> 3) * 1024 +
             (t & 7) * 128 + (((rg & 7) ^ (t & 7)) << 4);
    }
  };
  constexpr size_t lo_off = OUT == 0 ? (size_t)32 * KCH * 2 : (size_t)L3A_BYTES / 2;
  auto publish = [&](int b) {                          // this warp's units of block b of h_t are in tensor memory
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster_relaxed(a.leader_hq + b * 8);
  };
  auto pack_pair = [&](const float* hv, uint32_t* whi, uint32_t* wlo, int k) {
    const __half2 h2 = __floats2half2_rn(hv[2 * k], hv[2 * k + 1]);
    const float2 back = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(hv[2 * k] - back.x, hv[2 * k + 1] - back.y);
    whi[k] = *reinterpret_cast<const uint32_t*>(&h2);
    wlo[k] = *reinterpret_cast<const uint32_t*>(&l2);
  };

  for (int s = 0; s < T_STEPS; ++s) {
    const int t = a.dir ? (T_STEPS - 1 - s) : s;       // bw consumes t = 32..0 (model.py:306-312)
    const uint32_t h_st = lane_base + 256 + (s & 1) * 128;
    float hv_p[8];                                     // h_t of the previous slice, tail still to do
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = b & 1;
      {
        uint32_t& use = i ? use1 : use0;
        mbar_wait(&a.acc_full[i], use & 1);
        ++use;
        tc_fence_after();
      }
      stamp(s, 16 + b * 8);
      // this warp's whole share of the accumulator -> registers, then hand the accumulator back at once
      float v[4 * UW];
#pragma unroll
      for (int k = 0; k < UW / 4; ++k) tmem_ld16(lane_base + i * 128 + sub * UW * 4 + k * 16, v + 16 * k);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(i ? a.leader_acc_empty1 : a.leader_acc_empty0);
      stamp(s, 16 + b * 8 + 1);
#pragma unroll
      for (int sl = 0; sl < NSL; ++sl) {
        const bool first = b == 0 && sl == 0;          // compile-time after unrolling
        const int u0 = b * 32 + sub * UW + sl * 8;     // first hidden unit of this slice
        const int pb = sl ? b : b - 1, psl = sl ? sl - 1 : NSL - 1;     // the slice whose tail rides along
        const int pu0 = pb * 32 + sub * UW + psl * 8;
        float hv[8];
        uint32_t whi[4], wlo[4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float* vv = v + (sl * 8 + k) * 4;
#ifdef CLAIRB_EPI_NOCELL                                // timing probe: no transcendental work
          hv[k] = c[b][sl * 8 + k] = fmaf(vv[0], vv[1], fmaf(vv[2], vv[3], c[b][sl * 8 + k]));
#else
          if (BIAS) {
            const float4 bq = lds4(a.bias_base + (u0 + k) * 16);
            hv[k] = lstm_cell(vv[0] + bq.x, vv[1] + bq.y, vv[2] + bq.z, vv[3] + bq.w, c[b][sl * 8 + k]);
          } else {
            hv[k] = lstm_cell(vv[0], vv[1], vv[2], vv[3], c[b][sl * 8 + k]);
          }
#endif
          if (OUT == 1 || !PIPE) continue;             // plain tail below
#ifdef CLAIRB_EPI_NOTAIL                                // timing probe: no tail (h is not carried: results are wrong)
          if (!first && k == 3 && psl == NSL - 1) publish(pb);
          continue;
#endif
          if (!first) {
            if (k < 4) pack_pair(hv_p, whi, wlo, k);
            if (k == 3) {
              tmem_st4(h_st + (pu0 >> 1), whi);
              tmem_st4(h_st + 64 + (pu0 >> 1), wlo);
              if (psl == NSL - 1) publish(pb);
            }
            if (OUT == 2) {
              if (k == 4) { t8_stage1(whi, lane); t8_stage1(wlo, lane); }
              if (k == 5) { t8_stage2(whi, lane); t8_stage2(wlo, lane); }
              if (k == 6) { t8_stage3(whi, lane); t8_stage3(wlo, lane); }
            }
            if (k == 7) {
              uint8_t* out = out_ptr(t, pu0);
              *reinterpret_cast<uint4*>(out) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
              *reinterpret_cast<uint4*>(out + lo_off) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
            }
          } else {
            // part B of the previous step's last slice (nothing is owed at s = 0)
            if (OUT == 2) {
              if (k == 2) { t8_stage1(dhi, lane); t8_stage1(dlo, lane); }
              if (k == 3) { t8_stage2(dhi, lane); t8_stage2(dlo, lane); }
              if (k == 4) { t8_stage3(dhi, lane); t8_stage3(dlo, lane); }
            }
            if (k == 5 && s > 0) {
              uint8_t* out = out_ptr(t_prev, 3 * 32 + sub * UW + (NSL - 1) * 8);
              *reinterpret_cast<uint4*>(out) = make_uint4(dhi[0], dhi[1], dhi[2], dhi[3]);
              *reinterpret_cast<uint4*>(out + lo_off) = make_uint4(dlo[0], dlo[1], dlo[2], dlo[3]);
            }
          }
        }
        if (OUT == 1) {
          uint32_t w1[4], w2[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) pack_pair(hv, w1, w2, k);
          tmem_st4(h_st + (u0 >> 1), w1);
          tmem_st4(h_st + 64 + (u0 >> 1), w2);
          float* out = (float*)a.Hout + ((size_t)t * 2 * H + a.dir * H + u0) * a.np + (size_t)a.tile * 128 + r;
#pragma unroll
          for (int k = 0; k < 8; ++k) out[(size_t)k * a.np] = hv[k];
          if (sl == NSL - 1) publish(b);
        } else if (!PIPE) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pack_pair(hv, whi, wlo, k);
          tmem_st4(h_st + (u0 >> 1), whi);
          tmem_st4(h_st + 64 + (u0 >> 1), wlo);
          if (OUT == 2) {
            transpose8x8_h(whi, lane);
            transpose8x8_h(wlo, lane);
          }
          uint8_t* out = out_ptr(t, u0);
          *reinterpret_cast<uint4*>(out) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
          *reinterpret_cast<uint4*>(out + lo_off) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
          if (sl == NSL - 1) publish(b);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) hv_p[k] = hv[k];
        }
        stamp(s, 16 + b * 8 + 2 + sl);
      }
    }
    if (OUT != 1 && PIPE) {
      // last slice of the step: part A now, part B with the next step's first slice
#pragma unroll
      for (int k = 0; k < 4; ++k) pack_pair(hv_p, dhi, dlo, k);
      const int lu0 = 3 * 32 + sub * UW + (NSL - 1) * 8;
      tmem_st4(h_st + (lu0 >> 1), dhi);
      tmem_st4(h_st + 64 + (lu0 >> 1), dlo);
      publish(3);
      t_prev = t;
    }
    stamp(s, 16 + 3 * 8 + 7);
  }
  if (OUT != 1 && PIPE) {
    if (OUT == 2) {
      transpose8x8_h(dhi, lane);
      transpose8x8_h(dlo, lane);
    }
    uint8_t* out = out_ptr(t_prev, 3 * 32 + sub * UW + (NSL - 1) * 8);
    *reinterpret_cast<uint4*>(out) = make_uint4(dhi[0], dhi[1], dhi[2], dhi[3]);
    *reinterpret_cast<uint4*>(out + lo_off) = make_uint4(dlo[0], dlo[1], dlo[2], dlo[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// xproj_pair<KC>: Gx[a][dir][unit][row][gate] = A[a] . Wx[:, dir, unit*4+gate] + b      (a = t*NT + tile)
//   A tiles  : [a][hl][KC][128][8] fp16          (KC = K/8: 4 for layer 1, 32 for layer 2)
//   Wx       : [nb 4][q 2][hl][KC][128][8] fp16  N-block nb = dir*2 + half covers gate columns half*256..+256 of `dir`;
//              CTA q of the pair supplies rows q*128..+128 of the block
//   bias     : [nb 4][256] fp32 in the same column order
// Persistent CTA pairs: a pair keeps ONE N-block of Wx resident in shared memory (B operand, <=128 KB per CTA) and
// streams row-pairs (256 rows = two A tiles) through a ring of K=64 stages.  Pairs 4g..4g+3 walk the same row-pairs
// with the four different N-blocks at the same time, so each A tile is read from HBM once and hit in L2 three times.
// Warp roles: 0 = A producer, 1 = MMA issuer (leader) / stage relay (peer), 2..9 = epilogue (TMEM -> +bias -> Gx).
// ---------------------------------------------------------------------------------------------
constexpr int XP_SKC = 8;                              // k-chunks of 8 per ring stage: K = 64 per stage
constexpr int XP_RING = 3;
constexpr int XP_STAGE_BYTES = 2 * XP_SKC * KCH_BYTES; // [hl][8 kc][128][8] = 32 KB
constexpr int XP_THREADS = 352;                        // warps: producer, MMA / relay, 8 epilogue, second relay

template <int KC>
constexpr size_t xproj_smem_bytes() {
  return (size_t)2 * KC * KCH_BYTES + (size_t)XP_RING * XP_STAGE_BYTES + 256 * 4 + 256 + 1024;
}

template <int KC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XP_THREADS, 1)
xproj_pair(const __half* __restrict__ A, const __half* __restrict__ Wx, const float* __restrict__ bias,
           float* __restrict__ Gx, int num_row_pairs, int dbg) {
  constexpr int NST = KC / XP_SKC;                     // ring stages per row tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Bs = smem;                                  // [hl][KC][128][8]
  uint8_t* ring = Bs + 2 * KC * KCH_BYTES;
  float* bias_s = (float*)(ring + XP_RING * XP_STAGE_BYTES);
  uint64_t* bars = (uint64_t*)(bias_s + 256);
  uint64_t* full = bars;                               // [XP_RING] this CTA's stage landed (tx bytes)
  uint64_t* peer_full = bars + XP_RING;                // [XP_RING] (leader) peer's stage landed
  uint64_t* empty = bars + 2 * XP_RING;                // [XP_RING] stage consumed (MMA commit, both CTAs)
  uint64_t* acc_full = bars + 3 * XP_RING;             // [2] accumulator buffer complete (MMA commit, both CTAs)
  uint64_t* acc_empty = acc_full + 2;                  // [2] (leader) drained by all 8 epilogue warps of the pair
  uint64_t* b_full = acc_empty + 2;                    // [1] resident B landed
  uint32_t* tmem_slot = (uint32_t*)(b_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nb = cid & 3, grp = cid >> 2, ngrp = ncl >> 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < XP_RING; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&peer_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 16);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
    // resident B: this CTA's 128 rows of N-block nb
    const __half* src = Wx + ((size_t)nb * 2 + rank) * (2 * KC * KCH);
    mbar_expect_tx(b_full, 2 * KC * KCH_BYTES);
    for (int i = 0; i < 2 * KC * KCH_BYTES; i += 16384) bulk_g2s(Bs + i, (const uint8_t*)src + i, 16384, b_full);
  }
  if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
  for (int i = threadIdx.x; i < 256; i += XP_THREADS) bias_s[i] = bias[nb * 256 + i];
  __syncthreads();                                     // barrier inits visible before anyone polls them
  mbar_wait(b_full, 0);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // relay (peer CTA): forwards "stage landed" for every second stage (two relay threads alternate, so each release-arrive
  // -- ~700 cycles -- has two stage times to complete)
  auto relay_stages = [&](int which) {
    uint32_t use = 0;
    const uint32_t leader_peer_full = map_to_cta(smem_u32(peer_full), 0);
    for (int rp = grp; rp < num_row_pairs; rp += ngrp)
      for (int ks = 0; ks < NST; ++ks, ++use) {
        if ((int)(use & 1) != which) continue;
        const int slot = use % XP_RING;
        mbar_wait(&full[slot], (use / XP_RING) & 1);
        mbar_arrive_cluster(leader_peer_full + slot * 8);
      }
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---- A producer (each CTA loads its own 128 rows) ----
      uint32_t use = 0;
      for (int rp = grp; rp < num_row_pairs; rp += ngrp) {
        const __half* at = A + ((size_t)((dbg & 64) ? (rp & 7) : rp) * 2 + rank) * (2 * KC * KCH);
        {
          // pull the A tile two row-pairs ahead from HBM into L2 so the ring's bulk copies see L2 latency only
          const int rpn = rp + 2 * ngrp;
          if (rpn < num_row_pairs) {
            const uint8_t* an = (const uint8_t*)(A + ((size_t)rpn * 2 + rank) * (2 * KC * KCH));
            for (int i = 0; i < 2 * KC * KCH_BYTES; i += 32768)
              bulk_prefetch_l2(an + i, (2 * KC * KCH_BYTES - i) < 32768 ? (2 * KC * KCH_BYTES - i) : 32768);
          }
        }
        for (int ks = 0; ks < NST; ++ks, ++use) {
          const int slot = use % XP_RING;
          mbar_wait(&empty[slot], ((use / XP_RING) & 1) ^ 1);
          uint8_t* dst = ring + slot * XP_STAGE_BYTES;
          if (dbg & 8) { mbar_arrive(&full[slot]); continue; }
          mbar_expect_tx(&full[slot], XP_STAGE_BYTES);
          bulk_g2s(dst, at + (size_t)ks * XP_SKC * KCH, XP_SKC * KCH_BYTES, &full[slot]);
          bulk_g2s(dst + XP_SKC * KCH_BYTES, at + (size_t)KC * KCH + (size_t)ks * XP_SKC * KCH, XP_SKC * KCH_BYTES, &full[slot]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank == 1) {
        // ---- relay: tell the leader when this CTA's stage has landed.  The arrive MUST be a release at cluster scope:
        //      with a relaxed arrive the pair-MMA occasionally read this CTA's stage before the bulk copy's data was
        //      visible to it (rare, size-dependent logit errors ~3e-3; tools/parity_loop.py).  The release costs ~700
        //      cycles (ERRBAR + cluster fence), so stages are K = 64 (4 relays per unit, not 8) to keep this thread off
        //      the critical path. ----
        relay_stages(0);
      } else {
        // ---- MMA issuer (one thread feeds the tensor pipe: keep its instruction stream short -- descriptors are
        //      advanced with one add, ring slot / parity are carried instead of divided) ----
        const uint32_t idesc = make_idesc_f16(256, 256);
        const uint64_t b_hi0 = make_smem_desc(smem_u32(Bs), KCH_BYTES, 128);
        const uint64_t b_lo0 = desc_advance(b_hi0, KC * KCH_BYTES);
        const uint64_t a_ring = make_smem_desc(smem_u32(ring), KCH_BYTES, 128);
        uint32_t slot = 0, par = 0, it = 0;
        for (int rp = grp; rp < num_row_pairs; rp += ngrp, ++it) {
          const uint32_t buf = it & 1;
          mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem + buf * 256;
#pragma unroll
          for (int ks = 0; ks < NST; ++ks) {
            mbar_wait(&full[slot], par);
            mbar_wait(&peer_full[slot], par);
            tc_fence_after();
            const uint64_t a_hi0 = desc_advance(a_ring, slot * XP_STAGE_BYTES), a_lo0 = desc_advance(a_hi0, XP_SKC * KCH_BYTES);
#pragma unroll
            for (int kk = 0; kk < XP_SKC / 2; ++kk) {
              const uint64_t a_hi = desc_advance(a_hi0, kk * 2 * KCH_BYTES), a_lo = desc_advance(a_lo0, kk * 2 * KCH_BYTES);
              const uint64_t b_hi = desc_advance(b_hi0, (ks * XP_SKC + kk * 2) * KCH_BYTES);
              const uint64_t b_lo = desc_advance(b_lo0, (ks * XP_SKC + kk * 2) * KCH_BYTES);
              umma_f16_pair(d, a_hi, b_hi, idesc, (ks | kk) != 0);
              if (!(dbg & 2)) {
                umma_f16_pair(d, a_lo, b_hi, idesc, 1);
                umma_f16_pair(d, a_hi, b_lo, idesc, 1);
              }
            }
            umma_commit_pair(&empty[slot], 0b11);
            if (++slot == XP_RING) { slot = 0; par ^= 1; }
          }
          umma_commit_pair(&acc_full[buf], 0b11);
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    if (lane == 0 && rank == 1) relay_stages(1);
    __syncwarp();
  } else {
    // ---- epilogue: TMEM -> + bias -> Gx (float4 per hidden unit, 512 B per warp store) ----
    const int quarter = warp & 3;
    const int colhalf = (warp - 2) >> 2;               // two warps per lane quarter split the 256 columns
    const int r = quarter * 32 + lane;
    const int dir = nb >> 1, unit0 = (nb & 1) * 64 + colhalf * 32;
    const uint32_t leader_acc_empty[2] = {map_to_cta(smem_u32(&acc_empty[0]), 0), map_to_cta(smem_u32(&acc_empty[1]), 0)};
    uint32_t it = 0;
    for (int rp = grp; rp < num_row_pairs; rp += ngrp, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      float* out = Gx + ((size_t)((dbg & 32) ? (rp & 7) : rp) * 2 + rank) * GX_TILE_FLOATS + (size_t)dir * (GX_TILE_FLOATS / 2) +
                   (size_t)unit0 * 512 + r * 4;
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256 + colhalf * 128;
      const float* bs = bias_s + colhalf * 128;
#pragma unroll 2
      for (int c = 0; c < 128; c += 32) {
        if ((dbg & 128) && c) __nanosleep((dbg >> 8) * 100);      // pacing experiment
        float v[32];
        tmem_ld16(taddr + c, v);
        tmem_ld16(taddr + c + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 o = make_float4(v[4 * u] + bs[c + 4 * u], v[4 * u + 1] + bs[c + 4 * u + 1],
                                 v[4 * u + 2] + bs[c + 4 * u + 2], v[4 * u + 3] + bs[c + 4 * u + 3]);
          if (!(dbg & 1)) {
            if (dbg & 16) st_stream4(out + (size_t)(c / 4 + u) * 512, o);
            else *reinterpret_cast<float4*>(out + (size_t)(c / 4 + u) * 512) = o;
          }
        }
      }
      // the accumulator is drained (tcgen05.ld complete): no need to wait for the Gx stores before handing it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(leader_acc_empty[buf]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}


// ---------------------------------------------------------------------------------------------
// lstm_seq<FUSE_X, OUT, G>: the 33 dependent steps of one direction of one BiLSTM layer for a pair of site tiles
// (grid = (2 * pairs, 2 directions), cluster (2,1,1); CTA rank q owns tile 2*pair+q).
//   * h_t never touches shared memory: the epilogue packs it to fp16 hi/lo and writes it with tcgen05.st into
//     tensor memory, from where the next step's MMAs read it as the A operand (".ts" form).  Two h buffers
//     (2 x 128 columns) + two 128-column gate accumulators fill the 512 TMEM columns.
//   * a step is four gate blocks of 32 hidden units (N = 128 per pair); the accumulators ping-pong, so block b+1 is in
//     the tensor pipe while the epilogue turns block b into (c, h); nothing has to be parked in registers.
//   * FUSE_X (layer 1): the input projection x_t . W_x + b runs inside the kernel (K = 32 + 16: x_t, then two
//     constant-one columns that pick up the bias hi/lo rows of W_x), so Gx of layer 1 never exists in HBM.
//     x_t tiles arrive through a 2-stage bulk-copy ring.  !FUSE_X (layer 2): Gx (from xproj_pair) is added in the
//     epilogue, L2-prefetched two steps ahead and register double-buffered.
//   * h_t leaves for the next layer straight from registers: 16-byte stores, 512 contiguous bytes per warp.
//   Wh   : [dir][q][hl][b 4][kc 16][64 rows][8]  (row r of CTA q, block b = gate column b*128 + q*64 + r)
//   Wx   : [dir][q][hl][b 4][kc 6][64 rows][8]   (FUSE_X only; k = 32, 33 hold bias hi / lo)
//   X48  : [(t*NT+tile)][hl][kc 6][128][8]       (FUSE_X only; k = 32, 33 are 1.0)
// Warps: 0 = MMA issuer (leader CTA), 1..4G = epilogue (G warps per TMEM lane quarter), 4G+1 = loads (x_t / Gx ring).
// ---------------------------------------------------------------------------------------------
#ifndef CLAIRB_SEQ1_EARLY
#define CLAIRB_SEQ1_EARLY 1   // block-wise hand-off of h_t to the MMA issuer in layer 1 as well (0.483 -> 0.468 ms per chunk)
#endif
#ifndef CLAIRB_SEQ1_G
#define CLAIRB_SEQ1_G 2
#endif
#ifndef CLAIRB_SEQ1_ORDER
#define CLAIRB_SEQ1_ORDER 1   // issue order of blocks 0 / 1 in layer 1 (A/B switch; see the issuer)
#endif
constexpr int SEQ1_G = CLAIRB_SEQ1_G;                              // ... of the layer-1 launch (A/B switch)
constexpr int SEQ_G = 2;                                // epilogue warps per TMEM lane quarter (2 or 4)
constexpr int SEQ_THREADS = 32 * (2 + 4 * SEQ_G);
constexpr int SEQ_W_BYTES = 2 * 4 * 16 * 1024;          // 131072
constexpr int SEQ_WX_BYTES = 2 * 4 * 6 * 1024;          // 49152
constexpr int SEQ_X_STAGE = 2 * 6 * KCH_BYTES;          // 24576
constexpr int X48_TILE_HALVES = 2 * 6 * KCH;
constexpr int SEQ_G_STAGE = 16 * 128 * 16;              // Gx of 16 hidden units x 128 rows (float4 each) = 32768
constexpr int SEQ_G_RING = 3;
template <bool FUSE_X>
constexpr size_t seq_smem_bytes() {
  return (size_t)SEQ_W_BYTES + (FUSE_X ? SEQ_WX_BYTES + 2 * SEQ_X_STAGE : SEQ_G_RING * SEQ_G_STAGE) + 256 + 1024;
}

template <bool FUSE_X, int OUT, int G>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (2 + 4 * G), 1)
lstm_seq(const __half* __restrict__ Wh, const __half* __restrict__ Wx, const __half* __restrict__ X48,
         const float* __restrict__ Gx, void* __restrict__ Hout, int NT, int64_t np, int pf_dist, const int* __restrict__ x_lo_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Ws = smem;                                            // [hl][b][kc 16][64][8]
  uint8_t* Wxs = smem + SEQ_W_BYTES;                             // [hl][b][kc 6][64][8]
  uint8_t* Xs = Wxs + (FUSE_X ? SEQ_WX_BYTES : 0);               // [2 stages][hl][kc 6][128][8]
  uint8_t* Gs = Wxs;                                             // !FUSE_X: [3 stages][16 units][128 rows][4 floats]
  uint64_t* bars = (uint64_t*)(Xs + (FUSE_X ? 2 * SEQ_X_STAGE : SEQ_G_RING * SEQ_G_STAGE));
  uint64_t* acc_full = bars;           // [2] gate block complete (MMA commit, both CTAs)
  uint64_t* acc_empty = bars + 2;      // [2] (leader) accumulator drained by all 16 epilogue warps of the pair
  uint64_t* hq = bars + 18;            // [4] (leader) units 32b..32b+31 of h_t of both CTAs are in tensor memory
  uint64_t* x_full = bars + 5;         // [2] this CTA's x_t tile landed
  uint64_t* x_peer = bars + 7;         // [2] (leader) the peer's x_t tile landed
  uint64_t* x_empty = bars + 9;        // [2] x_t tile consumed (MMA commit, both CTAs)
  uint64_t* w_full = bars + 11;
  uint64_t* g_full = bars + 12;        // [3] Gx half-block landed
  uint64_t* g_empty = bars + 15;       // [3] Gx half-block consumed by its 4 epilogue warps
  uint32_t* tmem_slot = (uint32_t*)(bars + 22);

  // EARLY: hand h_t to the MMA issuer block by block, so that only the last two k-steps of a step's first gate block
  // wait for the end of the previous step (an early A/B had this slower for the fused-x kernel; with the current
  // epilogue it is 3 % faster there too, CLAIRB_SEQ1_EARLY)
  constexpr bool EARLY = CLAIRB_SEQ1_EARLY ? true : !FUSE_X;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.y;
  const int tile = blockIdx.x;                                   // = 2*pair + rank
#ifdef CLAIRB_TRACE
  // timeline probe (CLAIRB_S1_TRACE): the fused-x launch has no Gx, its parameter slot carries the stamp buffer
  long long* trace = FUSE_X ? (long long*)Gx : nullptr;
  const bool tr_on = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto stamp = [&](int s, int ev) { if (tr_on) trace[s * 64 + ev] = clock64(); };
#else
  auto stamp = [&](int, int) {};
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8 * G);
      mbar_init(&x_full[i], 1);
      mbar_init(&x_peer[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < SEQ_G_RING; ++i) {
      mbar_init(&g_full[i], 1);
      mbar_init(&g_empty[i], 4 * G);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&hq[i], 8 * G);
    mbar_init(w_full, 1);
    fence_barrier_init();
    const uint8_t* src = (const uint8_t*)Wh + ((size_t)dir * 2 + rank) * SEQ_W_BYTES;
    mbar_expect_tx(w_full, SEQ_W_BYTES + (FUSE_X ? SEQ_WX_BYTES : 0));
    for (int i = 0; i < SEQ_W_BYTES; i += 32768) bulk_g2s(Ws + i, src + i, 32768, w_full);
    if (FUSE_X) {
      const uint8_t* sx = (const uint8_t*)Wx + ((size_t)dir * 2 + rank) * SEQ_WX_BYTES;
      for (int i = 0; i < SEQ_WX_BYTES; i += 16384) bulk_g2s(Wxs + i, sx + i, 16384, w_full);
    }
  }
  if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
  __syncthreads();                                     // barrier inits visible before anyone polls them
  mbar_wait(w_full, 0);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;    // cols 0..127 acc0, 128..255 acc1, 256..383 h buffer 0 (hi 64 | lo 64), 384..511 h buffer 1

  if (warp == 0) {
    if (lane == 0 && rank == 0) {
      // ---- MMA issuer ----
      // Per step: four gate blocks b (accumulator b&1).  The recurrent part of block b contracts over all 128 units of
      // h_{s-1}, but those arrive block by block from the previous step's epilogue (hq[kq] = units 32kq..32kq+31 are in
      // tensor memory), so the k-steps of blocks 0 and 1 are issued as their slice of h becomes available; only the
      // last two k-steps (units 96..127) wait for the end of the previous step.
      const uint32_t idesc = make_idesc_f16(256, 128);
      const uint64_t wdesc = make_smem_desc(smem_u32(Ws), 1024, 128);        // recurrent kernel: 64-row k-chunks
      const uint64_t wxdesc = make_smem_desc(smem_u32(Wxs), 1024, 128);      // input kernel (FUSE_X)
      const uint64_t xdesc = make_smem_desc(smem_u32(Xs), KCH_BYTES, 128);   // x_t tile: 128-row k-chunks
      uint32_t use0 = 0, use1 = 0;
      // FUSE_X: prep_tiles48 reports whether any input value has a non-zero fp16 low part (never, for integer counts)
      const bool x_has_lo = FUSE_X && x_lo_flag != nullptr && *reinterpret_cast<const volatile int*>(x_lo_flag) != 0;
      for (int s = 0; s < T_STEPS; ++s) {
        if (!FUSE_X && s == 0) continue;               // h_{-1} = 0 and no x part: nothing to accumulate
        const uint32_t h_hi = tmem + 256 + ((s - 1) & 1) * 128, h_lo = h_hi + 64;
        const uint32_t hpar = (s - 1) & 1;
        auto acquire_acc = [&](int b) {
          const int i = b & 1;
          uint32_t& use = i ? use1 : use0;
          mbar_wait(&acc_empty[i], (use & 1) ^ 1);
          ++use;
          tc_fence_after();
        };
        auto x_part = [&](int b) {                     // x_t . W_x (+ bias through the constant-one columns)
          if (!FUSE_X) return;
          const int st = s & 1;
          if (b == 0) {
            mbar_wait(&x_full[st], (s >> 1) & 1);
            mbar_wait(&x_peer[st], (s >> 1) & 1);
            tc_fence_after();
          }
          const uint32_t d = tmem + (b & 1) * 128;
          const uint64_t a_hi0 = desc_advance(xdesc, st * SEQ_X_STAGE), a_lo0 = desc_advance(a_hi0, 6 * KCH_BYTES);
          const uint64_t b_hi0 = desc_advance(wxdesc, b * 6 * 1024), b_lo0 = desc_advance(b_hi0, 4 * 6 * 1024);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const uint64_t a_hi = desc_advance(a_hi0, j * 2 * KCH_BYTES), a_lo = desc_advance(a_lo0, j * 2 * KCH_BYTES);
            const uint64_t b_hi = desc_advance(b_hi0, j * 2 * 1024), b_lo = desc_advance(b_lo0, j * 2 * 1024);
            if (FUSE_X && (pf_dist & 2)) continue;     // timing experiment (CLAIRB_SEQ1_DBG): no tensor work
            umma_f16_pair(d, a_hi, b_hi, idesc, j != 0);
            // k-step 2 holds the constant-one columns against the bias rows: both low parts are zero by construction
            if (j < 2 && x_has_lo) umma_f16_pair(d, a_lo, b_hi, idesc, 1);
            if (j < 2) umma_f16_pair(d, a_hi, b_lo, idesc, 1);
          }
        };
        auto h_part = [&](int b, int j0, int j1) {     // k-steps j0..j1-1 of h_{s-1} . W_h for block b
          const uint32_t d = tmem + (b & 1) * 128;
          const uint64_t b_hi0 = desc_advance(wdesc, b * 16 * 1024), b_lo0 = desc_advance(b_hi0, 4 * 16 * 1024);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < j0 || j >= j1) continue;
            const uint64_t b_hi = desc_advance(b_hi0, j * 2 * 1024), b_lo = desc_advance(b_lo0, j * 2 * 1024);
            if (FUSE_X && (pf_dist & 2)) continue;
            umma_f16_pair_ts(d, h_hi + j * 8, b_hi, idesc, (FUSE_X || j != 0) ? 1u : 0u);
            umma_f16_pair_ts(d, h_lo + j * 8, b_hi, idesc, 1);
            umma_f16_pair_ts(d, h_hi + j * 8, b_lo, idesc, 1);
          }
        };
        auto wait_h = [&](int kq) {
          mbar_wait(&hq[kq], hpar);
          tc_fence_after();
        };
        if (s == 0) {                                  // FUSE_X only: input projection alone
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acquire_acc(b);
            x_part(b);
            umma_commit_pair(&acc_full[b & 1], 0b11);
          }
        } else if (!EARLY) {
          // whole-step hand-off: every block waits for all of h_{s-1} (signalled once per step on hq[3])
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acquire_acc(b);
            x_part(b);
            if (b == 0) wait_h(3);
            h_part(b, 0, 8);
            umma_commit_pair(&acc_full[b & 1], 0b11);
          }
        } else {
#if CLAIRB_SEQ1_ORDER == 0
          acquire_acc(0);
          stamp(s, 0);
          x_part(0);
          wait_h(0); stamp(s, 2); h_part(0, 0, 2);
          wait_h(1); h_part(0, 2, 4);
          wait_h(2); h_part(0, 4, 6);
          acquire_acc(1);
          stamp(s, 1);
          x_part(1);
          h_part(1, 0, 6);
          wait_h(3);
          stamp(s, 3);
          h_part(0, 6, 8);
          umma_commit_pair(&acc_full[0], 0b11);
          h_part(1, 6, 8);
          umma_commit_pair(&acc_full[1], 0b11);
          stamp(s, 4);
#else
          // Block 0 is completed as soon as the last quarter of h_{s-1} is published, and only then block 1 is issued: the
          // epilogue (which idled 2.4 k cycles per step behind the 54 recurrence-independent MMAs the old order issued
          // first) starts on block 0 one block's worth of tensor time earlier, and block 1 completes while it works.
          acquire_acc(0);
          stamp(s, 0);
          x_part(0);
          wait_h(0); stamp(s, 2); h_part(0, 0, 2);
          wait_h(1); h_part(0, 2, 4);
          wait_h(2); h_part(0, 4, 6);
          wait_h(3);
          stamp(s, 3);
          h_part(0, 6, 8);
          umma_commit_pair(&acc_full[0], 0b11);
          acquire_acc(1);
          stamp(s, 1);
          x_part(1);
          h_part(1, 0, 8);
          umma_commit_pair(&acc_full[1], 0b11);
          stamp(s, 4);
#endif
#pragma unroll
          for (int b = 2; b < 4; ++b) {
            acquire_acc(b);
            stamp(s, 6 + b);
            x_part(b);
            h_part(b, 0, 8);
            umma_commit_pair(&acc_full[b & 1], 0b11);
          }
          stamp(s, 12);
        }
        if (FUSE_X) umma_commit_pair(&x_empty[s & 1], 0b11);
      }
    }
    __syncwarp();
  } else if (warp == 1 + 4 * G) {
    if (lane == 0) {
      if (FUSE_X) {
        // ---- x_t ring: each CTA loads its own 128 rows; the peer tells the leader when its tile has landed ----
        for (int s = 0; s < T_STEPS; ++s) {
          const int st = s & 1, t = dir ? (T_STEPS - 1 - s) : s;
          mbar_wait(&x_empty[st], ((s >> 1) & 1) ^ 1);
          mbar_expect_tx(&x_full[st], SEQ_X_STAGE);
          bulk_g2s(Xs + st * SEQ_X_STAGE, X48 + ((size_t)t * NT + tile) * X48_TILE_HALVES, SEQ_X_STAGE, &x_full[st]);
          if (rank == 1) {
            mbar_wait(&x_full[st], (s >> 1) & 1);
            mbar_arrive_cluster(map_to_cta(smem_u32(&x_peer[st]), 0));   // release: see the relay note in xproj_pair
          }
        }
      } else {
        // ---- Gx ring: half-blocks of 16 hidden units (32 KB, contiguous in Gx) in consumption order ----
        uint32_t q = 0;
        for (int s = 0; s < T_STEPS; ++s) {
          const int t = dir ? (T_STEPS - 1 - s) : s;
          const float* g = Gx + ((size_t)t * NT + tile) * GX_TILE_FLOATS + (size_t)dir * (GX_TILE_FLOATS / 2);
          for (int hb = 0; hb < 8; ++hb, ++q) {
            const uint32_t st = q % SEQ_G_RING;
            mbar_wait(&g_empty[st], ((q / SEQ_G_RING) & 1) ^ 1);
            mbar_expect_tx(&g_full[st], SEQ_G_STAGE);
            bulk_g2s(Gs + st * SEQ_G_STAGE, g + (size_t)hb * (SEQ_G_STAGE / 4), SEQ_G_STAGE, &g_full[st]);
          }
        }
      }
    }
    __syncwarp();
  } else if constexpr (FUSE_X) {
    // ---- epilogue warps (lstm_epilogue: pipelined cell updates / tails, see there) ----
    static_assert(EARLY, "the shared epilogue publishes h_t block by block");
    EpiArgs ea;
    ea.tmem = tmem;
    ea.acc_full = acc_full;
    ea.leader_hq = map_to_cta(smem_u32(hq), 0);
    ea.leader_acc_empty0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
    ea.leader_acc_empty1 = map_to_cta(smem_u32(&acc_empty[1]), 0);
    ea.bias_base = 0;
    ea.Hout = Hout;
    ea.NT = NT;
    ea.np = np;
    ea.dir = dir;
    ea.tile = tile;
#ifdef CLAIRB_TRACE
    ea.trace = trace;
#else
    ea.trace = nullptr;
#endif
    lstm_epilogue<OUT, G, false, true>(ea, warp, lane);
  } else {
    // ---- epilogue warps of the Gx-fed cross-check kernel: G warps per TMEM lane quarter; a gate block (32 units) is two
    //      half-blocks of 16 units (= one Gx ring stage), and every warp takes UPS = 16/G units of EACH half-block, so
    //      all warps walk the ring stages in the same order ----
    constexpr int UPS = 16 / G;
    const int ew = warp - 1;
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int sub = ew >> 2;                           // 0..G-1
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t leader_hq = map_to_cta(smem_u32(hq), 0);
    const uint32_t leader_acc_empty0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
    const uint32_t leader_acc_empty1 = map_to_cta(smem_u32(&acc_empty[1]), 0);
    float c[4][2][UPS];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < UPS; ++i) c[b][j][i] = 0.f;

    uint32_t use0 = 0, use1 = 0, gq_idx = 0;
    const uint32_t gs_base = smem_u32(Gs);
    // Gx tiles (written by xproj_pair, HBM-resident) are pulled into L2 two steps ahead, paced by the step loop
    auto prefetch_gx = [&](int sp) {
      const int tp = dir ? (T_STEPS - 1 - sp) : sp;
      const float* g = Gx + ((size_t)tp * NT + tile) * GX_TILE_FLOATS + (size_t)dir * (GX_TILE_FLOATS / 2);
      for (int i = 0; i < 4; ++i) bulk_prefetch_l2(g + i * 16384, 65536);
    };
    if (!FUSE_X && ew == 0 && lane == 0) {
      for (int sp = 0; sp < pf_dist; ++sp) prefetch_gx(sp);
    }

    for (int s = 0; s < T_STEPS; ++s) {
      const int t = dir ? (T_STEPS - 1 - s) : s;       // bw consumes t = 32..0 (model.py:306-312)
      const bool have_acc = FUSE_X || s > 0;
      if (!FUSE_X && ew == 0 && lane == 0 && pf_dist > 0 && s + pf_dist < T_STEPS) prefetch_gx(s + pf_dist);
      const uint32_t h_st = lane_base + 256 + (s & 1) * 128;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = b & 1;
        if (have_acc) {
          uint32_t& use = i ? use1 : use0;
          mbar_wait(&acc_full[i], use & 1);
          ++use;
          tc_fence_after();
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int u0 = b * 32 + j * 16 + sub * UPS;  // first hidden unit of this thread's slice
          float4 gq[UPS];
          if (!FUSE_X) {
            const uint32_t gst = gq_idx % SEQ_G_RING;
            mbar_wait(&g_full[gst], (gq_idx / SEQ_G_RING) & 1);
            ++gq_idx;
            const uint32_t gsm = gs_base + gst * SEQ_G_STAGE + ((sub * UPS) * 128 + r) * 16;
#pragma unroll
            for (int k = 0; k < UPS; ++k) gq[k] = lds4(gsm + k * 128 * 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[gst]);
          }
          float v[4 * UPS];
          if (have_acc) {
#pragma unroll
            for (int k = 0; k < UPS / 4; ++k) tmem_ld16(lane_base + i * 128 + (j * 16 + sub * UPS) * 4 + k * 16, v + 16 * k);
            tmem_ld_wait();
            if (j == 1) {
              // this warp's part of the accumulator is in registers: hand it back to the MMA issuer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(i ? leader_acc_empty1 : leader_acc_empty0);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4 * UPS; ++k) v[k] = 0.f;
          }
          float hv[UPS];
#pragma unroll
          for (int k = 0; k < UPS; ++k) {
            float pi = v[4 * k], pg = v[4 * k + 1], pf = v[4 * k + 2], po = v[4 * k + 3];
            if (!FUSE_X) {
              pi += gq[k].x; pg += gq[k].y; pf += gq[k].z; po += gq[k].w;
            }
            hv[k] = lstm_cell(pi, pg, pf, po, c[b][j][k]);
          }
          // h_t of these units: fp16 hi/lo words (two units per word) -> tensor memory (next step's A operand) and
          // -> the next layer's operand tile in global memory
          uint32_t whi[UPS / 2], wlo[UPS / 2];
#pragma unroll
          for (int k = 0; k < UPS / 2; ++k) {
            const __half2 h2 = __floats2half2_rn(hv[2 * k], hv[2 * k + 1]);
            const float2 back = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(hv[2 * k] - back.x, hv[2 * k + 1] - back.y);
            whi[k] = *reinterpret_cast<const uint32_t*>(&h2);
            wlo[k] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          if constexpr (UPS == 8) {
            tmem_st4(h_st + (u0 >> 1), whi);
            tmem_st4(h_st + 64 + (u0 >> 1), wlo);
          } else {
            tmem_st2(h_st + (u0 >> 1), whi);
            tmem_st2(h_st + 64 + (u0 >> 1), wlo);
          }
          if (OUT == 0) {
            __half* out = (__half*)Hout + ((size_t)t * NT + tile) * (2 * 32 * KCH) + (size_t)(dir * 16 + (u0 >> 3)) * KCH + r * 8 + (u0 & 7);
            if constexpr (UPS == 8) {
              *reinterpret_cast<uint4*>(out) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
              *reinterpret_cast<uint4*>(out + 32 * KCH) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
            } else {
              *reinterpret_cast<uint2*>(out) = make_uint2(whi[0], whi[1]);
              *reinterpret_cast<uint2*>(out + 32 * KCH) = make_uint2(wlo[0], wlo[1]);
            }
          } else if (OUT == 1) {
            float* out = (float*)Hout + ((size_t)t * 2 * H + dir * H + u0) * np + (size_t)tile * 128 + r;
#pragma unroll
            for (int k = 0; k < UPS; ++k) out[(size_t)k * np] = hv[k];
          } else {
            // OUT == 2: MN-major SWIZZLE_128B tiles for the slice-dense MMA (K = time, M = site), per channel
            //   H2t[tile][c][hl][t/8 (5)][site/64 (2)][t%8][16-byte chunk ((site%64)/8) ^ (t%8)][site%8]
            // an 8x8 register transpose gives every lane one channel x 8 sites = one 16-byte chunk; the four row groups
            // of a warp land in one 64-byte run
            static_assert(OUT != 2 || UPS == 8, "the MN-major output needs 8 units per slice");
            if constexpr (UPS == 8) {
              transpose8x8_h(whi, lane);
              transpose8x8_h(wlo, lane);
              const int c = dir * H + u0 + (lane & 7);
              const int rg = r >> 3;                                  // site group of 8 within the tile (0..15)
              uint8_t* out = (uint8_t*)Hout + ((size_t)tile * 2 * H + c) * L3A_BYTES + (size_t)(t >> 3) * 2048 +
                             (rg >> 3) * 1024 + (t & 7) * 128 + (((rg & 7) ^ (t & 7)) << 4);
              *reinterpret_cast<uint4*>(out) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
              *reinterpret_cast<uint4*>(out + L3A_BYTES / 2) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
            }
          }
        }
        if (EARLY || b == 3) {
          // this warp's slice of h_t (block b; with !EARLY: all four blocks) is in tensor memory: the MMA issuer may
          // start contracting over it.  The payload is tensor memory, ordered by the tcgen05 fences.
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(leader_hq + b * 8);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

// ---------------------------------------------------------------------------------------------
// lstm_seq_x2<OUT, G>: layer 2 with its input projection streamed through the recurrent kernel (no Gx in HBM).
//   gates(t) = h1_t . W_x + b + h_{t-1} . W_h        K_x = 256: W_x hi/lo of one direction is 512 KB, i.e. 256 KB per
// CTA of the pair - it cannot be resident next to W_h (128 KB), so both operands of the x part stream through a ring:
//   stage (32 KB, K = 32) = [A hi 8K | A lo 8K | B hi 8K | B lo 8K]     (K = 16 stages were measured slower: 1.15 vs 1.09 ms)
//     A = this CTA's 128 rows of the h1_t tile (written by lstm_seq<FUSE_X>, K-major [hl][32 kc][128][8])
//     B = this CTA's 128 gate columns of TWO gate blocks: the x part is issued as M=256, N=256 pair-MMAs that fill both
//         ping-pong accumulators at once, so every A stage is used for 256 columns and h1_t streams twice per step
//         (block pairs (0,1) and (2,3)); 16 stages = 512 KB per CTA-step, all L2 hits except the first pass over h1_t.
//   Stages are loaded by BOTH CTAs with tensor-map TMA in the cta_group::2 form, whose completion bytes are signalled on
//   the LEADER's mbarrier (the plain bulk copy can only signal its own CTA, which needed a relay thread and a
//   cluster-scope release per stage: ~700 cycles on the ring's round trip).  The maps view H1 / W_x as [rows][1 KB].
// Issue order per step:  x(0,1) | h(0) | h(1) || h(2) | h(3)[k0] | x(2,3) | h(3)[k1..7]
//   x(0,1) of step s does not depend on the recurrence and fills the tensor pipe while the epilogue is still turning
//   blocks 2, 3 of step s-1 into h_{s-1}; in the second half h(2) goes first because it needs only ITS accumulator
//   drained, while the N=256 x part needs both; the two blocks of a pair complete 1.3-1.5 k cycles apart on purpose
//   (the epilogue handles them one after the other and hands the second accumulator back that much earlier).
// Epilogue warps read their whole share of an accumulator into registers and hand it back before doing any math.
// Accumulator column c of block b is gate column b*128 + c in both parts (x part: CTA q, row r -> block 2bp+q, c = r).
//   Wx   : [dir][q][bp 2][st 8][hl][kc 4][128 rows][8]    bias: [dir][512] (unit*4+gate order, gate-scaled)
// Warps: 0 = MMA issuer (leader), 1..4G = epilogue, 4G+1 = ring producer.
// ---------------------------------------------------------------------------------------------
#ifndef CLAIRB_SX_KC
#define CLAIRB_SX_KC 4
#endif
constexpr int SX_KC = CLAIRB_SX_KC;                     // k-chunks of 8 per ring stage (2: K = 16 per stage, 4: K = 32)
constexpr int SX_STAGE = 4 * SX_KC * KCH_BYTES;         // A hi | A lo | B hi | B lo
constexpr int SX_RING = 98304 / SX_STAGE;
constexpr int SX_NST = 32 / SX_KC;                      // stages per block pair (K = 256)
#ifndef CLAIRB_SX_G
#define CLAIRB_SX_G 2
#endif
constexpr int SX_G = CLAIRB_SX_G;                                 // epilogue warps per TMEM lane quarter (2 or 4)
constexpr int SX_THREADS = 32 * (2 + 4 * SX_G);
constexpr size_t seqx_smem_bytes() { return (size_t)SEQ_W_BYTES + SX_RING * SX_STAGE + 2048 + 256 + 128; }

// 2-D tiled tensor-map load, issued by each CTA of a pair for its own shared memory; completion bytes go to the
// mbarrier at the same offset in the LEADER CTA (cluster address `leader_bar`)
__device__ __forceinline__ void tma2_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
// arrive + expect_tx on an mbarrier of another CTA of the cluster (no payload of this thread to publish: relaxed)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}

template <int OUT, int G>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (2 + 4 * G), 1)
lstm_seq_x2(const __half* __restrict__ Wh, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const float* __restrict__ bias, const __half* __restrict__ H1, void* __restrict__ Hout, int NT, int64_t np,
            int dbg, long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint8_t* Ws = smem;                                            // [hl][b][kc 16][64][8]
  uint8_t* ring = smem + SEQ_W_BYTES;                            // [3 stages][A hi | A lo | B hi | B lo]
  float* bias_s = (float*)(ring + SX_RING * SX_STAGE);           // [128 units][4 gates]
  uint64_t* bars = (uint64_t*)(bias_s + 512);
  uint64_t* acc_full = bars;           // [2] gate block complete (MMA commit, both CTAs)
  uint64_t* acc_empty = bars + 2;      // [2] (leader) accumulator drained by all epilogue warps of the pair
  uint64_t* hq = bars + 4;             // [4] (leader) units 32b..32b+31 of h_t of both CTAs are in tensor memory
  uint64_t* full = bars + 8;           // [<=6] (leader) the stage of BOTH CTAs landed (2 arrivals + 2 x stage bytes)
  uint64_t* empty = bars + 14;         // [<=6] stage consumed (MMA commit, both CTAs)
  uint64_t* w_full = bars + 20;
  uint32_t* tmem_slot = (uint32_t*)(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.y;
  const int tile = blockIdx.x;                                   // = 2*pair + rank
  // timeline probe (CLAIRB_SX_TRACE): clock64 stamps of the leader CTA of pair 0, direction 0: [step][64 events]
#ifdef CLAIRB_TRACE
  const bool tr_on = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto stamp = [&](int s, int ev) { if (tr_on) trace[s * 64 + ev] = clock64(); };
#else
  auto stamp = [&](int, int) {};                       // probes compiled out (build with -DCLAIRB_TRACE to enable)
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8 * G);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&hq[i], 8 * G);
    for (int i = 0; i < SX_RING; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
    const uint8_t* src = (const uint8_t*)Wh + ((size_t)dir * 2 + rank) * SEQ_W_BYTES;
    mbar_expect_tx(w_full, SEQ_W_BYTES);
    for (int i = 0; i < SEQ_W_BYTES; i += 32768) bulk_g2s(Ws + i, src + i, 32768, w_full);
  }
  if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
  for (int i = threadIdx.x; i < 512; i += 32 * (2 + 4 * G)) bias_s[i] = bias[dir * 512 + i];
  __syncthreads();                                     // barrier inits visible before anyone polls them
  mbar_wait(w_full, 0);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;    // cols 0..127 acc0, 128..255 acc1, 256..383 h buffer 0 (hi 64 | lo 64), 384..511 h buffer 1

  if (warp == 0) {
    if (lane == 0 && rank == 0) {
      // ---- MMA issuer ----
      const uint32_t idesc_x = make_idesc_f16(256, 256), idesc_h = make_idesc_f16(256, 128);
      const uint64_t wdesc = make_smem_desc(smem_u32(Ws), 1024, 128);        // recurrent kernel: 64-row k-chunks
      const uint64_t rdesc = make_smem_desc(smem_u32(ring), KCH_BYTES, 128); // ring operands: 128-row k-chunks
      uint32_t slot = 0, par = 0, use0 = 0, use1 = 0;
      for (int s = 0; s < T_STEPS; ++s) {
        const uint32_t h_hi = tmem + 256 + ((s - 1) & 1) * 128, h_lo = h_hi + 64;
        const uint32_t hpar = (s - 1) & 1;
        auto h_part = [&](int b, int j0, int j1, bool fresh) {   // k-steps j0..j1-1 of h_{s-1} . W_h for block b
          const uint32_t d = tmem + (b & 1) * 128;
          const uint64_t b_hi0 = desc_advance(wdesc, b * 16 * 1024), b_lo0 = desc_advance(b_hi0, 4 * 16 * 1024);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < j0 || j >= j1) continue;
            const uint64_t b_hi = desc_advance(b_hi0, j * 2 * 1024), b_lo = desc_advance(b_lo0, j * 2 * 1024);
            if (dbg & 2) continue;                     // timing experiment: no tensor work
            umma_f16_pair_ts(d, h_hi + j * 8, b_hi, idesc_h, (fresh && j == 0) ? 0u : 1u);
            umma_f16_pair_ts(d, h_lo + j * 8, b_hi, idesc_h, 1);
            umma_f16_pair_ts(d, h_hi + j * 8, b_lo, idesc_h, 1);
          }
        };
        auto x_part = [&](bool fresh) {                // x part of one block pair: N = 256 into acc0|acc1
#pragma unroll 1
          for (int st = 0; st < SX_NST; ++st) {
            mbar_wait(&full[slot], par);
            tc_fence_after();
            const uint64_t a_hi0 = desc_advance(rdesc, slot * SX_STAGE), a_lo0 = desc_advance(a_hi0, SX_STAGE / 4);
            const uint64_t b_hi0 = desc_advance(a_hi0, SX_STAGE / 2), b_lo0 = desc_advance(a_hi0, 3 * (SX_STAGE / 4));
#pragma unroll
            for (int kk = 0; kk < SX_KC / 2; ++kk) {
              const uint64_t a_hi = desc_advance(a_hi0, kk * 2 * KCH_BYTES), a_lo = desc_advance(a_lo0, kk * 2 * KCH_BYTES);
              const uint64_t b_hi = desc_advance(b_hi0, kk * 2 * KCH_BYTES), b_lo = desc_advance(b_lo0, kk * 2 * KCH_BYTES);
              if (dbg & 2) continue;
              umma_f16_pair(tmem, a_hi, b_hi, idesc_x, (fresh && st == 0 && kk == 0) ? 0u : 1u);
              umma_f16_pair(tmem, a_lo, b_hi, idesc_x, 1);
              umma_f16_pair(tmem, a_hi, b_lo, idesc_x, 1);
            }
            umma_commit_pair(&empty[slot], 0b11);
            if (++slot == SX_RING) { slot = 0; par ^= 1; }
          }
        };
        auto acquire = [&](int i) {
          uint32_t& use = i ? use1 : use0;
          mbar_wait(&acc_empty[i], (use & 1) ^ 1);
          ++use;
          tc_fence_after();
        };
        auto wait_h = [&](int kq) {
          mbar_wait(&hq[kq], hpar);
          tc_fence_after();
        };
        // ---- blocks 0, 1: x part first (independent of the recurrence), then the recurrent parts one block after the
        //      other: block 0 completes a whole h part (1.5 k cycles) before block 1, so the epilogue - which handles
        //      the two blocks back to back - starts that much earlier and hands accumulator 1 back that much earlier ----
        acquire(0);
        acquire(1);
        stamp(s, 0);
        x_part(true);
        stamp(s, 1);
        if (s > 0) {
          wait_h(0); stamp(s, 2); h_part(0, 0, 2, false);
          wait_h(1); h_part(0, 2, 4, false);
          wait_h(2); h_part(0, 4, 6, false);
          wait_h(3); stamp(s, 3); h_part(0, 6, 8, false);
          umma_commit_pair(&acc_full[0], 0b11);
          h_part(1, 0, 8, false);
          umma_commit_pair(&acc_full[1], 0b11);
        } else {
          umma_commit_pair(&acc_full[0], 0b11);
          umma_commit_pair(&acc_full[1], 0b11);
        }
        stamp(s, 4);
        // ---- blocks 2, 3: h(2) needs only accumulator 0 (handed back early) and fills the wait for accumulator 1; the
        //      first k-step of h(3) initialises accumulator 1, the N=256 x part accumulates into both, block 2 is
        //      complete, and the rest of h(3) staggers block 3 behind it ----
        acquire(0);
        stamp(s, 8);
        if (s > 0) h_part(2, 0, 8, true);
        acquire(1);
        stamp(s, 9);
        if (s > 0) h_part(3, 0, 1, true);
        x_part(s == 0);
        umma_commit_pair(&acc_full[0], 0b11);
        if (s > 0) h_part(3, 1, 8, false);
        umma_commit_pair(&acc_full[1], 0b11);
        stamp(s, 12);
      }
    }
    __syncwarp();
  } else if (warp == 1 + 4 * G) {
    if (lane == 0) {
      // ---- ring producer: each CTA loads its own rows of h1_t and its own gate columns of W_x; both signal the leader ----
      const uint32_t leader_full = map_to_cta(smem_u32(full), 0);
      const int wrow0 = (dir * 2 + (int)rank) * (2 * SX_NST * 4 * SX_KC);   // 1 KB rows of W_x: 4*SX_KC per stage
      uint32_t slot = 0, par = 1;
      for (int s = 0; s < T_STEPS; ++s) {
        const int t = dir ? (T_STEPS - 1 - s) : s;
        const int arow0 = (t * NT + tile) * 128;                          // 1 KB rows of H1: 128 per tile (hi 64 | lo 64)
        if (s + 1 < T_STEPS) {
          // the next step's h1 tile: HBM -> L2 while this step streams
          const int tn = dir ? (T_STEPS - 2 - s) : s + 1;
          const uint8_t* an = (const uint8_t*)(H1 + ((size_t)tn * NT + tile) * (2 * 32 * KCH));
          for (int i = 0; i < 4; ++i) bulk_prefetch_l2(an + i * 32768, 32768);
        }
        for (int bp = 0; bp < 2; ++bp)
          for (int st = 0; st < SX_NST; ++st) {
            mbar_wait(&empty[slot], par);
            uint8_t* dst = ring + slot * SX_STAGE;
            const uint32_t bar = leader_full + slot * 8;
            if (dbg & 1) {                             // timing experiment: no operand traffic
              if (rank == 0) mbar_arrive(&full[slot]); else mbar_arrive_cluster_relaxed(bar);
            } else {
              if (rank == 0) mbar_expect_tx(&full[slot], SX_STAGE); else mbar_expect_tx_cluster(bar, SX_STAGE);
              tma2_pair(dst, &tmA, 0, arow0 + st * 2 * SX_KC, bar);
              tma2_pair(dst + SX_STAGE / 4, &tmA, 0, arow0 + 64 + st * 2 * SX_KC, bar);
              tma2_pair(dst + SX_STAGE / 2, &tmB, 0, wrow0 + (bp * SX_NST + st) * 4 * SX_KC, bar);
            }
            if (++slot == SX_RING) { slot = 0; par ^= 1; }
          }
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue warps (lstm_epilogue: pipelined cell updates / tails, see there) ----
    EpiArgs ea;
    ea.tmem = tmem;
    ea.acc_full = acc_full;
    ea.leader_hq = map_to_cta(smem_u32(hq), 0);
    ea.leader_acc_empty0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
    ea.leader_acc_empty1 = map_to_cta(smem_u32(&acc_empty[1]), 0);
    ea.bias_base = smem_u32(bias_s);
    ea.Hout = Hout;
    ea.NT = NT;
    ea.np = np;
    ea.dir = dir;
    ea.tile = tile;
    ea.trace = trace;
#ifdef CLAIRB_X2_PROBE_PLAIN                            // timing probe: layer 2 with layer 1's tail (no bias, no transposes); results are garbage
    lstm_epilogue<0, G, false, false>(ea, warp, lane);
#elif defined(CLAIRB_X2_PROBE_NOBIAS)
    lstm_epilogue<OUT, G, false, false>(ea, warp, lane);
#else
    lstm_epilogue<OUT, G, true, false>(ea, warp, lane);
#endif
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

// prep for the fused layer-1 kernel: like prep_tiles but K = 48 (k = 32, 33 are the constant-one bias columns)
template <typename TIn>
__global__ void __launch_bounds__(128) prep_tiles48(const TIn* __restrict__ x, __half* __restrict__ X48, int64_t n, int NT,
                                                   int* __restrict__ lo_flag) {
  const int tile = blockIdx.x, t = blockIdx.y, r = threadIdx.x;
  const int64_t site = (int64_t)tile * 128 + r;
  float v[32];
  if (site < n) {
    const TIn* src = x + site * SITE_ELEMS + t * F_IN;
    if constexpr (sizeof(TIn) == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 q = *reinterpret_cast<const float4*>(src + 4 * i);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q = *reinterpret_cast<const uint4*>(src + 8 * i);
        const int16_t* e = reinterpret_cast<const int16_t*>(&q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * i + j] = (float)e[j];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
  }
  __half* base = X48 + ((size_t)t * NT + tile) * X48_TILE_HALVES;
  uint32_t any_lo = 0;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint4 hi, lo;
    split8(v + 8 * kc, hi, lo);
    *reinterpret_cast<uint4*>(base + kc * KCH + r * 8) = hi;
    *reinterpret_cast<uint4*>(base + 6 * KCH + kc * KCH + r * 8) = lo;
    any_lo |= (lo.x | lo.y | lo.z | lo.w) & 0x7fff7fffu;        // -0.0 is still zero
  }
  // the generator's integer counts are exact in fp16 (|x| <= 2048): the layer-1 kernel then skips the x_lo . W_hi products
  if (any_lo) *lo_flag = 1;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<uint4*>(base + 4 * KCH + r * 8) = make_uint4(0x3C003C00u, 0, 0, 0);     // k = 32, 33: fp16 1.0
  *reinterpret_cast<uint4*>(base + 5 * KCH + r * 8) = zero;
  *reinterpret_cast<uint4*>(base + 10 * KCH + r * 8) = zero;
  *reinterpret_cast<uint4*>(base + 11 * KCH + r * 8) = zero;
}

// ---------------------------------------------------------------------------------------------
// l3l4_fused: slice-dense L3 (model.py:225-244, 464-471) chained into dense L4 (model.py:482-488) per 128-site tile.
//   for every channel c (256):  S_c[128 x 32] = selu(A_c[128 x t] . W3_c[t x o] + b3_c)      (o padded 30 -> 32)
//                               D4[128 x 192] += S_c . W4[(o,c), :]                          (L4 rows regrouped by channel)
//   then l4[128 x 192] = selu(D4 + b4) -> planes l4T[192][np] fp32 for the heads kernel.
// The [B,30,256] -> [B,7680] flatten (model.py:474-478, index o*256+c) never materialises: S_c goes TMEM -> registers ->
// fp16 hi/lo pairs back into TENSOR MEMORY, from where the L4 MMAs read it as their A operand.  A_c arrives MN-major,
// 128B-swizzled (K = time) from lstm_seq_x2; the k-group t = 40..47 of the third k-step points at a shared zero block
// through the descriptor's leading-dimension offset.
// What bounds it (timeline probe CLAIRB_LF_TRACE, tools/lf_trace.py): the time a channel keeps its ring stage - load
// ~1.0-1.5 k cycles, L3 ~0.9 k, epilogue ~1.9 k, L4 ~0.9 k - divided by the ring depth.  Hence
//   * 4 ring stages of (A tile + weight blob) = 205 KB: possible because S_c lives in tensor memory, not in smem;
//   * the MMAs of a tile are issued by TWO threads (warp 1: L3, warp 10: L4): one issuing thread needed ~100 cycles per
//     MMA plus ~400 cycles between bursts (it shares a scheduler with two epilogue warps) and capped the kernel at
//     2 x (400 + 640) = 2.05 k cycles per channel;
//   * L3 of a channel is two independent accumulation chains instead of one chain of 9 (links that accumulate into the
//     same columns are ~90 cycles apart however small N is):
//         D3[ 0..63] += A_hi . [W_hi ; W_lo]^T   (N = 64: hi.hi | hi.lo)      D3[64..95] += A_lo . W_hi^T   (N = 32)
//     interleaved over the 3 k-steps; the epilogue adds the three 32-column groups;
//   * every shared-memory descriptor is precomputed per ring stage and the loops are unrolled over the 4 stages.
// Warp roles: 0 = producer (one A tile + one weight blob per channel), 1 = L3 MMA issuer, 2..9 = epilogue: group
// g = (warp-2)/4 owns channels c = g mod 2 with its own D3 / S_c buffers, 10 = L4 MMA issuer.
// ---------------------------------------------------------------------------------------------
constexpr int LF_STAGES = 4;
constexpr int LF_STAGE_BYTES = L3A_BYTES + L3L4_BLOB_BYTES;          // 51328
constexpr int LF_THREADS = 352;
constexpr size_t l3l4_smem_bytes() { return (size_t)LF_STAGES * LF_STAGE_BYTES + 2048 + 256 + 1024; }

__global__ void __launch_bounds__(LF_THREADS, 1)
l3l4_fused(const __half* __restrict__ H2t, const uint8_t* __restrict__ blobs, const float* __restrict__ b4,
           float* __restrict__ l4T, __half* __restrict__ L4t, int64_t np, float* __restrict__ l3_dbg, int pf_dist,
           long long* __restrict__ trace, float w4_unscale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint8_t* zero = stages + LF_STAGES * LF_STAGE_BYTES; // 2 KB of zeros (k-group t = 40..47)
  uint64_t* bars = (uint64_t*)(zero + 2048);
  uint64_t* stage_full = bars;          // [4]
  uint64_t* stage_empty = bars + 4;     // [4]
  uint64_t* d3_full = bars + 8;         // [2]
  uint64_t* d3_empty = bars + 10;       // [2]
  uint64_t* a4_full = bars + 12;        // [2] S_c of group g is in tensor memory (4 epilogue warps)
  uint64_t* a4_empty = bars + 14;       // [2] S_c consumed (L4 MMA commit)
  uint64_t* d4_full = bars + 16;
  uint32_t* tmem_slot = (uint32_t*)(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
#ifdef CLAIRB_TRACE
  const bool tr_on = trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int c, int ev) { if (tr_on) trace[c * 16 + ev] = clock64(); };
#else
  auto stamp = [&](int, int) {};                       // probes compiled out (build with -DCLAIRB_TRACE to enable)
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < LF_STAGES; ++i) { mbar_init(&stage_full[i], 1); mbar_init(&stage_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d3_full[i], 1); mbar_init(&d3_empty[i], 4);
      mbar_init(&a4_full[i], 4); mbar_init(&a4_empty[i], 1);
    }
    mbar_init(d4_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2048 / 16; i += LF_THREADS) reinterpret_cast<uint4*>(zero)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;     // cols 0..191: D4, 192..287: D3[0] (hi.hi | hi.lo | lo.hi), 288..383: D3[1],
                                        // 384..415: S[0] (fp16 pairs: hi 16 cols | lo 16 cols), 416..447: S[1]

  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* a_src = (const uint8_t*)(H2t + (size_t)tile * 2 * H * L3A_HALVES);
      // pull the LSTM2 tiles of the next channels into L2 ahead of the bulk copies (the weight blobs are shared by all
      // CTAs and stay L2-resident by themselves)
      for (int c = 0; c < pf_dist && c < 2 * H; ++c) bulk_prefetch_l2(a_src + (size_t)c * L3A_BYTES, L3A_BYTES);
      for (int c = 0; c < 2 * H; ++c) {
        if (c + pf_dist < 2 * H && pf_dist > 0) bulk_prefetch_l2(a_src + (size_t)(c + pf_dist) * L3A_BYTES, L3A_BYTES);
        const int st = c % LF_STAGES;
        mbar_wait(&stage_empty[st], ((c / LF_STAGES) & 1) ^ 1);
        stamp(c, 0);
        uint8_t* dst = stages + st * LF_STAGE_BYTES;
        mbar_expect_tx(&stage_full[st], LF_STAGE_BYTES);
        bulk_g2s(dst, a_src + (size_t)c * L3A_BYTES, L3A_BYTES, &stage_full[st]);
        bulk_g2s(dst + L3A_BYTES, blobs + (size_t)c * L3L4_BLOB_BYTES, L3L4_BLOB_BYTES, &stage_full[st]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- L3 issuer ----
      const uint32_t idesc3a = make_idesc_f16_amn(128, 64), idesc3b = make_idesc_f16_amn(128, 32);
      const uint32_t zero_addr = smem_u32(zero);
      const uint64_t sw128 = (uint64_t)2 << 61;
      uint64_t a3[LF_STAGES][2][3], b3[LF_STAGES][3];
#pragma unroll
      for (int st = 0; st < LF_STAGES; ++st) {
        const uint32_t a_base = smem_u32(stages + st * LF_STAGE_BYTES), w_base = a_base + L3A_BYTES;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            // MN-major SWIZZLE_128B A: LBO = distance between the two 64-site atoms, SBO = distance between the two
            // time groups of this k-step (for t = 32..47 the second group is the shared zero block)
            const uint32_t addr = a_base + hl * (L3A_BYTES / 2) + j * 2 * 2048;
            a3[st][hl][j] = sw128 | (j < 2 ? make_smem_desc(addr, 1024, 2048) : make_smem_desc(addr, 1024, zero_addr - addr));
          }
          b3[st][j] = make_smem_desc(w_base + j * 2 * 1024, 1024, 128);      // [kc][hi 32 rows | lo 32 rows][8]
        }
      }
      for (int c0 = 0; c0 < 2 * H; c0 += LF_STAGES) {
#pragma unroll
        for (int u = 0; u < LF_STAGES; ++u) {
          const int c = c0 + u, st = u, g = u & 1;
          mbar_wait(&stage_full[st], (c / LF_STAGES) & 1);
          mbar_wait(&d3_empty[g], ((c >> 1) & 1) ^ 1);
          tc_fence_after();
          stamp(c, 2);
          const uint32_t d3 = tmem + 192 + g * 96;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            umma_f16(d3, a3[st][0][j], b3[st][j], idesc3a, j != 0);
            umma_f16(d3 + 64, a3[st][1][j], b3[st][j], idesc3b, j != 0);
          }
          umma_commit(&d3_full[g]);
          stamp(c, 3);
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    if (lane == 0) {
      // ---- L4 issuer: D4[128 x 192] += S_c . W4_c as soon as the epilogue has written S_c; frees the S_c buffer and
      //      the ring stage of channel c ----
      const uint32_t idesc4 = make_idesc_f16(128, 192);
      uint64_t b4d[LF_STAGES][2][2];
#pragma unroll
      for (int st = 0; st < LF_STAGES; ++st) {
        const uint32_t w_base = smem_u32(stages + st * LF_STAGE_BYTES) + L3A_BYTES + 2 * L3W_BYTES;
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            b4d[st][hl][kk] = make_smem_desc(w_base + hl * L4W_BYTES + kk * 2 * (192 * 16), 192 * 16, 128);
      }
      for (int c0 = 0; c0 < 2 * H; c0 += LF_STAGES) {
#pragma unroll
        for (int u = 0; u < LF_STAGES; ++u) {
          const int c = c0 + u, st = u, g = u & 1;
          mbar_wait(&a4_full[g], (c >> 1) & 1);
          tc_fence_after();
          stamp(c, 4);
          const uint32_t s_hi = tmem + 384 + g * 32, s_lo = s_hi + 16;     // A operand from tensor memory
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_f16_ts(tmem, s_hi + kk * 8, b4d[st][0][kk], idesc4, (c | kk) != 0);
            umma_f16_ts(tmem, s_lo + kk * 8, b4d[st][0][kk], idesc4, 1);
            umma_f16_ts(tmem, s_hi + kk * 8, b4d[st][1][kk], idesc4, 1);
          }
          umma_commit(&a4_empty[g]);
          umma_commit(&stage_empty[st]);
          stamp(c, 5);
        }
      }
      umma_commit(d4_full);
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int c = g; c < 2 * H; c += 2) {
      const int use = c >> 1, st = c % LF_STAGES;
      mbar_wait(&d3_full[g], use & 1);
      tc_fence_after();
      if (lane == 0 && quarter == 0) stamp(c, 8);
      float v[32];
      {
        // the three partial products of the hi/lo split sit in three 32-column groups: add them up
        float p1[32], p2[32];
        const uint32_t d3 = lane_addr + 192 + g * 96;
        tmem_ld16(d3, v);
        tmem_ld16(d3 + 16, v + 16);
        tmem_ld16(d3 + 32, p1);
        tmem_ld16(d3 + 48, p1 + 16);
        tmem_ld16(d3 + 64, p2);
        tmem_ld16(d3 + 80, p2 + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d3_empty[g]);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += p1[i] + p2[i];
      }
      const float* b3 = reinterpret_cast<const float*>(stages + st * LF_STAGE_BYTES + L3A_BYTES + 2 * L3W_BYTES + 2 * L4W_BYTES);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = v[i] + b3[i];
        v[i] = x >= 0.f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * (__expf(x) - 1.f);
      }
      if (l3_dbg != nullptr) {
        // parity hook only (clairb_get_layer): the [n,30,256] activations the production path never materialises
#pragma unroll
        for (int i = 0; i < L3_UNITS; ++i) l3_dbg[((size_t)i * 2 * H + c) * np + (size_t)tile * 128 + r] = v[i];
      }
      // S_c as fp16 hi/lo pairs -> tensor memory (the A operand of the L4 MMAs)
      uint32_t whi[16], wlo[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const __half2 h2 = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const float2 back = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * q] - back.x, v[2 * q + 1] - back.y);
        whi[q] = *reinterpret_cast<const uint32_t*>(&h2);
        wlo[q] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      if (lane == 0 && quarter == 0) stamp(c, 9);
      mbar_wait(&a4_empty[g], (use & 1) ^ 1);
      tc_fence_after();
      if (lane == 0 && quarter == 0) stamp(c, 10);
      tmem_st16(lane_addr + 384 + g * 32, whi);
      tmem_st16(lane_addr + 384 + g * 32 + 16, wlo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a4_full[g]);
      if (lane == 0) stamp(c, 11 + quarter);
    }
    // ---- L4 epilogue: group g takes columns g*96..+96 ----
    mbar_wait(d4_full, 0);
    tc_fence_after();
    float* out = l4T + (size_t)tile * 128 + r;
#pragma unroll
    for (int cb = 0; cb < 96; cb += 16) {
      const int col0 = g * 96 + cb;
      float v[16];
      tmem_ld16(lane_addr + col0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x = fmaf(v[i], w4_unscale, __ldg(b4 + col0 + i));      // W4 was uploaded scaled by a power of two
        v[i] = x >= 0.f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * (__expf(x) - 1.f);
        out[(size_t)(col0 + i) * np] = v[i];           // fp32 planes: parity hook / CUDA-core heads
      }
      // the same activations as the K-major fp16 hi/lo operand tile of heads_tc: [tile][hl][kc 24][128][8]
      __half* lt = L4t + (size_t)tile * (2 * 24 * KCH) + (size_t)(col0 / 8) * KCH + r * 8;
      uint4 hi, lo;
      split8(v, hi, lo);
      *reinterpret_cast<uint4*>(lt) = hi;
      *reinterpret_cast<uint4*>(lt + 24 * KCH) = lo;
      split8(v + 8, hi, lo);
      *reinterpret_cast<uint4*>(lt + KCH) = hi;
      *reinterpret_cast<uint4*>(lt + 25 * KCH) = lo;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// ---------------------------------------------------------------------------------------------
// heads_tc: L5_1..4 (192 -> 96, SELU), the four heads (96 -> 21/3/33/33, SELU) and their softmax
// (model.py:507-622) for one 128-site tile, on tensor cores.
//   L4t   : [tile][hl][kc 24][128][8] fp16   L4 activations as a K-major operand tile (written by l3l4_fused)
//   blob  : per head k: W5_k as two 48-column halves [hf][hl][kc 24][48][8], Whd_k [hl][kc 12][48][8] (rows >= n_k
//           zero), b5_k[96], bh_k[48]        (HEAD_BLOB_BYTES each)
// Per head: two L5 MMAs groups (N = 48 each) into D5 -> epilogue (+b5, SELU, fp16 hi/lo) -> smem operand -> head MMA
// (N = 48) into D6 -> epilogue (+bh, SELU = the reference's "logits"; softmax over the n_k real columns).
// Strictly sequential per tile (each stage is tiny); one CTA per tile, the grid is one wave.
// Warps: 0 = producer, 1 = MMA issuer, 2..5 = epilogue (one per TMEM lane quarter).
// ---------------------------------------------------------------------------------------------
constexpr int HD_A4_BYTES = 2 * 24 * KCH_BYTES;              // 98304
constexpr int HD_W5_BYTES = 2 * 24 * 48 * 16;                // 36864 (one 48-column half, hi|lo)
constexpr int HD_A5_BYTES = 2 * 12 * KCH_BYTES;              // 49152
constexpr int HD_WH_BYTES = 2 * 12 * 48 * 16;                // 18432
constexpr int HEAD_BLOB_BYTES = 2 * HD_W5_BYTES + HD_WH_BYTES + 96 * 4 + 48 * 4;   // 92736
constexpr int HD_THREADS = 192;
constexpr size_t heads_smem_bytes() { return (size_t)HD_A4_BYTES + HD_W5_BYTES + HD_A5_BYTES + HD_WH_BYTES + 1024 + 256 + 1024; }

__global__ void __launch_bounds__(HD_THREADS, 1)
heads_tc(const __half* __restrict__ L4t, const uint8_t* __restrict__ blobs, float* __restrict__ probs,
         float* __restrict__ logits, int64_t n, int64_t split_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* A4 = smem;
  uint8_t* W5 = A4 + HD_A4_BYTES;
  uint8_t* A5 = W5 + HD_W5_BYTES;
  uint8_t* WH = A5 + HD_A5_BYTES;
  float* bias = (float*)(WH + HD_WH_BYTES);            // [4][96 + 48] floats would not fit: per head, reloaded: [96 | 48]
  uint64_t* bars = (uint64_t*)((uint8_t*)bias + 1024);
  uint64_t* a4_full = bars;        // L4 tile landed
  uint64_t* w5_full = bars + 1;    // W5 half landed
  uint64_t* w5_free = bars + 2;    // W5 half consumed (MMA commit)
  uint64_t* d5_full = bars + 3;    // both halves of D5 complete
  uint64_t* d5_free = bars + 4;    // D5 drained (4 epilogue warps)
  uint64_t* a5_full = bars + 5;    // L5 operand tile written (4 epilogue warps)
  uint64_t* wh_full = bars + 6;    // head weights + biases landed
  uint64_t* d6_full = bars + 7;    // head accumulator complete (also: A5 and WH consumed)
  uint64_t* d6_free = bars + 8;    // D6 drained (4 epilogue warps)
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(a4_full, 1); mbar_init(w5_full, 1); mbar_init(w5_free, 1); mbar_init(d5_full, 1);
    mbar_init(d5_free, 4); mbar_init(a5_full, 4); mbar_init(wh_full, 1); mbar_init(d6_full, 1); mbar_init(d6_free, 4);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;     // cols 0..95: D5, 96..143: D6

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a4_full, HD_A4_BYTES);
      for (int i = 0; i < HD_A4_BYTES; i += 32768)
        bulk_g2s(A4 + i, (const uint8_t*)L4t + (size_t)tile * HD_A4_BYTES + i, 32768, a4_full);
      for (int k = 0; k < 4; ++k) {
        const uint8_t* blob = blobs + (size_t)k * HEAD_BLOB_BYTES;
        for (int hf = 0; hf < 2; ++hf) {
          const int u = k * 2 + hf;
          mbar_wait(w5_free, (u & 1) ^ 1);
          mbar_expect_tx(w5_full, HD_W5_BYTES);
          bulk_g2s(W5, blob + (size_t)hf * HD_W5_BYTES, HD_W5_BYTES, w5_full);
        }
        // head weights + the two bias vectors (contiguous in the blob); WH / bias are free once the epilogue of
        // head k-1 has read its accumulator and biases
        if (k > 0) mbar_wait(d6_free, (k - 1) & 1);
        mbar_expect_tx(wh_full, HD_WH_BYTES + 576);
        bulk_g2s(WH, blob + 2 * HD_W5_BYTES, HD_WH_BYTES + 576, wh_full);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 48);
      const uint64_t a4d = make_smem_desc(smem_u32(A4), KCH_BYTES, 128), a5d = make_smem_desc(smem_u32(A5), KCH_BYTES, 128);
      const uint64_t w5d = make_smem_desc(smem_u32(W5), 768, 128), whd = make_smem_desc(smem_u32(WH), 768, 128);
      mbar_wait(a4_full, 0);
      for (int k = 0; k < 4; ++k) {
        mbar_wait(d5_free, (k & 1) ^ 1);
        for (int hf = 0; hf < 2; ++hf) {
          const int u = k * 2 + hf;
          mbar_wait(w5_full, u & 1);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const uint64_t a_hi = desc_advance(a4d, j * 2 * KCH_BYTES), a_lo = desc_advance(a_hi, 24 * KCH_BYTES);
            const uint64_t b_hi = desc_advance(w5d, j * 2 * 768), b_lo = desc_advance(b_hi, 24 * 768);
            umma_f16(tmem + hf * 48, a_hi, b_hi, idesc, j != 0);
            umma_f16(tmem + hf * 48, a_lo, b_hi, idesc, 1);
            umma_f16(tmem + hf * 48, a_hi, b_lo, idesc, 1);
          }
          umma_commit(w5_free);
        }
        umma_commit(d5_full);
        mbar_wait(a5_full, k & 1);
        mbar_wait(wh_full, k & 1);
        mbar_wait(d6_free, (k & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint64_t a_hi = desc_advance(a5d, j * 2 * KCH_BYTES), a_lo = desc_advance(a_hi, 12 * KCH_BYTES);
          const uint64_t b_hi = desc_advance(whd, j * 2 * 768), b_lo = desc_advance(b_hi, 12 * 768);
          umma_f16(tmem + 96, a_hi, b_hi, idesc, j != 0);
          umma_f16(tmem + 96, a_lo, b_hi, idesc, 1);
          umma_f16(tmem + 96, a_hi, b_lo, idesc, 1);
        }
        umma_commit(d6_full);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int64_t site = (int64_t)tile * 128 + r;
    for (int k = 0; k < 4; ++k) {
      // ---- L5_k: D5 -> +b5 -> SELU -> fp16 hi/lo operand tile ----
      mbar_wait(d5_full, k & 1);
      tc_fence_after();
      mbar_wait(wh_full, k & 1);                       // biases of head k ride with the head weights
      const float* b5 = bias;                          // [96]
      const float* bh = bias + 96;                     // [48]
      if (k > 0) mbar_wait(d6_full, (k - 1) & 1);      // head k-1 has finished reading A5
#pragma unroll
      for (int cb = 0; cb < 96; cb += 16) {
        float v[16];
        tmem_ld16(lane_addr + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x = v[i] + b5[cb + i];
          v[i] = x >= 0.f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * (__expf(x) - 1.f);
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(A5 + (cb / 8) * KCH_BYTES + r * 16) = hi;
        *reinterpret_cast<uint4*>(A5 + 12 * KCH_BYTES + (cb / 8) * KCH_BYTES + r * 16) = lo;
        split8(v + 8, hi, lo);
        *reinterpret_cast<uint4*>(A5 + (cb / 8 + 1) * KCH_BYTES + r * 16) = hi;
        *reinterpret_cast<uint4*>(A5 + 12 * KCH_BYTES + (cb / 8 + 1) * KCH_BYTES + r * 16) = lo;
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(d5_free);
        mbar_arrive(a5_full);
      }
      // ---- head k: D6 -> +bh -> SELU (= the reference's *_logits) -> softmax ----
      mbar_wait(d6_full, k & 1);
      tc_fence_after();
      const int off = kHeadOff[k], cnt = kHeadOff[k + 1] - off;
      float z[48];
#pragma unroll
      for (int cb = 0; cb < 48; cb += 16) tmem_ld16(lane_addr + 96 + cb, z + cb);
      tmem_ld_wait();
      float m = -3.0e38f;
#pragma unroll
      for (int i = 0; i < 48; ++i) {
        const float x = z[i] + bh[i];
        z[i] = x >= 0.f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * (__expf(x) - 1.f);
        if (i < cnt) m = fmaxf(m, z[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d6_free);
      float sum = 0.f;
      float e[48];
#pragma unroll
      for (int i = 0; i < 48; ++i) {
        e[i] = i < cnt ? __expf(z[i] - m) : 0.f;
        sum += e[i];
      }
      const float inv = 1.f / sum;
      float* lg = logits + site * N_OUT + off;
#pragma unroll
      for (int i = 0; i < 48; ++i)
        if (i < cnt) lg[i] = z[i];
      if (site < n) {
        // packed [n][90] rows, or (split_rows > 0) four head-major arrays [split_rows][n_k] laid end to end - the layout
        // predict() hands back (clair/model.py:963), so the host never has to split rows
        float* pr = split_rows > 0 ? probs + (size_t)split_rows * off + site * cnt : probs + site * N_OUT + off;
#pragma unroll
        for (int i = 0; i < 48; ++i)
          if (i < cnt) pr[i] = e[i] * inv;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// A flat byte range viewed as a 2-D tensor [bytes/1024][256 x u32] for tensor-map TMA: a box of `box_rows` rows is
// box_rows KB of contiguous memory, landing contiguously in shared memory.  cuTensorMapEncodeTiled is fetched through
// the runtime (no link-time dependency on libcuda).
inline cudaError_t make_flat_map(CUtensorMap* tm, void* base, size_t bytes, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t st = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (st != cudaSuccess) return st;
    if (q != cudaDriverEntryPointSuccess || !p) return cudaErrorNotSupported;
    fn = (EncodeFn)p;
  }
  const cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 1024)};
  const cuuint64_t strides[1] = {1024};
  const cuuint32_t box[2] = {256, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

struct HostModel {
  const float* lstm_kernel[2][2];   // [layer][dir] TF layout [(Kx+128)][512], rows x first, columns gate-major i,c,f,o
  const float* lstm_bias[2][2];     // [512]
  const float* w3;                  // [256][33][30]  (L3/Unit_c/kernel)
  const float* b3;                  // [256][30]
  const float* W4;                  // [7680][192], row = o*256 + c
  const float* W5[4];               // L5_k/kernel [192][96]
  const float* b5[4];               // [96]
  const float* Whd[4];              // Prediction/*_logits/kernel [96][n_k]
  const float* bhd[4];              // [n_k]
};

struct Weights {
  __half* Wx2 = nullptr;                   // xproj_pair<32>, layer 2: [nb 4][q 2][hl][kc 32][128][8]
  float* bx2 = nullptr;                    // layer 2: [nb 4][256]
  uint8_t* Wx2s = nullptr;                 // lstm_seq_x2, layer 2: [dir][q][bp 2][st 8][hl][kc 4][128][8] fp16
  float* bx2s = nullptr;                   // layer 2: [dir][512] (unit*4+gate order, gate-scaled)
  CUtensorMap tmWx;                        // Wx2s as [rows][1 KB], box = 8 rows (one ring stage of B)
  __half* Whs[2] = {nullptr, nullptr};     // lstm_seq, per layer: [dir][q][hl][b 4][kc 16][64][8]
  __half* Wxf = nullptr;                   // lstm_seq<FUSE_X>, layer 1: [dir][q][hl][b 4][kc 6][64][8] (k 32,33 = bias hi,lo)
  uint8_t* l3l4 = nullptr;                 // [256] per-channel blobs (L3L4_BLOB_BYTES each)
  uint8_t* heads = nullptr;                // [4] per-head blobs (HEAD_BLOB_BYTES each)
  const float* b4 = nullptr;               // [192] (owned by the engine)
  float w4_unscale = 1.f;                  // 1 / (power-of-two scale folded into the W4 operand tiles)
};

struct Workspace {
  int64_t np_max = 0;
  __half* X48 = nullptr;     // [33*NT][hl][6][128][8]       layer-1 input tiles incl. the two bias columns
  __half* H1 = nullptr;      // [33*NT][hl][32][128][8]      LSTM1 output = A tiles of the layer-2 input projection
  float* Gx = nullptr;       // [33*NT][dir][unit][row][4]   layer-2 input projection (+bias, gate-scaled)
  __half* H2t = nullptr;     // [NT][256][hl][5][2][8][64]   LSTM2 output, MN-major 128B-swizzled per channel (t = 33..39 zero)
  __half* L4t = nullptr;     // [NT][hl][24][128][8]         L4 activations as the operand tile of heads_tc
  int sm_count = 148;
  CUtensorMap tmH1;          // H1 as [rows][1 KB], box = 4 rows (hi or lo half of one ring stage of A)
  long long* lf_trace = nullptr;   // CLAIRB_LF_TRACE=<file>: [256 channels][16] clock64 stamps of tile 0 of l3l4_fused
  long long* trace = nullptr; // CLAIRB_SX_TRACE=<file>: [33][64] clock64 stamps of one CTA pair of lstm_seq_x2 (dumped at destroy)
  int* x_lo_flag = nullptr;    // set by prep_tiles48 when an input value is not exact in fp16 (cleared before every chunk)
  long long* trace1 = nullptr; // CLAIRB_S1_TRACE=<file>: the same for the layer-1 launch of lstm_seq (CLAIRB_TRACE builds)
};

inline bool available() { return true; }

inline void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// gate column j (unit*4+gate) of a direction -> TF kernel column (gate*128+unit)
inline int tf_col(int j) { return (j & 3) * H + (j >> 2); }

template <typename Tp>
inline cudaError_t upload_vec(Tp** dptr, const std::vector<Tp>& h) {
  cudaError_t st = cudaMalloc((void**)dptr, h.size() * sizeof(Tp));
  if (st != cudaSuccess) return st;
  return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice);
}

inline cudaError_t build_weights(Weights& w, const HostModel& hm) {
  const int kx[2] = {F_IN, 2 * H};
  cudaError_t st;
#ifdef CLAIRB_CROSSCHECK
  // ---- layer-2 input projection (xproj_pair<32>): N-blocks of 256 gate columns, 128 rows per CTA ----
  {
    const int KC = 32;
    std::vector<__half> wx((size_t)4 * 2 * 2 * KC * KCH);
    std::vector<float> bx((size_t)4 * 256);
    for (int dir = 0; dir < 2; ++dir) {
      const float* K = hm.lstm_kernel[1][dir];
      const float* B = hm.lstm_bias[1][dir];
      for (int half = 0; half < 2; ++half) {
        const int nb = dir * 2 + half;
        for (int c = 0; c < 256; ++c) bx[(size_t)nb * 256 + c] = B[tf_col(half * 256 + c)] * GATE_SCALE[c & 3];
        for (int q = 0; q < 2; ++q)
          for (int row = 0; row < 128; ++row) {
            const int col = tf_col(half * 256 + q * 128 + row);
            const float gs = GATE_SCALE[row & 3];                 // fold the exp2 scaling of this gate into W
            for (int k = 0; k < kx[1]; ++k) {
              __half hi, lo;
              split_half(K[(size_t)k * G4 + col] * gs, hi, lo);
              const size_t base = (((size_t)nb * 2 + q) * 2) * KC * KCH + (size_t)(k / 8) * KCH + row * 8 + k % 8;
              wx[base] = hi;
              wx[base + (size_t)KC * KCH] = lo;
            }
          }
      }
    }
    if ((st = upload_vec(&w.Wx2, wx)) != cudaSuccess) return st;
    if ((st = upload_vec(&w.bx2, bx)) != cudaSuccess) return st;
  }
#endif
  // ---- layer-2 input projection streamed through lstm_seq_x2: per CTA q and block pair bp, row r = gate column
  //      (2bp+q)*128 + r; ring stages of K = 8*SX_KC: [hl][kc][128][8] ----
  {
    std::vector<__half> wx((size_t)2 * 2 * 2 * SX_NST * 2 * SX_KC * KCH);
    std::vector<float> bx((size_t)2 * 512);
    for (int dir = 0; dir < 2; ++dir) {
      const float* K = hm.lstm_kernel[1][dir];
      const float* B = hm.lstm_bias[1][dir];
      for (int j = 0; j < 512; ++j) bx[(size_t)dir * 512 + j] = B[tf_col(j)] * GATE_SCALE[j & 3];
      for (int q = 0; q < 2; ++q)
        for (int bp = 0; bp < 2; ++bp)
          for (int row = 0; row < 128; ++row) {
            const int col = tf_col((2 * bp + q) * 128 + row);
            const float gs = GATE_SCALE[row & 3];
            for (int k = 0; k < kx[1]; ++k) {
              __half hi, lo;
              split_half(K[(size_t)k * G4 + col] * gs, hi, lo);
              const int st = k / (8 * SX_KC), kc = (k % (8 * SX_KC)) / 8;
              const size_t base = ((((size_t)dir * 2 + q) * 2 + bp) * SX_NST + st) * (2 * SX_KC * KCH) + (size_t)kc * KCH + row * 8 + k % 8;
              wx[base] = hi;
              wx[base + SX_KC * KCH] = lo;
            }
          }
    }
    std::vector<uint8_t> raw((const uint8_t*)wx.data(), (const uint8_t*)wx.data() + wx.size() * 2);
    if ((st = upload_vec(&w.Wx2s, raw)) != cudaSuccess) return st;
    if ((st = upload_vec(&w.bx2s, bx)) != cudaSuccess) return st;
    if ((st = make_flat_map(&w.tmWx, w.Wx2s, raw.size(), 4 * SX_KC)) != cudaSuccess) return st;
  }
  // ---- lstm_seq layouts: gate blocks of 128 columns, 64 rows per CTA ----
  for (int l = 0; l < 2; ++l) {
    std::vector<__half> whs((size_t)2 * 2 * 2 * 4 * 16 * 512);
    std::vector<__half> wxf(l == 0 ? (size_t)2 * 2 * 2 * 4 * 6 * 512 : 0, __float2half_rn(0.f));
    for (int dir = 0; dir < 2; ++dir) {
      const float* K = hm.lstm_kernel[l][dir];
      const float* B = hm.lstm_bias[l][dir];
      for (int q = 0; q < 2; ++q)
        for (int b = 0; b < 4; ++b)
          for (int row = 0; row < 64; ++row) {
            const int col = tf_col(b * 128 + q * 64 + row);
            const float gs = GATE_SCALE[row & 3];
            const size_t cta = ((size_t)dir * 2 + q);
            for (int k = 0; k < H; ++k) {
              __half hi, lo;
              split_half(K[(size_t)(kx[l] + k) * G4 + col] * gs, hi, lo);
              const size_t off = cta * (2 * 4 * 16 * 512) + ((size_t)b * 16 + k / 8) * 512 + row * 8 + k % 8;
              whs[off] = hi;
              whs[off + (size_t)4 * 16 * 512] = lo;
            }
            if (l == 0) {
              for (int k = 0; k < F_IN + 2; ++k) {
                __half hi, lo;
                if (k < F_IN) {
                  split_half(K[(size_t)k * G4 + col] * gs, hi, lo);
                } else {
                  // bias rides on two constant-one input columns: row 32 = fp16(b), row 33 = fp16(b - fp16(b))
                  __half bh, bl;
                  split_half(B[col] * gs, bh, bl);
                  hi = k == F_IN ? bh : bl;
                  lo = __float2half_rn(0.f);
                }
                const size_t off = cta * (2 * 4 * 6 * 512) + ((size_t)b * 6 + k / 8) * 512 + row * 8 + k % 8;
                wxf[off] = hi;
                wxf[off + (size_t)4 * 6 * 512] = lo;
              }
            }
          }
    }
    if ((st = upload_vec(&w.Whs[l], whs)) != cudaSuccess) return st;
    if (l == 0 && (st = upload_vec(&w.Wxf, wxf)) != cudaSuccess) return st;
  }
  // ---- slice-dense + L4 blobs ----
  {
    // W4's entries are small (fan-in 7680: std ~0.013), so the low halves of an unscaled hi/lo split land in the fp16
    // subnormals (steps of 6e-8, i.e. ~18 bits of the weight instead of 22).  Scale by a power of two (exact) so that
    // the largest |w| sits near 2^10; the L4 epilogue multiplies the accumulator by the inverse.  (Not what limits L4
    // today: its 2e-5 comes from the tensor core's fp32 accumulation over 1536 chained MMAs per tile - summing 64-channel
    // segments in fp32 registers brought it to 5.7e-6 but cost 0.05-0.12 ms per chunk, tools/patches/.)
    float w4_max = 0.f;
    for (size_t i = 0; i < (size_t)L3_K * L4_UNITS; ++i) w4_max = fmaxf(w4_max, fabsf(hm.W4[i]));
    int e4 = 0;
    if (w4_max > 0.f) frexpf(w4_max, &e4);
    const float w4_scale = ldexpf(1.f, 10 - e4);
    w.w4_unscale = ldexpf(1.f, e4 - 10);
    std::vector<uint8_t> blob((size_t)2 * H * L3L4_BLOB_BYTES, 0);
    for (int c = 0; c < 2 * H; ++c) {
      uint8_t* b = blob.data() + (size_t)c * L3L4_BLOB_BYTES;
      __half* w3 = (__half*)b;
      __half* w4hi = (__half*)(b + 2 * L3W_BYTES);
      __half* w4lo = (__half*)(b + 2 * L3W_BYTES + L4W_BYTES);
      float* bb = (float*)(b + 2 * L3W_BYTES + 2 * L4W_BYTES);
      // W3 as ONE K-major B tile of 64 rows per k-chunk: rows 0..31 = hi[o], rows 32..63 = lo[o]
      for (int t = 0; t < T_STEPS; ++t)
        for (int o = 0; o < L3_UNITS; ++o) {
          const size_t idx = (size_t)(t / 8) * 64 * 8 + o * 8 + t % 8;
          split_half(hm.w3[((size_t)c * T_STEPS + t) * L3_UNITS + o], w3[idx], w3[idx + 32 * 8]);
        }
      for (int o = 0; o < L3_UNITS; ++o) {
        bb[o] = hm.b3[(size_t)c * L3_UNITS + o];
        for (int n = 0; n < L4_UNITS; ++n) {
          const size_t idx = (size_t)(o / 8) * 192 * 8 + n * 8 + o % 8;
          split_half(hm.W4[((size_t)o * 2 * H + c) * L4_UNITS + n] * w4_scale, w4hi[idx], w4lo[idx]);
        }
      }
    }
    if ((st = upload_vec(&w.l3l4, blob)) != cudaSuccess) return st;
  }
  // ---- L5 + head blobs ----
  {
    const int head_n[4] = {21, 3, 33, 33};
    std::vector<uint8_t> blob((size_t)4 * HEAD_BLOB_BYTES, 0);
    for (int k = 0; k < 4; ++k) {
      uint8_t* b = blob.data() + (size_t)k * HEAD_BLOB_BYTES;
      for (int hf = 0; hf < 2; ++hf) {
        __half* hi = (__half*)(b + (size_t)hf * HD_W5_BYTES);
        __half* lo = hi + 24 * 48 * 8;
        for (int row = 0; row < 48; ++row)
          for (int kk = 0; kk < L4_UNITS; ++kk) {
            const size_t idx = (size_t)(kk / 8) * 48 * 8 + row * 8 + kk % 8;
            split_half(hm.W5[k][(size_t)kk * L5_UNITS + hf * 48 + row], hi[idx], lo[idx]);
          }
      }
      __half* hi = (__half*)(b + 2 * HD_W5_BYTES);
      __half* lo = hi + 12 * 48 * 8;
      for (int o = 0; o < head_n[k]; ++o)
        for (int j = 0; j < L5_UNITS; ++j) {
          const size_t idx = (size_t)(j / 8) * 48 * 8 + o * 8 + j % 8;
          split_half(hm.Whd[k][(size_t)j * head_n[k] + o], hi[idx], lo[idx]);
        }
      float* bb = (float*)(b + 2 * HD_W5_BYTES + HD_WH_BYTES);
      for (int j = 0; j < L5_UNITS; ++j) bb[j] = hm.b5[k][j];
      for (int o = 0; o < head_n[k]; ++o) bb[96 + o] = hm.bhd[k][o];
    }
    if ((st = upload_vec(&w.heads, blob)) != cudaSuccess) return st;
  }
  return cudaSuccess;
}

inline void free_weights(Weights& w) {
  cudaFree(w.heads); w.heads = nullptr;
  cudaFree(w.l3l4); cudaFree(w.Wxf); cudaFree(w.Whs[0]); cudaFree(w.Whs[1]); cudaFree(w.Wx2); cudaFree(w.bx2);
  cudaFree(w.Wx2s); cudaFree(w.bx2s); w.Wx2s = nullptr; w.bx2s = nullptr;
  w.l3l4 = nullptr; w.Wxf = nullptr; w.Whs[0] = w.Whs[1] = nullptr; w.Wx2 = nullptr; w.bx2 = nullptr;
}

inline cudaError_t alloc_workspace(Workspace& ws, int64_t np_max, int device, bool need_gx) {
  ws.np_max = np_max;
  const size_t NT = (size_t)np_max / 128;
  cudaError_t st;
  if ((st = cudaMalloc((void**)&ws.X48, (size_t)T_STEPS * NT * X48_TILE_HALVES * 2)) != cudaSuccess) return st;
  if ((st = cudaMalloc((void**)&ws.H1, (size_t)T_STEPS * NT * 2 * 32 * KCH * 2)) != cudaSuccess) return st;
  if ((st = make_flat_map(&ws.tmH1, ws.H1, (size_t)T_STEPS * NT * 2 * 32 * KCH * 2, 2 * SX_KC)) != cudaSuccess) return st;
  // Gx (135 KB per site: 2.6 GB for a chunk) only exists on the two-kernel cross-check path of layer 2
  if (need_gx && (st = cudaMalloc((void**)&ws.Gx, (size_t)T_STEPS * NT * GX_TILE_FLOATS * 4)) != cudaSuccess) return st;
  if ((st = cudaMalloc((void**)&ws.H2t, NT * 2 * H * (size_t)L3A_BYTES)) != cudaSuccess) return st;
  // time steps 33..39 of the last time group are never written by lstm_seq and must read as zeros
  if ((st = cudaMemset(ws.H2t, 0, NT * 2 * H * (size_t)L3A_BYTES)) != cudaSuccess) return st;
  if ((st = cudaMalloc((void**)&ws.L4t, NT * (size_t)HD_A4_BYTES)) != cudaSuccess) return st;
  if ((st = cudaMalloc((void**)&ws.x_lo_flag, sizeof(int))) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(heads_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heads_smem_bytes())) != cudaSuccess) return st;
  cudaDeviceGetAttribute(&ws.sm_count, cudaDevAttrMultiProcessorCount, device);
  if (getenv("CLAIRB_LF_TRACE")) {
    if ((st = cudaMalloc((void**)&ws.lf_trace, 256 * 16 * sizeof(long long))) != cudaSuccess) return st;
    cudaMemset(ws.lf_trace, 0, 256 * 16 * sizeof(long long));
  }
  if (getenv("CLAIRB_SX_TRACE")) {
    if ((st = cudaMalloc((void**)&ws.trace, T_STEPS * 64 * sizeof(long long))) != cudaSuccess) return st;
    cudaMemset(ws.trace, 0, T_STEPS * 64 * sizeof(long long));
  }
  if (getenv("CLAIRB_S1_TRACE")) {
    if ((st = cudaMalloc((void**)&ws.trace1, T_STEPS * 64 * sizeof(long long))) != cudaSuccess) return st;
    cudaMemset(ws.trace1, 0, T_STEPS * 64 * sizeof(long long));
  }
#ifdef CLAIRB_CROSSCHECK
  if ((st = cudaFuncSetAttribute(xproj_pair<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xproj_smem_bytes<32>())) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(lstm_seq<false, 1, SEQ_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes<false>())) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(lstm_seq<false, 2, SEQ_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes<false>())) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(lstm_seq_x2<1, SX_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seqx_smem_bytes())) != cudaSuccess) return st;
#endif
  if ((st = cudaFuncSetAttribute(lstm_seq<true, 0, SEQ1_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes<true>())) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(l3l4_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l3l4_smem_bytes())) != cudaSuccess) return st;
  if ((st = cudaFuncSetAttribute(lstm_seq_x2<2, SX_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seqx_smem_bytes())) != cudaSuccess) return st;
  return cudaSuccess;
}

inline void free_workspace(Workspace& ws) {
  if (ws.lf_trace) {
    std::vector<long long> h(256 * 16);
    if (cudaMemcpy(h.data(), ws.lf_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(getenv("CLAIRB_LF_TRACE") ? getenv("CLAIRB_LF_TRACE") : "/dev/null", "w")) {
        for (int c = 0; c < 256; ++c)
          for (int e = 0; e < 16; ++e) fprintf(f, "%lld%c", h[c * 16 + e], e == 15 ? '\n' : ' ');
        fclose(f);
      }
    }
    cudaFree(ws.lf_trace);
    ws.lf_trace = nullptr;
  }
  if (ws.trace1) {
    std::vector<long long> h(T_STEPS * 64);
    if (cudaMemcpy(h.data(), ws.trace1, h.size() * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(getenv("CLAIRB_S1_TRACE") ? getenv("CLAIRB_S1_TRACE") : "/dev/null", "w")) {
        for (int s = 0; s < T_STEPS; ++s)
          for (int e = 0; e < 64; ++e) fprintf(f, "%lld%c", h[s * 64 + e], e == 63 ? '\n' : ' ');
        fclose(f);
      }
    }
    cudaFree(ws.trace1);
    ws.trace1 = nullptr;
  }
  if (ws.trace) {
    std::vector<long long> h(T_STEPS * 64);
    if (cudaMemcpy(h.data(), ws.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(getenv("CLAIRB_SX_TRACE") ? getenv("CLAIRB_SX_TRACE") : "/dev/null", "w")) {
        for (int s = 0; s < T_STEPS; ++s) {
          for (int e = 0; e < 64; ++e) fprintf(f, "%lld%c", h[s * 64 + e], e == 63 ? '\n' : ' ');
        }
        fclose(f);
      }
    }
    cudaFree(ws.trace);
    ws.trace = nullptr;
  }
  cudaFree(ws.X48); cudaFree(ws.Gx); cudaFree(ws.H1); cudaFree(ws.H2t); cudaFree(ws.L4t); cudaFree(ws.x_lo_flag);
  ws.x_lo_flag = nullptr;
  ws.L4t = nullptr; ws.X48 = nullptr; ws.Gx = nullptr; ws.H1 = nullptr; ws.H2t = nullptr;
}

// Both BiLSTM layers (+ slice-dense and L4 when fuse_tail) for np padded sites (np % 256 == 0).
//   fuse_tail = false: x -> h2 planes [33*256][np] fp32 (the CUDA-core slice-dense / L4 follow; parity cross-check)
//   fuse_tail = true : x -> probabilities [n][90] (+ logits [np][90]) through lstm_seq<.,2> + l3l4_fused + heads_tc
// `hook(id, begin)` brackets every launch for the per-kernel event timing
// (id: 0 prep_tiles, 1 lstm_seq1, 2 xproj2, 3 lstm_seq2, 4 l3l4_fused, 5 heads_tc, 6 lstm_seq_x2).
template <typename Hook>
inline cudaError_t forward_lstm(const Weights& w, Workspace& ws, const void* x_dev, int dtype_is_i16, int64_t n, int64_t np,
                                float* h2_planes, float* l4T, float* probs, float* logits, bool fuse_tail, bool l2_stream,
                                cudaStream_t st, int* launches, Hook&& hook, int64_t split_rows = 0) {
  const int NT = (int)(np / 128);
  const int num_row_pairs = T_STEPS * NT / 2;
  // persistent input-projection grid: whole groups of 4 CTA pairs (one pair per N-block), one pair per 2 SMs
  int ncl = (ws.sm_count / 2) / 4 * 4;
  if (ncl > 4 * num_row_pairs) ncl = 4 * num_row_pairs;
  if (ncl < 4) ncl = 4;
  static const int gx_pf = getenv("CLAIRB_GX_PF") ? atoi(getenv("CLAIRB_GX_PF")) : 0;   // Gx L2 prefetch distance (steps)
  static const int l3_pf = getenv("CLAIRB_L3_PF") ? atoi(getenv("CLAIRB_L3_PF")) : 8;   // l3l4 L2 prefetch distance (channels)
  static const int xp_dbg = getenv("CLAIRB_XP_DBG") ? atoi(getenv("CLAIRB_XP_DBG")) : 0;   // timing experiments only
  static const int sx_dbg = getenv("CLAIRB_SX_DBG") ? atoi(getenv("CLAIRB_SX_DBG")) : 0;   // timing experiments only
  static const int s1_dbg = getenv("CLAIRB_SEQ1_DBG") ? atoi(getenv("CLAIRB_SEQ1_DBG")) : 0; // timing experiments only
  dim3 gprep((unsigned)NT, T_STEPS);
  dim3 grec((unsigned)NT, 2);
  cudaMemsetAsync(ws.x_lo_flag, 0, sizeof(int), st);
  hook(0, true);
  if (dtype_is_i16) prep_tiles48<int16_t><<<gprep, 128, 0, st>>>((const int16_t*)x_dev, ws.X48, n, NT, ws.x_lo_flag);
  else prep_tiles48<float><<<gprep, 128, 0, st>>>((const float*)x_dev, ws.X48, n, NT, ws.x_lo_flag);
  hook(0, false);
  hook(1, true);   // layer 1: input projection fused into the recurrent kernel (no Gx round trip)
  lstm_seq<true, 0, SEQ1_G><<<grec, 32 * (2 + 4 * SEQ1_G), seq_smem_bytes<true>(), st>>>(w.Whs[0], w.Wxf, ws.X48, (const float*)ws.trace1, ws.H1, NT, np, s1_dbg, ws.x_lo_flag);
  hook(1, false);
  // layer 2: input projection streamed through the recurrent kernel (default), or the two-kernel path
  // xproj_pair -> Gx -> lstm_seq (CLAIRB_L2_STREAM=0: on-device cross-check)
  if (l2_stream) {
    hook(6, true);
    if (fuse_tail) lstm_seq_x2<2, SX_G><<<grec, SX_THREADS, seqx_smem_bytes(), st>>>(w.Whs[1], ws.tmH1, w.tmWx, w.bx2s, ws.H1, ws.H2t, NT, np, sx_dbg, ws.trace);
#ifdef CLAIRB_CROSSCHECK
    else lstm_seq_x2<1, SX_G><<<grec, SX_THREADS, seqx_smem_bytes(), st>>>(w.Whs[1], ws.tmH1, w.tmWx, w.bx2s, ws.H1, h2_planes, NT, np, sx_dbg, ws.trace);
#endif
    hook(6, false);
    *launches += 3;
  }
#ifdef CLAIRB_CROSSCHECK
  else {
    hook(2, true);
    xproj_pair<32><<<2 * ncl, XP_THREADS, xproj_smem_bytes<32>(), st>>>(ws.H1, w.Wx2, w.bx2, ws.Gx, num_row_pairs, xp_dbg);
    hook(2, false);
    hook(3, true);
    if (fuse_tail) lstm_seq<false, 2, SEQ_G><<<grec, SEQ_THREADS, seq_smem_bytes<false>(), st>>>(w.Whs[1], nullptr, nullptr, ws.Gx, ws.H2t, NT, np, gx_pf, nullptr);
    else lstm_seq<false, 1, SEQ_G><<<grec, SEQ_THREADS, seq_smem_bytes<false>(), st>>>(w.Whs[1], nullptr, nullptr, ws.Gx, h2_planes, NT, np, gx_pf, nullptr);
    hook(3, false);
    *launches += 4;
  }
#endif
  if (fuse_tail) {
    hook(4, true);
    l3l4_fused<<<(unsigned)NT, LF_THREADS, l3l4_smem_bytes(), st>>>(ws.H2t, w.l3l4, w.b4, l4T, ws.L4t, np, nullptr, l3_pf, ws.lf_trace, w.w4_unscale);
    hook(4, false);
    hook(5, true);
    heads_tc<<<(unsigned)NT, HD_THREADS, heads_smem_bytes(), st>>>(ws.L4t, w.heads, probs, logits, n, split_rows);
    hook(5, false);
    *launches += 2;
  }
  return cudaGetLastError();
}

// parity hook: LSTM1 output [33][n][256] fp32 rebuilt from the hi/lo operand tiles
inline cudaError_t get_lstm1(const Workspace& ws, int64_t n, int64_t np, float* out_host) {
  const size_t NT = (size_t)np / 128;
  std::vector<__half> buf((size_t)T_STEPS * NT * 2 * 32 * KCH);
  cudaError_t st = cudaMemcpy(buf.data(), ws.H1, buf.size() * 2, cudaMemcpyDeviceToHost);
  if (st != cudaSuccess) return st;
  for (int t = 0; t < T_STEPS; ++t)
    for (int64_t s = 0; s < n; ++s) {
      const size_t base = ((size_t)t * NT + (size_t)(s / 128)) * (2 * 32 * KCH);
      const int r = (int)(s % 128);
      for (int f = 0; f < 2 * H; ++f) {
        const size_t off = (size_t)(f / 8) * KCH + r * 8 + f % 8;
        out_host[((size_t)t * n + s) * 2 * H + f] =
            __half2float(buf[base + off]) + __half2float(buf[base + (size_t)32 * KCH + off]);
      }
    }
  return cudaSuccess;
}

// parity hook: re-run the fused slice-dense on the retained H2t with the L3 activations written out as planes
inline cudaError_t dump_l3(const Weights& w, const Workspace& ws, int64_t np, float* l4T, float* l3_planes) {
  l3l4_fused<<<(unsigned)(np / 128), LF_THREADS, l3l4_smem_bytes(), 0>>>(ws.H2t, w.l3l4, w.b4, l4T, ws.L4t, np, l3_planes, 0, nullptr, w.w4_unscale);
  cudaError_t st = cudaGetLastError();
  return st != cudaSuccess ? st : cudaDeviceSynchronize();
}

// parity hook: LSTM2 output [33][n][256] fp32 rebuilt from the MN-major swizzled hi/lo tiles
inline cudaError_t get_lstm2(const Workspace& ws, int64_t n, int64_t np, float* out_host) {
  const size_t NT = (size_t)np / 128;
  std::vector<__half> buf(NT * 2 * H * (size_t)L3A_HALVES);
  cudaError_t st = cudaMemcpy(buf.data(), ws.H2t, buf.size() * 2, cudaMemcpyDeviceToHost);
  if (st != cudaSuccess) return st;
  for (int t = 0; t < T_STEPS; ++t)
    for (int64_t s = 0; s < n; ++s) {
      const int r = (int)(s % 128), rg = r >> 3;
      for (int f = 0; f < 2 * H; ++f) {
        const size_t base = ((size_t)(s / 128) * 2 * H + f) * L3A_HALVES;
        const size_t off = (size_t)(t >> 3) * 1024 + (rg >> 3) * 512 + (t & 7) * 64 + (((rg & 7) ^ (t & 7)) << 3) + (r & 7);
        out_host[((size_t)t * n + s) * 2 * H + f] =
            __half2float(buf[base + off]) + __half2float(buf[base + L3A_HALVES / 2 + off]);
      }
    }
  return cudaSuccess;
}

}  // namespace tc
}  // namespace clairb
