// First-choice variant decision per site: which of the reference's ~1.2 k outcome products wins.
//
// Reference: clair/call_var.py
//   possible_outcome_probabilites_from :589-690 and the *_tuples_from helpers :344-424  (the outcome lists)
//   output_from, first pass of its loop :732-760  (maximum, then the is_* membership tests in elif order; list.index)
//   homo_SNP_bases_from / hetero_SNP_bases_from :60-67, read depth :1021-1024
// The reference evaluates every product in numpy.float32 (float32 * float32 stays float32) and associates left to
// right, so the products here are __fmul_rn in exactly that order and the result is bit-identical: the winner is the
// entry with the largest value and, among equal values, the smallest (category rank, index in the reference's list).
//
// One warp per site: the 90 probabilities sit in shared memory, every lane walks a 32-strided share of the 1179
// entries, then a shuffle reduction picks the winner.  HBM traffic is 360 B (+256 B: rows 16 and 17 of x) in and
// 32 B out per site; the arithmetic (about 3.6 k multiplies per site) is negligible next to it.
//
// Behind the winner, lane 0 also evaluates what output_with (clair/call_var.py:1002-1197) derives from it without any
// string in hand: the quality score (quality_score_from :568-586: float32 product p of the gt21 and genotype probabilities
// of the call, then float64 as the reference's pinned numpy 1.18 computes it - `1.0 - p` with a Python float promotes a
// float32 scalar to float64 there) and the supporting-read count (:1100-1151: sums over rows 16 / 17 of the tensor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace clairb {
namespace decide {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int REC_WORDS = 8;          // category, len1, len2, aux, max probability (f32 bits), read depth (f32 bits),
                                      // quality score, supporting reads (f32 bits)
// -10 * log(e, 10) as CPython evaluates it (clair/call_var.py:581): math.log(math.e, 10) = 0.4342944819032518
constexpr double QUAL_SCALE = -4.342944819032518;

struct Best {
  float v;
  uint32_t key;                       // category rank << 10 | index in the reference's list
};
__device__ __forceinline__ void consider(Best& b, float v, uint32_t key) {
  if (v > b.v || (v == b.v && key < b.key)) { b.v = v; b.key = key; }
}

template <typename TIn>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
decide_sites(const float* __restrict__ probs, const uint8_t* __restrict__ ref_base, const TIn* __restrict__ x,
             int32_t* __restrict__ rec, int64_t n, int64_t split_rows) {
  __shared__ float sp[WARPS_PER_BLOCK][96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t site = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  if (site >= n) return;
  float* p = sp[warp];
  // packed [n][90] rows, or (split_rows > 0) the four head-major arrays [split_rows][n_k] the heads kernel writes for
  // Clair.predict
  for (int i = lane; i < N_OUT; i += 32) {
    const int k = i < 21 ? 0 : i < 24 ? 1 : i < 57 ? 2 : 3;
    const int off = kHeadOff[k], cnt = kHeadOff[k + 1] - off;
    p[i] = split_rows > 0 ? probs[(size_t)split_rows * off + site * cnt + (i - off)] : probs[site * N_OUT + i];
  }
  __syncwarp();
  const float* gt21 = p;
  const float* vl1 = p + 24 + 16;     // index by signed length -16..16 (VariantLength.index_offset = 16)
  const float* vl2 = p + 57 + 16;
  const float homo_ref = p[21], homo_var = p[22], het_var = p[23];
  const int rb = ref_base[site] & 3;
  const int ref_gt21 = rb == 0 ? 0 : rb == 1 ? 4 : rb == 2 ? 7 : 9;           // AA CC GG TT
  const float vl0 = __fmul_rn(vl1[0], vl2[0]);                                  // :600-603

  Best b{-1.f, 0xffffffffu};
  // 0 reference (:606-608), 1 homo SNP (:610-612), 2 hetero SNP (:613-615)
  if (lane == 0) consider(b, __fmul_rn(__fmul_rn(vl0, homo_ref), gt21[ref_gt21]), 0u << 10);
  if (lane < 4) {
    const int g = lane == 0 ? 0 : lane == 1 ? 4 : lane == 2 ? 7 : 9;
    consider(b, __fmul_rn(__fmul_rn(vl0, homo_var), gt21[g]), (1u << 10) | lane);
  }
  if (lane < 6) {
    const int g = lane == 0 ? 1 : lane == 1 ? 2 : lane == 2 ? 3 : lane == 3 ? 5 : lane == 4 ? 6 : 8;
    consider(b, __fmul_rn(__fmul_rn(vl0, het_var), gt21[g]), (2u << 10) | lane);
  }
  // 3 homo Ins (:344-349, :617-621), 6 homo Del (:377-382, :645-649)
  if (lane < 16) {
    const int i = lane + 1;
    consider(b, __fmul_rn(__fmul_rn(vl1[i], vl2[i]), __fmul_rn(homo_var, gt21[15])), (3u << 10) | lane);
    consider(b, __fmul_rn(__fmul_rn(vl1[-i], vl2[-i]), __fmul_rn(homo_var, gt21[10])), (6u << 10) | lane);
  }
  // 4 hetero ACGT+Ins (:352-361, :630-639), 7 hetero ACGT+Del (:385-394, :658-667): index = (i-1)*4 + base
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = lane + 32 * k, i = (idx >> 2) + 1, base = idx & 3;
    const float pi = fmaxf(__fmul_rn(vl1[0], vl2[i]), __fmul_rn(vl1[i], vl2[0]));
    consider(b, __fmul_rn(__fmul_rn(pi, gt21[16 + base]), het_var), (4u << 10) | idx);
    const float pd = fmaxf(__fmul_rn(vl1[0], vl2[-i]), __fmul_rn(vl1[-i], vl2[0]));
    consider(b, __fmul_rn(__fmul_rn(pd, gt21[11 + base]), het_var), (7u << 10) | idx);
  }
  // 5 hetero InsIns (:364-374, :622-629), 8 hetero DelDel (:397-408, :650-657), 9 InsDel (:411-424, :670-678)
  const float e_insins = __fmul_rn(het_var, gt21[15]), e_deldel = __fmul_rn(het_var, gt21[10]);
  const float e_insdel = __fmul_rn(het_var, gt21[20]);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = lane + 32 * k, i = (idx >> 4) + 1, j = (idx & 15) + 1;
    consider(b, __fmul_rn(__fmul_rn(vl1[i], vl2[j]), e_insins), (5u << 10) | idx);
    if (i != j) {
      const int didx = (i - 1) * 15 + (j - 1) - (j > i ? 1 : 0);
      consider(b, __fmul_rn(__fmul_rn(vl1[-i], vl2[-j]), e_deldel), (8u << 10) | didx);
    }
    consider(b, __fmul_rn(__fmul_rn(vl1[i], vl2[-j]), e_insdel), (9u << 10) | (2 * idx));
    consider(b, __fmul_rn(__fmul_rn(vl1[-i], vl2[j]), e_insdel), (9u << 10) | (2 * idx + 1));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, off);
    const uint32_t ok = __shfl_xor_sync(0xffffffffu, b.key, off);
    consider(b, ov, ok);
  }
  if (lane == 0) {
    const int cat = b.key >> 10, idx = b.key & 1023;
    int len1 = 0, len2 = 0, aux = 0;
    if (cat == 0) {
      aux = ref_gt21;
    } else if (cat == 1) {                                   // label = arg-max over the subset, first maximum (:60-62)
      const int g[4] = {0, 4, 7, 9};
      int best = 0;
      for (int q = 1; q < 4; ++q) if (gt21[g[q]] > gt21[g[best]]) best = q;
      aux = g[best];
    } else if (cat == 2) {                                   // (:65-67)
      const int g[6] = {1, 2, 3, 5, 6, 8};
      int best = 0;
      for (int q = 1; q < 6; ++q) if (gt21[g[q]] > gt21[g[best]]) best = q;
      aux = g[best];
    } else if (cat == 3 || cat == 6) {
      len1 = idx + 1;
    } else if (cat == 4 || cat == 7) {
      len1 = (idx >> 2) + 1;
      aux = idx & 3;
    } else if (cat == 5) {
      const int i = (idx >> 4) + 1, j = (idx & 15) + 1;
      len1 = i < j ? i : j;
      len2 = i < j ? j : i;
    } else if (cat == 8) {
      const int i = idx / 15 + 1;
      int j = idx % 15 + 1;
      if (j >= i) ++j;
      len1 = i < j ? i : j;
      len2 = i < j ? j : i;
    } else {                                                 // InsDel tuples: (j, i) then (i, j) (:415-423)
      const int pair = idx >> 1, i = (pair >> 4) + 1, j = (pair & 15) + 1;
      len1 = (idx & 1) ? i : j;
      len2 = (idx & 1) ? j : i;
    }
    // read depth = sum(x[16,:,delete] + x[16,:,reference]), Python's left-to-right sum starting from 0 (:1021-1024)
    float depth = 0.f;
    if (x != nullptr) {
      const TIn* row = x + site * SITE_ELEMS + 16 * F_IN;
      for (int r = 0; r < 8; ++r) depth = __fadd_rn(depth, __fadd_rn((float)row[r * 4 + 2], (float)row[r * 4 + 0]));
    }
    // quality score of the call (:568-586).  The gt21 label / genotype the reference derives from its REF, ALT and genotype
    // strings (clair/task/gt21.py:64-108, genotype.py:20-33) follow from the category alone: reference -> (ref ref, 0/0);
    // SNPs -> (the label, 1/1 or 0/1 | 1/2 -> hetero); insertions -> InsIns / <base>Ins; deletions -> DelDel / <base>Del;
    // Ins+Del -> InsDel; every two-allele call scores with the hetero genotype (genotype_enum_for_task).
    const int q_gt21 = cat <= 2 ? aux : cat == 3 || cat == 5 ? 15 : cat == 4 ? 16 + aux : cat == 6 || cat == 8 ? 10 : cat == 7 ? 11 + aux : 20;
    const int q_geno = cat == 0 ? 0 : (cat == 1 || cat == 3 || cat == 6) ? 1 : 2;
    const double pq = (double)__fmul_rn(gt21[q_gt21], p[21 + q_geno]);
    double tq = QUAL_SCALE * log(((1.0 - pq) + 1e-300) / (pq + 1e-300)) + 16.0;
    tq = tq > 0.0 ? tq : 0.0;
    const int quality = (int)rint(tq * tq);                  // Python 3 round(): half to even
    // supporting reads (:1100-1151), sums in the reference's order (integer counts: exact in float32 either way)
    float support = 0.f;
    if (x != nullptr) {
      const TIn* c16 = x + site * SITE_ELEMS + 16 * F_IN;
      const TIn* c17 = c16 + F_IN;
      auto snp_support = [&](int base) {                     // SNP + reference channel of both strands of one base
        return __fadd_rn(__fadd_rn(__fadd_rn((float)c16[base * 4 + 3], (float)c16[(base + 4) * 4 + 3]), (float)c16[base * 4 + 0]),
                         (float)c16[(base + 4) * 4 + 0]);
      };
      auto column = [&](int channel) {
        float t = 0.f;
        for (int r = 0; r < 8; ++r) t = __fadd_rn(t, (float)c17[r * 4 + channel]);
        return t;
      };
      // label bases of a gt21 pair label 0..9 (AA AC AG AT CC CG CT GG GT TT)
      const int l1 = aux < 4 ? 0 : aux < 7 ? 1 : aux < 9 ? 2 : 3, l2 = aux < 4 ? aux : aux < 7 ? aux - 3 : aux < 9 ? aux - 5 : 3;
      if (cat == 0) {
        support = __fadd_rn((float)c16[rb * 4 + 0], (float)c16[(rb + 4) * 4 + 0]);
      } else if (cat == 1) {
        support = snp_support(l1);
      } else if (cat == 2) {
        if (l1 != rb && l2 != rb) support = __fadd_rn(snp_support(l1), snp_support(l2));
        else support = snp_support(l1 != rb ? l1 : l2);
      } else if (cat == 3 || cat == 5) {
        support = __fsub_rn(column(1), column(3));
      } else if (cat == 4) {
        support = __fadd_rn(__fsub_rn(column(1), column(3)), aux != rb ? snp_support(aux) : 0.f);
      } else if (cat == 6 || cat == 8) {
        support = column(2);
      } else if (cat == 7) {
        support = __fadd_rn(column(2), aux != rb ? snp_support(aux) : 0.f);
      } else {
        support = __fsub_rn(__fadd_rn(column(1), column(2)), column(3));
      }
    }
    int32_t* o = rec + site * REC_WORDS;
    o[0] = cat; o[1] = len1; o[2] = len2; o[3] = aux;
    o[4] = __float_as_int(b.v);
    o[5] = __float_as_int(depth);
    o[6] = quality;
    o[7] = __float_as_int(support);
  }
}

template <typename TIn>
inline cudaError_t launch(const float* probs, const uint8_t* ref_base, const TIn* x, int32_t* rec, int64_t n, cudaStream_t st,
                          int64_t split_rows = 0) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  decide_sites<TIn><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(probs, ref_base, x, rec, n, split_rows);
  return cudaGetLastError();
}

}  // namespace decide
}  // namespace clairb
