// CreateTensor on the device: alignments -> 33x8x4 pileup counts, one thread block per
// candidate site.  Replaces the record lists + generate_tensor of the reference
// (dataPrepScripts/CreateTensor.py:29-65 and the CIGAR walk :283-366).
//
// The reference walks every read once and appends a record to every candidate window the read is currently inside
// ("active set", :298-320), then folds each candidate's records into counts.  Seen from one candidate c (1-based centre;
// its window is the 33 zero-based reference positions c-17 .. c+15) and one read, that bookkeeping reduces to:
//   q = the first reference position at which the read opens the window
//       = max(POS, c-17) when the read covers a position in [c-17, c+16]      (left edge considered, the default, :88-94)
//       = c-17           when the read covers c-17, otherwise the read is ignored   (--stop_consider_left_edge, :96)
//   aligned base (M,=,X) at p in the window: counted                           (a window is opened before the append, :298-314)
//   deleted base (D) at p:                   counted iff p > q                 (opened after the append, :336-357)
//   inserted bases (I) before reference position p: counted iff p > q          (no open on an insertion, :322-334)
// Integer work, no reuse across sites: every block reads the ops / bases of the reads that overlap its window (a few KB,
// L2-resident for neighbouring candidates), accumulates the 1056 counters in shared memory and writes one coalesced
// 2,112-byte int16 row, optionally with channel 0 already subtracted from channels 1..3 (the transform
// utils.tensor_generator_from applies before the network: clair/utils.py:96-98).  HBM-bound by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clairb {
namespace ct {

constexpr int FLANK = 16;                 // shared/param.py:9
constexpr int N_POS = 2 * FLANK + 1;      // 33
constexpr int CELLS = N_POS * 8;          // (position, row) cells of 4 channels
constexpr int ELEMS = CELLS * 4;          // 1056
constexpr int OP_M = 0, OP_I = 1, OP_D = 2;   // ops kept by the host encoder; S only advances the query offset,
                                              // N / H / P advance nothing in the reference (:283-366) and are dropped
constexpr int THREADS = 128;

// shared/utils.py:24-27 (IUPAC code -> A,C,G,T row), lower case folded as SEQ.upper() / sequence.upper() do
// (CreateTensor.py:149,261); 255 = not a base, the record is skipped (:37-38)
// Branch-free: fold the case bit, index the 26 letters into two packed constants (a switch here diverges per lane and was
// 80 % of the kernel's instructions).
__host__ __device__ constexpr uint32_t base_valid_mask() {
  const char* k = "ACGTURYSWKMBDHVN";
  uint32_t m = 0;
  for (int i = 0; i < 16; ++i) m |= 1u << (k[i] - 'A');
  return m;
}
__host__ __device__ constexpr uint64_t base_row_bits() {
  const char* k = "ACGTURYSWKMBDHVN";
  const int v[16] = {0, 1, 2, 3, 3, 0, 1, 1, 0, 2, 0, 1, 0, 0, 0, 0};
  uint64_t b = 0;
  for (int i = 0; i < 16; ++i) b |= (uint64_t)v[i] << (2 * (k[i] - 'A'));
  return b;
}
__host__ __device__ __forceinline__ int base_row(uint8_t ch) {
  constexpr uint32_t VALID = base_valid_mask();
  constexpr uint64_t ROWS = base_row_bits();
  const unsigned t = (unsigned)(ch | 0x20) - (unsigned)'a';      // 'A'..'Z' and 'a'..'z' -> 0..25, everything else >= 26
  const unsigned tc = t < 26u ? t : 26u;
  const bool ok = t < 26u && ((VALID >> tc) & 1u);
  return ok ? (int)((ROWS >> (2 * tc)) & 3u) : 255;
}

struct Alignments {
  // reads that passed the mapping-quality filter and the depth cap (CreateTensor.py:264,274-281), ascending POS
  const int32_t* read_pos;       // [R]   0-based POS
  const int32_t* read_end;       // [R]   one past the last reference position an M/=/X/D op covers
  const int32_t* read_maxend;    // [R]   running maximum of read_end (binary-search key for the first overlapping read)
  const int32_t* read_op0;       // [R+1] first op of each read
  const uint8_t* read_strand;    // [R]   FLAG & 16 != 0
  const int32_t* op_ref;         // [O]   reference position at the start of the op
  const int32_t* op_qry;         // [O]   offset of the op's first query base in `seq`
  const int32_t* op_len;         // [O]   length << 2 | OP_*
  const uint8_t* seq;            // query bases of all reads, back to back
  const uint8_t* ref;            // reference_sequence
  int32_t ref_start0;            // reference_start_0_based (:223)
  int32_t ref_len;
  int32_t n_reads;
};

// Event slots of a (position, row) cell and the channels the reference derives from them (CreateTensor.py:41-50):
//   channel 0 = aligned bases counted at the reference row, 1 = aligned at the query row + inserted bases,
//   2 = aligned at the reference row + deleted bases, 3 = aligned at the query row.
constexpr int SLOT_MREF = 0, SLOT_INS = 1, SLOT_DEL = 2, SLOT_MQRY = 3;
__host__ __device__ __forceinline__ void channels_from_slots(const int* s, int* ch) {
  ch[0] = s[SLOT_MREF];
  ch[1] = s[SLOT_MQRY] + s[SLOT_INS];
  ch[2] = s[SLOT_MREF] + s[SLOT_DEL];
  ch[3] = s[SLOT_MQRY];
}

// Rows of the reference bases under a candidate's window (255 = not a base / outside the loaded reference): the same 33
// bytes serve every read of the site, so the kernel resolves them once per site into shared memory.
__host__ __device__ __forceinline__ uint8_t window_row(const Alignments& a, int center, int i) {
  const int ri = center - (FLANK + 1) + i - a.ref_start0;
  return (uint8_t)((ri >= 0 && ri < a.ref_len) ? base_row(a.ref[ri]) : 255);
}

// One read folded into one candidate's counters; `win` = window_row(a, center, 0..32).  `Add` supplies add(int element): an atomic on shared memory in the
// kernel, a plain increment when this header is compiled for the host by tests/harness.
template <typename Add>
__host__ __device__ __forceinline__ bool fold_read_ops(const Alignments& a, int r, int center, bool left_edge, const uint8_t* win,
                                                       Add& add) {
  const int w0 = center - (FLANK + 1);                 // first window position (0-based)
  const int pos = a.read_pos[r], end = a.read_end[r];
  int q;
  if (left_edge) {
    if (pos > w0 + 2 * FLANK + 1 || end <= w0) return false;      // no aligned/deleted base in [c-17, c+16]
    q = pos > w0 ? pos : w0;
  } else {
    if (pos > w0 || end <= w0) return false;
    q = w0;
  }
  if (q >= end) return false;
  const int wend = w0 + N_POS;                         // one past the last counted position
  const int strand = a.read_strand[r] ? 4 : 0;
  int lo = a.read_op0[r], hi = a.read_op0[r + 1];
  const int op_end = hi;
  // first op whose reference end lies beyond q (an insertion ends where it starts, so this is an M or D op)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int lc = a.op_len[mid];
    const int e = a.op_ref[mid] + (((lc & 3) == OP_I) ? 0 : (lc >> 2));
    if (e > q) hi = mid; else lo = mid + 1;
  }
  // The counters of a cell are kept as four event slots, two increments per aligned base instead of four:
  //   SLOT_MREF aligned bases by reference row, SLOT_INS inserted bases, SLOT_DEL deleted bases, SLOT_MQRY aligned bases by
  //   query row; channels_from_slots() turns them into the reference's four channels when the row is written.
  if (lo >= op_end) return true;
  int p0 = a.op_ref[lo], lc = a.op_len[lo], qo = a.op_qry[lo];
  for (int k = lo; k < op_end; ++k) {
    if (p0 >= wend) break;
    // the next op's fields are requested before this op's bases are counted (one load latency per op, overlapped)
    const int kn = k + 1 < op_end ? k + 1 : k;
    const int np0 = a.op_ref[kn], nlc = a.op_len[kn], nqo = a.op_qry[kn];
    const int len = lc >> 2, code = lc & 3;
    if (code == OP_M) {
      const int s = p0 > q ? p0 : q, t = p0 + len < wend ? p0 + len : wend;
      const uint8_t* qs = a.seq + (qo - p0);
      for (int pb = s; pb < t; pb += 8) {
        uint8_t qch[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) qch[j] = pb + j < t ? qs[pb + j] : (uint8_t)0;      // eight loads in flight
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int p = pb + j;
          if (p >= t) break;
          const int rb = win[p - w0];
          const int qb = base_row(qch[j]);
          if (rb == 255 || qb == 255) continue;
          const int cell = (p - w0) * 32 + strand * 4;
          add.add(cell + rb * 4 + SLOT_MREF);
          add.add(cell + qb * 4 + SLOT_MQRY);
        }
      }
    } else if (code == OP_D) {
      const int s = p0 > q + 1 ? p0 : q + 1, t = p0 + len < wend ? p0 + len : wend;
      for (int p = s; p < t; ++p) {
        const int rb = win[p - w0];
        if (rb == 255) continue;
        add.add((p - w0) * 32 + strand * 4 + rb * 4 + SLOT_DEL);
      }
    } else if (p0 > q) {                               // insertion before reference position p0
      const int i0 = p0 - w0;
      for (int t = 0; t < len; ++t) {
        const int qb = base_row(a.seq[qo + t]);
        if (qb == 255) continue;
        const int idx = i0 + t < N_POS - 1 ? i0 + t : N_POS - 1;     // min(position_index + queryAdv, 32), :46
        add.add(idx * 32 + strand * 4 + qb * 4 + SLOT_INS);
      }
    }
    p0 = np0; lc = nlc; qo = nqo;
  }
  return true;
}

// The same rule walked position-major: all lanes of a site step through the window positions together (most reads cover the
// whole window, so the lanes stay converged) and the op cursor advances as a side branch; the query base of the next
// position is requested one iteration ahead.  A/B against fold_read_ops with -DCLAIRB_CT_FLAT.
template <typename Add>
__host__ __device__ __forceinline__ bool fold_read_flat(const Alignments& a, int r, int center, bool left_edge, const uint8_t* win,
                                                        Add& add) {
  const int w0 = center - (FLANK + 1);
  const int pos = a.read_pos[r], end = a.read_end[r];
  int q;
  if (left_edge) {
    if (pos > w0 + 2 * FLANK + 1 || end <= w0) return false;
    q = pos > w0 ? pos : w0;
  } else {
    if (pos > w0 || end <= w0) return false;
    q = w0;
  }
  if (q >= end) return false;
  const int wend = w0 + N_POS;
  const int strand = (a.read_strand[r] ? 4 : 0) * 4;
  int lo = a.read_op0[r], hi = a.read_op0[r + 1];
  const int op_end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int lc = a.op_len[mid];
    const int e = a.op_ref[mid] + (((lc & 3) == OP_I) ? 0 : (lc >> 2));
    if (e > q) hi = mid; else lo = mid + 1;
  }
  if (lo >= op_end) return true;
  int k = lo;
  int p0 = a.op_ref[k], lc = a.op_len[k], qo = a.op_qry[k];
  int len = lc >> 2, code = lc & 3;                    // an M or D op that covers q
  uint8_t ahead = 0;
  bool have_ahead = false;
  for (int p = q; p < wend; ++p) {
    while (code == OP_I || p >= p0 + len) {            // move the cursor to the op that covers p
      if (code == OP_I && p0 > q) {                    // inserted bases sit before reference position p0 == p
        const int i0 = p0 - w0;
        for (int t = 0; t < len; ++t) {
          const int qb = base_row(a.seq[qo + t]);
          if (qb == 255) continue;
          const int idx = i0 + t < N_POS - 1 ? i0 + t : N_POS - 1;
          add.add(idx * 32 + strand + qb * 4 + SLOT_INS);
        }
      }
      if (++k >= op_end) return true;
      p0 = a.op_ref[k]; lc = a.op_len[k]; qo = a.op_qry[k];
      len = lc >> 2; code = lc & 3;
      have_ahead = false;
    }
    const int rb = win[p - w0];
    if (code == OP_M) {
      const uint8_t* qs = a.seq + (qo - p0);
      const uint8_t ch = have_ahead ? ahead : qs[p];
      have_ahead = p + 1 < p0 + len && p + 1 < wend;
      if (have_ahead) ahead = qs[p + 1];
      const int qb = base_row(ch);
      if (rb != 255 && qb != 255) {
        const int cell = (p - w0) * 32 + strand;
        add.add(cell + rb * 4 + SLOT_MREF);
        add.add(cell + qb * 4 + SLOT_MQRY);
      }
    } else if (p > q && rb != 255) {
      add.add((p - w0) * 32 + strand + rb * 4 + SLOT_DEL);
    }
  }
  return true;
}

template <typename Add>
__host__ __device__ __forceinline__ bool fold_read(const Alignments& a, int r, int center, bool left_edge, const uint8_t* win, Add& add) {
#ifdef CLAIRB_CT_FLAT
  return fold_read_flat(a, r, center, left_edge, win, add);
#else
  return fold_read_ops(a, r, center, left_edge, win, add);
#endif
}

// Reads that can open the window of `center`: [first, last) with first = the first read whose running maximum end reaches
// into the window and last = the first read that starts beyond the last position that opens it (reads in between that end
// before the window are rejected by fold_read).
__host__ __device__ __forceinline__ void read_range(const Alignments& a, int center, bool left_edge, int& first, int& last) {
  const int w0 = center - (FLANK + 1);
  int lo = 0, hi = a.n_reads;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a.read_maxend[mid] > w0) hi = mid; else lo = mid + 1;
  }
  first = lo;
  const int last_open = left_edge ? w0 + 2 * FLANK + 1 : w0;
  lo = 0, hi = a.n_reads;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a.read_pos[mid] > last_open) hi = mid; else lo = mid + 1;
  }
  last = lo;
}

// read_range by the whole warp: both searches advance together, 32 probes per step each (33-ary search: 3 steps + a final
// sweep for 10^4..10^5 reads instead of 2 x 14 dependent loads - the window's reads cannot be touched before this returns).
__device__ __forceinline__ int warp_narrow(const int32_t* __restrict__ arr, int key, int lane, int& lo, int& hi) {
  // one step on [lo, hi): probes p_i = lo + (hi-lo)(i+1)/33; predicate arr[p] > key is monotone
  const int p = lo + (int)(((int64_t)(hi - lo) * (lane + 1)) / 33);
  const unsigned m = __ballot_sync(0xffffffffu, arr[p] > key);
  const int prev = __shfl_up_sync(0xffffffffu, p, 1);
  int nlo, nhi;
  if (m == 0) {
    nlo = __shfl_sync(0xffffffffu, p, 31) + 1;
    nhi = hi;
  } else {
    const int j = __ffs(m) - 1;
    nhi = __shfl_sync(0xffffffffu, p, j);
    nlo = j ? __shfl_sync(0xffffffffu, prev, j) + 1 : lo;
  }
  lo = nlo;
  hi = nhi;
  return hi - lo;
}

__device__ __forceinline__ int warp_finish(const int32_t* __restrict__ arr, int key, int lane, int lo, int hi) {
  // hi - lo <= 32: the first index in [lo, hi) with arr[idx] > key, or hi
  const int idx = lo + lane;
  const unsigned m = __ballot_sync(0xffffffffu, idx < hi && arr[idx] > key);
  return m ? lo + __ffs(m) - 1 : hi;
}

__device__ __forceinline__ void read_range_warp(const Alignments& a, int center, bool left_edge, int lane, int& first, int& last) {
  const int w0 = center - (FLANK + 1);
  const int last_open = left_edge ? w0 + 2 * FLANK + 1 : w0;
  int lo1 = 0, hi1 = a.n_reads, lo2 = 0, hi2 = a.n_reads;
  while (hi1 - lo1 > 32 || hi2 - lo2 > 32) {
    if (hi1 - lo1 > 32) warp_narrow(a.read_maxend, w0, lane, lo1, hi1);
    if (hi2 - lo2 > 32) warp_narrow(a.read_pos, last_open, lane, lo2, hi2);
  }
  first = warp_finish(a.read_maxend, w0, lane, lo1, hi1);
  last = warp_finish(a.read_pos, last_open, lane, lo2, hi2);
}

struct SharedAdd {
  int* cnt;
  __device__ __forceinline__ void add(int i) { atomicAdd(cnt + i, 1); }
};

// flags
constexpr int F_LEFT_EDGE = 1;      // default of the reference (absence of --stop_consider_left_edge)
constexpr int F_SUBTRACT = 2;       // write channels 1..3 minus channel 0 (clair/utils.py:96-98) instead of raw counts

// x_out: [n_centers][1056] int16; meta: [n_centers][2] = reads that opened the window (0: the reference prints no row),
// depth at the centre position (depth[flanking_base_num], compared with --minCoverage at :55); overflow: set when a
// count does not fit int16.
// One WARP per candidate site (a window is reached by a few dozen reads: one or two rounds of 32 lanes, one read per lane),
// SITES_PER_BLOCK sites per block, each with its own 4.2 KB of counters; only warp-level synchronisation.
constexpr int SITES_PER_BLOCK = THREADS / 32;

// Every op must carry a known code and stay inside `seq` (create_tensors reads it unguarded).  bad[0] / bad[1] receive the index of
// an op with an unknown code / an op that reads past the end of SEQ (-1: none); create_tensors does nothing when either is set.
// 10^7 ops per region: on the host this check was 8 of the call's 21 ms, here it is one pass over 8 bytes per op.
__global__ void validate_ops(const int32_t* __restrict__ op_qry, const int32_t* __restrict__ op_len, int n_ops, int64_t seq_len,
                             int* __restrict__ bad) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_ops; k += gridDim.x * blockDim.x) {
    const int len = op_len[k] >> 2, code = op_len[k] & 3;
    if (code != OP_D && code != OP_M && code != OP_I) atomicMax(&bad[0], k);
    else if (len < 0 || (code != OP_D && (op_qry[k] < 0 || (int64_t)op_qry[k] + len > seq_len))) atomicMax(&bad[1], k);
  }
}

__global__ void __launch_bounds__(THREADS) create_tensors(Alignments a, const int32_t* __restrict__ centers, int n_centers,
                                                          int flags, int16_t* __restrict__ x_out, int32_t* __restrict__ meta,
                                                          int* __restrict__ overflow, const int* __restrict__ bad) {
  if (bad[0] >= 0 || bad[1] >= 0) return;                // validate_ops (same stream, just before) found a malformed op
  __shared__ int cnt_all[SITES_PER_BLOCK][ELEMS];
  __shared__ uint8_t win_all[SITES_PER_BLOCK][N_POS + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = cnt_all[warp];
  uint8_t* win = win_all[warp];
  const bool left_edge = flags & F_LEFT_EDGE;
  for (int ci = blockIdx.x * SITES_PER_BLOCK + warp; ci < n_centers; ci += gridDim.x * SITES_PER_BLOCK) {
    const int center = centers[ci];
    for (int i = lane; i < ELEMS; i += 32) cnt[i] = 0;
    for (int i = lane; i < N_POS; i += 32) win[i] = window_row(a, center, i);
    int first, last;
    read_range_warp(a, center, left_edge, lane, first, last);
    __syncwarp();
    SharedAdd add{cnt};
    int opened = 0;
    for (int r = first + lane; r < last; r += 32) opened += fold_read(a, r, center, left_edge, win, add) ? 1 : 0;
    opened = __reduce_add_sync(0xffffffffu, opened);
    __syncwarp();
    // one row out: 528 words of two int16 = channels (0,1) or (2,3) of a cell
    uint32_t* out = reinterpret_cast<uint32_t*>(x_out + (size_t)ci * ELEMS);
    bool ovf = false;
    for (int w = lane; w < ELEMS / 2; w += 32) {
      int ch[4];
      channels_from_slots(cnt + (w >> 1) * 4, ch);
      int v0 = (w & 1) ? ch[2] : ch[0], v1 = (w & 1) ? ch[3] : ch[1];
      if (flags & F_SUBTRACT) {
        if (w & 1) v0 -= ch[0];
        v1 -= ch[0];
      }
      ovf |= v0 > 32767 || v1 > 32767 || v0 < -32768 || v1 < -32768;
      out[w] = (uint32_t)(uint16_t)(int16_t)v0 | ((uint32_t)(uint16_t)(int16_t)v1 << 16);
    }
    if (ovf) *overflow = 1;
    // depth at the centre = aligned bases at position index 16 = sum over the 8 rows of channel 0
    int d = lane < 8 ? cnt[FLANK * 32 + lane * 4] : 0;
    d = __reduce_add_sync(0xffffffffu, d);
    if (lane == 0) {
      meta[2 * ci] = opened;
      meta[2 * ci + 1] = d;
    }
    __syncwarp();
  }
}

// rows[i] of the resident tensor block -> a dense block the forward reads (the candidates the reference would have printed
// and tensor_generator_from would have kept)
__global__ void gather_rows(const int16_t* __restrict__ x_all, const int64_t* __restrict__ rows, int64_t n,
                            int16_t* __restrict__ x_out) {
  const int64_t i = blockIdx.x;
  if (i >= n) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(x_all + (size_t)rows[i] * ELEMS);
  uint32_t* dst = reinterpret_cast<uint32_t*>(x_out + (size_t)i * ELEMS);
  for (int w = threadIdx.x; w < ELEMS / 2; w += blockDim.x) dst[w] = src[w];
}

}  // namespace ct
}  // namespace clairb
