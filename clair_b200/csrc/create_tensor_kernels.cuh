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

// Build switches (A/B, all parity-green on host and device): default = each lane follows its read's ops from global memory
// (0.58 ms on the 2 Mb ONT-like region of tools/ct_bench.py); -DCLAIRB_CT_FLAT = position-major walk from global memory
// (0.97 ms); -DCLAIRB_CT_STAGED = operands staged in shared memory, then all lanes walk the 33 positions together (0.75 ms).

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

// ---- the same rule from STAGED operands (-DCLAIRB_CT_STAGED; measured, not the default) -----------------------------------------
// fold_read_ops follows each read's own ops: the lanes of a warp sit in different op types and loop lengths (8.9 of 32 threads
// active per instruction, 9.3 k warp instructions per site).  Here a read is first staged for its window - the (length, code)
// words of the ops under the window and the query bytes they use, copied to shared memory once, with plain independent loads -
// and then all lanes step through the 33 window positions TOGETHER, the op cursor advancing from shared memory as a side branch.
// Result on B200 (ncu, profiles/r02_ct_staged_walk.txt): the positions are walked converged (20 of 32 lanes at the loop head:
// 40 reads are 32 + 8), but the kernel executes MORE warp instructions (10.8 k per site) and takes 0.75 ms against 0.58: staging
// and its consistency checks cost what the walk saves; inserted bases are still counted by one lane at a time (an indel every
// ~8 bases in the bench region: 10 % of the instructions at 1.0 threads each); fetching a query byte out of the staged words and
// turning it into a row is 30 % of the instructions; and converged lanes hit the SAME counter with their shared-memory atomics.
// Kept, parity-green on the host (tests/harness) and on the device, as the record of the attempt.
constexpr int STAGE_OPS = 20;             // ops kept per read (an ONT window of 33 bases holds 2 x Poisson(4) + 1 ops)
constexpr int STAGE_WORDS = 16;           // 64 query bytes per read, from a 4-byte aligned address
struct Staged {
  int q;                                  // first counted reference position
  int p0;                                 // reference position of the first staged op
  int qi;                                 // index, within the staged bytes, of the first staged op's first query base (can be < 0)
  int n_ops;
};
// -> 0: the read does not open the window (fold_read_ops would return false); 1: staged; 2: needs fold_read_ops - more ops or query
// bytes under the window than are staged, or op_ref / op_qry are not the running sums of the lengths before them (the staged walk
// rebuilds positions from lengths; the arrays of the host encoder always are, a foreign caller's may not be).
// Store: put_op(j, length_code), put_word(w, four query bytes).
template <typename Store>
__host__ __device__ __forceinline__ int stage_read(const Alignments& a, int r, int center, bool left_edge, Store& st, Staged& out) {
  const int w0 = center - (FLANK + 1);
  const int pos = a.read_pos[r], end = a.read_end[r];
  int q;
  if (left_edge) {
    if (pos > w0 + 2 * FLANK + 1 || end <= w0) return 0;
    q = pos > w0 ? pos : w0;
  } else {
    if (pos > w0 || end <= w0) return 0;
    q = w0;
  }
  if (q >= end) return 0;
  const int wend = w0 + N_POS;
  int lo = a.read_op0[r], hi = a.read_op0[r + 1];
  const int op_end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int lc = a.op_len[mid];
    const int e = a.op_ref[mid] + (((lc & 3) == OP_I) ? 0 : (lc >> 2));
    if (e > q) hi = mid; else lo = mid + 1;
  }
  out.q = q;
  out.n_ops = 0;
  out.p0 = 0;
  out.qi = 0;
  if (lo >= op_end) return 1;
  int run_p = a.op_ref[lo], run_q = a.op_qry[lo];
  const int lc0 = a.op_len[lo];
  // the first query byte the window can use: the base under q of an aligned op, or the next base behind a deletion
  const int first_byte = run_q + (((lc0 & 3) == OP_M) ? q - run_p : 0);
  const int base = first_byte & ~3;
  int need_end = first_byte;
  out.p0 = run_p;
  out.qi = run_q - base;
  int n = 0;
  for (int j0 = 0; j0 < STAGE_OPS; j0 += 4) {            // four ops' fields requested at a time
    int lc[4], pr[4], qq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = lo + j0 + u < op_end ? lo + j0 + u : op_end - 1;
      lc[u] = a.op_len[idx]; pr[u] = a.op_ref[idx]; qq[u] = a.op_qry[idx];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (lo + j0 + u >= op_end || run_p >= wend) { j0 = STAGE_OPS; break; }      // no more ops / the window is covered
      if (pr[u] != run_p || qq[u] != run_q) return 2;
      st.put_op(j0 + u, lc[u]);
      n = j0 + u + 1;
      const int len = lc[u] >> 2, code = lc[u] & 3;
      if (code == OP_M) { const int t = run_p + len < wend ? len : wend - run_p; if (run_q + t > need_end) need_end = run_q + t; }
      else if (code == OP_I && run_p > q && run_q + len > need_end) need_end = run_q + len;
      if (code != OP_I) run_p += len;
      if (code != OP_D) run_q += len;
    }
  }
  if (n == STAGE_OPS && lo + n < op_end && run_p < wend) return 2;
  if (need_end - base > 4 * STAGE_WORDS) return 2;
  out.n_ops = n;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(a.seq + base);       // `seq` carries 64 bytes of slack behind its end
  const int n_words = (need_end - base + 3) >> 2;
#pragma unroll 4
  for (int w = 0; w < STAGE_WORDS; ++w)
    if (w < n_words) st.put_word(w, words[w]);
  return 1;
}

// The walk of fold_read_flat over staged operands.  Load: op(j), byte(i).  Written so that the lanes of a warp stay in lockstep over
// the 33 positions (uniform loop bounds, the lane's own range as a predicate).
template <typename Add, typename Load>
__host__ __device__ __forceinline__ void fold_read_staged(const Staged& s, int center, int strand16, const uint8_t* win, Load& ld, Add& add) {
  const int w0 = center - (FLANK + 1), wend = w0 + N_POS, q = s.q;
  bool live = s.n_ops > 0;
  int k = 0, p0 = s.p0, qi = s.qi;
  int lc = live ? ld.op(0) : 0;
  int len = lc >> 2, code = lc & 3;
  for (int p = w0; p < wend; ++p) {
    ld.converge();                                       // the lanes that walk (uniform loop bounds) meet here every position ...
    const bool on = live && p >= q;
    if (on) {
      while (code == OP_I || p >= p0 + len) {            // move the cursor to the op that covers p
        if (code == OP_I && p0 > q) {                    // inserted bases sit before reference position p0 == p
          const int i0 = p0 - w0;
          for (int t = 0; t < len; ++t) {
            const int qb = base_row(ld.byte(qi + t));
            if (qb == 255) continue;
            const int idx = i0 + t < N_POS - 1 ? i0 + t : N_POS - 1;
            add.add(idx * 32 + strand16 + qb * 4 + SLOT_INS);
          }
        }
        if (code != OP_D) qi += len;
        if (code != OP_I) p0 += len;
        if (++k >= s.n_ops) { live = false; break; }
        lc = ld.op(k);
        len = lc >> 2; code = lc & 3;
      }
    }
    ld.converge();                                       // ... and again behind the cursor moves, which only some lanes make
    if (on && live) {
      const int rb = win[p - w0];
      if (code == OP_M) {
        const int qb = base_row(ld.byte(qi + (p - p0)));
        if (rb != 255 && qb != 255) {
          const int cell = (p - w0) * 32 + strand16;
          add.add(cell + rb * 4 + SLOT_MREF);
          add.add(cell + qb * 4 + SLOT_MQRY);
        }
      } else if (p > q && rb != 255) {
        add.add((p - w0) * 32 + strand16 + rb * 4 + SLOT_DEL);
      }
    }
  }
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
// a lane's staged operands: ops [STAGE_OPS][32 lanes], query words [STAGE_WORDS][32 lanes] (a lane's column: no bank conflicts)
struct SharedStage {
  int* ops;
  uint32_t* words;
  int lane;
  unsigned walkers;                                      // lanes inside fold_read_staged in this round
  __device__ __forceinline__ void converge() const { __syncwarp(walkers); }
  __device__ __forceinline__ void put_op(int j, int lc) { ops[j * 32 + lane] = lc; }
  __device__ __forceinline__ void put_word(int w, uint32_t v) { words[w * 32 + lane] = v; }
  __device__ __forceinline__ int op(int j) const { return ops[j * 32 + lane]; }
  __device__ __forceinline__ uint8_t byte(int i) const { return (uint8_t)(words[(i >> 2) * 32 + lane] >> (8 * (i & 3))); }
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
#ifdef CLAIRB_CT_STAGED
  __shared__ int ops_all[SITES_PER_BLOCK][STAGE_OPS * 32];
  __shared__ uint32_t words_all[SITES_PER_BLOCK][STAGE_WORDS * 32];
  SharedStage stage{ops_all[warp], words_all[warp], lane, 0u};
#endif
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
#ifndef CLAIRB_CT_STAGED
    for (int r = first + lane; r < last; r += 32) opened += fold_read(a, r, center, left_edge, win, add) ? 1 : 0;
#else
    for (int r0 = first; r0 < last; r0 += 32) {            // 32 reads per round: stage each, then walk the window together
      const int r = r0 + lane;
      Staged sg;
      const int state = r < last ? stage_read(a, r, center, left_edge, stage, sg) : 0;
      opened += state != 0;
      stage.walkers = __ballot_sync(0xffffffffu, state == 1);
      if (state == 1) fold_read_staged(sg, center, a.read_strand[r] ? 16 : 0, win, stage, add);
      else if (state == 2) fold_read_ops(a, r, center, left_edge, win, add);
      __syncwarp();
    }
#endif
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
