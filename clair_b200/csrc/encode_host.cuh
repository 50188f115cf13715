// Native encode of `samtools view` rows into the flat arrays clairb_create_tensors takes (host code, no device work).
// Replaces what the reference does per SAM row before and while it walks the CIGAR string
// (dataPrepScripts/CreateTensor.py:251-296): header rows skipped (:253), FLAG / POS / MAPQ / CIGAR / SEQ columns (:256-262),
// the mapping-quality filter (:264), the per-POS depth cap (:274-281), and the CIGAR grammar of :283-366 - every non-digit
// character is an op whose length is the digits before it; S advances the query, M = X advance both, I the query, D the
// reference, anything else (N, H, P, *) advances nothing.
// Pass 1 (serial) finds the rows and applies the filters - they depend on row order; pass 2 counts the kept ops of every
// read and pass 3 writes them, both on `threads` host threads.
#pragma once
#include <stdint.h>

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "create_tensor_kernels.cuh"

namespace clairb {
namespace sam {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

struct Read {
  const char* cigar;
  const char* seq;
  int32_t cigar_len, seq_len;
  int32_t pos;
  uint8_t strand;
  int64_t n_ops, ref_span;     // pass 2
};

struct Out {
  int32_t *read_pos, *read_end, *read_op0;
  uint8_t* read_strand;
  int32_t *op_ref, *op_qry, *op_len;
  uint8_t* seq;
};

enum { OK = 0, MALFORMED = 1, UNSORTED = 2, CIGAR_BEYOND_SEQ = 3, CAPACITY = 4, TOO_LARGE = 5 };

inline bool parse_int(const char* a, const char* b, int64_t* v) {
  if (a >= b) return false;
  bool neg = false;
  if (*a == '-' || *a == '+') { neg = *a == '-'; ++a; }
  if (a >= b || b - a > 18) return false;
  int64_t x = 0;
  for (; a < b; ++a) {
    const unsigned d = (unsigned)(*a - '0');
    if (d > 9u) return false;
    x = x * 10 + d;
  }
  *v = neg ? -x : x;
  return true;
}

// One CIGAR string.  fill == false: counts kept ops and the reference span, checks the query bounds.
// fill == true: writes ops at out->op_*[op_at...] (reference positions from `pos`, query offsets from `seq_at`).
inline int walk_cigar(const Read& r, bool fill, const Out* out, int64_t op_at, int64_t seq_at, int64_t* n_ops, int64_t* ref_span) {
  int64_t n = 0, ref = 0, qry = 0, ops = 0;
  for (int32_t i = 0; i < r.cigar_len; ++i) {
    const char c = r.cigar[i];
    const unsigned d = (unsigned)(c - '0');
    if (d <= 9u) {
      n = n * 10 + d;
      if (n >= ((int64_t)1 << 29)) return TOO_LARGE;
      continue;
    }
    int code = -1;
    if (c == 'M' || c == '=' || c == 'X') code = ct::OP_M;
    else if (c == 'I') code = ct::OP_I;
    else if (c == 'D') code = ct::OP_D;
    if (code >= 0 && n > 0) {
      if (code != ct::OP_D && qry + n > r.seq_len) return CIGAR_BEYOND_SEQ;
      if (fill) {
        out->op_ref[op_at + ops] = (int32_t)(r.pos + ref);
        out->op_qry[op_at + ops] = (int32_t)(seq_at + qry);
        out->op_len[op_at + ops] = (int32_t)((n << 2) | code);
      }
      ++ops;
    }
    if (c == 'S' || code == ct::OP_M || code == ct::OP_I) qry += n;
    if (code == ct::OP_M || code == ct::OP_D) ref += n;
    n = 0;
  }
  *n_ops = ops;
  *ref_span = ref;
  return OK;
}

template <typename F>
inline void parallel_for(int64_t n, int threads, F body) {
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    while (true) {
      const int64_t i0 = next.fetch_add(8);
      if (i0 >= n) return;
      for (int64_t i = i0; i < n && i < i0 + 8; ++i) body(i);
    }
  };
  int nt = threads < 1 ? 1 : threads;
  if (n < 32) nt = 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

// state: [0] previous_position, [1] depthCap (CreateTensor.py:249-250, carried across blocks of one stream),
//        [2] POS of the last kept read (sortedness across blocks).  Updated only when `out` is given.
inline int encode(const char* text, int64_t len, int min_mq, int dcov, int32_t* state, const Out* out, int64_t cap_reads,
                  int64_t cap_ops, int64_t cap_bases, int64_t* n_reads, int64_t* n_ops, int64_t* n_bases, int64_t* bad_line,
                  int threads) {
  std::vector<Read> reads;
  int64_t previous_position = state[0], depth_cap = state[1], last_pos = state[2];
  int64_t p = 0, line_no = 0;
  *bad_line = -1;
  while (p < len) {
    const char* nl = (const char*)memchr(text + p, '\n', (size_t)(len - p));
    const int64_t end = nl ? nl - text : len;
    const char* f[10][2];
    int nf = 0;
    int64_t q = p;
    while (nf < 10) {
      while (q < end && is_space(text[q])) ++q;
      if (q >= end) break;
      f[nf][0] = text + q;
      if (nf == 5 || nf == 9) {
        // CIGAR and SEQ are the long columns: jump to the next tab, and only rescan byte by byte when the span holds
        // another kind of white space (rows not written by samtools)
        const char* tab = (const char*)memchr(text + q, '\t', (size_t)(end - q));
        int64_t stop = tab ? tab - text : end;
        if (stop > q && text[stop - 1] == '\r') --stop;
        bool plain = true;
        for (const char ws : {' ', '\r', '\v', '\f'}) plain = plain && !memchr(text + q, ws, (size_t)(stop - q));
        if (plain) q = stop;
        else while (q < end && !is_space(text[q])) ++q;
      } else {
        while (q < end && !is_space(text[q])) ++q;
      }
      f[nf][1] = text + q;
      ++nf;
    }
    const int64_t this_line = line_no++;
    p = end + 1;
    if (nf == 0) { *bad_line = this_line; return MALFORMED; }        // the reference raises IndexError on an empty row
    if (f[0][0][0] == '@') continue;
    int64_t flag, pos1, mq;
    if (nf < 10 || !parse_int(f[1][0], f[1][1], &flag) || !parse_int(f[3][0], f[3][1], &pos1) || !parse_int(f[4][0], f[4][1], &mq)) {
      *bad_line = this_line;
      return MALFORMED;
    }
    if (mq < min_mq) continue;
    const int64_t pos = pos1 - 1;
    if (previous_position != pos) {
      previous_position = pos;
      depth_cap = 0;
    } else {
      depth_cap += 1;
      if (depth_cap >= dcov) continue;
    }
    if (pos < last_pos) { *bad_line = this_line; return UNSORTED; }
    if (pos < -1 || pos >= ((int64_t)1 << 31) - 64 || f[9][1] - f[9][0] >= ((int64_t)1 << 31) - 64 || f[5][1] - f[5][0] >= ((int64_t)1 << 31) - 64) {
      *bad_line = this_line;
      return TOO_LARGE;
    }
    last_pos = pos;
    Read r;
    r.cigar = f[5][0];
    r.cigar_len = (int32_t)(f[5][1] - f[5][0]);
    r.seq = f[9][0];
    r.seq_len = (int32_t)(f[9][1] - f[9][0]);
    r.pos = (int32_t)pos;
    r.strand = (flag & 16) ? 1 : 0;
    r.n_ops = r.ref_span = 0;
    reads.push_back(r);
  }
  const int64_t R = (int64_t)reads.size();
  std::atomic<int> status(0);
  std::atomic<int64_t> bad(-1);
  parallel_for(R, threads, [&](int64_t i) {
    Read& r = reads[(size_t)i];
    const int rc = walk_cigar(r, false, nullptr, 0, 0, &r.n_ops, &r.ref_span);
    if (rc) {
      int expected = 0;
      if (status.compare_exchange_strong(expected, rc)) bad.store(i);
    }
  });
  if (status.load()) { *bad_line = bad.load(); return status.load(); }
  int64_t O = 0, B = 0;
  for (const Read& r : reads) {
    O += r.n_ops;
    B += r.seq_len;
    if ((int64_t)r.pos + r.ref_span >= ((int64_t)1 << 31) - 64) return TOO_LARGE;
  }
  *n_reads = R;
  *n_ops = O;
  *n_bases = B;
  if (O >= ((int64_t)1 << 31) - 64 || B >= ((int64_t)1 << 31) - 64) return TOO_LARGE;
  if (!out) return OK;
  if (R > cap_reads || O > cap_ops || B > cap_bases) return CAPACITY;
  std::vector<int64_t> op0((size_t)R + 1), seq0((size_t)R + 1);
  op0[0] = seq0[0] = 0;
  for (int64_t i = 0; i < R; ++i) {
    op0[(size_t)i + 1] = op0[(size_t)i] + reads[(size_t)i].n_ops;
    seq0[(size_t)i + 1] = seq0[(size_t)i] + reads[(size_t)i].seq_len;
  }
  out->read_op0[R] = (int32_t)O;
  parallel_for(R, threads, [&](int64_t i) {
    const Read& r = reads[(size_t)i];
    out->read_pos[i] = r.pos;
    out->read_end[i] = (int32_t)(r.pos + r.ref_span);
    out->read_op0[i] = (int32_t)op0[(size_t)i];
    out->read_strand[i] = r.strand;
    int64_t no, rs;
    walk_cigar(r, true, out, op0[(size_t)i], seq0[(size_t)i], &no, &rs);
    memcpy(out->seq + seq0[(size_t)i], r.seq, (size_t)r.seq_len);
  });
  state[0] = (int32_t)previous_position;
  state[1] = (int32_t)(depth_cap > 0x7fffffff ? 0x7fffffff : depth_cap);
  state[2] = (int32_t)last_pos;
  return OK;
}

}  // namespace sam
}  // namespace clairb
