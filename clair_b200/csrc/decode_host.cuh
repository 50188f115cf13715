// Native text decode of CreateTensor rows (host code, no device work): the per-row `str.split()` +
// `np.array(list[str], float32)` of the reference generator (clair/utils.py:81-98) costs ~140 us per site in Python;
// this parses a whole predict-batch of lines in one call.
//
// Same semantics as the reference for the rows it accepts: a row is `ctg pos seq v0 ... v1055` separated by any
// whitespace; rows whose centre base seq[16] is not an IUPAC code are dropped (clair/utils.py:90); channels 1..3 of the
// kept rows have channel 0 subtracted (clair/utils.py:96-98).  Values are parsed as integers (what CreateTensor.py:60-65
// prints); a token that is not a plain integer goes through strtof, like numpy's string -> float32 conversion.
#pragma once
#include <stdint.h>

#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace clairb {
namespace decode {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
inline bool is_iupac(char c) { return c != '\0' && strchr("ACGTURYSWKMBDHVN", c) != nullptr; }

struct Line {
  int64_t begin, end;      // [begin, end) without the newline
  int64_t values;          // offset of the first byte after the sequence token
  int64_t out_row;         // row in x_out, -1 = dropped by the IUPAC filter
};

// values of one line -> out (or just validated when out == nullptr).  0 ok, 1 malformed, 2 not an int16 integer
template <typename TOut>
inline int parse_values(const char* text, const Line& ln, TOut* out) {
  int64_t p = ln.values;
  const int64_t end = ln.end;
  int nval = 0;
  while (true) {
    while (p < end && is_space(text[p])) ++p;
    if (p >= end) break;
    if (nval == SITE_ELEMS) return 1;                   // too many columns
    int64_t q = p;
    bool neg = false;
    if (text[q] == '-' || text[q] == '+') { neg = text[q] == '-'; ++q; }
    int64_t v = 0;
    int digits = 0;
    unsigned d;
    while (q < end && (d = (unsigned)(text[q] - '0')) <= 9u && digits < 18) { v = v * 10 + d; ++q; ++digits; }
    if (digits > 0 && (q == end || is_space(text[q]))) {
      if (neg) v = -v;
      if (sizeof(TOut) == 2 && (v < -32768 || v > 32767)) return 2;
      if (out) out[nval] = (TOut)v;
    } else {
      // general token (a float, an exponent, inf/nan): what numpy's str -> float32 accepts
      while (q < end && !is_space(text[q])) ++q;
      char buf[64];
      const int64_t tl = q - p;
      if (sizeof(TOut) == 2) return 2;
      if (tl <= 0 || tl >= (int64_t)sizeof buf) return 1;
      memcpy(buf, text + p, (size_t)tl);
      buf[tl] = 0;
      char* ep = nullptr;
      const float f = strtof(buf, &ep);
      if (ep != buf + tl) return 1;
      if (out) out[nval] = (TOut)f;
    }
    p = q;
    ++nval;
  }
  if (nval != SITE_ELEMS) return 1;
  if (out) {
    for (int i = 0; i < SITE_ELEMS; i += 4) {            // X[..., 1:] -= X[..., 0:1]
      const TOut r0 = out[i];
      out[i + 1] = (TOut)(out[i + 1] - r0);
      out[i + 2] = (TOut)(out[i + 2] - r0);
      out[i + 3] = (TOut)(out[i + 3] - r0);
    }
  }
  return 0;
}

// 0 ok, 1 malformed row (wrong column count / short sequence), 2 value not representable (int16 mode).
// Pass 1 (serial) finds the lines, their three info tokens and the filter decision, i.e. every row's slot in x_out;
// pass 2 parses the 1056 values of every line on `threads` host threads.
template <typename TOut>
inline int rows(const char* text, int64_t len, int64_t max_rows, TOut* x_out, int32_t* info_off, int64_t* rows_read,
                int64_t* rows_kept, int64_t* consumed, int64_t* bad_row, int threads) {
  std::vector<Line> lines;
  lines.reserve((size_t)(max_rows < 65536 ? max_rows : 65536));
  int64_t pos = 0, nkept = 0;
  *bad_row = -1;
  while ((int64_t)lines.size() < max_rows && pos < len) {
    const char* nl = (const char*)memchr(text + pos, '\n', (size_t)(len - pos));
    if (!nl) break;                                     // incomplete last line: the caller completes it
    const int64_t end = nl - text;
    int64_t p = pos;
    int32_t off[6];
    bool ok = true;
    for (int k = 0; k < 3 && ok; ++k) {                  // ctg, pos, seq
      while (p < end && is_space(text[p])) ++p;
      const int64_t a = p;
      while (p < end && !is_space(text[p])) ++p;
      ok = p > a;
      off[2 * k] = (int32_t)a;
      off[2 * k + 1] = (int32_t)p;
    }
    if (!ok || off[5] - off[4] <= 16) { *bad_row = (int64_t)lines.size(); return 1; }
    Line ln{pos, end, p, -1};
    if (is_iupac(text[off[4] + 16])) {
      ln.out_row = nkept;
      for (int k = 0; k < 6; ++k) info_off[nkept * 6 + k] = off[k];
      ++nkept;
    }
    lines.push_back(ln);
    pos = end + 1;
  }
  const int64_t n = (int64_t)lines.size();
  std::atomic<int64_t> next(0), first_bad(-1);
  std::atomic<int> status(0);
  auto work = [&]() {
    while (true) {
      const int64_t i0 = next.fetch_add(16);
      if (i0 >= n || status.load(std::memory_order_relaxed)) return;
      for (int64_t i = i0; i < n && i < i0 + 16; ++i) {
        const Line& ln = lines[(size_t)i];
        const int rc = parse_values<TOut>(text, ln, ln.out_row >= 0 ? x_out + ln.out_row * SITE_ELEMS : (TOut*)nullptr);
        if (rc) {
          int expected = 0;
          if (status.compare_exchange_strong(expected, rc)) first_bad.store(i);
          return;
        }
      }
    }
  };
  int nt = threads < 1 ? 1 : threads;
  if (n < 64) nt = 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  if (status.load()) { *bad_row = first_bad.load(); return status.load(); }
  *rows_read = n;
  *rows_kept = nkept;
  *consumed = pos;
  return 0;
}

}  // namespace decode
}  // namespace clairb
