// Native text rows of CreateTensor (host code, no device work): what the reference prints per tensor,
//   "%s %d %s %s" % (ctg_name, center, reference_sequence[window], " ".join("%d" % x for every count))
// (dataPrepScripts/CreateTensor.py:57-62) - 1,059 tokens per row, ~0.3 ms per row when formatted in Python.
// Pass 1 measures every row, pass 2 writes them in place, both on `threads` host threads.
#pragma once
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "encode_host.cuh"

namespace clairb {
namespace fmt {

inline int int_len(int64_t v) {
  int n = v < 0 ? 1 : 0;
  uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
  do { ++n; u /= 10; } while (u);
  return n;
}
inline char* put_int(char* p, int64_t v) {
  const int n = int_len(v);
  uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
  char* e = p + n;
  do { *--e = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) *--e = '-';
  return p + n;
}

// rows: n sites; x [n][1056] int16; window text = ref[start[i] .. start[i]+33) cut at ref_len.
// out == nullptr: only *out_len (bytes needed).  Returns 0, or 1 when out_cap is too small.
inline int rows(const char* ctg, const int64_t* positions, const char* ref, int64_t ref_len, const int64_t* start, const int16_t* x,
                int64_t n, char* out, int64_t out_cap, int64_t* out_len, int threads) {
  const int64_t ctg_len = (int64_t)strlen(ctg);
  std::vector<int64_t> at((size_t)n + 1);
  auto window = [&](int64_t i, int64_t* a, int64_t* b) {
    int64_t s = start[i] < 0 ? 0 : start[i], e = start[i] + 33;
    if (s > ref_len) s = ref_len;
    if (e > ref_len) e = ref_len;
    if (e < s) e = s;
    *a = s;
    *b = e;
  };
  sam::parallel_for(n, threads, [&](int64_t i) {
    int64_t a, b;
    window(i, &a, &b);
    int64_t len = ctg_len + 1 + int_len(positions[i]) + 1 + (b - a) + 1;
    const int16_t* r = x + i * 1056;
    for (int k = 0; k < 1056; ++k) len += int_len(r[k]);
    len += 1055 + 1;                                     // separators and the newline
    at[(size_t)i + 1] = len;
  });
  at[0] = 0;
  for (int64_t i = 0; i < n; ++i) at[(size_t)i + 1] += at[(size_t)i];
  *out_len = at[(size_t)n];
  if (!out) return 0;
  if (out_cap < at[(size_t)n]) return 1;
  sam::parallel_for(n, threads, [&](int64_t i) {
    char* p = out + at[(size_t)i];
    memcpy(p, ctg, (size_t)ctg_len);
    p += ctg_len;
    *p++ = ' ';
    p = put_int(p, positions[i]);
    *p++ = ' ';
    int64_t a, b;
    window(i, &a, &b);
    memcpy(p, ref + a, (size_t)(b - a));
    p += b - a;
    const int16_t* r = x + i * 1056;
    for (int k = 0; k < 1056; ++k) {
      *p++ = ' ';
      p = put_int(p, r[k]);
    }
    *p++ = '\n';
  });
  return 0;
}

// VCF rows of reference / SNP calls (host code): what output_with prints per site,
//   "%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f"   (clair/call_var.py:1184-1197)
// for calls whose REF is one base and whose ALT is one base or "X,Y".  Rows are separated by '\n' (none after the last:
// the caller's print adds it).  row_end[i] = offset one past row i.  Returns 0, or 1 when out_cap is too small.
inline int vcf_rows(int64_t n, const char* ctg_blob, const int32_t* ctg_off, const int64_t* pos, const uint8_t* ref, const uint8_t* alt,
                    const int32_t* quality, const uint8_t* filter_code, const uint8_t* gt_code, const int32_t* depth, const double* af,
                    char* out, int64_t out_cap, int64_t* out_len, int64_t* row_end, int threads) {
  static const char* const kFilter[3] = {".", "PASS", "LowQual"};
  static const char* const kGenotype[6] = {"0/0", "1/1", "0/1", "1/2", "0", "1"};
  std::vector<int64_t> at((size_t)n + 1);
  auto af_text = [](double v, char* buf) { return snprintf(buf, 32, "%.4f", v); };
  sam::parallel_for(n, threads, [&](int64_t i) {
    char buf[32];
    int64_t len = (ctg_off[i + 1] - ctg_off[i]) + 1 + int_len(pos[i]) + 3 + 1 + 1 + (int64_t)strlen((const char*)alt + 4 * i) + 1 +
                  int_len(quality[i]) + 1 + (int64_t)strlen(kFilter[filter_code[i]]) + 3 + 12 +
                  (int64_t)strlen(kGenotype[gt_code[i]]) + 1 + int_len(quality[i]) + 1 + int_len(depth[i]) + 1 + af_text(af[i], buf);
    at[(size_t)i + 1] = len + 1;                          // + the row separator
  });
  at[0] = 0;
  for (int64_t i = 0; i < n; ++i) at[(size_t)i + 1] += at[(size_t)i];
  *out_len = n ? at[(size_t)n] - 1 : 0;
  if (!out) return 0;
  if (out_cap < at[(size_t)n]) return 1;
  sam::parallel_for(n, threads, [&](int64_t i) {
    char* p = out + at[(size_t)i];
    const int64_t cl = ctg_off[i + 1] - ctg_off[i];
    memcpy(p, ctg_blob + ctg_off[i], (size_t)cl);
    p += cl;
    *p++ = '\t';
    p = put_int(p, pos[i]);
    memcpy(p, "\t.\t", 3);
    p += 3;
    *p++ = (char)ref[i];
    *p++ = '\t';
    for (const char* a = (const char*)alt + 4 * i; *a; ++a) *p++ = *a;
    *p++ = '\t';
    p = put_int(p, quality[i]);
    *p++ = '\t';
    for (const char* a = kFilter[filter_code[i]]; *a; ++a) *p++ = *a;
    memcpy(p, "\t.\tGT:GQ:DP:AF\t", 15);
    p += 15;
    for (const char* a = kGenotype[gt_code[i]]; *a; ++a) *p++ = *a;
    *p++ = ':';
    p = put_int(p, quality[i]);
    *p++ = ':';
    p = put_int(p, depth[i]);
    *p++ = ':';
    char buf[32];
    const int k = af_text(af[i], buf);
    memcpy(p, buf, (size_t)k);
    p += k;
    *p++ = '\n';
    row_end[i] = at[(size_t)i + 1] - 1;
  });
  return 0;
}

}  // namespace fmt
}  // namespace clairb
