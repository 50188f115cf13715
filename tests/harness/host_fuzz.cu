// TEST INFRASTRUCTURE.  The host-side decoders of the C-ABI library (Blosc frames, SAM rows, row formatter) under
// AddressSanitizer + UBSan on random and damaged input: any read or write outside a buffer aborts the run
// (tests/test_host_fuzz.py builds and runs this; exact-size destination buffers so that an overrun is visible).
#include "../../clair_b200/csrc/blosc_host.cuh"
#include "../../clair_b200/csrc/encode_host.cuh"
#include "../../clair_b200/csrc/format_host.cuh"
#include "../../clair_b200/csrc/decode_host.cuh"
#include <cstdio>
#include <random>
#include <string>
using namespace clairb;
int main() {
  std::mt19937 rng(7);
  // 1) blosc: a hand-made valid frame (one block, literals + overlapping match), then random damage
  std::vector<uint8_t> block = {0x4F, 'a', 'b', 'c', 'd', 4, 0, (uint8_t)(196 - 4 - 15), 0x50, 'v', 'w', 'x', 'y', 'z'};
  std::vector<uint8_t> frame(16 + 4 + 4 + block.size());
  uint32_t nbytes = 205, cbytes = (uint32_t)frame.size();
  frame[0] = 2; frame[1] = 1; frame[2] = 0x10 | (1 << 5); frame[3] = 1;
  memcpy(&frame[4], &nbytes, 4); memcpy(&frame[8], &nbytes, 4); memcpy(&frame[12], &cbytes, 4);
  uint32_t start = 20; memcpy(&frame[16], &start, 4);
  int32_t c = (int32_t)block.size(); memcpy(&frame[20], &c, 4);
  memcpy(&frame[24], block.data(), block.size());
  std::vector<uint8_t> out(4096);
  int64_t n = 0;
  int rc = blosc::decompress(frame.data(), frame.size(), out.data(), 205, &n);
  printf("valid frame rc=%d n=%lld first=%c last=%c\n", rc, (long long)n, out[0], out[204]);
  int ok = 0, err = 0;
  for (int it = 0; it < 60000; ++it) {
    std::vector<uint8_t> d(frame);
    int k = 1 + rng() % 4;
    for (int j = 0; j < k; ++j) d[rng() % d.size()] = (uint8_t)rng();
    if (rng() % 5 == 0) d.resize(rng() % d.size());
    std::vector<uint8_t> dst(256);                       // exact-size destination buffers so ASan sees any overrun
    int64_t m = 0;
    int r = blosc::decompress(d.data(), (int64_t)d.size(), dst.data(), (int64_t)dst.size(), &m);
    (r == 0 ? ok : err)++;
  }
  printf("blosc fuzz: ok=%d err=%d\n", ok, err);
  // 2) SAM encoder: random rows
  const char* ops = "MIDNSHP=X*B";
  ok = err = 0;
  for (int it = 0; it < 6000; ++it) {
    std::string text;
    int rows = 1 + rng() % 4, pos = 1;
    for (int r = 0; r < rows; ++r) {
      pos += rng() % 3;
      std::string cigar;
      int nops = rng() % 6;
      for (int j = 0; j < nops; ++j) { cigar += std::to_string(rng() % 30); cigar += ops[rng() % 11]; }
      if (cigar.empty()) cigar = "*";
      std::string seq(rng() % 40, 'A');
      if (seq.empty()) seq = "*";
      text += "r\t" + std::to_string(rng() % 4096) + "\tc\t" + std::to_string(pos) + "\t" + std::to_string(rng() % 61) + "\t" + cigar + "\t*\t0\t0\t" + seq + "\t*";
      if (rng() % 7 == 0) text.resize(rng() % (text.size() + 1));      // truncate somewhere
      text += (rng() % 9 == 0) ? "\r\n" : "\n";
    }
    int32_t state[3] = {0, 0, INT32_MIN};
    int64_t R = 0, O = 0, B = 0, bad = -1;
    int r1 = sam::encode(text.data(), (int64_t)text.size(), 0, 250, state, nullptr, 0, 0, 0, &R, &O, &B, &bad, 2);
    if (r1) { ++err; continue; }
    std::vector<int32_t> rp(R), re(R), ro(R + 1), a1(O), a2(O), a3(O);
    std::vector<uint8_t> rs(R), sq(B);
    sam::Out o{rp.data(), re.data(), ro.data(), rs.data(), a1.data(), a2.data(), a3.data(), sq.data()};
    int r2 = sam::encode(text.data(), (int64_t)text.size(), 0, 250, state, &o, R, O, B, &R, &O, &B, &bad, 2);
    (r2 == 0 ? ok : err)++;
  }
  printf("sam fuzz: ok=%d err=%d\n", ok, err);
  // 2b) tensor-row decoder (clairb_decode_rows): valid rows with random damage and truncation
  ok = err = 0;
  for (int it = 0; it < 3000; ++it) {
    std::string text;
    int rows = 1 + rng() % 3;
    for (int r = 0; r < rows; ++r) {
      text += "chr1 " + std::to_string(rng() % 100000) + " ACGTACGTACGTACGTNACGTACGTACGTACGTA";
      for (int k = 0; k < 1056; ++k) text += " " + std::to_string((int)(rng() % 200) - 20);
      text += "\n";
    }
    int k = rng() % 4;
    for (int j = 0; j < k; ++j) text[rng() % text.size()] = " \t-9aN.\n"[rng() % 9];
    if (rng() % 4 == 0) text.resize(rng() % text.size());
    std::vector<int16_t> xo((size_t)rows * 1056 + 1056);
    std::vector<int32_t> info((size_t)(rows + 2) * 6);
    int64_t rr = 0, rk = 0, cons = 0, bad = -1;
    int r = decode::rows<int16_t>(text.data(), (int64_t)text.size(), rows, xo.data(), info.data(), &rr, &rk, &cons, &bad, 2);
    (r == 0 ? ok : err)++;
  }
  printf("decode fuzz: ok=%d err=%d\n", ok, err);
  // 3) formatter at the edges of the reference text
  const char* ref = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT";
  int64_t positions[3] = {17, 30, 1000000000}, starts[3] = {0, 20, 39};
  std::vector<int16_t> x(3 * 1056, (int16_t)-32768);
  int64_t need = 0;
  fmt::rows("chr", positions, ref, 40, starts, x.data(), 3, nullptr, 0, &need, 2);
  std::vector<char> buf(need);
  int r3 = fmt::rows("chr", positions, ref, 40, starts, x.data(), 3, buf.data(), need, &need, 2);
  printf("format rc=%d bytes=%lld\n", r3, (long long)need);
  return 0;
}
