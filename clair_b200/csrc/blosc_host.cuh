// Blosc1 frame reader (host code, no device work) for the reference's training / evaluation bins: every block of 500 sites
// is `blosc.pack_array(array, cname='lz4hc', clevel=9, shuffle=NOSHUFFLE)` (clair/utils.py:47-48, 184-220), i.e. a pickled
// numpy array inside one Blosc1 frame.  python-blosc 1.8.3 / c-blosc 1.x are not vendored with the reference and absent
// here, so this follows the published container and LZ4 block formats:
//   header (16 bytes): version, versionlz, flags, typesize, nbytes u32le, blocksize u32le, cbytes u32le
//     flags: 0x01 byte shuffle, 0x02 memcpyed (payload stored as is), 0x04 bit shuffle, 0x10 blocks are not split,
//            bits 5..7 codec format (0 blosclz, 1 lz4 / lz4hc, 2 snappy, 3 zlib, 4 zstd)
//   then one i32le start offset per block, then the blocks; a block is `typesize` separately compressed streams when it is
//   split (typesize <= 16, blocksize / typesize >= 128, not the short last block, flag 0x10 clear), else one stream; each
//   stream = i32le compressed size + payload, stored verbatim when the size equals the stream's uncompressed size.
// Only the LZ4 codec (what the reference writes) and stored data are decoded; shuffled frames are un-shuffled.
#pragma once
#include <stdint.h>

#include <cstring>
#include <vector>

namespace clairb {
namespace blosc {

enum { OK = 0, TRUNCATED = 1, UNSUPPORTED = 2, CORRUPT = 3, CAPACITY = 4 };
constexpr int MAX_SPLITS = 16, MIN_BUFFERSIZE = 128, HEADER = 16;

inline uint32_t u32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct Info {
  int version, versionlz, flags, typesize;
  int64_t nbytes, blocksize, cbytes;
};

inline int info(const uint8_t* src, int64_t len, Info* out) {
  if (len < HEADER) return TRUNCATED;
  out->version = src[0];
  out->versionlz = src[1];
  out->flags = src[2];
  out->typesize = src[3];
  out->nbytes = u32le(src + 4);
  out->blocksize = u32le(src + 8);
  out->cbytes = u32le(src + 12);
  return OK;
}

// One LZ4 block (sequences of token, literals, 2-byte offset, match); exact output size known.  Bounds-checked.
inline int lz4_block(const uint8_t* src, int64_t n, uint8_t* dst, int64_t want) {
  int64_t i = 0, o = 0;
  while (i < n) {
    const int token = src[i++];
    int64_t lit = token >> 4;
    if (lit == 15) {
      int b;
      do {
        if (i >= n) return CORRUPT;
        b = src[i++];
        lit += b;
      } while (b == 255);
    }
    if (i + lit > n || o + lit > want) return CORRUPT;
    memcpy(dst + o, src + i, (size_t)lit);
    i += lit;
    o += lit;
    if (i >= n) break;                                   // the last sequence holds literals only
    if (i + 2 > n) return CORRUPT;
    const int64_t off = src[i] | (src[i + 1] << 8);
    i += 2;
    int64_t mlen = (token & 15) + 4;
    if ((token & 15) == 15) {
      int b;
      do {
        if (i >= n) return CORRUPT;
        b = src[i++];
        mlen += b;
      } while (b == 255);
    }
    if (off == 0 || off > o || o + mlen > want) return CORRUPT;
    for (int64_t k = 0; k < mlen; ++k) dst[o + k] = dst[o + k - off];     // may overlap: byte by byte
    o += mlen;
  }
  return o == want ? OK : CORRUPT;
}

inline void unshuffle(const uint8_t* src, uint8_t* dst, int64_t n, int typesize) {
  const int64_t items = n / typesize;
  for (int64_t j = 0; j < typesize; ++j)
    for (int64_t i = 0; i < items; ++i) dst[i * typesize + j] = src[j * items + i];
  memcpy(dst + items * typesize, src + items * typesize, (size_t)(n - items * typesize));
}

inline int decompress(const uint8_t* src, int64_t len, uint8_t* dst, int64_t cap, int64_t* nbytes) {
  Info h;
  if (int rc = info(src, len, &h)) return rc;
  *nbytes = h.nbytes;
  if (h.cbytes > len) return TRUNCATED;
  if (h.nbytes > cap) return CAPACITY;
  if (h.nbytes == 0) return OK;
  if (h.flags & 0x2) {                                   // memcpyed
    if (HEADER + h.nbytes > h.cbytes) return CORRUPT;
    memcpy(dst, src + HEADER, (size_t)h.nbytes);
    return OK;
  }
  if (h.flags & 0x4) return UNSUPPORTED;                 // bit shuffle
  if ((h.flags >> 5) != 1) return UNSUPPORTED;           // not LZ4
  if (h.blocksize <= 0 || h.typesize <= 0) return CORRUPT;
  const int64_t nblocks = (h.nbytes + h.blocksize - 1) / h.blocksize;
  if (HEADER + 4 * nblocks > h.cbytes) return CORRUPT;
  const bool dont_split = h.flags & 0x10, shuffled = h.flags & 0x1;
  std::vector<uint8_t> tmp;
  if (shuffled) tmp.resize((size_t)h.blocksize);
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t bsize = b == nblocks - 1 && h.nbytes % h.blocksize ? h.nbytes % h.blocksize : h.blocksize;
    const bool leftover = bsize != h.blocksize;
    int64_t at = u32le(src + HEADER + 4 * b);
    if (at < HEADER + 4 * nblocks || at > h.cbytes) return CORRUPT;
    const int nsplits = (!dont_split && h.typesize <= MAX_SPLITS && h.blocksize / h.typesize >= MIN_BUFFERSIZE && !leftover) ? h.typesize : 1;
    const int64_t neblock = bsize / nsplits;
    uint8_t* out = shuffled ? tmp.data() : dst + b * h.blocksize;
    for (int s = 0; s < nsplits; ++s) {
      if (at + 4 > h.cbytes) return CORRUPT;
      const int64_t c = (int32_t)u32le(src + at);
      at += 4;
      if (c < 0 || at + c > h.cbytes) return CORRUPT;
      if (c == neblock) memcpy(out + s * neblock, src + at, (size_t)c);
      else if (int rc = lz4_block(src + at, c, out + s * neblock, neblock)) return rc;
      at += c;
    }
    if (shuffled) unshuffle(tmp.data(), dst + b * h.blocksize, bsize, h.typesize);
  }
  return OK;
}

}  // namespace blosc
}  // namespace clairb
