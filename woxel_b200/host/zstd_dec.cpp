// Zstandard decoder for Blosc frames (c-blosc codec 4: every stream of a block is one frame written by ZSTD_compress).
// The reference hands Blosc frames to the c-blosc C library (src/vdb/read.rs:514-533), which bundles libzstd; neither is in this
// image, so this is a decoder written from the published format (RFC 8878): frames (magic, header descriptor, window, dictionary id,
// content size, optional checksum), raw / RLE / compressed blocks, the literals section (raw, RLE, Huffman-coded with 1 or 4
// streams, tree reuse), Huffman tree descriptions (direct or FSE-coded weights), the sequences section (predefined / RLE / FSE /
// repeated tables, three interleaved FSE states on a backward bit stream), repeat offsets and sequence execution.
// Not supported (reported as corrupt): dictionaries (c-blosc never uses one).  The optional content checksum is skipped, not verified.
// Every read and write is bounds-checked: a corrupt stream is an error, never a crash.
// Checked against libzstd's own output (pyarrow's "zstd" codec, levels 1..19): tests/golden/blosc_frames.npz.
#include <cstdint>
#include <cstring>
#include <vector>

namespace woxel::vdb {
namespace {

struct ZErr {};  // any malformed input
[[noreturn]] void corrupt() { throw ZErr{}; }
inline void need(bool ok) {
  if (!ok) corrupt();
}
inline int high_bit(uint64_t v) {  // position of the highest set bit; v != 0
  return 63 - __builtin_clzll(v);
}

// bits [off, off + nb) of a little-endian bit string, nb <= 57 (callers ask for at most 32); everything must lie inside [0, 8 * len)
inline uint64_t bits_le(const uint8_t* p, size_t len, int nb, int64_t off) {
  if (nb == 0) return 0;
  need(off >= 0 && (uint64_t)off + (uint64_t)nb <= 8ull * len);
  const size_t byte = (size_t)off >> 3;
  const int sh = (int)(off & 7);
  uint64_t v = 0;
  const size_t take = (size_t)((sh + nb + 7) >> 3);
  for (size_t k = 0; k < take; ++k) v |= (uint64_t)p[byte + k] << (8 * k);
  return (v >> sh) & ((1ull << nb) - 1ull);
}

struct FwdBits {  // forward reader (FSE table descriptions)
  const uint8_t* p;
  size_t len;
  int64_t pos = 0;
  uint64_t read(int nb) {
    const uint64_t v = bits_le(p, len, nb, pos);
    pos += nb;
    return v;
  }
  void rewind(int nb) { pos -= nb; }
  size_t bytes_used() const { return (size_t)((pos + 7) >> 3); }
};

struct BackBits {  // backward reader: the stream ends with a 1 bit below the zero padding of its last byte
  const uint8_t* p;
  size_t len;
  int64_t off;
  BackBits(const uint8_t* src, size_t n) : p(src), len(n) {
    need(n > 0 && src[n - 1] != 0);
    off = (int64_t)n * 8 - (8 - high_bit(src[n - 1]));
  }
  // nb bits above the new position; bits before the start of the stream read as zeros (the format relies on it at the very end)
  uint64_t read(int nb) {
    off -= nb;
    if (off >= 0) return bits_le(p, len, nb, off);
    const int64_t under = -off;
    if (under >= nb) return 0;
    return bits_le(p, len, nb - (int)under, 0) << under;
  }
};

// ---- FSE ------------------------------------------------------------------------------------------------------------
struct FseTable {
  int log = 0;
  std::vector<uint8_t> sym, nbits;
  std::vector<uint16_t> base;
  bool valid = false;
};

void fse_build(FseTable& t, const int16_t* freq, int nsym, int log) {
  const int size = 1 << log;
  t.log = log;
  t.sym.assign(size, 0), t.nbits.assign(size, 0), t.base.assign(size, 0);
  std::vector<uint16_t> next(nsym, 0);
  int high = size;
  for (int s = 0; s < nsym; ++s)
    if (freq[s] == -1) {
      need(high > 0);
      t.sym[--high] = (uint8_t)s, next[s] = 1;
    }
  const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; ++s) {
    if (freq[s] <= 0) continue;
    next[s] = (uint16_t)freq[s];
    for (int i = 0; i < freq[s]; ++i) {
      t.sym[pos] = (uint8_t)s;
      do pos = (pos + step) & mask;
      while (pos >= high);
    }
  }
  need(pos == 0);
  for (int i = 0; i < size; ++i) {
    const uint16_t d = next[t.sym[i]]++;
    need(d != 0);
    t.nbits[i] = (uint8_t)(log - high_bit(d));
    t.base[i] = (uint16_t)(((uint32_t)d << t.nbits[i]) - (uint32_t)size);
  }
  t.valid = true;
}

void fse_rle(FseTable& t, uint8_t symbol) {
  t.log = 0;
  t.sym.assign(1, symbol), t.nbits.assign(1, 0), t.base.assign(1, 0);
  t.valid = true;
}

// Normalised counts of an FSE table description; returns the bytes it occupies.
size_t fse_read_description(FseTable& t, const uint8_t* src, size_t n, int max_log, int max_sym) {
  FwdBits b{src, n};
  const int log = 5 + (int)b.read(4);
  need(log <= max_log);
  int remaining = 1 << log, s = 0;
  int16_t freq[256];
  while (remaining > 0 && s <= max_sym) {
    const int nb = high_bit((uint64_t)remaining + 1) + 1;
    int val = (int)b.read(nb);
    const int lower = (1 << (nb - 1)) - 1, threshold = (1 << nb) - 1 - (remaining + 1);
    if ((val & lower) < threshold) {
      b.rewind(1);
      val &= lower;
    } else if (val > lower) {
      val -= threshold;
    }
    const int proba = val - 1;
    remaining -= proba < 0 ? -proba : proba;
    freq[s++] = (int16_t)proba;
    if (proba == 0) {
      int repeat = (int)b.read(2);
      for (;;) {
        for (int i = 0; i < repeat && s <= max_sym; ++i) freq[s++] = 0;
        if (repeat != 3) break;
        repeat = (int)b.read(2);
      }
    }
  }
  need(remaining == 0 && s <= max_sym + 1);
  fse_build(t, freq, s, log);
  return b.bytes_used();
}

struct FseState {
  const FseTable* t;
  uint32_t s = 0;
  void init(const FseTable& table, BackBits& b) {
    t = &table;
    s = (uint32_t)b.read(table.log);
  }
  uint8_t peek() const { return t->sym[s]; }
  void update(BackBits& b) { s = (uint32_t)t->base[s] + (uint32_t)b.read(t->nbits[s]); }
};

// ---- Huffman --------------------------------------------------------------------------------------------------------
struct HufTable {
  int max_bits = 0;
  std::vector<uint8_t> sym, nbits;
  bool valid = false;
};

void huf_build(HufTable& h, const uint8_t* weights, int nw) {  // nw weights, the last symbol's is implied
  uint32_t sum = 0;
  for (int i = 0; i < nw; ++i) {
    need(weights[i] <= 11);
    if (weights[i]) sum += 1u << (weights[i] - 1);
  }
  need(sum != 0);
  const int max_bits = high_bit(sum) + 1;
  need(max_bits <= 11);
  const uint32_t left = (1u << max_bits) - sum;
  need((left & (left - 1)) == 0);  // a power of two
  uint8_t bits[256];
  need(nw < 256);
  for (int i = 0; i < nw; ++i) bits[i] = weights[i] ? (uint8_t)(max_bits + 1 - weights[i]) : 0;
  bits[nw] = (uint8_t)(max_bits + 1 - (high_bit(left) + 1));
  const int nsym = nw + 1;
  uint32_t rank_count[13] = {0}, rank_idx[13] = {0};
  for (int i = 0; i < nsym; ++i) rank_count[bits[i]]++;
  const int size = 1 << max_bits;
  h.max_bits = max_bits;
  h.sym.assign(size, 0), h.nbits.assign(size, 0);
  rank_idx[max_bits] = 0;
  for (int i = max_bits; i >= 1; --i) {
    rank_idx[i - 1] = rank_idx[i] + rank_count[i] * (1u << (max_bits - i));
    need(rank_idx[i - 1] <= (uint32_t)size);
    memset(h.nbits.data() + rank_idx[i], i, rank_idx[i - 1] - rank_idx[i]);
  }
  need(rank_idx[0] == (uint32_t)size);
  for (int i = 0; i < nsym; ++i) {
    if (!bits[i]) continue;
    const uint32_t code = rank_idx[bits[i]], len = 1u << (max_bits - bits[i]);
    need(code + len <= (uint32_t)size);
    memset(h.sym.data() + code, i, len);
    rank_idx[bits[i]] += len;
  }
  h.valid = true;
}

// Huffman tree description; returns the bytes it occupies.
size_t huf_read_tree(HufTable& h, const uint8_t* src, size_t n) {
  need(n >= 1);
  const int header = src[0];
  uint8_t weights[256];
  int nw = 0;
  size_t used;
  if (header >= 128) {  // 4-bit weights, high nibble first
    nw = header - 127;
    const size_t bytes = (size_t)(nw + 1) / 2;
    need(1 + bytes <= n);
    for (int i = 0; i < nw; ++i) weights[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    used = 1 + bytes;
  } else {  // FSE-coded weights, two interleaved states
    need(header >= 1 && (size_t)header + 1 <= n);
    const uint8_t* p = src + 1;
    FseTable t;
    const size_t d = fse_read_description(t, p, header, 7, 255);  // (the format allows 6; weights above 11 are refused by huf_build)
    need(d < (size_t)header);
    BackBits b(p + d, header - d);
    FseState s1, s2;
    s1.init(t, b), s2.init(t, b);
    for (;;) {
      need(nw < 255);
      weights[nw++] = s1.peek();
      s1.update(b);
      if (b.off < 0) {
        need(nw < 255);
        weights[nw++] = s2.peek();
        break;
      }
      need(nw < 255);
      weights[nw++] = s2.peek();
      s2.update(b);
      if (b.off < 0) {
        need(nw < 255);
        weights[nw++] = s1.peek();
        break;
      }
    }
    used = 1 + (size_t)header;
  }
  huf_build(h, weights, nw);
  return used;
}

void huf_decode_stream(const HufTable& h, const uint8_t* src, size_t n, uint8_t* dst, size_t count) {
  BackBits b(src, n);
  const uint32_t mask = (1u << h.max_bits) - 1u;
  uint32_t state = (uint32_t)b.read(h.max_bits);
  size_t o = 0;
  while (b.off > -(int64_t)h.max_bits) {
    need(o < count);
    dst[o++] = h.sym[state];
    const int nb = h.nbits[state];
    state = ((state << nb) + (uint32_t)b.read(nb)) & mask;
  }
  need(b.off == -(int64_t)h.max_bits && o == count);
}

// ---- frame ----------------------------------------------------------------------------------------------------------
const int16_t kLLDefault[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
const int16_t kMLDefault[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
const int16_t kOFDefault[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
const uint32_t kLLBase[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
const uint8_t kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const uint32_t kMLBase[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41,
                              43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
const uint8_t kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
                             2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

struct FrameState {  // what persists from block to block of one frame
  HufTable huf;
  FseTable ll, of, ml;
  uint64_t rep[3] = {1, 4, 8};
};

struct Out {
  uint8_t* dst;
  size_t cap, pos, frame_start;
  void put(const uint8_t* s, size_t n) {
    need(n <= cap - pos);
    if (n) memcpy(dst + pos, s, n);
    pos += n;
  }
  void fill(uint8_t v, size_t n) {
    need(n <= cap - pos);
    if (n) memset(dst + pos, v, n);
    pos += n;
  }
  void match(uint64_t offset, size_t n) {
    need(offset != 0 && offset <= pos - frame_start && n <= cap - pos);
    const uint8_t* s = dst + pos - offset;
    uint8_t* d = dst + pos;
    for (size_t k = 0; k < n; ++k) d[k] = s[k];  // byte by byte: a match may overlap what it writes
    pos += n;
  }
};

constexpr size_t kBlockMax = 128 * 1024;

// Literals section: fills `lit`; returns the bytes of the section.
size_t read_literals(FrameState& fs, const uint8_t* src, size_t n, std::vector<uint8_t>& lit) {
  need(n >= 1);
  const int type = src[0] & 3, fmt = (src[0] >> 2) & 3;
  if (type <= 1) {  // raw / RLE
    size_t hdr, size;
    if ((fmt & 1) == 0) hdr = 1, size = src[0] >> 3;
    else if (fmt == 1) need(n >= 2), hdr = 2, size = (src[0] >> 4) | ((size_t)src[1] << 4);
    else need(n >= 3), hdr = 3, size = (src[0] >> 4) | ((size_t)src[1] << 4) | ((size_t)src[2] << 12);
    need(size <= kBlockMax);
    if (type == 0) {
      need(hdr + size <= n);
      lit.assign(src + hdr, src + hdr + size);
      return hdr + size;
    }
    need(hdr + 1 <= n);
    lit.assign(size, src[hdr]);
    return hdr + 1;
  }
  size_t hdr, regen, comp;
  int streams;
  if (fmt <= 1) {
    need(n >= 3);
    const uint32_t h = src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
    hdr = 3, streams = fmt == 0 ? 1 : 4, regen = (h >> 4) & 0x3ff, comp = (h >> 14) & 0x3ff;
  } else if (fmt == 2) {
    need(n >= 4);
    const uint32_t h = src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    hdr = 4, streams = 4, regen = (h >> 4) & 0x3fff, comp = h >> 18;
  } else {
    need(n >= 5);
    const uint64_t h = src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
    hdr = 5, streams = 4, regen = (size_t)((h >> 4) & 0x3ffff), comp = (size_t)((h >> 22) & 0x3ffff);
  }
  need(regen <= kBlockMax && hdr + comp <= n);
  const uint8_t* p = src + hdr;
  size_t left = comp;
  if (type == 2) {
    const size_t used = huf_read_tree(fs.huf, p, left);
    p += used, left -= used;
  } else {
    need(fs.huf.valid);  // treeless: the previous block's tree
  }
  lit.resize(regen);
  if (streams == 1) {
    huf_decode_stream(fs.huf, p, left, lit.data(), regen);
  } else {
    need(left >= 6);
    const size_t s1 = p[0] | ((size_t)p[1] << 8), s2 = p[2] | ((size_t)p[3] << 8), s3 = p[4] | ((size_t)p[5] << 8);
    need(6 + s1 + s2 + s3 < left && s1 && s2 && s3);
    const size_t s4 = left - 6 - s1 - s2 - s3;
    const size_t seg = (regen + 3) / 4;
    need(3 * seg <= regen);
    const uint8_t* q = p + 6;
    huf_decode_stream(fs.huf, q, s1, lit.data(), seg);
    huf_decode_stream(fs.huf, q + s1, s2, lit.data() + seg, seg);
    huf_decode_stream(fs.huf, q + s1 + s2, s3, lit.data() + 2 * seg, seg);
    huf_decode_stream(fs.huf, q + s1 + s2 + s3, s4, lit.data() + 3 * seg, regen - 3 * seg);
  }
  return hdr + comp;
}

// One table of the sequences section according to its compression mode; advances p.
void read_seq_table(FseTable& t, int mode, const uint8_t*& p, size_t& left, const int16_t* def, int def_n, int def_log, int max_log, int max_sym) {
  switch (mode) {
    case 0: fse_build(t, def, def_n, def_log); break;
    case 1:
      need(left >= 1 && p[0] <= max_sym);
      fse_rle(t, p[0]);
      p += 1, left -= 1;
      break;
    case 2: {
      const size_t used = fse_read_description(t, p, left, max_log, max_sym);
      need(used <= left);
      p += used, left -= used;
      break;
    }
    default: need(t.valid);  // repeat
  }
}

void decode_block(FrameState& fs, const uint8_t* src, size_t n, Out& out) {
  std::vector<uint8_t> lit;
  const size_t lsz = read_literals(fs, src, n, lit);
  const uint8_t* p = src + lsz;
  size_t left = n - lsz;
  need(left >= 1);
  size_t nseq = p[0];
  if (nseq == 0) {
    need(left == 1);
    out.put(lit.data(), lit.size());
    return;
  }
  if (nseq < 128) p += 1, left -= 1;
  else if (nseq < 255) need(left >= 2), nseq = ((nseq - 128) << 8) + p[1], p += 2, left -= 2;
  else need(left >= 3), nseq = (size_t)p[1] + ((size_t)p[2] << 8) + 0x7f00, p += 3, left -= 3;
  need(left >= 1);
  const int modes = p[0];
  need((modes & 3) == 0);
  p += 1, left -= 1;
  read_seq_table(fs.ll, modes >> 6, p, left, kLLDefault, 36, 6, 9, 35);
  read_seq_table(fs.of, (modes >> 4) & 3, p, left, kOFDefault, 29, 5, 8, 31);
  read_seq_table(fs.ml, (modes >> 2) & 3, p, left, kMLDefault, 53, 6, 9, 52);
  BackBits b(p, left);
  FseState ll, of, ml;
  ll.init(fs.ll, b), of.init(fs.of, b), ml.init(fs.ml, b);
  size_t lpos = 0;
  for (size_t k = 0; k < nseq; ++k) {
    const int ofc = of.peek(), llc = ll.peek(), mlc = ml.peek();
    need(ofc <= 31 && llc <= 35 && mlc <= 52);
    const uint64_t ofv = (1ull << ofc) + b.read(ofc);
    const size_t mlen = kMLBase[mlc] + (size_t)b.read(kMLBits[mlc]);
    const size_t llen = kLLBase[llc] + (size_t)b.read(kLLBits[llc]);
    need(b.off >= 0);
    if (k + 1 < nseq) ll.update(b), ml.update(b), of.update(b);
    uint64_t offset;
    if (ofv > 3) {
      offset = ofv - 3;
      fs.rep[2] = fs.rep[1], fs.rep[1] = fs.rep[0], fs.rep[0] = offset;
    } else {
      uint32_t idx = (uint32_t)ofv - 1u + (llen == 0 ? 1u : 0u);
      if (idx == 0) {
        offset = fs.rep[0];
      } else {
        offset = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
        need(offset != 0);
        if (idx > 1) fs.rep[2] = fs.rep[1];
        fs.rep[1] = fs.rep[0], fs.rep[0] = offset;
      }
    }
    need(llen <= lit.size() - lpos);
    out.put(lit.data() + lpos, llen);
    lpos += llen;
    out.match(offset, mlen);
  }
  need(b.off == 0);
  out.put(lit.data() + lpos, lit.size() - lpos);
}

// One frame starting at src; returns the bytes it occupies.
size_t decode_frame(const uint8_t* src, size_t n, Out& out) {
  need(n >= 4);
  const uint32_t magic = src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
  if ((magic & 0xfffffff0u) == 0x184d2a50u) {  // skippable frame
    need(n >= 8);
    const size_t sz = src[4] | ((size_t)src[5] << 8) | ((size_t)src[6] << 16) | ((size_t)src[7] << 24);
    need(sz <= n - 8);
    return 8 + sz;
  }
  need(magic == 0xfd2fb528u && n >= 5);
  const int fhd = src[4];
  const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did_flag = fhd & 3;
  need((fhd & 0x08) == 0);
  need(did_flag == 0);  // a dictionary is needed: not supported
  size_t at = 5;
  if (!single) at += 1;  // window descriptor: the output buffer is the window
  const size_t fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
  need(at + fcs_bytes <= n);
  uint64_t fcs = 0;
  for (size_t k = 0; k < fcs_bytes; ++k) fcs |= (uint64_t)src[at + k] << (8 * k);
  if (fcs_bytes == 2) fcs += 256;
  at += fcs_bytes;
  if (fcs_bytes) need(fcs <= out.cap - out.pos);
  out.frame_start = out.pos;
  FrameState fs;
  for (;;) {
    need(at + 3 <= n);
    const uint32_t bh = src[at] | ((uint32_t)src[at + 1] << 8) | ((uint32_t)src[at + 2] << 16);
    at += 3;
    const int last = bh & 1, type = (bh >> 1) & 3;
    const size_t bsize = bh >> 3;
    need(bsize <= kBlockMax);
    if (type == 0) {
      need(bsize <= n - at);
      out.put(src + at, bsize);
      at += bsize;
    } else if (type == 1) {
      need(at < n);
      out.fill(src[at], bsize);
      at += 1;
    } else if (type == 2) {
      need(bsize <= n - at);
      const size_t before = out.pos;
      decode_block(fs, src + at, bsize, out);
      need(out.pos - before <= kBlockMax);
      at += bsize;
    } else {
      corrupt();
    }
    if (last) break;
  }
  if (fcs_bytes) need(out.pos - out.frame_start == fcs);
  if (checksum) {
    need(at + 4 <= n);
    at += 4;  // xxh64 of the content, low 32 bits: not verified
  }
  return at;
}

}  // namespace

// All frames of [src, src + n) into dst (exactly cap bytes expected); false when the stream is malformed or decodes to another size.
bool zstd_decompress_block(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
  try {
    Out out{dst, cap, 0, 0};
    size_t at = 0;
    need(n > 0);
    while (at < n) at += decode_frame(src + at, n - at, out);
    return out.pos == cap;
  } catch (const ZErr&) {
    return false;
  }
}

}  // namespace woxel::vdb
