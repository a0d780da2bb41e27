// inflate.hpp — raw DEFLATE (RFC 1951) / gzip (RFC 1952) decoder of the file drivers (fastq.cpp: .fastq.gz streams;
// bam.cpp: BGZF blocks).  The reference reads both through flate2 / htslib (src/parse/fastq.rs:21-43,
// src/parse/sorted_bam_reader.rs:22-41); here inflating is the first stage of the host pipeline that feeds the device,
// and zlib's byte-at-a-time state machine (0.2–0.4 GB/s per thread on read data) was the bound of both drivers.
//
// Shape of the decoder (the well-known fast-inflate recipe, written for this input model):
//   * the whole compressed input is in memory (a mapped file, or one BGZF block), so the decoder never suspends inside a
//     symbol for want of input — running out of input is an error, not a state;
//   * a 64-bit bit buffer refilled with one unaligned 8-byte load, at most twice per literal run + match;
//   * one table lookup per symbol: 11-bit primary table for literal/length codes, 8-bit for distance codes, second-level
//     tables for the longer codes; an entry carries the base value, the number of extra bits and the bits to consume;
//   * literal entries of the primary table carry up to THREE literals when their codes fit into its 11 bits together
//     (read data is four-letter sequence and a handful of quality characters: codes of two or three bits), so a run of
//     literals costs one lookup per two or three bytes instead of one per byte;
//   * matches copied eight bytes at a time;
//   * a careful loop (byte-wise refill, every access bounds-checked, symbol-exact stop) near the end of the input or of the
//     output window, so that the output can be produced in pieces: run() stops BETWEEN symbols when the next one does not
//     fit and continues in the next window, which must begin with the previous 32 KiB of output.
// Accepts what zlib accepts: over-subscribed or incomplete code sets are errors, except a literal/distance set made of a
// single one-bit code and an empty distance set (a block of literals only); symbols 286/287 and distances 30/31 are
// errors when they are decoded.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>   // crc32_z only

namespace nbz {

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;

enum { INF_MORE = 0, INF_END = 1, INF_STOP = 2, INF_ERROR = -1 };

// base values and extra bits of the length symbols 257..285 and the distance symbols 0..29 (RFC 1951 3.2.5)
static const u16 LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const u8 LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const u16 DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const u8 DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

class Inflater {
 public:
  // [in, in_end): the whole remaining compressed input
  void start(const u8* in, const u8* in_end) { origin_ = in_ = in; end_ = in_end; bb_ = 0; bc_ = 0; in_block_ = false; final_ = false; stored_left_ = 0; btype_ = 0; stop_bit_ = ~0ull; stop_out_ = ~(size_t)0; }
  // ---- for the parallel reader (pgunzip.hpp): decoding that starts at a block header at any BIT of `data`, stops in front
  // of the first block header at or behind bit `stop` (run / run16 return INF_STOP there), and reports where it is
  void start_at_bit(const u8* data, const u8* in_end, u64 bit, u64 stop, size_t stop_out = ~(size_t)0) {   // stop_out: also stop in front of the first block header behind that many output elements
    start(data + (bit >> 3), in_end); origin_ = data; stop_bit_ = stop; stop_out_ = stop_out;
    if (bit & 7) { if (need(8)) take((u32)(bit & 7)); else in_ = end_; }
  }
  u64 bit_pos() const { return (u64)(in_ - origin_) * 8 - bc_; }       // of the next unread bit (meaningful between blocks)
  bool header_at_bit(const u8* data, const u8* in_end, u64 bit) { start_at_bit(data, in_end, bit, ~0ull); return block_header() && btype_ == 2 && !final_; }   // is there a sound non-final dynamic block header?
  // The same decoding into 16-bit symbols, for a stream entered in the middle, where the 32 KiB before the entry point are
  // unknown: [out - 32768, out) must exist and hold the values 0x8000 + i (i = 0..32767) at the first call, so that a match
  // reaching back behind the entry point copies PLACEHOLDERS for those bytes like any other symbol; whoever knows the real
  // window replaces every value >= 0x8000 by window[value - 0x8000] afterwards.  base: the start of that placeholder window.
  inline int run16(const u16* base, u16*& out, u16* out_end);
  // Decodes into [out, out_end); [base, out) must hold the output so far (at least its last 32 KiB).  Returns INF_END at
  // the end of the final block (out = end of the data; next_byte() = the first byte behind the deflate stream), INF_MORE
  // when the next symbol does not fit into the window (out = what was produced; call again with a new window),
  // INF_ERROR on a damaged or truncated stream.
  inline int run(const u8* base, u8*& out, u8* out_end);
  const u8* next_byte() const { return in_ - (bc_ >> 3); }

 private:
  static const int LP = 11, DP = 8;                   // primary table bits
  static const int LCAP = 2400, DCAP = 420;           // >= the largest table any code set of 288 / 32 symbols needs (2342 / 402)
  static const u32 L_LIT = 1u << 31, L_EXC = 1u << 30, L_SUB = 1u << 29, L_EOB = 1u << 28;   // EXC alone: invalid code
  static const u32 D_SUB = 1u << 31, D_BAD = 1u << 15;
  // entry layouts (bits):   literal  31 LIT | 30-24 third | 23-16 second | 15-8 first | 5-4 literals - 1 | 3-0 bits to consume
  //                         length   24-16 base | 12-8 extra bits | 3-0 consume      second level pointer  27-16 start | 11-8 index bits | 3-0 consume
  //                         distance 30-16 base | 11-8 extra bits | 3-0 consume      (D_SUB: 24-16 start | 11-8 index bits)
  static const ptrdiff_t FAST_IN = 16, FAST_OUT = 258 + 32;   // margins of the fast loop: two refills; nine literals + the longest match + the stores' overrun
  const u8* in_ = nullptr; const u8* end_ = nullptr; u64 bb_ = 0; u32 bc_ = 0;   // invariant between calls: bits of bb_ above bc_ are zero
  const u8* origin_ = nullptr; u64 stop_bit_ = ~0ull; size_t stop_out_ = ~(size_t)0;
  bool in_block_ = false, final_ = false; int btype_ = 0; u32 stored_left_ = 0;
  const u32* lt_ = nullptr; const u32* dt_ = nullptr;
  u32 ltab_[LCAP]; u32 dtab_[DCAP];

  struct Fixed {
    u32 lt[LCAP]; u32 dt[DCAP];
    Fixed() {
      u8 l[288]; for (int i = 0; i < 288; i++) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      u8 d[32]; memset(d, 5, 32);
      build(l, 288, 0, lt, LCAP); build(d, 32, 1, dt, DCAP);
    }
  };
  static const Fixed& fixed() { static const Fixed f; return f; }

  static u32 rev16(u32 c, int len) {
    c = ((c & 0x5555) << 1) | ((c >> 1) & 0x5555); c = ((c & 0x3333) << 2) | ((c >> 2) & 0x3333);
    c = ((c & 0x0F0F) << 4) | ((c >> 4) & 0x0F0F); c = ((c & 0x00FF) << 8) | ((c >> 8) & 0x00FF);
    return c >> (16 - len);
  }
  static u64 load64(const u8* p) { u64 v; memcpy(&v, p, 8); return v; }   // (little-endian host)
  // Canonical Huffman decode table over lens[0..n).  kind 0: literal/length, 1: distance.  false: not a usable code set.
  static inline bool build(const u8* lens, int n, int kind, u32* tab, int cap);
  bool need(u32 n) { while (bc_ < n) { if (in_ >= end_) return false; bb_ |= (u64)*in_++ << bc_; bc_ += 8; } return true; }
  u32 take(u32 n) { u32 v = (u32)(bb_ & ((1ull << n) - 1)); bb_ >>= n; bc_ -= n; return v; }
  bool get(u32 n, u32& v) { if (!need(n)) return false; v = take(n); return true; }
  inline bool block_header();
  template <typename T> inline int huff(const T* base, T*& out, T* out_end);   // 0: block ended, 1: window full, -1: error.  T: u8, or u16 for run16
  static void put_lits(u8* out, u32 w) { memcpy(out, &w, 4); }                   // three literals (or fewer and zeros) packed into w's bytes
  static void put_lits(u16* out, u32 w) { const u64 x = (u64)(w & 0xFF) | ((u64)((w >> 8) & 0xFF) << 16) | ((u64)(w >> 16) << 32); memcpy(out, &x, 8); }
  template <typename T> static void copy_match(T* dst, const T* src, T* end, u32 dist) {       // 8 bytes at a time; may write up to 16 bytes - 1 element past end
    const u32 W = 8 / sizeof(T);
    if (dist >= W) {
      memcpy(dst, src, 8); memcpy(dst + W, src + W, 8);
      if (end - dst > (ptrdiff_t)(2 * W)) { dst += 2 * W; src += 2 * W; do { memcpy(dst, src, 8); dst += W; src += W; } while (dst < end); }
    } else if (dist == 1) { const T v = *src; if (sizeof(T) == 1) memset(dst, (int)v, (size_t)(end - dst)); else do { *dst++ = v; } while (dst < end); }
    else { do { *dst++ = *src++; } while (dst < end); }
  }
};

inline bool Inflater::build(const u8* lens, int n, int kind, u32* tab, int cap) {
  const int P = kind ? DP : LP;
  u16 count[16] = {0}; for (int i = 0; i < n; i++) count[lens[i] & 15]++;
  const int n_codes = n - count[0];
  u32 left = 1u << 15;
  for (int l = 1; l <= 15; l++) { const u32 use = (u32)count[l] << (15 - l); if (use > left) return false; left -= use; }   // over-subscribed
  const bool incomplete = left != 0;
  if (incomplete && !(n_codes == 0 && kind == 1) && !(n_codes == 1 && count[1] == 1)) return false;
  if (incomplete) { const u32 bad = kind ? (D_BAD | 1u) : (L_EXC | 1u); for (int i = 0; i < (1 << P); i++) tab[i] = bad; }   // an unused code: consumes a bit, flagged
  // symbols sorted by (length, value), with their bit-reversed canonical codes (the stream carries codes MSB first)
  u16 sorted[288], rcode[288]; u8 slen[288];
  { u16 at[16]; u16 o = 0; for (int l = 1; l <= 15; l++) { at[l] = o; o = (u16)(o + count[l]); }
    for (int i = 0; i < n; i++) if (lens[i] & 15) sorted[at[lens[i] & 15]++] = (u16)i; }
  { u32 code = 0; int k = 0; for (int l = 1; l <= 15; l++) { for (int c = 0; c < count[l]; c++, k++, code++) { rcode[k] = (u16)rev16(code, l); slen[k] = (u8)l; } code <<= 1; } }
  auto entry = [&](int sym) -> u32 {        // everything but the bits to consume
    if (kind) { if (sym >= 30) return D_BAD; return ((u32)DIST_BASE[sym] << 16) | ((u32)DIST_EXTRA[sym] << 8); }
    if (sym < 256) return L_LIT | ((u32)sym << 8);
    if (sym == 256) return L_EXC | L_EOB;
    if (sym >= 286) return L_EXC;
    return ((u32)LEN_BASE[sym - 257] << 16) | ((u32)LEN_EXTRA[sym - 257] << 8);
  };
  int k = 0;
  for (; k < n_codes && slen[k] <= P; k++) {
    const u32 e = entry(sorted[k]) | slen[k];
    for (u32 i = rcode[k]; i < (1u << P); i += 1u << slen[k]) tab[i] = e;
  }
  // longer codes: one second-level table per P-bit prefix, indexed by the bits behind the prefix.  Codes of one prefix are
  // neighbours in the sorted order with non-decreasing lengths, so the last one of the run gives the table's size.
  int next = 1 << P;
  while (k < n_codes) {
    const u32 prefix = rcode[k] & ((1u << P) - 1);
    int e_ = k; while (e_ + 1 < n_codes && (rcode[e_ + 1] & ((1u << P) - 1)) == prefix) e_++;
    const int bits = slen[e_] - P;
    if (next + (1 << bits) > cap) return false;
    tab[prefix] = (kind ? D_SUB : (L_EXC | L_SUB)) | ((u32)next << 16) | ((u32)bits << 8) | (u32)P;
    for (; k <= e_; k++) {
      const u32 e = entry(sorted[k]) | (u32)(slen[k] - P);
      for (u32 i = rcode[k] >> P; i < (1u << bits); i += 1u << (slen[k] - P)) tab[next + i] = e;
    }
    next += 1 << bits;
  }
  // literal runs: an index whose low bits are a literal's code and whose next bits are ANOTHER literal's whole code decodes both
  // (and a third, below 128, when it fits too).  Entries of codes shorter than P are replicated over all higher index bits, so
  // tab[i >> l1] — the index with the first code shifted out, zeros shifted in — is the next symbol's entry whenever that
  // symbol's code is no longer than the P - l1 bits that are really known.
  if (!kind) {
    u32 tmp[1 << LP];
    for (u32 i = 0; i < (1u << LP); i++) {
      const u32 e1 = tab[i]; tmp[i] = e1;
      if (!(e1 & L_LIT)) continue;
      const u32 l1 = e1 & 15;
      const u32 e2 = tab[i >> l1]; const u32 l2 = e2 & 15;
      if (!(e2 & L_LIT) || l1 + l2 > (u32)LP) continue;
      u32 e = L_LIT | (e1 & 0xFF00) | ((e2 & 0xFF00) << 8) | (1u << 4) | (l1 + l2);
      const u32 e3 = tab[i >> (l1 + l2)]; const u32 l3 = e3 & 15;
      if ((e3 & L_LIT) && l1 + l2 + l3 <= (u32)LP && !(e3 & 0x8000)) e = L_LIT | (e1 & 0xFF00) | ((e2 & 0xFF00) << 8) | ((e3 & 0x7F00) << 16) | (2u << 4) | (l1 + l2 + l3);
      tmp[i] = e;
    }
    memcpy(tab, tmp, sizeof tmp);
  }
  return true;
}

inline bool Inflater::block_header() {
  u32 v;
  if (!get(3, v)) return false;
  final_ = v & 1; btype_ = (int)(v >> 1);
  if (btype_ == 0) {
    take(bc_ & 7);                                   // to the byte boundary
    u32 len, nlen;
    if (!get(16, len) || !get(16, nlen) || (len ^ 0xFFFF) != nlen) return false;
    in_ -= bc_ >> 3; bb_ = 0; bc_ = 0;               // the bit buffer holds whole bytes only: give them back
    stored_left_ = len;
    return true;
  }
  if (btype_ == 1) { lt_ = fixed().lt; dt_ = fixed().dt; return true; }
  if (btype_ != 2) return false;
  u32 hlit, hdist, hclen;
  if (!get(5, hlit) || !get(5, hdist) || !get(4, hclen)) return false;
  hlit += 257; hdist += 1; hclen += 4;
  if (hlit > 286 || hdist > 30) return false;
  static const u8 ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  u8 pl[19] = {0};
  for (u32 i = 0; i < hclen; i++) { if (!get(3, v)) return false; pl[ORDER[i]] = (u8)v; }
  // code-length code: at most 7 bits, one direct table; must be complete (zlib)
  u16 ptab[128];
  { u32 left = 1u << 7; int cnt[8] = {0}; for (int i = 0; i < 19; i++) cnt[pl[i]]++;
    for (int l = 1; l <= 7; l++) { const u32 use = (u32)cnt[l] << (7 - l); if (use > left) return false; left -= use; }
    if (left) return false;
    u32 code = 0;
    for (int l = 1; l <= 7; l++) { for (int s = 0; s < 19; s++) if (pl[s] == l) { const u32 r = rev16(code, l); for (u32 i = r; i < 128; i += 1u << l) ptab[i] = (u16)((s << 8) | l); code++; } code <<= 1; } }
  u8 lens[288 + 32]; const u32 total = hlit + hdist; u32 i = 0;
  while (i < total) {
    need(7);                                         // (fewer bits near the end of the input: checked against the code's length)
    const u16 e = ptab[bb_ & 127]; const u32 l = e & 0xFF, s = e >> 8;
    if (l > bc_) return false;
    take(l);
    if (s < 16) { lens[i++] = (u8)s; continue; }
    u32 rep; u8 val = 0;
    if (s == 16) { if (i == 0 || !get(2, v)) return false; rep = 3 + v; val = lens[i - 1]; }
    else if (s == 17) { if (!get(3, v)) return false; rep = 3 + v; }
    else { if (!get(7, v)) return false; rep = 11 + v; }
    if (i + rep > total) return false;
    memset(lens + i, val, rep); i += rep;
  }
  if (lens[256] == 0) return false;                  // no end-of-block code
  if (!build(lens, (int)hlit, 0, ltab_, LCAP) || !build(lens + hlit, (int)hdist, 1, dtab_, DCAP)) return false;
  lt_ = ltab_; dt_ = dtab_;
  return true;
}

template <typename T> inline int Inflater::huff(const T* base, T*& out_ref, T* out_end) {
  const u32* const lt = lt_; const u32* const dt = dt_;
  u64 bb = bb_; u32 bc = bc_; const u8* in = in_; T* out = out_ref;
  const u32 LMASK = (1u << LP) - 1, DMASK = (1u << DP) - 1;
  int rc = -1;
#define NBZ_REFILL() do { bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; } while (0)
  // ---- fast loop: room for two refills and nine literals + the longest match (plus the stores' overrun) without looking.
  // The next symbol's entry is looked up BEFORE the match is copied and before the refill (whose load does not touch the
  // bits the lookup used: above bc the buffer already holds stream bits), so the table load overlaps both.
#define NBZ_LITS() do { bb >>= (e & 15); bc -= (e & 15); put_lits(out, (e >> 8) & 0x7FFFFF); out += 1 + ((e >> 4) & 3); } while (0)
  if (end_ - in >= FAST_IN && out_end - out >= FAST_OUT) {
    NBZ_REFILL();
    u32 e = lt[bb & LMASK];
    for (;;) {                                       // here: refilled (bc >= 56), e = entry of the next symbol
      if ((int32_t)e < 0) {                          // literals; up to three entries (<= 11 bits, <= 3 literals each) per refill
        NBZ_LITS(); e = lt[bb & LMASK];
        if ((int32_t)e < 0) {
          NBZ_LITS(); e = lt[bb & LMASK];
          if ((int32_t)e < 0) { NBZ_LITS(); e = lt[bb & LMASK]; }
        }
        NBZ_REFILL();
        if ((int32_t)e < 0) goto next;
      }
      if (e & L_EXC) {
        if (e & L_SUB) {
          bb >>= LP; bc -= LP;
          e = lt[((e >> 16) & 0xFFF) + (u32)(bb & ((1u << ((e >> 8) & 0xF)) - 1))];
          if ((int32_t)e < 0) { NBZ_LITS(); e = lt[bb & LMASK]; NBZ_REFILL(); goto next; }
        }
        if (e & L_EXC) {
          if (!(e & L_EOB)) goto done;               // invalid code
          bb >>= (u8)e; bc -= (u8)e; rc = 0; goto done;
        }
      }
      {
        bb >>= (u8)e; bc -= (u8)e;
        const u32 xb = (e >> 8) & 0x1F;
        const u32 len = ((e >> 16) & 0x1FF) + (u32)(bb & ((1u << xb) - 1)); bb >>= xb; bc -= xb;
        u32 d = dt[bb & DMASK];
        if (d & D_SUB) { bb >>= DP; bc -= DP; d = dt[((d >> 16) & 0x1FF) + (u32)(bb & ((1u << ((d >> 8) & 0xF)) - 1))]; }
        if (d & D_BAD) goto done;
        bb >>= (u8)d; bc -= (u8)d;
        const u32 db = (d >> 8) & 0xF;
        const u32 dist = ((d >> 16) & 0x7FFF) + (u32)(bb & ((1u << db) - 1)); bb >>= db; bc -= db;
        if (dist > (size_t)(out - base)) goto done;  // before the start of the output
        e = lt[bb & LMASK];                          // (at most 48 bits used since the refill: 16 stream bits are left)
        T* const dst = out; out += len;
        copy_match(dst, dst - dist, out, dist);         // (writes up to 15 elements past the match: inside FAST_OUT)
        NBZ_REFILL();
      }
    next:
      if (!(end_ - in >= FAST_IN && out_end - out >= FAST_OUT)) break;
    }
  }
#undef NBZ_REFILL
#undef NBZ_LITS
  // ---- careful loop: symbol by symbol, nothing read or written outside the buffers, stops exactly where the window ends
  bb &= (1ull << bc) - 1;                            // (bc <= 63 everywhere; the fast refill leaves stream bits above bc)
  for (;;) {
    while (bc <= 55 && in < end_) { bb |= (u64)*in++ << bc; bc += 8; }
    const u64 s_bb = bb; const u32 s_bc = bc; const u8* const s_in = in;      // to take the symbol back when it does not fit
    int left = (int)bc;                              // goes negative when a code runs past the end of the input
    u32 e = lt[bb & LMASK];
    if ((int32_t)e >= 0 && (e & L_SUB)) { bb >>= LP; left -= LP; e = lt[((e >> 16) & 0xFFF) + (u32)(bb & ((1u << ((e >> 8) & 0xF)) - 1))]; }
    bb >>= (e & 15); left -= (int)(e & 15);
    if (left < 0) goto done;
    if ((int32_t)e < 0) {
      const u32 n = 1 + ((e >> 4) & 3);
      if ((size_t)(out_end - out) < n) { bb = s_bb; bc = s_bc; in = s_in; rc = 1; goto done; }
      u32 w = e >> 8; for (u32 i = 0; i < n; i++, w >>= 8) *out++ = (T)(i == 2 ? (w & 0x7F) : (w & 0xFF));
      bc = (u32)left; continue;
    }
    if (e & L_EXC) { if (e & L_EOB) { bc = (u32)left; rc = 0; } goto done; }
    const u32 xb = (e >> 8) & 0x1F;
    const u32 len = ((e >> 16) & 0x1FF) + (u32)(bb & ((1u << xb) - 1)); bb >>= xb; left -= (int)xb;
    u32 d = dt[bb & DMASK];
    if (d & D_SUB) { bb >>= DP; left -= DP; d = dt[((d >> 16) & 0x1FF) + (u32)(bb & ((1u << ((d >> 8) & 0xF)) - 1))]; }
    if (d & D_BAD) goto done;
    bb >>= (u8)d; left -= (int)(u8)d;
    const u32 db = (d >> 8) & 0xF;
    const u32 dist = ((d >> 16) & 0x7FFF) + (u32)(bb & ((1u << db) - 1)); bb >>= db; left -= (int)db;
    if (left < 0) goto done;
    if (dist > (size_t)(out - base)) goto done;
    if (len > (size_t)(out_end - out)) { bb = s_bb; bc = s_bc; in = s_in; rc = 1; goto done; }
    { const T* src = out - dist; for (u32 i = 0; i < len; i++) out[i] = src[i]; out += len; }
    bc = (u32)left;
  }
done:
  bb_ = bb & ((1ull << bc) - 1); bc_ = bc; in_ = in; out_ref = out;
  return rc;
}

inline int Inflater::run(const u8* base, u8*& out, u8* out_end) {
  for (;;) {
    if (!in_block_) {
      if (final_) return INF_END;
      if (bit_pos() >= stop_bit_ || (size_t)(out - base) >= stop_out_) return INF_STOP;
      if (!block_header()) return INF_ERROR;
      in_block_ = true;
    }
    if (btype_ == 0) {
      size_t n = stored_left_; if (n > (size_t)(out_end - out)) n = (size_t)(out_end - out);
      if (n > (size_t)(end_ - in_)) return INF_ERROR;
      memcpy(out, in_, n); out += n; in_ += n; stored_left_ -= (u32)n;
      if (stored_left_) return INF_MORE;
      in_block_ = false; continue;
    }
    const int r = huff<u8>(base, out, out_end);
    if (r < 0) return INF_ERROR;
    if (r == 1) return INF_MORE;
    in_block_ = false;
  }
}

inline int Inflater::run16(const u16* base, u16*& out, u16* out_end) {
  for (;;) {
    if (!in_block_) {
      if (final_) return INF_END;
      if (bit_pos() >= stop_bit_ || (size_t)(out - base) >= stop_out_) return INF_STOP;
      if (!block_header()) return INF_ERROR;
      in_block_ = true;
    }
    if (btype_ == 0) {
      size_t n = stored_left_; if (n > (size_t)(out_end - out)) n = (size_t)(out_end - out);
      if (n > (size_t)(end_ - in_)) return INF_ERROR;
      for (size_t i = 0; i < n; i++) out[i] = in_[i];
      out += n; in_ += n; stored_left_ -= (u32)n;
      if (stored_left_) return INF_MORE;
      in_block_ = false; continue;
    }
    const int r = huff<u16>(base, out, out_end);
    if (r < 0) return INF_ERROR;
    if (r == 1) return INF_MORE;
    in_block_ = false;
  }
}

// gzip member header at p (RFC 1952 2.3): the first byte of the deflate stream, or nullptr (not gzip / cut off)
inline const u8* gzip_header(const u8* p, const u8* end) {
  if (end - p < 10 || p[0] != 0x1F || p[1] != 0x8B || p[2] != 8 || (p[3] & 0xE0)) return nullptr;
  const u8 flg = p[3]; p += 10;
  if (flg & 4) { if (end - p < 2) return nullptr; const size_t xlen = p[0] | ((size_t)p[1] << 8); p += 2; if ((size_t)(end - p) < xlen) return nullptr; p += xlen; }
  for (int f = 8; f <= 16; f <<= 1) if (flg & f) { const u8* z = (const u8*)memchr(p, 0, (size_t)(end - p)); if (!z) return nullptr; p = z + 1; }   // FNAME, FCOMMENT
  if (flg & 2) { if (end - p < 2) return nullptr; p += 2; }
  return p;
}

// Concatenated gzip members held in memory, decoded into windows the caller provides (each window continues the output:
// when a member runs on from the previous window, [base, out) must hold the previous 32 KiB).  CRC-32 and ISIZE of every
// member are checked.  Bytes behind the last member that do not start another one are ignored, like gzip(1) and gzread.
class GzipStream {
 public:
  void open(const u8* data, size_t n) { p_ = data; end_ = data + n; in_member_ = false; done_ = false; members_ = 0; }
  bool done() const { return done_; }
  // Produces into [out, out_end) until the window is full or the input ends; returns the bytes produced, -1 on a damaged
  // stream.  After the last byte done() is true.
  ptrdiff_t read(const u8* base, u8* out, u8* out_end) {
    u8* const out0 = out;
    const u8* mbase = base;
    while (!done_) {
      if (!in_member_) {
        if (p_ == end_) { if (!members_) return -1; done_ = true; break; }          // (an empty file is not gzip)
        const u8* d = gzip_header(p_, end_);
        if (!d) { if (!members_) return -1; done_ = true; break; }                   // trailing bytes that are no member
        inf_.start(d, end_); in_member_ = true; crc_ = crc32_z(0L, nullptr, 0); size_ = 0; mbase = out;
      }
      u8* q = out;
      const int r = inf_.run(mbase, q, out_end);
      if (r == INF_ERROR) return -1;
      crc_ = crc32_z(crc_, out, (size_t)(q - out)); size_ += (u64)(q - out); out = q;
      if (r == INF_MORE) break;
      const u8* t = inf_.next_byte();
      if (end_ - t < 8) return -1;
      u32 want_crc, want_size; memcpy(&want_crc, t, 4); memcpy(&want_size, t + 4, 4);
      if (want_crc != (u32)crc_ || want_size != (u32)size_) return -1;
      p_ = t + 8; in_member_ = false; members_++;
    }
    return out - out0;
  }
 private:
  const u8* p_ = nullptr; const u8* end_ = nullptr; Inflater inf_; bool in_member_ = false, done_ = false; unsigned long crc_ = 0; u64 size_ = 0, members_ = 0;
};

}  // namespace nbz
