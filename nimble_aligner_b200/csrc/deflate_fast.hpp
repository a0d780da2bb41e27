// deflate_fast.hpp — gzip members for the BAM driver's TSV.gz output (bam.cpp's rows stage; flate2's GzEncoder in
// /root/reference/src/process/bam.rs:22-42, 90-121 and src/utils.rs).  Once the row formatter stopped being the cost, zlib's
// deflate of the rows — about 400 bytes per output row, two records' worth of metadata — was the largest single share of the
// driver's host time.  The rows are extremely regular (a row repeats most of the row before it, field by field), which a
// compressor written for them exploits cheaply:
//   * greedy LZ77 over a 32 KiB window with ONE candidate per position from a hash of four bytes, and before that a try at the
//     distance of the previous match (a row that keeps matching the row above it continues at the same distance after the
//     258-byte cap or after a field that differs);
//   * matches are extended eight bytes at a time; only a few positions inside a match are entered into the hash table;
//   * tokens are collected per block (up to 64 Ki tokens), then written with dynamic Huffman codes built from the block's own
//     statistics (length-limited to 15 bits by the usual Kraft repair), code lengths run-length coded as RFC 1951 prescribes.
// Any inflater reads the result (tests: zlib and inflate.hpp on every corpus file); the CRC-32 is zlib's.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>   // crc32_z only

#include <algorithm>
#include <string>
#include <vector>

namespace nbz {

class FastDeflate {
 public:
  // appends one gzip member holding data[0, n) to out (several members when n is beyond 1 GiB: positions are 32-bit, and
  // concatenated members are one valid gzip stream)
  void gzip_member(const uint8_t* data, size_t n, std::string& out) {
    const size_t piece = (size_t)1 << 30;
    while (n > piece) { one_member(data, piece, out); data += piece; n -= piece; }
    one_member(data, n, out);
  }
  void one_member(const uint8_t* data, size_t n, std::string& out) {
    static const uint8_t hdr[10] = {0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF};
    out.append((const char*)hdr, 10);
    deflate(data, n, out);
    const uint32_t crc = (uint32_t)crc32_z(0L, data, n), isize = (uint32_t)n;
    char t[8]; memcpy(t, &crc, 4); memcpy(t + 4, &isize, 4); out.append(t, 8);
  }
  // appends a raw deflate stream (final block included)
  void deflate(const uint8_t* d, size_t n, std::string& out) {
    w_.begin(out);
    if (n == 0) { w_.put(1, 1); w_.put(0, 2); w_.align(); w_.put(0, 16); w_.put(0xFFFF, 16); w_.end(); return; }   // one empty stored block
    if (head_.empty()) head_.assign(HSIZE, 0);
    // positions are stored + 1 (0 = never seen) relative to `d`; a part of the TSV is tens of megabytes, far below 2^32
    std::fill(head_.begin(), head_.end(), 0u);
    tokens_.clear(); tokens_.reserve(MAX_TOKENS);
    memset(lfreq_, 0, sizeof lfreq_); memset(dfreq_, 0, sizeof dfreq_);
    size_t i = 0, last_dist = 0;
    const size_t safe = n >= 12 ? n - 12 : 0;        // positions with 8 readable bytes behind a 4-byte hash
    while (i < n) {
      size_t len = 0, dist = 0;
      if (i < safe) {
        const uint32_t cur = load32(d + i);
        if (last_dist && load32(d + i - last_dist) == cur) { dist = last_dist; len = extend(d, i, i - last_dist, n); }
        const uint32_t h = (cur * 0x9E3779B1u) >> (32 - HBITS);
        const uint32_t c1 = head_[h]; head_[h] = (uint32_t)(i + 1);
        if (len < 32 && c1 && i + 1 - c1 <= 32768 && load32(d + c1 - 1) == cur) { const size_t l2 = extend(d, i, c1 - 1, n); if (l2 > len) { len = l2; dist = i + 1 - c1; } }
      }
      if (len >= 4) {
        tokens_.push_back(0x80000000u | (uint32_t)((len - 3) << 16) | (uint32_t)(dist - 1));
        lfreq_[257 + T().len_sym[len - 3]]++; dfreq_[dist_sym((uint32_t)dist)]++;
        // a few positions of the match go into the table: its second byte and its last four (where the next row's match may start)
        if (i + 1 < safe) { head_[(load32(d + i + 1) * 0x9E3779B1u) >> (32 - HBITS)] = (uint32_t)(i + 2); }
        for (size_t p = i + len - 3; p < i + len && p < safe; p++) if (p > i + 1) head_[(load32(d + p) * 0x9E3779B1u) >> (32 - HBITS)] = (uint32_t)(p + 1);
        i += len; last_dist = dist;
      } else { tokens_.push_back(d[i]); lfreq_[d[i]]++; i++; }
      if (tokens_.size() >= MAX_TOKENS) { write_block(false); tokens_.clear(); memset(lfreq_, 0, sizeof lfreq_); memset(dfreq_, 0, sizeof dfreq_); }
    }
    write_block(true);
    w_.end();
  }

 private:
  static const int HBITS = 15; static const size_t HSIZE = (size_t)1 << HBITS; static const size_t MAX_TOKENS = 1u << 16;
  struct Writer {
    std::string* out = nullptr; uint64_t acc = 0; int cnt = 0; char buf[4096]; size_t used = 0;
    void begin(std::string& o) { out = &o; acc = 0; cnt = 0; used = 0; }
    void put(uint32_t v, int nbits) {              // nbits <= 32
      acc |= (uint64_t)v << cnt; cnt += nbits;
      if (cnt >= 32) { if (used + 4 > sizeof buf) { out->append(buf, used); used = 0; } const uint32_t lo = (uint32_t)acc; memcpy(buf + used, &lo, 4); used += 4; acc >>= 32; cnt -= 32; }
    }
    void align() { if (cnt & 7) put(0, 8 - (cnt & 7)); }
    void end() { align(); while (cnt > 0) { if (used + 1 > sizeof buf) { out->append(buf, used); used = 0; } buf[used++] = (char)(acc & 0xFF); acc >>= 8; cnt -= 8; } out->append(buf, used); used = 0; cnt = 0; acc = 0; }
  };
  Writer w_; std::vector<uint32_t> head_, tokens_; uint32_t lfreq_[288], dfreq_[32];

  static uint32_t load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
  // length of the common run of d[i..] and d[c..] (c < i), at most 258 and never past n
  static size_t extend(const uint8_t* d, size_t i, size_t c, size_t n) {
    const size_t maxl = n - i < 258 ? n - i : 258; size_t l = 0;
    while (l + 8 <= maxl) { const uint64_t x = load64(d + i + l) ^ load64(d + c + l); if (x) return l + ((size_t)__builtin_ctzll(x) >> 3); l += 8; }
    while (l < maxl && d[i + l] == d[c + l]) l++;
    return l;
  }
  struct Tables {
    uint8_t len_sym[256], len_extra[256]; uint16_t len_base[256]; uint8_t dist_lo[512];   // dist_lo: symbol of dist-1 < 256, and of (dist-1) >> 7 behind it
    uint16_t dbase[30]; uint8_t dextra[30];
    Tables() {
      static const uint16_t LB[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
      static const uint8_t LE[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
      static const uint16_t DB[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
      static const uint8_t DE[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
      for (int l = 3; l <= 258; l++) { int s = 28; while (LB[s] > l) s--; if (l == 258) s = 28; len_sym[l - 3] = (uint8_t)s; len_extra[l - 3] = LE[s]; len_base[l - 3] = LB[s]; }
      for (int s = 0; s < 30; s++) { dbase[s] = DB[s]; dextra[s] = DE[s]; }
      for (int v = 0; v < 256; v++) { int s = 29; while (DB[s] > v + 1) s--; dist_lo[v] = (uint8_t)s; }
      for (int v = 0; v < 256; v++) { const int dist = (v << 7) + 1; int s = 29; while (DB[s] > dist) s--; dist_lo[256 + v] = (uint8_t)s; }   // for dist-1 >= 256 the symbol depends on (dist-1) >> 7 only
    }
  };
  static const Tables& T() { static const Tables t; return t; }
  static uint32_t dist_sym(uint32_t dist) { const uint32_t v = dist - 1; return v < 256 ? T().dist_lo[v] : T().dist_lo[256 + (v >> 7)]; }

  // code lengths (<= maxbits) for the symbols with freq > 0; at least two codes so that every decoder takes the set
  static void huff_lengths(const uint32_t* freq, int n, int maxbits, uint8_t* lens) {
    struct Node { uint64_t f; int a, b; };
    std::vector<std::pair<uint64_t, int>> syms; for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) syms.push_back({freq[i], i}); }
    if (syms.empty()) { lens[0] = 1; return; }
    if (syms.size() == 1) { lens[syms[0].second] = 1; lens[syms[0].second ? 0 : 1] = 1; return; }      // (a second, unused code makes the set complete)
    std::sort(syms.begin(), syms.end());
    const int m = (int)syms.size(); std::vector<Node> nodes(2 * m); for (int i = 0; i < m; i++) nodes[i] = {syms[i].first, -1, -1};
    int leaf = 0, inner = m, next = m;               // two queues: sorted leaves, and inner nodes in creation order (also sorted)
    auto take = [&]() { if (leaf < m && (inner >= next || nodes[leaf].f <= nodes[inner].f)) return leaf++; return inner++; };
    for (int it = 0; it + 1 < m; it++) { const int a = take(), b = take(); nodes[next] = {nodes[a].f + nodes[b].f, a, b}; next++; }
    std::vector<int> depth(next, 0);
    for (int v = next - 1; v >= m; v--) { depth[nodes[v].a] = depth[v] + 1; depth[nodes[v].b] = depth[v] + 1; }
    // limit: fold the lengths above maxbits into maxbits, then repair the Kraft sum by lengthening the shortest codes that can give
    std::vector<int> count(maxbits + 2, 0);
    for (int i = 0; i < m; i++) count[std::min(depth[i], maxbits)]++;
    uint64_t total = 0; for (int l = 1; l <= maxbits; l++) total += (uint64_t)count[l] << (maxbits - l);
    while (total > ((uint64_t)1 << maxbits)) {
      count[maxbits]--;
      for (int l = maxbits - 1; l > 0; l--) if (count[l]) { count[l]--; count[l + 1] += 2; break; }
      total--;
    }
    // the rarest symbols get the longest codes
    int k = 0; for (int l = maxbits; l >= 1; l--) for (int c = 0; c < count[l]; c++) lens[syms[k++].second] = (uint8_t)l;
  }
  static void canonical(const uint8_t* lens, int n, uint16_t* codes) {           // bit-reversed: the writer emits LSB first
    uint32_t next[17] = {0}, cnt[17] = {0}; for (int i = 0; i < n; i++) cnt[lens[i]]++; cnt[0] = 0;
    uint32_t code = 0; for (int l = 1; l <= 15; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
    for (int i = 0; i < n; i++) { const int l = lens[i]; if (!l) { codes[i] = 0; continue; } uint32_t c = next[l]++, r = 0; for (int b = 0; b < l; b++) { r = (r << 1) | (c & 1); c >>= 1; } codes[i] = (uint16_t)r; }
  }

  void write_block(bool final) {
    lfreq_[256]++;
    uint8_t ll[288], dl[32]; huff_lengths(lfreq_, 286, 15, ll); huff_lengths(dfreq_, 30, 15, dl);
    uint16_t lc[288], dc[32]; canonical(ll, 286, lc); canonical(dl, 30, dc);
    int hlit = 286; while (hlit > 257 && !ll[hlit - 1]) hlit--;
    int hdist = 30; while (hdist > 1 && !dl[hdist - 1]) hdist--;
    // code lengths, run-length coded (symbols 16 / 17 / 18), with their own Huffman code of at most 7 bits
    uint8_t seq[320]; int ns = 0; for (int i = 0; i < hlit; i++) seq[ns++] = ll[i]; for (int i = 0; i < hdist; i++) seq[ns++] = dl[i];
    struct Rl { uint8_t sym, extra; }; Rl rl[320]; int nr = 0; uint32_t pfreq[19] = {0};
    for (int i = 0; i < ns;) {
      int j = i; while (j < ns && seq[j] == seq[i]) j++;
      int run = j - i; const uint8_t v = seq[i];
      if (v == 0) { while (run >= 11) { const int r = std::min(run, 138); rl[nr++] = {18, (uint8_t)(r - 11)}; run -= r; } if (run >= 3) { rl[nr++] = {17, (uint8_t)(run - 3)}; run = 0; } }
      else if (run >= 4) { rl[nr++] = {v, 0}; run--; while (run >= 3) { const int r = std::min(run, 6); rl[nr++] = {16, (uint8_t)(r - 3)}; run -= r; } }
      while (run-- > 0) rl[nr++] = {v, 0};
      i = j;
    }
    for (int i = 0; i < nr; i++) pfreq[rl[i].sym]++;
    uint8_t pl[19]; huff_lengths(pfreq, 19, 7, pl); uint16_t pc[19]; canonical(pl, 19, pc);
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19; while (hclen > 4 && !pl[ORDER[hclen - 1]]) hclen--;
    w_.put(final ? 1 : 0, 1); w_.put(2, 2); w_.put((uint32_t)(hlit - 257), 5); w_.put((uint32_t)(hdist - 1), 5); w_.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) w_.put(pl[ORDER[i]], 3);
    for (int i = 0; i < nr; i++) { w_.put(pc[rl[i].sym], pl[rl[i].sym]); if (rl[i].sym == 16) w_.put(rl[i].extra, 2); else if (rl[i].sym == 17) w_.put(rl[i].extra, 3); else if (rl[i].sym == 18) w_.put(rl[i].extra, 7); }
    const Tables& t = T();
    for (const uint32_t tok : tokens_) {
      if (!(tok & 0x80000000u)) { w_.put(lc[tok], ll[tok]); continue; }
      const uint32_t l3 = (tok >> 16) & 0xFF, dist = (tok & 0xFFFF) + 1;
      const uint32_t ls = 257 + t.len_sym[l3]; w_.put(lc[ls], ll[ls]); if (t.len_extra[l3]) w_.put(l3 + 3 - t.len_base[l3], t.len_extra[l3]);
      const uint32_t ds = dist_sym(dist); w_.put(dc[ds], dl[ds]); if (t.dextra[ds]) w_.put(dist - t.dbase[ds], t.dextra[ds]);
    }
    w_.put(lc[256], ll[256]);
  }
};

}  // namespace nbz
