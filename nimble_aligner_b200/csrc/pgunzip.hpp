// pgunzip.hpp — one gzip file inflated by several threads (fastq.cpp: the .fastq.gz feeder; niffler + flate2's place in
// /root/reference/src/parse/fastq.rs:21-43, which inflate on the one thread that also parses).
//
// A deflate stream has no entry points, but a decoder CAN start at any block header if it treats the 32 KiB in front of it
// as unknown: it decodes into 16-bit symbols in which a byte copied from that unknown window is a placeholder naming the
// window position (Inflater::run16), and whoever later knows the window's real content substitutes them.  So the
// compressed file is cut into byte ranges ("chunks"), and for chunk k a worker
//   1. searches the first position in its range that holds a sound non-final dynamic block header (every field of the
//      header is validated: complete code-length code, complete literal/length and distance codes, an end-of-block code),
//   2. decodes from there, speculatively, until the first block boundary at or behind the end of its range,
//   3. waits for chunk k-1's worker to publish where ITS decoding really ended and the last 32 KiB of its text,
//   4. accepts its own speculative symbols only if they begin exactly there — otherwise it decodes the gap (or, when its
//      guess was no block boundary at all, its whole range) the ordinary way from the published position —, publishes its own
//      end position and window for chunk k+1 right away, and only then substitutes the placeholders of its symbols.
// Steps 1, 2 and the substitution run on all threads at once; the chain through step 3/4 carries 32 KiB per chunk.  Nothing
// depends on a guess being right: text is only ever emitted for a decode that started at a position the chain reached, and
// every gzip member's CRC-32 and ISIZE are checked over the emitted text (per-chunk CRCs put together with crc32_combine).
// Members (concatenated .gz, bgzip) are walked by the chain; the decode of a chunk stops at a member's end.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "inflate.hpp"

namespace nbz {

struct PgChunk {
  std::vector<u8> text;          // this chunk's part of the inflated file
  int status = 1;                // 1: more chunks follow, 0: last chunk, -1: damaged stream
  void* user_ptr = nullptr; size_t user_a = 0, user_b = 0;   // for the post-processing hook (fastq.cpp: the records parsed by the worker)
};

class ParallelGunzip {
 public:
  ~ParallelGunzip() {
    finish();
    if (getenv("NB_GZ_STATS") && n_) fprintf(stderr, "parallel gunzip: %zu chunks of %zu KiB on %d threads: %zu entered at a guessed block / member header, %zu guesses discarded, %.1f MB decoded the ordinary way\n",
                                             tasks_.size(), C_ >> 10, T_, (size_t)n_spec_, (size_t)n_bad_, (double)serial_bytes_ / 1e6),
                                     fprintf(stderr, "  thread time: find %.2fs, entered decode %.2fs, waiting for the chain %.2fs, on the chain %.2fs, placeholders %.2fs, crc %.2fs, parse %.2fs\n",
                                             ns_find_ / 1e9, ns_spec_ / 1e9, ns_wait_ / 1e9, ns_chain_ / 1e9, ns_resolve_ / 1e9, ns_crc_ / 1e9, ns_post_ / 1e9);
  }
  static size_t chunks_of(size_t n, size_t chunk_bytes) { const size_t c = chunk_bytes < 4096 ? 4096 : chunk_bytes; return n ? (n + c - 1) / c : 1; }
  // post: called by the worker on each finished chunk (off the chain, before the chunk is handed out)
  void open(const u8* data, size_t n, int threads, size_t chunk_bytes, std::function<void(size_t, PgChunk&)> post = nullptr) {
    data_ = data; n_ = n; C_ = chunk_bytes < 4096 ? 4096 : chunk_bytes; T_ = threads < 1 ? 1 : threads; post_ = post;
    const size_t nt = chunks_of(n_, C_);
    tasks_.clear(); for (size_t i = 0; i < nt; i++) tasks_.emplace_back(new Task());
    look_ = (size_t)T_ * 2 + 2;
    for (int t = 0; t < T_; t++) th_.emplace_back([this] { work(); });
  }
  // chunks in file order; blocks until the next one is complete.  nullptr after the last one was handed out.
  PgChunk* next() {
    Task* tp;
    { std::unique_lock<std::mutex> lk(m_); if (delivered_ >= tasks_.size()) return nullptr; tp = tasks_[delivered_].get(); cv_.wait(lk, [&] { return tp->done; }); }
    Task& t = *tp;
    // member checksums over the emitted text, in file order
    for (const Seg& s : t.segs) {
      crc_ = crc32_combine(crc_, s.crc, (z_off_t)s.len); size_ += s.len;
      if (s.ends_member) { if ((u32)crc_ != s.want_crc || (u32)size_ != s.want_size) t.out.status = -1; crc_ = 0; size_ = 0; }
    }
    if (failed_) t.out.status = -1;
    if (t.out.status < 0) failed_ = true;
    delivered_++;
    return &t.out;
  }
  void recycle(PgChunk* c) {
    { std::lock_guard<std::mutex> lk(m_); recycled_++; if (c->text.capacity()) { c->text.clear(); text_pool_.push_back(std::move(c->text)); } }
    std::vector<u8>().swap(c->text);
    cv_.notify_all();
  }
  void finish() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) if (t.joinable()) t.join();
    th_.clear();
  }

 private:
  struct Seg { size_t len; u32 crc; bool ends_member; u32 want_crc, want_size; };
  struct Chain {                 // where the decode stands behind a chunk
    bool in_member = false; u64 bit = 0;       // in a member: the next block header's bit
    size_t hdr = 0;                             // between members: the byte where the next member's header is expected
    std::vector<u8> window;                     // the last <= 32 KiB of the current member's text
    u64 members = 0; bool done = false, error = false;
  };
  struct Task { PgChunk out; std::vector<Seg> segs; Chain chain; bool chain_ready = false, done = false; };

  const u8* data_ = nullptr; size_t n_ = 0, C_ = 0; int T_ = 1; size_t look_ = 4; std::function<void(size_t, PgChunk&)> post_;
  std::vector<std::unique_ptr<Task>> tasks_; std::vector<std::thread> th_;
  std::mutex m_; std::condition_variable cv_; size_t next_task_ = 0, recycled_ = 0; bool stop_ = false;
  std::vector<std::vector<u8>> text_pool_;
  std::atomic<size_t> n_spec_{0}, n_bad_{0}, serial_bytes_{0};
  std::atomic<u64> ns_find_{0}, ns_spec_{0}, ns_wait_{0}, ns_chain_{0}, ns_resolve_{0}, ns_crc_{0}, ns_post_{0};   // NB_GZ_STATS
  static u64 now_ns() { return (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  size_t delivered_ = 0; unsigned long crc_ = 0; u64 size_ = 0; bool failed_ = false;   // consumer side

  size_t out_cap() const { const size_t c = C_ * 16; return c < ((size_t)8 << 20) ? ((size_t)8 << 20) : c; }   // text one chunk may hold (+ one deflate block)
  u64 peek_bits(u64 bit) const {             // 57 bits from `bit` on (zeros behind the end of the data)
    const size_t b = (size_t)(bit >> 3); u64 v = 0;
    if (b + 8 <= n_) memcpy(&v, data_ + b, 8); else if (b < n_) memcpy(&v, data_ + b, n_ - b);
    return v >> (bit & 7);
  }
  // first bit in [from, to) with a sound non-final dynamic block header; ~0 when there is none
  u64 find_block(u64 from, u64 to, Inflater& scratch) const {
    for (u64 bit = from; bit < to; bit++) {
      const u64 v = peek_bits(bit);
      if ((v & 7) != 4) continue;                                   // BFINAL = 0, BTYPE = 2
      if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;  // HLIT, HDIST
      const u32 hclen = (u32)((v >> 13) & 15) + 4;
      const u64 w = peek_bits(bit + 17);                            // the code-length code's lengths: complete?
      u32 kraft = 0; for (u32 i = 0; i < hclen; i++) { const u32 l = (u32)(w >> (3 * i)) & 7; if (l) kraft += 128u >> l; }
      if (kraft != 128) continue;
      if (scratch.header_at_bit(data_, data_ + n_, bit)) return bit;
    }
    return ~0ull;
  }

  void work() {
    std::unique_ptr<Inflater> inf(new Inflater()), scratch(new Inflater());
    std::vector<u16> sym; std::vector<u8> btext;     // this worker's symbol / member-text buffers, kept between chunks (fresh pages are not free)
    for (;;) {
      size_t k;
      { std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (next_task_ < tasks_.size() && next_task_ < recycled_ + look_); });
        if (stop_) return;
        k = next_task_++; }
      run_task(k, *inf, *scratch, sym, btext);
    }
  }

  // waits for chunk k-1's chain; false when the reader is being shut down
  bool wait_chain(size_t k, Chain& cs) {
    if (k == 0) { cs = Chain(); return true; }
    std::unique_lock<std::mutex> lk(m_);
    Task& p = *tasks_[k - 1];
    cv_.wait(lk, [&] { return stop_ || p.chain_ready; });
    if (stop_) return false;
    cs = p.chain;
    return true;
  }
  void publish_chain(Task& t, const Chain& cs) { { std::lock_guard<std::mutex> lk(m_); t.chain = cs; t.chain_ready = true; } cv_.notify_all(); }
  static void push_window(std::vector<u8>& w, const u8* p, size_t n) {       // w = last 32 KiB of (w + p[0..n))
    if (n >= 32768) { w.assign(p + n - 32768, p + n); return; }
    if (w.size() + n > 32768) w.erase(w.begin(), w.begin() + (w.size() + n - 32768));
    w.insert(w.end(), p, p + n);
  }

  void run_task(size_t k, Inflater& inf, Inflater& scratch, std::vector<u16>& sym, std::vector<u8>& btext) {
    Task* tp; bool is_last;
    { std::lock_guard<std::mutex> lk(m_); tp = tasks_[k].get(); is_last = k + 1 == tasks_.size(); }   // (the list may grow: see the end of the chain loop)
    Task& t = *tp;
    const u64 lo = (u64)k * C_ * 8, limit = is_last ? ~0ull : (u64)(k + 1) * C_ * 8;
    // ---- speculative decode of this range, entered at a guessed block header
    size_t nsym = 0; u64 s = ~0ull, e = 0; int s_end = INF_ERROR; const u8* s_next = nullptr;
    u64 t0 = now_ns();
    // A file of many members (bgzip: one per 64 KiB; cat of .gz files) needs no guessing about windows: a member starts with an
    // empty one.  The worker looks for a member header in its range — the first candidate whose member decodes to the CRC-32 and
    // ISIZE of its trailer — and decodes whole members from there, as bytes; the chain takes them if it arrives at that very
    // byte between two members.
    std::vector<std::pair<size_t, std::pair<u32, u32>>> bmembers;   // (end offset in btext, (CRC-32, ISIZE) of the trailer)
    size_t b_start = ~(size_t)0, b_end = 0, b_used = 0;
    if (k > 0 && T_ > 1 && lo < (u64)n_ * 8) {
      const size_t lo_b = (size_t)(lo >> 3), hi_b = limit == ~0ull ? n_ : (size_t)(limit >> 3);
      for (size_t h = lo_b; h + 18 <= n_ && h < hi_b; h++) {
        const u8* q = (const u8*)memchr(data_ + h, 0x1F, (hi_b < n_ - 17 ? hi_b : n_ - 17) - h); if (!q) break;
        h = (size_t)(q - data_);
        if (q[1] != 0x8B || q[2] != 8 || (q[3] & 0xE0)) continue;
        const u8* d = gzip_header(q, data_ + n_); if (!d) continue;
        if (btext.size() < ((size_t)1 << 20)) btext.resize((size_t)1 << 20);
        inf.start(d, data_ + n_);
        u8* o = btext.data(); int r;
        for (;;) { r = inf.run(btext.data(), o, btext.data() + btext.size()); if (r != INF_MORE || btext.size() >= out_cap()) break; const size_t at = (size_t)(o - btext.data()); btext.resize(btext.size() * 2); o = btext.data() + at; }
        const u8* tr = inf.next_byte();
        if (r != INF_END || (size_t)(data_ + n_ - tr) < 8) continue;
        u32 wc, ws; memcpy(&wc, tr, 4); memcpy(&ws, tr + 4, 4);
        if (ws != (u32)(o - btext.data()) || wc != (u32)crc32_z(0L, btext.data(), (size_t)(o - btext.data()))) continue;
        b_start = h; break;
      }
      if (b_start != ~(size_t)0) {
        const u8* p = data_ + b_start; const size_t cap = out_cap(); bool good = true;
        if (btext.size() < cap + (1u << 20)) btext.resize(cap + (1u << 20));
        size_t used = 0;
        while (good && p < data_ + n_ && (size_t)(p - data_) < hi_b && used < cap) {
          const u8* d = gzip_header(p, data_ + n_); if (!d) break;
          inf.start(d, data_ + n_);
          u8* o = btext.data() + used; int r;
          for (;;) { r = inf.run(btext.data() + used, o, btext.data() + btext.size()); if (r != INF_MORE) break; const size_t at = (size_t)(o - btext.data()); btext.resize(btext.size() * 2); o = btext.data() + at; }
          const u8* tr = inf.next_byte();
          if (r != INF_END || (size_t)(data_ + n_ - tr) < 8) { good = false; break; }
          u32 wc, ws; memcpy(&wc, tr, 4); memcpy(&ws, tr + 4, 4);
          used = (size_t)(o - btext.data()); bmembers.push_back({used, {wc, ws}}); p = tr + 8;
        }
        if (!good || bmembers.empty()) { b_start = ~(size_t)0; bmembers.clear(); } else { b_end = (size_t)(p - data_); b_used = used; }
      }
      { const u64 t1 = now_ns(); ns_find_ += t1 - t0; t0 = t1; }
    }
    if (k > 0 && T_ > 1 && lo < (u64)n_ * 8 && b_start == ~(size_t)0) {
      { u64 to = limit == ~0ull ? (u64)n_ * 8 : limit; if (to - lo > ((u64)4 << 20)) to = lo + ((u64)4 << 20);      // (no block header in half a megabyte: stored or giant blocks — left to the chain)
        s = find_block(lo, to, scratch); }
      { const u64 t1 = now_ns(); ns_find_ += t1 - t0; t0 = t1; }
      if (s != ~0ull) {
        if (sym.size() < 32768 + C_ * 6 + 4096) sym.resize(32768 + C_ * 6 + 4096);
        for (u32 i = 0; i < 32768; i++) sym[i] = (u16)(0x8000 + i);
        inf.start_at_bit(data_, data_ + n_, s, limit);
        const size_t cap = out_cap();
        u16* o = sym.data() + 32768;
        for (;;) {
          const int r = inf.run16(sym.data(), o, sym.data() + sym.size());
          if (r != INF_MORE) { s_end = r; break; }
          const size_t at = (size_t)(o - sym.data());
          if (at > 32768 + cap) { s_end = INF_ERROR; break; }     // far more text than read data expands to: a wrong guess, or data of another kind — left to the ordinary decode, which works in bounded pieces
          sym.resize(sym.size() * 2); o = sym.data() + at;
        }
        nsym = (size_t)(o - sym.data()) - 32768;
        if (s_end == INF_STOP) e = inf.bit_pos(); else if (s_end == INF_END) s_next = inf.next_byte(); else s = ~0ull;
      }
    }
    // ---- the chain: from where chunk k-1 really ended to the end of this range
    { const u64 t1 = now_ns(); ns_spec_ += t1 - t0; t0 = t1; }
    Chain cs;
    if (!wait_chain(k, cs)) return;
    { const u64 t1 = now_ns(); ns_wait_ += t1 - t0; t0 = t1; }
    struct Piece { bool spec; size_t off, len; std::vector<u8> window; };      // parts of this chunk's text, in order
    std::vector<u8>& text = t.out.text; t.segs.clear();
    { std::lock_guard<std::mutex> lk(m_); if (!text_pool_.empty()) { text = std::move(text_pool_.back()); text_pool_.pop_back(); } }   // (a buffer whose pages exist already)
    text.clear();
    std::vector<Piece> pieces; std::vector<size_t> seg_end; std::vector<Seg> seg_info;   // (segment ends as offsets into text)
    bool spec_used = false, guess_used = false, capped = false; const bool have_guess = s != ~0ull || b_start != ~(size_t)0; std::vector<u8> buf;
    auto end_member = [&](const u8* trailer) -> bool {               // trailer: 8 bytes behind the deflate stream
      if ((size_t)(data_ + n_ - trailer) < 8) return false;
      Seg g; g.len = 0; g.crc = 0; g.ends_member = true; memcpy(&g.want_crc, trailer, 4); memcpy(&g.want_size, trailer + 4, 4);
      seg_end.push_back(text.size()); seg_info.push_back(g);
      cs.in_member = false; cs.hdr = (size_t)(trailer + 8 - data_); cs.members++; cs.window.clear();
      return true;
    };
    while (!cs.done && !cs.error) {
      if (!cs.in_member) {
        if (cs.hdr >= n_) { if (!cs.members) cs.error = true; cs.done = true; break; }
        if ((u64)cs.hdr * 8 >= limit) break;                          // the next member begins in a later chunk's range
        if (b_start != ~(size_t)0 && cs.hdr == b_start) {              // the members this worker decoded on its own begin exactly here
          const size_t off = text.size();
          text.insert(text.end(), btext.begin(), btext.begin() + b_used);
          for (const auto& mb : bmembers) { Seg g; g.len = 0; g.crc = 0; g.ends_member = true; g.want_crc = mb.second.first; g.want_size = mb.second.second; seg_end.push_back(off + mb.first); seg_info.push_back(g); }
          cs.hdr = b_end; cs.members += bmembers.size(); cs.window.clear(); b_start = ~(size_t)0; guess_used = true; n_spec_++;
          continue;
        }
        if (b_start != ~(size_t)0 && cs.hdr > b_start) b_start = ~(size_t)0;   // that signature was not a member boundary
        const u8* d = gzip_header(data_ + cs.hdr, data_ + n_);
        if (!d) { if (!cs.members) cs.error = true; cs.done = true; break; }      // bytes behind the last member that are no member: ignored
        cs.in_member = true; cs.bit = (u64)(d - data_) * 8; cs.window.clear();
        continue;
      }
      if (cs.bit >= limit) break;
      if (s != ~0ull && !spec_used && cs.bit == s) {
        // the guess was a block boundary the real decode arrives at: this chunk's symbols are its text
        spec_used = guess_used = true; n_spec_++;
        Piece p; p.spec = true; p.off = text.size(); p.len = nsym; p.window = cs.window;
        text.resize(text.size() + nsym);
        // the window behind the piece: only its last 32 KiB have to be real bytes now
        { const size_t tail = nsym < 32768 ? nsym : 32768; std::vector<u8> last(tail);
          if (!resolve(sym.data() + 32768 + nsym - tail, tail, p.window, last.data())) { cs.error = true; break; }
          push_window(cs.window, last.data(), tail); }
        pieces.push_back(std::move(p));
        if (s_end == INF_STOP) cs.bit = e; else if (!end_member(s_next)) { cs.error = true; break; }
        continue;
      }
      // ordinary decode from the chain's position: up to the guess (when it is still ahead), else to the end of the range
      const u64 stop = (s != ~0ull && !spec_used && cs.bit < s) ? s : limit;
      if (s != ~0ull && !spec_used && cs.bit > s) s = ~0ull;         // the real decode went past the guess: it was no boundary
      const size_t w = cs.window.size(), off = text.size();
      inf.start_at_bit(data_, data_ + n_, cs.bit, stop, w + out_cap());       // (bounded memory on data that expands a thousandfold: the next chunk's worker carries on from where this stops)
      if (buf.size() < w + 65536) buf.resize(w + C_ * 5 + 65536);
      if (w) memcpy(buf.data(), cs.window.data(), w);
      u8* o = buf.data() + w; int r;
      for (;;) {
        r = inf.run(buf.data(), o, buf.data() + buf.size());
        if (r != INF_MORE) break;
        const size_t at = (size_t)(o - buf.data()); buf.resize(buf.size() * 2); o = buf.data() + at;
      }
      if (r == INF_ERROR) { cs.error = true; break; }
      const size_t got = (size_t)(o - buf.data()) - w; serial_bytes_ += got;
      text.insert(text.end(), buf.data() + w, buf.data() + w + got);
      Piece p; p.spec = false; p.off = off; p.len = got; pieces.push_back(std::move(p));
      push_window(cs.window, buf.data() + w, got);
      if (r == INF_STOP) { cs.bit = inf.bit_pos(); if (got >= out_cap()) { capped = true; break; } } else if (!end_member(inf.next_byte())) { cs.error = true; break; }
    }
    if (is_last && capped && !cs.error && !cs.done) {   // the last range still holds more text than a chunk may: one more chunk behind this one carries on
      { std::lock_guard<std::mutex> lk(m_); tasks_.emplace_back(new Task()); }
      is_last = false;
    }
    if (is_last && !cs.error && !cs.done) cs.error = true;   // the data ended inside a member
    if (have_guess && !guess_used) n_bad_++;
    publish_chain(t, cs);
    { const u64 t1 = now_ns(); ns_chain_ += t1 - t0; t0 = t1; }
    // ---- off the chain: placeholders -> bytes, checksums
    bool ok = !cs.error;
    for (const Piece& p : pieces) if (ok && p.spec) ok = resolve(sym.data() + 32768, p.len, p.window, text.data() + p.off);
    { const u64 t1 = now_ns(); ns_resolve_ += t1 - t0; t0 = t1; }
    if (ok) {
      size_t a = 0;
      for (size_t i = 0; i <= seg_end.size(); i++) {
        const size_t b = i < seg_end.size() ? seg_end[i] : text.size();
        Seg g; if (i < seg_end.size()) g = seg_info[i]; else { g.ends_member = false; g.want_crc = g.want_size = 0; }
        g.len = b - a; g.crc = (u32)crc32_z(0L, text.data() + a, b - a);
        if (g.len || g.ends_member) t.segs.push_back(g);
        a = b;
      }
    }
    t.out.status = !ok ? -1 : (is_last ? 0 : 1);
    { const u64 t1 = now_ns(); ns_crc_ += t1 - t0; t0 = t1; }
    if (post_) post_(k, t.out);
    { const u64 t1 = now_ns(); ns_post_ += t1 - t0; t0 = t1; }
    { std::lock_guard<std::mutex> lk(m_); t.done = true; }
    cv_.notify_all();
  }

  // symbols -> bytes with the real window (the last window.size() <= 32768 bytes in front of the entry point); false when a
  // placeholder names a byte in front of the member's start
  static bool resolve(const u16* sym, size_t n, const std::vector<u8>& window, u8* out) {
    const size_t w = window.size(); const u32 first = 0x8000u + (u32)(32768 - w);       // smallest valid placeholder = window[0]
    const u8* wp = window.data();
    if (w == 32768) {
      // the usual case, and the hot one: on read data placeholders do not die out (sequence is compressed as matches into earlier
      // sequence, so bytes copied from the unknown window keep being copied) — no branch per symbol, one table of every value a symbol can take
      std::vector<u8> lut(65536); for (u32 v = 0; v < 256; v++) lut[v] = (u8)v;
      memcpy(lut.data() + 0x8000, wp, 32768);
      const u8* L = lut.data(); size_t i = 0;
      for (; i + 8 <= n; i += 8) { out[i] = L[sym[i]]; out[i + 1] = L[sym[i + 1]]; out[i + 2] = L[sym[i + 2]]; out[i + 3] = L[sym[i + 3]]; out[i + 4] = L[sym[i + 4]]; out[i + 5] = L[sym[i + 5]]; out[i + 6] = L[sym[i + 6]]; out[i + 7] = L[sym[i + 7]]; }
      for (; i < n; i++) out[i] = L[sym[i]];
      return true;
    }
    u32 bad = 0;                                     // near a member's start the window is shorter: a placeholder may name a byte in front of it
    for (size_t i = 0; i < n; i++) { const u32 v = sym[i]; if (v < 0x8000u) out[i] = (u8)v; else if (v >= first) out[i] = wp[v - first]; else bad = 1; }
    return !bad;
  }
};

}  // namespace nbz
