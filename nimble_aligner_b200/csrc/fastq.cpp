// fastq.cpp — FASTQ(.gz) feeder, FASTQ-mode TSV writer and the process::fastq::process driver.
// Mirrors /root/reference/src/parse/fastq.rs:8-43 (niffler gz autodetect + bio fastq records -> DnaString::from_acgt_bytes),
// src/utils.rs:27-51 (write_to_tsv: append, header iff empty, features TAB-joined) and src/process/fastq.rs:7-30.
#include <zlib.h>

#include <cstdio>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host.hpp"

using namespace nb;

namespace {

// One FASTQ(.gz) stream parsed in place: lines are views into the read buffer (a partial line at the end of the buffer
// is moved to the front before the next gzread), no per-line strings.
struct FastqReader {
  gzFile f = nullptr; std::string path; std::vector<char> buf; size_t pos = 0, len = 0; bool eof = false;
  bool open(const std::string& p) { path = p; f = gzopen(p.c_str(), "rb"); if (!f) return false; gzbuffer(f, 1 << 20); buf.resize(1 << 23); return true; }
  ~FastqReader() { if (f) gzclose(f); }
  // one line without its terminator as a view into buf; false at end of input; a line longer than the buffer grows it
  bool line(const char*& out, size_t& n) {
    for (;;) {
      const char* s = buf.data() + pos; const char* nl = len > pos ? (const char*)memchr(s, '\n', len - pos) : nullptr;
      if (nl) { out = s; n = (size_t)(nl - s); pos += n + 1; if (n && out[n - 1] == '\r') n--; return true; }
      if (eof) { if (pos >= len) return false; out = s; n = len - pos; pos = len; if (n && out[n - 1] == '\r') n--; return true; }
      if (pos) { memmove(buf.data(), buf.data() + pos, len - pos); len -= pos; pos = 0; }
      if (len == buf.size()) buf.resize(buf.size() * 2);
      int got = gzread(f, buf.data() + len, (unsigned)std::min<size_t>(buf.size() - len, 1u << 30));
      if (got <= 0) eof = true; else len += (size_t)got;
    }
  }
};

struct Pinned {
  u8* p = nullptr; size_t cap = 0;
  ~Pinned() { nb_host_free(p); }
  bool ensure(size_t n, size_t keep) { if (n <= cap) return true; size_t nc = std::max(n, cap * 2 + (1 << 20)); u8* q = (u8*)nb_host_alloc(nc); if (!q) return false; if (keep) memcpy(q, p, keep); nb_host_free(p); p = q; cap = nc; return true; }
};

// up to `want` records of one stream, bases concatenated in pinned memory
struct Block { Pinned seq; std::vector<u64> off; u64 n = 0; u32 maxlen = 0; size_t used = 0; int status = 1; };   // status: 1 more may follow, 0 end of file, -1 malformed, -2 out of pinned memory

// fills b from r; multi-line records are accepted like bio::io::fastq does
void parse_block(FastqReader& r, Block& b, u64 want) {
  b.n = 0; b.maxlen = 0; b.used = 0; b.status = 1; b.off.assign(1, 0);
  const char* s; size_t n;
  while (b.n < want) {
    do { if (!r.line(s, n)) { b.status = 0; return; } } while (n == 0);
    if (s[0] != '@') { b.status = -1; return; }
    size_t start = b.used;
    for (;;) {
      if (!r.line(s, n)) { b.status = -1; return; }
      if (n && s[0] == '+') break;
      if (!b.seq.ensure(b.used + n + 64, b.used)) { b.status = -2; return; }
      memcpy(b.seq.p + b.used, s, n); b.used += n;
    }
    size_t slen = b.used - start, qlen = 0;
    while (qlen < slen) { if (!r.line(s, n)) { b.status = -1; return; } qlen += n; }
    if (qlen != slen) { b.status = -1; return; }
    b.off.push_back(b.used); b.maxlen = std::max<u32>(b.maxlen, (u32)slen); b.n++;
  }
}

// a reader thread per input file: blocks circulate between `free` and `filled`
struct Feeder {
  FastqReader rd; Block blocks[3]; std::vector<Block*> free_q, filled_q; std::mutex m; std::condition_variable cv; bool stop = false; std::thread th; u64 want;
  void run() {
    for (;;) {
      Block* b;
      { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return stop || !free_q.empty(); }); if (stop) return; b = free_q.back(); free_q.pop_back(); }
      parse_block(rd, *b, want);
      { std::lock_guard<std::mutex> lk(m); filled_q.push_back(b); }
      cv.notify_all();
      if (b->status != 1) return;
    }
  }
  void start(u64 w) { want = w; for (auto& b : blocks) free_q.push_back(&b); th = std::thread([this] { run(); }); }
  Block* pop() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return !filled_q.empty(); }); Block* b = filled_q.front(); filled_q.erase(filled_q.begin()); return b; }
  void recycle(Block* b) { { std::lock_guard<std::mutex> lk(m); free_q.push_back(b); } cv.notify_all(); }
  void finish() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); if (th.joinable()) th.join(); }
};

int write_tsv(const std::string& path, const nb_library* lib, const nb_counts& cts) {  // utils::write_to_tsv
  FILE* f = fopen(path.c_str(), "ab");
  if (!f) return fail(NB_ERR_IO, "Unable to open file " + path);
  fseek(f, 0, SEEK_END);
  if (ftell(f) == 0) fputs("feature\tscore\n", f);
  for (u64 r = 0; r < cts.n_rows; r++) {
    u32 cs = cts.row_callset[r];
    for (u64 i = cts.callset_off[cs]; i < cts.callset_off[cs + 1]; i++) { fputs(nb_library_group_name(lib, cts.callset_items[i]), f); fputc('\t', f); }
    fprintf(f, "%lld\n", (long long)cts.row_count[r]);
  }
  fclose(f);
  return NB_OK;
}

}  // namespace

extern "C" int nb_write_fastq_tsv(const char* path, const nb_library* lib, const nb_counts* counts) {
  if (!path || !lib || !counts) return fail(NB_ERR_INVALID, "null argument");
  return write_tsv(path, lib, *counts);
}

// process::fastq::process with the library loop of src/bin/main.rs:95-133 in front of it.
extern "C" int nb_process_fastq(const char* const* input_files, uint32_t n_inputs, const char* const* reference_json, const char* const* output_paths,
                                uint32_t n_refs, int strand_filter, int num_cores, int device) {
  if (!input_files || !reference_json || !output_paths || n_inputs < 1 || n_inputs > 2 || n_refs < 1) return fail(NB_ERR_INVALID, "need 1-2 inputs and >=1 reference/output pair");
  const u64 BATCH = 1u << 19;
  for (u32 li = 0; li < n_refs; li++) {
    nb_library* lib = nullptr; nb_index* ix = nullptr; nb_ctx* ctx = nullptr;
    int rc = nb_library_load_json(reference_json[li], strand_filter, &lib);
    if (rc == NB_OK) rc = nb_index_build(lib, num_cores, &ix);
    if (rc == NB_OK) rc = nb_ctx_create(ix, lib, device, nullptr, &ctx);
    if (rc == NB_OK) rc = nb_ctx_set_option(ctx, "max_batch_pairs", BATCH);
    // one reader thread per input file parses blocks of BATCH records into pinned memory while the device works
    Feeder fd[2]; bool paired = n_inputs > 1;
    if (rc == NB_OK && !fd[0].rd.open(input_files[0])) rc = fail(NB_ERR_IO, std::string("could not open ") + input_files[0]);
    if (rc == NB_OK && paired && !fd[1].rd.open(input_files[1])) rc = fail(NB_ERR_IO, std::string("could not open ") + input_files[1]);
    if (rc == NB_OK) { fd[0].start(BATCH); if (paired) fd[1].start(BATCH); }
    while (rc == NB_OK) {
      Block* a = fd[0].pop(); Block* b2 = paired ? fd[1].pop() : nullptr;
      if (a->status == -2 || (b2 && b2->status == -2)) rc = fail(NB_ERR_CUDA, "pinned allocation failed");
      else if (a->status == -1) rc = fail(NB_ERR_PARSE, "Error -- could not parse read. Input R1 data malformed.");
      else if (b2 && b2->status == -1) rc = fail(NB_ERR_PARSE, "Error -- could not parse reverse read. Input R2 data malformed.");
      else if (b2 && (a->n != b2->n || a->status != b2->status)) rc = fail(NB_ERR_PARSE, "Error -- read and reverse read files do not have matching lengths: ");
      if (rc != NB_OK) break;
      if (a->n) {
        nb_batch b; memset(&b, 0, sizeof b);
        b.n_pairs = a->n; b.location = NB_MEM_HOST; b.max_read_len = std::max<u32>(a->maxlen, b2 ? b2->maxlen : 0);
        b.r1 = a->seq.p; b.r1_off = a->off.data();
        if (b2) { b.r2 = b2->seq.p; b.r2_off = b2->off.data(); }
        rc = nb_align_batch(ctx, &b, nullptr, nullptr);
        if (rc == NB_OK) rc = nb_ctx_sync(ctx);   // the copies out of these blocks are done: they can be refilled (a batch is ~5 ms on the device, ~100 ms to parse)
      }
      bool last = a->status == 0;
      fd[0].recycle(a); if (b2) fd[1].recycle(b2);
      if (last) break;
    }
    fd[0].finish(); if (paired) fd[1].finish();
    nb_counts cts;
    if (rc == NB_OK) rc = nb_counts_finalize(ctx, &cts);
    if (rc == NB_OK) rc = write_tsv(output_paths[li], lib, cts);
    nb_ctx_free(ctx); nb_index_free(ix); nb_library_free(lib);
    if (rc != NB_OK) return rc;
  }
  return NB_OK;
}
