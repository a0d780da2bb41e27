// fastq.cpp — FASTQ(.gz) feeder, FASTQ-mode TSV writer and the process::fastq::process driver.
// Mirrors /root/reference/src/parse/fastq.rs:8-43 (niffler gz autodetect + bio fastq records -> DnaString::from_acgt_bytes),
// src/utils.rs:27-51 (write_to_tsv: append, header iff empty, features TAB-joined) and src/process/fastq.rs:7-30.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host.hpp"

using namespace nb;

namespace {

struct FastqReader {
  gzFile f = nullptr; std::string path; std::vector<char> buf; size_t pos = 0, len = 0; bool eof = false;
  bool open(const std::string& p) { path = p; f = gzopen(p.c_str(), "rb"); if (!f) return false; gzbuffer(f, 1 << 20); buf.resize(1 << 22); return true; }
  ~FastqReader() { if (f) gzclose(f); }
  bool fill() { if (eof) return false; int n = gzread(f, buf.data(), (unsigned)buf.size()); if (n <= 0) { eof = true; return false; } pos = 0; len = (size_t)n; return true; }
  // reads one line without the terminator; false at end of input
  bool line(std::string& out) {
    out.clear(); bool any = false;
    for (;;) {
      if (pos >= len && !fill()) return any;
      any = true;
      const char* s = buf.data() + pos; const char* nl = (const char*)memchr(s, '\n', len - pos);
      if (nl) { out.append(s, nl - s); pos += (nl - s) + 1; if (!out.empty() && out.back() == '\r') out.pop_back(); return true; }
      out.append(s, len - pos); pos = len;
    }
  }
  // 1 record, 0 end of file, -1 malformed
  int next(std::string& seq, std::string& qual, std::string& tmp) {
    do { if (!line(tmp)) return 0; } while (tmp.empty());
    if (tmp[0] != '@') return -1;
    seq.clear(); qual.clear();
    for (;;) { if (!line(tmp)) return -1; if (!tmp.empty() && tmp[0] == '+') break; seq += tmp; }
    while (qual.size() < seq.size()) { if (!line(tmp)) return -1; qual += tmp; }
    if (qual.size() != seq.size()) return -1;
    return 1;
  }
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

struct Pinned {
  u8* p = nullptr; size_t cap = 0;
  ~Pinned() { nb_host_free(p); }
  bool ensure(size_t n, size_t keep) { if (n <= cap) return true; size_t nc = std::max(n, cap * 2 + (1 << 20)); u8* q = (u8*)nb_host_alloc(nc); if (!q) return false; if (keep) memcpy(q, p, keep); nb_host_free(p); p = q; cap = nc; return true; }
};

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
    FastqReader r1, r2; bool paired = n_inputs > 1;
    if (rc == NB_OK && !r1.open(input_files[0])) rc = fail(NB_ERR_IO, std::string("could not open ") + input_files[0]);
    if (rc == NB_OK && paired && !r2.open(input_files[1])) rc = fail(NB_ERR_IO, std::string("could not open ") + input_files[1]);
    // two pinned buffer sets: the library may still be copying set i while set i^1 is filled
    Pinned seq[2][2]; std::vector<u64> off[2][2];
    std::string s, q, tmp; int cur = 0; bool done = false;
    while (rc == NB_OK && !done) {
      size_t used[2] = {0, 0}; u64 n = 0; u32 maxlen = 0;
      off[cur][0].assign(1, 0); off[cur][1].assign(1, 0);
      while (n < BATCH) {
        int g = r1.next(s, q, tmp);
        if (g == 0) { done = true; if (paired && r2.next(s, q, tmp) == 1) rc = fail(NB_ERR_PARSE, "Error -- read and reverse read files do not have matching lengths: "); break; }
        if (g < 0) { rc = fail(NB_ERR_PARSE, "Error -- could not parse read. Input R1 data malformed."); break; }
        if (!seq[cur][0].ensure(used[0] + s.size(), used[0])) { rc = fail(NB_ERR_CUDA, "pinned allocation failed"); break; }
        memcpy(seq[cur][0].p + used[0], s.data(), s.size()); used[0] += s.size(); off[cur][0].push_back(used[0]); maxlen = std::max<u32>(maxlen, (u32)s.size());
        if (paired) {
          g = r2.next(s, q, tmp);
          if (g == 0) { rc = fail(NB_ERR_PARSE, "Error -- read and reverse read files do not have matching lengths: "); break; }
          if (g < 0) { rc = fail(NB_ERR_PARSE, "Error -- could not parse reverse read. Input R2 data malformed."); break; }
          if (!seq[cur][1].ensure(used[1] + s.size(), used[1])) { rc = fail(NB_ERR_CUDA, "pinned allocation failed"); break; }
          memcpy(seq[cur][1].p + used[1], s.data(), s.size()); used[1] += s.size(); off[cur][1].push_back(used[1]); maxlen = std::max<u32>(maxlen, (u32)s.size());
        }
        n++;
      }
      if (rc != NB_OK || n == 0) break;
      nb_batch b; memset(&b, 0, sizeof b);
      b.n_pairs = n; b.location = NB_MEM_HOST; b.max_read_len = maxlen;
      b.r1 = seq[cur][0].p; b.r1_off = off[cur][0].data();
      if (paired) { b.r2 = seq[cur][1].p; b.r2_off = off[cur][1].data(); }
      rc = nb_align_batch(ctx, &b, nullptr, nullptr);
      cur ^= 1;
      if (rc == NB_OK && cur == 0) rc = nb_ctx_sync(ctx);   // both sets in flight: wait before refilling set 0
    }
    nb_counts cts;
    if (rc == NB_OK) rc = nb_counts_finalize(ctx, &cts);
    if (rc == NB_OK) rc = write_tsv(output_paths[li], lib, cts);
    nb_ctx_free(ctx); nb_index_free(ix); nb_library_free(lib);
    if (rc != NB_OK) return rc;
  }
  return NB_OK;
}
