// library.cpp — nimble reference-library JSON -> Reference + AlignFilterConfig, and the integer roll-up tables.
// Mirrors reference_library::get_reference_library (/root/reference/src/reference_library.rs:20-226).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

#include "host.hpp"

namespace nb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) { g_err = msg; return code; }

// lexical-sort 0.3.1 natural_lexical_cmp (src/align.rs:846), ASCII restatement: case-folded comparison, digit runs
// compared by numeric value, remaining ties broken by byte order so the order is total.
int natural_lexical_cmp(const std::string& a, const std::string& b) {
  auto lower = [](unsigned char c) { return (unsigned char)((c >= 'A' && c <= 'Z') ? c + 32 : c); };
  auto digit = [](unsigned char c) { return c >= '0' && c <= '9'; };
  size_t i = 0, j = 0, na = a.size(), nb_ = b.size();
  while (i < na && j < nb_) {
    unsigned char x = (unsigned char)a[i], y = (unsigned char)b[j];
    if (digit(x) && digit(y)) {
      while (i < na && a[i] == '0') i++;
      while (j < nb_ && b[j] == '0') j++;
      size_t ie = i, je = j;
      while (ie < na && digit((unsigned char)a[ie])) ie++;
      while (je < nb_ && digit((unsigned char)b[je])) je++;
      if (ie - i != je - j) return (ie - i) < (je - j) ? -1 : 1;
      for (size_t k = 0; k < ie - i; k++)
        if (a[i + k] != b[j + k]) return a[i + k] < b[j + k] ? -1 : 1;
      i = ie; j = je;
      continue;
    }
    unsigned char fx = lower(x), fy = lower(y);
    if (fx != fy) return fx < fy ? -1 : 1;
    i++; j++;
  }
  if (i < na) return 1;
  if (j < nb_) return -1;
  int c = a.compare(b);
  return c < 0 ? -1 : (c > 0 ? 1 : 0);
}

// ------------------------------------------------------------------ minimal JSON (serde_json::Value subset)
struct JVal {
  enum T { Null, Bool, Int, Float, Str, Arr, Obj } t = Null;
  bool b = false; i64 i = 0; double f = 0; std::string s;
  std::vector<JVal> a; std::vector<std::pair<std::string, JVal>> o;
  const JVal& key(const char* k) const { static const JVal null; if (t != Obj) return null; for (auto& kv : o) if (kv.first == k) return kv.second; return null; }
  JVal& mkey(const char* k) { static thread_local JVal null; null = JVal(); if (t != Obj) return null; for (auto& kv : o) if (kv.first == k) return kv.second; return null; }
  const JVal& at(size_t n) const { static const JVal null; if (t != Arr || n >= a.size()) return null; return a[n]; }
};
struct JParser {
  const char* p; const char* e; std::string err;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool fail(const char* m) { if (err.empty()) err = m; return false; }
  static void utf8(std::string& s, unsigned cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
  }
  bool hex4(unsigned& v) { if (e - p < 4) return fail("bad \\u escape"); v = 0; for (int k = 0; k < 4; k++) { char c = *p++; v <<= 4; if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10; else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else return fail("bad \\u escape"); } return true; }
  bool str(std::string& out) {
    if (p >= e || *p != '"') return fail("expected string");
    p++;
    while (p < e && *p != '"') {
      if (*p == '\\') {
        p++; if (p >= e) return fail("bad escape");
        char c = *p++;
        switch (c) {
          case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break; case 'r': out += '\r'; break; case 't': out += '\t'; break;
          case 'u': { unsigned v; if (!hex4(v)) return false;
            if (v >= 0xD800 && v < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; unsigned lo; if (!hex4(lo)) return false; v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00); }
            utf8(out, v); break; }
          default: return fail("bad escape");
        }
      } else {   // append the whole run up to the next quote or escape (a sequence value is one run of a kilobase)
        const char* q = p;
        while (q < e && *q != '"' && *q != '\\') q++;
        out.append(p, q - p); p = q;
      }
    }
    if (p >= e) return fail("unterminated string");
    p++; return true;
  }
  bool val(JVal& v, int depth = 0) {
    if (depth > 64) return fail("nesting too deep");
    ws(); if (p >= e) return fail("unexpected end");
    char c = *p;
    if (c == '{') {
      v.t = JVal::Obj; p++; ws();
      if (p < e && *p == '}') { p++; return true; }
      for (;;) { ws(); std::string k; if (!str(k)) return false; ws(); if (p >= e || *p != ':') return fail("expected ':'"); p++; JVal x; if (!val(x, depth + 1)) return false; v.o.emplace_back(std::move(k), std::move(x)); ws(); if (p < e && *p == ',') { p++; continue; } if (p < e && *p == '}') { p++; return true; } return fail("expected ',' or '}'"); }
    }
    if (c == '[') {
      v.t = JVal::Arr; p++; ws();
      if (p < e && *p == ']') { p++; return true; }
      for (;;) { JVal x; if (!val(x, depth + 1)) return false; v.a.push_back(std::move(x)); ws(); if (p < e && *p == ',') { p++; continue; } if (p < e && *p == ']') { p++; return true; } return fail("expected ',' or ']'"); }
    }
    if (c == '"') { v.t = JVal::Str; return str(v.s); }
    if (e - p >= 4 && !strncmp(p, "true", 4)) { v.t = JVal::Bool; v.b = true; p += 4; return true; }
    if (e - p >= 5 && !strncmp(p, "false", 5)) { v.t = JVal::Bool; v.b = false; p += 5; return true; }
    if (e - p >= 4 && !strncmp(p, "null", 4)) { v.t = JVal::Null; p += 4; return true; }
    if (c == '-' || (c >= '0' && c <= '9')) {
      const char* s = p; bool is_float = false;
      if (*p == '-') p++;
      while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) { if (*p == '.' || *p == 'e' || *p == 'E') is_float = true; p++; }
      std::string num(s, p);
      if (is_float) { v.t = JVal::Float; v.f = strtod(num.c_str(), nullptr); }
      else { v.t = JVal::Int; errno = 0; v.i = strtoll(num.c_str(), nullptr, 10); if (errno) { v.t = JVal::Float; v.f = strtod(num.c_str(), nullptr); } }
      return true;
    }
    return fail("unexpected character");
  }
};

// utils::revcomp, src/utils.rs:61-94 (panics on non-DNA; N and others complement to 'N')
static bool revcomp(const std::string& s, std::string& out, std::string& err) {
  static const struct Tab { char t[256]; Tab() { memset(t, 0, sizeof t); const char* a = "actguACTGUNn"; const char* b = "tgacaTGACANN"; for (int i = 0; a[i]; i++) t[(unsigned char)a[i]] = b[i]; } } tab;
  size_t n = s.size();
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    char c = s[n - 1 - i], r = tab.t[(unsigned char)c];
    if (!r) { err = std::string("Input sequence base is not DNA: ") + c; return false; }
    out[i] = r;
  }
  return true;
}

static int sanity(const nb_config& c) {  // src/reference_library.rs:209-226
  if (!(c.score_percent >= 0.0 && c.score_percent <= 1.0)) return fail(NB_ERR_CONFIG, "Error -- score_percent must be between 0 and 1");
  if (c.score_filter < 0) return fail(NB_ERR_CONFIG, "Error -- score_filter must be positive");
  if (!(c.trim_strictness >= 0.0 && c.trim_strictness <= 1.0)) return fail(NB_ERR_CONFIG, "Error -- trim_strictness must be between 0 and 1");
  return NB_OK;
}

// fs::read_to_string (src/reference_library.rs:21) refuses a file that is not UTF-8: so does this loader
static bool valid_utf8(const char* t, size_t n) {
  const unsigned char* p = (const unsigned char*)t; const unsigned char* e = p + n;
  while (p < e) {
    if (e - p >= 8) { u64 w; memcpy(&w, p, 8); if (!(w & 0x8080808080808080ULL)) { p += 8; continue; } }
    unsigned c = *p;
    if (c < 0x80) { p++; continue; }
    int need; unsigned cp, lo;
    if ((c & 0xE0) == 0xC0) { need = 1; cp = c & 0x1F; lo = 0x80; }
    else if ((c & 0xF0) == 0xE0) { need = 2; cp = c & 0x0F; lo = 0x800; }
    else if ((c & 0xF8) == 0xF0) { need = 3; cp = c & 0x07; lo = 0x10000; }
    else return false;
    if (e - p <= need) return false;
    for (int i = 1; i <= need; i++) { if ((p[i] & 0xC0) != 0x80) return false; cp = (cp << 6) | (p[i] & 0x3F); }
    if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp < 0xE000)) return false;
    p += need + 1;
  }
  return true;
}

static int parse_library(const char* text, size_t len, int strand_filter, nb_library** out) {   // (string values are moved, not copied, out of the parsed tree: a 200 k-transcript library is 230 MB of them)
  if (strand_filter < 0 || strand_filter > 3) return fail(NB_ERR_INVALID, "Could not parse strand_filter option.");
  if (!valid_utf8(text, len)) return fail(NB_ERR_PARSE, "Error -- could not read reference library: stream did not contain valid UTF-8");
  JParser jp{text, text + len, ""};
  JVal v;
  if (!jp.val(v)) return fail(NB_ERR_PARSE, "Error -- could not parse reference library JSON: " + jp.err);
  jp.ws();
  if (jp.p != jp.e) return fail(NB_ERR_PARSE, "Error -- could not parse reference library JSON: trailing characters");
  const JVal& c = v.at(0);
  nb_config cfg; memset(&cfg, 0, sizeof cfg);
  auto as_f64 = [&](const char* k, double& o) { const JVal& x = c.key(k); if (x.t == JVal::Float) o = x.f; else if (x.t == JVal::Int) o = (double)x.i; else return false; return true; };
  auto as_i64 = [&](const char* k, i64& o) { const JVal& x = c.key(k); if (x.t != JVal::Int) return false; o = x.i; return true; };
  auto as_bool = [&](const char* k, bool& o) { const JVal& x = c.key(k); if (x.t != JVal::Bool) return false; o = x.b; return true; };
  i64 iv; bool bv;
  if (!as_f64("score_percent", cfg.score_percent)) return fail(NB_ERR_PARSE, "Error -- could not parse score_percent as f64");
  if (!as_i64("score_filter", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse score_filter as int64");
  cfg.score_filter = (int32_t)iv;
  if (!as_i64("score_threshold", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse score_threshold as int64");
  cfg.score_threshold = (u64)iv;
  if (!as_i64("num_mismatches", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse num_mismatches as int64");
  cfg.num_mismatches = (u64)iv;
  if (!as_bool("discard_multiple_matches", bv)) return fail(NB_ERR_PARSE, "Error -- could not parse discard_multiple_mismatches as boolean");
  cfg.discard_multiple_matches = bv;
  if (!as_bool("require_valid_pair", bv)) return fail(NB_ERR_PARSE, "Error -- could not parse require_valid_pair as boolean");
  cfg.require_valid_pair = bv;
  if (!as_i64("discard_multi_hits", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse discard_multi_hits as int64");
  cfg.discard_multi_hits = (u64)iv;
  if (!as_i64("intersect_level", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse intersect_level as int64");
  i64 intersect = iv;
  if (!as_i64("max_hits_to_report", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse max_hits_to_report as int64");
  cfg.max_hits_to_report = (u64)iv;
  if (intersect < 0 || intersect > 2) return fail(NB_ERR_PARSE, "Error -- invalid intersect level in config file. Please choose intersect level 0, 1, or 2.");
  cfg.intersect_level = (int32_t)intersect;
  const JVal& g = c.key("group_on");
  if (g.t != JVal::Str) return fail(NB_ERR_PARSE, "Error -- could not parse group_on as string");
  std::string group_on = g.s;
  if (!as_i64("trim_target_length", iv)) return fail(NB_ERR_PARSE, "Error -- could not parse trim_target_length as usize");
  cfg.trim_target_length = (u64)iv;
  if (!as_f64("trim_strictness", cfg.trim_strictness)) return fail(NB_ERR_PARSE, "Error -- could not parse trim_strictness as f64");
  cfg.strand_filter = strand_filter;
  cfg.discard_nonzero_mismatch = 0;  // src/reference_library.rs:116

  static thread_local JVal jnull; jnull = JVal();
  JVal& r = (v.t == JVal::Arr && v.a.size() > 1) ? v.a[1] : jnull;
  auto to_strs = [&](JVal& x, const char* name, std::vector<std::string>& o) -> bool {
    if (x.t != JVal::Arr) { set_error(std::string("Error -- could not parse ") + name + " as array"); return false; }
    o.reserve(x.a.size());
    for (auto& s : x.a) { if (s.t != JVal::Str) { set_error(std::string("Error -- could not parse ") + name + " element as a string"); return false; } o.push_back(std::move(s.s)); }
    std::vector<JVal>().swap(x.a);
    return true;
  };
  std::vector<std::string> headers;
  if (!to_strs(r.mkey("headers"), "headers", headers)) return NB_ERR_PARSE;
  auto col_index = [&](const std::string& h) -> int { for (size_t i = 0; i < headers.size(); i++) if (headers[i] == h) return (int)i; return -1; };
  int name_idx = col_index("sequence_name");
  if (name_idx < 0) return fail(NB_ERR_PARSE, "Could not find header sequence_name");
  int gidx = name_idx;
  if (!group_on.empty()) { gidx = col_index(group_on); if (gidx < 0) return fail(NB_ERR_PARSE, "Error -- could not find column for group_on " + group_on); }
  int seq_idx = col_index("sequence");
  if (seq_idx < 0) return fail(NB_ERR_PARSE, "Error -- could not find sequences column");
  JVal& cols = r.mkey("columns");
  if (cols.t != JVal::Arr) return fail(NB_ERR_PARSE, "Error -- could not parse columns as array");
  std::vector<std::vector<std::string>> columns;
  for (auto& cj : cols.a) { columns.emplace_back(); if (!to_strs(cj, "column", columns.back())) return NB_ERR_PARSE; }
  if (columns.size() < headers.size() || columns.empty()) return fail(NB_ERR_PARSE, "Error -- fewer columns than headers");
  size_t n_rows = columns[0].size();
  for (auto& cc : columns) if (cc.size() != n_rows) return fail(NB_ERR_PARSE, "Error -- columns have different lengths");
  cfg.reference_genome_size = columns[name_idx].size();

  nb_library* lib = new nb_library();
  lib->cfg = cfg; lib->headers = headers; lib->group_on = (u32)gidx; lib->name_idx = (u32)name_idx; lib->seq_idx = (u32)seq_idx;
  lib->columns.assign(columns.size(), std::vector<std::string>());
  for (auto& cc : lib->columns) cc.reserve(2 * n_rows);
  std::string rc, err;
  for (size_t row = 0; row < n_rows; row++) {  // src/reference_library.rs:130-153
    std::string& seq = columns[seq_idx][row];
    for (char& ch : seq) { if (ch == 'U') ch = 'T'; else if (ch == 'u') ch = 't'; }
    if (!revcomp(seq, rc, err)) { delete lib; return fail(NB_ERR_PARSE, err); }
    for (size_t ci = 0; ci < columns.size(); ci++) {
      std::vector<std::string>& dst = lib->columns[ci];
      if ((int)ci == seq_idx) { dst.push_back(std::move(seq)); dst.push_back(rc); }
      else if ((int)ci == name_idx) { dst.push_back(columns[ci][row] + "\xC2\xA7rev"); dst.push_back(std::move(columns[ci][row])); std::swap(dst[dst.size() - 2], dst[dst.size() - 1]); }
      else { dst.push_back(columns[ci][row]); dst.push_back(std::move(columns[ci][row])); }
    }
  }
  int rcode = sanity(cfg);
  if (rcode != NB_OK) { delete lib; return rcode; }
  *out = lib;
  return NB_OK;
}

}  // namespace nb

using namespace nb;

static bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

// Integer tables replacing the per-pair string work: filter_read_calls_with_orientation's base name (strip_suffix
// "§rev", src/align.rs:149), parse_calls' (base, is_rev) (src/align.rs:276-285), unmap's first-row lookup
// (src/align.rs:851-864) and the roll-up group string (src/align.rs:810-836) ranked by natural_lexical_cmp.
void nb_library::finalize() {
  const std::vector<std::string>& names = columns[name_idx];
  const std::vector<std::string>& groups = columns[group_on];
  u32 n = (u32)names.size();
  no_dedup = headers[group_on] == "nt_sequence";
  injective = true; irregular_reason.clear();
  std::unordered_map<std::string, u32> fid_of, first_row;
  row_fid.assign(n, 0); row_rev.assign(n, 0);
  std::vector<std::string> fnames;
  for (u32 r = 0; r < n; r++) first_row.emplace(names[r], r);
  for (u32 r = 0; r < n; r++) {
    const std::string& nm = names[r];
    std::string b1 = ends_with(nm, "\xC2\xA7rev") ? nm.substr(0, nm.size() - 5) : nm;
    std::string b2 = nm; bool rev = false;
    if (ends_with(nm, "rev")) { rev = true; while (ends_with(b2, "rev")) b2.resize(b2.size() - 3); while (ends_with(b2, "\xC2\xA7")) b2.resize(b2.size() - 2); }
    if (b1 != b2 && injective) { injective = false; irregular_reason = "sequence_name '" + nm + "' ends in 'rev' without the single \xC2\xA7rev suffix (parse_calls and strip_suffix disagree)"; }
    auto it = fid_of.find(b2);
    u32 f;
    if (it == fid_of.end()) { f = (u32)fnames.size(); fid_of.emplace(b2, f); fnames.push_back(b2); } else f = it->second;
    row_fid[r] = f; row_rev[r] = rev;
  }
  n_features = (u32)fnames.size();
  row_of.assign(2 * (size_t)n_features, NONE32);
  for (u32 r = 0; r < n; r++) {
    u32& slot = row_of[2 * (size_t)row_fid[r] + row_rev[r]];
    if (slot != NONE32) { if (injective) { injective = false; irregular_reason = "duplicate sequence_name '" + names[r] + "'"; } }
    else slot = r;
  }
  // group strings of the row unmap() finds for each feature
  std::vector<std::string> gs(n_features); std::vector<char> has(n_features, 0);
  std::map<std::string, u32> distinct;
  for (u32 f = 0; f < n_features; f++) {
    auto it = first_row.find(fnames[f]);
    if (it == first_row.end()) continue;
    u32 r = it->second;
    // nt_sequence header: one-to-one translation from the name column (src/align.rs:810-816); else the group_on
    // value, falling back to the feature name when it is empty (src/align.rs:822-829)
    gs[f] = no_dedup ? names[r] : (groups[r].empty() ? names[r] : groups[r]);
    has[f] = 1; distinct.emplace(gs[f], 0);
  }
  group_names.clear();
  for (auto& kv : distinct) group_names.push_back(kv.first);
  std::sort(group_names.begin(), group_names.end(), [](const std::string& a, const std::string& b) { return natural_lexical_cmp(a, b) < 0; });
  for (u32 i = 0; i < group_names.size(); i++) distinct[group_names[i]] = i;
  {  // rank of every group name in plain byte order (Vec<String> Ord, utils::sort_score_vector): callsets sort on integers
    std::vector<u32> ord(group_names.size()); for (u32 i = 0; i < ord.size(); i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](u32 a, u32 b) { return group_names[a] < group_names[b]; });
    group_byte_rank.assign(group_names.size(), 0); for (u32 i = 0; i < ord.size(); i++) group_byte_rank[ord[i]] = i;
  }
  feat_group.assign(n_features, NONE32);
  for (u32 f = 0; f < n_features; f++) if (has[f]) feat_group[f] = distinct[gs[f]];
  derived = true;
}

extern "C" {

const char* nb_last_error(void) { return nb::g_err.c_str(); }
const char* nb_version(void) { return "nimble_b200 0.1 (sm_100a)"; }
const char* nb_reason_str(int r) {
  static const char* s[] = {"Score Below Threshold", "Discarded Multiple Match", "Discarded Nonzero Mismatch", "No Match",
    "No Match and Score Below Threshold", "Different Filter Reasons", "Required Valid Pair Not Matching", "Force Intersect Failure",
    "Short Read", "Max Hits Exceeded", "Low Entropy", "Successful Match", "Strandedness Filtered", "Equivalence Class Empty After Filters",
    "Above Mismatch Threshold", "SKipped Align Due To Unpaired Dummy Read", "None"};
  return (r >= 0 && r <= 16) ? s[r] : "?";
}

int nb_library_parse_json(const char* text, size_t len, int strand_filter, nb_library** out) {
  if (!text || !out) return fail(NB_ERR_INVALID, "null argument");
  return parse_library(text, len, strand_filter, out);
}
int nb_library_load_json(const char* path, int strand_filter, nb_library** out) {
  if (!path || !out) return fail(NB_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(NB_ERR_IO, std::string("Error -- could not read reference library ") + path);
  std::string s; char buf[1 << 16]; size_t got;
  if (fseek(f, 0, SEEK_END) == 0) { long sz = ftell(f); if (sz > 0) s.reserve((size_t)sz); rewind(f); }
  while ((got = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, got);
  bool bad = ferror(f) != 0; fclose(f);
  if (bad) return fail(NB_ERR_IO, std::string("Error -- could not read reference library ") + path);
  return parse_library(s.data(), s.size(), strand_filter, out);
}
int nb_library_from_columns(const char* const* headers, uint32_t n_headers, const char* const* const* columns, uint32_t n_rows,
                            uint32_t group_on, uint32_t sequence_name_idx, uint32_t sequence_idx, const nb_config* cfg, nb_library** out) {
  if (!headers || !columns || !cfg || !out) return fail(NB_ERR_INVALID, "null argument");
  if (group_on >= n_headers || sequence_name_idx >= n_headers || sequence_idx >= n_headers) return fail(NB_ERR_INVALID, "column index out of range");
  nb_library* lib = new nb_library();
  lib->cfg = *cfg; lib->group_on = group_on; lib->name_idx = sequence_name_idx; lib->seq_idx = sequence_idx;
  for (uint32_t h = 0; h < n_headers; h++) { lib->headers.emplace_back(headers[h]); lib->columns.emplace_back(); for (uint32_t r = 0; r < n_rows; r++) lib->columns.back().emplace_back(columns[h][r]); }
  lib->cfg.reference_genome_size = n_rows;
  *out = lib;
  return NB_OK;
}
void nb_library_free(nb_library* l) { delete l; }
int nb_library_get_config(const nb_library* l, nb_config* out) { if (!l || !out) return fail(NB_ERR_INVALID, "null argument"); *out = l->cfg; return NB_OK; }
int nb_library_set_config(nb_library* l, const nb_config* cfg) {
  if (!l || !cfg) return fail(NB_ERR_INVALID, "null argument");
  if (cfg->intersect_level < 0 || cfg->intersect_level > 2 || cfg->strand_filter < 0 || cfg->strand_filter > 3) return fail(NB_ERR_INVALID, "invalid intersect_level / strand_filter");
  int rc = sanity(*cfg); if (rc) return rc;
  l->cfg = *cfg; return NB_OK;
}
uint32_t nb_library_n_rows(const nb_library* l) { return l ? l->n_rows() : 0; }
uint32_t nb_library_n_headers(const nb_library* l) { return l ? (uint32_t)l->headers.size() : 0; }
const char* nb_library_header(const nb_library* l, uint32_t c) { return (l && c < l->headers.size()) ? l->headers[c].c_str() : nullptr; }
const char* nb_library_value(const nb_library* l, uint32_t c, uint32_t r) { return (l && c < l->columns.size() && r < l->columns[c].size()) ? l->columns[c][r].c_str() : nullptr; }
uint32_t nb_library_group_on(const nb_library* l) { return l->group_on; }
uint32_t nb_library_sequence_name_idx(const nb_library* l) { return l->name_idx; }
uint32_t nb_library_sequence_idx(const nb_library* l) { return l->seq_idx; }
int nb_library_push_column(nb_library* l, const char* header, const char* const* values, uint32_t n_rows, int set_group_on) {
  if (!l || !header || !values) return fail(NB_ERR_INVALID, "null argument");
  if (n_rows != l->n_rows()) return fail(NB_ERR_INVALID, "column length does not match the library");
  if (l->headers.size() != l->columns.size()) return fail(NB_ERR_INVALID, "library has unnamed columns; cannot push a column");
  l->headers.emplace_back(header); l->columns.emplace_back();
  for (uint32_t r = 0; r < n_rows; r++) l->columns.back().emplace_back(values[r]);
  if (set_group_on) l->group_on = (u32)l->columns.size() - 1;
  l->derived = false;
  return NB_OK;
}
uint32_t nb_library_n_groups(const nb_library* l) { if (!l->derived) const_cast<nb_library*>(l)->finalize(); return (uint32_t)l->group_names.size(); }
const char* nb_library_group_name(const nb_library* l, uint32_t g) { if (!l->derived) const_cast<nb_library*>(l)->finalize(); return g < l->group_names.size() ? l->group_names[g].c_str() : nullptr; }

}  // extern "C"
