// Sanitizer harness of the file drivers' inflate (csrc/inflate.hpp, csrc/pgunzip.hpp): every file given on the command line is
// decoded by zlib, by the parallel reader (1 and 3 threads, small chunks) and by the windowed serial reader, then damaged copies
// of it go through all three — any outcome is fine for those except a crash, a sanitizer report, text that differs from zlib's
// when both accept, or the parallel and the serial reader disagreeing.  Built and run by tests/test_inflate_cpu.py with
// -fsanitize=address,undefined (and by hand with -fsanitize=thread).
#include "../../nimble_aligner_b200/csrc/pgunzip.hpp"
#include "../../nimble_aligner_b200/csrc/deflate_fast.hpp"
#include <cstdio>
#include <random>
#include <string>
using namespace nbz;
static std::vector<u8> slurp(const char* p) { FILE* f = fopen(p, "rb"); std::vector<u8> v; u8 b[65536]; size_t n; while ((n = fread(b, 1, sizeof b, f)) > 0) v.insert(v.end(), b, b + n); fclose(f); return v; }
static bool zref(const std::vector<u8>& in, std::vector<u8>& out) {   // zlib, concatenated members, trailing garbage ignored after >= 1 member
  out.clear(); size_t pos = 0; int members = 0;
  while (pos < in.size()) {
    z_stream zs; memset(&zs, 0, sizeof zs); if (inflateInit2(&zs, 31) != Z_OK) return false;
    zs.next_in = (Bytef*)in.data() + pos; zs.avail_in = (uInt)(in.size() - pos); int rc;
    do { u8 buf[1 << 16]; zs.next_out = buf; zs.avail_out = sizeof buf; rc = inflate(&zs, Z_NO_FLUSH); if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); return members > 0 && zs.total_out == 0 && false; } out.insert(out.end(), buf, buf + (sizeof buf - zs.avail_out)); } while (rc != Z_STREAM_END);
    pos = in.size() - zs.avail_in; inflateEnd(&zs); members++;
    if (pos + 2 > in.size() || in[pos] != 0x1f || in[pos + 1] != 0x8b) break;
  }
  return members > 0;
}
static int ours_par(const std::vector<u8>& in, int T, size_t C, std::vector<u8>& out) {
  out.clear(); ParallelGunzip pg; pg.open(in.data(), in.size(), T, C);
  while (PgChunk* c = pg.next()) { if (c->status < 0) return -1; out.insert(out.end(), c->text.begin(), c->text.end()); pg.recycle(c); }
  return 0;
}
static int ours_ser(const std::vector<u8>& in, size_t window, std::vector<u8>& out) {
  out.clear(); std::unique_ptr<GzipStream> gz(new GzipStream()); gz->open(in.data(), in.size());
  const size_t H = 32768; std::vector<u8> a(H + window), b(H + window); u8* cur = a.data(); u8* prev = nullptr; size_t pl = 0;
  while (!gz->done()) { if (prev) memcpy(cur, prev + pl, H); ptrdiff_t n = gz->read(prev ? cur : cur + H, cur + H, cur + H + window); if (n < 0) return -1; out.insert(out.end(), cur + H, cur + H + n); prev = cur; pl = n; cur = cur == a.data() ? b.data() : a.data(); }
  return 0;
}
int main(int argc, char** argv) {
  std::mt19937 rng(7); int bad = 0; long fuzz_err = 0, fuzz_ok = 0;
  for (int i = 1; i < argc; i++) {
    std::vector<u8> in = slurp(argv[i]), want, got;
    bool ok = zref(in, want);
    for (int T : {1, 3}) for (size_t C : {(size_t)4096, (size_t)30000}) { int r = ours_par(in, T, C, got); if ((r == 0) != ok || (ok && got != want)) { printf("%s: parallel T=%d C=%zu mismatch (r=%d ok=%d)\n", argv[i], T, C, r, ok); bad++; } }
    for (size_t w : {(size_t)300, (size_t)5000, (size_t)(1 << 20)}) { int r = ours_ser(in, w, got); if ((r == 0) != ok || (ok && got != want)) { printf("%s: serial window %zu mismatch\n", argv[i], w); bad++; } }
    if (ok) {   // the BAM driver's compressor on the same text: zlib and the serial reader must give it back
      std::string z; FastDeflate fd; fd.gzip_member(want.data(), want.size(), z);
      std::vector<u8> zin(z.begin(), z.end()), back;
      if (!zref(zin, back) || back != want) { printf("%s: fast deflate -> zlib mismatch\n", argv[i]); bad++; }
      if (ours_ser(zin, 70000, got) != 0 || got != want) { printf("%s: fast deflate -> own inflate mismatch\n", argv[i]); bad++; }
    }
    if (in.size() > 5000000 || want.size() > 5000000) continue;
    for (int f = 0; f < 12; f++) {      // damaged copies: any outcome but a crash / sanitizer report / different text than zlib when both accept
      std::vector<u8> d = in; if (f % 3 == 0) d.resize(rng() % d.size()); else for (int k = 0; k < 1 + (int)(rng() % 3); k++) d[rng() % d.size()] = (u8)rng();
      std::vector<u8> w2; bool zok = zref(d, w2);
      int r = ours_par(d, 3, 4096, got); if (r == 0 && zok && got != w2) { printf("%s: fuzz %d parallel differs from zlib\n", argv[i], f); bad++; }
      int r2 = ours_ser(d, 4096, got); if (r2 == 0 && zok && got != w2) { printf("%s: fuzz %d serial differs from zlib\n", argv[i], f); bad++; }
      if (r == 0) fuzz_ok++; else fuzz_err++;
      if ((r == 0) != (r2 == 0)) { printf("%s: fuzz %d parallel and serial disagree (%d %d)\n", argv[i], f, r, r2); bad++; }
    }
  }
  printf("done: %d problems; fuzz accepted %ld rejected %ld\n", bad, fuzz_ok, fuzz_err);
  return bad != 0;
}
