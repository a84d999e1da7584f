// shim/align_simd_vsg.cpp — seam 1 of the drop-in boundary: a replacement for the
// reference's src/core/align_simd.cpp that keeps its exact C++ call surface
//     search16_init / search16_exit / search16_qprep / search16      (src/core/align_simd.hpp:76-108)
// and forwards the work to libvsg.so (include/vsg.h).  Compile it against the reference's own headers
// and link it INSTEAD of align_simd.cpp.o; nothing else in the reference changes.
// Callers that reach this seam: core/searchcore.cpp:768,892; core/search.cpp:147,168;
// core/cluster.cpp:217,239,743; commands/allpairs_global.cpp:351,420,422,563; core/chimera.cpp:1899-2078.
//
// Ownership and error conventions are the reference's: CIGARs are returned as xmalloc'd C strings the
// caller xfree()s (align_simd.cpp:1855-1857), "cannot align" is the SHRT_MAX sentinel with an empty
// CIGAR (align_simd.cpp:1838-1846), unrecoverable conditions end in fatal() (utils/fatal.cpp:67).
#include "vsearch.h"
#include "core/align_simd.hpp"
#include "core/db.hpp"
#include "utils/fatal.hpp"

#include "vsg.h"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

struct s16info_s {
  vsg_ctx * ctx = nullptr;
  vsg_seqset * query = nullptr;
  char * qseq = nullptr;
  int qlen = 0;
};

namespace {

// The reference's Database is shared read-only by all worker threads (LIBRARY_API.md:962-999): keep
// one device-resident copy per Database object, re-uploaded if the object is seen to have changed
// (cluster_* appends centroids between rounds).  Mirrors are reference counted: a search16 call keeps
// the one it started with alive even if another thread replaces it meanwhile.
struct DbMirror {
  Database const * db = nullptr;
  uint64_t count = 0;
  char const * first = nullptr;
  uint64_t nucleotides = 0;
  vsg_seqset * set = nullptr;
  ~DbMirror() { if (set != nullptr) { vsg_seqset_destroy(set); } }
};
std::mutex g_mutex;
std::shared_ptr<DbMirror> g_mirror;

int device_ordinal()
{
  const char * e = std::getenv("VSG_DEVICE");
  return e != nullptr ? std::atoi(e) : 0;
}

void die(const char * what)
{
  std::string const m = std::string(what) + ": " + vsg_last_error();
  fatal(m.c_str());
}

std::shared_ptr<DbMirror> mirror_of(vsg_ctx * ctx, Database const & db)
{
  std::lock_guard<std::mutex> const lock(g_mutex);
  uint64_t const n = db.getsequencecount();
  char const * const first = n > 0 ? db.getsequence(0) : nullptr;
  if (g_mirror && g_mirror->db == &db && g_mirror->count == n && g_mirror->first == first &&
      g_mirror->nucleotides == db.getnucleotidecount()) {
    return g_mirror;
  }
  std::vector<int64_t> off(n);
  std::vector<int32_t> len(n);
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; i++) { off[i] = static_cast<int64_t>(total); len[i] = static_cast<int32_t>(db.getsequencelen(i)); total += db.getsequencelen(i); }
  std::vector<char> cat(total + 1);
  for (uint64_t i = 0; i < n; i++) { std::memcpy(cat.data() + off[i], db.getsequence(i), static_cast<size_t>(len[i])); }
  auto m = std::make_shared<DbMirror>();
  if (vsg_seqset_create(ctx, cat.data(), off.data(), len.data(), static_cast<int64_t>(n), 1, &m->set) != VSG_OK) { die("vsg_seqset_create"); }
  m->db = &db; m->count = n; m->first = first; m->nucleotides = db.getnucleotidecount();
  g_mirror = m;   // the previous mirror goes away when its last user lets go of it
  return m;
}

}  // namespace

auto search16_init(int64_t score_match, int64_t score_mismatch,
                   int64_t penalty_gap_open_query_left, int64_t penalty_gap_open_target_left,
                   int64_t penalty_gap_open_query_interior, int64_t penalty_gap_open_target_interior,
                   int64_t penalty_gap_open_query_right, int64_t penalty_gap_open_target_right,
                   int64_t penalty_gap_extension_query_left, int64_t penalty_gap_extension_target_left,
                   int64_t penalty_gap_extension_query_interior, int64_t penalty_gap_extension_target_interior,
                   int64_t penalty_gap_extension_query_right, int64_t penalty_gap_extension_target_right,
                   bool score_n_mismatch) -> struct s16info_s *
{
  vsg_scoring sc;
  int64_t const v[14] = {score_match, score_mismatch,
                         penalty_gap_open_query_left, penalty_gap_open_target_left,
                         penalty_gap_open_query_interior, penalty_gap_open_target_interior,
                         penalty_gap_open_query_right, penalty_gap_open_target_right,
                         penalty_gap_extension_query_left, penalty_gap_extension_target_left,
                         penalty_gap_extension_query_interior, penalty_gap_extension_target_interior,
                         penalty_gap_extension_query_right, penalty_gap_extension_target_right};
  std::memcpy(sc.v, v, sizeof v);
  sc.n_mismatch = score_n_mismatch ? 1 : 0;
  auto * s = new s16info_s();
  if (vsg_ctx_create(device_ordinal(), &sc, &s->ctx) != VSG_OK) { die("vsg_ctx_create"); }
  return s;
}

auto search16_exit(s16info_s * s) -> void
{
  if (s->query != nullptr) { vsg_seqset_destroy(s->query); }
  vsg_ctx_destroy(s->ctx);
  delete s;
}

auto search16_qprep(s16info_s * s, char * qseq, int qlen) -> void
{
  s->qseq = qseq;
  s->qlen = qlen;
  if (s->query != nullptr) { vsg_seqset_destroy(s->query); s->query = nullptr; }
  int64_t const off = 0;
  int32_t const len = qlen;
  if (vsg_seqset_create(s->ctx, qseq, &off, &len, 1, 1, &s->query) != VSG_OK) { die("vsg_seqset_create"); }
}

auto search16(s16info_s * s, unsigned int sequences, unsigned int const * seqnos,
              CELL * pscores, unsigned short * paligned, unsigned short * pmatches,
              unsigned short * pmismatches, unsigned short * pgaps, char ** pcigar,
              struct Database const & db) -> void
{
  if (sequences == 0) { return; }
  std::shared_ptr<DbMirror> const mirror = mirror_of(s->ctx, db);   // held for the duration of the call
  vsg_seqset * const targets = mirror->set;
  std::vector<uint32_t> qidx(sequences, 0U);
  int64_t cap = 16;
  for (unsigned int i = 0; i < sequences; i++) { cap += static_cast<int64_t>(s->qlen) + static_cast<int64_t>(db.getsequencelen(seqnos[i])) + 2; }
  std::vector<char> cig(static_cast<size_t>(cap));
  std::vector<int64_t> coff(static_cast<size_t>(sequences) + 1);
  if (vsg_align_pairs(s->ctx, s->query, targets, sequences, qidx.data(), seqnos, pscores, paligned, pmatches,
                      pmismatches, pgaps, nullptr, cig.data(), cap, coff.data()) != VSG_OK) {
    die("vsg_align_pairs");
  }
  for (unsigned int i = 0; i < sequences; i++) {
    char const * const c = cig.data() + coff[i];
    size_t const l = std::strlen(c);
    pcigar[i] = static_cast<char *>(xmalloc(l + 1));
    std::memcpy(pcigar[i], c, l + 1);
  }
}
