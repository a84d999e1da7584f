// shim/search_batch_vsg.cpp — seam 2 of the drop-in boundary: a replacement for the
// reference's library entry point
//     search_batch(Parameters const&, Dbindex const&, Database const&, seqs, heads, lens, sizes, n,
//                  search_result_s* results, max_per_query, counts)        (src/core/search.hpp:135-145,
//                                                                            src/core/search.cpp:511-593)
// with the same signature, the same result records (search.hpp:67-80) and the same error convention
// (fatal(), utils/fatal.cpp:67), forwarding the whole batch to libvsg.so (include/vsg.h): queries are
// uploaded once, DUST-masked on the device, ranked, aligned and accept/reject-replayed there
// (vsg_search_batch), and only the hit table comes back.  Compile against the reference's headers and
// link so that this definition wins over the one in core/search.cpp.o (oracle/Makefile weakens that
// symbol with objcopy; a maintainer would delete the body from search.cpp).  See INTEGRATION.md.
//
// What stays on the host, on the reference's own code: the LinearMemoryAligner for pairs the 16-bit
// aligner defers (core/searchcore.cpp:806-832) — registered as the library's fallback callback — and
// hardmask() for --qmask soft --hardmask.
#include "vsearch_api.h"
#include "core/linmemalign.hpp"
#include "core/mask.hpp"
#include "core/searchcore.hpp"
#include "utils/fatal.hpp"
#include "utils/reverse_complement.hpp"
#include "utils/string_alloc.hpp"

#include "vsg.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

[[noreturn]] void die(const char * what)
{
  std::string const m = std::string("GPU search_batch: ") + what + ": " + vsg_last_error();
  fatal(m.c_str());
  std::abort();
}

// VSG_DEVICES=0,1,2,... : the GPUs this process may use (queries are sharded across them); default: VSG_DEVICE or 0
std::vector<int> device_list()
{
  std::vector<int> d;
  if (const char * e = std::getenv("VSG_DEVICES")) {
    const char * p = e;
    while (*p != '\0') {
      d.push_back(std::atoi(p));
      while (*p != '\0' && *p != ',') { ++p; }
      if (*p == ',') { ++p; }
    }
  }
  if (d.empty()) {
    const char * e = std::getenv("VSG_DEVICE");
    d.push_back(e != nullptr ? std::atoi(e) : 0);
  }
  return d;
}

// The database and its index are shared read-only objects of the embedding application
// (LIBRARY_API.md:962-999).  One device mirror (context + packed sequences + k-mer index) is kept per
// process and rebuilt when the Database object, its contents, the index width or the scoring change.
struct Mirror {
  Database const * db = nullptr;
  uint64_t count = 0, nucleotides = 0;
  char const * first = nullptr;
  unsigned int wordlength = 0;
  int mask_lower = 0;
  vsg_scoring scoring{};
  vsg_group * group = nullptr;   // one context + database copy + index per device
  std::vector<int64_t> sizes;    // db.getabundance
  std::vector<int64_t> labels;   // header identity: equal numbers <=> equal headers
  std::unordered_map<std::string, int64_t> label_of;
  void drop()
  {
    if (group != nullptr) { vsg_group_destroy(group); group = nullptr; }
    sizes.clear(); labels.clear(); label_of.clear();
  }
};
Mirror g_mirror;
std::mutex g_mutex;

vsg_scoring scoring_of(Parameters const & p)
{
  vsg_scoring sc;
  int64_t const v[14] = {p.opt_match, p.opt_mismatch,
                         p.opt_gap_open_query_left, p.opt_gap_open_target_left,
                         p.opt_gap_open_query_interior, p.opt_gap_open_target_interior,
                         p.opt_gap_open_query_right, p.opt_gap_open_target_right,
                         p.opt_gap_extension_query_left, p.opt_gap_extension_target_left,
                         p.opt_gap_extension_query_interior, p.opt_gap_extension_target_interior,
                         p.opt_gap_extension_query_right, p.opt_gap_extension_target_right};
  std::memcpy(sc.v, v, sizeof v);   // the argument order of search16_init (core/search.cpp:147-165)
  sc.n_mismatch = p.opt_n_mismatch ? 1 : 0;
  return sc;
}

Mirror & mirror_of(Parameters const & p, Dbindex const & dbindex, Database const & db)
{
  Mirror & m = g_mirror;
  uint64_t const n = db.getsequencecount();
  char const * const first = n > 0 ? db.getsequence(0) : nullptr;
  vsg_scoring const sc = scoring_of(p);
  int const mask_lower = (p.opt_dbmask != Masking::none) ? 1 : 0;
  if (m.group != nullptr && m.db == &db && m.count == n && m.first == first && m.nucleotides == db.getnucleotidecount() &&
      m.wordlength == dbindex.wordlength && m.mask_lower == mask_lower && std::memcmp(&m.scoring, &sc, sizeof sc) == 0) {
    return m;
  }
  m.drop();
  std::vector<int64_t> off(n);
  std::vector<int32_t> len(n);
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; i++) { off[i] = static_cast<int64_t>(total); len[i] = static_cast<int32_t>(db.getsequencelen(i)); total += db.getsequencelen(i); }
  std::vector<char> cat(total + 1);
  m.sizes.resize(n); m.labels.resize(n);
  for (uint64_t i = 0; i < n; i++) {
    std::memcpy(cat.data() + off[i], db.getsequence(i), static_cast<size_t>(len[i]));   // case = the database's soft mask
    m.sizes[i] = static_cast<int64_t>(db.getabundance(i));
    auto const it = m.label_of.emplace(std::string(db.getheader(i)), static_cast<int64_t>(m.label_of.size()));
    m.labels[i] = it.first->second;
  }
  // one upload, device-to-device copies to the other GPUs, an index per GPU (the database arrives already masked)
  std::vector<int> const devs = device_list();
  if (vsg_group_create(devs.data(), static_cast<int>(devs.size()), &sc, cat.data(), off.data(), len.data(), static_cast<int64_t>(n),
                       static_cast<int>(dbindex.wordlength), mask_lower, 0, &m.group) != VSG_OK) { die("vsg_group_create"); }
  m.db = &db; m.count = n; m.first = first; m.nucleotides = db.getnucleotidecount();
  m.wordlength = dbindex.wordlength; m.mask_lower = mask_lower; m.scoring = sc;
  return m;
}

// alignment_uses_forbidden_gap (core/searchcore.cpp:612-660): '*' penalties forbid a gap class (open) or
// runs longer than one (extension); I = gap in the query, D = gap in the target; the first CIGAR run is a
// left-terminal gap, the last a right-terminal one
bool uses_forbidden_gap(char const * cigar, Parameters const & p)
{
  bool first = true;
  char const * c = cigar;
  while (*c != '\0') {
    int64_t run = 0;
    bool digits = false;
    while (*c >= '0' && *c <= '9') { run = run * 10 + (*c - '0'); ++c; digits = true; }
    if (!digits) { run = 1; }
    char const op = *c++;
    bool const last = (*c == '\0');
    if (op == 'I' || op == 'D') {
      bool const q = (op == 'I');
      bool const open_inf = q ? (first ? p.opt_gap_open_query_left_infinite : last ? p.opt_gap_open_query_right_infinite : p.opt_gap_open_query_interior_infinite)
                              : (first ? p.opt_gap_open_target_left_infinite : last ? p.opt_gap_open_target_right_infinite : p.opt_gap_open_target_interior_infinite);
      bool const ext_inf = q ? (first ? p.opt_gap_extension_query_left_infinite : last ? p.opt_gap_extension_query_right_infinite : p.opt_gap_extension_query_interior_infinite)
                             : (first ? p.opt_gap_extension_target_left_infinite : last ? p.opt_gap_extension_target_right_infinite : p.opt_gap_extension_target_interior_infinite);
      if (open_inf || (ext_inf && run > 1)) { return true; }
    }
    first = false;
  }
  return false;
}

struct FallbackEnv {
  Parameters const * p;
  Database const * db;
  std::vector<std::string> const * queries;   // masked plus-strand queries as searched
};

// the reference's own answer for a pair its 16-bit aligner defers (core/searchcore.cpp:806-832)
int lma_fallback(void * user, int64_t query, int32_t strand, int64_t target, int64_t * out)
{
  FallbackEnv const & env = *static_cast<FallbackEnv *>(user);
  std::string const & plus = (*env.queries)[static_cast<size_t>(query)];
  std::string q = plus;
  if (strand != 0) { reverse_complement(&q[0], plus.c_str(), static_cast<int64_t>(plus.size())); }
  char const * const dseq = env.db->getsequence(static_cast<uint64_t>(target));
  auto const dlen = static_cast<int64_t>(env.db->getsequencelen(static_cast<uint64_t>(target)));
  struct Scoring scoring = scoring_from_options(*env.p);
  LinearMemoryAligner lma(scoring);   // one per call: the callback runs on several library threads
  char * const cigar = xstrdup(lma.align(q.c_str(), dseq, static_cast<int64_t>(q.size()), dlen));
  int64_t sc = 0, al = 0, ma = 0, mi = 0, ga = 0;
  lma.alignstats(cigar, q.c_str(), dseq, &sc, &al, &ma, &mi, &ga);
  out[0] = sc; out[1] = al; out[2] = ma; out[3] = mi; out[4] = ga;
  // terminal runs as align_trim reads them off the CIGAR (core/searchcore.cpp:357-417)
  auto run_at = [&](char const * s, int64_t & len, char & op) {
    len = 0;
    while (*s >= '0' && *s <= '9') { len = len * 10 + (*s - '0'); ++s; }
    if (len == 0) { len = 1; }
    op = *s;
  };
  int64_t l0 = 0, l1 = 0; char o0 = 0, o1 = 0;
  size_t const n = std::strlen(cigar);
  out[5] = out[6] = out[7] = out[8] = 0;
  if (n > 0) {
    run_at(cigar, l0, o0);
    size_t st = n - 1;
    while (st > 0 && cigar[st - 1] >= '0' && cigar[st - 1] <= '9') { st--; }
    run_at(cigar + st, l1, o1);
    if (o0 == 'D') { out[5] = l0; } else if (o0 == 'I') { out[6] = l0; }
    if (o1 == 'D') { out[7] = l1; } else if (o1 == 'I') { out[8] = l1; }
  }
  out[9] = (env.p->opt_gap_penalty_has_infinite && uses_forbidden_gap(cigar, *env.p)) ? 1 : 0;
  xfree(cigar);
  return 0;
}

}  // namespace

auto search_batch(struct Parameters const & parameters,
                  struct Dbindex const & dbindex,
                  struct Database const & db,
                  const char ** query_seqs,
                  const char ** query_heads,
                  const int * query_lens,
                  const int64_t * query_sizes,
                  int query_count,
                  struct search_result_s * results,
                  int max_results_per_query,
                  int * result_counts) -> void
{
  std::lock_guard<std::mutex> const lock(g_mutex);   // the reference's search_batch is not re-entrant either
  for (int i = 0; i < query_count; i++) { result_counts[i] = 0; }
  if (query_count <= 0 || db.getsequencecount() == 0) { return; }
  Parameters const & p = parameters;
  // the library path does not clamp the limits to the database size (search.cpp:523-531): with a limit of
  // zero search_onequery's loop never runs (searchcore.cpp:915-918)
  if (p.opt_maxaccepts <= 0 || p.opt_maxrejects <= 0) { return; }
  if (p.opt_qmask == Masking::dust && p.opt_hardmask) { fatal("GPU search_batch: --qmask dust with --hardmask is not offered on this path"); }

  Mirror & m = mirror_of(p, dbindex, db);

  // queries: one packed upload; masking as search_batch_worker_fn does it (search.cpp:437-449)
  size_t const nq = static_cast<size_t>(query_count);
  std::vector<int64_t> off(nq);
  std::vector<int32_t> len(nq);
  int64_t total = 0;
  for (size_t i = 0; i < nq; i++) { off[i] = total; len[i] = query_lens[i]; total += query_lens[i]; }
  std::vector<char> cat(static_cast<size_t>(total) + 1);
  for (size_t i = 0; i < nq; i++) { std::memcpy(cat.data() + off[i], query_seqs[i], static_cast<size_t>(len[i])); }
  bool const soft_hard = (p.opt_qmask == Masking::soft) && p.opt_hardmask;
  if (soft_hard) {
    for (size_t i = 0; i < nq; i++) {
      std::vector<char> tmp(cat.begin() + off[i], cat.begin() + off[i] + len[i]);
      tmp.push_back('\0');
      hardmask(tmp.data(), len[i]);
      std::memcpy(cat.data() + off[i], tmp.data(), static_cast<size_t>(len[i]));
    }
  }
  vsg_search_opts o;
  vsg_search_opts_default(&o);
  o.id = p.opt_id; o.weak_id = p.opt_weak_id;
  o.maxaccepts = static_cast<int32_t>(p.opt_maxaccepts); o.maxrejects = static_cast<int32_t>(p.opt_maxrejects);
  o.wordlength = static_cast<int32_t>(dbindex.wordlength);
  o.minwordmatches = static_cast<int32_t>(p.opt_minwordmatches);
  o.iddef = static_cast<int32_t>(p.opt_iddef);
  o.strand_both = p.opt_strand ? 1 : 0;
  o.mask_lower = (p.opt_qmask != Masking::none) ? 1 : 0;
  o.qmask_dust = (p.opt_qmask == Masking::dust) ? 1 : 0;
  o.minqt = p.opt_minqt; o.maxqt = p.opt_maxqt; o.minsl = p.opt_minsl; o.maxsl = p.opt_maxsl;
  o.maxid = p.opt_maxid; o.mid = p.opt_mid; o.query_cov = p.opt_query_cov; o.target_cov = p.opt_target_cov;
  o.maxsubs = p.opt_maxsubs; o.maxgaps = p.opt_maxgaps; o.mincols = p.opt_mincols; o.maxdiffs = p.opt_maxdiffs;
  o.leftjust = p.opt_leftjust != 0 ? 1 : 0; o.rightjust = p.opt_rightjust != 0 ? 1 : 0;
  o.unoise = (p.opt_cluster_unoise != nullptr) ? 1 : 0; o.unoise_alpha = p.opt_unoise_alpha;
  o.maxqsize = p.opt_maxqsize; o.mintsize = p.opt_mintsize;
  o.minsizeratio = p.opt_minsizeratio; o.maxsizeratio = p.opt_maxsizeratio;
  if (p.opt_idprefix > 2147483647 || p.opt_idsuffix > 2147483647) { fatal("GPU search_batch: --idprefix/--idsuffix out of range"); }
  o.idprefix = static_cast<int32_t>(p.opt_idprefix); o.idsuffix = static_cast<int32_t>(p.opt_idsuffix);
  o.self = p.opt_self != 0 ? 1 : 0; o.selfid = p.opt_selfid != 0 ? 1 : 0;
  o.query_sizes = query_sizes;
  o.target_sizes = m.sizes.data();
  std::vector<int64_t> qlabels;
  if (o.self != 0) {
    qlabels.resize(nq);
    for (size_t i = 0; i < nq; i++) {
      auto const it = m.label_of.find(std::string(query_heads[i]));
      qlabels[i] = it != m.label_of.end() ? it->second : -1 - static_cast<int64_t>(i);   // a header no target carries
    }
    o.query_labels = qlabels.data();
    o.target_labels = m.labels.data();
  }

  // deferred pairs: the query text the reference would hand to its LinearMemoryAligner (case is immaterial there)
  std::vector<std::string> masked(nq);
  for (size_t i = 0; i < nq; i++) { masked[i].assign(cat.data() + off[i], static_cast<size_t>(len[i])); }
  FallbackEnv env{&p, &db, &masked};
  vsg_group_set_fallback(m.group, lma_fallback, &env);

  std::vector<vsg_search_result> r(nq * static_cast<size_t>(max_results_per_query));
  // queries: one packed host buffer; every device uploads its contiguous share, DUST-masks it on the device
  // (search.cpp:437-449) and searches it
  int const rc = vsg_group_search(m.group, cat.data(), off.data(), len.data(), query_count, p.opt_qmask == Masking::dust ? 1 : 0,
                                  &o, r.data(), max_results_per_query, result_counts, nullptr);
  vsg_group_set_fallback(m.group, nullptr, nullptr);
  if (rc != VSG_OK) { die("vsg_group_search"); }
  for (size_t i = 0; i < nq; i++) {
    for (int j = 0; j < result_counts[i]; j++) {
      vsg_search_result const & x = r[i * static_cast<size_t>(max_results_per_query) + static_cast<size_t>(j)];
      search_result_s & y = results[i * static_cast<size_t>(max_results_per_query) + static_cast<size_t>(j)];
      y.target = x.target; y.id = x.id; y.matches = x.matches; y.mismatches = x.mismatches; y.gaps = x.gaps;
      y.alignment_length = x.alignment_length; y.query_length = x.query_length; y.target_length = x.target_length;
      y.accepted = x.accepted != 0; y.strand = x.strand;
    }
  }
}
