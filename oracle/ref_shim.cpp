/* oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * A C-ABI window (for ctypes) onto the UNMODIFIED reference, linked against the
 * reference objects that oracle/Makefile compiles from /root/reference/src into
 * oracle/_ref/libvsearch_ref.a.  Nothing here restates an algorithm: every entry
 * point just marshals plain buffers into the reference's own C++ interface
 *   search16_init/qprep/search16       (src/core/align_simd.hpp:76-108)
 *   unique_count                        (src/core/unique.hpp)
 *   Dbindex::prepare/add_all_sequences  (src/core/dbindex.hpp)
 *   search_topscores                    (src/core/searchcore.hpp:178)
 *   search_session_single               (src/core/search.hpp:107)
 * so that tests can pin oracle/*.c and the CUDA path against the real thing.
 * The result is oracle/_ref/libvsref.so; it is never linked into the product.
 */
#include "vsearch_api.h"
#include "core/align_simd.hpp"
#include "core/searchcore.hpp"
#include "core/search_internal.hpp"
#include "core/minheap.hpp"
#include "core/unique.hpp"
#include "core/mask.hpp"
#include "core/linmemalign.hpp"
#include "utils/string_alloc.hpp"

#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <memory>
#include <thread>

/* Link-time instrumentation of the UNMODIFIED reference: oracle/Makefile links libvsref.so with
   -Wl,--wrap=<search16>, so every call the reference makes to search16 lands here first.  We only
   count the work (pairs and DP cells = qlen*dlen handed to the SIMD aligner) and
   pass the call through to the real function. */
#include <atomic>
static std::atomic<long long> g_s16_pairs{0}, g_s16_cells{0}, g_s16_calls{0};
static thread_local int t_qlen = 0;
extern "C" {
void __real__Z14search16_qprepP9s16info_sPci(s16info_s *, char *, int);
void __wrap__Z14search16_qprepP9s16info_sPci(s16info_s * s, char * q, int qlen)
{
  t_qlen = qlen;
  __real__Z14search16_qprepP9s16info_sPci(s, q, qlen);
}
void __real__Z8search16P9s16info_sjPKjPsPtS4_S4_S4_PPcRK8Database(
    s16info_s *, unsigned int, unsigned int const *, CELL *, unsigned short *, unsigned short *,
    unsigned short *, unsigned short *, char **, Database const &);
void __wrap__Z8search16P9s16info_sjPKjPsPtS4_S4_S4_PPcRK8Database(
    s16info_s * s, unsigned int n, unsigned int const * seqnos, CELL * sc, unsigned short * a,
    unsigned short * m, unsigned short * mm, unsigned short * g, char ** cig, Database const & db)
{
  long long cells = 0;
  for (unsigned int i = 0; i < n; i++) { cells += static_cast<long long>(t_qlen) * static_cast<long long>(db.getsequencelen(seqnos[i])); }
  g_s16_pairs += n; g_s16_cells += cells; g_s16_calls += 1;
  __real__Z8search16P9s16info_sjPKjPsPtS4_S4_S4_PPcRK8Database(s, n, seqnos, sc, a, m, mm, g, cig, db);
}
void vsref_work_reset(void) { g_s16_pairs = 0; g_s16_cells = 0; g_s16_calls = 0; }
void vsref_work_get(long long * pairs, long long * cells, long long * calls)
{ *pairs = g_s16_pairs; *cells = g_s16_cells; *calls = g_s16_calls; }
}

namespace {

struct RefDb {
  Parameters params;
  Database db;
  Dbindex dbindex;
  bool session = false;
  int tophits = 0;
};

void fill_db(Database & db, int n, const char * cat, const int64_t * off, const int * len)
{
  db.init();
  for (int i = 0; i < n; i++) {
    std::string head = "t" + std::to_string(i);
    std::string seq(cat + off[i], static_cast<size_t>(len[i]));
    db.add(false, head.c_str(), seq.c_str(), nullptr, head.size(), seq.size(), 1);
  }
}

}  // namespace

extern "C" {

/* search16 on one query against n targets.  pen[14] is the search16_init argument
   list in its own order (match, mismatch, 6 opens q_l,t_l,q_i,t_i,q_r,t_r, 6 extensions
   in the same order) — i.e. ALREADY "open excludes the first extension" as
   vsearch_apply_defaults_fixups leaves them.  cigars are written NUL-terminated at
   cigar + i*cigar_stride. */
int vsref_search16(const int64_t * pen, int n_mismatch,
                   const char * qseq, int qlen,
                   int n, const char * tcat, const int64_t * toff, const int * tlen,
                   short * scores, unsigned short * aligned, unsigned short * matches,
                   unsigned short * mismatches, unsigned short * gaps,
                   char * cigar, int64_t cigar_stride)
{
  Database db;
  fill_db(db, n, tcat, toff, tlen);
  s16info_s * s = search16_init(pen[0], pen[1], pen[2], pen[3], pen[4], pen[5], pen[6], pen[7],
                                pen[8], pen[9], pen[10], pen[11], pen[12], pen[13],
                                n_mismatch != 0);
  std::vector<char> q(qseq, qseq + qlen);
  q.push_back(0);
  search16_qprep(s, q.data(), qlen);
  std::vector<unsigned int> seqnos(static_cast<size_t>(n));
  for (int i = 0; i < n; i++) { seqnos[static_cast<size_t>(i)] = static_cast<unsigned int>(i); }
  std::vector<char *> pc(static_cast<size_t>(n), nullptr);
  search16(s, static_cast<unsigned int>(n), seqnos.data(), scores, aligned, matches,
           mismatches, gaps, pc.data(), db);
  int rc = 0;
  for (int i = 0; i < n; i++) {
    char * c = pc[static_cast<size_t>(i)];
    size_t const l = std::strlen(c);
    if (static_cast<int64_t>(l) + 1 > cigar_stride) { rc = -1; }
    else { std::memcpy(cigar + i * cigar_stride, c, l + 1); }
    xfree(c);
  }
  search16_exit(s);
  db.clear();
  return rc;
}

/* distinct k-mers of one sequence in first-occurrence order; mask_lower != 0 selects the
   soft-masking map (Masking::dust / soft), else only non-ACGTU windows are skipped. */
int vsref_unique_count(int wordlength, const char * seq, int len, int mask_lower,
                       unsigned int * out, int cap)
{
  uhandle_s * uh = unique_init();
  unsigned int n = 0;
  unsigned int const * list = nullptr;
  std::vector<char> s(seq, seq + len);
  s.push_back(0);
  unique_count(uh, wordlength, len, s.data(), &n, &list,
               mask_lower != 0 ? Masking::dust : Masking::none);
  int const m = static_cast<int>(n) < cap ? static_cast<int>(n) : cap;
  for (int i = 0; i < m; i++) { out[i] = list[i]; }
  unique_exit(uh);
  return static_cast<int>(n);
}

/* DUST soft-masking of one sequence in place (core/mask.cpp:79-188 via dust()); used to
   generate pre-masked golden fixtures: the product takes soft-masked input as given. */
void vsref_dust(char * seq, int len)
{
  Parameters p;
  p.opt_hardmask = false;
  dust(seq, len, p);
}

/* Build a reference Database + Dbindex from plain buffers and open a library session.
   Only one may exist at a time (the reference serialises sessions, vsearch.cc:283). */
void * vsref_db_create(int n, const char * cat, const int64_t * off, const int * len,
                       int wordlength, double id, int maxaccepts, int maxrejects,
                       int minwordmatches /* <0: default table */, int dust /* 0: qmask/dbmask none */,
                       int strand_both, int iddef)
{
  RefDb * r = new RefDb();
  Parameters & p = r->params;
  p.opt_wordlength = wordlength;
  p.opt_id = id;
  p.opt_maxaccepts = maxaccepts;
  p.opt_maxrejects = maxrejects;
  p.opt_minwordmatches = minwordmatches;
  p.opt_threads = 1;
  p.opt_strand = (strand_both != 0);
  p.opt_iddef = iddef;
  if (dust == 0) { p.opt_qmask = Masking::none; p.opt_dbmask = Masking::none; }
  vsearch_session_begin(p);
  r->session = true;
  fill_db(r->db, n, cat, off, len);
  if (p.opt_dbmask == Masking::dust) { dust_all(r->db, p); }
  r->dbindex.prepare(1, p.opt_dbmask, r->db, p);
  r->dbindex.add_all_sequences(p.opt_dbmask, r->db, p);
  int const seqcount = static_cast<int>(r->db.getsequencecount());
  /* same clamp as search_prep / search_session_init (usearch_global.cpp:598-614) */
  int64_t th = p.opt_maxaccepts + p.opt_maxrejects + static_cast<int64_t>(MAXDELAYED);
  if (th > seqcount) { th = seqcount; }
  r->tophits = static_cast<int>(th);
  return r;
}

/* optional accept/reject filters of the session (Parameters fields of the same names); order:
   minqt maxqt minsl maxsl maxid mid query_cov target_cov maxsubs maxgaps mincols maxdiffs leftjust rightjust */
void vsref_db_set_filters(void * h, const double * v)
{
  Parameters & p = static_cast<RefDb *>(h)->params;
  p.opt_minqt = v[0]; p.opt_maxqt = v[1]; p.opt_minsl = v[2]; p.opt_maxsl = v[3];
  p.opt_maxid = v[4]; p.opt_mid = v[5]; p.opt_query_cov = v[6]; p.opt_target_cov = v[7];
  p.opt_maxsubs = static_cast<int64_t>(v[8]); p.opt_maxgaps = static_cast<int64_t>(v[9]);
  p.opt_mincols = static_cast<int64_t>(v[10]); p.opt_maxdiffs = static_cast<int64_t>(v[11]);
  p.opt_leftjust = static_cast<int64_t>(v[12]); p.opt_rightjust = static_cast<int64_t>(v[13]);
}

void vsref_db_free(void * h)
{
  RefDb * r = static_cast<RefDb *>(h);
  r->dbindex.clear();
  r->db.clear();
  if (r->session) { vsearch_session_end(); }
  delete r;
}

int vsref_db_tophits(void * h) { return static_cast<RefDb *>(h)->tophits; }

/* search_topscores for one (already masked, plus-strand) query: best-first list of
   (seqno,count,length) exactly as minheap_poplast would hand them out. */
int vsref_db_topscores(void * h, const char * qseq, int qlen,
                       unsigned int * seqno, unsigned int * count, unsigned int * length)
{
  RefDb * r = static_cast<RefDb *>(h);
  searchinfo_s si;
  int const seqcount = static_cast<int>(r->db.getsequencecount());
  search_thread_init(&si, seqcount, r->tophits, r->params, r->dbindex, r->db);
  std::vector<char> q(qseq, qseq + qlen);
  q.push_back(0);
  si.qsequence = q.data();
  si.qseqlen = qlen;
  unique_count(si.uh, static_cast<int>(r->dbindex.wordlength), qlen, q.data(),
               &si.kmersamplecount, &si.kmersample, r->params.opt_qmask);
  search_topscores(&si);
  int n = 0;
  while (not minheap_isempty(si.m)) {
    elem_t const e = minheap_poplast(si.m);
    seqno[n] = e.seqno; count[n] = e.count; length[n] = e.length;
    ++n;
  }
  si.qsequence = nullptr;  /* borrowed */
  search_thread_exit(&si);
  return n;
}

/* search_session_single over a batch of queries (the sequential form of search_batch,
   search.cpp:511); results[q*max_results + j], fields flattened to plain arrays. */
void vsref_db_search(void * h, int nq, const char * qcat, const int64_t * qoff, const int * qlen,
                     int max_results, int * counts,
                     int * target, double * id, int * matches, int * mismatches, int * gaps,
                     int * alnlen, int * accepted, int * strand)
{
  RefDb * r = static_cast<RefDb *>(h);
  search_session_s * ss = search_session_alloc();
  search_session_init(ss, r->params, r->dbindex, r->db);
  std::vector<search_result_s> res(static_cast<size_t>(max_results));
  for (int q = 0; q < nq; q++) {
    std::string seq(qcat + qoff[q], static_cast<size_t>(qlen[q]));
    std::string head = "q" + std::to_string(q);
    int c = 0;
    search_session_single(ss, seq.c_str(), head.c_str(), qlen[q], 1, res.data(), max_results, &c);
    counts[q] = c;
    for (int j = 0; j < c; j++) {
      size_t const o = static_cast<size_t>(q) * static_cast<size_t>(max_results) + static_cast<size_t>(j);
      search_result_s const & x = res[static_cast<size_t>(j)];
      target[o] = x.target; id[o] = x.id; matches[o] = x.matches; mismatches[o] = x.mismatches;
      gaps[o] = x.gaps; alnlen[o] = x.alignment_length; accepted[o] = x.accepted ? 1 : 0;
      strand[o] = x.strand;
    }
  }
  search_session_cleanup(ss);
  search_session_free(ss);
}

/* the reference's own multi-threaded search_batch (core/search.cpp:511-593) on `threads` host
   threads; returns the number of queries with at least one hit.  Results are discarded except
   for the first hit's target (first_target[q], -1 if none): this entry point exists to TIME the
   reference and to count its search16 workload. */
int vsref_db_search_batch(void * h, int nq, const char * qcat, const int64_t * qoff, const int * qlen,
                          int threads, int * first_target)
{
  RefDb * r = static_cast<RefDb *>(h);
  std::vector<std::string> seqs(static_cast<size_t>(nq)), heads(static_cast<size_t>(nq));
  std::vector<const char *> ps(static_cast<size_t>(nq)), ph(static_cast<size_t>(nq));
  std::vector<int64_t> sizes(static_cast<size_t>(nq), 1);
  for (int q = 0; q < nq; q++) {
    seqs[static_cast<size_t>(q)].assign(qcat + qoff[q], static_cast<size_t>(qlen[q]));
    heads[static_cast<size_t>(q)] = "q" + std::to_string(q);
    ps[static_cast<size_t>(q)] = seqs[static_cast<size_t>(q)].c_str();
    ph[static_cast<size_t>(q)] = heads[static_cast<size_t>(q)].c_str();
  }
  Parameters p = r->params;
  p.opt_threads = threads;
  std::vector<search_result_s> res(static_cast<size_t>(nq));
  std::vector<int> counts(static_cast<size_t>(nq), 0);
  search_batch(p, r->dbindex, r->db, ps.data(), ph.data(), qlen, sizes.data(), nq, res.data(), 1, counts.data());
  int hits = 0;
  for (int q = 0; q < nq; q++) {
    first_target[q] = counts[static_cast<size_t>(q)] > 0 ? res[static_cast<size_t>(q)].target : -1;
    hits += counts[static_cast<size_t>(q)] > 0;
  }
  return hits;
}

/* the same multi-threaded search_batch with every result record kept (results[q*max_results + j], fields
   flattened to plain arrays as in vsref_db_search): the at-scale parity tests compare full rows. */
void vsref_db_search_batch_rows(void * h, int nq, const char * qcat, const int64_t * qoff, const int * qlen,
                                int threads, int max_results, int * counts,
                                int * target, double * id, int * matches, int * mismatches, int * gaps,
                                int * alnlen, int * accepted, int * strand)
{
  RefDb * r = static_cast<RefDb *>(h);
  std::vector<std::string> seqs(static_cast<size_t>(nq)), heads(static_cast<size_t>(nq));
  std::vector<const char *> ps(static_cast<size_t>(nq)), ph(static_cast<size_t>(nq));
  std::vector<int64_t> sizes(static_cast<size_t>(nq), 1);
  for (int q = 0; q < nq; q++) {
    seqs[static_cast<size_t>(q)].assign(qcat + qoff[q], static_cast<size_t>(qlen[q]));
    heads[static_cast<size_t>(q)] = "q" + std::to_string(q);
    ps[static_cast<size_t>(q)] = seqs[static_cast<size_t>(q)].c_str();
    ph[static_cast<size_t>(q)] = heads[static_cast<size_t>(q)].c_str();
  }
  Parameters p = r->params;
  p.opt_threads = threads;
  std::vector<search_result_s> res(static_cast<size_t>(nq) * static_cast<size_t>(max_results));
  search_batch(p, r->dbindex, r->db, ps.data(), ph.data(), qlen, sizes.data(), nq, res.data(), max_results, counts);
  for (int q = 0; q < nq; q++) {
    for (int j = 0; j < counts[q]; j++) {
      size_t const o = static_cast<size_t>(q) * static_cast<size_t>(max_results) + static_cast<size_t>(j);
      search_result_s const & x = res[o];
      target[o] = x.target; id[o] = x.id; matches[o] = x.matches; mismatches[o] = x.mismatches;
      gaps[o] = x.gaps; alnlen[o] = x.alignment_length; accepted[o] = x.accepted ? 1 : 0; strand[o] = x.strand;
    }
  }
}

/* LinearMemoryAligner::align + alignstats (core/linmemalign.cpp) with the session's scoring: the
   reference's answer for pairs its 16-bit aligner defers.  out = {score, alnlen, matches,
   mismatches, gaps}; the CIGAR goes to cigar (cap bytes). */
int vsref_lma(void * h, const char * q, int qlen, const char * t, int tlen, long long * out, char * cigar, int cap)
{
  RefDb * r = static_cast<RefDb *>(h);
  struct Scoring scoring = scoring_from_options(r->params);
  LinearMemoryAligner lma(scoring);
  std::string qs(q, static_cast<size_t>(qlen)), ts(t, static_cast<size_t>(tlen));
  char * c = xstrdup(lma.align(qs.c_str(), ts.c_str(), qlen, tlen));
  int64_t sc = 0, al = 0, ma = 0, mi = 0, ga = 0;
  lma.alignstats(c, qs.c_str(), ts.c_str(), &sc, &al, &ma, &mi, &ga);
  out[0] = sc; out[1] = al; out[2] = ma; out[3] = mi; out[4] = ga;
  int rc = 0;
  if (static_cast<int>(std::strlen(c)) + 1 > cap) { rc = -1; } else { std::strcpy(cigar, c); }
  xfree(c);
  return rc;
}

/* the per-query body of allpairs_thread_run (commands/allpairs_global.cpp:395-431) for rows
   [row0, row0+nrows) of the session's database on `threads` host threads: search16_qprep + ONE
   search16 call against all later sequences.  Exists to TIME the reference's aligner on the dense
   all-pairs workload (results are discarded; the --wrap counter records pairs and cells). */
long long vsref_allpairs_rows(void * h, int row0, int nrows, int threads)
{
  RefDb * r = static_cast<RefDb *>(h);
  Database & db = r->db;
  Parameters const & p = r->params;
  int const n = static_cast<int>(db.getsequencecount());
  std::atomic<int> next{row0};
  std::atomic<long long> pairs{0};
  auto worker = [&]() {
    s16info_s * s = search16_init(p.opt_match, p.opt_mismatch, p.opt_gap_open_query_left, p.opt_gap_open_target_left,
                                  p.opt_gap_open_query_interior, p.opt_gap_open_target_interior,
                                  p.opt_gap_open_query_right, p.opt_gap_open_target_right,
                                  p.opt_gap_extension_query_left, p.opt_gap_extension_target_left,
                                  p.opt_gap_extension_query_interior, p.opt_gap_extension_target_interior,
                                  p.opt_gap_extension_query_right, p.opt_gap_extension_target_right, p.opt_n_mismatch);
    std::vector<unsigned int> seqnos;
    std::vector<CELL> sc; std::vector<unsigned short> a, m, mm, g; std::vector<char *> cg;
    for (;;) {
      int const i = next.fetch_add(1);
      if (i >= row0 + nrows || i >= n) { break; }
      int const cnt = n - i - 1;
      if (cnt <= 0) { continue; }
      seqnos.resize(cnt); sc.resize(cnt); a.resize(cnt); m.resize(cnt); mm.resize(cnt); g.resize(cnt); cg.assign(cnt, nullptr);
      for (int j = 0; j < cnt; j++) { seqnos[j] = static_cast<unsigned int>(i + 1 + j); }
      search16_qprep(s, db.mutatesequence(i), static_cast<int>(db.getsequencelen(i)));
      search16(s, static_cast<unsigned int>(cnt), seqnos.data(), sc.data(), a.data(), m.data(), mm.data(), g.data(), cg.data(), db);
      for (int j = 0; j < cnt; j++) { if (cg[j] != nullptr) { xfree(cg[j]); } }
      pairs += cnt;
    }
    search16_exit(s);
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) { pool.emplace_back(worker); }
  for (auto & th : pool) { th.join(); }
  return pairs.load();
}

}  // extern "C"
