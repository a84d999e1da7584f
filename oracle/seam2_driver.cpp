/* oracle/seam2_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * One program, linked twice by oracle/Makefile:
 *   _ref/seam2_driver_ref   against the UNMODIFIED reference (libvsearch_ref.a)
 *   _ref/seam2_driver_gpu   against the same objects with search_batch() replaced by
 *                           shim/search_batch_vsg.cpp (+ libvsg.so)
 * It drives the reference's library API exactly as api_examples/example_search.cc does
 * (vsearch_session_begin, Database::add, dust_all, Dbindex::prepare/add_all_sequences, search_batch,
 * src/core/search.hpp:135-145) with options given as key=value arguments and prints every result
 * record in full precision, so that tests/test_seam2_gpu.py can diff the two.
 *
 *   seam2_driver db.fasta queries.fasta [key=value ...]
 * FASTA headers may carry ";size=N" (abundance).  Keys: id weak_id maxaccepts maxrejects wordlength
 * strand(0/1) qmask dbmask (none|dust|soft) hardmask threads max_results iddef self selfid idprefix
 * idsuffix maxqsize mintsize minsizeratio maxsizeratio minqt maxqt minsl maxsl maxid mid query_cov
 * target_cov maxsubs maxgaps mincols maxdiffs leftjust rightjust infinite(=qi|ti|ql|... gap-open classes,
 * comma separated) infinite_ext(=...) gapopen_i gapext_i match mismatch
 */
#include "vsearch_api.h"
#include "core/mask.hpp"

#include <cinttypes>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

struct Rec { std::string head, seq; int64_t size = 1; };

static std::vector<Rec> read_fasta(const char * path)
{
  std::vector<Rec> out;
  std::ifstream in(path);
  if (!in) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') { line.pop_back(); }
    if (line.empty()) { continue; }
    if (line[0] == '>') {
      Rec r;
      r.head = line.substr(1);
      size_t const p = r.head.find(";size=");
      if (p != std::string::npos) { r.size = std::atoll(r.head.c_str() + p + 6); }
      out.push_back(r);
    } else if (!out.empty()) {
      out.back().seq += line;
    }
  }
  return out;
}

static Masking mask_of(const char * v)
{
  if (std::strcmp(v, "none") == 0) { return Masking::none; }
  if (std::strcmp(v, "soft") == 0) { return Masking::soft; }
  return Masking::dust;
}

int main(int argc, char ** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: %s db.fasta queries.fasta [key=value ...]\n", argv[0]); return 2; }
  Parameters p;
  p.opt_wordlength = 8;
  p.opt_id = 0.9;
  p.opt_threads = 4;
  int max_results = 8;
  for (int a = 3; a < argc; a++) {
    char * eq = std::strchr(argv[a], '=');
    if (eq == nullptr) { std::fprintf(stderr, "bad argument %s\n", argv[a]); return 2; }
    std::string const k(argv[a], static_cast<size_t>(eq - argv[a]));
    const char * v = eq + 1;
    if (k == "id") { p.opt_id = std::atof(v); }
    else if (k == "weak_id") { p.opt_weak_id = std::atof(v); }
    else if (k == "maxaccepts") { p.opt_maxaccepts = std::atoll(v); }
    else if (k == "maxrejects") { p.opt_maxrejects = std::atoll(v); }
    else if (k == "wordlength") { p.opt_wordlength = std::atoll(v); }
    else if (k == "strand") { p.opt_strand = std::atoi(v) != 0; }
    else if (k == "qmask") { p.opt_qmask = mask_of(v); }
    else if (k == "dbmask") { p.opt_dbmask = mask_of(v); }
    else if (k == "hardmask") { p.opt_hardmask = std::atoi(v) != 0; }
    else if (k == "threads") { p.opt_threads = std::atoll(v); }
    else if (k == "max_results") { max_results = std::atoi(v); }
    else if (k == "iddef") { p.opt_iddef = std::atoll(v); }
    else if (k == "self") { p.opt_self = std::atoll(v); }
    else if (k == "selfid") { p.opt_selfid = std::atoll(v); }
    else if (k == "idprefix") { p.opt_idprefix = std::atoll(v); }
    else if (k == "idsuffix") { p.opt_idsuffix = std::atoll(v); }
    else if (k == "maxqsize") { p.opt_maxqsize = std::atoll(v); }
    else if (k == "mintsize") { p.opt_mintsize = std::atoll(v); }
    else if (k == "minsizeratio") { p.opt_minsizeratio = std::atof(v); }
    else if (k == "maxsizeratio") { p.opt_maxsizeratio = std::atof(v); }
    else if (k == "minqt") { p.opt_minqt = std::atof(v); }
    else if (k == "maxqt") { p.opt_maxqt = std::atof(v); }
    else if (k == "minsl") { p.opt_minsl = std::atof(v); }
    else if (k == "maxsl") { p.opt_maxsl = std::atof(v); }
    else if (k == "maxid") { p.opt_maxid = std::atof(v); }
    else if (k == "mid") { p.opt_mid = std::atof(v); }
    else if (k == "query_cov") { p.opt_query_cov = std::atof(v); }
    else if (k == "target_cov") { p.opt_target_cov = std::atof(v); }
    else if (k == "maxsubs") { p.opt_maxsubs = std::atoll(v); }
    else if (k == "maxgaps") { p.opt_maxgaps = std::atoll(v); }
    else if (k == "mincols") { p.opt_mincols = std::atoll(v); }
    else if (k == "maxdiffs") { p.opt_maxdiffs = std::atoll(v); }
    else if (k == "leftjust") { p.opt_leftjust = std::atoll(v); }
    else if (k == "rightjust") { p.opt_rightjust = std::atoll(v); }
    else if (k == "unoise_alpha") { p.opt_cluster_unoise = const_cast<char *>("unoise"); p.opt_unoise_alpha = std::atof(v); }
    else if (k == "match") { p.opt_match = std::atoll(v); }
    else if (k == "mismatch") { p.opt_mismatch = std::atoll(v); }
    else if (k == "gapopen_i") { p.opt_gap_open_query_interior = p.opt_gap_open_target_interior = std::atoll(v); }
    else if (k == "gapext_i") { p.opt_gap_extension_query_interior = p.opt_gap_extension_target_interior = std::atoll(v); }
    else if (k == "infinite" || k == "infinite_ext") {
      /* what cli.cc:211-229 does for '*': the penalty becomes INT_MAX and the class is flagged */
      bool const open = (k == "infinite");
      std::string s(v);
      size_t pos = 0;
      while (pos <= s.size()) {
        size_t const e = s.find(',', pos);
        std::string const c = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
        int64_t const inf = INT_MAX;
#define CLASS(tag, field) if (c == tag) { if (open) { p.opt_gap_open_##field = inf; p.opt_gap_open_##field##_infinite = true; } \
                                           else { p.opt_gap_extension_##field = inf; p.opt_gap_extension_##field##_infinite = true; } }
        CLASS("ql", query_left) CLASS("qi", query_interior) CLASS("qr", query_right)
        CLASS("tl", target_left) CLASS("ti", target_interior) CLASS("tr", target_right)
#undef CLASS
        p.opt_gap_penalty_has_infinite = true;
        if (e == std::string::npos) { break; }
        pos = e + 1;
      }
    }
    else { std::fprintf(stderr, "unknown key %s\n", k.c_str()); return 2; }
  }

  vsearch_session_begin(p);
  std::vector<Rec> const refs = read_fasta(argv[1]);
  std::vector<Rec> const qs = read_fasta(argv[2]);
  Database db;
  db.init();
  for (auto const & r : refs) {
    db.add(false, r.head.c_str(), r.seq.c_str(), nullptr, r.head.size(), r.seq.size(), r.size);
  }
  if (p.opt_dbmask == Masking::dust) { dust_all(db, p); }
  else if ((p.opt_dbmask == Masking::soft) && p.opt_hardmask) { hardmask_all(db); }
  Dbindex dbindex;
  dbindex.prepare(1, p.opt_dbmask, db, p);
  dbindex.add_all_sequences(p.opt_dbmask, db, p);

  int const nq = static_cast<int>(qs.size());
  std::vector<const char *> q_seqs(nq), q_heads(nq);
  std::vector<int> q_lens(nq);
  std::vector<int64_t> q_sizes(nq);
  for (int i = 0; i < nq; i++) {
    q_seqs[i] = qs[i].seq.c_str(); q_heads[i] = qs[i].head.c_str();
    q_lens[i] = static_cast<int>(qs[i].seq.size()); q_sizes[i] = qs[i].size;
  }
  std::vector<search_result_s> res(static_cast<size_t>(nq) * max_results);
  std::vector<int> counts(nq, 0);
  /* twice: the second call must find the device mirror of the first still valid */
  for (int rep = 0; rep < 2; rep++) {
    search_batch(p, dbindex, db, q_seqs.data(), q_heads.data(), q_lens.data(), q_sizes.data(), nq,
                 res.data(), max_results, counts.data());
  }
  for (int i = 0; i < nq; i++) {
    for (int j = 0; j < counts[i]; j++) {
      search_result_s const & r = res[static_cast<size_t>(i) * max_results + j];
      std::printf("%s\t%s\t%.17g\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", q_heads[i], db.getheader(r.target), r.id, r.matches,
                  r.mismatches, r.gaps, r.alignment_length, r.query_length, r.target_length, r.accepted ? 1 : 0, r.strand);
    }
  }
  dbindex.clear();
  db.clear();
  vsearch_session_end();
  return 0;
}
