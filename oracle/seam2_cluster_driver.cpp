/* oracle/seam2_cluster_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Linked twice by oracle/Makefile: _ref/seam2_cluster_driver_ref against the UNMODIFIED reference,
 * _ref/seam2_cluster_driver_gpu against the same objects with cluster_session_* / cluster_assign_* replaced by
 * shim/cluster_session_vsg.cpp (+ libvsg.so).  Drives the reference's incremental clustering API exactly as
 * api_examples/example_cluster.cc does (Database::add, dust_all, sortbylength, Dbindex::prepare,
 * cluster_session_init, cluster_assign_batch / cluster_assign_single, src/core/cluster.hpp:78-118) and prints every
 * result record, so that tests/test_seam2_gpu.py can diff the two.
 *
 *   seam2_cluster_driver reads.fasta [key=value ...]
 * Keys: id maxaccepts maxrejects wordlength threads qmask(none|dust) iddef chunk (sequences per
 * cluster_assign_batch call; 0 = all at once; -1 = cluster_assign_single one by one) minsl maxqt mid
 */
#include "vsearch_api.h"
#include "core/mask.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

int main(int argc, char ** argv)
{
  if (argc < 2) { std::fprintf(stderr, "usage: %s reads.fasta [key=value ...]\n", argv[0]); return 2; }
  Parameters p;
  p.opt_wordlength = 8;
  p.opt_id = 0.97;
  p.opt_maxaccepts = 1;
  p.opt_maxrejects = 8;
  p.opt_threads = 4;
  int chunk = 0;
  for (int a = 2; a < argc; a++) {
    char * eq = std::strchr(argv[a], '=');
    if (eq == nullptr) { std::fprintf(stderr, "bad argument %s\n", argv[a]); return 2; }
    std::string const k(argv[a], static_cast<size_t>(eq - argv[a]));
    const char * v = eq + 1;
    if (k == "id") { p.opt_id = std::atof(v); }
    else if (k == "maxaccepts") { p.opt_maxaccepts = std::atoll(v); }
    else if (k == "maxrejects") { p.opt_maxrejects = std::atoll(v); }
    else if (k == "wordlength") { p.opt_wordlength = std::atoll(v); }
    else if (k == "threads") { p.opt_threads = std::atoll(v); }
    else if (k == "qmask") { p.opt_qmask = std::strcmp(v, "none") == 0 ? Masking::none : Masking::dust; }
    else if (k == "iddef") { p.opt_iddef = std::atoll(v); }
    else if (k == "minsl") { p.opt_minsl = std::atof(v); }
    else if (k == "maxqt") { p.opt_maxqt = std::atof(v); }
    else if (k == "mid") { p.opt_mid = std::atof(v); }
    else if (k == "unoise_alpha") { p.opt_cluster_unoise = const_cast<char *>("unoise"); p.opt_unoise_alpha = std::atof(v); }
    else if (k == "sizeorder") { p.opt_sizeorder = std::atoi(v) != 0; }
    else if (k == "chunk") { chunk = std::atoi(v); }
    else { std::fprintf(stderr, "unknown key %s\n", k.c_str()); return 2; }
  }
  vsearch_session_begin(p);
  Database db;
  db.init();
  {
    std::ifstream in(argv[1]);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::string line, head, seq;
    auto flush = [&]() {
      if (head.empty()) { return; }
      int64_t size = 1;   // ";size=N" in the header = the abundance (what --sizein reads)
      size_t const at = head.find(";size=");
      if (at != std::string::npos) { size = std::max<int64_t>(1, std::atoll(head.c_str() + at + 6)); }
      db.add(false, head.c_str(), seq.c_str(), nullptr, head.size(), seq.size(), size);
    };
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') { line.pop_back(); }
      if (line.empty()) { continue; }
      if (line[0] == '>') { flush(); head = line.substr(1); seq.clear(); } else { seq += line; }
    }
    flush();
  }
  dust_all(db, p);
  db.sortbylength(p);
  Dbindex dbindex;
  dbindex.prepare(1, p.opt_qmask, db, p);
  struct cluster_session_s * cs = cluster_session_alloc();
  cluster_session_init(cs, p, dbindex, db);
  int const n = static_cast<int>(db.getsequencecount());
  std::vector<cluster_result_s> r(static_cast<size_t>(n));
  if (chunk < 0) {
    for (int i = 0; i < n; i++) { cluster_assign_single(cs, i, &r[static_cast<size_t>(i)]); }
  } else {
    int const step = chunk == 0 ? n : chunk;
    for (int s = 0; s < n; s += step) { cluster_assign_batch(cs, s, std::min(step, n - s), r.data() + s); }
  }
  for (int i = 0; i < n; i++) {
    cluster_result_s const & x = r[static_cast<size_t>(i)];
    std::printf("%d\t%s\t%d\t%d\t%d\t%s\t%.10f\t%s\t%d\n", i, db.getheader(static_cast<uint64_t>(i)), x.is_centroid ? 1 : 0, x.cluster_id,
                x.centroid_seqno, x.centroid_label, x.identity, x.cigar[0] != 0 ? x.cigar : "*", x.cigar_truncated ? 1 : 0);
  }
  cluster_session_cleanup(cs);
  cluster_session_free(cs);
  dbindex.clear();
  db.clear();
  vsearch_session_end();
  return 0;
}
