/* examples/usearch_udb.c — `vsearch --usearch_global QUERIES --db DB.udb --id ID --blast6out OUT` on the GPU(s), in
 * plain C against include/vsg.h: what a maintainer's command-level binding of this library looks like.
 *
 * The calls are the ones tests/test_udb_gpu.py drives through ctypes and compares byte for byte with the reference
 * CLI: vsg_udb_open (= udb_detect_isudb + udb_read, core/udb.cpp:120-578), vsg_group_create_udb (upload, device index,
 * check against the stored index; database copied to every listed GPU), vsg_usearch_stream (= the query loop of
 * commands/usearch_global.cpp:376-534 with results_show_blast6out_one, core/results.cpp:221-271).
 *
 *   gcc -O2 -Iinclude examples/usearch_udb.c -Lvsearch_b200/csrc -lvsg -Wl,-rpath,$PWD/vsearch_b200/csrc -o usearch_udb
 *   ./usearch_udb db.udb queries.fasta out.b6 0.9 [gpu,gpu,...]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vsg.h"

static int fail(const char * what)
{
  fprintf(stderr, "%s: %s\n", what, vsg_last_error());
  return 1;
}

int main(int argc, char ** argv)
{
  if (argc < 5) {
    fprintf(stderr, "usage: %s DB.udb QUERIES.fasta OUT.blast6 ID [gpu[,gpu...]]\n", argv[0]);
    return 2;
  }
  int devices[16], ndev = 0;
  if (argc > 5) {
    char * list = argv[5];
    for (char * tok = strtok(list, ","); tok != NULL && ndev < 16; tok = strtok(NULL, ",")) { devices[ndev++] = atoi(tok); }
  }
  if (ndev == 0) { devices[ndev++] = 0; }

  if (vsg_udb_detect(argv[1]) != 1) { fprintf(stderr, "%s is not a UDB file\n", argv[1]); return 2; }
  vsg_udb * udb = NULL;
  if (vsg_udb_open(argv[1], &udb) != VSG_OK) { return fail("vsg_udb_open"); }
  vsg_udb_info info;
  vsg_udb_info_get(udb, &info);
  fprintf(stderr, "%lld sequences, %lld nt, word length %d\n", (long long)info.sequences, (long long)info.nucleotides, info.wordlength);

  /* the reference's default scoring after its fix-ups (vsearch.cc:250-259): match 2, mismatch -4, interior gaps
     20 + 2 per residue, terminal gaps 2 + 1 — in search16_init's argument order (core/align_simd.hpp:76-91) */
  vsg_scoring sc;
  int64_t const v[14] = {2, -4, 2, 2, 20, 20, 2, 2, 1, 1, 2, 2, 1, 1};
  memcpy(sc.v, v, sizeof v);
  sc.n_mismatch = 0;

  vsg_group * group = NULL;
  if (vsg_group_create_udb(devices, ndev, &sc, udb, &group) != VSG_OK) { return fail("vsg_group_create_udb"); }

  const char ** labels = (const char **)malloc(sizeof(char *) * (size_t)info.sequences);
  if (labels == NULL) { return 1; }
  for (int64_t i = 0; i < info.sequences; i++) { labels[i] = vsg_udb_header(udb, i); }

  vsg_search_opts o;
  vsg_search_opts_default(&o);
  o.id = atof(argv[4]);
  o.wordlength = info.wordlength;
  o.mask_lower = 1;   /* --qmask dust (the default): queries are DUST-masked on the device, masked symbols seed no words */
  o.qmask_dust = 1;

  vsg_stream_stats st;
  if (vsg_usearch_stream(group, labels, argv[2], &o, /*qmask_dust=*/1, /*notrunclabels=*/0, /*batch_queries=*/65536,
                         /*maxhits=*/0, /*output_no_hits=*/0, argv[3], &st) != VSG_OK) { return fail("vsg_usearch_stream"); }
  fprintf(stderr, "%lld queries, %lld matched, %lld rows in %.2f s (parse %.2f s, search %.2f s, write %.2f s busy)\n",
          (long long)st.queries, (long long)st.matched, (long long)st.rows, st.wall_s, st.parse_s, st.search_s, st.write_s);
  free(labels);
  vsg_group_destroy(group);
  vsg_udb_close(udb);
  return 0;
}
