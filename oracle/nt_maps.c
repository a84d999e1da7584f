/* oracle/nt_maps.c — TEST INFRASTRUCTURE ONLY.
 * Nucleotide symbol maps of the vsearch hot path, derived from the IUPAC definitions
 * rather than tabulated (reference tables: src/utils/maps.cpp:75-118 chrmap_4bit,
 * 121-151 chrmap_complement, 153-185 chrmap_2bit, 188-205 chrmap_ambiguous_4bit,
 * 208-266 chrmap_mask_ambig / chrmap_mask_lower).  Pinned against the reference in
 * tests/test_oracle_vs_ref.py (all 256 byte values through search16 / unique_count).
 */
#include "oracle.h"

static unsigned char upper(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }

/* 4-bit code = set of bases {A=1,C=2,G=4,T=8} the IUPAC letter stands for; case-insensitive;
   every non-IUPAC byte is 0 (maps.cpp:75-118). */
unsigned char oracle_map_4bit(unsigned char c)
{
  switch (upper(c)) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;
    case 'T': case 'U': return 8;
    case 'M': return 1 | 2;       /* A or C */
    case 'R': return 1 | 4;       /* A or G */
    case 'S': return 2 | 4;       /* C or G */
    case 'V': return 1 | 2 | 4;   /* not T  */
    case 'W': return 1 | 8;       /* A or T */
    case 'Y': return 2 | 8;       /* C or T */
    case 'H': return 1 | 2 | 8;   /* not G  */
    case 'K': return 4 | 8;       /* G or T */
    case 'D': return 1 | 4 | 8;   /* not C  */
    case 'B': return 2 | 4 | 8;   /* not A  */
    case 'N': return 15;
    default:  return 0;
  }
}

/* 2-bit code A0 C1 G2 T/U3, everything else 0 (maps.cpp:153-185) */
unsigned int oracle_map_2bit(unsigned char c)
{
  switch (upper(c)) {
    case 'C': return 1; case 'G': return 2; case 'T': case 'U': return 3;
    default: return 0;
  }
}

static int is_acgtu(unsigned char u) { return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U'; }

/* 1 = exclude from k-mer sampling.  ambig: anything but ACGTU in either case
   (maps.cpp:208-236); lower: additionally all lower-case letters (maps.cpp:239-266). */
unsigned int oracle_map_mask_ambig(unsigned char c) { return is_acgtu(upper(c)) ? 0U : 1U; }
unsigned int oracle_map_mask_lower(unsigned char c) { return is_acgtu(c) ? 0U : 1U; }

/* every 4-bit code except the four single bases is "ambiguous", code 0 included
   (maps.cpp:188-205) */
int oracle_is_ambiguous_4bit(unsigned int code)
{
  return !(code == 1 || code == 2 || code == 4 || code == 8);
}

/* complement keeps the case of IUPAC letters; U/u -> A/a; 'n' -> 'n'; every other byte
   (including non-IUPAC letters of either case) -> 'N' (maps.cpp:121-151) */
char oracle_complement(unsigned char c)
{
  unsigned char const u = upper(c);
  int const lower = (c != u);
  char r;
  switch (u) {
    case 'A': r = 'T'; break; case 'C': r = 'G'; break; case 'G': r = 'C'; break;
    case 'T': case 'U': r = 'A'; break;
    case 'M': r = 'K'; break; case 'K': r = 'M'; break;
    case 'R': r = 'Y'; break; case 'Y': r = 'R'; break;
    case 'S': r = 'S'; break; case 'W': r = 'W'; break;
    case 'V': r = 'B'; break; case 'B': r = 'V'; break;
    case 'H': r = 'D'; break; case 'D': r = 'H'; break;
    case 'N': r = 'N'; break;
    default: return 'N';
  }
  return lower ? (char)(r + 32) : r;
}
