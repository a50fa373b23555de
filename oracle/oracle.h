/*
 * oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of what the reference's three public entry points compute
 * (/root/reference/src/library/msufsort/msufsort.h:57-75).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product library
 * (msufsort_b200/lib/libb200sa.so) never does and has no CPU path of its own.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (i)  the unmodified reference compiled from /root/reference into oracle/_ref/libmsufsort_ref.so
 *        (recipe: oracle/Makefile) on the same bytes, and
 *   (ii) the known-answer FNV-1a-64 digests in tests/golden/kat.json, which were produced by that
 *        reference build with tests/golden/make_golden.py.
 */
#ifndef B200SA_ORACLE_H
#define B200SA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SA with the reference's conventions (msufsort.cpp:1730-1767, :1720): n+1 entries, sa[0]=n.
 * Returns 0 on success, -1 on allocation failure / bad args.  n <= 2^31-2. */
int oracle_make_suffix_array(const uint8_t* text, int64_t n, int32_t* sa_out);

/* Brute-force SA for tiny n: qsort with the ordering of the reference's validator
 * (main.cpp:210-232 compare(): first differing byte decides, a proper prefix is smaller). */
int oracle_make_suffix_array_bruteforce(const uint8_t* text, int64_t n, int32_t* sa_out);

/* Forward BWT in place, returns the sentinel row (msufsort.cpp:1771-1817; copy-out :1811-1815).
 * Returns -1 on failure. */
int32_t oracle_forward_bwt(uint8_t* text_inout, int64_t n);

/* BWT from a finished SA into a separate buffer (same definition). */
int32_t oracle_bwt_from_sa(const uint8_t* text, int64_t n, const int32_t* sa, uint8_t* bwt_out);

/* Inverse BWT in place (msufsort.cpp:1821-2096: psi table build :1880-1915, walk from
 * index[0] = sentinelIndex :1922-2015).  Returns 0 / -1. */
int oracle_reverse_bwt(uint8_t* bwt_inout, int64_t n, int32_t sentinel_index);

/* O(n) SA validator used for sizes beyond the reference (SURVEY.md §8c last row). Returns the
 * number of offending rows (0 = correct), or -1 on allocation failure. */
int64_t oracle_check_suffix_array(const uint8_t* text, int64_t n, const int32_t* sa);

/* LCP array, n+1 entries aligned with the SA: lcp[0] = lcp[1] = 0, lcp[r] = lcp(SA[r-1], SA[r]).
 * oracle_make_lcp_array restates the reference demo (main.cpp:16-64; its output[i] = lcp[i+2]);
 * oracle_lcp_kasai is an independent linear-time cross-check.  Return 0 / -1. */
int oracle_make_lcp_array(const uint8_t* text, int64_t n, const int32_t* sa, int32_t* lcp_out);
int oracle_lcp_kasai(const uint8_t* text, int64_t n, const int32_t* sa, int32_t* lcp_out);

/* FNV-1a-64 (h=0xcbf29ce484222325; h^=b; h*=0x100000001b3) over raw bytes. */
uint64_t oracle_fnv1a64(const void* data, int64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
