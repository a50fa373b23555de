/*
 * oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C CPU restatement of the results of the reference's three entry points.
 *
 * The reference sorts with an "improved two-stage" scheme: multikey quicksort of B* suffixes
 * (msufsort.cpp:488-642) + tandem-repeat completion (:316-484) + two induction passes
 * (:646-1017).  Because the virtual sentinel makes all n+1 suffixes distinct, the suffix array is
 * UNIQUE (the reference's own validator demands strict '<' between neighbours, main.cpp:261-264),
 * so the oracle restates WHAT is computed with an independent algorithm (SA-IS, induced sorting,
 * Nong/Zhang/Chan 2009) rather than re-deriving the 1 500-line sorter.  It is also a different
 * algorithm family from the GPU path (prefix doubling), so agreement between the two is meaningful.
 * The BWT and inverse BWT functions follow the reference's definitions line by line.
 *
 * Parity status: pinned against oracle/_ref (the unmodified reference built from /root/reference)
 * and tests/golden/kat.json — see tests/test_oracle.py.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* SA-IS over an int32 or uint8 string whose last symbol is a unique, smallest sentinel.        */

#define CHR(i) (cs == 4 ? ((const int32_t*)s)[(i)] : (int32_t)((const uint8_t*)s)[(i)])
#define IS_LMS(i) ((i) > 0 && typ[(i)] && !typ[(i) - 1])

static void bucket_bounds(const void* s, int32_t* bkt, int32_t n, int32_t K, int cs, int want_end)
{
    int32_t i, sum = 0;
    for (i = 0; i <= K; ++i) bkt[i] = 0;
    for (i = 0; i < n; ++i) bkt[CHR(i)]++;
    for (i = 0; i <= K; ++i) {
        sum += bkt[i];
        bkt[i] = want_end ? sum : sum - bkt[i];
    }
}

static void induce_L(const uint8_t* typ, int32_t* SA, const void* s, int32_t* bkt, int32_t n, int32_t K, int cs)
{
    int32_t i, j;
    bucket_bounds(s, bkt, n, K, cs, 0);
    for (i = 0; i < n; ++i) {
        j = SA[i] - 1;
        if (j >= 0 && !typ[j]) SA[bkt[CHR(j)]++] = j;
    }
}

static void induce_S(const uint8_t* typ, int32_t* SA, const void* s, int32_t* bkt, int32_t n, int32_t K, int cs)
{
    int32_t i, j;
    bucket_bounds(s, bkt, n, K, cs, 1);
    for (i = n - 1; i >= 0; --i) {
        j = SA[i] - 1;
        if (j >= 0 && typ[j]) SA[--bkt[CHR(j)]] = j;
    }
}

/* s[0..n-1], s[n-1] unique smallest; symbols in [0,K]; SA has n entries. */
static int sais_rec(const void* s, int32_t* SA, int32_t n, int32_t K, int cs)
{
    int32_t i, j, n1, name, prev;
    uint8_t* typ;
    int32_t* bkt;
    int32_t *SA1, *s1;

    if (n == 1) { SA[0] = 0; return 0; }
    typ = (uint8_t*)malloc((size_t)n);
    bkt = (int32_t*)malloc(sizeof(int32_t) * ((size_t)K + 1));
    if (!typ || !bkt) { free(typ); free(bkt); return -1; }

    /* S-type = 1, L-type = 0 */
    typ[n - 1] = 1;
    typ[n - 2] = 0;
    for (i = n - 3; i >= 0; --i) {
        int32_t a = CHR(i), b = CHR(i + 1);
        typ[i] = (uint8_t)((a < b || (a == b && typ[i + 1])) ? 1 : 0);
    }

    /* stage 1: sort LMS substrings */
    bucket_bounds(s, bkt, n, K, cs, 1);
    for (i = 0; i < n; ++i) SA[i] = -1;
    for (i = 1; i < n; ++i)
        if (IS_LMS(i)) SA[--bkt[CHR(i)]] = i;
    induce_L(typ, SA, s, bkt, n, K, cs);
    induce_S(typ, SA, s, bkt, n, K, cs);

    /* compact sorted LMS substrings, name them */
    n1 = 0;
    for (i = 0; i < n; ++i)
        if (IS_LMS(SA[i])) SA[n1++] = SA[i];
    for (i = n1; i < n; ++i) SA[i] = -1;
    name = 0;
    prev = -1;
    for (i = 0; i < n1; ++i) {
        int32_t pos = SA[i];
        int diff = 0;
        int32_t d;
        for (d = 0; d < n; ++d) {
            if (prev == -1 || CHR(pos + d) != CHR(prev + d) || typ[pos + d] != typ[prev + d]) {
                diff = 1;
                break;
            } else if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) {
                break;
            }
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0) SA[j--] = SA[i];

    /* stage 2: solve the reduced problem */
    SA1 = SA;
    s1 = SA + n - n1;
    if (name < n1) {
        if (sais_rec(s1, SA1, n1, name - 1, 4) != 0) { free(typ); free(bkt); return -1; }
    } else {
        for (i = 0; i < n1; ++i) SA1[s1[i]] = i;
    }

    /* stage 3: induce the result */
    bucket_bounds(s, bkt, n, K, cs, 1);
    for (i = 1, j = 0; i < n; ++i)
        if (IS_LMS(i)) s1[j++] = i;
    for (i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (i = n1; i < n; ++i) SA[i] = -1;
    for (i = n1 - 1; i >= 0; --i) {
        j = SA[i];
        SA[i] = -1;
        SA[--bkt[CHR(j)]] = j;
    }
    induce_L(typ, SA, s, bkt, n, K, cs);
    induce_S(typ, SA, s, bkt, n, K, cs);

    free(typ);
    free(bkt);
    return 0;
}

/* msufsort.cpp:1730-1767: result has n+1 entries; :1720 puts (inputSize | flag) at SA[0] and the
 * induction passes strip flags (:860,:961), so SA[0] == n; the remaining entries are the suffixes
 * in the order defined by main.cpp:210-232.  Shifting every byte up by one and appending symbol 0
 * reproduces "virtual sentinel smaller than 0x00" exactly. */
int oracle_make_suffix_array(const uint8_t* text, int64_t n, int32_t* sa_out)
{
    int32_t* s;
    int64_t i;
    int rc;
    if (n < 0 || n > 2147483646LL || !sa_out || (n > 0 && !text)) return -1;
    if (n == 0) { sa_out[0] = 0; return 0; }
    s = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    if (!s) return -1;
    for (i = 0; i < n; ++i) s[i] = (int32_t)text[i] + 1;
    s[n] = 0;
    rc = sais_rec(s, sa_out, (int32_t)(n + 1), 256, 4);
    free(s);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* brute force (tiny n)                                                                        */

static const uint8_t* g_bf_text;
static int64_t g_bf_n;

/* main.cpp:210-232: walk while equal; running off the end first means smaller. */
static int bf_compare(const void* pa, const void* pb)
{
    int64_t a = *(const int32_t*)pa, b = *(const int32_t*)pb;
    if (a == b) return 0;
    while (a < g_bf_n && b < g_bf_n && g_bf_text[a] == g_bf_text[b]) { ++a; ++b; }
    if (a == g_bf_n) return -1;
    if (b == g_bf_n) return 1;
    return g_bf_text[a] < g_bf_text[b] ? -1 : 1;
}

int oracle_make_suffix_array_bruteforce(const uint8_t* text, int64_t n, int32_t* sa_out)
{
    int64_t i;
    if (n < 0 || !sa_out) return -1;
    sa_out[0] = (int32_t)n;
    for (i = 0; i < n; ++i) sa_out[i + 1] = (int32_t)i;
    g_bf_text = text;
    g_bf_n = n;
    qsort(sa_out + 1, (size_t)n, sizeof(int32_t), bf_compare);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* forward BWT                                                                                 */

/* msufsort.cpp:1811-1815: walk the n+1 rows, skip the row whose suffix is 0 (that row number is
 * the return value, :1816), emit the byte preceding each other suffix; row 0 (suffix n) emits the
 * last byte of the text. */
int32_t oracle_bwt_from_sa(const uint8_t* text, int64_t n, const int32_t* sa, uint8_t* bwt_out)
{
    int64_t i;
    int32_t sentinel = 0;
    uint8_t* w = bwt_out;
    for (i = 0; i <= n; ++i) {
        if (sa[i] == 0) sentinel = (int32_t)i;
        else *w++ = text[sa[i] - 1];
    }
    return sentinel;
}

int32_t oracle_forward_bwt(uint8_t* text_inout, int64_t n)
{
    int32_t* sa;
    uint8_t* tmp;
    int32_t sentinel;
    if (n < 0 || n > 2147483646LL) return -1;
    if (n == 0) return 0;
    sa = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    tmp = (uint8_t*)malloc((size_t)n);
    if (!sa || !tmp || oracle_make_suffix_array(text_inout, n, sa) != 0) { free(sa); free(tmp); return -1; }
    sentinel = oracle_bwt_from_sa(text_inout, n, sa, tmp);
    memcpy(text_inout, tmp, (size_t)n);
    free(sa);
    free(tmp);
    return sentinel;
}

/* ------------------------------------------------------------------------------------------ */
/* inverse BWT                                                                                 */

/* Single-threaded restatement of msufsort.cpp:1821-2096.
 *   :1880-1889  F-column starts: n = 1 because row 0 is the sentinel's row.
 *   :1891       index[0] = {sentinelIndex, ...}: the successor of the sentinel row is the row of
 *               suffix 0.
 *   :1898-1915  for each BWT byte i (row r = i + (i >= sentinelIndex)), k = cursor[byte]++ and
 *               index[k].value = r  — i.e. psi[k] = r, the row of the suffix one position later;
 *               the symbol stored with index[k] is the BWT byte of row k.
 *   :1988-2015  walk: starting at i = index[0].value, repeatedly emit index[i].symbol (no emit on
 *               the sentinel row) and follow index[i].value.
 * The reference decodes from 256*threads start rows and stitches (:1922-2095); with one walker
 * that reduces to the loop below.  Emitting L[row] while following psi yields T[0], T[1], ...
 * because L[psi(k)] is the byte that precedes suffix SA[k]+1, i.e. T[SA[k]].  */
int oracle_reverse_bwt(uint8_t* bwt_inout, int64_t n, int32_t sentinel_index)
{
    int64_t cursor[256];
    int64_t i, run, row;
    int32_t* psi;
    uint8_t* out;
    if (n < 0 || n > 2147483646LL) return -1;
    if (n == 0) return 0;
    if (sentinel_index < 1 || (int64_t)sentinel_index > n) return -1;
    psi = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    out = (uint8_t*)malloc((size_t)n);
    if (!psi || !out) { free(psi); free(out); return -1; }
    memset(cursor, 0, sizeof(cursor));
    for (i = 0; i < n; ++i) cursor[bwt_inout[i]]++;
    run = 1;
    for (i = 0; i < 256; ++i) { int64_t c = cursor[i]; cursor[i] = run; run += c; }
    psi[0] = sentinel_index;
    for (i = 0; i < n; ++i) {
        row = i + (i >= sentinel_index);
        psi[cursor[bwt_inout[i]]++] = (int32_t)row;
    }
    row = psi[0];
    for (i = 0; i < n; ++i) {
        row = psi[row];                      /* row of suffix i+1 */
        out[i] = bwt_inout[row - (row > sentinel_index)]; /* its preceding byte = T[i] */
    }
    memcpy(bwt_inout, out, (size_t)n);
    free(psi);
    free(out);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */

int64_t oracle_check_suffix_array(const uint8_t* text, int64_t n, const int32_t* sa)
{
    int32_t* isa;
    int64_t i, bad = 0;
    if (n < 0) return -1;
    if (sa[0] != (int32_t)n) ++bad;
    isa = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    if (!isa) return -1;
    for (i = 0; i <= n; ++i) isa[i] = -1;
    for (i = 0; i <= n; ++i) {
        int64_t v = sa[i];
        if (v < 0 || v > n || isa[v] != -1) ++bad;
        else isa[v] = (int32_t)i;
    }
    if (bad) { free(isa); return bad; }
    for (i = 1; i < n; ++i) {
        int64_t a = sa[i], b = sa[i + 1];
        if (text[a] > text[b]) ++bad;
        else if (text[a] == text[b] && !(isa[a + 1] < isa[b + 1])) ++bad;
    }
    free(isa);
    return bad;
}

/* ---------------------------------------------------------------------------------------------
 * LCP array.  Two restatements:
 *  (1) oracle_make_lcp_array: the reference demo's construction (main.cpp:16-64): match_length()
 *      compares from a known common-prefix length; lcp() halves a range of the suffix array, using
 *      the match length of its two end points as the starting length for the left half.
 *  (2) oracle_lcp_kasai: Kasai et al.'s linear-time algorithm (different family, cross-check).
 * Output convention of this repository (include/b200sa.h): n+1 entries aligned with the SA,
 * lcp[0] = lcp[1] = 0, lcp[r] = lcp(SA[r-1], SA[r]).  The reference's output[i] is lcp[i+2]. */
static int32_t match_length_(const uint8_t* t, int64_t n, int32_t a, int32_t b, int32_t len)
{
    if (a > b) { int32_t x = a; a = b; b = x; }
    while ((int64_t)b + len < n && t[a + len] == t[b + len]) ++len;  /* main.cpp:27-37, byte-wise */
    return len;
}

static void lcp_range_(const uint8_t* t, int64_t n, const int32_t* sa, int32_t* out, int64_t lo, int64_t size, int32_t cur)
{
    /* out[lo+i] = lcp(sa[lo+i], sa[lo+i+1]) for i < size, every pair shares at least `cur` bytes (main.cpp:42-64) */
    while (size > 4) {
        const int64_t mid = size / 2;
        const int32_t next = match_length_(t, n, sa[lo], sa[lo + mid], cur);
        lcp_range_(t, n, sa, out, lo, mid, next);
        lo += mid;
        size -= mid;
    }
    for (int64_t i = 0; i < size; ++i) out[lo + i] = match_length_(t, n, sa[lo + i], sa[lo + i + 1], cur);
}

int oracle_make_lcp_array(const uint8_t* text, int64_t n, const int32_t* sa, int32_t* lcp_out)
{
    if (n < 0 || !sa || !lcp_out) return -1;
    lcp_out[0] = 0;
    if (n >= 1) lcp_out[1] = 0;
    if (n >= 2) {
        /* rows 1..n hold real suffixes; pair (r, r+1) goes to lcp_out[r+1]: shift the output by one */
        lcp_range_(text, n, sa + 1, lcp_out + 2, 0, n - 1, 0);
    }
    return 0;
}

int oracle_lcp_kasai(const uint8_t* text, int64_t n, const int32_t* sa, int32_t* lcp_out)
{
    if (n < 0 || !sa || !lcp_out) return -1;
    int32_t* isa = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!isa) return -1;
    for (int64_t r = 0; r <= n; ++r) isa[sa[r]] = (int32_t)r;
    lcp_out[0] = 0;
    int64_t l = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = isa[i];          /* r >= 1: row 0 is the empty suffix */
        const int64_t j = sa[r - 1];
        if (j == n) l = 0;
        else while (i + l < n && j + l < n && text[i + l] == text[j + l]) ++l;
        lcp_out[r] = (int32_t)l;
        if (l > 0) --l;
    }
    free(isa);
    return 0;
}

uint64_t oracle_fnv1a64(const void* data, int64_t nbytes)
{
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xcbf29ce484222325ULL;
    int64_t i;
    for (i = 0; i < nbytes; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}
