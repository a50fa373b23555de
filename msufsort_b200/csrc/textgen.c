/*
 * textgen.c — deterministic synthetic inputs for the SA / BWT / unBWT path.
 *
 * The reference ships no inputs; its self-test draws rand()%sym bytes (main.cpp:274-286).  These
 * closed-form generators give the same bytes to the GPU path, the oracle and the reference build.
 * PRNG: counter-based splitmix64 — word k = mix(seed + (k+1)*0x9E3779B97F4A7C15); byte i of a
 * stream is byte (i mod 8) of word i/8 (little-endian).
 *
 * Built into msufsort_b200/lib/libb200sa_textgen.so (host only, no CUDA).
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define TG_API __attribute__((visibility("default")))

static inline uint64_t mix64(uint64_t x)
{
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t sm_word(uint64_t seed, uint64_t k) { return mix64(seed + (k + 1) * 0x9E3779B97F4A7C15ULL); }

/* config 1: uniform random bytes */
TG_API void textgen_rand(uint8_t* out, int64_t n, uint64_t seed)
{
    int64_t i = 0, k = 0;
    for (; i + 8 <= n; i += 8, ++k) { uint64_t w = sm_word(seed, (uint64_t)k); memcpy(out + i, &w, 8); }
    if (i < n) { uint64_t w = sm_word(seed, (uint64_t)k); memcpy(out + i, &w, (size_t)(n - i)); }
}

/* uniform over an alphabet of `sigma` symbols starting at byte value `base` (sigma=1 -> constant) */
TG_API void textgen_alphabet(uint8_t* out, int64_t n, uint64_t seed, int sigma, int base)
{
    int64_t i;
    if (sigma < 1) sigma = 1;
    for (i = 0; i < n; ++i) {
        uint64_t w = sm_word(seed, (uint64_t)(i >> 2));
        out[i] = (uint8_t)(base + (int)(((w >> (16 * (i & 3))) & 0xFFFF) % (uint64_t)sigma));
    }
}

/* config 2: order-3 Markov "English-like" text over a-z + space.  Context = previous 3 symbols as a
 * base-27 integer (start 0).  From each context four favoured successors s_j = mix(seed,ctx,j) % 27
 * are taken with probability 1/2, 1/4, 1/8, 1/16; the remaining 1/16 is uniform over the alphabet. */
TG_API void textgen_markov3(uint8_t* out, int64_t n, uint64_t seed)
{
    static const char sym[28] = "abcdefghijklmnopqrstuvwxyz ";
    uint32_t ctx = 0;
    int64_t i;
    for (i = 0; i < n; ++i) {
        uint64_t r = sm_word(seed, (uint64_t)i);
        uint32_t u = (uint32_t)(r & 15u), s;
        int j = u < 8 ? 0 : u < 12 ? 1 : u < 14 ? 2 : u == 14 ? 3 : -1;
        if (j >= 0) s = (uint32_t)(mix64(seed ^ (((uint64_t)ctx * 4u + (uint64_t)j + 1u) * 0xD6E8FEB86659FD93ULL)) % 27u);
        else s = (uint32_t)((r >> 8) % 27u);
        out[i] = (uint8_t)sym[s];
        ctx = (ctx * 27u + s) % 19683u;
    }
}

/* config 3: uniform ACGT, then `repeats` copied segments: for each, L log-uniform in [2^8, 2^14),
 * source and destination uniform; every 128th copied base is substituted by a different base. */
TG_API void textgen_acgt_rep(uint8_t* out, int64_t n, uint64_t seed, int64_t repeats)
{
    static const char b[5] = "ACGT";
    int64_t i, r;
    for (i = 0; i < n; ++i) {
        uint64_t w = sm_word(seed, (uint64_t)(i >> 5));
        out[i] = (uint8_t)b[(w >> (2 * (i & 31))) & 3u];
    }
    for (r = 0; r < repeats; ++r) {
        uint64_t w0 = sm_word(seed ^ 0xA5A5A5A5DEADBEEFULL, (uint64_t)(4 * r));
        uint64_t w1 = sm_word(seed ^ 0xA5A5A5A5DEADBEEFULL, (uint64_t)(4 * r + 1));
        uint64_t w2 = sm_word(seed ^ 0xA5A5A5A5DEADBEEFULL, (uint64_t)(4 * r + 2));
        uint64_t w3 = sm_word(seed ^ 0xA5A5A5A5DEADBEEFULL, (uint64_t)(4 * r + 3));
        int e = 8 + (int)(w0 % 6u);
        int64_t L = ((int64_t)1 << e) + (int64_t)(w1 % ((uint64_t)1 << e));
        int64_t src, dst, q;
        if (L >= n) continue;
        src = (int64_t)(w2 % (uint64_t)(n - L));
        dst = (int64_t)(w3 % (uint64_t)(n - L));
        memmove(out + dst, out + src, (size_t)L);
        for (q = 127; q < L; q += 128) {
            uint8_t c = out[dst + q];
            int idx = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
            out[dst + q] = (uint8_t)b[(idx + 1 + (int)(mix64(w0 + (uint64_t)q) % 3u)) & 3];
        }
    }
}

/* config 5a: T[i] = P[i mod p], P = p symbols drawn from "abcd" (forced non-constant for p > 1) */
TG_API void textgen_periodic(uint8_t* out, int64_t n, uint64_t seed, int64_t p)
{
    uint8_t* P;
    int64_t i;
    if (p < 1) p = 1;
    P = (uint8_t*)malloc((size_t)p);
    if (!P) { memset(out, 'a', (size_t)n); return; }
    for (i = 0; i < p; ++i) P[i] = (uint8_t)('a' + (int)((sm_word(seed, (uint64_t)i) >> 7) & 3u));
    if (p > 1) {
        int constant = 1;
        for (i = 1; i < p; ++i) if (P[i] != P[0]) constant = 0;
        if (constant) P[p - 1] = (uint8_t)(P[0] == 'a' ? 'b' : 'a');
    }
    for (i = 0; i < n; ++i) out[i] = P[i % p];
    free(P);
}

/* config 5b: Fibonacci word: a,b = "b","a"; while len(b) < n: a,b = b,b+a; T = b[:n] */
TG_API void textgen_fib(uint8_t* out, int64_t n)
{
    int64_t la = 1, lb = 1;   /* lengths of a and b; b lives in out[0..lb) */
    if (n <= 0) return;
    out[0] = 'a';             /* b = "a", a = "b" */
    while (lb < n) {
        int64_t take = la < n - lb ? la : n - lb;
        /* after the first step a is always the previous b, i.e. the length-la prefix of b */
        if (lb == 1) out[1] = 'b';
        else memmove(out + lb, out, (size_t)take);
        { int64_t nb = lb + la; la = lb; lb = nb; }
    }
}

/* text whose last `tail` bytes are 0x00 (byte-0-versus-sentinel edge case), rest uniform over
 * `sigma` symbols starting at 0 */
TG_API void textgen_zero_tail(uint8_t* out, int64_t n, uint64_t seed, int sigma, int64_t tail)
{
    int64_t i;
    textgen_alphabet(out, n, seed, sigma, 0);
    if (tail > n) tail = n;
    for (i = n - tail; i < n; ++i) out[i] = 0;
}

/* the reference's own self-test input (main.cpp:274-286): srand(seed); byte = rand() % sigma */
TG_API void textgen_reference_selftest(uint8_t* out, int64_t n, unsigned seed, int sigma)
{
    int64_t i;
    srand(seed);
    for (i = 0; i < n; ++i) out[i] = (uint8_t)(rand() % sigma);
}

TG_API uint64_t textgen_fnv1a64(const void* data, int64_t nbytes)
{
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xcbf29ce484222325ULL;
    int64_t i;
    for (i = 0; i < nbytes; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}
