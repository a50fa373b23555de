// lcp_kernels.cuh — longest-common-prefix array from a finished suffix array.
//
// Reference role: the demo's LCP construction (main.cpp:16-159: match_length :16-38, the divide-and-conquer
// lcp() :42-64, lcp_multithreaded :68-105).  The reference compares neighbouring suffixes byte by byte and
// carries the match length down a binary recursion over the suffix array (O(n log n) compares in the good
// case, O(n * LCP) on repetitive inputs).  Here the work is done in TEXT order (the permuted LCP array):
//
//   phi[i]  = the suffix that precedes suffix i in the suffix array          (scatter of SA, bucketed)
//   plcp[i] = lcp(i, phi[i]),   and   plcp[i] >= plcp[i-d] - d   for every d  (Kasai's inequality)
//   lcp[r]  = plcp[SA[r]]                                                     (gather)
//
// The inequality is used hierarchically so that every position is an independent thread: level L handles the
// positions that are odd multiples of S = 2^L, each starting from the lower bound given by the position S to
// its left, which a coarser level has already finished (one launch pair per level: a fused shared-memory kernel
// for the five finest levels measured 8.9 ms against 7.2 ms for the five launches, profiles/r01_negative_results.md).  The sum of all extensions of one level is < 3n
// bytes whatever the text looks like, and the extension of one position beyond a small budget is handed to a
// second kernel in which a whole CTA compares 8 KB per iteration, so periodic / Fibonacci inputs (lcp ~ n)
// cost a bounded number of streaming passes instead of n * LCP byte compares.
#pragma once
#include "common.cuh"

namespace b200sa {

static const int LC_THREADS = 256;
static const u32 LC_BUDGET = 256;      // bytes one thread extends a match before handing it to a CTA
static const int LC_UNITS = 4;         // 8-byte units per thread per iteration of the CTA compare
static const u32 LC_NONE = 0xffffffffu;

// 8 text bytes starting at byte x (little endian: byte x is bits 0..7).  `words` is the text pointer rounded
// down to 4 bytes, `off` what was rounded away.  All of x..x+7 must be valid text bytes; only aligned words
// that contain at least one of them are read (an allocation never ends inside an aligned word).
__device__ __forceinline__ u64 lc_load8(const u32* __restrict__ words, u32 off, u32 x)
{
    const u64 X = (u64)x + off;
#ifdef B200SA_EMU_ASAN
    // sanitizer build of the emulator: tests hand in host arrays of odd sizes as "device" text, so only the 8 bytes themselves
    // are touched here (the word-wise form below may read up to 3 bytes past them inside the last aligned word)
    u64 v;
    memcpy(&v, (const u8*)words + X, 8);
    return v;
#endif
    const u32* w = words + (X >> 2);
    const u32 sh = (u32)(X & 3u) * 8u;
    const u32 w0 = w[0], w1 = w[1];
    u32 lo = w0, hi = w1;
    if (sh) {
        const u32 w2 = w[2];
        lo = __funnelshift_r(w0, w1, sh);
        hi = __funnelshift_r(w1, w2, sh);
    }
    return ((u64)hi << 32) | lo;
}

// One level of the hierarchy.  Sample j of the level is text position first + j * step (top level: first = 0,
// from scratch; other levels: first = S, step = 2S, lower bound from position p - S).  A match that is still
// running after LC_BUDGET bytes is appended to the overflow list as (p, bytes matched so far).
__global__ void __launch_bounds__(LC_THREADS)
k_plcp_level(const u8* __restrict__ text, u32 n, const u32* __restrict__ phi, u32* __restrict__ plcp,
             u32 first, u32 step, u32 back /*0 = top level*/, u32 nsamples,
             u32* __restrict__ ovf_pos, u32* __restrict__ ovf_len, u32* __restrict__ ovf_count)
{
    const u32 off = (u32)((uintptr_t)text & 3u);
    const u32* words = (const u32*)(text - off);
    for (u32 j = blockIdx.x * LC_THREADS + threadIdx.x; j < nsamples; j += gridDim.x * LC_THREADS) {
        const u32 p = first + j * step;  // < n by construction of nsamples
        const u32 q = phi[p];
        u32 l = 0;
        if (back) {
            const u32 prev = plcp[p - back];
            l = prev > back ? prev - back : 0u;
        }
        if (q >= n) { plcp[p] = 0; continue; }  // preceded by the empty suffix (row 0)
        const u32 maxl = n - (p > q ? p : q);   // the shorter of the two suffixes
        const u32 stop = (maxl - l) > LC_BUDGET ? l + LC_BUDGET : maxl;
        bool open = true;  // no mismatch found yet
        while (open && l + 8u <= stop) {
            const u64 x = lc_load8(words, off, p + l) ^ lc_load8(words, off, q + l);
            if (x) { l += (u32)(__ffsll((long long)x) - 1) >> 3; open = false; }
            else l += 8u;
        }
        while (open && l < stop) {
            if (text[p + l] != text[q + l]) open = false;
            else ++l;
        }
        if (open && l < maxl) {
            const u32 e = atomicAdd(ovf_count, 1u);
            ovf_pos[e] = p;
            ovf_len[e] = l;
        } else {
            plcp[p] = l;
        }
    }
}

// A whole CTA extends ONE match: 256 threads x LC_UNITS x 8 bytes per iteration.  Must be called by all threads of
// the CTA with identical arguments; returns lcp(p, q) given that the first l bytes are known to match.
__device__ __forceinline__ u32 lc_cta_extend(const u8* __restrict__ text, const u32* __restrict__ words, u32 off, u32 n,
                                             u32 p, u32 q, u32 l, u32* s_min /*[LC_THREADS/32]*/, u32* s_res)
{
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 maxl = n - (p > q ? p : q);
    const u32 chunk = LC_THREADS * LC_UNITS * 8u;
    u32 result = LC_NONE;
    while (result == LC_NONE) {
        // first mismatching byte among this thread's units of [l, l + chunk), LC_NONE if all equal
        u64 xa[LC_UNITS], xb[LC_UNITS];
        u32 mine = LC_NONE;
#pragma unroll
        for (int k = 0; k < LC_UNITS; ++k) {
            const u64 u = (u64)l + ((u32)k * LC_THREADS + tid) * 8u;
            const bool ok = u + 8u <= (u64)maxl;
            xa[k] = ok ? lc_load8(words, off, p + (u32)u) : 0ull;
            xb[k] = ok ? lc_load8(words, off, q + (u32)u) : 0ull;
        }
#pragma unroll
        for (int k = LC_UNITS - 1; k >= 0; --k) {
            const u64 x = xa[k] ^ xb[k];
            if (x) mine = l + ((u32)k * LC_THREADS + tid) * 8u + ((u32)(__ffsll((long long)x) - 1) >> 3);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const u32 o = __shfl_xor_sync(B200SA_FULL_MASK, mine, d);
            mine = mine < o ? mine : o;
        }
        if (lane == 0) s_min[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            u32 m = LC_NONE;
            for (int w = 0; w < LC_THREADS / 32; ++w) m = m < s_min[w] ? m : s_min[w];
            if (m == LC_NONE) {
                // whole chunk equal; the last (partial) chunk leaves fewer than 8 bytes to thread 0
                const u64 next = (u64)l + chunk;
                if (next + 8u > (u64)maxl) {
                    u32 t = l + ((maxl - l) & ~7u);
                    while (t < maxl && text[p + t] == text[q + t]) ++t;
                    m = t;
                }
            }
            *s_res = m;
        }
        __syncthreads();
        result = *s_res;  // rewritten only after the next call's / iteration's first barrier
        l += chunk;
    }
    return result;
}

// Overflow list of one level: one CTA per entry at a time.
__global__ void __launch_bounds__(LC_THREADS)
k_plcp_overflow(const u8* __restrict__ text, u32 n, const u32* __restrict__ phi, u32* __restrict__ plcp,
                const u32* __restrict__ ovf_pos, const u32* __restrict__ ovf_len, const u32* __restrict__ ovf_count)
{
    __shared__ u32 s_min[LC_THREADS / 32];
    __shared__ u32 s_res;
    const u32 off = (u32)((uintptr_t)text & 3u);
    const u32* words = (const u32*)(text - off);
    const u32 count = *ovf_count;
    for (u32 e = blockIdx.x; e < count; e += gridDim.x) {
        const u32 p = ovf_pos[e];
        const u32 result = lc_cta_extend(text, words, off, n, p, phi[p], ovf_len[e], s_min, &s_res);
        if (threadIdx.x == 0) plcp[p] = result;
    }
}

// Direct route for texts whose LCP values are small (experimental, B200SA_LCP_DIRECT=1; not yet measured): row r compares
// suffixes SA[r-1] and SA[r] byte-wise up to LD_BUDGET bytes.  Thread r's right-hand suffix is thread r+1's left-hand
// one, so a row costs ONE random text access (the other hits L1/L2) instead of the three random accesses per position of
// the PLCP route (phi scatter, probe, gather).  Rows that exhaust the budget are counted and listed; the host then
// either finishes the few of them with the CTA-wide compare or, when they are many (repetitive text), discards this
// pass and runs the PLCP route, whose cost is bounded whatever the text looks like.
static const u32 LD_BUDGET = 64;

__global__ void __launch_bounds__(256)
k_lcp_direct(const u8* __restrict__ text, u32 n, const i32* __restrict__ sa, i32* __restrict__ lcp,
             u32* __restrict__ ovf_rows, u32 ovf_cap, u32* __restrict__ ovf_count)
{
    const u32 off = (u32)((uintptr_t)text & 3u);
    const u32* words = (const u32*)(text - off);
    const u64 total = (u64)n + 1;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (u64)gridDim.x * blockDim.x) {
        if (r < 2) { lcp[r] = 0; continue; }  // row 0 is the empty suffix
        const u32 p = (u32)sa[r - 1], q = (u32)sa[r];
        const u32 maxl = n - (p > q ? p : q);
        const u32 stop = maxl > LD_BUDGET ? LD_BUDGET : maxl;
        u32 l = 0;
        bool open = true;
        while (open && l + 8u <= stop) {
            const u64 x = lc_load8(words, off, p + l) ^ lc_load8(words, off, q + l);
            if (x) { l += (u32)(__ffsll((long long)x) - 1) >> 3; open = false; }
            else l += 8u;
        }
        while (open && l < stop) {
            if (text[p + l] != text[q + l]) open = false;
            else ++l;
        }
        lcp[r] = (i32)l;
        if (open && l < maxl) {  // budget exhausted, the match is still running
            const u32 e = atomicAdd(ovf_count, 1u);
            if (e < ovf_cap) ovf_rows[e] = (u32)r;
        }
    }
}

// the listed rows, one CTA at a time: lcp[r] = lcp(SA[r-1], SA[r]) continued from the budget
__global__ void __launch_bounds__(LC_THREADS)
k_lcp_direct_finish(const u8* __restrict__ text, u32 n, const i32* __restrict__ sa, i32* __restrict__ lcp,
                    const u32* __restrict__ ovf_rows, u32 count)
{
    __shared__ u32 s_min[LC_THREADS / 32];
    __shared__ u32 s_res;
    const u32 off = (u32)((uintptr_t)text & 3u);
    const u32* words = (const u32*)(text - off);
    for (u32 e = blockIdx.x; e < count; e += gridDim.x) {
        const u32 r = ovf_rows[e];
        const u32 result = lc_cta_extend(text, words, off, n, (u32)sa[r - 1], (u32)sa[r], LD_BUDGET, s_min, &s_res);
        if (threadIdx.x == 0) lcp[r] = (i32)result;
    }
}

// lcp[r] = lcp(SA[r-1], SA[r]) = plcp[SA[r]] for r = 1..n; lcp[0] = 0 (row 0 is the empty suffix).
__global__ void __launch_bounds__(256)
k_lcp_gather(const i32* __restrict__ sa, u32 n, const u32* __restrict__ plcp, i32* __restrict__ lcp)
{
    const u64 total = (u64)n + 1;
    for (u64 r0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 4; r0 < total; r0 += (u64)gridDim.x * blockDim.x * 4) {
        u32 s[4], v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = (r0 + k < total) ? (u32)sa[r0 + k] : n;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = s[k] < n ? plcp[s[k]] : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r0 + k < total) lcp[r0 + k] = (i32)v[k];
    }
}

}  // namespace b200sa
