// bwt_kernels.cuh — forward BWT gather and inverse BWT (psi table, walkers, list ranking).
#pragma once
#include "common.cuh"

namespace b200sa {

// ---------------------------------------------------------------------------------------------
// Forward BWT: out[o] = T[SA[row] - 1] with row = o + (o >= s), s = row of suffix 0 = rank[0].
// Replaces second_stage_its_as_burrows_wheeler_transform (msufsort.cpp:1453-1492) and the serial
// copy-out loop (:1811-1815).  Row 0 holds SA = n, so it emits T[n-1] without a special case.
// Each thread produces 4 consecutive output bytes per step (one 32-bit store); SA is read with
// consecutive lanes on consecutive 16-byte chunks, the text bytes are gathered.
// Once the text is larger than L2 every gathered byte costs one random 32-byte DRAM sector (256 MiB: 4.2 ms, ~64 G
// gathers/s — the random-access rate of HBM, the same rate the LCP probes and the inverse BWT see).  The alternative
// measured in round 2 — text order: one radix sweep of (rank[i], T[i-1]) pairs by row window, then a scatter inside
// L2-resident windows — LOST on B200: the sweep takes 1.46 ms, but 2^28 one-byte stores take 3.3 ms (sub-word stores are
// read-modify-write operations for the ECC-protected L2: ncu shows 6.9 GB of DRAM traffic for 1.6 GB of algorithmic
// bytes), 4.3 ms with 32-bit OR reductions; 4.9 / 5.8 ms against 4.2 ms for this gather (profiles/r02_negative_results.md).
static const int BW_THREADS = 256;
static const int BW_STEPS = 4;  // independent 4-byte groups in flight per thread

// Suffix array entries of the four rows behind output bytes o .. o+3 (row = o + (o >= s)).  One 4-byte load per row made
// the gather passes L2-REQUEST bound (2^28 requests per pass: 1.07 ms per pass of the windowed kernel whatever it hit).  Where
// the group is complete and sa + o is 16-byte aligned the four (or five) entries come with one 128-bit load plus at most
// one more word: rows before the sentinel row are sa[o .. o+3], rows behind it sa[o+1 .. o+4].
__device__ __forceinline__ void bw_load_rows(const i32* __restrict__ sa, u32 o, u32 o_end, u32 s, bool vec_ok, u32 v[4])
{
    if (vec_ok && o + 4u <= o_end) {
        const int4 q = ld_stream((const int4*)(sa + o));
        const u32 a[4] = {(u32)q.x, (u32)q.y, (u32)q.z, (u32)q.w};
        if (o + 3u < s) {
            v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3];
        } else {
            const u32 e = (u32)ld_stream(sa + o + 4u);  // o + 4 <= o_end <= n: inside the n + 1 entries
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] = (o + (u32)b >= s) ? (b < 3 ? a[b + 1] : e) : a[b];
        }
    } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const u32 ob = o + (u32)b;
            v[b] = ob < o_end ? (u32)ld_stream(sa + ob + (ob >= s ? 1u : 0u)) : 1u;  // 1: a harmless text position for bytes that are not produced
        }
    }
}

// Produces output bytes [o_begin, o_end) (the whole transform: 0 and n; a sharded run: the bytes
// of the rows this GPU owns).
__global__ void __launch_bounds__(BW_THREADS)
k_bwt_gather(const u8* __restrict__ text, const i32* __restrict__ sa, const u32* __restrict__ rank,
             u32 o_begin, u32 o_end, u8* __restrict__ out, i32* __restrict__ sentinel_out)
{
    const u32 s = rank[0];
    if (blockIdx.x == 0 && threadIdx.x == 0 && sentinel_out) *sentinel_out = (i32)s;
    const u32 span = o_end - o_begin;
    const u32 ngroups = (u32)div_up_u64(span, 4);
    const bool out_aligned = (((uintptr_t)(out + o_begin)) & 3u) == 0;
    const bool vec_ok = (((uintptr_t)(sa + o_begin)) & 15u) == 0;
    for (u32 g0 = (blockIdx.x * BW_THREADS + threadIdx.x); g0 < ngroups; g0 += gridDim.x * BW_THREADS * BW_STEPS) {
        u32 packed[BW_STEPS];
        u32 v[BW_STEPS][4];
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            if (g < ngroups) bw_load_rows(sa, o_begin + g * 4u, o_end, s, vec_ok, v[st]);
        }
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            u32 w = 0;
            if (g < ngroups) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const u32 o = o_begin + g * 4u + (u32)b;
                    if (o < o_end) w |= (u32)text[v[st][b] - 1u] << (8 * b);
                }
            }
            packed[st] = w;
        }
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            if (g < ngroups) {
                const u32 o = o_begin + g * 4u;
                if (out_aligned && o + 4u <= o_end) {
                    *(u32*)(out + o) = packed[st];
                } else {
                    for (u32 b = 0; b < 4u && o + b < o_end; ++b) out[o + b] = (u8)(packed[st] >> (8 * b));
                }
            }
        }
    }
}

// The same gather for texts larger than L2, in passes over TEXT WINDOWS that fit L2: pass p reads the suffix array rows
// again (streaming, evict-first) but gathers only the bytes whose text position lies in window p, so every gather hits a
// window that stays resident in the 126 MB L2.  Measured on B200: 224 G gathers/s when the text fits L2 (64 MiB) against
// 64 G/s at 256 MiB and 43 G/s at 1 GiB in one pass.  A pass costs about 1 ms per 2^28 rows, so the trade pays for up to
// four windows (256 MiB in three windows: 3.2 instead of 4.35 ms) and not beyond (host: Engine::bwt_max_passes).  The first pass writes whole output words (bytes outside its window as 0), the later
// passes OR their bytes in; out must be 4-byte aligned (the host falls back to the single pass otherwise).
__global__ void __launch_bounds__(BW_THREADS)
k_bwt_gather_window(const u8* __restrict__ text, const i32* __restrict__ sa, const u32* __restrict__ rank,
                    u32 o_begin, u32 o_end, u8* __restrict__ out, i32* __restrict__ sentinel_out, u32 win_lo, u32 win_hi, int first_pass)
{
    const u32 s = rank[0];
    if (blockIdx.x == 0 && threadIdx.x == 0 && sentinel_out) *sentinel_out = (i32)s;
    const u32 span = o_end - o_begin;
    const u32 ngroups = (u32)div_up_u64(span, 4);
    u32* out32 = (u32*)(out + o_begin);
    const bool vec_ok = (((uintptr_t)(sa + o_begin)) & 15u) == 0;
    for (u32 g0 = (blockIdx.x * BW_THREADS + threadIdx.x); g0 < ngroups; g0 += gridDim.x * BW_THREADS * BW_STEPS) {
        u32 packed[BW_STEPS];
        u32 v[BW_STEPS][4];
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            if (g < ngroups) bw_load_rows(sa, o_begin + g * 4u, o_end, s, vec_ok, v[st]);
        }
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            u32 w = 0;
            if (g < ngroups) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const u32 o = o_begin + g * 4u + (u32)b;
                    const u32 t = v[st][b] - 1u;
                    if (o < o_end && t >= win_lo && t < win_hi) w |= (u32)text[t] << (8 * b);
                }
            }
            packed[st] = w;
        }
#pragma unroll
        for (int st = 0; st < BW_STEPS; ++st) {
            const u32 g = g0 + (u32)st * gridDim.x * BW_THREADS;
            if (g < ngroups) {
                const u32 o = g * 4u;
                if (o + 4u <= span) {
                    if (first_pass) st_stream(out32 + g, packed[st]);
                    else if (packed[st]) out32[g] |= packed[st];
                } else {
                    // the last, partial word: bytes, so that nothing beyond out[o_end - 1] is touched
                    for (u32 b = 0; b < 4u && o + b < span; ++b) {
                        const u8 v = (u8)(packed[st] >> (8 * b));
                        if (first_pass) out[o_begin + o + b] = v;
                        else if (v) out[o_begin + o + b] |= v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// *differ |= (a[0..n) != b[0..n)).  Both buffers are workspace allocations (16-byte aligned).  Used by the host entry points
// to recognise that the text they are handed is byte for byte the one whose suffix array is still resident (the
// reference's users call make_suffix_array and forward_burrows_wheeler_transform on the same bytes: one sort serves both).
__global__ void __launch_bounds__(256)
k_bytes_differ(const u8* __restrict__ a, const u8* __restrict__ b, u64 n, u32* __restrict__ differ)
{
    const u64 nvec = n / 16;
    bool d = false;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (u64)gridDim.x * blockDim.x) {
        const uint4 x = ((const uint4*)a)[v], y = ((const uint4*)b)[v];
        d |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
    }
    if (blockIdx.x == 0)
        for (u64 i = nvec * 16 + threadIdx.x; i < n; i += blockDim.x) d |= a[i] != b[i];
    if (d) atomicOr(differ, 1u);
}

// ---------------------------------------------------------------------------------------------
// Inverse BWT.
//
// Rows of the conceptual (n+1)-row matrix: row 0 is the sentinel's row ("" / suffix n), L[row] for
// row != s is bwt[row - (row > s)], L[s] is the sentinel.  psi[k] = the row r with LF(r) = k, i.e.
// the row of the suffix one text position to the right of row k's suffix; psi[0] = s.  psi[1..n]
// is exactly a stable 8-bit counting sort of the rows by their BWT byte — one onesweep pass with
// generated values (radix_sort.cuh), the reference's phase C (msufsort.cpp:1898-1915).  The first
// byte of row k's suffix is the symbol whose F-column range contains k (fstart[]), so no symbol is
// stored next to psi.
//
// Decoding follows psi from row s (suffix 0) and emits F[row] at every step.  It is split over
// walkers seeded at every D-th row plus row s (the reference seeds 256 x threads walkers,
// :1922-1944).  D is a power of two, so "is this row a seed" is (row & (D-1)) == 0 || row == s: psi
// carries no mark bit and all 32 bits of an entry are row number (texts up to 2^32 - 8194 bytes).
// Pass A measures each walker's segment (length, successor walker) and decodes it into a private
// window; a pointer-jumping list ranking turns the segment chain into text offsets; pass B copies
// the windows to their final positions (and finishes the few overlong segments).
//
// Untrusted input: psi is a permutation of the rows whatever the bytes are, but only a real BWT
// makes it ONE cycle 0 -> s -> ... -> 0.  On anything else some walkers sit on cycles that never
// reach the terminal (row 0) and the list ranking adds lengths around them, so pass B first checks,
// per walker, that its chain ended in the terminal and that its segment lies inside the text
// (len <= dist <= n), and that the walker of row s is n bytes from the end.  A violation raises the
// `bad` flag (the host entry point then fails with B200SA_EINVAL) and the walker writes nothing.

// fstart[c] = first F-row of symbol c (rows 1..n), fstart[256] = n+1.  bins = exclusive byte counts.
__global__ void k_unbwt_fstart(const u32* __restrict__ bins, u32 n, u32* __restrict__ fstart, u32* __restrict__ psi, u32 s, u32* __restrict__ bad)
{
    const u32 t = threadIdx.x;
    if (t < 256) fstart[t] = bins[t] + 1u;
    if (t == 0) { fstart[256] = n + 1u; psi[0] = s; *bad = 0; }
}

// walker w < nreg starts at row w*D (walker 0 = row 0 is the terminal, it never walks);
// walker nreg (only if s % D != 0) starts at row s.  D = 1 << dshift.
__device__ __forceinline__ u32 ub_walker_row(u32 w, u32 nreg, int dshift, u32 s) { return w < nreg ? w << dshift : s; }
__device__ __forceinline__ bool ub_is_seed(u32 row, u32 dmask, u32 s) { return (row & dmask) == 0u || row == s; }
__device__ __forceinline__ u32 ub_row_walker(u32 row, u32 nreg, int dshift, u32 dmask)
{
    return (row & dmask) == 0u ? row >> dshift : nreg;  // only called on seed rows
}

// symbol of F-row `row` (row >= 1): the largest c with f[c] <= row; f = 257-entry table in shared memory
__device__ __forceinline__ u32 ub_row_symbol(const u32* f, u32 row)
{
    u32 lo = 0, hi = 256;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const u32 mid = (lo + hi) >> 1;
        if (f[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// Pass A: walkers [w_begin, w_end) (a sharded run gives every GPU a slice; the psi table is replicated)
// follow psi until the next seed row, record segment length and successor, and decode on the way: the
// bytes go to the walker's private window of `cap` bytes in `scratch` (8 bytes per store); a walker
// whose segment is longer than its window remembers the row where it stopped storing (ovf_row) and only
// counts from there.  Segment lengths are close to exponentially distributed with mean D, so with
// cap = 4 D about 2 % of the bytes are decoded a second time by k_unbwt_place instead of 100 % in the
// first version (measure pass + emit pass).  The next psi load is issued before the symbol search of
// the current row, so the 8-step search hides under the DRAM latency of the dependent load.
static const int UW_THREADS = 128;
static const u32 UB_NO_OVERFLOW = 0xffffffffu;

__global__ void __launch_bounds__(UW_THREADS)
k_unbwt_walk(const u32* __restrict__ psi, const u32* __restrict__ fstart, u32 w_begin, u32 w_end, u32 nreg, int dshift, u32 s,
             u32 cap, u8* __restrict__ scratch, u32* __restrict__ seg_len, u32* __restrict__ seg_next, u32* __restrict__ ovf_row)
{
    __shared__ u32 s_f[257];
    for (u32 i = threadIdx.x; i < 257; i += UW_THREADS) s_f[i] = fstart[i];
    __syncthreads();
    const u32 w = w_begin + blockIdx.x * UW_THREADS + threadIdx.x;
    if (w >= w_end) return;
    if (w == 0) { seg_len[0] = 0; seg_next[0] = 0; ovf_row[0] = UB_NO_OVERFLOW; return; }  // terminal node points to itself
    const u32 dmask = (1u << dshift) - 1u;
    u64* win = (u64*)(scratch + (u64)w * cap);  // cap is a multiple of 8 and scratch is 256-byte aligned
    u32 cur = ub_walker_row(w, nreg, dshift, s);
    u32 nxt = psi[cur];
    u32 len = 0, ovf = UB_NO_OVERFLOW;
    u64 acc = 0;
    do {
        const u32 nxt2 = psi[nxt];                    // dependent load first ...
        if (len < cap) {
            const u64 c = ub_row_symbol(s_f, cur);    // ... symbol search while it is in flight
            acc |= c << (8 * (len & 7u));
            if ((len & 7u) == 7u) { win[len >> 3] = acc; acc = 0; }
        } else if (len == cap) {
            ovf = cur;                                // first row whose byte did not fit
        }
        ++len;
        cur = nxt;
        nxt = nxt2;
    } while (!ub_is_seed(cur, dmask, s));
    if (len < cap && (len & 7u)) win[len >> 3] = acc;  // partial last word (inside the window: cap % 8 == 0)
    seg_len[w] = len;
    seg_next[w] = ub_row_walker(cur, nreg, dshift, dmask);
    ovf_row[w] = ovf;
}

// Pointer jumping (Wyllie): dist[w] = bytes emitted from walker w to the end of the text.
// One step: (next, dist) <- (next[next], dist + dist[next]); the terminal has next = itself, dist 0.
__global__ void __launch_bounds__(256)
k_unbwt_jump(const u32* __restrict__ next_in, const u32* __restrict__ dist_in,
             u32* __restrict__ next_out, u32* __restrict__ dist_out, u32 nwalkers)
{
    const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwalkers) return;
    const u32 nx = next_in[w];
    dist_out[w] = dist_in[w] + dist_in[nx];
    next_out[w] = next_in[nx];
}

// Where pass B stores text byte p.  LocalOut: this GPU's buffer.  ShardedOut: the text is sharded by position over the GPUs
// of the box (byte p lives on GPU p >> shift, at offset p of that GPU's peer-mapped buffer) and a walker's bytes are
// stored over NVLink in runs of a whole segment.
struct LocalOut {
    u8* out;
    __device__ __forceinline__ u8& operator()(u32 p) const { return out[p]; }
};
struct ShardedOut {
    u8* base[kMaxPeers];
    int shift;
    __device__ __forceinline__ u8& operator()(u32 p) const { return base[p >> shift][p]; }
};

// Sharded inverse BWT: the (length, successor) entries of walkers [w_begin, w_end) go to word w (lengths) and word W + w
// (successors) of every peer's inbox payload.
struct PeerBcast {
    u32* dst[kMaxPeers];  // null for this GPU itself
    int nparts;
};
__global__ void __launch_bounds__(256)
k_peer_bcast_segments(const u32* __restrict__ seg_len, const u32* __restrict__ seg_next, u32 w_begin, u32 w_end, u32 W, PeerBcast pb)
{
    for (u32 w = w_begin + blockIdx.x * blockDim.x + threadIdx.x; w < w_end; w += gridDim.x * blockDim.x) {
        const u32 l = seg_len[w], x = seg_next[w];
        for (int g = 0; g < pb.nparts; ++g) {
            if (pb.dst[g]) { pb.dst[g][w] = l; pb.dst[g][W + w] = x; }
        }
    }
}

// Pass B: one warp per walker copies the decoded window to its final place (text offset n - dist[w]);
// the few walkers that outgrew their window continue decoding from ovf_row straight into the text.
// start_walker = the walker seeded at row s (text offset 0): it must be exactly n bytes from the end.
static const int UP_THREADS = 256;

template <typename OutT>
__global__ void __launch_bounds__(UP_THREADS)
k_unbwt_place(const u32* __restrict__ psi, const u32* __restrict__ fstart, const u32* __restrict__ dist, const u32* __restrict__ final_next,
              const u32* __restrict__ seg_len, const u32* __restrict__ ovf_row, const u8* __restrict__ scratch, u32 cap,
              u32 w_begin, u32 w_end, u32 n, u32 start_walker, OutT out, u32* __restrict__ bad)
{
    __shared__ u32 s_f[257];
    for (u32 i = threadIdx.x; i < 257; i += UP_THREADS) s_f[i] = fstart[i];
    __syncthreads();
    const u32 lane = threadIdx.x & 31u;
    const u32 w = w_begin + blockIdx.x * (UP_THREADS / 32) + (threadIdx.x >> 5);
    if (w >= w_end || w == 0) return;
    const u32 len = seg_len[w];
    const u32 d = dist[w];
    if (final_next[w] != 0u || d > n || len > d || (w == start_walker && d != n)) {  // not a BWT: see the note on untrusted input
        if (lane == 0) atomicOr(bad, 1u);
        return;
    }
    const u32 pos = n - d;
    const u32 stored = len < cap ? len : cap;
    const u8* src = scratch + (u64)w * cap;
    for (u32 i = lane; i < stored; i += 32u) out(pos + i) = src[i];
    if (len > cap && lane == 0) {
        u32 cur = ovf_row[w];
        u32 nxt = psi[cur];
        u32 o = pos + cap;
        for (u32 k = cap; k < len; ++k) {
            const u32 nxt2 = psi[nxt];
            out(o++) = (u8)ub_row_symbol(s_f, cur);
            cur = nxt;
            nxt = nxt2;
        }
    }
}

}  // namespace b200sa
