// batch_kernels.cuh — many independent blocks sorted as ONE launch sequence.
//
// The reference handles a batch of blocks (the block-sorting-compressor use of the forward / reverse
// transforms, main.cpp:466-487) by calling the library once per block.  On a GPU a block of a few hundred KiB
// cannot fill 148 SMs and every call pays ~80 kernel launches and a handful of host round trips, so the
// blocks are laid out back to back with ONE separator slot after each block ("expanded" coordinates:
// block b occupies [start_b, end_b), its separator sits at end_b = offsets[b+1] + b, start_b = offsets[b] + b)
// and sorted together by the ordinary doubling rounds with
//
//   key(p) = block(p) << (k*bits + len_bits) | k symbols clamped at the block's end | min(end_b - p, k)
//
// The block number is the most significant part of the key, therefore (1) suffixes of different blocks are
// never compared (a batch of near-identical blocks does not create deep doubling rounds), (2) the rows of
// block b come out contiguous, at exactly the expanded coordinates of the block, with the separator (the
// block's empty suffix, key = b | 0 | 0) first — i.e. SA''[1 + start_b + i] - start_b is the reference's
// suffix array of block b, SA_b[0] = n_b included — and (3) rank[end_b] is smaller than the rank of every
// real suffix of block b, which is all the later rounds need from the virtual sentinel.
#pragma once
#include "common.cuh"
#include "sa_kernels.cuh"
#include "bwt_kernels.cuh"

namespace b200sa {

// smallest b with ends[b] >= p   (ends is increasing; p <= ends[count-1])
__device__ __forceinline__ u32 bt_block_of(const u32* __restrict__ ends, u32 count, u32 p)
{
    u32 lo = 0, hi = count - 1u;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (ends[mid] >= p) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

// packed blocks (no separators) -> expanded text; separator bytes are 0 (never part of a key)
static const int BE_THREADS = 256;
static const int BE_IPT = 16;

__global__ void __launch_bounds__(BE_THREADS)
k_batch_expand(const u8* __restrict__ packed, const u32* __restrict__ ends, u32 count, u32 N, u8* __restrict__ out)
{
    const u32 chunks = (u32)div_up_u64(N, BE_IPT);
    for (u32 c = blockIdx.x * BE_THREADS + threadIdx.x; c < chunks; c += gridDim.x * BE_THREADS) {
        const u32 p0 = c * (u32)BE_IPT;
        u32 b = bt_block_of(ends, count, p0);
        u32 e = ends[b];
#pragma unroll
        for (int i = 0; i < BE_IPT; ++i) {
            const u32 p = p0 + (u32)i;
            if (p < N) {
                while (e < p) { ++b; e = ends[b]; }
                out[p] = (p == e) ? (u8)0 : packed[p - b];
            }
        }
    }
}

// Initial keys of the expanded text (see k_pack_keys for the staging / write-out scheme).
// RADIX: mixed-radix symbol part (see k_pack_keys); `bits` then holds the width of the symbol part, pw[j] = B^j.
template <bool RADIX>
__global__ void __launch_bounds__(PK_THREADS)
k_pack_keys_batch(const u8* __restrict__ text, u32 n, const u8* __restrict__ code, int bits, int k, int len_bits, u64 B,
                  const u64* __restrict__ pw, const u32* __restrict__ ends, u32 count, u64* __restrict__ keys)
{
    __shared__ u8 s_code[256];
    __shared__ __align__(16) u8 s_sym[PK_TILE + PK_HALO];
    __shared__ u64 s_out[PK_THREADS / 32][PK_IPT * 33];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    __shared__ u64 s_pw[66];
    s_code[tid] = RADIX ? (u8)(code[tid] + 1u) : code[tid];
    if (RADIX && tid < 66u) s_pw[tid] = pw[tid];
    __syncthreads();
    const u32 ntiles = (u32)div_up_u64(n, PK_TILE);
    const u64 sym_mask = (RADIX || k * bits >= 64) ? ~0ull : ((1ull << (k * bits)) - 1ull);
    const int bshift = RADIX ? bits : k * bits + len_bits;  // < 64 whenever count > 1 (plan_alphabet reserves the block bits)
    const u64 top_pow = RADIX ? s_pw[k - 1] : 0ull;
    for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u32 base = t * (u32)PK_TILE;
        const bool aligned = (((uintptr_t)(text + base)) & 15u) == 0;
        for (u32 v = tid; v < (u32)(PK_TILE + PK_HALO) / 16; v += PK_THREADS) {
            const u32 g = base + v * 16;
            if (aligned && g + 16 <= n) {
                const uint4 q = *(const uint4*)(text + g);
                const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) s_sym[v * 16 + i] = s_code[(w[i >> 2] >> (8 * (i & 3))) & 255u];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) s_sym[v * 16 + i] = (g + i < n) ? s_code[text[g + i]] : (u8)0;
            }
        }
        __syncthreads();
        const u32 p0 = tid * PK_IPT;
        u64 win = 0;
        for (int j = 0; j < k; ++j) win = RADIX ? win * B + (u64)s_sym[p0 + j] : ((win << bits) | (u64)s_sym[p0 + j]);
        u32 b = 0, e = 0;
        if (base + p0 < n) { b = bt_block_of(ends, count, base + p0); e = ends[b]; }
#pragma unroll
        for (int i = 0; i < PK_IPT; ++i) {
            const u32 gp = base + p0 + i;
            u64 key = ~0ull;
            if (gp < n) {
                while (e < gp) { ++b; e = ends[b]; }
                const u32 rem = e - gp;
                const u32 r = rem < (u32)k ? rem : (u32)k;
                // symbols beyond the block's end belong to the separator / the next block: drop them
                if (RADIX) {
                    const u64 w = r == (u32)k ? win : (r ? win - win % s_pw[(u32)k - r] : 0ull);
                    key = (bshift < 64 ? ((u64)b << bshift) : 0ull) | w;
                } else {
                    const u64 w = r == (u32)k ? win : (r ? (win & (~0ull << (((u32)k - r) * (u32)bits))) : 0ull);
                    key = (bshift < 64 ? ((u64)b << bshift) : 0ull) | (w << len_bits) | (u64)r;
                }
            }
            s_out[warp][i * 33 + lane] = key;
            win = RADIX ? (win - (u64)s_sym[p0 + i] * top_pow) * B + (u64)s_sym[p0 + i + k]
                        : (((win << bits) | (u64)s_sym[p0 + i + k]) & sym_mask);
        }
        __syncwarp();
        const u32 wb = base + warp * (32u * PK_IPT);
#pragma unroll
        for (int r = 0; r < PK_IPT; ++r) {
            const u32 q = (u32)r * 32u + lane;
            const u32 src_lane = q / PK_IPT, src_i = q % PK_IPT;
            const u32 gp = wb + q;
            if (gp < n) st_stream(keys + gp, s_out[warp][src_i * 33 + src_lane]);
        }
        __syncthreads();
    }
}

// Suffix arrays in block-local coordinates: out[j] = SA''[j + 1] - start(block of row j), j = 0..N-1.
// Row j of the global order belongs to the block whose expanded range contains j (see the header comment).
static const int BL_THREADS = 256;
static const int BL_IPT = 8;

__global__ void __launch_bounds__(BL_THREADS)
k_batch_localize(const i32* __restrict__ sa, const u32* __restrict__ ends, u32 count, u32 N, i32* __restrict__ out)
{
    const u32 chunks = (u32)div_up_u64(N, BL_IPT);
    for (u32 c = blockIdx.x * BL_THREADS + threadIdx.x; c < chunks; c += gridDim.x * BL_THREADS) {
        const u32 j0 = c * (u32)BL_IPT;
        u32 b = bt_block_of(ends, count, j0);
        u32 e = ends[b];
        u32 start = b ? ends[b - 1u] + 1u : 0u;
#pragma unroll
        for (int i = 0; i < BL_IPT; ++i) {
            const u32 j = j0 + (u32)i;
            if (j < N) {
                while (e < j) { start = e + 1u; ++b; e = ends[b]; }
                out[j] = (i32)((u32)sa[j + 1u] - start);
            }
        }
    }
}

// Per-block sentinel row: the local row of the block's suffix 0 = rank[start_b] - 1 - start_b (ranks are exact
// global rows once the sort has finished; row 0 of the global order is the virtual sentinel of the whole text).
__global__ void __launch_bounds__(256)
k_batch_sentinels(const u32* __restrict__ rank, const u32* __restrict__ ends, u32 count, i32* __restrict__ sent)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const u32 start = b ? ends[b - 1u] + 1u : 0u;
    sent[b] = ends[b] > start ? (i32)(rank[start] - 1u - start) : 0;
}

// Forward BWT of every block, packed like the input: byte x of the output belongs to block b with
// offs[b] <= x < offs[b+1]; local o = x - offs[b]; row = o + (o >= s_b); out[x] = T''[SA''[1 + start_b + row] - 1].
__global__ void __launch_bounds__(BW_THREADS)
k_bwt_gather_batch(const u8* __restrict__ text, const i32* __restrict__ sa, const u32* __restrict__ offs, const u32* __restrict__ ends,
                   const i32* __restrict__ sent, u32 count, u32 total, u8* __restrict__ out)
{
    const u32 ngroups = (u32)div_up_u64(total, 4);
    const bool out_aligned = (((uintptr_t)out) & 3u) == 0;
    for (u32 g = blockIdx.x * BW_THREADS + threadIdx.x; g < ngroups; g += gridDim.x * BW_THREADS) {
        const u32 x0 = g * 4u;
        // block of x0: smallest b with offs[b+1] > x0
        u32 lo = 0, hi = count - 1u;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (offs[mid + 1u] > x0) hi = mid; else lo = mid + 1u;
        }
        u32 b = lo;
        u32 ob = offs[b], oe = offs[b + 1u];
        u32 s = (u32)sent[b];
        u32 w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 x = x0 + (u32)i;
            if (x < total) {
                while (x >= oe) { ++b; ob = oe; oe = offs[b + 1u]; s = (u32)sent[b]; }
                const u32 o = x - ob;
                const u32 row = ob + b + o + (o >= s ? 1u : 0u);  // start_b = offs[b] + b
                const u32 v = (u32)sa[row + 1u];
                w |= (u32)text[v - 1u] << (8 * i);
            }
        }
        if (out_aligned && x0 + 4u <= total) {
            *(u32*)(out + x0) = w;
        } else {
            for (u32 i = 0; i < 4u && x0 + i < total; ++i) out[x0 + i] = (u8)(w >> (8 * i));
        }
    }
    (void)ends;
}


// =============================================================================================
// Inverse BWT of a batch.  Rows live in the same expanded coordinates as the forward transform: block b owns
// rows [start_b, end_b], local row 0 (= start_b) is the row of the block's empty suffix, the block's text
// starts at local row s_b (its sentinel index).  One table entry per row, 8 bytes:
//     bits 0..31  psi''[row]  (the row of the suffix one text position to the right; same block)
//     bits 32..39 first byte of the row's suffix (what the walker emits when it stands on the row)
//     bit  40     the row is a walker seed / a terminal
// psi and the symbols of all blocks come from ONE stable sort of the BWT bytes by (block << 8 | byte) with the
// rows as values — the batched form of the reference's phase C (msufsort.cpp:1898-1915) — so no per-block
// F-column table is needed.  Walkers: every D-th row, every block's start row and every block's row 0
// (terminal); walker ids are [0, nreg) regular, [nreg, nreg + count) block starts, [nreg + count, nreg + 2 count)
// terminals.  A regular walker whose row is also a start or a terminal row is dead (the special walker owns it).
static const u64 UBB_SYM_SHIFT = 32;
static const u64 UBB_MARK = 1ull << 40;

__device__ __forceinline__ u32 ubb_block_start(const u32* __restrict__ ends, u32 b) { return b ? ends[b - 1u] + 1u : 0u; }

// sort input: key = block << 8 | byte, value = global row of the byte
__global__ void __launch_bounds__(256)
k_ubb_gen(const u8* __restrict__ bwt, const u32* __restrict__ offs, const i32* __restrict__ sent, u32 count, u32 total,
          u32* __restrict__ keys, u32* __restrict__ vals)
{
    const u32 ngroups = (u32)div_up_u64(total, 4);
    for (u32 g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        const u32 x0 = g * 4u;
        u32 lo = 0, hi = count - 1u;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (offs[mid + 1u] > x0) hi = mid; else lo = mid + 1u;
        }
        u32 b = lo, ob = offs[b], oe = offs[b + 1u], s = (u32)sent[b];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 x = x0 + (u32)i;
            if (x < total) {
                while (x >= oe) { ++b; ob = oe; oe = offs[b + 1u]; s = (u32)sent[b]; }
                const u32 o = x - ob;
                keys[x] = (b << 8) | (u32)bwt[x];
                vals[x] = ob + b + o + (o >= s ? 1u : 0u);
            }
        }
    }
}

// table rows 1..n_b of every block from the sorted pairs; sorted position x of block b is row x + b + 1
__global__ void __launch_bounds__(256)
k_ubb_table(const u32* __restrict__ keys, const u32* __restrict__ vals, u32 total, u64* __restrict__ table)
{
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
        const u32 k = ld_stream(keys + x);
        table[x + (k >> 8) + 1u] = ((u64)(k & 255u) << UBB_SYM_SHIFT) | (u64)ld_stream(vals + x);
    }
}

__device__ __forceinline__ u32 ubb_row_walker(u32 row, const u32* __restrict__ ends, const i32* __restrict__ sent, u32 count, u32 nreg, u32 D)
{
    const u32 b = bt_block_of(ends, count, row);
    const u32 start = ubb_block_start(ends, b);
    if (row == start) return nreg + count + b;
    if (row == start + (u32)sent[b]) return nreg + b;
    return row / D;
}

// seed row of walker w; returns false for walkers that do not walk (terminals, dead duplicates, empty blocks)
__device__ __forceinline__ bool ubb_walker_seed(u32 w, const u32* __restrict__ ends, const i32* __restrict__ sent, u32 count, u32 nreg, u32 D,
                                                u32 N, u32* row_out, u32* block_out)
{
    if (w < nreg) {
        const u32 row = w * D;
        if (row >= N) return false;
        const u32 b = bt_block_of(ends, count, row);
        const u32 start = ubb_block_start(ends, b);
        *row_out = row; *block_out = b;
        return row != start && row != start + (u32)sent[b];
    }
    if (w < nreg + count) {
        const u32 b = w - nreg;
        const u32 start = ubb_block_start(ends, b);
        *row_out = start + (u32)sent[b]; *block_out = b;
        return ends[b] > start;  // empty block: nothing to decode
    }
    const u32 b = w - nreg - count;
    *row_out = ubb_block_start(ends, b); *block_out = b;
    return false;
}

// row 0 of every block: psi = the block's start row; then the seed / terminal marks
__global__ void __launch_bounds__(256)
k_ubb_rows0(const u32* __restrict__ ends, const i32* __restrict__ sent, u32 count, u64* __restrict__ table)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const u32 start = ubb_block_start(ends, b);
    table[start] = (u64)(start + (u32)sent[b]) | UBB_MARK;
}

__global__ void __launch_bounds__(256)
k_ubb_mark(u64* __restrict__ table, const u32* __restrict__ ends, const i32* __restrict__ sent, u32 count, u32 nreg, u32 D, u32 N)
{
    const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nreg + count) return;  // terminals are marked by k_ubb_rows0
    u32 row = 0, b = 0;
    if (ubb_walker_seed(w, ends, sent, count, nreg, D, N, &row, &b)) atomicOr((u32*)(table + row) + 1, (u32)(UBB_MARK >> 32));  // high word = symbol | mark
}

__global__ void __launch_bounds__(UW_THREADS)
k_ubb_walk(const u64* __restrict__ table, const u32* __restrict__ ends, const i32* __restrict__ sent, u32 count, u32 nreg, u32 D, u32 N,
           u32 nwalkers, u32 cap, u8* __restrict__ scratch, u32* __restrict__ seg_len, u32* __restrict__ seg_next, u32* __restrict__ ovf_row)
{
    const u32 w = blockIdx.x * UW_THREADS + threadIdx.x;
    if (w >= nwalkers) return;
    u32 cur = 0, blk = 0;
    if (!ubb_walker_seed(w, ends, sent, count, nreg, D, N, &cur, &blk)) {
        seg_len[w] = 0; seg_next[w] = w; ovf_row[w] = UB_NO_OVERFLOW;  // terminal / dead: points to itself
        return;
    }
    u64* win = (u64*)(scratch + (u64)w * cap);
    u64 e = table[cur];
    u32 len = 0, ovf = UB_NO_OVERFLOW;
    u64 acc = 0;
    do {
        const u32 nxt = (u32)e;
        const u64 e2 = table[nxt];
        if (len < cap) {
            acc |= ((e >> UBB_SYM_SHIFT) & 255ull) << (8 * (len & 7u));
            if ((len & 7u) == 7u) { win[len >> 3] = acc; acc = 0; }
        } else if (len == cap) {
            ovf = cur;
        }
        ++len;
        cur = nxt;
        e = e2;
    } while (!(e & UBB_MARK));
    if (len < cap && (len & 7u)) win[len >> 3] = acc;
    seg_len[w] = len;
    seg_next[w] = ubb_row_walker(cur, ends, sent, count, nreg, D);
    ovf_row[w] = ovf;
}

// one warp per walker: window -> offs[block + 1] - dist[w] in the packed output.  Untrusted input (see bwt_kernels.cuh): a
// walker writes only if its chain ended in its own block's terminal and its segment lies inside the block; the block's
// start walker must be exactly the block's size away from the end.  Anything else raises `bad`.
__global__ void __launch_bounds__(UP_THREADS)
k_ubb_place(const u64* __restrict__ table, const u32* __restrict__ ends, const u32* __restrict__ offs, const i32* __restrict__ sent, u32 count,
            u32 nreg, u32 D, u32 N, u32 nwalkers, const u32* __restrict__ dist, const u32* __restrict__ final_next, const u32* __restrict__ seg_len,
            const u32* __restrict__ ovf_row, const u8* __restrict__ scratch, u32 cap, u8* __restrict__ out, u32* __restrict__ bad)
{
    const u32 lane = threadIdx.x & 31u;
    const u32 w = blockIdx.x * (UP_THREADS / 32) + (threadIdx.x >> 5);
    if (w >= nwalkers) return;
    u32 row = 0, blk = 0;
    if (!ubb_walker_seed(w, ends, sent, count, nreg, D, N, &row, &blk)) return;
    const u32 len = seg_len[w];
    const u32 d = dist[w], nb = offs[blk + 1u] - offs[blk];
    if (final_next[w] != nreg + count + blk || d > nb || len > d || (w == nreg + blk && d != nb)) {
        if (lane == 0) atomicOr(bad, 1u);
        return;
    }
    const u32 pos = offs[blk + 1u] - d;
    const u32 stored = len < cap ? len : cap;
    const u8* src = scratch + (u64)w * cap;
    for (u32 i = lane; i < stored; i += 32u) out[pos + i] = src[i];
    if (len > cap && lane == 0) {
        u32 cur = ovf_row[w];
        u64 e = table[cur];
        u32 o = pos + cap;
        for (u32 k = cap; k < len; ++k) {
            const u32 nxt = (u32)e;
            const u64 e2 = table[nxt];
            out[o++] = (u8)(e >> UBB_SYM_SHIFT);
            cur = nxt;
            e = e2;
        }
    }
}

}  // namespace b200sa
