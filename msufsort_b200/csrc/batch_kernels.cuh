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
__global__ void __launch_bounds__(PK_THREADS)
k_pack_keys_batch(const u8* __restrict__ text, u32 n, const u8* __restrict__ code, int bits, int k, int len_bits,
                  const u32* __restrict__ ends, u32 count, u64* __restrict__ keys)
{
    __shared__ u8 s_code[256];
    __shared__ __align__(16) u8 s_sym[PK_TILE + PK_HALO];
    __shared__ u64 s_out[PK_THREADS / 32][PK_IPT * 33];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    s_code[tid] = code[tid];
    __syncthreads();
    const u32 ntiles = (u32)div_up_u64(n, PK_TILE);
    const u64 sym_mask = (k * bits >= 64) ? ~0ull : ((1ull << (k * bits)) - 1ull);
    const int bshift = k * bits + len_bits;  // < 64 whenever count > 1 (plan_alphabet reserves the block bits)
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
        for (int j = 0; j < k; ++j) win = (win << bits) | (u64)s_sym[p0 + j];
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
                const u64 w = r == (u32)k ? win : (r ? (win & (~0ull << (((u32)k - r) * (u32)bits))) : 0ull);
                key = (bshift < 64 ? ((u64)b << bshift) : 0ull) | (w << len_bits) | (u64)r;
            }
            s_out[warp][i * 33 + lane] = key;
            win = ((win << bits) | (u64)s_sym[p0 + i + k]) & sym_mask;
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

}  // namespace b200sa
