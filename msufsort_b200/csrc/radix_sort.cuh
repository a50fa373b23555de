// radix_sort.cuh — stable LSD radix sort for (key, u32 value) pairs, 8-bit digits, one scatter
// sweep per digit ("onesweep": chained scan with decoupled look-back across tiles).
//
// Replaces, in the reference, the role of multikey_quicksort (msufsort.cpp:488-642): the reference
// orders suffixes by recursive 7-way partitioning on 4-byte words with one random text read per
// element per level; here every doubling round orders (group, rank[i+h], i) tuples with a handful
// of streaming passes whose cost is independent of the input's LCP structure.
//
// Per pass and tile (THREADS x IPT consecutive elements, warp-striped so that element order inside
// a warp is (item, lane)):
//   1. warp-level ranking: eight ballots give the lanes holding the same digit; the lowest such
//      lane bumps the warp's private counter in shared memory; rank = old count + #peers below me.
//   2. per-digit reduction over the warps (thread d owns digit d), published as this tile's PARTIAL
//      descriptor; look-back over the predecessors' descriptors until an INCLUSIVE one is found;
//      publish INCLUSIVE.  A descriptor is one 64-bit word (2 flag bits | 62-bit count).
//   3. exclusive scan of the tile's digit counts -> every element's slot in shared memory; keys and
//      values are staged there in digit order and written out with consecutive threads touching
//      consecutive global addresses inside each digit run.
// Tile ids are handed out by an atomic ticket, so every predecessor of a running tile is running or
// finished and the look-back cannot deadlock.
#pragma once
#include "common.cuh"

namespace b200sa {

static const int RS_RADIX_BITS = 8;
static const int RS_RADIX = 256;
static const int RS_THREADS = 256;
static const int RS_IPT = 16;
static const int RS_TILE = RS_THREADS * RS_IPT;  // 4096 pairs per tile
static const int RS_MIN_BLOCKS = 3;              // 3 CTAs/SM -> <= 85 registers per thread
static const int RS_MAX_PASSES = 8;

static const u64 RS_FLAG_PARTIAL = 1ull << 62;
static const u64 RS_FLAG_INCLUSIVE = 1ull << 63;
static const u64 RS_VALUE_MASK = (1ull << 62) - 1;

template <typename KeyT>
__host__ __device__ constexpr size_t rs_pass_smem_bytes()
{
    return (size_t)(RS_THREADS / 32) * RS_RADIX * 4 + 3 * RS_RADIX * 4 + 16 * 4 + (size_t)RS_TILE * sizeof(KeyT) + (size_t)RS_TILE * 4;
}

template <typename KeyT>
__device__ __forceinline__ u32 rs_digit(KeyT key, int shift)
{
    return (u32)(key >> shift) & (u32)(RS_RADIX - 1);
}

// ---------------------------------------------------------------------------------------------
// Digit histograms for `npasses` consecutive digits starting at begin_bit: one read of the keys.
// ghist[p*256 + d] += count.
//
// Every warp owns a private [npasses][256] counter table in shared memory and updates it with plain
// (non-atomic) read-modify-writes: eight ballots group the lanes holding the same digit, the
// lowest lane of each group adds the group size.  Shared-memory atomics on scattered addresses cost
// ~2 cycles per lane on this part (B300_MICROARCH.md, "ATOMS spread-addr") — the first version of
// this kernel used them and took 16 ms for 2^28 keys x 8 digits; plain LDS/STS are bank-limited only.
static const int RH_THREADS = 128;
static const int RH_WARPS = RH_THREADS / 32;
static const int RH_IPT = 8;

__host__ __device__ constexpr size_t rh_smem_bytes(int npasses) { return (size_t)RH_WARPS * npasses * RS_RADIX * 4; }

template <typename KeyT>
__global__ void __launch_bounds__(RH_THREADS)
k_radix_hist(const KeyT* __restrict__ keys, u32 m, int begin_bit, int npasses, u32* __restrict__ ghist)
{
    B200SA_DYN_SMEM(smem);
    u32* sh = (u32*)smem;  // [RH_WARPS][npasses][256]
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 per_warp = (u32)npasses * RS_RADIX;
    for (u32 i = tid; i < RH_WARPS * per_warp; i += RH_THREADS) sh[i] = 0;
    __syncthreads();
    u32* mine = sh + warp * per_warp;
    // a warp takes chunks of 32*RH_IPT consecutive keys
    const u32 chunk = 32u * RH_IPT;
    const u32 nchunks = (u32)div_up_u64(m, chunk);
    for (u32 c = blockIdx.x * RH_WARPS + warp; c < nchunks; c += gridDim.x * RH_WARPS) {
        const u32 base = c * chunk + lane;
        KeyT k[RH_IPT];
        bool ok[RH_IPT];
#pragma unroll
        for (int i = 0; i < RH_IPT; ++i) {
            const u32 idx = base + (u32)i * 32u;
            ok[i] = idx < m;
            k[i] = ok[i] ? ld_stream(keys + idx) : (KeyT)0;
        }
        for (int p = 0; p < npasses; ++p) {
            const int shift = begin_bit + p * RS_RADIX_BITS;
            u32* row = mine + p * RS_RADIX;
#pragma unroll
            for (int i = 0; i < RH_IPT; ++i) {
                const u32 d = rs_digit<KeyT>(k[i], shift);
                // out-of-range lanes are masked out of everybody's peer set
                const u32 peers = warp_peers_digit8(d) & __ballot_sync(B200SA_FULL_MASK, ok[i]);
                const u32 leader = (u32)__ffs((int)peers) - 1u;
                if (ok[i] && lane == leader) row[d] += (u32)__popc(peers);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (u32 i = tid; i < per_warp; i += RH_THREADS) {
        u32 c = 0;
#pragma unroll
        for (int w = 0; w < RH_WARPS; ++w) c += sh[w * per_warp + i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// Turns each pass's 256 counts into exclusive offsets, in place.  grid = npasses, block = 256.
__global__ void __launch_bounds__(RS_RADIX)
k_radix_scan_bins(u32* __restrict__ ghist)
{
    __shared__ u32 wtot[RS_RADIX / 32];
    u32* h = ghist + blockIdx.x * RS_RADIX;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 c = h[tid];
    const u32 incl = warp_incl_scan_u32(c);
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    u32 prefix = 0;
    for (u32 w = 0; w < warp; ++w) prefix += wtot[w];
    h[tid] = prefix + incl - c;
}

// ---------------------------------------------------------------------------------------------
// One scatter sweep on digit (key >> shift) & 255.
//   vin == nullptr  -> the value of element i is i + (i >= gen_skip)   (element indices / BWT rows)
//   WRITE_KEYS=false-> only the permuted values are written (psi table of the inverse BWT)
template <typename KeyT, bool WRITE_KEYS>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
k_onesweep_pass(const KeyT* __restrict__ kin, KeyT* __restrict__ kout,
                const u32* __restrict__ vin, u32* __restrict__ vout,
                u32 m, int shift, u32 gen_skip,
                const u32* __restrict__ bins, u64* __restrict__ status, u32* __restrict__ tile_counter)
{
    constexpr int THREADS = RS_THREADS, IPT = RS_IPT, WARPS = THREADS / 32, TILE = THREADS * IPT;
    B200SA_DYN_SMEM(smem);
    u32* whist = (u32*)smem;             // [WARPS][256] per-warp digit counters, later warp-exclusive prefixes
    u32* s_cnt = whist + WARPS * RS_RADIX;  // [256] tile digit counts
    u32* s_coff = s_cnt + RS_RADIX;      // [256] exclusive scan of s_cnt (slot of the digit run in smem)
    u32* s_gdelta = s_coff + RS_RADIX;   // [256] global offset of the digit run minus s_coff
    u32* s_wtot = s_gdelta + RS_RADIX;   // [16]
    KeyT* skeys = (KeyT*)(s_wtot + 16);  // [TILE]
    u32* svals = (u32*)(skeys + TILE);   // [TILE]
    __shared__ u32 s_tile;

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (u32 i = tid; i < (u32)(WARPS * RS_RADIX); i += THREADS) whist[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * (u32)TILE;
    const u32 valid = min((u32)TILE, m - base);

    // ---- load (warp-striped): element order inside the tile is (warp, item, lane)
    KeyT key[IPT];
    u32 val[IPT];
    u32 pos[IPT];
    const u32 wbase = warp * (32u * IPT) + lane;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const u32 li = wbase + (u32)k * 32u;
        if (li < valid) {
            const u32 gi = base + li;
            key[k] = ld_stream(kin + gi);
            val[k] = vin ? ld_stream(vin + gi) : gi + (gi >= gen_skip ? 1u : 0u);
        } else {
            key[k] = (KeyT)~(KeyT)0;  // digit 255 in every pass: pads sort to the very end of the tile
            val[k] = 0;
        }
    }

    // ---- 1. rank inside the warp
    u32* mywh = whist + warp * RS_RADIX;
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const u32 d = rs_digit<KeyT>(key[k], shift);
        const u32 peers = warp_peers_digit8(d);
        const u32 leader = (u32)__ffs((int)peers) - 1u;
        u32 prev = 0;
        if (lane == leader) {
            prev = mywh[d];
            mywh[d] = prev + (u32)__popc(peers);
        }
        prev = __shfl_sync(B200SA_FULL_MASK, prev, (int)leader);
        pos[k] = prev + (u32)__popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- 2. per-digit: warp-exclusive prefixes, tile count; publish the PARTIAL descriptor early
    u32 my_cnt = 0;
    if (tid < (u32)RS_RADIX) {
        u32 acc = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const u32 c = whist[w * RS_RADIX + tid];
            whist[w * RS_RADIX + tid] = acc;
            acc += c;
        }
        my_cnt = acc;
        st_relaxed_u64(status + (u64)tile * RS_RADIX + tid, (tile == 0 ? RS_FLAG_INCLUSIVE : RS_FLAG_PARTIAL) | (u64)acc);
        // ---- 3a. exclusive scan of the 256 tile counts (8 full warps)
        const u32 incl = warp_incl_scan_u32(acc);
        if (lane == 31) s_wtot[warp] = incl;
        s_cnt[tid] = incl - acc;  // warp-local exclusive, fixed up below
    }
    __syncthreads();
    if (tid < (u32)RS_RADIX) {
        u32 prefix = 0;
        for (u32 w = 0; w < warp; ++w) prefix += s_wtot[w];
        s_coff[tid] = s_cnt[tid] + prefix;
    }
    __syncthreads();

    // ---- stage keys in digit order (needs only tile-local offsets; predecessors keep publishing meanwhile)
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const u32 d = rs_digit<KeyT>(key[k], shift);
        pos[k] += s_coff[d] + mywh[d];
        skeys[pos[k]] = key[k];
    }

    // ---- 4. look-back for digit tid: four predecessor descriptors in flight per step
    if (tid < (u32)RS_RADIX) {
        u64 excl = 0;
        if (tile != 0) {
            const u64* col = status + tid;
            i64 t = (i64)tile - 1;
            bool done = false;
            while (!done) {
                u64 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (t - i >= 0) ? ld_relaxed_u64(col + (u64)(t - i) * RS_RADIX) : 0ull;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!done) {
                        u64 x = v[i];
                        while ((x >> 62) == 0) x = ld_relaxed_u64(col + (u64)(t - i) * RS_RADIX);
                        excl += x & RS_VALUE_MASK;
                        if (x & RS_FLAG_INCLUSIVE) done = true;
                    }
                }
                t -= 4;
            }
            st_relaxed_u64(status + (u64)tile * RS_RADIX + tid, RS_FLAG_INCLUSIVE | (excl + (u64)my_cnt));
        }
        // global start of this tile's run of digit tid, minus its slot in shared memory
        s_gdelta[tid] = bins[tid] + (u32)excl - s_coff[tid];
    }
    __syncthreads();
    if (WRITE_KEYS) {
#pragma unroll 4
        for (u32 j = tid; j < valid; j += THREADS) {
            const KeyT kk = skeys[j];
            const u32 d = rs_digit<KeyT>(kk, shift);
            st_stream(kout + (s_gdelta[d] + j), kk);
        }
    }
    // ---- stage and write values
#pragma unroll
    for (int k = 0; k < IPT; ++k) svals[pos[k]] = val[k];
    __syncthreads();
#pragma unroll 4
    for (u32 j = tid; j < valid; j += THREADS) {
        const u32 d = rs_digit<KeyT>(skeys[j], shift);
        st_stream(vout + (s_gdelta[d] + j), svals[j]);
    }
}

}  // namespace b200sa
