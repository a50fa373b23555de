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
// Tuning knobs (overridable at compile time for experiments: make NVCC_DEFS="-DB200SA_RS_IPT=8 ...").
#ifndef B200SA_RS_IPT
#define B200SA_RS_IPT 16
#endif
#ifndef B200SA_RS_MIN_BLOCKS
#define B200SA_RS_MIN_BLOCKS 4
#endif
#ifndef B200SA_RS_LOOKBACK_DEPTH
#define B200SA_RS_LOOKBACK_DEPTH 4
#endif
#ifndef B200SA_RS_PEERS_ATOMIC_OR
#define B200SA_RS_PEERS_ATOMIC_OR 0
#endif
#ifndef B200SA_RS_THREADS
#define B200SA_RS_THREADS 256
#endif
// 1: keep the sixteen within-warp ranks of a thread (each < 32 * IPT <= 65535) two to a register: the sweep then fits its
// 80-register budget with at most 4 bytes of spills (80 bytes otherwise).  Measured on B200 (profiles/r02_knobs.txt, 2^28 pairs
// per sweep): 1.76 ms against 1.89 ms; the persistent variant with next-tile key prefetch measured 1.81 ms (1.82 ms with
// packed ranks) and was removed.
#ifndef B200SA_RS_PACK_POS
#define B200SA_RS_PACK_POS 1
#endif
#if B200SA_RS_PACK_POS
#define RS_POS_DECL(IPT) u32 pos2[((IPT) + 1) / 2]; for (int i_ = 0; i_ < ((IPT) + 1) / 2; ++i_) pos2[i_] = 0
#define RS_POS_SET(k, v) pos2[(k) >> 1] |= (u32)(v) << (((k) & 1) * 16)
#define RS_POS_GET(k) ((pos2[(k) >> 1] >> (((k) & 1) * 16)) & 0xffffu)
#else
#define RS_POS_DECL(IPT) u32 pos[IPT]
#define RS_POS_SET(k, v) pos[k] = (v)
#define RS_POS_GET(k) pos[k]
#endif
// 1: a full tile's keys (and values) are fetched by ONE bulk asynchronous copy each (cp.async.bulk, completion on an mbarrier)
// into the staging buffers in shared memory and read from there, instead of sixteen 8-byte + sixteen 4-byte global loads per
// thread (SASS: UBLKCP.S.G, SYNCS.ARRIVE.TRANS64, SYNCS.PHASECHK.TRANS64.TRYWAIT).  The copies are issued by the thread that
// takes the tile ticket, before the counters are cleared, and the raw tile is consumed into registers before the same
// buffers receive the digit-ordered tile.  Measured on B200 (profiles/r02_sweep_bulk.txt): 1.794 ms against 1.819 ms per
// 2^28-pair sweep, 30.26 against 30.47 ms per SA + BWT step.  -DB200SA_RS_BULK_LOAD=0 restores the per-thread loads.
#ifndef B200SA_RS_BULK_LOAD
#define B200SA_RS_BULK_LOAD 1
#endif
// 1: the per-warp digit counters are 16-bit (a warp sees at most 32 x IPT = 512 keys, a tile 4096): 4 KB instead of 8 KB of
// shared memory per CTA, which is what a FOURTH resident CTA per SM needs next to the 48 KB of staging buffers (together with
// B200SA_RS_MIN_BLOCKS = 4, i.e. 64 registers per thread: the u64 sweep then spills 72 bytes per thread, and still wins).
// Measured on B200 (profiles/r02_sweep_4cta.txt): 1.708 ms against 1.796 ms per 2^28-pair sweep with three CTAs per SM,
// 27.13 against 27.98 ms per SA + BWT step.  (-DB200SA_RS_WHIST_U16=0 -DB200SA_RS_MIN_BLOCKS=3 restores the old shape.)
#ifndef B200SA_RS_WHIST_U16
#define B200SA_RS_WHIST_U16 1
#endif
#if B200SA_RS_WHIST_U16
typedef u16 rs_whist_t;
#else
typedef u32 rs_whist_t;
#endif
static const int RS_THREADS = B200SA_RS_THREADS;
static const int RS_IPT = B200SA_RS_IPT;
static const int RS_TILE = RS_THREADS * RS_IPT;  // 4096 pairs per tile
static const int RS_MIN_BLOCKS = B200SA_RS_MIN_BLOCKS;  // 4 CTAs/SM -> 64 registers per thread
static const int RS_MAX_PASSES = 8;

static const u64 RS_FLAG_PARTIAL = 1ull << 62;
static const u64 RS_FLAG_INCLUSIVE = 1ull << 63;
static const u64 RS_VALUE_MASK = (1ull << 62) - 1;

template <typename KeyT>
__host__ __device__ constexpr size_t rs_pass_smem_bytes()
{
    return (size_t)(RS_THREADS / 32) * RS_RADIX * (sizeof(rs_whist_t) + (B200SA_RS_PEERS_ATOMIC_OR ? 4 : 0)) + 3 * RS_RADIX * 4 + 16 * 4 +
           (size_t)RS_TILE * sizeof(KeyT) + (size_t)RS_TILE * 4;
}

#if B200SA_RS_BULK_LOAD && !defined(B200SA_EMU)
__device__ __forceinline__ u32 rs_smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(u64* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(rs_smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(u64* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(rs_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rs_bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rs_smem_addr(dst)), "l"(src), "r"(bytes), "r"(rs_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(u64* bar, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(rs_smem_addr(bar)), "r"(parity) : "memory");
}
#endif

template <typename KeyT>
__device__ __forceinline__ u32 rs_digit(KeyT key, int shift)
{
    return (u32)(key >> shift) & (u32)(RS_RADIX - 1);
}

// ---------------------------------------------------------------------------------------------
// Digit histograms for `npasses` consecutive digits starting at begin_bit: one read of the keys.
// ghist[p*256 + d] += count.
//
// Every warp owns a private [npasses][256] counter table in shared memory and bumps it with plain
// shared-memory atomics.  Measured on B200 (tools/ubench): ATOMS.ADD costs 1.1 (all lanes on one
// word) to 2.4 (32 random words) SM-cycles per warp instruction, eight VOTE.BALLOTs ~20 and
// MATCH.ANY ~60 — the first versions of this kernel grouped equal digits with match_any (13.6 ms
// for 2^28 keys x 8 digits) and then with ballots (7.8 ms).
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
        const bool full = (c + 1u) * chunk <= m;  // warp-uniform
        KeyT k[RH_IPT];
#pragma unroll
        for (int i = 0; i < RH_IPT; ++i) {
            const u32 idx = base + (u32)i * 32u;
            k[i] = (full || idx < m) ? ld_stream(keys + idx) : (KeyT)0;
        }
#pragma unroll
        for (int i = 0; i < RH_IPT; ++i) {
            const bool ok = full || (base + (u32)i * 32u < m);
#pragma unroll
            for (int p = 0; p < RS_MAX_PASSES; ++p) {
                if (p < npasses && ok) atomicAdd(&mine[p * RS_RADIX + rs_digit<KeyT>(k[i], begin_bit + p * RS_RADIX_BITS)], 1u);
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
#ifdef B200SA_PHASE_TIMING
// debug build only: cycles spent per phase of the sweep, summed over warp 0 of every 16th tile
__device__ unsigned long long g_phase_cycles[8];
#define PT_MARK(i) do { if (pt_on) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(now_ - pt_t)); pt_t = now_; } } while (0)
#else
#define PT_MARK(i) do { } while (0)
#endif

//   STABLE=false    -> a multi-split: pairs with equal digits may change their relative order.  That is all the ISA update and
//                      the routing of pairs to their owner GPU need (they only bucket pairs for a scatter), and it replaces
//                      the within-warp ranking — eight ballots, per-warp counters and their reduction over the warps, 40 % of
//                      the sweep's instructions — by ONE shared-memory atomicAdd per pair on the tile's digit counters.
template <typename KeyT, bool WRITE_KEYS, bool STABLE = true>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
k_onesweep_pass(const KeyT* __restrict__ kin, KeyT* __restrict__ kout,
                const u32* __restrict__ vin, u32* __restrict__ vout,
                u32 m, int shift, u32 gen_skip,
                const u32* __restrict__ bins, u64* __restrict__ status, u32* __restrict__ tile_counter)
{
    constexpr int THREADS = RS_THREADS, IPT = RS_IPT, WARPS = THREADS / 32, TILE = THREADS * IPT;
    B200SA_DYN_SMEM(smem);
    rs_whist_t* whist = (rs_whist_t*)smem;  // [WARPS][256] per-warp digit counters, later warp-exclusive prefixes
    u32* s_cnt = (u32*)(whist + WARPS * RS_RADIX);  // [256] tile digit counts
    u32* s_coff = s_cnt + RS_RADIX;      // [256] exclusive scan of s_cnt (slot of the digit run in smem)
    u32* s_gdelta = s_coff + RS_RADIX;   // [256] global offset of the digit run minus s_coff
    u32* s_wtot = s_gdelta + RS_RADIX;   // [16]
#if B200SA_RS_PEERS_ATOMIC_OR
    u32* wmask = s_wtot + 16;            // [WARPS][256] per-warp peer masks (all zero between uses)
    KeyT* skeys = (KeyT*)(wmask + WARPS * RS_RADIX);  // [TILE]
#else
    KeyT* skeys = (KeyT*)(s_wtot + 16);  // [TILE]
#endif
    u32* svals = (u32*)(skeys + TILE);   // [TILE]
    __shared__ u32 s_tile;
#if B200SA_RS_BULK_LOAD && !defined(B200SA_EMU)
    __shared__ __align__(8) u64 s_bar;
#endif

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#if B200SA_RS_BULK_LOAD && !defined(B200SA_EMU)
    constexpr bool BULK = sizeof(KeyT) >= 4;  // bulk copies move multiples of 16 bytes from 16-byte aligned addresses
    if (BULK && tid == 0) {
        s_tile = atomicAdd(tile_counter, 1u);
        rs_mbar_init(&s_bar, 1u);
        const u32 tile0 = s_tile, base0 = tile0 * (u32)TILE;
        // a full tile whose sources are 16-byte aligned: one copy for the keys, one for the values, both signalling s_bar
        if (m - base0 >= (u32)TILE && ((((uintptr_t)kin) | ((uintptr_t)vin)) & 15u) == 0) {
            rs_mbar_expect_tx(&s_bar, (u32)(TILE * sizeof(KeyT)) + (vin ? (u32)(TILE * 4) : 0u));
            rs_bulk_g2s(skeys, kin + base0, (u32)(TILE * sizeof(KeyT)), &s_bar);
            if (vin) rs_bulk_g2s(svals, vin + base0, (u32)(TILE * 4), &s_bar);
        }
    }
    if (!BULK && tid == 0) s_tile = atomicAdd(tile_counter, 1u);
#else
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
#endif
    if (STABLE) {
        for (u32 i = tid; i < (u32)(WARPS * RS_RADIX); i += THREADS) {
            whist[i] = 0;
#if B200SA_RS_PEERS_ATOMIC_OR
            wmask[i] = 0;
#endif
        }
    } else {
        if (tid < (u32)RS_RADIX) s_cnt[tid] = 0;
    }
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * (u32)TILE;
    const u32 valid = min((u32)TILE, m - base);
#ifdef B200SA_PHASE_TIMING
    const bool pt_on = (tid == 0) && ((tile & 15u) == 0) && sizeof(KeyT) == 8;
    long long pt_t = clock64();
    if (pt_on) atomicAdd(&g_phase_cycles[7], 1ull);
#endif

    // ---- load (warp-striped): element order inside the tile is (warp, item, lane)
    KeyT key[IPT];
    u32 val[IPT];
    RS_POS_DECL(IPT);
    const u32 wbase = warp * (32u * IPT) + lane;
    const bool full = valid == (u32)TILE;  // block-uniform: no bounds checks on the common path
#if B200SA_RS_BULK_LOAD && !defined(B200SA_EMU)
    const bool bulk = BULK && full && ((((uintptr_t)kin) | ((uintptr_t)vin)) & 15u) == 0;  // block-uniform, same test as the issuing thread
    if (bulk) {
        rs_mbar_wait(&s_bar, 0u);
#pragma unroll
        for (int k = 0; k < IPT; ++k) key[k] = skeys[wbase + (u32)k * 32u];
    } else
#endif
    if (full) {
        const KeyT* kp = kin + base + wbase;
#pragma unroll
        for (int k = 0; k < IPT; ++k) key[k] = ld_stream(kp + k * 32);
    } else {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const u32 li = wbase + (u32)k * 32u;
            // pads get digit 255 in every pass: they sort to the very end of the tile
            key[k] = li < valid ? ld_stream(kin + base + li) : (KeyT)~(KeyT)0;
        }
    }

    if (sizeof(KeyT) == 8) { volatile u32 sink = (u32)key[IPT - 1]; (void)sink; }
    PT_MARK(0);  // keys arrived
    // ---- 1. rank inside the warp.  Peers = lanes holding my digit.  Two interchangeable ways to find
    // them (tools/ubench on B200, SM-cycles per 32 keys at full occupancy): eight ballots 20.7,
    // atomicOr into a shared mask word + read back 7.4, MATCH.ANY 60.  Inside this kernel the shared
    // memory pipe is the scarcer resource (staging + counters already need ~25 wavefronts per 32
    // keys), so the ballot form is faster in situ (R0 sweep 2.0 ms vs 2.4 ms) and is the default.
    // All peers then read the warp's running count for the digit (one broadcast word) and the lowest
    // peer bumps it by the group size.
    rs_whist_t* mywh = whist + warp * RS_RADIX;
    const u32 lt = lanemask_lt();
#if B200SA_RS_PEERS_ATOMIC_OR
    u32* mymask = wmask + warp * RS_RADIX;
    const u32 mybit = 1u << lane;
#endif
    if (STABLE) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const u32 d = rs_digit<KeyT>(key[k], shift);
#if B200SA_RS_PEERS_ATOMIC_OR
            atomicOr(&mymask[d], mybit);
            __syncwarp();
            const u32 peers = mymask[d];
#elif defined(B200SA_ABLATE_BALLOTS)
            const u32 peers = 1u << lane;  // timing experiment only: wrong ranks
#else
            const u32 peers = warp_peers_digit8(d);
#endif
            const u32 prev = mywh[d];
            __syncwarp();
            const u32 below = (u32)__popc(peers & lt);
            if (below == 0) {
#if B200SA_RS_PEERS_ATOMIC_OR
                mymask[d] = 0;
#endif
                mywh[d] = (rs_whist_t)(prev + (u32)__popc(peers));
            }
            __syncwarp();
            RS_POS_SET(k, prev + below);
        }
    } else {
        // rank inside the tile = the old value of the tile's digit counter; pads of the last tile are not counted
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const bool real = full || (wbase + (u32)k * 32u < valid);
            const u32 p = real ? atomicAdd(&s_cnt[rs_digit<KeyT>(key[k], shift)], 1u) : 0u;
            RS_POS_SET(k, p);
        }
    }
    PT_MARK(1);  // ranking
    // values are fetched only now: during ranking they would cost 16 more live registers (spills at
    // the 80-register budget of 3 CTAs/SM); their latency overlaps the tile-level scan below
#if B200SA_RS_BULK_LOAD && !defined(B200SA_EMU)
    if (bulk && vin) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) val[k] = svals[wbase + (u32)k * 32u];
    } else
#endif
    if (vin) {
        if (full) {
            const u32* vp = vin + base + wbase;
#pragma unroll
            for (int k = 0; k < IPT; ++k) val[k] = ld_stream(vp + k * 32);
        } else {
#pragma unroll
            for (int k = 0; k < IPT; ++k) {
                const u32 li = wbase + (u32)k * 32u;
                val[k] = li < valid ? ld_stream(vin + base + li) : 0u;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const u32 gi = base + wbase + (u32)k * 32u;
            val[k] = gi + (gi >= gen_skip ? 1u : 0u);
        }
    }
    __syncthreads();

    // ---- 2. per-digit: warp-exclusive prefixes, tile count; publish the PARTIAL descriptor early
    u32 my_cnt = 0;
    if (tid < (u32)RS_RADIX) {
        u32 acc = 0;
        if (STABLE) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const u32 c = whist[w * RS_RADIX + tid];
                whist[w * RS_RADIX + tid] = (rs_whist_t)acc;
                acc += c;
            }
        } else {
            acc = s_cnt[tid];
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
        const u32 off = s_cnt[tid] + prefix;
        s_coff[tid] = off;
        if (STABLE) {
            // fold the digit's slot into every warp's exclusive prefix: staging then needs one lookup per key
#pragma unroll
            for (int w = 0; w < WARPS; ++w) whist[w * RS_RADIX + tid] = (rs_whist_t)(whist[w * RS_RADIX + tid] + off);
        }
    }
    __syncthreads();

    PT_MARK(2);  // tile-level combine + scan (incl. barrier waits)
    // ---- stage keys and values in digit order (needs only tile-local offsets; predecessors keep
    // publishing their descriptors meanwhile)
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const u32 d = rs_digit<KeyT>(key[k], shift);
        if (STABLE) {
            const u32 p = RS_POS_GET(k) + mywh[d];
            skeys[p] = key[k];
            svals[p] = val[k];
        } else if (full || (wbase + (u32)k * 32u < valid)) {
            const u32 p = RS_POS_GET(k) + s_coff[d];
            skeys[p] = key[k];
            svals[p] = val[k];
        }
    }

    PT_MARK(3);  // staging
    // ---- 4. look-back for digit tid: B200SA_RS_LOOKBACK_DEPTH predecessor descriptors in flight per step
    if (tid < (u32)RS_RADIX) {
        u64 excl = 0;
#ifdef B200SA_ABLATE_LOOKBACK
        if (false) {
#else
        if (tile != 0) {
#endif
            const u64* col = status + tid;
            i64 t = (i64)tile - 1;
            bool done = false;
            while (!done) {
                // LB descriptors in flight per step: the inclusive frontier trails a running tile by some
                // tens of tiles (clock64 phase timing: with 4 in flight the walk back cost 6.3 k of the 25 k
                // cycles of a tile's life); keys/values/ranks are dead by now, so registers are free
                constexpr int LB = B200SA_RS_LOOKBACK_DEPTH;
                u64 v[LB];
#pragma unroll
                for (int i = 0; i < LB; ++i) v[i] = (t - i >= 0) ? ld_relaxed_u64(col + (u64)(t - i) * RS_RADIX) : 0ull;
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    if (!done) {
                        u64 x = v[i];
                        while ((x >> 62) == 0) x = ld_relaxed_u64(col + (u64)(t - i) * RS_RADIX);
                        excl += x & RS_VALUE_MASK;
                        if (x & RS_FLAG_INCLUSIVE) done = true;
                    }
                }
                t -= LB;
            }
            st_relaxed_u64(status + (u64)tile * RS_RADIX + tid, RS_FLAG_INCLUSIVE | (excl + (u64)my_cnt));
        }
        // global start of this tile's run of digit tid, minus its slot in shared memory
        s_gdelta[tid] = bins[tid] + (u32)excl - s_coff[tid];
    }
    PT_MARK(4);  // look-back
    __syncthreads();
    PT_MARK(5);  // barrier after look-back
    // ---- write out: consecutive threads take consecutive slots, i.e. consecutive addresses inside a digit run
    if (full) {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const u32 j = tid + (u32)i * THREADS;
            const KeyT kk = skeys[j];
            const u32 g = s_gdelta[rs_digit<KeyT>(kk, shift)] + j;
            if (WRITE_KEYS) st_stream(kout + g, kk);
            st_stream(vout + g, svals[j]);
        }
    } else {
        for (u32 j = tid; j < valid; j += THREADS) {
            const KeyT kk = skeys[j];
            const u32 g = s_gdelta[rs_digit<KeyT>(kk, shift)] + j;
            if (WRITE_KEYS) st_stream(kout + g, kk);
            st_stream(vout + g, svals[j]);
        }
    }
    PT_MARK(6);  // write-out (issue only; stores retire asynchronously)
}

}  // namespace b200sa
