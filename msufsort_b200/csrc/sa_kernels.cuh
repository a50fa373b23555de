// sa_kernels.cuh — prefix-doubling suffix sorting kernels (everything except the radix sort).
//
// Data model of one doubling round with depth h (all arrays in HBM, struct-of-arrays):
//   idx[j]   u32  suffix start of the tuple in slot j of the ACTIVE array (suffixes not yet unique),
//                 in current suffix-array order
//   slot[j]  u32  global suffix-array position (0-based among the n real suffixes) of active slot j;
//                 sorting only permutes tuples inside a group, so this map is invariant in a round
//   gid[j]   u32  dense number of the group slot j belongs to (non-decreasing in j)
//   rank[i]  u32  current rank of suffix i = 1 + global position of its group's head; rank[n] = 0
//                 (the empty suffix / virtual sentinel, row 0 of the reference's SA)
//   key[j]   u64  (gid[j] << rank_bits) | rank[idx[j] + h]   — what the radix sort orders
// After the sort, neighbouring equal keys form the refined groups; groups of size one are final:
// SA[slot + 1] = idx, and they are compacted away.
//
// Reference roles replaced: first_stage_its (msufsort.cpp:1559-1726) by k_byte_hist/k_pack_keys,
// the depth step of multikey_quicksort (:629-637) by k_build_keys, the "partitionSize < 2" exit
// (:516) plus the spread of sorted suffixes (:1702-1720) by the rerank kernels.  Tandem-repeat
// handling (:316-484) and both induction passes (:646-1017) have no counterpart: doubling sorts
// every suffix and is insensitive to periodicity.
#pragma once
#include "common.cuh"

namespace b200sa {

// ---------------------------------------------------------------------------------------------
// Byte histogram of the text (alphabet discovery).  Each thread takes 16 consecutive bytes with one
// 128-bit load and adds run lengths, so constant texts cost one shared atomic per 16 bytes.
static const int BH_THREADS = 256;

__global__ void __launch_bounds__(BH_THREADS)
k_byte_hist(const u8* __restrict__ text, u64 n, u32* __restrict__ hist /*256, u64 counts as 2xu32? no: u32 wraps at 2^32 > n*/)
{
    __shared__ u32 sh[256];
    const u32 tid = threadIdx.x;
    sh[tid] = 0;
    __syncthreads();
    const bool aligned = (((uintptr_t)text) & 15u) == 0;
    const u64 nvec = n / 16;
    for (u64 v = (u64)blockIdx.x * BH_THREADS + tid; v < nvec; v += (u64)gridDim.x * BH_THREADS) {
        u32 w[4];
        if (aligned) {
            const uint4 q = *(const uint4*)(text + v * 16);
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
            const u8* p = text + v * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = (u32)p[4 * i] | ((u32)p[4 * i + 1] << 8) | ((u32)p[4 * i + 2] << 16) | ((u32)p[4 * i + 3] << 24);
        }
        u32 cur = w[0] & 255u, run = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const u32 c = (w[i >> 2] >> (8 * (i & 3))) & 255u;
            if (c == cur) {
                ++run;
            } else {
                atomicAdd(&sh[cur], run);
                cur = c;
                run = 1;
            }
        }
        atomicAdd(&sh[cur], run);
    }
    // tail bytes
    if (blockIdx.x == 0) {
        for (u64 i = nvec * 16 + tid; i < n; i += BH_THREADS) atomicAdd(&sh[text[i]], 1u);
    }
    __syncthreads();
    if (sh[tid]) atomicAdd(&hist[tid], sh[tid]);
}

// ---------------------------------------------------------------------------------------------
// Initial keys.  key(i) = ( sum_{j<k} code[T[i+j]] << bits*(k-1-j) ) << len_bits | min(n-i, k),
// code[] = dense symbol numbers (0-based), positions past the end contribute 0.  Two suffixes that
// agree on the packed symbols but differ in clamped length are ordered shorter-first, which is the
// virtual-sentinel rule of the reference (main.cpp:227-230; sentinel smaller than 0x00), and every
// suffix shorter than k gets a key of its own.
static const int PK_THREADS = 256;
static const int PK_IPT = 16;
static const int PK_TILE = PK_THREADS * PK_IPT;  // 4096 positions
static const int PK_HALO = 64;                   // k <= 58

// RADIX = true (experimental): key = sum_{j<k} digit_j * B^(k-1-j) with digit = code + 1 inside the text and 0 past its end
// (B = sigma + 1, top_pow = B^(k-1)) — the same order, no length field, no bits wasted on odd alphabet sizes.
//
// RANGE = true (sharded runs): only the keys inside [lo, hi) (hi inclusive for the last part) are kept, as (key, suffix) pairs
// compacted through an atomic cursor.  Every GPU of a sharded sort holds the whole text and runs this over all n positions:
// it reads n bytes and writes n/G pairs, instead of packing all n keys and filtering them in a second pass (round 1: 1.35 of
// 8 kernel-ms per step on eight GPUs).  The order of the kept pairs is irrelevant (equal keys form one group whatever
// their order).
// symbol at tile position p in the padded shared-memory layout of k_pack_keys
#define PK_SYM(p) ((u64)s_sym[(p) + (((p) >> 4) << 2)])
template <bool RADIX, bool RANGE>
__global__ void __launch_bounds__(PK_THREADS)
k_pack_keys(const u8* __restrict__ text, u32 n, const u8* __restrict__ code, int bits, int k, int len_bits, u64 B, u64 top_pow,
            u64* __restrict__ keys, u64 lo, u64 hi, int hi_inclusive, u32* __restrict__ out_idx, u32* __restrict__ cursor)
{
    __shared__ u8 s_code[256];
    // symbols of the tile (+ halo) as dense codes.  Thread t works on positions 16 t .. 16 t + 15 (+ k of halo), so a plain
    // layout puts the 32 lanes of a warp on 8 banks (4-way conflicts on every byte access: ncu had this kernel bound by
    // shared-memory wavefronts).  Four pad bytes after every 16 symbols make the lanes' word indices 5 t + c: conflict-free.
    __shared__ __align__(16) u8 s_sym[(PK_TILE + PK_HALO) / 16 * 20];
    __shared__ u64 s_out[RANGE ? 1 : PK_THREADS / 32][RANGE ? 1 : PK_IPT * 33];
    __shared__ u32 s_wsum[PK_THREADS / 32];
    __shared__ u32 s_base, s_total;
    __shared__ u64 s_ck[RANGE ? PK_TILE : 1];  // kept keys of the tile, compacted, so that the global stores are coalesced
    __shared__ u16 s_cp[RANGE ? PK_TILE : 1];  // their positions inside the tile
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    s_code[tid] = RADIX ? (u8)(code[tid] + 1u) : code[tid];  // sigma <= 255 in the mixed-radix layout: digits fit a byte
    __syncthreads();
    const u32 ntiles = (u32)div_up_u64(n, PK_TILE);
    const u64 sym_mask = (k * bits >= 64) ? ~0ull : ((1ull << (k * bits)) - 1ull);
    for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u32 base = t * (u32)PK_TILE;
        // stage the tile's symbols (+halo), already translated to dense codes; 0 past the end
        const bool aligned = (((uintptr_t)(text + base)) & 15u) == 0;
        for (u32 v = tid; v < (u32)(PK_TILE + PK_HALO) / 16; v += PK_THREADS) {
            const u32 g = base + v * 16;
            u32 w[4] = {0, 0, 0, 0};
            u32 c[4] = {0, 0, 0, 0};
            if (aligned && g + 16 <= n) {
                const uint4 q = *(const uint4*)(text + g);
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i >> 2] |= (u32)s_code[(w[i >> 2] >> (8 * (i & 3))) & 255u] << (8 * (i & 3));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i >> 2] |= (u32)((g + i < n) ? s_code[text[g + i]] : (u8)0) << (8 * (i & 3));
            }
            u32* dst = (u32*)(s_sym + v * 20u);  // 20 v is a multiple of 4
            dst[0] = c[0]; dst[1] = c[1]; dst[2] = c[2]; dst[3] = c[3];
        }
        __syncthreads();
        // thread handles 16 consecutive positions with a sliding window
        const u32 p0 = tid * PK_IPT;
        u64 win = 0;
        for (int j = 0; j < k; ++j) win = RADIX ? win * B + (u64)PK_SYM(p0 + (u32)j) : ((win << bits) | (u64)PK_SYM(p0 + (u32)j));
        if (RANGE) {
            u64 kreg[PK_IPT];
            u32 keep = 0;
#pragma unroll
            for (int i = 0; i < PK_IPT; ++i) {
                const u32 gp = base + p0 + i;
                u64 key;
                if (RADIX) {
                    key = win;
                    win = (win - (u64)PK_SYM(p0 + (u32)i) * top_pow) * B + (u64)PK_SYM(p0 + (u32)i + (u32)k);
                } else {
                    const u32 rem = gp < n ? n - gp : 0u;
                    key = (win << len_bits) | (u64)(rem < (u32)k ? rem : (u32)k);
                    win = ((win << bits) | (u64)PK_SYM(p0 + (u32)i + (u32)k)) & sym_mask;
                }
                kreg[i] = key;
                if (gp < n && key >= lo && (hi_inclusive ? key <= hi : key < hi)) keep |= 1u << i;
            }
            const u32 cnt = (u32)__popc(keep);
            const u32 incl = warp_incl_scan_u32(cnt);
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (tid == 0) {
                u32 tot = 0;
                for (int w = 0; w < PK_THREADS / 32; ++w) { const u32 c = s_wsum[w]; s_wsum[w] = tot; tot += c; }
                s_total = tot;
                s_base = tot ? atomicAdd(cursor, tot) : 0u;  // one cursor bump per tile
            }
            __syncthreads();
            u32 d = s_wsum[warp] + incl - cnt;
#pragma unroll
            for (int i = 0; i < PK_IPT; ++i) {
                if ((keep >> i) & 1u) {
                    s_ck[d] = kreg[i];
                    s_cp[d] = (u16)(p0 + (u32)i);
                    ++d;
                }
            }
            __syncthreads();
            const u32 tot = s_total, gb = s_base;
            for (u32 j = tid; j < tot; j += PK_THREADS) {
                keys[gb + j] = s_ck[j];
                out_idx[gb + j] = base + (u32)s_cp[j];
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int i = 0; i < PK_IPT; ++i) {
                const u32 gp = base + p0 + i;
                if (RADIX) {
                    s_out[warp][i * 33 + lane] = win;
                    win = (win - (u64)PK_SYM(p0 + (u32)i) * top_pow) * B + (u64)PK_SYM(p0 + (u32)i + (u32)k);
                } else {
                    const u32 rem = gp < n ? n - gp : 0u;
                    const u64 key = (win << len_bits) | (u64)(rem < (u32)k ? rem : (u32)k);
                    s_out[warp][i * 33 + lane] = key;
                    win = ((win << bits) | (u64)PK_SYM(p0 + (u32)i + (u32)k)) & sym_mask;
                }
            }
            __syncwarp();
            // the warp's 512 keys are positions base + warp*512 + lane*16 + i; write them coalesced
            const u32 wb = base + warp * (32u * PK_IPT);
#pragma unroll
            for (int r = 0; r < PK_IPT; ++r) {
                const u32 q = (u32)r * 32u + lane;  // 0..511 within the warp's chunk
                const u32 src_lane = q / PK_IPT, src_i = q % PK_IPT;
                const u32 gp = wb + q;
                if (gp < n) st_stream(keys + gp, s_out[warp][src_i * 33 + src_lane]);
            }
            __syncthreads();
        }
    }
}

#undef PK_SYM

// ---------------------------------------------------------------------------------------------
// Sharded runs: the splitters come from a sorted regular sample of the initial keys.  The sample keys are computed straight
// from the text (the same key as k_pack_keys: same text + same kernels on every GPU => identical splitters everywhere, no
// communication); k_pack_keys<.., RANGE> then keeps this GPU's key range.
template <bool RADIX>
__global__ void __launch_bounds__(256)
k_sample_keys_text(const u8* __restrict__ text, u32 n, const u8* __restrict__ code, int bits, int k, int len_bits, u64 B,
                   u32 nsample, u32 stride, u64* __restrict__ out)
{
    __shared__ u8 s_code[256];
    s_code[threadIdx.x] = RADIX ? (u8)(code[threadIdx.x] + 1u) : code[threadIdx.x];
    __syncthreads();
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsample) return;
    const u64 p = (u64)i * stride;
    u64 win = 0;
    for (int j = 0; j < k; ++j) {
        const u64 q = p + (u64)j;
        const u64 sym = q < (u64)n ? (u64)s_code[text[q]] : 0ull;
        win = RADIX ? win * B + sym : ((win << bits) | sym);
    }
    const u64 rem = (u64)n - p;
    out[i] = RADIX ? win : ((win << len_bits) | (rem < (u64)k ? rem : (u64)k));
}

// owner-sharded ISA (multi-GPU): positions whose ranks the next round will read, and the serving gather
__global__ void __launch_bounds__(256)
k_make_requests(const u32* __restrict__ idx, u32 m, u32 h, u32 n, u32* __restrict__ pos_out)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        // rank[n] = 0 is known everywhere: a read past the end needs no lookup; ask for the suffix's own rank
        // instead (harmless, keeps every request below n so that position >> shift is a valid owner)
        const u32 i = idx[j];
        const u64 p = (u64)i + h;
        pos_out[j] = p < n ? (u32)p : i;
    }
}

__global__ void __launch_bounds__(256)
k_gather_u32(const u32* __restrict__ pos, u32 count, const u32* __restrict__ table, u32* __restrict__ out)
{
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) out[j] = table[pos[j]];
}

// ---------------------------------------------------------------------------------------------
// rank[n] = 0 (sentinel row), SA[0] = n.
__global__ void k_sa_init(u32* __restrict__ rank, i32* __restrict__ sa, u32 n)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rank[n] = 0;
        sa[0] = (i32)n;
    }
}

// ---------------------------------------------------------------------------------------------
// key[j] = (gid[j] << rank_bits) | rank[idx[j] + h].  Inside a non-singleton group every suffix has
// at least h real symbols, so idx+h <= n; the clamp is defensive.
static const int BK_THREADS = 256;
static const int BK_IPT = 4;

template <typename RankT>
__global__ void __launch_bounds__(BK_THREADS)
k_build_keys(const u32* __restrict__ idx, const u32* __restrict__ gid, RankT rank,
             u32 m, u32 n, u32 h, int rank_bits, u64* __restrict__ keys)
{
    const u32 tile = BK_THREADS * BK_IPT;
    const u32 ntiles = (u32)div_up_u64(m, tile);
    for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u32 base = t * tile + threadIdx.x;
        u32 i[BK_IPT], g[BK_IPT], r[BK_IPT];
#pragma unroll
        for (int q = 0; q < BK_IPT; ++q) {
            const u32 j = base + (u32)q * BK_THREADS;
            i[q] = j < m ? ld_stream(idx + j) : 0u;
            g[q] = j < m ? ld_stream(gid + j) : 0u;
        }
#pragma unroll
        for (int q = 0; q < BK_IPT; ++q) {
            r[q] = rank((u64)i[q] + h);
        }
#pragma unroll
        for (int q = 0; q < BK_IPT; ++q) {
            const u32 j = base + (u32)q * BK_THREADS;
            if (j < m) st_stream(keys + j, ((u64)g[q] << rank_bits) | (u64)r[q]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Re-ranking after a sort: ONE sweep over the m sorted tuples (chained scan, decoupled look-back).
//   head[j]   = key[j] != key[j-1]            (j == 0 is a head)
//   single[j] = head[j] && head[j+1]          (j == m-1: head[m] counts as true)
//   kept[j]   = !single[j]
// Scanned along j: #kept (compaction slot), #kept heads (new dense group id), last head slot (start
// of my group -> new rank).  Items are striped over the block (item = q*256 + tid), so flags come
// from warp ballots, neighbours from shuffles, and every global access is coalesced; a tile's three
// aggregates travel in two self-validating 64-bit descriptors.
static const int RR_THREADS = 256;
static const int RR_IPT = 16;
static const int RR_TILE = RR_THREADS * RR_IPT;  // 4096 tuples per tile
static const int RR_WARPS = RR_THREADS / 32;
static const int RR_CHUNKS = RR_IPT * RR_WARPS;  // 128 warp-rows of 32 items per tile
#ifndef B200SA_RR_MIN_BLOCKS
#define B200SA_RR_MIN_BLOCKS 6
#endif
#ifndef B200SA_RR_PREFETCH
#define B200SA_RR_PREFETCH 1
#endif
static const int RR_MIN_BLOCKS = B200SA_RR_MIN_BLOCKS;

// descriptor A: flag(2) | kept heads (31) | kept, low 31 bits;  descriptor B: flag(2) | kept, bit 31 | 1 + last head slot (32).
// Counts of up to 2^32-1 tuples fit (the uint32 entry points): a kept group has at least two members, so the number
// of kept heads stays below 2^31, and the 32nd bit of the kept count travels in the spare bits of B.
static const u64 RR_FLAG_PARTIAL = 1ull << 62;
static const u64 RR_FLAG_INCLUSIVE = 2ull << 62;
__host__ __device__ __forceinline__ u64 rr_pack_a(u64 flag, u32 kept, u32 kheads) { return flag | ((u64)kheads << 31) | (u64)(kept & 0x7fffffffu); }
__host__ __device__ __forceinline__ u64 rr_pack_b(u64 flag, u32 kept, u32 last_head1) { return flag | ((u64)(kept >> 31) << 32) | (u64)last_head1; }
__host__ __device__ __forceinline__ u32 rr_ab_kept(u64 a, u64 b) { return (u32)(a & 0x7fffffffull) | ((u32)((b >> 32) & 1ull) << 31); }
__host__ __device__ __forceinline__ u32 rr_a_kheads(u64 a) { return (u32)((a >> 31) & 0x7fffffffull); }
__host__ __device__ __forceinline__ u32 rr_b_last_head1(u64 b) { return (u32)(b & 0xffffffffull); }

//   slot_in == nullptr  -> round 0: active slot j is global position slot_base + j (slot_base = number
//       of suffixes owned by lower-numbered parts in a sharded run, else 0)
//   newrank_out != nullptr -> new ranks are written in slot order (coalesced) instead of being
//       scattered into rank[]; the caller then runs the bucketed ISA update (k_scatter_pairs)
//   gstart_out[g] = first active slot of new group g (gstart_out[groups] = next m)
//   info[0] = #kept (next m), info[1] = #kept heads (next group count), written by the last tile
__global__ void __launch_bounds__(RR_THREADS, RR_MIN_BLOCKS)
k_rerank(const u64* __restrict__ keys, const u32* __restrict__ idx_in, const u32* __restrict__ slot_in, u32 slot_base, u32 m,
         u64* __restrict__ desc /*[2][ntiles]*/, u32 ntiles, u32* __restrict__ tile_counter,
         u32* __restrict__ rank, u32* __restrict__ newrank_out, i32* __restrict__ sa,
         u32* __restrict__ idx_out, u32* __restrict__ slot_out, u32* __restrict__ gid_out, u32* __restrict__ gstart_out,
         u32* __restrict__ info)
{
    __shared__ u32 s_k[RR_CHUNKS], s_kh[RR_CHUNKS], s_lh[RR_CHUNKS];  // per chunk: kept, kept heads, 1+last head (tile-local)
    __shared__ u32 s_tile, s_pre_k, s_pre_kh, s_pre_lh;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * (u32)RR_TILE;

    // ---- flags from coalesced key loads + neighbour shuffles
    u64 key[RR_IPT];
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        const u32 j = base + (u32)q * RR_THREADS + tid;
        key[q] = j < m ? ld_stream(keys + j) : 0ull;
    }
    u32 bal_head[RR_IPT], bal_single[RR_IPT];  // warp-uniform ballots per row
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        const u32 j = base + (u32)q * RR_THREADS + tid;
        u64 prev = __shfl_up_sync(B200SA_FULL_MASK, key[q], 1);
        u64 next = __shfl_down_sync(B200SA_FULL_MASK, key[q], 1);
        const bool valid = j < m;
        bool head = false, head_next = true;
        if (valid) {
            if (lane == 0) prev = j > 0 ? keys[j - 1] : ~key[q];
            if (lane == 31) next = (j + 1 < m) ? keys[j + 1] : ~key[q];
            head = (j == 0) || (prev != key[q]);
            head_next = (j + 1 >= m) || (next != key[q]);
        }
        bal_head[q] = __ballot_sync(B200SA_FULL_MASK, head);
        bal_single[q] = __ballot_sync(B200SA_FULL_MASK, head && head_next);
    }
    // ---- chunk aggregates (chunk c = q*8 + warp is item order)
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < RR_IPT; ++q) {
            const u32 rowbase = base + (u32)q * RR_THREADS + warp * 32u;
            const u32 nvalid = rowbase >= m ? 0u : min(32u, m - rowbase);
            const u32 vmask = nvalid >= 32u ? 0xffffffffu : ((1u << nvalid) - 1u);
            const u32 kept = ~bal_single[q] & vmask;
            const u32 c = (u32)q * RR_WARPS + warp;
            s_k[c] = (u32)__popc(kept);
            s_kh[c] = (u32)__popc(kept & bal_head[q]);
            s_lh[c] = bal_head[q] ? ((u32)q * RR_THREADS + warp * 32u + (31u - (u32)__clz((int)bal_head[q])) + 1u) : 0u;
        }
    }
    __syncthreads();
#if B200SA_RR_PREFETCH
    // the suffixes of the tile are fetched now (the key registers are dead): their latency hides behind the look-back below
    u32 sfxr[RR_IPT];
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        const u32 j = base + (u32)q * RR_THREADS + tid;
        sfxr[q] = j < m ? ld_stream(idx_in + j) : 0u;
    }
#endif
    // ---- warp 0: exclusive scan of the 128 chunk aggregates, then look-back for the tile prefix
    if (warp == 0) {
        u32 k4[4], h4[4], l4[4];
        u32 sk = 0, sh = 0, sl = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 c = lane * 4u + (u32)i;
            k4[i] = s_k[c]; h4[i] = s_kh[c]; l4[i] = s_lh[c];
            sk += k4[i]; sh += h4[i]; sl = sl > l4[i] ? sl : l4[i];
        }
        const u32 ik = warp_incl_scan_u32(sk), ih = warp_incl_scan_u32(sh), il = warp_incl_scan_max_u32(sl);
        u32 ek = ik - sk, eh = ih - sh;
        u32 el = __shfl_up_sync(B200SA_FULL_MASK, il, 1);
        if (lane == 0) el = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 c = lane * 4u + (u32)i;
            s_k[c] = ek; s_kh[c] = eh; s_lh[c] = el;
            ek += k4[i]; eh += h4[i]; el = el > l4[i] ? el : l4[i];
        }
        const u32 tot_k = __shfl_sync(B200SA_FULL_MASK, ik, 31);
        const u32 tot_kh = __shfl_sync(B200SA_FULL_MASK, ih, 31);
        const u32 tot_lh_local = __shfl_sync(B200SA_FULL_MASK, il, 31);
        const u32 tot_lh = tot_lh_local ? base + tot_lh_local : 0u;  // 1 + global slot of the tile's last head
        u64* da = desc;
        u64* db = desc + ntiles;
        u32 pre_k = 0, pre_kh = 0, pre_lh = 0;
        if (tile == 0) {
            if (lane == 0) {
                st_relaxed_u64(da + tile, rr_pack_a(RR_FLAG_INCLUSIVE, tot_k, tot_kh));
                st_relaxed_u64(db + tile, rr_pack_b(RR_FLAG_INCLUSIVE, tot_k, tot_lh));
            }
        } else {
            if (lane == 0) {
                st_relaxed_u64(da + tile, rr_pack_a(RR_FLAG_PARTIAL, tot_k, tot_kh));
                st_relaxed_u64(db + tile, rr_pack_b(RR_FLAG_PARTIAL, tot_k, tot_lh));
            }
            // window look-back: lane l inspects tile (t - l)
            i64 t = (i64)tile - 1;
            for (;;) {
                const i64 mine = t - (i64)lane;
                const bool have = mine >= 0;
                u64 a = 0, b = 0;
                if (have) {
                    do {
                        a = ld_relaxed_u64(da + mine);
                        b = ld_relaxed_u64(db + mine);
                    } while ((a >> 62) == 0 || (a >> 62) != (b >> 62));
                }
                const u32 incl = __ballot_sync(B200SA_FULL_MASK, have && (a >> 62) == 2);
                const u32 cutoff = incl ? (u32)__ffs((int)incl) - 1u : 31u;
                const bool use = have && lane <= cutoff;
                u32 ck = use ? rr_ab_kept(a, b) : 0u, ch = use ? rr_a_kheads(a) : 0u;
                u32 cl = use ? rr_b_last_head1(b) : 0u;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    ck += __shfl_xor_sync(B200SA_FULL_MASK, ck, d);
                    ch += __shfl_xor_sync(B200SA_FULL_MASK, ch, d);
                    const u32 o = __shfl_xor_sync(B200SA_FULL_MASK, cl, d);
                    cl = cl > o ? cl : o;
                }
                pre_k += ck; pre_kh += ch; pre_lh = pre_lh > cl ? pre_lh : cl;
                if (incl) break;
                t -= 32;  // no inclusive descriptor in this window: tile 0 is always inclusive, so t stays >= 0
            }
            if (lane == 0) {
                const u32 inc_lh = tot_lh > pre_lh ? tot_lh : pre_lh;
                st_relaxed_u64(da + tile, rr_pack_a(RR_FLAG_INCLUSIVE, pre_k + tot_k, pre_kh + tot_kh));
                st_relaxed_u64(db + tile, rr_pack_b(RR_FLAG_INCLUSIVE, pre_k + tot_k, inc_lh));
            }
        }
        if (lane == 0) {
            s_pre_k = pre_k; s_pre_kh = pre_kh; s_pre_lh = pre_lh;
            if (tile == ntiles - 1) {
                info[0] = pre_k + tot_k;
                info[1] = pre_kh + tot_kh;
                gstart_out[pre_kh + tot_kh] = pre_k + tot_k;  // end marker: gstart[groups] = active tuples
            }
        }
    }
    __syncthreads();
    // ---- outputs
    const u32 pre_k = s_pre_k, pre_kh = s_pre_kh, pre_lh = s_pre_lh;
    const u32 lt = lanemask_lt(), le = lt | (1u << lane);
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        const u32 j = base + (u32)q * RR_THREADS + tid;
        if (j < m) {
            const u32 c = (u32)q * RR_WARPS + warp;
            const u32 rowbase = (u32)q * RR_THREADS + warp * 32u;
            const u32 hb = bal_head[q] & le;
            u32 hs1;  // 1 + global slot of my group's head
            if (hb) hs1 = base + rowbase + (31u - (u32)__clz((int)hb)) + 1u;
            else if (s_lh[c]) hs1 = base + s_lh[c];
            else hs1 = pre_lh;
            const u32 hs = hs1 - 1u;
            const u32 gpos = slot_in ? slot_in[hs] : slot_base + hs;
#if B200SA_RR_PREFETCH
            const u32 sfx = sfxr[q];
#else
            const u32 sfx = ld_stream(idx_in + j);
#endif
            if (newrank_out) st_stream(newrank_out + j, gpos + 1u);  // ISA update deferred: bucketed scatter
            else rank[sfx] = gpos + 1u;
            const bool single = (bal_single[q] >> lane) & 1u;
            if (single) {
                sa[gpos + 1u] = (i32)sfx;
            } else {
                const u32 keptb = ~bal_single[q];
                const u32 dest = pre_k + s_k[c] + (u32)__popc(keptb & lt);
                const u32 heads = pre_kh + s_kh[c] + (u32)__popc(keptb & bal_head[q] & le);
                st_stream(idx_out + dest, sfx);
                st_stream(slot_out + dest, slot_in ? ld_stream(slot_in + j) : slot_base + j);
                st_stream(gid_out + dest, heads - 1u);
                if ((bal_head[q] >> lane) & 1u) gstart_out[heads - 1u] = dest;  // first slot of the new group
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Rounds >= 1 when groups are small (after round 0 of a 256 MiB English-like text 98.6 % of the
// active suffixes sit in groups of <= 16): instead of radix-sorting (gid, rank[i+h]) over 55 bits,
// every group is sorted where it lies.  One thread takes one group of up to GS_TINY members
// (gather the second keys, insertion sort, write suffixes and keys back in place — the array then
// looks exactly as the radix sort would have left it); groups up to GS_MEDIUM go to one CTA each
// (bitonic sort in shared memory); larger ones are listed for the host, which radix-sorts their
// slot ranges individually.  Plays the role of multikey_insertion_sort (msufsort.cpp:223-312) for
// partitions under the reference's insertion_sort_threshold, at one text access per member instead
// of one per compared byte.
static const int GS_TINY = 32;
static const int GS_MEDIUM = 4096;
static const int GS_THREADS = 128;

// counters: [0] #medium groups, [1] #huge groups, [2] tuples in medium groups, [3] tuples in huge groups
template <typename RankT>
__global__ void __launch_bounds__(GS_THREADS)
k_group_sort_tiny(const u32* __restrict__ gstart, u32 groups, u32* __restrict__ idx, RankT rank,
                  u32 n, u32 h, int rank_bits, u64* __restrict__ keys, u32 tiny_max, u32 medium_max,
                  u32* __restrict__ medium_list, u32* __restrict__ huge_list, u32* __restrict__ counters)
{
    const u32 g = blockIdx.x * GS_THREADS + threadIdx.x;
    if (g >= groups) return;
    const u32 s = gstart[g], sz = gstart[g + 1] - s;
    if (sz > tiny_max) {
        if (sz <= medium_max) {
            medium_list[atomicAdd(&counters[0], 1u)] = g;
            atomicAdd(&counters[2], sz);
        } else {
            huge_list[atomicAdd(&counters[1], 1u)] = g;
            atomicAdd(&counters[3], sz);
        }
        return;
    }
    u32 a[GS_TINY], k2[GS_TINY];
    for (u32 i = 0; i < sz; ++i) a[i] = idx[s + i];
    for (u32 i = 0; i < sz; ++i) k2[i] = rank((u64)a[i] + h);
    for (u32 i = 1; i < sz; ++i) {
        const u32 x = k2[i], y = a[i];
        u32 j = i;
        while (j > 0 && k2[j - 1] > x) { k2[j] = k2[j - 1]; a[j] = a[j - 1]; --j; }
        k2[j] = x;
        a[j] = y;
    }
    const u64 hi = (u64)g << rank_bits;
    for (u32 i = 0; i < sz; ++i) {
        idx[s + i] = a[i];
        keys[s + i] = hi | (u64)k2[i];
    }
}

// one CTA per listed group (GS_TINY < size <= GS_MEDIUM): bitonic sort of (rank[i+h] << 32 | suffix)
static const int GM_THREADS = 256;

template <typename RankT>
__global__ void __launch_bounds__(GM_THREADS)
k_group_sort_medium(const u32* __restrict__ list, const u32* __restrict__ gstart, u32* __restrict__ idx,
                    RankT rank, u32 n, u32 h, int rank_bits, u64* __restrict__ keys)
{
    __shared__ u64 buf[GS_MEDIUM];
    const u32 g = list[blockIdx.x];
    const u32 s = gstart[g], sz = gstart[g + 1] - s;
    u32 P = 2;
    while (P < sz) P <<= 1;
    for (u32 i = threadIdx.x; i < P; i += GM_THREADS) {
        u64 v = ~0ull;
        if (i < sz) {
            const u32 sfx = idx[s + i];
            v = ((u64)rank((u64)sfx + h) << 32) | sfx;
        }
        buf[i] = v;
    }
    __syncthreads();
    for (u32 k = 2; k <= P; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 i = threadIdx.x; i < P; i += GM_THREADS) {
                const u32 partner = i ^ j;
                if (partner > i) {
                    const u64 x = buf[i], y = buf[partner];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { buf[i] = y; buf[partner] = x; }
                }
            }
            __syncthreads();
        }
    }
    const u64 hi = (u64)g << rank_bits;
    for (u32 i = threadIdx.x; i < sz; i += GM_THREADS) {
        const u64 v = buf[i];
        idx[s + i] = (u32)v;
        keys[s + i] = hi | (v >> 32);
    }
}

// ---------------------------------------------------------------------------------------------
// ISA update, second half: rank[key[j]] = val[j] over pairs that one radix sweep has bucketed by the
// top 8 bits of the suffix index.  A direct scatter of 2^28 ranks in suffix-array order measured
// 12.9 ms (every 4-byte write dirties a different 32-byte sector of a 1 GiB array: DRAM
// read-modify-write); bucketed, all CTAs write into a window of n/256 entries that lives in L2.
static const int SP_THREADS = 256;
static const int SP_IPT = 8;

__global__ void __launch_bounds__(SP_THREADS)
k_scatter_pairs(const u32* __restrict__ key, const u32* __restrict__ val, u32 m, u32* __restrict__ rank)
{
    const u32 tile = SP_THREADS * SP_IPT;
    const u32 base = blockIdx.x * tile + threadIdx.x;
    u32 k[SP_IPT], v[SP_IPT];
#pragma unroll
    for (int q = 0; q < SP_IPT; ++q) {
        const u32 j = base + (u32)q * SP_THREADS;
        if (j < m) { k[q] = ld_stream(key + j); v[q] = ld_stream(val + j); }
    }
#pragma unroll
    for (int q = 0; q < SP_IPT; ++q) {
        const u32 j = base + (u32)q * SP_THREADS;
        if (j < m) rank[k[q]] = v[q];
    }
}

// The owner's side of the sharded ISA: apply what the G sources left in this GPU's inbox.  Every source's run arrives
// bucketed by the top bits of the suffix index (B buckets per owner) together with its B + 1 bucket offsets (inbox header).
// The runs are applied in BUCKET-major order — bucket 0 of all sources, bucket 1 of all sources, ... — by one launch, so
// that at any moment all CTAs store into the same L2-sized window of the ISA shard.  (Source-major order, one pass over the
// whole shard per source, measured 56 G pairs/s on eight GPUs against 184 G/s for the single-GPU bucketed scatter.)
static const int kInboxCountWords = 64;     // header: pair count per source ...
static const int kInboxOffsStride = 260;    // ... then per source up to 257 bucket offsets (relative to the source's run)
static const int kInboxMaxEntries = 512;    // (bucket, source) pairs of one owner

struct InboxArgs {
    const u32* keys[kMaxPeers];   // region of source s in this GPU's inbox
    const u32* vals[kMaxPeers];
    const u32* header;            // this GPU's inbox header
    int nparts;
    int buckets;                  // B
};

// plan layout (u32 words): [0 .. E] tile prefix, [kInboxMaxEntries + 1 + e] start of entry e inside its region, [2 * kInboxMaxEntries + 1 + e] count
__global__ void __launch_bounds__(kInboxMaxEntries)
k_inbox_plan(InboxArgs a, u32* __restrict__ plan)
{
    __shared__ u32 s_w[kInboxMaxEntries / 32];
    const u32 e = threadIdx.x, lane = e & 31u, warp = e >> 5;
    const u32 E = (u32)(a.buckets * a.nparts);
    u32 tiles = 0;
    if (e < E) {
        const u32 j = e / (u32)a.nparts, src = e % (u32)a.nparts;
        const u32* offs = a.header + kInboxCountWords + src * kInboxOffsStride;
        const u32 lo = offs[j], hi = offs[j + 1u];
        plan[kInboxMaxEntries + 1 + e] = lo;
        plan[2 * kInboxMaxEntries + 1 + e] = hi - lo;
        tiles = (u32)div_up_u64(hi - lo, SP_THREADS * SP_IPT);
    }
    const u32 incl = warp_incl_scan_u32(tiles);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u32 prefix = 0;
    for (u32 w = 0; w < warp; ++w) prefix += s_w[w];
    if (e < E) plan[e] = prefix + incl - tiles;
    if (e == E - 1u) plan[E] = prefix + incl;
}

__global__ void __launch_bounds__(SP_THREADS)
k_scatter_plan(InboxArgs a, const u32* __restrict__ plan, u32* __restrict__ rank)
{
    const u32 E = (u32)(a.buckets * a.nparts);
    if (blockIdx.x >= plan[E]) return;  // the grid is an upper bound of the tile count
    u32 lo = 0, hi = E;                 // the entry e with plan[e] <= blockIdx.x < plan[e + 1]
    while (hi - lo > 1u) {
        const u32 mid = (lo + hi) >> 1;
        if (plan[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const u32 e = lo, src = e % (u32)a.nparts;
    const u32 start = plan[kInboxMaxEntries + 1 + e], m = plan[2 * kInboxMaxEntries + 1 + e];
    const u32* __restrict__ key = a.keys[src] + start;
    const u32* __restrict__ val = a.vals[src] + start;
    const u32 tile = SP_THREADS * SP_IPT;
    const u32 base = (blockIdx.x - plan[e]) * tile + threadIdx.x;
    u32 k[SP_IPT], v[SP_IPT];
#pragma unroll
    for (int q = 0; q < SP_IPT; ++q) {
        const u32 j = base + (u32)q * SP_THREADS;
        if (j < m) { k[q] = ld_stream(key + j); v[q] = ld_stream(val + j); }
    }
#pragma unroll
    for (int q = 0; q < SP_IPT; ++q) {
        const u32 j = base + (u32)q * SP_THREADS;
        if (j < m) rank[k[q]] = v[q];
    }
}

// Sharded ISA in peer memory, write phase: the pairs arrive here sorted by the top 8 bits of the suffix index, hence
// routed by owner (a run of consecutive digits per owner); run d is copied into this GPU's region of owner d's inbox with consecutive threads on consecutive
// addresses, i.e. full 128-byte stores over NVLink (direct 4-byte stores of the ranks into the owners' ISA arrays
// measured 9.5 ms for 1.3e8 pairs on two GPUs — NVLink moves small scattered stores at a few G/s).
struct PeerSend {
    u32* keys[kMaxPeers];        // destination of run d: keys, values
    u32* vals[kMaxPeers];
    u32* count_slot[kMaxPeers];  // where owner d reads how many pairs this GPU sent
    u32* offs_slot[kMaxPeers];   // ... and the B + 1 bucket offsets of the run (relative to its start)
    int nparts;
    int me;                      // sending GPU: its copy loop starts with the run of owner me + 1 (see k_peer_send)
    int per_owner_log;           // the input is sorted by a 256-way digit; owner d holds digits [d << per_owner_log, (d+1) << per_owner_log)
};

__global__ void __launch_bounds__(256)
k_peer_send(const u32* __restrict__ key, const u32* __restrict__ val, u32 m, const u32* __restrict__ bins /*256 exclusive digit offsets*/,
            PeerSend ps)
{
    __shared__ u32 s_off[kMaxPeers + 1];
    if (threadIdx.x <= (u32)kMaxPeers) {
        const u32 first_digit = threadIdx.x << ps.per_owner_log;
        s_off[threadIdx.x] = (threadIdx.x < (u32)ps.nparts && first_digit < 256u) ? bins[first_digit] : m;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (threadIdx.x < (u32)ps.nparts) *ps.count_slot[threadIdx.x] = s_off[threadIdx.x + 1] - s_off[threadIdx.x];
        const u32 B = 1u << ps.per_owner_log;
        for (u32 t = threadIdx.x; t < (u32)ps.nparts * (B + 1u); t += blockDim.x) {
            const u32 d = t / (B + 1u), j = t % (B + 1u);
            const u32 digit = (d << ps.per_owner_log) + j;
            u32 off = digit < 256u ? bins[digit] : m;
            off = off < s_off[d + 1] ? off : s_off[d + 1];  // the last owner's run ends at m
            ps.offs_slot[d][j] = off - s_off[d];
        }
    }
    // every GPU walks its runs starting with its right-hand neighbour's, so at any moment the G senders store into
    // G different inboxes (all starting with owner 0 measured G-fold ingress contention on that GPU's links)
    const u32 rot = s_off[(ps.me + 1) % ps.nparts];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const u32 j = i + rot < m ? i + rot : i + rot - m;
        int d = 0;
#pragma unroll
        for (int g = 1; g < kMaxPeers; ++g) d += (g < ps.nparts && s_off[g] <= j) ? 1 : 0;
        const u32 dst = j - s_off[d];
        ps.keys[d][dst] = ld_stream(key + j);
        ps.vals[d][dst] = ld_stream(val + j);
    }
}

// ---------------------------------------------------------------------------------------------
// O(n) validator (role of validate_suffix_array, main.cpp:236-270).
__global__ void __launch_bounds__(256)
k_check_scatter(const i32* __restrict__ sa, u32 n, u32* __restrict__ isa, unsigned long long* __restrict__ bad)
{
    const u64 total = (u64)n + 1;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (u64)gridDim.x * blockDim.x) {
        const u32 v = (u32)sa[r];  // int32 and uint32 suffix arrays alike: a negative int32 entry is > n
        bool ok = v <= n;
        if (r == 0 && v != n) ok = false;
        if (ok) {
            const u32 old = atomicExch(&isa[v], (u32)r);
            if (old != 0xffffffffu) ok = false;
        }
        if (!ok) atomicAdd(bad, 1ull);
    }
}

__global__ void __launch_bounds__(256)
k_check_order(const u8* __restrict__ text, const i32* __restrict__ sa, u32 n, const u32* __restrict__ isa,
              unsigned long long* __restrict__ bad)
{
    // rows 1..n-1 compared with their successor
    for (u64 r = 1 + (u64)blockIdx.x * blockDim.x + threadIdx.x; r < (u64)n; r += (u64)gridDim.x * blockDim.x) {
        const u32 a = (u32)sa[r], b = (u32)sa[r + 1];
        if (a >= n || b >= n) { atomicAdd(bad, 1ull); continue; }
        const u8 ca = text[a], cb = text[b];
        bool ok = ca < cb || (ca == cb && isa[a + 1] < isa[b + 1]);
        if (!ok) atomicAdd(bad, 1ull);
    }
}

}  // namespace b200sa
