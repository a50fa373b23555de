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

__global__ void __launch_bounds__(PK_THREADS)
k_pack_keys(const u8* __restrict__ text, u32 n, const u8* __restrict__ code, int bits, int k, int len_bits,
            u64* __restrict__ keys)
{
    __shared__ u8 s_code[256];
    __shared__ __align__(16) u8 s_sym[PK_TILE + PK_HALO];
    __shared__ u64 s_out[PK_THREADS / 32][PK_IPT * 33];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    s_code[tid] = code[tid];
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
            if (aligned && g + 16 <= n) {
                const uint4 q = *(const uint4*)(text + g);
                w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
#pragma unroll
                for (int i = 0; i < 16; ++i) s_sym[v * 16 + i] = s_code[(w[i >> 2] >> (8 * (i & 3))) & 255u];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) s_sym[v * 16 + i] = (g + i < n) ? s_code[text[g + i]] : (u8)0;
            }
        }
        __syncthreads();
        // thread handles 16 consecutive positions with a sliding window
        const u32 p0 = tid * PK_IPT;
        u64 win = 0;
        for (int j = 0; j < k; ++j) win = (win << bits) | (u64)s_sym[p0 + j];
#pragma unroll
        for (int i = 0; i < PK_IPT; ++i) {
            const u32 gp = base + p0 + i;
            const u32 rem = gp < n ? n - gp : 0u;
            const u64 key = (win << len_bits) | (u64)(rem < (u32)k ? rem : (u32)k);
            s_out[warp][i * 33 + lane] = key;
            win = ((win << bits) | (u64)s_sym[p0 + i + k]) & sym_mask;
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

__global__ void __launch_bounds__(BK_THREADS)
k_build_keys(const u32* __restrict__ idx, const u32* __restrict__ gid, const u32* __restrict__ rank,
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
            const u64 p = (u64)i[q] + h;
            r[q] = rank[p < n ? p : n];
        }
#pragma unroll
        for (int q = 0; q < BK_IPT; ++q) {
            const u32 j = base + (u32)q * BK_THREADS;
            if (j < m) st_stream(keys + j, ((u64)g[q] << rank_bits) | (u64)r[q]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Re-ranking after a sort: a three-kernel segmented scan over the m sorted tuples.
//   head[j]   = key[j] != key[j-1]            (j == 0 is a head)
//   single[j] = head[j] && head[j+1]          (j == m-1: head[m] counts as true)
//   kept[j]   = !single[j]
// scanned quantities: #kept (compaction slot), #kept heads (new dense group id), max head slot
// (start of my group).  Block aggregates are combined by one small block; the apply kernel then
// recomputes the flags and produces all outputs in one sweep.
static const int RR_THREADS = 256;
static const int RR_IPT = 8;
static const int RR_TILE = RR_THREADS * RR_IPT;  // 2048 tuples per block

struct RerankFlags {
    u32 head;    // bit q: item q is a head
    u32 single;  // bit q: item q is a singleton group
};

// Thread owns items j0 .. j0+RR_IPT-1 (blocked); reads keys j0-1 .. j0+RR_IPT.
__device__ __forceinline__ RerankFlags rr_flags(const u64* __restrict__ keys, u32 m, u32 j0)
{
    RerankFlags f;
    f.head = 0;
    f.single = 0;
    if (j0 >= m) return f;
    u64 k[RR_IPT + 2];
#pragma unroll
    for (int q = 0; q < RR_IPT + 2; ++q) {
        const i64 j = (i64)j0 + q - 1;
        k[q] = (j >= 0 && j < (i64)m) ? keys[j] : 0ull;
    }
    u32 headx = 0;  // bit q: item j0+q is a head, for q in 0..RR_IPT (one past the end included)
#pragma unroll
    for (int q = 0; q <= RR_IPT; ++q) {
        const u32 j = j0 + (u32)q;
        const bool h = (j == 0) || (j >= m) || (k[q + 1] != k[q]);
        headx |= (h ? 1u : 0u) << q;
    }
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        if (j0 + (u32)q < m) {
            const u32 h = (headx >> q) & 1u, hn = (headx >> (q + 1)) & 1u;
            f.head |= h << q;
            f.single |= (h & hn) << q;
        }
    }
    return f;
}

// aggregates: agg_cnt[b] = (#kept heads << 32) | #kept ; agg_max[b] = 1 + max head slot in block (0 = none)
__global__ void __launch_bounds__(RR_THREADS)
k_rerank_reduce(const u64* __restrict__ keys, u32 m, u64* __restrict__ agg_cnt, u32* __restrict__ agg_max)
{
    __shared__ u64 s_cnt[RR_THREADS / 32];
    __shared__ u32 s_max[RR_THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 j0 = blockIdx.x * (u32)RR_TILE + tid * RR_IPT;
    const RerankFlags f = rr_flags(keys, m, j0);
    const u32 nvalid = j0 >= m ? 0u : min((u32)RR_IPT, m - j0);
    const u32 validmask = nvalid >= 32 ? ~0u : ((1u << nvalid) - 1u);
    const u32 kept = ~f.single & validmask;
    u64 cnt = (u64)__popc(kept) | ((u64)__popc(kept & f.head) << 32);
    u32 mx = f.head ? j0 + (31u - (u32)__clz((int)f.head)) + 1u : 0u;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        cnt += __shfl_xor_sync(B200SA_FULL_MASK, cnt, d);
        const u32 o = __shfl_xor_sync(B200SA_FULL_MASK, mx, d);
        mx = mx > o ? mx : o;
    }
    if (lane == 0) { s_cnt[warp] = cnt; s_max[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        u64 c = 0;
        u32 x = 0;
        for (int w = 0; w < RR_THREADS / 32; ++w) { c += s_cnt[w]; x = x > s_max[w] ? x : s_max[w]; }
        agg_cnt[blockIdx.x] = c;
        agg_max[blockIdx.x] = x;
    }
}

// Exclusive scan of the block aggregates by one block; also publishes the round totals
// info[0] = #kept (next m), info[1] = #kept heads (next group count).
static const int RS2_THREADS = 1024;

__global__ void __launch_bounds__(RS2_THREADS)
k_rerank_scan_blocks(u64* __restrict__ agg_cnt, u32* __restrict__ agg_max, u32 nblocks, u32* __restrict__ info)
{
    __shared__ u64 s_wc[RS2_THREADS / 32];
    __shared__ u32 s_wm[RS2_THREADS / 32];
    __shared__ u64 s_carry_c;
    __shared__ u32 s_carry_m;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) { s_carry_c = 0; s_carry_m = 0; }
    __syncthreads();
    for (u32 base = 0; base < nblocks; base += RS2_THREADS) {
        const u32 b = base + tid;
        const u64 c = b < nblocks ? agg_cnt[b] : 0ull;
        const u32 x = b < nblocks ? agg_max[b] : 0u;
        const u64 ic = warp_incl_scan_u64(c);
        const u32 ix = warp_incl_scan_max_u32(x);
        if (lane == 31) { s_wc[warp] = ic; s_wm[warp] = ix; }
        __syncthreads();
        u64 pc = s_carry_c;
        u32 pm = s_carry_m;
        for (u32 w = 0; w < warp; ++w) { pc += s_wc[w]; pm = pm > s_wm[w] ? pm : s_wm[w]; }
        // exclusive values: everything before element b
        const u64 ec = pc + ic - c;
        u32 ex = pm;
        {
            const u32 up = __shfl_up_sync(B200SA_FULL_MASK, ix, 1);
            if (lane > 0) ex = ex > up ? ex : up;
        }
        if (b < nblocks) { agg_cnt[b] = ec; agg_max[b] = ex; }
        __syncthreads();
        if (tid == RS2_THREADS - 1) {
            s_carry_c = pc + ic;
            s_carry_m = pm > ix ? pm : ix;
        }
        __syncthreads();
    }
    if (tid == 0) {
        info[0] = (u32)(s_carry_c & 0xffffffffull);
        info[1] = (u32)(s_carry_c >> 32);
    }
}

// Apply: new ranks into the ISA, final SA entries for singletons, compaction of the rest.
//   slot_in == nullptr  -> round 0: active slot j is global position j
__global__ void __launch_bounds__(RR_THREADS)
k_rerank_apply(const u64* __restrict__ keys, const u32* __restrict__ idx_in, const u32* __restrict__ slot_in, u32 m,
               const u64* __restrict__ agg_cnt, const u32* __restrict__ agg_max,
               u32* __restrict__ rank, i32* __restrict__ sa,
               u32* __restrict__ idx_out, u32* __restrict__ slot_out, u32* __restrict__ gid_out)
{
    __shared__ u64 s_wc[RR_THREADS / 32];
    __shared__ u32 s_wm[RR_THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 j0 = blockIdx.x * (u32)RR_TILE + tid * RR_IPT;
    const RerankFlags f = rr_flags(keys, m, j0);
    const u32 nvalid = j0 >= m ? 0u : min((u32)RR_IPT, m - j0);
    const u32 validmask = nvalid >= 32 ? ~0u : ((1u << nvalid) - 1u);
    const u32 kept = ~f.single & validmask;
    const u64 cnt = (u64)__popc(kept) | ((u64)__popc(kept & f.head) << 32);
    const u32 mx = f.head ? j0 + (31u - (u32)__clz((int)f.head)) + 1u : 0u;

    const u64 ic = warp_incl_scan_u64(cnt);
    const u32 ix = warp_incl_scan_max_u32(mx);
    if (lane == 31) { s_wc[warp] = ic; s_wm[warp] = ix; }
    __syncthreads();
    u64 pc = agg_cnt[blockIdx.x];
    u32 pm = agg_max[blockIdx.x];
    for (u32 w = 0; w < warp; ++w) { pc += s_wc[w]; pm = pm > s_wm[w] ? pm : s_wm[w]; }
    u64 run_c = pc + ic - cnt;  // exclusive (#kept, #kept heads) before my first item
    u32 run_m = pm;             // 1 + head slot of the group open before my first item
    {
        const u32 up = __shfl_up_sync(B200SA_FULL_MASK, ix, 1);
        if (lane > 0) run_m = run_m > up ? run_m : up;
    }
    if (nvalid == 0) return;

    u32 dest = (u32)(run_c & 0xffffffffull);
    u32 heads = (u32)(run_c >> 32);
#pragma unroll
    for (int q = 0; q < RR_IPT; ++q) {
        if ((u32)q < nvalid) {
            const u32 j = j0 + (u32)q;
            const u32 is_head = (f.head >> q) & 1u, is_single = (f.single >> q) & 1u;
            if (is_head) run_m = j + 1u;
            const u32 hs = run_m - 1u;  // head slot of my group (always defined: slot 0 is a head)
            const u32 gpos = slot_in ? slot_in[hs] : hs;
            const u32 sfx = idx_in[j];
            rank[sfx] = gpos + 1u;
            if (is_single) {
                sa[gpos + 1u] = (i32)sfx;
            } else {
                heads += is_head;
                idx_out[dest] = sfx;
                slot_out[dest] = slot_in ? slot_in[j] : j;
                gid_out[dest] = heads - 1u;
                ++dest;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// O(n) validator (role of validate_suffix_array, main.cpp:236-270).
__global__ void __launch_bounds__(256)
k_check_scatter(const i32* __restrict__ sa, u32 n, u32* __restrict__ isa, unsigned long long* __restrict__ bad)
{
    const u64 total = (u64)n + 1;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (u64)gridDim.x * blockDim.x) {
        const i64 v = sa[r];
        bool ok = v >= 0 && v <= (i64)n;
        if (r == 0 && v != (i64)n) ok = false;
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
