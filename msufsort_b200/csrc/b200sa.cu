// b200sa.cu — engine implementation and the C ABI declared in include/b200sa.h.
//
// Build (product):  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//                        -Xcompiler -fPIC -o msufsort_b200/lib/libb200sa.so b200sa.cu
// There is no CPU path in this file: every entry point needs a CUDA device.
#include "engine.cuh"

#include <mutex>
#include <new>

namespace b200sa {

// ---------------------------------------------------------------------------------------------
// errors

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---------------------------------------------------------------------------------------------
// buffers

int DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    // round up so that slowly growing inputs do not reallocate every call
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return set_error(B200SA_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return 0;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ---------------------------------------------------------------------------------------------
// alphabet plan: dense symbol codes; as many symbols as fit next to the clamped-length field

AlphabetPlan plan_alphabet(const u32* hist, int reserved_bits, int max_key_bits, bool allow_radix)
{
    if (max_key_bits < 16 || max_key_bits > 64) max_key_bits = 64;
    AlphabetPlan a;
    int sigma = 0;
    for (int c = 0; c < 256; ++c) {
        a.code[c] = (u8)(sigma > 255 ? 255 : sigma);
        if (hist[c]) ++sigma;
    }
    // bytes that do not occur keep the code of the next larger occurring byte; never looked up
    if (sigma < 1) sigma = 1;
    a.sigma = sigma;
    int bits = 1;
    while ((1 << bits) < sigma) ++bits;
    a.bits = bits;
    int k = 1;
    for (int cand = 1; cand <= 58; ++cand)
        if (cand * bits + bit_length_u64((u64)cand) <= max_key_bits - reserved_bits) k = cand;
    a.k = k;
    a.len_bits = bit_length_u64((u64)k);
    a.radix = 0;
    a.key_bits = a.k * a.bits + a.len_bits;
    memset(a.pow, 0, sizeof(a.pow));
    if (allow_radix && sigma <= 255) {
        // the most digits of base sigma + 1 whose number fits the same bit budget
        const int budget = max_key_bits - reserved_bits;
        const unsigned __int128 limit = (unsigned __int128)1 << budget;
        const u64 B = (u64)sigma + 1;
        unsigned __int128 p = 1;
        int kr = 0;
        u64 pw[66] = {1};
        while (kr < 64 && p * B <= limit) { p *= B; ++kr; pw[kr] = (u64)p; }  // p == 2^64 cannot occur: B is not a power of two when it matters
        if (kr > a.k && kr <= 58 && p - 1 <= (unsigned __int128)~0ull) {
            a.radix = B;
            a.k = kr;
            a.len_bits = 0;
            a.bits = 0;
            a.key_bits = bit_length_u64((u64)(p - 1));
            memcpy(a.pow, pw, sizeof(pw));
        }
    }
    return a;
}

// ---------------------------------------------------------------------------------------------
// engine lifetime

int Engine::init(int dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return set_error(B200SA_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (dev < 0 || dev >= count) return set_error(B200SA_EINVAL, "device %d out of range (0..%d)", dev, count - 1);
    device = dev;
    B200SA_CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    B200SA_CU(cudaGetDeviceProperties(&prop, dev));
    num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : kNumSMs;
    B200SA_CU(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    B200SA_CU(cudaHostAlloc((void**)&h_pinned, 512 * sizeof(u32), cudaHostAllocDefault));
    memset(&prof, 0, sizeof(prof));
    // tuning knobs (tests lower them to drive the bucketed ISA update at small n)
    if (const char* e1 = getenv("B200SA_ISA_DIRECT_BYTES")) isa_direct_bytes = (size_t)strtoull(e1, nullptr, 10);
    if (const char* e2 = getenv("B200SA_ISA_MIN_UPDATES")) isa_min_updates = (u32)strtoul(e2, nullptr, 10);
    if (const char* e3 = getenv("B200SA_GROUPSORT_AVG")) groupsort_max_avg = (u32)strtoul(e3, nullptr, 10);
    if (const char* e4 = getenv("B200SA_GROUPSORT_TINY")) groupsort_tiny = (u32)strtoul(e4, nullptr, 10);
    if (const char* e5 = getenv("B200SA_GROUPSORT_MEDIUM")) groupsort_medium = (u32)strtoul(e5, nullptr, 10);
    if (const char* e8 = getenv("B200SA_MAX_KEY_BITS")) max_key_bits = atoi(e8);
    if (const char* e9 = getenv("B200SA_PACK_RADIX")) pack_radix = atoi(e9) != 0;
    if (const char* e10 = getenv("B200SA_RS_PERSISTENT")) rs_persistent = atoi(e10) != 0;
    if (const char* e12 = getenv("B200SA_LCP_DIRECT")) lcp_direct = atoi(e12) != 0;
    if (const char* e11 = getenv("B200SA_NUM_SMS")) { const int v = atoi(e11); if (v >= 1 && v <= 1024) num_sms = v; }  // tests: small persistent grids
    if (const char* e6 = getenv("B200SA_UNBWT_CAP_MULT")) unbwt_cap_mult = (u32)strtoul(e6, nullptr, 10);
    if (unbwt_cap_mult < 1) unbwt_cap_mult = 1;
    if (groupsort_tiny > (u32)GS_TINY) groupsort_tiny = GS_TINY;
    if (groupsort_medium > (u32)GS_MEDIUM) groupsort_medium = GS_MEDIUM;
    // the scatter kernels use more than the default 48 KB of dynamic shared memory
    {
        auto k64 = k_onesweep_pass<u64, true>;
        auto k64p = k_onesweep_pass_persistent<u64, true>;
        B200SA_CU(cudaFuncSetAttribute(k64p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u64>()));
        auto k8 = k_onesweep_pass<u8, false>;
        auto k32 = k_onesweep_pass<u32, true>;
        B200SA_CU(cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u32>()));
        B200SA_CU(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u64>()));
        B200SA_CU(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u8>()));
    }
    return 0;
}

int Engine::release_workspace()
{
    if (peer.active) peer_detach();  // the ISA array and the inbox go away: peers must re-attach before the next sharded sort
    for (int i = 0; i < 2; ++i) { keys[i].release(); idx[i].release(); slot[i].release(); }
    gid.release(); gstart.release(); glist.release(); rank.release(); sa_ws.release(); sortmeta.release(); agg_cnt.release(); agg_max.release();
    misc.release(); text_ws.release(); bwt_ws.release(); walk.release();
    batch_text.release(); batch_meta.release(); batch_out.release();
    peer_inbox.release();
    return 0;
}

void Engine::shutdown()
{
    cudaSetDevice(device);
    peer_detach();
    release_workspace();
    for (auto& s : spans) { event_pool.push_back(s.a); event_pool.push_back(s.b); }
    spans.clear();
    for (auto ev : event_pool) cudaEventDestroy(ev);
    event_pool.clear();
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    if (own_stream) cudaStreamDestroy(own_stream);
    own_stream = nullptr;
}

// ---------------------------------------------------------------------------------------------
// profiling: CUDA events on the launching stream around each phase

int Engine::phase_begin(int phase, cudaStream_t st)
{
    open_phase = phase;
    if (!profiling) return 0;
    cudaEvent_t a;
    if (!event_pool.empty()) { a = event_pool.back(); event_pool.pop_back(); }
    else B200SA_CU(cudaEventCreate(&a));
    B200SA_CU(cudaEventRecord(a, st));
    open_event = a;
    return 0;
}

int Engine::phase_end(cudaStream_t st)
{
    if (!profiling) { open_phase = -1; return 0; }
    cudaEvent_t b;
    if (!event_pool.empty()) { b = event_pool.back(); event_pool.pop_back(); }
    else B200SA_CU(cudaEventCreate(&b));
    B200SA_CU(cudaEventRecord(b, st));
    spans.push_back(Span{open_event, b, open_phase});
    open_phase = -1;
    open_event = nullptr;
    return 0;
}

int Engine::collect_profile()
{
    for (auto& s : spans) {
        B200SA_CU(cudaEventSynchronize(s.b));
        float ms = 0.f;
        B200SA_CU(cudaEventElapsedTime(&ms, s.a, s.b));
        prof.ms[s.phase] += (double)ms;
        event_pool.push_back(s.a);
        event_pool.push_back(s.b);
    }
    spans.clear();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// radix sort driver

// sortmeta layout: [ghist u32 8*256][tile counters u32 16][status u64 passes*tiles*256]
static const size_t kSortMetaHeader = (size_t)RS_MAX_PASSES * RS_RADIX * 4 + 16 * 4;

int Engine::radix_sort_pairs(u64* keys2[2], u32* vals2[2], bool gen_vals, u32 m, int begin_bit, int end_bit,
                             int* result_side, cudaStream_t st)
{
    *result_side = 0;
    if (m == 0 || end_bit <= begin_bit) return 0;
    const int passes = (end_bit - begin_bit + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
    if (passes > RS_MAX_PASSES) return set_error(B200SA_EINTERNAL, "radix sort over %d bits needs more than %d passes", end_bit - begin_bit, RS_MAX_PASSES);
    const u32 tiles = (u32)div_up_u64(m, RS_TILE);
    const size_t status_bytes = (size_t)passes * tiles * RS_RADIX * sizeof(u64);
    B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
    u32* ghist = sortmeta.as<u32>();
    u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
    u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
    B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
    prof.memsets++;

    B200SA_TRY(phase_begin(B200SA_PH_SORT_HIST, st));
    {
        const u32 htiles = (u32)div_up_u64(m, RH_THREADS * RH_IPT);
        const u32 grid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
        auto kh = k_radix_hist<u64>;
        B200SA_LAUNCH(kh, grid, RH_THREADS, rh_smem_bytes(passes), st, keys2[0], m, begin_bit, passes, ghist);
        count_launch(B200SA_PH_SORT_HIST);
        B200SA_LAUNCH(k_radix_scan_bins, passes, RS_RADIX, 0, st, ghist);
        count_launch(B200SA_PH_SORT_HIST);
        prof.alg_bytes[B200SA_PH_SORT_HIST] += (u64)m * 8;
    }
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());

    int side = 0;
    for (int p = 0; p < passes; ++p) {
        B200SA_TRY(phase_begin(B200SA_PH_SORT_PASS, st));
        const u32* vin = (p == 0 && gen_vals) ? nullptr : vals2[side];
        if (rs_persistent) {
            auto kp = k_onesweep_pass_persistent<u64, true>;
            const u32 grid = tiles < (u32)(num_sms * RS_MIN_BLOCKS) ? tiles : (u32)(num_sms * RS_MIN_BLOCKS);
            B200SA_LAUNCH(kp, grid, RS_THREADS, rs_pass_smem_bytes<u64>(), st,
                          keys2[side], keys2[side ^ 1], vin, vals2[side ^ 1], m, begin_bit + p * RS_RADIX_BITS, 0xffffffffu,
                          ghist + p * RS_RADIX, status + (size_t)p * tiles * RS_RADIX, counters + p, tiles);
        } else {
            auto kp = k_onesweep_pass<u64, true>;
            B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u64>(), st,
                          keys2[side], keys2[side ^ 1], vin, vals2[side ^ 1], m, begin_bit + p * RS_RADIX_BITS, 0xffffffffu,
                          ghist + p * RS_RADIX, status + (size_t)p * tiles * RS_RADIX, counters + p);
        }
        count_launch(B200SA_PH_SORT_PASS);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_SORT_PASS] += (u64)m * (vin ? 24 : 20);
        prof.sort_passes++;
        prof.sorted_tuples += m;
        side ^= 1;
    }
    B200SA_CU(cudaGetLastError());
    *result_side = side;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// rerank driver

// ISA update rank[idx[j]] = val[j] for `count` pairs.  Large arrays go through one radix sweep on the
// top 8 bits of the suffix index so that the scatter proper works inside an L2-resident window.
int Engine::isa_update(const u32* d_idx, const u32* d_val, u32 count, u32 n, u32* bk_key, u32* bk_val, bool all_suffixes, cudaStream_t st,
                       u32* target)
{
    if (count == 0) return 0;
    if (!target) target = rank.as<u32>();
    const bool bucketed = ((u64)n * 4 > isa_direct_bytes) && (count >= isa_min_updates) && bk_key && bk_val;
    B200SA_TRY(phase_begin(B200SA_PH_ISA, st));
    if (bucketed) {
        const int nbits = bit_length_u64((u64)n - 1);
        const int shift = nbits > RS_RADIX_BITS ? nbits - RS_RADIX_BITS : 0;
        const u32 tiles = (u32)div_up_u64(count, RS_TILE);
        const size_t status_bytes = (size_t)tiles * RS_RADIX * sizeof(u64);
        B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
        u32* ghist = sortmeta.as<u32>();
        u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
        u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
        B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
        prof.memsets++;
        if (all_suffixes && count == n) {
            // every suffix 0..n-1 appears exactly once (round 0 on one GPU): bucket b starts at min(b << shift, n)
            u32* h_bins = h_pinned + 64;
            for (u32 b = 0; b < 256; ++b) {
                const u64 lo = (u64)b << shift;
                h_bins[b] = (u32)(lo < n ? lo : n);
            }
            B200SA_CU(cudaMemcpyAsync(ghist, h_bins, 256 * 4, cudaMemcpyHostToDevice, st));
        } else {
            const u32 htiles = (u32)div_up_u64(count, RH_THREADS * RH_IPT);
            const u32 hgrid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
            auto kh = k_radix_hist<u32>;
            B200SA_LAUNCH(kh, hgrid, RH_THREADS, rh_smem_bytes(1), st, d_idx, count, shift, 1, ghist);
            count_launch(B200SA_PH_ISA);
            B200SA_LAUNCH(k_radix_scan_bins, 1, RS_RADIX, 0, st, ghist);
            count_launch(B200SA_PH_ISA);
        }
        auto kp = k_onesweep_pass<u32, true>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u32>(), st, d_idx, bk_key, d_val, bk_val,
                      count, shift, 0xffffffffu, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_ISA);
        B200SA_LAUNCH(k_scatter_pairs, (u32)div_up_u64(count, SP_THREADS * SP_IPT), SP_THREADS, 0, st, (const u32*)bk_key,
                      (const u32*)bk_val, count, target);
        count_launch(B200SA_PH_ISA);
        prof.alg_bytes[B200SA_PH_ISA] += (u64)count * (4 + 8 + 8 + 8 + 4);
    } else {
        B200SA_LAUNCH(k_scatter_pairs, (u32)div_up_u64(count, SP_THREADS * SP_IPT), SP_THREADS, 0, st, d_idx, d_val, count, target);
        count_launch(B200SA_PH_ISA);
        prof.alg_bytes[B200SA_PH_ISA] += (u64)count * 12;
    }
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    return 0;
}

// Rerank driver.  mode: 0 = the kernel scatters new ranks into rank[] itself (small ISA);
// 1 = new ranks are written in slot order and applied here through isa_update(); 2 = written in slot
// order and left for the caller (sharded runs exchange them between GPUs first).
int Engine::rerank(const u64* keys_sorted, const u32* idx_sorted, const u32* slot_in, u32 slot_base, u32 m, u32 n, i32* d_sa,
                   u32* idx_out, u32* slot_out, u64* free_keys, int mode, u32* next_m, u32* next_groups, cudaStream_t st)
{
    *next_m = 0;
    *next_groups = 0;
    if (m == 0) return 0;
    const u32 ntiles = (u32)div_up_u64(m, RR_TILE);
    // agg_cnt holds the two descriptor arrays followed by the tile ticket counter
    const size_t desc_bytes = (size_t)ntiles * 2 * sizeof(u64);
    B200SA_TRY(agg_cnt.ensure(desc_bytes + 64));
    u64* desc = agg_cnt.as<u64>();
    u32* ticket = (u32*)((u8*)agg_cnt.p + desc_bytes);
    u32* d_info = misc.as<u32>() + 512;  // see ensure_sa_workspace for the misc layout
    B200SA_CU(cudaMemsetAsync(agg_cnt.p, 0, desc_bytes + 64, st));
    prof.memsets++;
    if (mode == 1 && !(((u64)n * 4 > isa_direct_bytes) && (m >= isa_min_updates))) mode = 0;
    u32* newrank = mode ? (u32*)free_keys : nullptr;  // [m]
    B200SA_TRY(phase_begin(B200SA_PH_RERANK, st));
    B200SA_LAUNCH(k_rerank, ntiles, RR_THREADS, 0, st, keys_sorted, idx_sorted, slot_in, slot_base, m, desc, ntiles, ticket,
                  rank.as<u32>(), newrank, d_sa, idx_out, slot_out, gid.as<u32>(), gstart.as<u32>(), d_info);
    count_launch(B200SA_PH_RERANK);
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned, d_info, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    if (mode == 1) {
        B200SA_TRY(agg_max.ensure((size_t)m * 4 + 64));
        B200SA_TRY(isa_update(idx_sorted, newrank, m, n, agg_max.as<u32>(), (u32*)free_keys + m, slot_in == nullptr && m == n, st));
    }
    B200SA_CU(cudaStreamSynchronize(st));
    *next_m = h_pinned[0];
    *next_groups = h_pinned[1];
    prof.alg_bytes[B200SA_PH_RERANK] += (u64)m * (8 + 4 + 4 + 4) + (u64)(*next_m) * 12 + (u64)(m - *next_m) * 4;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// suffix array

int Engine::ensure_sa_workspace(u64 n)
{
    for (int i = 0; i < 2; ++i) {
        B200SA_TRY(keys[i].ensure((size_t)n * 8 + 64));
        B200SA_TRY(idx[i].ensure((size_t)n * 4 + 64));
        B200SA_TRY(slot[i].ensure((size_t)n * 4 + 64));
    }
    B200SA_TRY(gid.ensure((size_t)n * 4 + 64));
    B200SA_TRY(gstart.ensure(((size_t)n / 2 + 2) * 4 + 64));  // a kept group has >= 2 members
    B200SA_TRY(rank.ensure(((size_t)n + 1) * 4 + 64));
    // misc: [0..255] byte histogram u32, [256..319] symbol codes (256 bytes), [512..] round info,
    // [520..521] validator counter (u64), [528] sentinel row
    B200SA_TRY(misc.ensure(4096));
    return 0;
}

// The sort is a small state machine so that the single-GPU entry point and the sharded (multi-GPU)
// driver share every step: begin (alphabet, keys, optional key-range filter, first sort) ->
// round 0 ranking -> doubling rounds.  With nparts > 1 a context sorts only the suffixes whose
// initial key falls into its splitter range; those form whole groups, so every later sort is local
// and only ISA updates travel between GPUs (msufsort_b200/sharded.py).
int Engine::sort_begin(const u8* d_text, u32 n, i32* d_sa, int part, int nparts, u32* n_local, cudaStream_t st)
{
    B200SA_TRY(ensure_sa_workspace(n));
    ss = SortState();
    ss.d_text = d_text; ss.n = n; ss.d_sa = d_sa; ss.part = part; ss.nparts = nparts;
    if (next_batch.count) {
        if (nparts != 1) return set_error(B200SA_EINVAL, "a batched sort cannot be sharded");
        ss.batch_ends = next_batch.d_ends;
        ss.batch_count = next_batch.count;
        ss.batch_bits = bit_length_u64((u64)next_batch.count - 1);
        next_batch = BatchDesc();
    }
    u32* d_hist = misc.as<u32>();
    u8* d_code = (u8*)(misc.as<u32>() + 256);

    // ---- alphabet
    B200SA_TRY(phase_begin(B200SA_PH_ALPHABET, st));
    B200SA_CU(cudaMemsetAsync(d_hist, 0, 256 * sizeof(u32), st));
    prof.memsets++;
    {
        const u64 nvec = (u64)n / 16 + 1;
        const u32 grid = (u32)(div_up_u64(nvec, BH_THREADS) < (u64)(num_sms * 8) ? div_up_u64(nvec, BH_THREADS) : (u64)(num_sms * 8));
        B200SA_LAUNCH(k_byte_hist, grid, BH_THREADS, 0, st, d_text, (u64)n, d_hist);
        count_launch(B200SA_PH_ALPHABET);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_ALPHABET] += n;
    u32 h_hist[256];
    B200SA_CU(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    if (ss.batch_count) h_hist[0] -= ss.batch_count;  // the separator slots of a batch are not symbols
    ss.plan = plan_alphabet(h_hist, ss.batch_bits, max_key_bits, pack_radix);
    const AlphabetPlan& plan = ss.plan;
    B200SA_CU(cudaMemcpyAsync(d_code, plan.code, 256, cudaMemcpyHostToDevice, st));

    // ---- initial keys for every suffix
    B200SA_LAUNCH(k_sa_init, 1, 32, 0, st, rank.as<u32>(), d_sa, n);
    count_launch(B200SA_PH_PACK);
    B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
    {
        const u32 tiles = (u32)div_up_u64(n, PK_TILE);
        const u32 grid = tiles < (u32)(num_sms * 4) ? tiles : (u32)(num_sms * 4);
        if (plan.radix) {
            u64* d_pow = (u64*)(misc.as<u32>() + 320);  // 66 words of u64 behind the symbol codes
            B200SA_CU(cudaMemcpyAsync(d_pow, plan.pow, sizeof(plan.pow), cudaMemcpyHostToDevice, st));
            if (ss.batch_count) {
                auto kp = k_pack_keys_batch<true>;
                B200SA_LAUNCH(kp, grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.key_bits, plan.k, 0, plan.radix, (const u64*)d_pow,
                              ss.batch_ends, ss.batch_count, keys[0].as<u64>());
            } else {
                auto kp = k_pack_keys<true>;
                B200SA_LAUNCH(kp, grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, 0, plan.k, 0, plan.radix, plan.pow[plan.k - 1], keys[0].as<u64>());
            }
        } else if (ss.batch_count) {
            auto kp = k_pack_keys_batch<false>;
            B200SA_LAUNCH(kp, grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, (const u64*)nullptr,
                          ss.batch_ends, ss.batch_count, keys[0].as<u64>());
        } else {
            auto kp = k_pack_keys<false>;
            B200SA_LAUNCH(kp, grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, (u64)0, keys[0].as<u64>());
        }
        count_launch(B200SA_PH_PACK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_PACK] += (u64)n * 9;
    B200SA_CU(cudaGetLastError());

    u64* k2[2] = {keys[0].as<u64>(), keys[1].as<u64>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    const int key_bits = plan.key_bits + ss.batch_bits;
    int side = 0;
    u32 count = n;
    if (nparts > 1) {
        // ---- splitters from a sorted sample of the keys (same text + same kernels on every GPU =>
        // identical splitters everywhere, no communication), then keep only this part's key range
        u32 nsample = n < (1u << 18) ? n : (1u << 18);
        const u32 stride = n / nsample;
        u64* smp[2] = {(u64*)slot[0].p, (u64*)slot[1].p};  // the slot maps are not in use yet (8 * 2^18 bytes each fits: n >= nsample)
        B200SA_TRY(slot[0].ensure((size_t)nsample * 8 + 64));
        B200SA_TRY(slot[1].ensure((size_t)nsample * 8 + 64));
        smp[0] = (u64*)slot[0].p; smp[1] = (u64*)slot[1].p;
        B200SA_TRY(agg_max.ensure((size_t)nsample * 8 + 64));
        u32* sv[2] = {agg_max.as<u32>(), agg_max.as<u32>() + nsample};
        B200SA_LAUNCH(k_sample_keys, (u32)div_up_u64(nsample, 256), 256, 0, st, (const u64*)k2[0], nsample, stride, smp[0]);
        count_launch(B200SA_PH_PACK);
        int sside = 0;
        B200SA_TRY(radix_sort_pairs(smp, sv, true, nsample, 0, key_bits, &sside, st));
        std::vector<u64> h_sample(nsample);
        B200SA_CU(cudaMemcpyAsync(h_sample.data(), smp[sside], (size_t)nsample * 8, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        const u64 lo = part == 0 ? 0ull : h_sample[(size_t)((u64)part * nsample / nparts)];
        const u64 hi = part == nparts - 1 ? ~0ull : h_sample[(size_t)((u64)(part + 1) * nsample / nparts)];
        const bool hi_inclusive = part == nparts - 1;
        u32* d_count = misc.as<u32>() + 536;
        B200SA_CU(cudaMemsetAsync(d_count, 0, 4, st));
        prof.memsets++;
        B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
        B200SA_LAUNCH(k_filter_range, (u32)div_up_u64(n, FR_THREADS * FR_IPT), FR_THREADS, 0, st, (const u64*)k2[0], n, lo, hi,
                      hi_inclusive ? 1 : 0, k2[1], v2[1], d_count);
        count_launch(B200SA_PH_PACK);
        B200SA_TRY(phase_end(st));
        B200SA_CU(cudaMemcpyAsync(h_pinned + 16, d_count, 4, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        count = h_pinned[16];
        prof.alg_bytes[B200SA_PH_PACK] += (u64)n * 8 + (u64)count * 12;
        // sort this part: input on side 1
        u64* kk[2] = {k2[1], k2[0]};
        u32* vv[2] = {v2[1], v2[0]};
        int rs = 0;
        B200SA_TRY(radix_sort_pairs(kk, vv, false, count, 0, key_bits, &rs, st));
        side = 1 ^ rs;
    } else {
        B200SA_TRY(radix_sort_pairs(k2, v2, true, n, 0, key_bits, &side, st));
    }
    prof.rounds++;
    prof.active_tuples += count;
    ss.n_local = count;
    ss.sorted_side = side;
    ss.stage = 1;
    *n_local = count;
    return 0;
}

int Engine::sort_round0(u32 slot_base, u32* m_local, cudaStream_t st)
{
    if (ss.stage != 1) return set_error(B200SA_EINVAL, "sort_round0 called out of order");
    u64* k2[2] = {keys[0].as<u64>(), keys[1].as<u64>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    const int side = ss.sorted_side;
    ss.cur_slot = 0;
    // sorted tuples are on `side`; the compacted active array goes to the other side
    B200SA_TRY(rerank(k2[side], v2[side], nullptr, slot_base, ss.n_local, ss.n, ss.d_sa, v2[side ^ 1], slot[0].as<u32>(),
                      k2[side ^ 1], ss.nparts > 1 ? 2 : 1, &ss.m, &ss.groups, st));
    ss.upd_idx = v2[side];
    ss.upd_rank = (const u32*)k2[side ^ 1];
    ss.upd_count = ss.nparts > 1 ? ss.n_local : 0;
    ss.act = side ^ 1;
    ss.rank_bits = bit_length_u64(ss.n);
    ss.h = (u64)ss.plan.k;
    ss.stage = 2;
    ss.guard = 0;
    *m_local = ss.m;
    return 0;
}

int Engine::sort_round(u32* m_local, cudaStream_t st)
{
    if (ss.stage != 2) return set_error(B200SA_EINVAL, "sort_round called out of order");
    ss.upd_count = 0;
    if (ss.m == 0) { *m_local = 0; ss.h *= 2; return 0; }  // nothing left here; peers may still be working
    const u32 m = ss.m, n = ss.n;
    if (++ss.guard > 64) return set_error(B200SA_EINTERNAL, "prefix doubling did not converge (m=%u, h=%llu)", m, (unsigned long long)ss.h);
    if (ss.groups == 0 || ss.h > (u64)n) return set_error(B200SA_EINTERNAL, "inconsistent round state (m=%u groups=%u h=%llu)", m, ss.groups, (unsigned long long)ss.h);
    u64* k2[2] = {keys[0].as<u64>(), keys[1].as<u64>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    const int act = ss.act;
    const int gid_bits = ss.groups <= 1 ? 0 : bit_length_u64((u64)ss.groups - 1);
    int sorted_side = -1;
    // where rank[suffix + h] is read from: this GPU's array, or the shards of all GPUs through peer memory
    const bool use_peer = peer.active && ss.nparts > 1;
    if (use_peer && (peer.view.n != n || peer.nparts != ss.nparts)) return set_error(B200SA_EINVAL, "peer ISA attached for another text size / GPU count");
    LocalRank local_rank{rank.as<u32>(), n};
    PeerRank peer_rank{peer.view};

    // ---- small groups: sort every group where it lies (no radix sweeps)
    if (groupsort_max_avg > 0 && (u64)m <= (u64)ss.groups * groupsort_max_avg) {
        const u32 G = ss.groups;
        B200SA_TRY(glist.ensure(((size_t)G * 2 + 16) * 4));
        u32* medium_list = glist.as<u32>();
        u32* huge_list = glist.as<u32>() + G;
        u32* d_cnt = misc.as<u32>() + 540;  // 4 counters
        B200SA_CU(cudaMemsetAsync(d_cnt, 0, 16, st));
        prof.memsets++;
        B200SA_TRY(phase_begin(B200SA_PH_SEGSORT, st));
        if (use_peer) {
            auto kt = k_group_sort_tiny<PeerRank>;
            B200SA_LAUNCH(kt, (u32)div_up_u64(G, GS_THREADS), GS_THREADS, 0, st, (const u32*)gstart.as<u32>(), G, v2[act],
                          peer_rank, n, (u32)ss.h, ss.rank_bits, k2[act], groupsort_tiny, groupsort_medium, medium_list, huge_list, d_cnt);
        } else {
            auto kt = k_group_sort_tiny<LocalRank>;
            B200SA_LAUNCH(kt, (u32)div_up_u64(G, GS_THREADS), GS_THREADS, 0, st, (const u32*)gstart.as<u32>(), G, v2[act],
                          local_rank, n, (u32)ss.h, ss.rank_bits, k2[act], groupsort_tiny, groupsort_medium, medium_list, huge_list, d_cnt);
        }
        count_launch(B200SA_PH_SEGSORT);
        B200SA_CU(cudaMemcpyAsync(h_pinned + 20, d_cnt, 16, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        const u32 nmedium = h_pinned[20], nhuge = h_pinned[21];
        prof.alg_bytes[B200SA_PH_SEGSORT] += (u64)G * 4 + (u64)m * 20;
        if (nhuge <= 64) {
            if (nmedium) {
                if (use_peer) {
                    auto km = k_group_sort_medium<PeerRank>;
                    B200SA_LAUNCH(km, nmedium, GM_THREADS, 0, st, (const u32*)medium_list, (const u32*)gstart.as<u32>(), v2[act], peer_rank, n,
                                  (u32)ss.h, ss.rank_bits, k2[act]);
                } else {
                    auto km = k_group_sort_medium<LocalRank>;
                    B200SA_LAUNCH(km, nmedium, GM_THREADS, 0, st, (const u32*)medium_list, (const u32*)gstart.as<u32>(), v2[act], local_rank, n,
                                  (u32)ss.h, ss.rank_bits, k2[act]);
                }
                count_launch(B200SA_PH_SEGSORT);
            }
            B200SA_TRY(phase_end(st));
            B200SA_CU(cudaGetLastError());
            if (nhuge) {
                // the few groups too large for one CTA: radix-sort their slot ranges one by one on the second key
                std::vector<u32> hl(nhuge), gs((size_t)2 * nhuge);
                B200SA_CU(cudaMemcpyAsync(hl.data(), huge_list, (size_t)nhuge * 4, cudaMemcpyDeviceToHost, st));
                B200SA_CU(cudaStreamSynchronize(st));
                for (u32 q = 0; q < nhuge; ++q)
                    B200SA_CU(cudaMemcpyAsync(&gs[2 * q], gstart.as<u32>() + hl[q], 8, cudaMemcpyDeviceToHost, st));
                B200SA_CU(cudaStreamSynchronize(st));
                for (u32 q = 0; q < nhuge; ++q) {
                    const u32 s0 = gs[2 * q], sz = gs[2 * q + 1] - s0;
                    B200SA_TRY(phase_begin(B200SA_PH_BUILD, st));
                    const u32 tiles = (u32)div_up_u64(sz, BK_THREADS * BK_IPT);
                    const u32 grid = tiles < (u32)(num_sms * 8) ? tiles : (u32)(num_sms * 8);
                    if (use_peer) {
                        auto kb = k_build_keys<PeerRank>;
                        B200SA_LAUNCH(kb, grid, BK_THREADS, 0, st, (const u32*)(v2[act] + s0), (const u32*)(gid.as<u32>() + s0), peer_rank, sz, n,
                                      (u32)ss.h, ss.rank_bits, k2[act] + s0);
                    } else {
                        auto kb = k_build_keys<LocalRank>;
                        B200SA_LAUNCH(kb, grid, BK_THREADS, 0, st, (const u32*)(v2[act] + s0), (const u32*)(gid.as<u32>() + s0), local_rank, sz, n,
                                      (u32)ss.h, ss.rank_bits, k2[act] + s0);
                    }
                    count_launch(B200SA_PH_BUILD);
                    B200SA_TRY(phase_end(st));
                    u64* kk[2] = {k2[act] + s0, k2[act ^ 1] + s0};
                    u32* vv[2] = {v2[act] + s0, v2[act ^ 1] + s0};
                    int rs = 0;
                    B200SA_TRY(radix_sort_pairs(kk, vv, false, sz, 0, ss.rank_bits, &rs, st));
                    if (rs) {
                        B200SA_CU(cudaMemcpyAsync(kk[0], kk[1], (size_t)sz * 8, cudaMemcpyDeviceToDevice, st));
                        B200SA_CU(cudaMemcpyAsync(vv[0], vv[1], (size_t)sz * 4, cudaMemcpyDeviceToDevice, st));
                    }
                }
            }
            sorted_side = act;
        } else {
            // too many large groups for per-group launches: fall through to the radix path, which
            // rebuilds the keys (the in-place work done so far only permuted members inside groups)
            B200SA_TRY(phase_end(st));
        }
    }

    if (sorted_side < 0) {
        B200SA_TRY(phase_begin(B200SA_PH_BUILD, st));
        {
            const u32 tiles = (u32)div_up_u64(m, BK_THREADS * BK_IPT);
            const u32 grid = tiles < (u32)(num_sms * 8) ? tiles : (u32)(num_sms * 8);
            if (use_peer) {
                auto kb = k_build_keys<PeerRank>;
                B200SA_LAUNCH(kb, grid, BK_THREADS, 0, st, (const u32*)v2[act], (const u32*)gid.as<u32>(), peer_rank, m, n, (u32)ss.h,
                              ss.rank_bits, k2[act]);
            } else {
                auto kb = k_build_keys<LocalRank>;
                B200SA_LAUNCH(kb, grid, BK_THREADS, 0, st, (const u32*)v2[act], (const u32*)gid.as<u32>(), local_rank, m, n, (u32)ss.h,
                              ss.rank_bits, k2[act]);
            }
            count_launch(B200SA_PH_BUILD);
        }
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_BUILD] += (u64)m * 20;
        u64* kk[2] = {k2[act], k2[act ^ 1]};
        u32* vv[2] = {v2[act], v2[act ^ 1]};
        int rs = 0;
        B200SA_TRY(radix_sort_pairs(kk, vv, false, m, 0, ss.rank_bits + gid_bits, &rs, st));
        sorted_side = act ^ rs;
    }
    u32 m2 = 0, g2 = 0;
    B200SA_TRY(rerank(k2[sorted_side], v2[sorted_side], slot[ss.cur_slot].as<u32>(), 0, m, n, ss.d_sa, v2[sorted_side ^ 1],
                      slot[ss.cur_slot ^ 1].as<u32>(), k2[sorted_side ^ 1], ss.nparts > 1 ? 2 : 1, &m2, &g2, st));
    prof.rounds++;
    prof.active_tuples += m;
    ss.upd_idx = v2[sorted_side];
    ss.upd_rank = (const u32*)k2[sorted_side ^ 1];
    ss.upd_count = ss.nparts > 1 ? m : 0;
    ss.act = sorted_side ^ 1;
    ss.cur_slot ^= 1;
    ss.m = m2;
    ss.groups = g2;
    ss.h *= 2;
    *m_local = m2;
    return 0;
}

int Engine::suffix_array_dev(const u8* d_text, i64 n64, i32* d_sa, cudaStream_t st, i64 max_n)
{
    if (n64 < 0 || n64 > max_n) return set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n64, (long long)max_n);
    if (!d_sa || (n64 > 0 && !d_text)) return set_error(B200SA_EINVAL, "null pointer");
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64;
    if (n == 0) {
        B200SA_CU(cudaMemsetAsync(d_sa, 0, sizeof(i32), st));
        B200SA_CU(cudaStreamSynchronize(st));
        return 0;
    }
    u32 n_local = 0, m = 0;
    B200SA_TRY(sort_begin(d_text, n, d_sa, 0, 1, &n_local, st));
    B200SA_TRY(sort_round0(0, &m, st));
    while (m > 0) B200SA_TRY(sort_round(&m, st));
    ss.stage = 3;
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ISA sharded over the GPUs of one box, accessed through peer memory (CUDA IPC; msufsort_b200/sharded.py isa="peer").
// Two allocations per GPU are mapped by all peers: the ISA array (peers LOAD rank[suffix + h] from it) and an inbox
// (peers STORE the new ranks of the suffixes this GPU owns into it, in bulk; the owner then applies them locally).

static const size_t kInboxHeader = 256;  // u32 count[kMaxPeers] written by the sources, padded

int Engine::peer_export(u64 n, unsigned char* handles_out)
{
    if (n == 0 || n > (u64)B200SA_MAX_N_INT32 || !handles_out) return set_error(B200SA_EINVAL, "bad argument");
    B200SA_CU(cudaSetDevice(device));
    B200SA_TRY(ensure_sa_workspace(n));  // the ISA array keeps its address as long as n does not grow
    B200SA_TRY(peer_inbox.ensure(kInboxHeader + (size_t)n * 8 + 64));
    cudaIpcMemHandle_t h[2];
    static_assert(sizeof(h[0]) == 64, "IPC handle size");
    B200SA_CU(cudaIpcGetMemHandle(&h[0], rank.p));
    B200SA_CU(cudaIpcGetMemHandle(&h[1], peer_inbox.p));
    memcpy(handles_out, h, 128);
    return 0;
}

int Engine::peer_detach()
{
    for (auto& o : peer.opened) cudaIpcCloseMemHandle(o.second);
    peer.opened.clear();
    peer = PeerState();
    return 0;
}

int Engine::peer_attach(int part, int nparts, int shift, u64 n, const unsigned char* handles)
{
    if (nparts < 2 || nparts > kMaxPeers || part < 0 || part >= nparts || shift < 0 || shift > 31 || !handles || n == 0 ||
        n > (u64)B200SA_MAX_N_INT32)
        return set_error(B200SA_EINVAL, "bad argument (at most %d GPUs)", kMaxPeers);
    if (((n - 1) >> shift) >= (u64)nparts) return set_error(B200SA_EINVAL, "shift %d does not spread %llu positions over %d GPUs", shift, (unsigned long long)n, nparts);
    B200SA_CU(cudaSetDevice(device));
    B200SA_TRY(ensure_sa_workspace(n));
    B200SA_TRY(peer_inbox.ensure(kInboxHeader + (size_t)n * 8 + 64));
    // mappings of an earlier attach are reused when the peer still exports the same allocation
    std::vector<std::pair<std::string, void*>> keep;
    PeerState next;
    auto open_one = [&](const unsigned char* hb, void** out) -> int {
        const std::string key((const char*)hb, 64);
        void* ptr = nullptr;
        for (auto& o : peer.opened)
            if (o.second && o.first == key) { ptr = o.second; o.second = nullptr; break; }
        if (!ptr) {
            cudaIpcMemHandle_t h;
            memcpy(&h, hb, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return set_error(B200SA_ECOMM, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
            }
        }
        keep.emplace_back(key, ptr);
        *out = ptr;
        return 0;
    };
    int rc = 0;
    for (int g = 0; g < nparts && rc == 0; ++g) {
        if (g == part) { next.view.base[g] = rank.as<u32>(); next.inbox[g] = peer_inbox.as<u8>(); continue; }
        void *pr = nullptr, *pi = nullptr;
        rc = open_one(handles + (size_t)g * 128, &pr);
        if (rc == 0) rc = open_one(handles + (size_t)g * 128 + 64, &pi);
        next.view.base[g] = (u32*)pr;
        next.inbox[g] = (u8*)pi;
    }
    for (auto& o : peer.opened)
        if (o.second) cudaIpcCloseMemHandle(o.second);
    peer.opened.clear();
    if (rc != 0) {
        for (auto& k : keep) cudaIpcCloseMemHandle(k.second);
        peer = PeerState();
        return rc;
    }
    next.opened = keep;
    next.active = true;
    next.part = part;
    next.nparts = nparts;
    next.view.shift = shift;
    next.view.n = (u32)n;
    peer = next;
    return 0;
}

// inbox regions: source s stores its pairs at kInboxHeader + 8 * (suffixes owned by lower-numbered sources) of EVERY
// destination's inbox (keys first, then values); a source never sends more pairs than it owns suffixes
int Engine::peer_layout(const i64* counts, int nparts)
{
    if (!peer.active || nparts != peer.nparts || !counts) return set_error(B200SA_EINVAL, "no peer ISA attached for %d GPUs", nparts);
    u64 acc = 0;
    for (int g = 0; g < nparts; ++g) {
        if (counts[g] < 0) return set_error(B200SA_EINVAL, "negative count");
        peer.region_off[g] = kInboxHeader + 8 * acc;
        peer.region_cap[g] = (u32)counts[g];
        acc += (u64)counts[g];
    }
    if (acc != peer.view.n) return set_error(B200SA_EINVAL, "suffix counts of the parts add up to %llu, not n = %u", (unsigned long long)acc, peer.view.n);
    peer.laid_out = true;
    return 0;
}

// Write phase, part 1: route the (suffix, rank) pairs of the last round0 / round step by owner (one radix sweep) and
// store every owner's run into this GPU's region of that owner's inbox — coalesced 128-byte stores over NVLink.
int Engine::peer_scatter(cudaStream_t st)
{
    if (!peer.active || !peer.laid_out || ss.stage < 2) return set_error(B200SA_EINVAL, "no sharded sort with a peer ISA in progress");
    const u32 count = ss.upd_count;
    const int G = peer.nparts, me = peer.part;
    if (count > peer.region_cap[me]) return set_error(B200SA_EINTERNAL, "%u updates exceed this GPU's inbox region (%u)", count, peer.region_cap[me]);
    // ONE sweep on the top 8 bits of the suffix index both routes by owner (a bucket never straddles two owners:
    // bucket width 2^bshift divides the shard width 2^shift) and pre-buckets every run for the owner's scatter
    const int nbits = bit_length_u64((u64)ss.n - 1);
    int bshift = nbits > RS_RADIX_BITS ? nbits - RS_RADIX_BITS : 0;
    if (bshift > peer.view.shift) bshift = peer.view.shift;
    const int per_owner_log = peer.view.shift - bshift;  // buckets per owner = 2^per_owner_log
    PeerSend ps;
    for (int g = 0; g < kMaxPeers; ++g) {
        ps.keys[g] = g < G ? (u32*)(peer.inbox[g] + peer.region_off[me]) : nullptr;
        ps.vals[g] = g < G ? ps.keys[g] + peer.region_cap[me] : nullptr;
        ps.count_slot[g] = g < G ? (u32*)peer.inbox[g] + me : nullptr;
    }
    ps.nparts = G;
    ps.me = me;
    ps.per_owner_log = per_owner_log;
    const u32 tiles = (u32)div_up_u64(count ? count : 1, RS_TILE);
    const size_t status_bytes = (size_t)tiles * RS_RADIX * sizeof(u64);
    B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
    u32* ghist = sortmeta.as<u32>();
    B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
    prof.memsets++;
    B200SA_TRY(phase_begin(B200SA_PH_ISA, st));
    if (count) {
        B200SA_TRY(agg_max.ensure((size_t)count * 4 + 64));
        u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
        u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
        u32* bk_key = agg_max.as<u32>();
        u32* bk_val = (u32*)ss.upd_rank + count;  // second half of the key buffer the new ranks sit in
        const u32 htiles = (u32)div_up_u64(count, RH_THREADS * RH_IPT);
        const u32 hgrid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
        auto kh = k_radix_hist<u32>;
        B200SA_LAUNCH(kh, hgrid, RH_THREADS, rh_smem_bytes(1), st, ss.upd_idx, count, bshift, 1, ghist);
        count_launch(B200SA_PH_ISA);
        B200SA_LAUNCH(k_radix_scan_bins, 1, RS_RADIX, 0, st, ghist);
        count_launch(B200SA_PH_ISA);
        auto kp = k_onesweep_pass<u32, true>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u32>(), st, ss.upd_idx, bk_key, ss.upd_rank, bk_val, count, bshift,
                      0xffffffffu, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_ISA);
        const u32 want = (u32)div_up_u64(count, 256 * 4);
        const u32 grid = want < (u32)(num_sms * 16) ? want : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_peer_send, grid, 256, 0, st, (const u32*)bk_key, (const u32*)bk_val, count, (const u32*)ghist, ps);
        count_launch(B200SA_PH_ISA);
        prof.alg_bytes[B200SA_PH_ISA] += (u64)count * (4 + 16 + 16);
    } else {
        // nothing to send this round: the counts the owners read must still be reset (ghist is all zero)
        B200SA_LAUNCH(k_peer_send, 1, 256, 0, st, (const u32*)nullptr, (const u32*)nullptr, 0u, (const u32*)ghist, ps);
        count_launch(B200SA_PH_ISA);
    }
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));  // stores to peer memory have landed when the kernel has completed
    return 0;
}

// Write phase, part 2 (after a barrier): apply what the peers left in this GPU's inbox to its ISA shard.
int Engine::peer_apply(cudaStream_t st)
{
    if (!peer.active || !peer.laid_out) return set_error(B200SA_EINVAL, "no peer ISA attached");
    const int G = peer.nparts;
    B200SA_CU(cudaMemcpyAsync(h_pinned + 400, peer_inbox.p, kMaxPeers * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    for (int s = 0; s < G; ++s) {
        const u32 cnt = h_pinned[400 + s];
        if (cnt == 0) continue;
        if (cnt > peer.region_cap[s]) return set_error(B200SA_EINTERNAL, "GPU %d announced %u pairs for a region of %u", s, cnt, peer.region_cap[s]);
        const u32* keys_s = (const u32*)(peer_inbox.as<u8>() + peer.region_off[s]);
        const u32* vals_s = keys_s + peer.region_cap[s];
        // the run arrives bucketed by the top bits of the suffix index: the stores walk through L2-sized windows
        B200SA_TRY(phase_begin(B200SA_PH_ISA, st));
        B200SA_LAUNCH(k_scatter_pairs, (u32)div_up_u64(cnt, SP_THREADS * SP_IPT), SP_THREADS, 0, st, keys_s, vals_s, cnt, rank.as<u32>());
        count_launch(B200SA_PH_ISA);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_ISA] += (u64)cnt * 12;
    }
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// forward BWT

// output bytes [o_begin, o_end) from a finished SA (+ rank[0] = sentinel row); the sentinel row lands
// in h_pinned[8]
int Engine::bwt_rows(const u8* d_text, u32 n, const i32* d_sa, u32 o_begin, u32 o_end, u8* d_bwt, cudaStream_t st)
{
    i32* d_sent = (i32*)(misc.as<u32>() + 528);
    B200SA_TRY(phase_begin(B200SA_PH_BWT, st));
    {
        const u32 groups = (u32)div_up_u64(o_end - o_begin, 4);
        const u32 want = (u32)div_up_u64(groups, BW_THREADS * BW_STEPS);
        const u32 grid = want < (u32)(num_sms * 16) ? (want ? want : 1u) : (u32)(num_sms * 16);
        // rank[0] (the sentinel row) lives on GPU 0 when the ISA is sharded in peer memory
        const u32* rank0 = (peer.active && ss.nparts > 1) ? (const u32*)peer.view.base[0] : (const u32*)rank.as<u32>();
        B200SA_LAUNCH(k_bwt_gather, grid, BW_THREADS, 0, st, d_text, d_sa, rank0, o_begin, o_end, d_bwt, d_sent);
        count_launch(B200SA_PH_BWT);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_BWT] += (u64)(o_end - o_begin) * 6;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned + 8, d_sent, sizeof(i32), cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    (void)n;
    return 0;
}

int Engine::bwt_dev(const u8* d_text, i64 n64, u8* d_bwt, i32* d_sa_or_null, i64* sentinel_host, cudaStream_t st, i64 max_n)
{
    if (n64 < 0 || n64 > max_n) return set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n64, (long long)max_n);
    if (n64 > 0 && (!d_text || !d_bwt)) return set_error(B200SA_EINVAL, "null pointer");
    if (n64 == 0) {
        if (sentinel_host) *sentinel_host = 0;
        if (d_sa_or_null) return suffix_array_dev(d_text, 0, d_sa_or_null, st);
        return 0;
    }
    const u32 n = (u32)n64;
    i32* d_sa = d_sa_or_null;
    if (!d_sa) {
        B200SA_TRY(sa_ws.ensure(((size_t)n + 1) * 4));
        d_sa = sa_ws.as<i32>();
    }
    B200SA_TRY(suffix_array_dev(d_text, n64, d_sa, st, max_n));
    B200SA_TRY(bwt_rows(d_text, n, d_sa, 0, n, d_bwt, st));
    if (sentinel_host) *sentinel_host = (i64)h_pinned[8];
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// inverse BWT

// Three steps so that a sharded run can split the walkers over GPUs: build (psi table, F-column table,
// seed marks: replicated), measure (a slice of the walkers), finish (list ranking over all walkers, then
// emit the slice's bytes).
int Engine::unbwt_build(const u8* d_bwt, u32 n, u32 s, u32* nwalkers_out, cudaStream_t st)
{
    B200SA_TRY(keys[0].ensure(((size_t)n + 1) * 4 + 64));
    B200SA_TRY(misc.ensure(8192));
    us = UnbwtState();
    us.n = n; us.s = s;
    u32* psi = keys[0].as<u32>();
    u32* fstart = misc.as<u32>() + 600;  // 257 words

    // ---- psi table: one stable counting-sort sweep of the rows by BWT byte
    const u32 tiles = (u32)div_up_u64(n, RS_TILE);
    const size_t status_bytes = (size_t)tiles * RS_RADIX * sizeof(u64);
    B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
    u32* ghist = sortmeta.as<u32>();
    u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
    u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
    B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
    prof.memsets++;
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_BUILD, st));
    {
        const u32 htiles = (u32)div_up_u64(n, RH_THREADS * RH_IPT);
        const u32 grid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
        auto kh = k_radix_hist<u8>;
        B200SA_LAUNCH(kh, grid, RH_THREADS, rh_smem_bytes(1), st, d_bwt, n, 0, 1, ghist);
        count_launch(B200SA_PH_UNBWT_BUILD);
        B200SA_LAUNCH(k_radix_scan_bins, 1, RS_RADIX, 0, st, ghist);
        count_launch(B200SA_PH_UNBWT_BUILD);
        B200SA_LAUNCH(k_unbwt_fstart, 1, 256, 0, st, (const u32*)ghist, n, fstart);
        count_launch(B200SA_PH_UNBWT_BUILD);
        auto kp = k_onesweep_pass<u8, false>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u8>(), st, d_bwt, (u8*)nullptr, (const u32*)nullptr, psi + 1,
                      n, 0, s, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_UNBWT_BUILD);
    }
    // ---- walkers: every D-th row plus row s
    u32 D = (u32)div_up_u64((u64)n + 1, (u64)1 << 21);
    if (D < 64) D = 64;
    us.D = D;
    us.nreg = (u32)div_up_u64((u64)n + 1, D);
    us.nwalkers = us.nreg + ((s % D) != 0 ? 1u : 0u);
    B200SA_TRY(walk.ensure((size_t)us.nwalkers * 5 * 4 + 64));
    us.cap = (unbwt_cap_mult * D + 7u) & ~7u;  // window bytes per walker, 8-byte granular (64-bit stores)
    B200SA_LAUNCH(k_unbwt_mark, (u32)div_up_u64(us.nwalkers, 256), 256, 0, st, psi, us.nwalkers, us.nreg, D, s);
    count_launch(B200SA_PH_UNBWT_BUILD);
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_BUILD] += (u64)n * 6;
    B200SA_CU(cudaGetLastError());
    us.stage = 1;
    *nwalkers_out = us.nwalkers;
    return 0;
}

int Engine::unbwt_measure(u32 w_begin, u32 w_end, cudaStream_t st)
{
    if (us.stage < 1 || w_begin > w_end || w_end > us.nwalkers) return set_error(B200SA_EINVAL, "unbwt_measure: bad state or range");
    if (w_end == w_begin) return 0;
    const size_t W = us.nwalkers;
    u32* nx0 = walk.as<u32>();
    u32* ds0 = walk.as<u32>() + 2 * W;
    u32* ovf = walk.as<u32>() + 4 * W;
    // decode windows: cap bytes per walker of this slice, addressed by the global walker number
    B200SA_TRY(keys[1].ensure((size_t)W * us.cap + 256));
    B200SA_TRY(idx[0].ensure((size_t)W * 4 + 64));
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_WALK, st));
    B200SA_LAUNCH(k_unbwt_walk, (u32)div_up_u64(w_end - w_begin, UW_THREADS), UW_THREADS, 0, st, (const u32*)keys[0].as<u32>(),
                  (const u32*)(misc.as<u32>() + 600), w_begin, w_end, us.nreg, us.D, us.s, us.cap, keys[1].as<u8>(), ds0, nx0, ovf);
    count_launch(B200SA_PH_UNBWT_WALK);
    // the list ranking below overwrites the lengths: keep a copy for the placement pass
    B200SA_CU(cudaMemcpyAsync(idx[0].as<u32>() + w_begin, ds0 + w_begin, (size_t)(w_end - w_begin) * 4, cudaMemcpyDeviceToDevice, st));
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    return 0;
}

int Engine::unbwt_finish(u32 w_begin, u32 w_end, u8* d_out, cudaStream_t st)
{
    if (us.stage < 1 || w_begin > w_end || w_end > us.nwalkers) return set_error(B200SA_EINVAL, "unbwt_finish: bad state or range");
    const size_t W = us.nwalkers;
    u32* nx[2] = {walk.as<u32>(), walk.as<u32>() + W};
    u32* ds[2] = {walk.as<u32>() + 2 * W, walk.as<u32>() + 3 * W};
    u32* ovf = walk.as<u32>() + 4 * W;
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_WALK, st));
    const u32 g256 = (u32)div_up_u64(W, 256);
    int cur = 0;
    const int jumps = bit_length_u64(W);
    for (int it = 0; it < jumps; ++it) {
        B200SA_LAUNCH(k_unbwt_jump, g256, 256, 0, st, (const u32*)nx[cur], (const u32*)ds[cur], nx[cur ^ 1], ds[cur ^ 1], (u32)W);
        count_launch(B200SA_PH_UNBWT_WALK);
        cur ^= 1;
    }
    if (w_end > w_begin) {
        B200SA_LAUNCH(k_unbwt_place, (u32)div_up_u64(w_end - w_begin, UP_THREADS / 32), UP_THREADS, 0, st, (const u32*)keys[0].as<u32>(),
                      (const u32*)(misc.as<u32>() + 600), (const u32*)ds[cur], (const u32*)idx[0].as<u32>(), (const u32*)ovf,
                      (const u8*)keys[1].as<u8>(), us.cap, w_begin, w_end, us.n, d_out);
        count_launch(B200SA_PH_UNBWT_WALK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_WALK] += (u64)us.n * 7;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    // the jumps consumed the measured segments: a second finish needs a new measure pass
    us.stage = 0;
    return 0;
}

int Engine::unbwt_dev(const u8* d_bwt, i64 n64, i32 sentinel, u8* d_out, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32) return set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n64);
    if (n64 == 0) return 0;
    if (!d_bwt || !d_out) return set_error(B200SA_EINVAL, "null pointer");
    if (sentinel < 1 || (i64)sentinel > n64) return set_error(B200SA_EINVAL, "sentinel index %d outside [1, n]", sentinel);
    B200SA_CU(cudaSetDevice(device));
    u32 W = 0;
    B200SA_TRY(unbwt_build(d_bwt, (u32)n64, (u32)sentinel, &W, st));
    B200SA_TRY(unbwt_measure(0, W, st));
    B200SA_TRY(unbwt_finish(0, W, d_out, st));
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// validator

int Engine::check_sa_dev(const u8* d_text, i64 n64, const i32* d_sa, i64* bad_rows, cudaStream_t st, i64 max_n)
{
    if (n64 < 0 || n64 > max_n || !d_sa || !bad_rows) return set_error(B200SA_EINVAL, "bad argument");
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64;
    B200SA_TRY(rank.ensure(((size_t)n + 1) * 4 + 64));
    B200SA_TRY(misc.ensure(4096));
    u32* isa = rank.as<u32>();
    unsigned long long* d_bad = (unsigned long long*)(misc.as<u32>() + 520);
    B200SA_CU(cudaMemsetAsync(isa, 0xff, ((size_t)n + 1) * 4, st));
    B200SA_CU(cudaMemsetAsync(d_bad, 0, 8, st));
    prof.memsets += 2;
    B200SA_TRY(phase_begin(B200SA_PH_CHECK, st));
    const u32 grid = (u32)(div_up_u64((u64)n + 1, 256) < (u64)(num_sms * 16) ? div_up_u64((u64)n + 1, 256) : (u64)(num_sms * 16));
    B200SA_LAUNCH(k_check_scatter, grid, 256, 0, st, d_sa, n, isa, d_bad);
    count_launch(B200SA_PH_CHECK);
    if (n > 1) {
        B200SA_LAUNCH(k_check_order, grid, 256, 0, st, d_text, d_sa, n, (const u32*)isa, d_bad);
        count_launch(B200SA_PH_CHECK);
    }
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    unsigned long long h_bad = 0;
    B200SA_CU(cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    *bad_rows = (i64)h_bad;
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LCP array (lcp_kernels.cuh): phi scatter, hierarchical PLCP levels, gather through the SA

int Engine::lcp_dev(const u8* d_text, i64 n64, const i32* d_sa, i32* d_lcp, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32) return set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n64);
    if (!d_sa || !d_lcp || (n64 > 0 && !d_text)) return set_error(B200SA_EINVAL, "null pointer");
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64;
    if (n == 0) {
        B200SA_CU(cudaMemsetAsync(d_lcp, 0, sizeof(i32), st));
        B200SA_CU(cudaStreamSynchronize(st));
        return 0;
    }
    if (lcp_direct && n >= 2) {
        // ---- direct route first (see k_lcp_direct); falls through to the PLCP route when too many rows outgrow the budget
        const u32 ovf_cap = n / 64 + 1024;
        B200SA_TRY(slot[1].ensure((size_t)ovf_cap * 4 + 64));
        B200SA_TRY(misc.ensure(4096));
        u32* d_ovf = misc.as<u32>() + 1000;
        B200SA_CU(cudaMemsetAsync(d_ovf, 0, 4, st));
        prof.memsets++;
        B200SA_TRY(phase_begin(B200SA_PH_LCP, st));
        const u32 want = (u32)div_up_u64((u64)n + 1, 256);
        const u32 grid = want < (u32)(num_sms * 16) ? want : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_lcp_direct, grid, 256, 0, st, d_text, n, d_sa, d_lcp, slot[1].as<u32>(), ovf_cap, d_ovf);
        count_launch(B200SA_PH_LCP);
        B200SA_TRY(phase_end(st));
        B200SA_CU(cudaMemcpyAsync(h_pinned + 440, d_ovf, 4, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        const u32 novf = h_pinned[440];
        if (novf <= ovf_cap) {
            if (novf) {
                B200SA_TRY(phase_begin(B200SA_PH_LCP, st));
                const u32 g2 = novf < (u32)(num_sms * 4) ? novf : (u32)(num_sms * 4);
                B200SA_LAUNCH(k_lcp_direct_finish, g2, LC_THREADS, 0, st, d_text, n, d_sa, d_lcp, (const u32*)slot[1].as<u32>(), novf);
                count_launch(B200SA_PH_LCP);
                B200SA_TRY(phase_end(st));
            }
            prof.alg_bytes[B200SA_PH_LCP] += (u64)n * (4 + 8 + 4);
            B200SA_CU(cudaGetLastError());
            B200SA_CU(cudaStreamSynchronize(st));
            if (profiling) B200SA_TRY(collect_profile());
            return 0;
        }
    }
    // workspace: phi -> gid, plcp -> slot[0], overflow list -> slot[1], bucketed scatter scratch -> agg_max / keys[0]
    B200SA_TRY(gid.ensure((size_t)n * 4 + 64));
    B200SA_TRY(slot[0].ensure((size_t)n * 4 + 64));
    B200SA_TRY(slot[1].ensure(((size_t)n / 2 + 1024 + 2) * 8 + 64));
    B200SA_TRY(agg_max.ensure((size_t)n * 4 + 64));
    B200SA_TRY(keys[0].ensure((size_t)n * 4 + 64));
    B200SA_TRY(misc.ensure(4096));
    u32* phi = gid.as<u32>();
    u32* plcp = slot[0].as<u32>();
    const size_t ovf_cap = (size_t)n / 2 + 1024 + 2;
    u32* ovf_pos = slot[1].as<u32>();
    u32* ovf_len = slot[1].as<u32>() + ovf_cap;
    u32* d_cnt = misc.as<u32>() + 960;  // one overflow counter per level (<= 33 levels)
    B200SA_CU(cudaMemsetAsync(d_cnt, 0, 40 * sizeof(u32), st));
    prof.memsets++;

    // ---- phi[SA[r]] = SA[r-1]: the pairs are two shifted views of the suffix array itself
    B200SA_TRY(isa_update((const u32*)d_sa + 1, (const u32*)d_sa, n, n, agg_max.as<u32>(), keys[0].as<u32>(), true, st, phi));

    // ---- PLCP, coarse to fine
    B200SA_TRY(phase_begin(B200SA_PH_LCP, st));
    u64 top = 1;
    while (div_up_u64(n, top) > 1024) top <<= 1;
    int level = 0;
    auto run_level = [&](u32 first, u64 step, u32 back, u32 ns) -> int {
        if (ns == 0) return 0;
        const u32 want = (u32)div_up_u64(ns, LC_THREADS);
        const u32 grid = want < (u32)(num_sms * 8) ? want : (u32)(num_sms * 8);
        B200SA_LAUNCH(k_plcp_level, grid, LC_THREADS, 0, st, d_text, n, (const u32*)phi, plcp, first, (u32)step, back, ns, ovf_pos, ovf_len,
                      d_cnt + level);
        count_launch(B200SA_PH_LCP);
        B200SA_LAUNCH(k_plcp_overflow, (u32)(num_sms * 4), LC_THREADS, 0, st, d_text, n, (const u32*)phi, plcp, (const u32*)ovf_pos,
                      (const u32*)ovf_len, (const u32*)(d_cnt + level));
        count_launch(B200SA_PH_LCP);
        ++level;
        return 0;
    };
    B200SA_TRY(run_level(0, top, 0, (u32)div_up_u64(n, top)));
    for (u64 S = top >> 1; S >= 1; S >>= 1) {
        const u32 ns = (u64)n > S ? (u32)(((u64)n - S - 1) / (2 * S) + 1) : 0u;
        B200SA_TRY(run_level((u32)S, 2 * S, (u32)S, ns));
    }
    // ---- lcp[r] = plcp[SA[r]]
    {
        const u64 quads = div_up_u64((u64)n + 1, 4);
        const u32 want = (u32)div_up_u64(quads, 256);
        const u32 grid = want < (u32)(num_sms * 16) ? (want ? want : 1u) : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_lcp_gather, grid, 256, 0, st, d_sa, n, (const u32*)plcp, d_lcp);
        count_launch(B200SA_PH_LCP);
    }
    B200SA_TRY(phase_end(st));
    // phi read + plcp write + text probes (one sector-sized access per position, counted as 8 bytes) + SA/plcp/lcp of the gather
    prof.alg_bytes[B200SA_PH_LCP] += (u64)n * (4 + 4 + 8 + 4 + 4 + 4);
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// batch of independent blocks (batch_kernels.cuh)

// block tables of a batch in device memory: [ends u32 count][offs u32 count+1][sent i32 count]
int Engine::batch_tables(const i64* offsets, u32 count, const i32* sentinels_or_null, u32** d_ends, u32** d_offs, i32** d_sent, cudaStream_t st)
{
    const size_t words = (size_t)count * 3 + 1;
    B200SA_TRY(batch_meta.ensure(words * 4 + 64));
    *d_ends = batch_meta.as<u32>();
    *d_offs = *d_ends + count;
    *d_sent = (i32*)(*d_offs + count + 1);
    std::vector<u32> h(words, 0u);
    for (u32 b = 0; b < count; ++b) h[b] = (u32)offsets[b + 1] + b;
    for (u32 b = 0; b <= count; ++b) h[count + b] = (u32)offsets[b];
    if (sentinels_or_null)
        for (u32 b = 0; b < count; ++b) h[(size_t)2 * count + 1 + b] = (u32)sentinels_or_null[b];
    B200SA_CU(cudaMemcpyAsync(*d_ends, h.data(), words * 4, cudaMemcpyHostToDevice, st));
    B200SA_CU(cudaStreamSynchronize(st));  // h goes out of scope
    return 0;
}

static int check_batch_offsets(const i64* offsets, i64 count64)
{
    if (count64 < 0 || count64 > ((i64)1 << 24) || (count64 > 0 && !offsets))
        return set_error(B200SA_EINVAL, "block count %lld outside [0, 2^24] or null offsets", (long long)count64);
    if (count64 == 0) return 0;
    if (offsets[0] != 0) return set_error(B200SA_EINVAL, "offsets[0] must be 0");
    for (i64 b = 0; b < count64; ++b)
        if (offsets[b + 1] < offsets[b]) return set_error(B200SA_EINVAL, "offsets must be non-decreasing (block %lld)", (long long)b);
    if (offsets[count64] + count64 > B200SA_MAX_N_INT32)
        return set_error(B200SA_EINVAL, "batch of %lld bytes in %lld blocks exceeds 2^31-2 rows", (long long)offsets[count64], (long long)count64);
    return 0;
}


int Engine::batch_dev(const u8* d_packed, const i64* offsets, i64 count64, u8* d_bwt_out, i32* d_sa_out, i32* sentinels_host, cudaStream_t st)
{
    B200SA_TRY(check_batch_offsets(offsets, count64));
    if (count64 == 0) return 0;
    const u32 count = (u32)count64;
    const i64 total64 = offsets[count];
    if (total64 > 0 && !d_packed) return set_error(B200SA_EINVAL, "null pointer");
    B200SA_CU(cudaSetDevice(device));
    const u32 total = (u32)total64, N = total + count;
    u32 *d_ends = nullptr, *d_offs = nullptr;
    i32* d_sent = nullptr;
    B200SA_TRY(batch_tables(offsets, count, nullptr, &d_ends, &d_offs, &d_sent, st));
    B200SA_TRY(batch_text.ensure((size_t)N + 64));
    B200SA_TRY(sa_ws.ensure(((size_t)N + 1) * 4));
    u8* d_text = batch_text.as<u8>();
    i32* d_sa = sa_ws.as<i32>();
    B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
    {
        const u32 want = (u32)div_up_u64(div_up_u64(N, BE_IPT), BE_THREADS);
        const u32 grid = want < (u32)(num_sms * 8) ? want : (u32)(num_sms * 8);
        B200SA_LAUNCH(k_batch_expand, grid, BE_THREADS, 0, st, d_packed, (const u32*)d_ends, count, N, d_text);
        count_launch(B200SA_PH_PACK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_PACK] += (u64)total + N;

    next_batch.d_ends = d_ends;
    next_batch.count = count;
    u32 n_local = 0, m = 0;
    B200SA_TRY(sort_begin(d_text, N, d_sa, 0, 1, &n_local, st));
    B200SA_TRY(sort_round0(0, &m, st));
    while (m > 0) B200SA_TRY(sort_round(&m, st));
    ss.stage = 3;

    if (d_sa_out) {
        B200SA_TRY(phase_begin(B200SA_PH_BWT, st));
        const u32 want = (u32)div_up_u64(div_up_u64(N, BL_IPT), BL_THREADS);
        const u32 grid = want < (u32)(num_sms * 8) ? want : (u32)(num_sms * 8);
        B200SA_LAUNCH(k_batch_localize, grid, BL_THREADS, 0, st, (const i32*)d_sa, (const u32*)d_ends, count, N, d_sa_out);
        count_launch(B200SA_PH_BWT);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_BWT] += (u64)N * 8;
    }
    if (d_bwt_out || sentinels_host) {
        B200SA_TRY(phase_begin(B200SA_PH_BWT, st));
        B200SA_LAUNCH(k_batch_sentinels, (u32)div_up_u64(count, 256), 256, 0, st, (const u32*)rank.as<u32>(), (const u32*)d_ends, count, d_sent);
        count_launch(B200SA_PH_BWT);
        if (d_bwt_out && total) {
            const u32 want = (u32)div_up_u64(div_up_u64(total, 4), BW_THREADS);
            const u32 grid = want < (u32)(num_sms * 16) ? want : (u32)(num_sms * 16);
            B200SA_LAUNCH(k_bwt_gather_batch, grid, BW_THREADS, 0, st, (const u8*)d_text, (const i32*)d_sa, (const u32*)d_offs,
                          (const u32*)d_ends, (const i32*)d_sent, count, total, d_bwt_out);
            count_launch(B200SA_PH_BWT);
            prof.alg_bytes[B200SA_PH_BWT] += (u64)total * 6;
        }
        B200SA_TRY(phase_end(st));
        if (sentinels_host) B200SA_CU(cudaMemcpyAsync(sentinels_host, d_sent, (size_t)count * 4, cudaMemcpyDeviceToHost, st));
    }
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// Inverse BWT of a batch (batch_kernels.cuh, second half): one sort builds psi and the symbols of all blocks,
// one walk decodes all blocks.
int Engine::unbwt_batch_dev(const u8* d_bwt, const i64* offsets, i64 count64, const i32* sentinels, u8* d_out, cudaStream_t st)
{
    B200SA_TRY(check_batch_offsets(offsets, count64));
    if (count64 == 0) return 0;
    if (!sentinels) return set_error(B200SA_EINVAL, "null sentinel indices");
    const u32 count = (u32)count64;
    for (u32 b = 0; b < count; ++b) {
        const i64 nb = offsets[b + 1] - offsets[b];
        if (nb > 0 && (sentinels[b] < 1 || (i64)sentinels[b] > nb))
            return set_error(B200SA_EINVAL, "sentinel index %d of block %u outside [1, %lld]", sentinels[b], b, (long long)nb);
    }
    const u32 total = (u32)offsets[count], N = total + count;
    if (total == 0) return 0;
    if (!d_bwt || !d_out) return set_error(B200SA_EINVAL, "null pointer");
    B200SA_CU(cudaSetDevice(device));
    u32 *d_ends = nullptr, *d_offs = nullptr;
    i32* d_sent = nullptr;
    B200SA_TRY(batch_tables(offsets, count, sentinels, &d_ends, &d_offs, &d_sent, st));

    // ---- sort (block << 8 | byte, row): 1 + ceil(block bits / 8) sweeps of u32 pairs
    const int key_bits = 8 + bit_length_u64((u64)count - 1);
    const int passes = (key_bits + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
    B200SA_TRY(keys[0].ensure((size_t)total * 4 + 64));
    B200SA_TRY(keys[1].ensure((size_t)total * 4 + 64));
    B200SA_TRY(idx[0].ensure((size_t)total * 4 + 64));
    B200SA_TRY(idx[1].ensure((size_t)total * 4 + 64));
    B200SA_TRY(batch_out.ensure((size_t)N * 8 + 64));
    u32* k2[2] = {keys[0].as<u32>(), keys[1].as<u32>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    u64* table = batch_out.as<u64>();
    const u32 tiles = (u32)div_up_u64(total, RS_TILE);
    const size_t status_bytes = (size_t)passes * tiles * RS_RADIX * sizeof(u64);
    B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
    u32* ghist = sortmeta.as<u32>();
    u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
    u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
    B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
    prof.memsets++;
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_BUILD, st));
    const u32 g4 = (u32)div_up_u64(div_up_u64(total, 4), 256);
    const u32 grid4 = g4 < (u32)(num_sms * 16) ? g4 : (u32)(num_sms * 16);
    B200SA_LAUNCH(k_ubb_gen, grid4, 256, 0, st, d_bwt, (const u32*)d_offs, (const i32*)d_sent, count, total, k2[0], v2[0]);
    count_launch(B200SA_PH_UNBWT_BUILD);
    {
        const u32 htiles = (u32)div_up_u64(total, RH_THREADS * RH_IPT);
        const u32 hgrid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
        auto kh = k_radix_hist<u32>;
        B200SA_LAUNCH(kh, hgrid, RH_THREADS, rh_smem_bytes(passes), st, (const u32*)k2[0], total, 0, passes, ghist);
        count_launch(B200SA_PH_UNBWT_BUILD);
        B200SA_LAUNCH(k_radix_scan_bins, passes, RS_RADIX, 0, st, ghist);
        count_launch(B200SA_PH_UNBWT_BUILD);
    }
    int side = 0;
    for (int p = 0; p < passes; ++p) {
        auto kp = k_onesweep_pass<u32, true>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u32>(), st, (const u32*)k2[side], k2[side ^ 1], (const u32*)v2[side], v2[side ^ 1],
                      total, p * RS_RADIX_BITS, 0xffffffffu, (const u32*)(ghist + p * RS_RADIX), status + (size_t)p * tiles * RS_RADIX, counters + p);
        count_launch(B200SA_PH_UNBWT_BUILD);
        side ^= 1;
    }
    {
        const u32 g1 = (u32)div_up_u64(total, 256);
        const u32 grid1 = g1 < (u32)(num_sms * 16) ? g1 : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_ubb_table, grid1, 256, 0, st, (const u32*)k2[side], (const u32*)v2[side], total, table);
        count_launch(B200SA_PH_UNBWT_BUILD);
        B200SA_LAUNCH(k_ubb_rows0, (u32)div_up_u64(count, 256), 256, 0, st, (const u32*)d_ends, (const i32*)d_sent, count, table);
        count_launch(B200SA_PH_UNBWT_BUILD);
    }
    // ---- walkers
    u32 D = (u32)div_up_u64((u64)N, (u64)1 << 21);
    if (D < 64) D = 64;
    const u32 nreg = (u32)div_up_u64(N, D);
    const u32 W = nreg + 2 * count;
    // the longest chain is the largest block's: (rows / D) regular walkers + start + terminal; no segment is longer than
    // the largest block either, which bounds the decode windows when a batch consists of very many tiny blocks
    u64 max_rows = 0;
    for (u32 b = 0; b < count; ++b) { const u64 r = (u64)(offsets[b + 1] - offsets[b]) + 1; max_rows = r > max_rows ? r : max_rows; }
    u32 cap = (unbwt_cap_mult * D + 7u) & ~7u;
    if ((u64)cap > ((max_rows + 7) & ~(u64)7)) cap = (u32)((max_rows + 7) & ~(u64)7);
    B200SA_TRY(walk.ensure((size_t)W * 5 * 4 + 64));
    B200SA_LAUNCH(k_ubb_mark, (u32)div_up_u64(nreg + count, 256), 256, 0, st, table, (const u32*)d_ends, (const i32*)d_sent, count, nreg, D, N);
    count_launch(B200SA_PH_UNBWT_BUILD);
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_BUILD] += (u64)total * (1 + 8 + 4 + 16 * (u64)passes + 8 + 8);
    B200SA_CU(cudaGetLastError());

    // the sort buffers are free again: decode windows go to keys[0], the saved segment lengths to idx[0]
    B200SA_TRY(keys[0].ensure((size_t)W * cap + 256));
    B200SA_TRY(idx[0].ensure((size_t)W * 4 + 64));
    u32* nx[2] = {walk.as<u32>(), walk.as<u32>() + W};
    u32* ds[2] = {walk.as<u32>() + 2 * (size_t)W, walk.as<u32>() + 3 * (size_t)W};
    u32* ovf = walk.as<u32>() + 4 * (size_t)W;
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_WALK, st));
    B200SA_LAUNCH(k_ubb_walk, (u32)div_up_u64(W, UW_THREADS), UW_THREADS, 0, st, (const u64*)table, (const u32*)d_ends, (const i32*)d_sent, count,
                  nreg, D, N, W, cap, keys[0].as<u8>(), ds[0], nx[0], ovf);
    count_launch(B200SA_PH_UNBWT_WALK);
    B200SA_CU(cudaMemcpyAsync(idx[0].p, ds[0], (size_t)W * 4, cudaMemcpyDeviceToDevice, st));
    const u32 g256 = (u32)div_up_u64(W, 256);
    int cur = 0;
    const int jumps = bit_length_u64(max_rows / D + 3);
    for (int it = 0; it < jumps; ++it) {
        B200SA_LAUNCH(k_unbwt_jump, g256, 256, 0, st, (const u32*)nx[cur], (const u32*)ds[cur], nx[cur ^ 1], ds[cur ^ 1], W);
        count_launch(B200SA_PH_UNBWT_WALK);
        cur ^= 1;
    }
    B200SA_LAUNCH(k_ubb_place, (u32)div_up_u64(W, UP_THREADS / 32), UP_THREADS, 0, st, (const u64*)table, (const u32*)d_ends, (const u32*)d_offs,
                  (const i32*)d_sent, count, nreg, D, N, W, (const u32*)ds[cur], (const u32*)idx[0].as<u32>(), (const u32*)ovf,
                  (const u8*)keys[0].as<u8>(), cap, d_out);
    count_launch(B200SA_PH_UNBWT_WALK);
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_WALK] += (u64)total * 9;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

}  // namespace b200sa

// =============================================================================================
// C ABI

using b200sa::Engine;

struct b200sa_ctx {
    Engine eng;
};

extern "C" {

int b200sa_version(void) { return 100; }

const char* b200sa_last_error(void) { return b200sa::g_err; }

int b200sa_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int b200sa_create(b200sa_ctx** out, int device)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    *out = nullptr;
    b200sa_ctx* c = new (std::nothrow) b200sa_ctx();
    if (!c) return b200sa::set_error(B200SA_ENOMEM, "out of host memory");
    int rc = c->eng.init(device);
    if (rc != 0) { delete c; return rc; }
    *out = c;
    return 0;
}

void b200sa_destroy(b200sa_ctx* ctx)
{
    if (!ctx) return;
    ctx->eng.shutdown();
    delete ctx;
}

int b200sa_release_workspace(b200sa_ctx* ctx)
{
    if (!ctx) return b200sa::set_error(B200SA_EINVAL, "null context");
    return ctx->eng.release_workspace();
}

#define B200SA_NEED_CTX(ctx) \
    if (!(ctx)) return b200sa::set_error(B200SA_EINVAL, "null context (call b200sa_create first)")

int b200sa_suffix_array_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, int32_t* d_sa_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.suffix_array_dev(d_text, n, d_sa_out, ctx->eng.pick(stream));
}

int b200sa_bwt_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, uint8_t* d_bwt_out, int32_t* d_sa_out,
                   int32_t* sentinel_index_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n > 0 && d_text == d_bwt_out) return b200sa::set_error(B200SA_EINVAL, "d_bwt_out must not alias d_text");
    i64 s64 = 0;
    B200SA_TRY(ctx->eng.bwt_dev(d_text, n, d_bwt_out, d_sa_out, &s64, ctx->eng.pick(stream)));
    if (sentinel_index_out) *sentinel_index_out = (int32_t)s64;
    return 0;
}

// ---- wide-index superset: uint32 suffix arrays, n up to B200SA_MAX_N_UINT32 ---------------------------

int b200sa_suffix_array_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, uint32_t* d_sa_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.suffix_array_dev(d_text, n, (i32*)d_sa_out, ctx->eng.pick(stream), B200SA_MAX_N_UINT32);
}

int b200sa_bwt_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, uint8_t* d_bwt_out, uint32_t* d_sa_out,
                       int64_t* sentinel_index_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n > 0 && d_text == d_bwt_out) return b200sa::set_error(B200SA_EINVAL, "d_bwt_out must not alias d_text");
    return ctx->eng.bwt_dev(d_text, n, d_bwt_out, (i32*)d_sa_out, sentinel_index_out, ctx->eng.pick(stream), B200SA_MAX_N_UINT32);
}

int b200sa_check_suffix_array_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const uint32_t* d_sa, int64_t* bad_rows_out,
                                      void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.check_sa_dev(d_text, n, (const i32*)d_sa, bad_rows_out, ctx->eng.pick(stream), B200SA_MAX_N_UINT32);
}

int b200sa_unbwt_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n, int32_t sentinel_index, uint8_t* d_text_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n > 0 && d_bwt == d_text_out) return b200sa::set_error(B200SA_EINVAL, "d_text_out must not alias d_bwt");
    return ctx->eng.unbwt_dev(d_bwt, n, sentinel_index, d_text_out, ctx->eng.pick(stream));
}

int b200sa_check_suffix_array_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const int32_t* d_sa,
                                  int64_t* bad_rows_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.check_sa_dev(d_text, n, d_sa, bad_rows_out, ctx->eng.pick(stream));
}

int b200sa_lcp_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const int32_t* d_sa, int32_t* d_lcp_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.lcp_dev(d_text, n, d_sa, d_lcp_out, ctx->eng.pick(stream));
}

// ---- host-buffer entry points ----------------------------------------------------------------

int b200sa_suffix_array_bwt(b200sa_ctx* ctx, const uint8_t* text, int64_t n, int32_t* sa_out, uint8_t* bwt_out,
                            int32_t* sentinel_index_out)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n);
    if (n > 0 && !text) return b200sa::set_error(B200SA_EINVAL, "null text");
    if (n == 0) {
        if (sa_out) sa_out[0] = 0;
        if (sentinel_index_out) *sentinel_index_out = 0;
        // still require a device: this library never computes on the CPU
        return 0;
    }
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n));
    B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
    B200SA_CU(cudaMemcpyAsync(e.text_ws.p, text, (size_t)n, cudaMemcpyHostToDevice, st));
    const bool want_bwt = bwt_out != nullptr || sentinel_index_out != nullptr;
    if (want_bwt) {
        B200SA_TRY(e.bwt_ws.ensure((size_t)n));
        i64 sentinel = 0;
        B200SA_TRY(e.bwt_dev(e.text_ws.as<u8>(), n, e.bwt_ws.as<u8>(), e.sa_ws.as<i32>(), &sentinel, st));
        if (sentinel_index_out) *sentinel_index_out = (int32_t)sentinel;
        if (bwt_out) B200SA_CU(cudaMemcpyAsync(bwt_out, e.bwt_ws.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    } else {
        B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st));
    }
    if (sa_out) B200SA_CU(cudaMemcpyAsync(sa_out, e.sa_ws.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_suffix_array_bwt_u32(b200sa_ctx* ctx, const uint8_t* text, int64_t n, uint32_t* sa_out, uint8_t* bwt_out,
                                int64_t* sentinel_index_out)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > B200SA_MAX_N_UINT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^32-8194]", (long long)n);
    if (n > 0 && !text) return b200sa::set_error(B200SA_EINVAL, "null text");
    if (n == 0) {
        if (sa_out) sa_out[0] = 0;
        if (sentinel_index_out) *sentinel_index_out = 0;
        return 0;
    }
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n));
    B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
    B200SA_CU(cudaMemcpyAsync(e.text_ws.p, text, (size_t)n, cudaMemcpyHostToDevice, st));
    if (bwt_out || sentinel_index_out) {
        B200SA_TRY(e.bwt_ws.ensure((size_t)n));
        i64 sentinel = 0;
        B200SA_TRY(e.bwt_dev(e.text_ws.as<u8>(), n, e.bwt_ws.as<u8>(), e.sa_ws.as<i32>(), &sentinel, st, B200SA_MAX_N_UINT32));
        if (sentinel_index_out) *sentinel_index_out = sentinel;
        if (bwt_out) B200SA_CU(cudaMemcpyAsync(bwt_out, e.bwt_ws.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    } else {
        B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st, B200SA_MAX_N_UINT32));
    }
    if (sa_out) B200SA_CU(cudaMemcpyAsync(sa_out, e.sa_ws.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_suffix_array(b200sa_ctx* ctx, const uint8_t* text, int64_t n, int32_t* sa_out)
{
    if (!sa_out) return b200sa::set_error(B200SA_EINVAL, "null sa_out");
    return b200sa_suffix_array_bwt(ctx, text, n, sa_out, nullptr, nullptr);
}

int b200sa_bwt(b200sa_ctx* ctx, uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out)
{
    if (!sentinel_index_out) return b200sa::set_error(B200SA_EINVAL, "null sentinel_index_out");
    return b200sa_suffix_array_bwt(ctx, text_inout, n, nullptr, text_inout, sentinel_index_out);
}

int b200sa_unbwt(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int32_t sentinel_index)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n);
    if (n == 0) return 0;
    if (!bwt_inout) return b200sa::set_error(B200SA_EINVAL, "null buffer");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n));
    B200SA_TRY(e.bwt_ws.ensure((size_t)n));
    B200SA_CU(cudaMemcpyAsync(e.bwt_ws.p, bwt_inout, (size_t)n, cudaMemcpyHostToDevice, st));
    B200SA_TRY(e.unbwt_dev(e.bwt_ws.as<u8>(), n, sentinel_index, e.text_ws.as<u8>(), st));
    B200SA_CU(cudaMemcpyAsync(bwt_inout, e.text_ws.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_lcp(b200sa_ctx* ctx, const uint8_t* text, int64_t n, const int32_t* sa, int32_t* sa_out, int32_t* lcp_out)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n);
    if (!lcp_out || (n > 0 && !text)) return b200sa::set_error(B200SA_EINVAL, "null pointer");
    if (n == 0) {
        lcp_out[0] = 0;
        if (sa_out) sa_out[0] = 0;
        return 0;
    }
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n));
    B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
    B200SA_TRY(e.keys[1].ensure(((size_t)n + 1) * 4 + 64));  // LCP staging: keys[1] is not used by lcp_dev
    B200SA_CU(cudaMemcpyAsync(e.text_ws.p, text, (size_t)n, cudaMemcpyHostToDevice, st));
    if (sa) B200SA_CU(cudaMemcpyAsync(e.sa_ws.p, sa, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, st));
    else B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st));
    if (sa_out) B200SA_CU(cudaMemcpyAsync(sa_out, e.sa_ws.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    B200SA_TRY(e.lcp_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), e.keys[1].as<i32>(), st));
    B200SA_CU(cudaMemcpyAsync(lcp_out, e.keys[1].p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

// ---- batches of independent blocks ------------------------------------------------------------------

int b200sa_batch_dev(b200sa_ctx* ctx, const uint8_t* d_blocks, const int64_t* offsets, int64_t count, uint8_t* d_bwt_out,
                     int32_t* d_sa_out, int32_t* sentinel_index_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (d_bwt_out && d_bwt_out == d_blocks) return b200sa::set_error(B200SA_EINVAL, "d_bwt_out must not alias d_blocks");
    return ctx->eng.batch_dev(d_blocks, offsets, count, d_bwt_out, d_sa_out, sentinel_index_out, ctx->eng.pick(stream));
}

int b200sa_unbwt_batch_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, const int64_t* offsets, int64_t count, const int32_t* sentinel_index,
                           uint8_t* d_text_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (d_text_out && d_text_out == d_bwt) return b200sa::set_error(B200SA_EINVAL, "d_text_out must not alias d_bwt");
    return ctx->eng.unbwt_batch_dev(d_bwt, offsets, count, sentinel_index, d_text_out, ctx->eng.pick(stream));
}

static int batch_host(b200sa_ctx* ctx, const uint8_t* blocks, const int64_t* offsets, int64_t count, uint8_t* bwt_out, int32_t* sa_out,
                      int32_t* sentinel_index_out)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (count < 0 || (count > 0 && !offsets)) return b200sa::set_error(B200SA_EINVAL, "bad block table");
    if (count == 0) return 0;
    const int64_t total = offsets[count];
    if (total < 0 || total + count > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "batch too large for 32-bit suffix indices");
    if (total > 0 && !blocks) return b200sa::set_error(B200SA_EINVAL, "null blocks");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)total + 64));
    if (total) B200SA_CU(cudaMemcpyAsync(e.text_ws.p, blocks, (size_t)total, cudaMemcpyHostToDevice, st));
    u8* d_bwt = nullptr;
    i32* d_sa = nullptr;
    if (bwt_out) { B200SA_TRY(e.bwt_ws.ensure((size_t)total + 64)); d_bwt = e.bwt_ws.as<u8>(); }
    if (sa_out) { B200SA_TRY(e.batch_out.ensure(((size_t)total + (size_t)count) * 4 + 64)); d_sa = e.batch_out.as<i32>(); }
    B200SA_TRY(e.batch_dev(e.text_ws.as<u8>(), offsets, count, d_bwt, d_sa, sentinel_index_out, st));
    if (bwt_out && total) B200SA_CU(cudaMemcpyAsync(bwt_out, d_bwt, (size_t)total, cudaMemcpyDeviceToHost, st));
    if (sa_out) B200SA_CU(cudaMemcpyAsync(sa_out, d_sa, ((size_t)total + (size_t)count) * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_suffix_array_batch(b200sa_ctx* ctx, const uint8_t* blocks, const int64_t* offsets, int64_t count, int32_t* sa_out)
{
    if (count > 0 && !sa_out) return b200sa::set_error(B200SA_EINVAL, "null sa_out");
    return batch_host(ctx, blocks, offsets, count, nullptr, sa_out, nullptr);
}

int b200sa_bwt_batch(b200sa_ctx* ctx, uint8_t* blocks_inout, const int64_t* offsets, int64_t count, int32_t* sentinel_index_out)
{
    if (count > 0 && !sentinel_index_out) return b200sa::set_error(B200SA_EINVAL, "null sentinel_index_out");
    return batch_host(ctx, blocks_inout, offsets, count, blocks_inout, nullptr, sentinel_index_out);
}

int b200sa_unbwt_batch(b200sa_ctx* ctx, uint8_t* blocks_inout, const int64_t* offsets, int64_t count, const int32_t* sentinel_index)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (count < 0 || (count > 0 && (!offsets || !sentinel_index))) return b200sa::set_error(B200SA_EINVAL, "bad block table");
    if (count == 0) return 0;
    if (offsets[0] != 0) return b200sa::set_error(B200SA_EINVAL, "offsets[0] must be 0");
    const int64_t total = offsets[count];
    for (int64_t b = 0; b < count; ++b) {
        const int64_t nb = offsets[b + 1] - offsets[b];
        if (nb < 0 || nb > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "bad size of block %lld", (long long)b);
        if (nb > 0 && (sentinel_index[b] < 1 || (int64_t)sentinel_index[b] > nb))
            return b200sa::set_error(B200SA_EINVAL, "sentinel index %d of block %lld outside [1, %lld]", sentinel_index[b], (long long)b, (long long)nb);
    }
    if (total > 0 && !blocks_inout) return b200sa::set_error(B200SA_EINVAL, "null blocks");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.bwt_ws.ensure((size_t)total + 64));
    B200SA_TRY(e.text_ws.ensure((size_t)total + 64));
    if (total) B200SA_CU(cudaMemcpyAsync(e.bwt_ws.p, blocks_inout, (size_t)total, cudaMemcpyHostToDevice, st));
    B200SA_TRY(e.unbwt_batch_dev(e.bwt_ws.as<u8>(), offsets, count, sentinel_index, e.text_ws.as<u8>(), st));
    if (total) B200SA_CU(cudaMemcpyAsync(blocks_inout, e.text_ws.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

// ---- streaming pipeline over batches --------------------------------------------------------------
//
// `depth` contexts (each: own stream, own workspace, own worker thread) take submitted batches from one queue.
// While one context sorts, another uploads its next batch and a third downloads its results, so the copy
// engines and the SMs overlap across batches; inside a batch nothing changes.  Host buffers handed to submit
// must stay valid until the ticket has been waited for; pinned memory makes the copies truly asynchronous.

}  // extern "C"

#include <condition_variable>
#include <deque>
#include <map>
#include <set>
#include <thread>

struct b200sa_pipeline {
    struct Job {
        int64_t ticket;
        int kind;  // 0 forward BWT, 1 inverse BWT, 2 suffix arrays
        uint8_t* blocks;
        const int64_t* offsets;
        int64_t count;
        int32_t* sentinels;  // out (forward) / in (inverse)
        int32_t* sa_out;
    };
    std::vector<b200sa_ctx*> ctxs;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Job> queue;
    std::map<int64_t, std::pair<int, std::string>> done;  // ticket -> (status, message), until somebody waits for it
    std::set<int64_t> open_tickets;                       // submitted and not yet collected by wait / drain
    int64_t next_ticket = 1;
    int64_t in_flight = 0;
    bool stopping = false;

    void run(size_t w)
    {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stopping || !queue.empty(); });
                if (queue.empty()) return;
                job = queue.front();
                queue.pop_front();
            }
            int rc;
            {
#ifdef B200SA_EMU
                static std::mutex emu_mu;  // the CPU emulator is single-threaded
                std::lock_guard<std::mutex> g(emu_mu);
#endif
                if (job.kind == 0) rc = b200sa_bwt_batch(ctxs[w], job.blocks, job.offsets, job.count, job.sentinels);
                else if (job.kind == 1) rc = b200sa_unbwt_batch(ctxs[w], job.blocks, job.offsets, job.count, job.sentinels);
                else rc = b200sa_suffix_array_batch(ctxs[w], job.blocks, job.offsets, job.count, job.sa_out);
            }
            std::string msg = rc ? b200sa_last_error() : "";
            {
                std::lock_guard<std::mutex> lk(mu);
                done[job.ticket] = std::make_pair(rc, msg);
                --in_flight;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" {

int b200sa_pipeline_create(b200sa_pipeline** out, int device, int depth)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    *out = nullptr;
    if (depth < 1 || depth > 8) return b200sa::set_error(B200SA_EINVAL, "pipeline depth %d outside [1, 8]", depth);
    b200sa_pipeline* p = new (std::nothrow) b200sa_pipeline();
    if (!p) return b200sa::set_error(B200SA_ENOMEM, "out of host memory");
    for (int i = 0; i < depth; ++i) {
        b200sa_ctx* c = nullptr;
        const int rc = b200sa_create(&c, device);
        if (rc != 0) {
            for (auto* q : p->ctxs) b200sa_destroy(q);
            delete p;
            return rc;
        }
        p->ctxs.push_back(c);
    }
    for (int i = 0; i < depth; ++i) p->workers.emplace_back([p, i] { p->run((size_t)i); });
    *out = p;
    return 0;
}

void b200sa_pipeline_destroy(b200sa_pipeline* p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stopping = true;  // workers finish what is queued, then leave
    }
    p->cv_work.notify_all();
    for (auto& t : p->workers) t.join();
    for (auto* c : p->ctxs) b200sa_destroy(c);
    delete p;
}

static int pipeline_submit(b200sa_pipeline* p, int kind, uint8_t* blocks, const int64_t* offsets, int64_t count, int32_t* sentinels,
                           int32_t* sa_out, int64_t* ticket_out)
{
    if (!p || !ticket_out) return b200sa::set_error(B200SA_EINVAL, "null pipeline or ticket pointer");
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->stopping) return b200sa::set_error(B200SA_EINVAL, "pipeline is shutting down");
        *ticket_out = p->next_ticket++;
        p->open_tickets.insert(*ticket_out);
        p->queue.push_back(b200sa_pipeline::Job{*ticket_out, kind, blocks, offsets, count, sentinels, sa_out});
        ++p->in_flight;
    }
    p->cv_work.notify_one();
    return 0;
}

int b200sa_pipeline_submit_bwt(b200sa_pipeline* p, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                               int32_t* sentinel_index_out, int64_t* ticket_out)
{
    return pipeline_submit(p, 0, blocks_inout, offsets, count, sentinel_index_out, nullptr, ticket_out);
}

int b200sa_pipeline_submit_unbwt(b200sa_pipeline* p, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                 const int32_t* sentinel_index, int64_t* ticket_out)
{
    return pipeline_submit(p, 1, blocks_inout, offsets, count, const_cast<int32_t*>(sentinel_index), nullptr, ticket_out);
}

int b200sa_pipeline_submit_suffix_array(b200sa_pipeline* p, const uint8_t* blocks, const int64_t* offsets, int64_t count,
                                        int32_t* sa_out, int64_t* ticket_out)
{
    return pipeline_submit(p, 2, const_cast<uint8_t*>(blocks), offsets, count, nullptr, sa_out, ticket_out);
}

int b200sa_pipeline_wait(b200sa_pipeline* p, int64_t ticket)
{
    if (!p) return b200sa::set_error(B200SA_EINVAL, "null pipeline");
    std::unique_lock<std::mutex> lk(p->mu);
    if (p->open_tickets.count(ticket) == 0)
        return b200sa::set_error(B200SA_EINVAL, "ticket %lld is unknown or has already been collected", (long long)ticket);
    p->cv_done.wait(lk, [&] { return p->done.count(ticket) != 0; });
    auto it = p->done.find(ticket);
    const int rc = it->second.first;
    if (rc) b200sa::set_error(rc, "%s", it->second.second.c_str());
    p->done.erase(it);
    p->open_tickets.erase(ticket);
    return rc;
}

int b200sa_pipeline_drain(b200sa_pipeline* p)
{
    if (!p) return b200sa::set_error(B200SA_EINVAL, "null pipeline");
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->in_flight == 0; });
    int first = 0;
    for (auto& kv : p->done)
        if (kv.second.first && !first) { first = kv.second.first; b200sa::set_error(first, "%s", kv.second.second.c_str()); }
    p->done.clear();
    p->open_tickets.clear();
    return first;
}

// ---- sharded (multi-GPU) building blocks ---------------------------------------------------------

int b200sa_shard_begin(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, int32_t* d_sa, int part, int nparts,
                       int64_t* n_local_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n <= 0 || n > B200SA_MAX_N_INT32 || !d_text || !d_sa || !n_local_out || nparts < 1 || part < 0 || part >= nparts)
        return b200sa::set_error(B200SA_EINVAL, "bad argument");
    Engine& e = ctx->eng;
    B200SA_CU(cudaSetDevice(e.device));
    u32 nl = 0;
    B200SA_TRY(e.sort_begin(d_text, (u32)n, d_sa, part, nparts, &nl, e.pick(stream)));
    *n_local_out = nl;
    return 0;
}

int b200sa_shard_round0(b200sa_ctx* ctx, int64_t slot_base, int64_t* m_local_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (!m_local_out || slot_base < 0) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    u32 m = 0;
    B200SA_TRY(ctx->eng.sort_round0((u32)slot_base, &m, ctx->eng.pick(stream)));
    *m_local_out = m;
    return 0;
}

int b200sa_shard_round(b200sa_ctx* ctx, int64_t* m_local_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (!m_local_out) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    u32 m = 0;
    B200SA_TRY(ctx->eng.sort_round(&m, ctx->eng.pick(stream)));
    *m_local_out = m;
    return 0;
}

int b200sa_shard_updates(b200sa_ctx* ctx, const uint32_t** d_idx_out, const uint32_t** d_rank_out, int64_t* count_out)
{
    B200SA_NEED_CTX(ctx);
    if (!d_idx_out || !d_rank_out || !count_out) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    *d_idx_out = ctx->eng.ss.upd_idx;
    *d_rank_out = ctx->eng.ss.upd_rank;
    *count_out = ctx->eng.ss.upd_count;
    return 0;
}

int b200sa_shard_copy_updates(b200sa_ctx* ctx, uint32_t* d_idx_dst, uint32_t* d_rank_dst, int64_t capacity, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    const int64_t count = e.ss.upd_count;
    if (count > capacity || (count > 0 && (!d_idx_dst || !d_rank_dst))) return b200sa::set_error(B200SA_EINVAL, "destination too small");
    if (count == 0) return 0;
    cudaStream_t st = e.pick(stream);
    B200SA_CU(cudaMemcpyAsync(d_idx_dst, e.ss.upd_idx, (size_t)count * 4, cudaMemcpyDeviceToDevice, st));
    B200SA_CU(cudaMemcpyAsync(d_rank_dst, e.ss.upd_rank, (size_t)count * 4, cudaMemcpyDeviceToDevice, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_shard_apply_updates(b200sa_ctx* ctx, const uint32_t* d_idx, const uint32_t* d_rank, int64_t count, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (count < 0 || (count > 0 && (!d_idx || !d_rank))) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    if (e.ss.stage < 1) return b200sa::set_error(B200SA_EINVAL, "no sharded sort in progress");
    if (count == 0) return 0;
    cudaStream_t st = e.pick(stream);
    B200SA_TRY(e.agg_max.ensure((size_t)count * 4 + 64));
    B200SA_TRY(e.walk.ensure((size_t)count * 4 + 64));
    B200SA_TRY(e.isa_update(d_idx, d_rank, (u32)count, e.ss.n, e.agg_max.as<u32>(), e.walk.as<u32>(), false, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_shard_peer_export(b200sa_ctx* ctx, int64_t n, uint8_t* handle_out)
{
    B200SA_NEED_CTX(ctx);
    if (n <= 0) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    return ctx->eng.peer_export((u64)n, handle_out);
}

int b200sa_shard_peer_attach(b200sa_ctx* ctx, int part, int nparts, int shift, int64_t n, const uint8_t* handles)
{
    B200SA_NEED_CTX(ctx);
    if (n <= 0) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    return ctx->eng.peer_attach(part, nparts, shift, (u64)n, handles);
}

int b200sa_shard_peer_scatter(b200sa_ctx* ctx, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.peer_scatter(ctx->eng.pick(stream));
}

int b200sa_shard_peer_layout(b200sa_ctx* ctx, const int64_t* counts, int nparts)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.peer_layout(counts, nparts);
}

int b200sa_shard_peer_apply(b200sa_ctx* ctx, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.peer_apply(ctx->eng.pick(stream));
}

int b200sa_shard_peer_detach(b200sa_ctx* ctx)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.peer_detach();
}

int b200sa_shard_bwt(b200sa_ctx* ctx, int64_t row_begin, int64_t row_end, uint8_t* d_bwt, int64_t* out_begin, int64_t* out_end,
                     int32_t* sentinel_index_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (e.ss.stage < 2) return b200sa::set_error(B200SA_EINVAL, "no finished sharded sort");
    const int64_t n = e.ss.n;
    if (row_begin < 0 || row_end < row_begin || row_end > n + 1 || !d_bwt || !out_begin || !out_end)
        return b200sa::set_error(B200SA_EINVAL, "bad argument");
    cudaStream_t st = e.pick(stream);
    // sentinel row s = rank[0]; rows [row_begin,row_end) minus row s map to bytes [rb - (rb > s), re - (re > s))
    const u32* rank0 = (e.peer.active && e.ss.nparts > 1) ? (const u32*)e.peer.view.base[0] : (const u32*)e.rank.as<u32>();
    B200SA_CU(cudaMemcpyAsync(e.h_pinned + 9, rank0, 4, cudaMemcpyDefault, st));
    B200SA_CU(cudaStreamSynchronize(st));
    const int64_t s = e.h_pinned[9];
    const int64_t ob = row_begin - (row_begin > s ? 1 : 0), oe = row_end - (row_end > s ? 1 : 0);
    if (oe > ob) B200SA_TRY(e.bwt_rows(e.ss.d_text, (u32)n, e.ss.d_sa, (u32)ob, (u32)oe, d_bwt, st));
    *out_begin = ob;
    *out_end = oe;
    if (sentinel_index_out) *sentinel_index_out = (int32_t)s;
    if (e.profiling) B200SA_TRY(e.collect_profile());
    return 0;
}

// Stable partition of (key, value) pairs by (key >> shift) & 255 — the routing step of every all-to-all of the
// owner-sharded ISA (bucket = owning GPU).  counts_out: 256 host words.  d_vals may be NULL.
int b200sa_shard_partition(b200sa_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_vals, int64_t count, int shift,
                           uint32_t* d_keys_out, uint32_t* d_vals_out, uint32_t* counts_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (count < 0 || shift < 0 || shift > 31 || !counts_out || (count > 0 && (!d_keys || !d_keys_out || !d_vals_out)))
        return b200sa::set_error(B200SA_EINVAL, "bad argument");
    for (int i = 0; i < 256; ++i) counts_out[i] = 0;
    if (count == 0) return 0;
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.pick(stream);
    const u32 m = (u32)count;
    const u32 tiles = (u32)b200sa::div_up_u64(m, b200sa::RS_TILE);
    const size_t status_bytes = (size_t)tiles * b200sa::RS_RADIX * sizeof(u64);
    B200SA_TRY(e.sortmeta.ensure(b200sa::kSortMetaHeader + status_bytes));
    u32* ghist = e.sortmeta.as<u32>();
    u32* counters = ghist + b200sa::RS_MAX_PASSES * b200sa::RS_RADIX;
    u64* status = (u64*)((u8*)e.sortmeta.p + b200sa::kSortMetaHeader);
    B200SA_CU(cudaMemsetAsync(e.sortmeta.p, 0, b200sa::kSortMetaHeader + status_bytes, st));
    e.prof.memsets++;
    B200SA_TRY(e.phase_begin(B200SA_PH_ISA, st));
    const u32 htiles = (u32)b200sa::div_up_u64(m, b200sa::RH_THREADS * b200sa::RH_IPT);
    const u32 hgrid = htiles < (u32)(e.num_sms * 6) ? htiles : (u32)(e.num_sms * 6);
    auto kh = b200sa::k_radix_hist<u32>;
    B200SA_LAUNCH(kh, hgrid, b200sa::RH_THREADS, b200sa::rh_smem_bytes(1), st, d_keys, m, shift, 1, ghist);
    e.count_launch(B200SA_PH_ISA);
    B200SA_CU(cudaMemcpyAsync(counts_out, ghist, 256 * 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    B200SA_LAUNCH(b200sa::k_radix_scan_bins, 1, b200sa::RS_RADIX, 0, st, ghist);
    e.count_launch(B200SA_PH_ISA);
    auto kp = b200sa::k_onesweep_pass<u32, true>;
    B200SA_LAUNCH(kp, tiles, b200sa::RS_THREADS, b200sa::rs_pass_smem_bytes<u32>(), st, d_keys, d_keys_out, d_vals, d_vals_out,
                  m, shift, 0xffffffffu, (const u32*)ghist, status, counters);
    e.count_launch(B200SA_PH_ISA);
    B200SA_TRY(e.phase_end(st));
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

// Positions (suffix + h, clamped to n) whose ranks the next doubling round of this part will read.
int b200sa_shard_requests(b200sa_ctx* ctx, uint32_t* d_pos_out, int64_t capacity, int64_t* count_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (e.ss.stage != 2 || !count_out) return b200sa::set_error(B200SA_EINVAL, "no sharded sort in its ranking rounds");
    const u32 m = e.ss.m;
    *count_out = m;
    if (m == 0) return 0;
    if ((int64_t)m > capacity || !d_pos_out) return b200sa::set_error(B200SA_EINVAL, "destination too small");
    cudaStream_t st = e.pick(stream);
    const u32 grid = (u32)b200sa::div_up_u64(m, 256) < (u32)(e.num_sms * 8) ? (u32)b200sa::div_up_u64(m, 256) : (u32)(e.num_sms * 8);
    B200SA_LAUNCH(b200sa::k_make_requests, grid, 256, 0, st, (const u32*)e.idx[e.ss.act].as<u32>(), m, (u32)e.ss.h, e.ss.n, d_pos_out);
    e.count_launch(B200SA_PH_BUILD);
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

// out[j] = rank[pos[j]] — serves the lookups of other GPUs from this context's ISA shard.
int b200sa_shard_gather_ranks(b200sa_ctx* ctx, const uint32_t* d_pos, int64_t count, uint32_t* d_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (count < 0 || (count > 0 && (!d_pos || !d_out))) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    if (e.ss.stage < 1) return b200sa::set_error(B200SA_EINVAL, "no sharded sort in progress");
    if (count == 0) return 0;
    cudaStream_t st = e.pick(stream);
    const u32 c = (u32)count;
    const u32 grid = (u32)b200sa::div_up_u64(c, 256) < (u32)(e.num_sms * 8) ? (u32)b200sa::div_up_u64(c, 256) : (u32)(e.num_sms * 8);
    B200SA_LAUNCH(b200sa::k_gather_u32, grid, 256, 0, st, d_pos, c, (const u32*)e.rank.as<u32>(), d_out);
    e.count_launch(B200SA_PH_ISA);
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_unbwt_shard_build(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n, int32_t sentinel_index, int64_t* nwalkers_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n <= 0 || n > B200SA_MAX_N_INT32 || !d_bwt || !nwalkers_out) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    if (sentinel_index < 1 || (int64_t)sentinel_index > n) return b200sa::set_error(B200SA_EINVAL, "sentinel index %d outside [1, n]", sentinel_index);
    Engine& e = ctx->eng;
    B200SA_CU(cudaSetDevice(e.device));
    u32 W = 0;
    B200SA_TRY(e.unbwt_build(d_bwt, (u32)n, (u32)sentinel_index, &W, e.pick(stream)));
    *nwalkers_out = W;
    return 0;
}

int b200sa_unbwt_shard_measure(b200sa_ctx* ctx, int64_t w_begin, int64_t w_end, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (w_begin < 0 || w_end < w_begin) return b200sa::set_error(B200SA_EINVAL, "bad range");
    Engine& e = ctx->eng;
    B200SA_TRY(e.unbwt_measure((u32)w_begin, (u32)w_end, e.pick(stream)));
    B200SA_CU(cudaStreamSynchronize(e.pick(stream)));
    return 0;
}

// direction 0: copy this context's measured (length, successor) entries [w_begin,w_end) OUT to caller buffers;
// direction 1: copy a peer's entries IN
int b200sa_unbwt_shard_segments(b200sa_ctx* ctx, int direction, int64_t w_begin, int64_t w_end, uint32_t* d_len, uint32_t* d_next, void* stream)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (e.us.stage < 1 || w_begin < 0 || w_end < w_begin || w_end > (int64_t)e.us.nwalkers || !d_len || !d_next)
        return b200sa::set_error(B200SA_EINVAL, "bad state or argument");
    const size_t W = e.us.nwalkers, cnt = (size_t)(w_end - w_begin);
    if (cnt == 0) return 0;
    u32* nx0 = e.walk.as<u32>() + w_begin;
    u32* ds0 = e.walk.as<u32>() + 2 * W + w_begin;
    cudaStream_t st = e.pick(stream);
    if (direction == 0) {
        B200SA_CU(cudaMemcpyAsync(d_len, ds0, cnt * 4, cudaMemcpyDeviceToDevice, st));
        B200SA_CU(cudaMemcpyAsync(d_next, nx0, cnt * 4, cudaMemcpyDeviceToDevice, st));
    } else {
        B200SA_CU(cudaMemcpyAsync(ds0, d_len, cnt * 4, cudaMemcpyDeviceToDevice, st));
        B200SA_CU(cudaMemcpyAsync(nx0, d_next, cnt * 4, cudaMemcpyDeviceToDevice, st));
    }
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_unbwt_shard_finish(b200sa_ctx* ctx, int64_t w_begin, int64_t w_end, uint8_t* d_text_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (w_begin < 0 || w_end < w_begin || !d_text_out) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    Engine& e = ctx->eng;
    B200SA_TRY(e.unbwt_finish((u32)w_begin, (u32)w_end, d_text_out, e.pick(stream)));
    if (e.profiling) B200SA_TRY(e.collect_profile());
    return 0;
}

// ---- instrumentation ---------------------------------------------------------------------------

int b200sa_set_profiling(b200sa_ctx* ctx, int enabled)
{
    B200SA_NEED_CTX(ctx);
    ctx->eng.profiling = enabled != 0;
    return 0;
}

int b200sa_profile_reset(b200sa_ctx* ctx)
{
    B200SA_NEED_CTX(ctx);
    memset(&ctx->eng.prof, 0, sizeof(ctx->eng.prof));
    return 0;
}

int b200sa_profile_get(b200sa_ctx* ctx, b200sa_profile* out)
{
    B200SA_NEED_CTX(ctx);
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out");
    B200SA_TRY(ctx->eng.collect_profile());
    *out = ctx->eng.prof;
    return 0;
}

uint64_t b200sa_launch_count(b200sa_ctx* ctx) { return ctx ? ctx->eng.total_launches : 0; }

#ifdef B200SA_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int b200sa_debug_phase_cycles(unsigned long long* out8, int reset)
{
    if (cudaMemcpyFromSymbol(out8, b200sa::g_phase_cycles, 64) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(b200sa::g_phase_cycles, z, 64); }
    return 0;
}
#endif

// ---- building blocks ---------------------------------------------------------------------------

// round trip of the two rerank look-back descriptors (sa_kernels.cuh) for one (kept, kept heads, last head) triple:
// out = {kept, kept heads, 1 + last head slot, flags of A, flags of B}; lets the CPU tier check counts >= 2^31
int b200sa_debug_rerank_descriptor(uint32_t kept, uint32_t kheads, uint32_t last_head1, uint64_t* out5)
{
    if (!out5 || kheads > 0x7fffffffu) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    const u64 a = b200sa::rr_pack_a(b200sa::RR_FLAG_INCLUSIVE, kept, kheads);
    const u64 b = b200sa::rr_pack_b(b200sa::RR_FLAG_INCLUSIVE, kept, last_head1);
    out5[0] = b200sa::rr_ab_kept(a, b);
    out5[1] = b200sa::rr_a_kheads(a);
    out5[2] = b200sa::rr_b_last_head1(b);
    out5[3] = a >> 62;
    out5[4] = b >> 62;
    return 0;
}

int b200sa_radix_sort_pairs_dev(b200sa_ctx* ctx, uint64_t* d_keys, uint64_t* d_keys_alt, uint32_t* d_vals, uint32_t* d_vals_alt,
                                int64_t m, int begin_bit, int end_bit, int* result_in_alt, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (m < 0 || m > 0xfffffffeLL - b200sa::RS_TILE) return b200sa::set_error(B200SA_EINVAL, "m out of range");
    if (begin_bit < 0 || end_bit > 64 || !d_keys || !d_keys_alt || !d_vals_alt || !result_in_alt)
        return b200sa::set_error(B200SA_EINVAL, "bad argument");
    Engine& e = ctx->eng;
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.pick(stream);
    u64* k2[2] = {d_keys, d_keys_alt};
    // with generated values the first pass reads no value array; later passes ping-pong between
    // d_vals_alt and a scratch array on the input side
    u32* side0_vals = d_vals;
    if (!d_vals) {
        B200SA_TRY(e.idx[0].ensure((size_t)m * 4 + 64));
        side0_vals = e.idx[0].as<u32>();
    }
    u32* v2[2] = {side0_vals, d_vals_alt};
    int side = 0;
    B200SA_TRY(e.radix_sort_pairs(k2, v2, d_vals == nullptr, (u32)m, begin_bit, end_bit, &side, st));
    B200SA_CU(cudaStreamSynchronize(st));
    if (!d_vals && side == 0 && m > 0) {
        // sorted values ended up in the scratch array: hand them back through d_vals_alt
        B200SA_CU(cudaMemcpyAsync(d_vals_alt, side0_vals, (size_t)m * 4, cudaMemcpyDeviceToDevice, st));
        B200SA_CU(cudaStreamSynchronize(st));
    }
    *result_in_alt = side;
    if (e.profiling) B200SA_TRY(e.collect_profile());
    return 0;
}

}  // extern "C"
