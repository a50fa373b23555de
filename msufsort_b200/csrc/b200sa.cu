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

AlphabetPlan plan_alphabet(const u32* hist)
{
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
        if (cand * bits + bit_length_u64((u64)cand) <= 64) k = cand;
    a.k = k;
    a.len_bits = bit_length_u64((u64)k);
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
    B200SA_CU(cudaHostAlloc((void**)&h_pinned, 64 * sizeof(u32), cudaHostAllocDefault));
    memset(&prof, 0, sizeof(prof));
    // tuning knobs (tests lower them to drive the bucketed ISA update at small n)
    if (const char* e1 = getenv("B200SA_ISA_DIRECT_BYTES")) isa_direct_bytes = (size_t)strtoull(e1, nullptr, 10);
    if (const char* e2 = getenv("B200SA_ISA_MIN_UPDATES")) isa_min_updates = (u32)strtoul(e2, nullptr, 10);
    // the scatter kernels use more than the default 48 KB of dynamic shared memory
    {
        auto k64 = k_onesweep_pass<u64, true>;
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
    for (int i = 0; i < 2; ++i) { keys[i].release(); idx[i].release(); slot[i].release(); }
    gid.release(); rank.release(); sa_ws.release(); sortmeta.release(); agg_cnt.release(); agg_max.release();
    misc.release(); text_ws.release(); bwt_ws.release(); walk.release();
    return 0;
}

void Engine::shutdown()
{
    cudaSetDevice(device);
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
        auto kp = k_onesweep_pass<u64, true>;
        const u32* vin = (p == 0 && gen_vals) ? nullptr : vals2[side];
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u64>(), st,
                      keys2[side], keys2[side ^ 1], vin, vals2[side ^ 1], m, begin_bit + p * RS_RADIX_BITS, 0xffffffffu,
                      ghist + p * RS_RADIX, status + (size_t)p * tiles * RS_RADIX, counters + p);
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

int Engine::rerank(const u64* keys_sorted, const u32* idx_sorted, const u32* slot_in, u32 m, u32 n, i32* d_sa,
                   u32* idx_out, u32* slot_out, u64* free_keys, u32* next_m, u32* next_groups, cudaStream_t st)
{
    const u32 ntiles = (u32)div_up_u64(m, RR_TILE);
    // agg_cnt holds the two descriptor arrays followed by the tile ticket counter
    const size_t desc_bytes = (size_t)ntiles * 2 * sizeof(u64);
    B200SA_TRY(agg_cnt.ensure(desc_bytes + 64));
    u64* desc = agg_cnt.as<u64>();
    u32* ticket = (u32*)((u8*)agg_cnt.p + desc_bytes);
    u32* d_info = misc.as<u32>() + 512;  // see ensure_sa_workspace for the misc layout
    B200SA_CU(cudaMemsetAsync(agg_cnt.p, 0, desc_bytes + 64, st));
    prof.memsets++;
    // The ISA (rank[]) is updated by a bucketed scatter when it is too large to live in L2 and there
    // are enough updates to pay for the extra sweep; otherwise directly from the rerank kernel.
    const bool bucketed = ((u64)n * 4 > isa_direct_bytes) && (m >= isa_min_updates);
    u32* newrank = bucketed ? (u32*)free_keys : nullptr;          // [m]
    u32* bk_val = bucketed ? (u32*)free_keys + m : nullptr;       // [m]
    u32* bk_key = nullptr;
    if (bucketed) {
        B200SA_TRY(agg_max.ensure((size_t)m * 4 + 64));
        bk_key = agg_max.as<u32>();
    }
    B200SA_TRY(phase_begin(B200SA_PH_RERANK, st));
    B200SA_LAUNCH(k_rerank, ntiles, RR_THREADS, 0, st, keys_sorted, idx_sorted, slot_in, m, desc, ntiles, ticket,
                  rank.as<u32>(), newrank, d_sa, idx_out, slot_out, gid.as<u32>(), d_info);
    count_launch(B200SA_PH_RERANK);
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned, d_info, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    if (bucketed) {
        // one radix sweep of (suffix, new rank) pairs on the top 8 bits of the suffix index ...
        const int nbits = bit_length_u64((u64)n - 1);
        const int shift = nbits > RS_RADIX_BITS ? nbits - RS_RADIX_BITS : 0;
        const u32 tiles = (u32)div_up_u64(m, RS_TILE);
        const size_t status_bytes = (size_t)tiles * RS_RADIX * sizeof(u64);
        B200SA_TRY(sortmeta.ensure(kSortMetaHeader + status_bytes));
        u32* ghist = sortmeta.as<u32>();
        u32* counters = ghist + RS_MAX_PASSES * RS_RADIX;
        u64* status = (u64*)((u8*)sortmeta.p + kSortMetaHeader);
        B200SA_CU(cudaMemsetAsync(sortmeta.p, 0, kSortMetaHeader + status_bytes, st));
        prof.memsets++;
        B200SA_TRY(phase_begin(B200SA_PH_ISA, st));
        const u32 htiles = (u32)div_up_u64(m, RH_THREADS * RH_IPT);
        const u32 hgrid = htiles < (u32)(num_sms * 6) ? htiles : (u32)(num_sms * 6);
        auto kh = k_radix_hist<u32>;
        B200SA_LAUNCH(kh, hgrid, RH_THREADS, rh_smem_bytes(1), st, idx_sorted, m, shift, 1, ghist);
        count_launch(B200SA_PH_ISA);
        B200SA_LAUNCH(k_radix_scan_bins, 1, RS_RADIX, 0, st, ghist);
        count_launch(B200SA_PH_ISA);
        auto kp = k_onesweep_pass<u32, true>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u32>(), st, idx_sorted, bk_key, (const u32*)newrank, bk_val,
                      m, shift, 0xffffffffu, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_ISA);
        // ... then the scatter proper, now confined to an L2-resident window at any moment
        B200SA_LAUNCH(k_scatter_pairs, (u32)div_up_u64(m, SP_THREADS * SP_IPT), SP_THREADS, 0, st, (const u32*)bk_key,
                      (const u32*)bk_val, m, rank.as<u32>());
        count_launch(B200SA_PH_ISA);
        B200SA_TRY(phase_end(st));
        B200SA_CU(cudaGetLastError());
        prof.alg_bytes[B200SA_PH_ISA] += (u64)m * (4 + 8 + 8 + 8 + 4);
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
    B200SA_TRY(rank.ensure(((size_t)n + 1) * 4 + 64));
    // misc: [0..255] byte histogram u32, [256..319] symbol codes (256 bytes), [512..] round info,
    // [520..521] validator counter (u64), [528] sentinel row
    B200SA_TRY(misc.ensure(4096));
    return 0;
}

int Engine::suffix_array_dev(const u8* d_text, i64 n64, i32* d_sa, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32) return set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n64);
    if (!d_sa || (n64 > 0 && !d_text)) return set_error(B200SA_EINVAL, "null pointer");
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64;
    if (n == 0) {
        B200SA_CU(cudaMemsetAsync(d_sa, 0, sizeof(i32), st));
        B200SA_CU(cudaStreamSynchronize(st));
        return 0;
    }
    B200SA_TRY(ensure_sa_workspace(n));
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
    const AlphabetPlan plan = plan_alphabet(h_hist);
    B200SA_CU(cudaMemcpyAsync(d_code, plan.code, 256, cudaMemcpyHostToDevice, st));

    // ---- round 0: pack, sort, rank
    B200SA_LAUNCH(k_sa_init, 1, 32, 0, st, rank.as<u32>(), d_sa, n);
    count_launch(B200SA_PH_PACK);
    B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
    {
        const u32 tiles = (u32)div_up_u64(n, PK_TILE);
        const u32 grid = tiles < (u32)(num_sms * 4) ? tiles : (u32)(num_sms * 4);
        B200SA_LAUNCH(k_pack_keys, grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits,
                      keys[0].as<u64>());
        count_launch(B200SA_PH_PACK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_PACK] += (u64)n * 9;
    B200SA_CU(cudaGetLastError());

    u64* k2[2] = {keys[0].as<u64>(), keys[1].as<u64>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    int side = 0;
    B200SA_TRY(radix_sort_pairs(k2, v2, true, n, 0, plan.bits * plan.k + plan.len_bits, &side, st));
    prof.rounds++;
    prof.active_tuples += n;

    u32 m = 0, groups = 0;
    int cur_slot = 0;  // slot[cur_slot] holds the slot map of the active array
    // sorted tuples are on `side`; the compacted active array goes to the other side
    B200SA_TRY(rerank(k2[side], v2[side], nullptr, n, n, d_sa, v2[side ^ 1], slot[cur_slot].as<u32>(), k2[side ^ 1], &m, &groups, st));
    int act = side ^ 1;  // side holding the active idx array

    // ---- doubling rounds
    const int rank_bits = bit_length_u64(n);
    u64 h = (u64)plan.k;
    int guard = 0;
    while (m > 0) {
        if (++guard > 64) return set_error(B200SA_EINTERNAL, "prefix doubling did not converge (m=%u, h=%llu)", m, (unsigned long long)h);
        if (groups == 0 || h > (u64)n) return set_error(B200SA_EINTERNAL, "inconsistent round state (m=%u groups=%u h=%llu)", m, groups, (unsigned long long)h);
        const int gid_bits = groups <= 1 ? 0 : bit_length_u64((u64)groups - 1);
        B200SA_TRY(phase_begin(B200SA_PH_BUILD, st));
        {
            const u32 tiles = (u32)div_up_u64(m, BK_THREADS * BK_IPT);
            const u32 grid = tiles < (u32)(num_sms * 8) ? tiles : (u32)(num_sms * 8);
            B200SA_LAUNCH(k_build_keys, grid, BK_THREADS, 0, st, (const u32*)v2[act], (const u32*)gid.as<u32>(),
                          (const u32*)rank.as<u32>(), m, n, (u32)h, rank_bits, k2[act]);
            count_launch(B200SA_PH_BUILD);
        }
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_BUILD] += (u64)m * 20;
        u64* kk[2] = {k2[act], k2[act ^ 1]};
        u32* vv[2] = {v2[act], v2[act ^ 1]};
        int rs = 0;
        B200SA_TRY(radix_sort_pairs(kk, vv, false, m, 0, rank_bits + gid_bits, &rs, st));
        const int sorted_side = act ^ rs;
        u32 m2 = 0, g2 = 0;
        B200SA_TRY(rerank(k2[sorted_side], v2[sorted_side], slot[cur_slot].as<u32>(), m, n, d_sa, v2[sorted_side ^ 1],
                          slot[cur_slot ^ 1].as<u32>(), k2[sorted_side ^ 1], &m2, &g2, st));
        prof.rounds++;
        prof.active_tuples += m;
        act = sorted_side ^ 1;
        cur_slot ^= 1;
        m = m2;
        groups = g2;
        h *= 2;
    }
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// forward BWT

int Engine::bwt_dev(const u8* d_text, i64 n64, u8* d_bwt, i32* d_sa_or_null, i32* sentinel_host, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32) return set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n64);
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
    B200SA_TRY(suffix_array_dev(d_text, n64, d_sa, st));
    i32* d_sent = (i32*)(misc.as<u32>() + 528);
    B200SA_TRY(phase_begin(B200SA_PH_BWT, st));
    {
        const u32 groups = (u32)div_up_u64(n, 4);
        const u32 want = (u32)div_up_u64(groups, BW_THREADS * BW_STEPS);
        const u32 grid = want < (u32)(num_sms * 16) ? (want ? want : 1u) : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_bwt_gather, grid, BW_THREADS, 0, st, d_text, (const i32*)d_sa, (const u32*)rank.as<u32>(), n, d_bwt, d_sent);
        count_launch(B200SA_PH_BWT);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_BWT] += (u64)n * 6;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned + 8, d_sent, sizeof(i32), cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    if (sentinel_host) *sentinel_host = (i32)h_pinned[8];
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// inverse BWT

int Engine::unbwt_dev(const u8* d_bwt, i64 n64, i32 sentinel, u8* d_out, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32) return set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n64);
    if (n64 == 0) return 0;
    if (!d_bwt || !d_out) return set_error(B200SA_EINVAL, "null pointer");
    if (sentinel < 1 || (i64)sentinel > n64) return set_error(B200SA_EINVAL, "sentinel index %d outside [1, n]", sentinel);
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64, s = (u32)sentinel;
    B200SA_TRY(keys[0].ensure(((size_t)n + 1) * 4 + 64));
    B200SA_TRY(misc.ensure(4096));
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
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_BUILD] += (u64)n * 6;
    B200SA_CU(cudaGetLastError());

    // ---- walkers
    u32 D = (u32)div_up_u64((u64)n + 1, (u64)1 << 21);
    if (D < 64) D = 64;
    const u32 nreg = (u32)div_up_u64((u64)n + 1, D);
    const u32 nwalkers = nreg + ((s % D) != 0 ? 1u : 0u);
    B200SA_TRY(walk.ensure((size_t)nwalkers * 4 * 4 + 64));
    u32* nx[2] = {walk.as<u32>(), walk.as<u32>() + (size_t)nwalkers};
    u32* ds[2] = {walk.as<u32>() + 2 * (size_t)nwalkers, walk.as<u32>() + 3 * (size_t)nwalkers};
    B200SA_TRY(phase_begin(B200SA_PH_UNBWT_WALK, st));
    {
        const u32 g256 = (u32)div_up_u64(nwalkers, 256);
        const u32 gw = (u32)div_up_u64(nwalkers, UW_THREADS);
        B200SA_LAUNCH(k_unbwt_mark, g256, 256, 0, st, psi, nwalkers, nreg, D, s);
        count_launch(B200SA_PH_UNBWT_WALK);
        B200SA_LAUNCH(k_unbwt_measure, gw, UW_THREADS, 0, st, (const u32*)psi, nwalkers, nreg, D, s, ds[0], nx[0]);
        count_launch(B200SA_PH_UNBWT_WALK);
        int cur = 0;
        const int jumps = bit_length_u64(nwalkers);
        for (int it = 0; it < jumps; ++it) {
            B200SA_LAUNCH(k_unbwt_jump, g256, 256, 0, st, (const u32*)nx[cur], (const u32*)ds[cur], nx[cur ^ 1], ds[cur ^ 1], nwalkers);
            count_launch(B200SA_PH_UNBWT_WALK);
            cur ^= 1;
        }
        B200SA_LAUNCH(k_unbwt_emit, gw, UW_THREADS, 0, st, (const u32*)psi, (const u32*)fstart, (const u32*)ds[cur],
                      nwalkers, nreg, D, s, n, d_out);
        count_launch(B200SA_PH_UNBWT_WALK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_WALK] += (u64)n * 9;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    if (profiling) B200SA_TRY(collect_profile());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// validator

int Engine::check_sa_dev(const u8* d_text, i64 n64, const i32* d_sa, i64* bad_rows, cudaStream_t st)
{
    if (n64 < 0 || n64 > B200SA_MAX_N_INT32 || !d_sa || !bad_rows) return set_error(B200SA_EINVAL, "bad argument");
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
    return ctx->eng.bwt_dev(d_text, n, d_bwt_out, d_sa_out, sentinel_index_out, ctx->eng.pick(stream));
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
        int32_t sentinel = 0;
        B200SA_TRY(e.bwt_dev(e.text_ws.as<u8>(), n, e.bwt_ws.as<u8>(), e.sa_ws.as<i32>(), &sentinel, st));
        if (sentinel_index_out) *sentinel_index_out = sentinel;
        if (bwt_out) B200SA_CU(cudaMemcpyAsync(bwt_out, e.bwt_ws.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    } else {
        B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st));
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

// ---- building blocks ---------------------------------------------------------------------------

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
