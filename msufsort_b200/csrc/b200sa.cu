// b200sa.cu — engine implementation and the C ABI declared in include/b200sa.h.
//
// Build (product):  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//                        -Xcompiler -fPIC -o msufsort_b200/lib/libb200sa.so b200sa.cu
// There is no CPU path in this file: every entry point needs a CUDA device.
#include "engine.cuh"

#ifdef B200SA_EMU_ASAN
#include <sanitizer/asan_interface.h>
#endif
#include <mutex>
#include <new>
#include <thread>
#include <atomic>

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
#ifdef B200SA_EMU_ASAN
    bytes = (bytes + 3) & ~(size_t)3;  // whole words, as every real device allocation
    // AddressSanitizer build of the emulator (tests only): the bytes past the size asked for LAST are poisoned, so a kernel that
    // runs over the logical end of a reused buffer is reported although the allocation is larger
    if (bytes <= cap) {
        __asan_unpoison_memory_region(p, cap);
        if (bytes < cap) __asan_poison_memory_region((char*)p + bytes, cap - bytes);
        return 0;
    }
    if (p) { __asan_unpoison_memory_region(p, cap); cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes;
#else
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    // round up so that slowly growing inputs do not reallocate every call
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
#endif
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
#ifdef B200SA_EMU_ASAN
    if (p) __asan_unpoison_memory_region(p, cap);
#endif
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
    B200SA_CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
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
    if (const char* e12 = getenv("B200SA_LCP_DIRECT")) lcp_direct = atoi(e12) != 0;
    if (const char* e11 = getenv("B200SA_NUM_SMS")) { const int v = atoi(e11); if (v >= 1 && v <= 1024) num_sms = v; }  // tests: small persistent grids
    {
        // staging threads for pageable host buffers: half the host cores, between 2 and 8 (measured on a 16-core box, 1 GiB
        // download into a std::vector: 4 threads 213 ms, 8 threads 192 ms per make_suffix_array call)
        const unsigned hw = std::thread::hardware_concurrency();
        copy_threads = hw >= 16 ? 8 : (hw >= 4 ? (int)(hw / 2) : 2);
    }
    if (const char* e13 = getenv("B200SA_COPY_THREADS")) copy_threads = atoi(e13);
    if (const char* e17 = getenv("B200SA_BWT_WINDOW_BYTES")) bwt_window_bytes = (size_t)strtoull(e17, nullptr, 10);
    if (const char* e18 = getenv("B200SA_BWT_MAX_PASSES")) bwt_max_passes = (u32)strtoul(e18, nullptr, 10);
    if (const char* e16 = getenv("B200SA_ISA_PULL_FRACTION")) isa_pull_fraction = (u32)strtoul(e16, nullptr, 10);
    if (const char* e6 = getenv("B200SA_UNBWT_CAP_MULT")) unbwt_cap_mult = (u32)strtoul(e6, nullptr, 10);
    if (unbwt_cap_mult < 1) unbwt_cap_mult = 1;
    if (groupsort_tiny > (u32)GS_TINY) groupsort_tiny = GS_TINY;
    if (groupsort_medium > (u32)GS_MEDIUM) groupsort_medium = GS_MEDIUM;
    // the scatter kernels use more than the default 48 KB of dynamic shared memory
    {
        auto k64 = k_onesweep_pass<u64, true>;
        auto k8 = k_onesweep_pass<u8, false>;
        auto k32 = k_onesweep_pass<u32, true>;
        auto k32u = k_onesweep_pass<u32, true, false>;
        B200SA_CU(cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u32>()));
        B200SA_CU(cudaFuncSetAttribute(k32u, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u32>()));
        B200SA_CU(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u64>()));
        B200SA_CU(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_pass_smem_bytes<u8>()));
    }
    return 0;
}

int Engine::release_workspace()
{
    sa_cache.valid = false;  // the resident text and suffix array go away
    if (peer.active) peer_detach();  // the ISA array and the inbox go away: peers must re-attach before the next sharded sort
    for (int i = 0; i < 2; ++i) { keys[i].release(); idx[i].release(); slot[i].release(); }
    gid.release(); gstart.release(); glist.release(); rank.release(); sa_ws.release(); sortmeta.release(); agg_cnt.release(); agg_max.release();
    misc.release(); text_ws.release(); bwt_ws.release(); walk.release();
    batch_text.release(); batch_meta.release(); batch_out.release();
    peer_inbox.release();
    peer_out.release();
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
#ifndef B200SA_EMU
    for (int t = 0; t < stage.threads; ++t) {
        for (int b = 0; b < 2; ++b) { cudaFreeHost(stage.buf[t][b]); cudaEventDestroy(stage.done[t][b]); }
        cudaStreamDestroy(stage.stream[t]);
    }
    if (stage.fence) cudaEventDestroy(stage.fence);
    stage = HostStage();
#endif
    if (own_stream) cudaStreamDestroy(own_stream);
    own_stream = nullptr;
    if (copy_stream) cudaStreamDestroy(copy_stream);
    copy_stream = nullptr;
}

// ---------------------------------------------------------------------------------------------
// host <-> device transfers (see engine.cuh)

#ifndef B200SA_EMU
static bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
#endif

int Engine::stage_ensure()
{
#ifndef B200SA_EMU
    if (stage.threads) return 0;
    int T = copy_threads < 1 ? 1 : (copy_threads > HostStage::kMaxThreads ? HostStage::kMaxThreads : copy_threads);
    B200SA_CU(cudaEventCreateWithFlags(&stage.fence, cudaEventDisableTiming));
    for (int t = 0; t < T; ++t) {
        B200SA_CU(cudaStreamCreateWithFlags(&stage.stream[t], cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            B200SA_CU(cudaHostAlloc(&stage.buf[t][b], stage.chunk, cudaHostAllocDefault));
            B200SA_CU(cudaEventCreateWithFlags(&stage.done[t][b], cudaEventDisableTiming));
        }
    }
    stage.threads = T;
#endif
    return 0;
}

// direction 0: host -> device, 1: device -> host.  Thread t takes chunks t, t + T, ...; per chunk the PCIe copy and the host
// memcpy run on alternating staging buffers.
#ifndef B200SA_EMU
static cudaError_t staged_worker(Engine::HostStage& hs, int device, int t, int T, int direction, char* dev, char* host, size_t bytes)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    const size_t chunk = hs.chunk;
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    cudaStream_t s = hs.stream[t];
    if (direction == 0) {
        int b = 0;
        for (size_t c = (size_t)t; c < nchunks; c += (size_t)T, b ^= 1) {
            const size_t off = c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
            if ((e = cudaEventSynchronize(hs.done[t][b])) != cudaSuccess) return e;  // the copy that last used this buffer
            memcpy(hs.buf[t][b], host + off, len);
            if ((e = cudaMemcpyAsync(dev + off, hs.buf[t][b], len, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(hs.done[t][b], s)) != cudaSuccess) return e;
        }
        return cudaStreamSynchronize(s);
    }
    // device -> host: keep one copy in flight while the previous chunk is moved out of its staging buffer
    size_t c = (size_t)t;
    int b = 0;
    if (c < nchunks) {
        const size_t off = c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
        if ((e = cudaMemcpyAsync(hs.buf[t][b], dev + off, len, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(hs.done[t][b], s)) != cudaSuccess) return e;
    }
    for (; c < nchunks; c += (size_t)T, b ^= 1) {
        const size_t nc = c + (size_t)T;
        if (nc < nchunks) {
            const size_t off = nc * chunk, len = bytes - off < chunk ? bytes - off : chunk;
            if ((e = cudaMemcpyAsync(hs.buf[t][b ^ 1], dev + off, len, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(hs.done[t][b ^ 1], s)) != cudaSuccess) return e;
        }
        if ((e = cudaEventSynchronize(hs.done[t][b])) != cudaSuccess) return e;
        const size_t off = c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
        memcpy(host + off, hs.buf[t][b], len);
    }
    return cudaSuccess;
}

static int staged_copy(Engine& e, int direction, void* dev, void* host, size_t bytes, cudaStream_t st, bool independent)
{
    B200SA_TRY(e.stage_ensure());
    Engine::HostStage& hs = e.stage;
    const int T = hs.threads;
    // the workers' streams start after everything already enqueued on st (the producer of a download, the last reader of an
    // upload target)
    if (!independent) {
        B200SA_CU(cudaEventRecord(hs.fence, st));
        for (int t = 0; t < T; ++t) B200SA_CU(cudaStreamWaitEvent(hs.stream[t], hs.fence, 0));
    }
    cudaError_t err[Engine::HostStage::kMaxThreads];
    std::thread th[Engine::HostStage::kMaxThreads];
    for (int t = 1; t < T; ++t)
        th[t] = std::thread([&, t] { err[t] = staged_worker(hs, e.device, t, T, direction, (char*)dev, (char*)host, bytes); });
    err[0] = staged_worker(hs, e.device, 0, T, direction, (char*)dev, (char*)host, bytes);
    for (int t = 1; t < T; ++t) th[t].join();
    for (int t = 0; t < T; ++t)
        if (err[t] != cudaSuccess) return set_error(B200SA_ECUDA, "staged host copy failed: %s", cudaGetErrorString(err[t]));
    // every worker synchronised its stream (uploads) or its last event (downloads): the data is in place; st needs no wait
    return 0;
}
#endif

int Engine::copy_in(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return 0;
#ifndef B200SA_EMU
    if (copy_threads > 0 && bytes >= copy_staged_min && !host_pointer_is_pinned(h_src)) return staged_copy(*this, 0, d_dst, (void*)h_src, bytes, st, false);
#endif
    B200SA_CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

int Engine::copy_out(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st, bool independent)
{
    if (bytes == 0) return 0;
#ifndef B200SA_EMU
    if (copy_threads > 0 && bytes >= copy_staged_min && !host_pointer_is_pinned(h_dst)) return staged_copy(*this, 1, (void*)d_src, h_dst, bytes, st, independent);
#endif
    B200SA_CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, independent ? copy_stream : st));
    return 0;
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
        auto kp = k_onesweep_pass<u64, true>;
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
        auto kp = k_onesweep_pass<u32, true, false>;  // a multi-split suffices: the pairs are only bucketed for the scatter
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
    // a rank of a sharded sort driven by sharded_sort() counts only its slice of the text; the counts are summed over the ranks
    const bool split_hist = shard_comm != nullptr && nparts > 1;
    const u64 h_lo = split_hist ? (u64)n * (u64)part / (u64)nparts : 0ull;
    const u64 h_hi = split_hist ? (u64)n * (u64)(part + 1) / (u64)nparts : (u64)n;
    {
        const u64 nvec = (h_hi - h_lo) / 16 + 1;
        const u32 grid = (u32)(div_up_u64(nvec, BH_THREADS) < (u64)(num_sms * 8) ? div_up_u64(nvec, BH_THREADS) : (u64)(num_sms * 8));
        B200SA_LAUNCH(k_byte_hist, grid, BH_THREADS, 0, st, d_text + h_lo, h_hi - h_lo, d_hist);
        count_launch(B200SA_PH_ALPHABET);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_ALPHABET] += h_hi - h_lo;
    u32 h_hist[256];
    B200SA_CU(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    if (split_hist) {
        static_assert(sizeof(h_hist) <= (size_t)kCommSlotBytes, "histogram fits a comm slot");
        std::vector<u32> all((size_t)256 * nparts);
        B200SA_TRY(shard_comm->allgather(h_hist, sizeof(h_hist), all.data()));
        for (int c = 0; c < 256; ++c) {
            u32 t = 0;
            for (int g = 0; g < nparts; ++g) t += all[(size_t)g * 256 + c];
            h_hist[c] = t;
        }
    }
    if (ss.batch_count) h_hist[0] -= ss.batch_count;  // the separator slots of a batch are not symbols
    ss.plan = plan_alphabet(h_hist, ss.batch_bits, max_key_bits, pack_radix);
    const AlphabetPlan& plan = ss.plan;
    B200SA_CU(cudaMemcpyAsync(d_code, plan.code, 256, cudaMemcpyHostToDevice, st));

    // ---- initial keys
    B200SA_LAUNCH(k_sa_init, 1, 32, 0, st, rank.as<u32>(), d_sa, n);
    count_launch(B200SA_PH_PACK);
    u64* k2[2] = {keys[0].as<u64>(), keys[1].as<u64>()};
    u32* v2[2] = {idx[0].as<u32>(), idx[1].as<u32>()};
    const int key_bits = plan.key_bits + ss.batch_bits;
    const u32 pk_tiles = (u32)div_up_u64(n, PK_TILE);
    const u32 pk_grid = pk_tiles < (u32)(num_sms * 4) ? pk_tiles : (u32)(num_sms * 4);
    u64* d_pow = (u64*)(misc.as<u32>() + 320);  // 66 words of u64 behind the symbol codes (mixed-radix keys)
    if (plan.radix) B200SA_CU(cudaMemcpyAsync(d_pow, plan.pow, sizeof(plan.pow), cudaMemcpyHostToDevice, st));
    int side = 0;
    u32 count = n;
    if (nparts > 1) {
        // ---- splitters from a sorted regular sample of the keys, then pack only this part's key range
        u32 nsample = n < (1u << 18) ? n : (1u << 18);
        const u32 stride = n / nsample;
        B200SA_TRY(slot[0].ensure((size_t)nsample * 8 + 64));  // the slot maps are not in use yet
        B200SA_TRY(slot[1].ensure((size_t)nsample * 8 + 64));
        u64* smp[2] = {(u64*)slot[0].p, (u64*)slot[1].p};
        B200SA_TRY(agg_max.ensure((size_t)nsample * 8 + 64));
        u32* sv[2] = {agg_max.as<u32>(), agg_max.as<u32>() + nsample};
        if (plan.radix) {
            auto ks = k_sample_keys_text<true>;
            B200SA_LAUNCH(ks, (u32)div_up_u64(nsample, 256), 256, 0, st, d_text, n, (const u8*)d_code, 0, plan.k, 0, plan.radix, nsample, stride, smp[0]);
        } else {
            auto ks = k_sample_keys_text<false>;
            B200SA_LAUNCH(ks, (u32)div_up_u64(nsample, 256), 256, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, nsample, stride, smp[0]);
        }
        count_launch(B200SA_PH_PACK);
        int sside = 0;
        B200SA_TRY(radix_sort_pairs(smp, sv, true, nsample, 0, key_bits, &sside, st));
        // only this part's two splitters travel to the host (pinned words 450..453)
        u64* h_split = (u64*)(h_pinned + 450);
        h_split[0] = 0ull;
        h_split[1] = ~0ull;
        if (part > 0) B200SA_CU(cudaMemcpyAsync(&h_split[0], smp[sside] + (size_t)((u64)part * nsample / nparts), 8, cudaMemcpyDeviceToHost, st));
        if (part < nparts - 1)
            B200SA_CU(cudaMemcpyAsync(&h_split[1], smp[sside] + (size_t)((u64)(part + 1) * nsample / nparts), 8, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        const u64 lo = h_split[0], hi = h_split[1];
        const int hi_inclusive = part == nparts - 1 ? 1 : 0;
        u32* d_count = misc.as<u32>() + 536;
        B200SA_CU(cudaMemsetAsync(d_count, 0, 4, st));
        prof.memsets++;
        B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
        if (plan.radix) {
            auto kp = k_pack_keys<true, true>;
            B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, 0, plan.k, 0, plan.radix, plan.pow[plan.k - 1], k2[1],
                          lo, hi, hi_inclusive, v2[1], d_count);
        } else {
            auto kp = k_pack_keys<false, true>;
            B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, (u64)0, k2[1],
                          lo, hi, hi_inclusive, v2[1], d_count);
        }
        count_launch(B200SA_PH_PACK);
        B200SA_TRY(phase_end(st));
        B200SA_CU(cudaGetLastError());
        B200SA_CU(cudaMemcpyAsync(h_pinned + 16, d_count, 4, cudaMemcpyDeviceToHost, st));
        B200SA_CU(cudaStreamSynchronize(st));
        count = h_pinned[16];
        prof.alg_bytes[B200SA_PH_PACK] += (u64)n + (u64)count * 12;
        // sort this part: input on side 1
        u64* kk[2] = {k2[1], k2[0]};
        u32* vv[2] = {v2[1], v2[0]};
        int rs = 0;
        B200SA_TRY(radix_sort_pairs(kk, vv, false, count, 0, key_bits, &rs, st));
        side = 1 ^ rs;
    } else {
        B200SA_TRY(phase_begin(B200SA_PH_PACK, st));
        if (plan.radix) {
            if (ss.batch_count) {
                auto kp = k_pack_keys_batch<true>;
                B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.key_bits, plan.k, 0, plan.radix, (const u64*)d_pow,
                              ss.batch_ends, ss.batch_count, k2[0]);
            } else {
                auto kp = k_pack_keys<true, false>;
                B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, 0, plan.k, 0, plan.radix, plan.pow[plan.k - 1], k2[0],
                              (u64)0, (u64)0, 0, (u32*)nullptr, (u32*)nullptr);
            }
        } else if (ss.batch_count) {
            auto kp = k_pack_keys_batch<false>;
            B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, (const u64*)nullptr,
                          ss.batch_ends, ss.batch_count, k2[0]);
        } else {
            auto kp = k_pack_keys<false, false>;
            B200SA_LAUNCH(kp, pk_grid, PK_THREADS, 0, st, d_text, n, (const u8*)d_code, plan.bits, plan.k, plan.len_bits, (u64)0, (u64)0, k2[0],
                          (u64)0, (u64)0, 0, (u32*)nullptr, (u32*)nullptr);
        }
        count_launch(B200SA_PH_PACK);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_PACK] += (u64)n * 9;
        B200SA_CU(cudaGetLastError());
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
    const bool use_peer = peer.active && peer.has_isa && ss.nparts > 1;
    if (use_peer && (peer.view.n != n || peer.nparts != ss.nparts)) return set_error(B200SA_EINVAL, "peer ISA attached for another text size / GPU count");
    LocalRank local_rank{rank.as<u32>(), n};
    PeerRank peer_rank{peer.view};
    // Sharded runs read rank[suffix + h] from the owners' HBM with 4-byte remote loads — fine while few suffixes are active,
    // but a round over huge groups (periodic / Fibonacci texts: every suffix active for ~log n rounds) measured 10 G loads/s
    // per GPU, half of the 2 GiB Fibonacci run on eight GPUs.  When this GPU is about to read more than n/12 ranks, pulling
    // the peers' shards in bulk (4 n bytes over NVLink at several hundred GB/s) and gathering locally is cheaper.  The copies
    // land in the unused part of this GPU's own full-size rank array; shards only change between rounds, behind barriers.
    bool pulled = false;
    auto pull_peer_shards = [&]() -> int {
        if (pulled) return 0;
        B200SA_TRY(phase_begin(B200SA_PH_PEER_SEND, st));
        for (int k = 1; k < peer.nparts; ++k) {
            const int g = (peer.part + k) % peer.nparts;
            const u64 lo = (u64)g << peer.view.shift, hi = ((u64)(g + 1) << peer.view.shift) < (u64)n ? ((u64)(g + 1) << peer.view.shift) : (u64)n;
            if (lo < hi) B200SA_CU(cudaMemcpyAsync(rank.as<u32>() + lo, peer.view.base[g] + lo, (size_t)(hi - lo) * 4, cudaMemcpyDefault, st));
        }
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_PEER_SEND] += (u64)n * 4;
        pulled = true;
        return 0;
    };
    const bool bulk_isa = use_peer && isa_pull_fraction > 0 && (u64)m * isa_pull_fraction > (u64)n;

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
                if (bulk_isa) B200SA_TRY(pull_peer_shards());
                for (u32 q = 0; q < nhuge; ++q) {
                    const u32 s0 = gs[2 * q], sz = gs[2 * q + 1] - s0;
                    B200SA_TRY(phase_begin(B200SA_PH_BUILD, st));
                    const u32 tiles = (u32)div_up_u64(sz, BK_THREADS * BK_IPT);
                    const u32 grid = tiles < (u32)(num_sms * 8) ? tiles : (u32)(num_sms * 8);
                    if (use_peer && !bulk_isa) {
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
        if (bulk_isa) B200SA_TRY(pull_peer_shards());
        B200SA_TRY(phase_begin(B200SA_PH_BUILD, st));
        {
            const u32 tiles = (u32)div_up_u64(m, BK_THREADS * BK_IPT);
            const u32 grid = tiles < (u32)(num_sms * 8) ? tiles : (u32)(num_sms * 8);
            if (use_peer && !bulk_isa) {
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

#include "engine_peer.inl"

// ---------------------------------------------------------------------------------------------
// forward BWT

// output bytes [o_begin, o_end) from a finished SA (+ rank[0] = sentinel row); the sentinel row lands
// in h_pinned[8]
int Engine::bwt_rows(const u8* d_text, u32 n, const i32* d_sa, u32 o_begin, u32 o_end, u8* d_bwt, cudaStream_t st, bool defer_sync)
{
    i32* d_sent = (i32*)(misc.as<u32>() + 528);
    B200SA_TRY(phase_begin(B200SA_PH_BWT, st));
    {
        const u32 groups = (u32)div_up_u64(o_end - o_begin, 4);
        const u32 want = (u32)div_up_u64(groups, BW_THREADS * BW_STEPS);
        const u32 grid = want < (u32)(num_sms * 16) ? (want ? want : 1u) : (u32)(num_sms * 16);
        // rank[0] (the sentinel row) lives on GPU 0 when the ISA is sharded in peer memory
        const u32* rank0 = (peer.active && peer.has_isa && ss.nparts > 1) ? (const u32*)peer.view.base[0] : (const u32*)rank.as<u32>();
        // texts that do not fit L2: one pass per text window (bwt_kernels.cuh); otherwise, and for unaligned outputs, one pass
        // (a pass costs about 1 ms per 2^28 rows whatever it hits, so more than bwt_max_passes windows lose against the one
        // pass of random DRAM sectors: 2^30 bytes in 8 windows 31.6 ms, in one pass 25.5 ms)
        u32 passes = bwt_window_bytes ? (u32)div_up_u64(n, bwt_window_bytes) : 1u;
        if (passes > bwt_max_passes || o_end - o_begin < 64u) passes = 1u;
        if (passes <= 1u) {
            B200SA_LAUNCH(k_bwt_gather, grid, BW_THREADS, 0, st, d_text, d_sa, rank0, o_begin, o_end, d_bwt, d_sent);
            count_launch(B200SA_PH_BWT);
        } else {
            // the windowed passes combine 32-bit words: the (up to three) bytes before the first aligned output word go through
            // the single-pass kernel
            const u32 head = (u32)((4u - (u32)(((uintptr_t)(d_bwt + o_begin)) & 3u)) & 3u);
            if (head) {
                B200SA_LAUNCH(k_bwt_gather, 1, BW_THREADS, 0, st, d_text, d_sa, rank0, o_begin, o_begin + head, d_bwt, d_sent);
                count_launch(B200SA_PH_BWT);
            }
            const u32 win = (u32)div_up_u64(n, passes);
            for (u32 p = 0; p < passes; ++p) {
                const u64 lo = (u64)p * win, hi = lo + win < (u64)n ? lo + win : (u64)n;
                B200SA_LAUNCH(k_bwt_gather_window, grid, BW_THREADS, 0, st, d_text, d_sa, rank0, o_begin + head, o_end, d_bwt, d_sent, (u32)lo, (u32)hi,
                              p == 0 ? 1 : 0);
                count_launch(B200SA_PH_BWT);
            }
            prof.alg_bytes[B200SA_PH_BWT] += (u64)(o_end - o_begin) * (4 + 2) * (passes - 1);  // suffix array re-read, output read-modify-write
        }
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_BWT] += (u64)(o_end - o_begin) * 6;
    B200SA_CU(cudaGetLastError());
    (void)n;
    if (defer_sync) return 0;
    B200SA_CU(cudaMemcpyAsync(h_pinned + 8, d_sent, sizeof(i32), cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
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

// Three steps so that a sharded run can split the walkers over GPUs: build (psi table, F-column table:
// replicated), measure (a slice of the walkers), finish (list ranking over all walkers, validity check, then
// emit the slice's bytes).
static const int kUnbwtBadWord = 544;  // misc word raised by k_unbwt_place when the input is not a BWT

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
        B200SA_LAUNCH(k_unbwt_fstart, 1, 256, 0, st, (const u32*)ghist, n, fstart, psi, s, misc.as<u32>() + kUnbwtBadWord);
        count_launch(B200SA_PH_UNBWT_BUILD);
        auto kp = k_onesweep_pass<u8, false>;
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u8>(), st, d_bwt, (u8*)nullptr, (const u32*)nullptr, psi + 1,
                      n, 0, s, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_UNBWT_BUILD);
    }
    B200SA_TRY(phase_end(st));
    // ---- walkers: every D-th row (D a power of two: seed test = mask, no mark bit in psi) plus row s
    int dshift = 6;
    while (((u64)n + 1) >> dshift > ((u64)1 << 21)) ++dshift;
    us.dshift = dshift;
    us.D = 1u << dshift;
    us.nreg = (u32)div_up_u64((u64)n + 1, us.D);
    us.nwalkers = us.nreg + ((s & (us.D - 1u)) != 0 ? 1u : 0u);
    B200SA_TRY(walk.ensure((size_t)us.nwalkers * 5 * 4 + 64));
    us.cap = (unbwt_cap_mult * us.D + 7u) & ~7u;  // window bytes per walker, 8-byte granular (64-bit stores)
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
                  (const u32*)(misc.as<u32>() + 600), w_begin, w_end, us.nreg, us.dshift, us.s, us.cap, keys[1].as<u8>(), ds0, nx0, ovf);
    count_launch(B200SA_PH_UNBWT_WALK);
    // the list ranking below overwrites the lengths: keep a copy for the placement pass
    B200SA_CU(cudaMemcpyAsync(idx[0].as<u32>() + w_begin, ds0 + w_begin, (size_t)(w_end - w_begin) * 4, cudaMemcpyDeviceToDevice, st));
    B200SA_TRY(phase_end(st));
    B200SA_CU(cudaGetLastError());
    return 0;
}

int Engine::unbwt_finish(u32 w_begin, u32 w_end, u8* d_out, cudaStream_t st, const ShardedOut* so)
{
    if (us.stage < 1 || w_begin > w_end || w_end > us.nwalkers) return set_error(B200SA_EINVAL, "unbwt_finish: bad state or range");
    const size_t W = us.nwalkers;
    u32* nx[2] = {walk.as<u32>(), walk.as<u32>() + W};
    u32* ds[2] = {walk.as<u32>() + 2 * W, walk.as<u32>() + 3 * W};
    u32* ovf = walk.as<u32>() + 4 * W;
    u32* d_bad = misc.as<u32>() + kUnbwtBadWord;
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
        const u32 start_walker = (us.s & (us.D - 1u)) ? us.nreg : (us.s >> us.dshift);
        const u32 grid = (u32)div_up_u64(w_end - w_begin, UP_THREADS / 32);
        if (so) {
            auto kp = k_unbwt_place<ShardedOut>;
            B200SA_LAUNCH(kp, grid, UP_THREADS, 0, st, (const u32*)keys[0].as<u32>(),
                          (const u32*)(misc.as<u32>() + 600), (const u32*)ds[cur], (const u32*)nx[cur], (const u32*)idx[0].as<u32>(), (const u32*)ovf,
                          (const u8*)keys[1].as<u8>(), us.cap, w_begin, w_end, us.n, start_walker, *so, d_bad);
        } else {
            auto kp = k_unbwt_place<LocalOut>;
            B200SA_LAUNCH(kp, grid, UP_THREADS, 0, st, (const u32*)keys[0].as<u32>(),
                          (const u32*)(misc.as<u32>() + 600), (const u32*)ds[cur], (const u32*)nx[cur], (const u32*)idx[0].as<u32>(), (const u32*)ovf,
                          (const u8*)keys[1].as<u8>(), us.cap, w_begin, w_end, us.n, start_walker, LocalOut{d_out}, d_bad);
        }
        count_launch(B200SA_PH_UNBWT_WALK);
    }
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_WALK] += (u64)us.n * 7;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned + 24, d_bad, 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    // the jumps consumed the measured segments: a second finish needs a new measure pass
    us.stage = 0;
    if (h_pinned[24] != 0)
        return set_error(B200SA_EINVAL, "the input is not a Burrows-Wheeler transform (its LF mapping does not form one cycle through the sentinel row); "
                                        "the output buffer holds no valid text");
    return 0;
}

int Engine::unbwt_dev(const u8* d_bwt, i64 n64, i64 sentinel, u8* d_out, cudaStream_t st, i64 max_n)
{
    if (n64 < 0 || n64 > max_n) return set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n64, (long long)max_n);
    if (n64 == 0) return 0;
    if (!d_bwt || !d_out) return set_error(B200SA_EINVAL, "null pointer");
    if (sentinel < 1 || sentinel > n64) return set_error(B200SA_EINVAL, "sentinel index %lld outside [1, n]", (long long)sentinel);
    B200SA_CU(cudaSetDevice(device));
    u32 W = 0;
    B200SA_TRY(unbwt_build(d_bwt, (u32)n64, (u32)sentinel, &W, st));
    B200SA_TRY(unbwt_measure(0, W, st));
    const int rc = unbwt_finish(0, W, d_out, st);
    if (profiling) B200SA_TRY(collect_profile());
    return rc;
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

#include "engine_lcp.inl"
#include "engine_batch.inl"
#include "engine_shard.inl"

}  // namespace b200sa

#include "c_abi.inl"
#include "c_abi_group.inl"
