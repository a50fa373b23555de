// c_abi_group.inl — the multi-GPU part of the C ABI of include/b200sa.h: b200sa_comm_*, b200sa_shard_sort / _unbwt,
// b200sa_group_* (one host thread + one context per device) and the context-free b200sa_*_gpus calls.
// Included by b200sa.cu right after c_abi.inl (one translation unit).

#include <map>

extern "C" {

// ---- one text over several GPUs, driven from C++ (engine_shard.inl, comm.cuh) ------------------------

struct b200sa_comm {
    b200sa::Comm* c;
};

int b200sa_comm_create_local(b200sa_comm** out, int nranks)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    b200sa::Comm* cs[b200sa::kMaxPeers];
    B200SA_TRY(b200sa::comm_create_local(cs, nranks));
    for (int r = 0; r < nranks; ++r) out[r] = new b200sa_comm{cs[r]};
    return 0;
}

int b200sa_comm_create_shm(b200sa_comm** out, const char* name, int rank, int nranks)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    *out = nullptr;
    b200sa::Comm* c = nullptr;
    B200SA_TRY(b200sa::comm_create_shm(&c, name, rank, nranks));
    *out = new b200sa_comm{c};
    return 0;
}

void b200sa_comm_destroy(b200sa_comm* comm)
{
    if (!comm) return;
    b200sa::comm_destroy(comm->c);
    delete comm;
}

int b200sa_comm_set_timeout_ms(b200sa_comm* comm, int timeout_ms)
{
    if (!comm || timeout_ms <= 0) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    comm->c->timeout_ms = timeout_ms;
    return 0;
}

int b200sa_comm_barrier(b200sa_comm* comm)
{
    if (!comm) return b200sa::set_error(B200SA_EINVAL, "null comm");
    return comm->c->barrier();
}

int b200sa_comm_allreduce_sum(b200sa_comm* comm, int64_t value, int64_t* sum_out)
{
    if (!comm || !sum_out) return b200sa::set_error(B200SA_EINVAL, "null pointer");
    return comm->c->allreduce_sum(value, sum_out);
}

int b200sa_shard_sort(b200sa_ctx* ctx, b200sa_comm* comm, const uint8_t* d_text, int64_t n, int32_t* d_sa, uint8_t* d_bwt,
                      int64_t* info_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (!comm || !info_out) return b200sa::set_error(B200SA_EINVAL, "null pointer");
    Engine::ShardInfo info;
    const int rc = ctx->eng.sharded_sort(*comm->c, d_text, n, d_sa, d_bwt, &info, ctx->eng.pick(stream));
    info_out[0] = info.row_begin; info_out[1] = info.row_end; info_out[2] = info.out_begin; info_out[3] = info.out_end;
    info_out[4] = info.sentinel; info_out[5] = info.rounds; info_out[6] = info.sent_bytes; info_out[7] = info.n_local;
    return rc;
}

int b200sa_shard_unbwt(b200sa_ctx* ctx, b200sa_comm* comm, const uint8_t* d_bwt, int64_t n, int32_t sentinel_index, uint8_t* d_text_out,
                       int gather_all, int64_t* slice_begin_out, int64_t* slice_end_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (!comm) return b200sa::set_error(B200SA_EINVAL, "null comm");
    if (n > 0 && d_bwt == d_text_out) return b200sa::set_error(B200SA_EINVAL, "d_text_out must not alias d_bwt");
    int64_t b = 0, e = 0;
    const int rc = ctx->eng.sharded_unbwt(*comm->c, d_bwt, n, sentinel_index, d_text_out, gather_all != 0, &b, &e, ctx->eng.pick(stream));
    if (slice_begin_out) *slice_begin_out = b;
    if (slice_end_out) *slice_end_out = e;
    return rc;
}

}  // extern "C"

// A group = G contexts (one per listed device; a device may be listed more than once) + one comm; every call runs one
// host thread per context.  This is what the facade uses for MSUFSORT_NUM_GPUS > 1 and what a C or C++ caller shards with.
struct b200sa_group {
    std::vector<b200sa_ctx*> ctxs;
    std::vector<b200sa_comm*> comms;
    std::mutex mu;  // one job at a time, like one msufsort object
    // what the last sharded sort left resident on the contexts (Engine::sa_cache with sharded = true): every context holds the
    // whole text and its rows of the suffix array.  A following call that brings the same bytes — the drop-in sequence
    // make_suffix_array, forward_burrows_wheeler_transform — reuses that sort, as the single-GPU entry points do.
    std::vector<Engine::ShardInfo> resident;

    // Upload of a host buffer every rank needs in full: rank r moves only its 1/G slice over its own PCIe link, then pulls the
    // other slices from its peers' HBM over NVLink (G x less host-memory and PCIe traffic than G full uploads).  `which`
    // selects the per-context destination buffer (it has been sized by the caller).
    typedef b200sa::DevBuf& (*BufOf)(Engine&);
    int upload_shared(int r, const uint8_t* host, int64_t n, BufOf which)
    {
        const int G = (int)ctxs.size();
        Engine& e = ctxs[(size_t)r]->eng;
        cudaStream_t st = e.own_stream;
        auto lo = [&](int g) { return (size_t)((unsigned __int128)n * (unsigned)g / (unsigned)G) & ~(size_t)15; };
        auto hi = [&](int g) { return g == G - 1 ? (size_t)n : lo(g + 1); };
        u8* mine = which(e).as<u8>();
        if (hi(r) > lo(r)) B200SA_TRY(e.copy_in(mine + lo(r), host + lo(r), hi(r) - lo(r), st));
        B200SA_CU(cudaStreamSynchronize(st));
        B200SA_TRY(comms[(size_t)r]->c->barrier());  // every slice is in its owner's HBM
        for (int k = 1; k < G; ++k) {
            const int g = (r + k) % G;               // start with the right-hand neighbour: G readers on G different sources
            if (hi(g) > lo(g))
                B200SA_CU(cudaMemcpyAsync(mine + lo(g), which(ctxs[(size_t)g]->eng).as<u8>() + lo(g), hi(g) - lo(g), cudaMemcpyDefault, st));
        }
        B200SA_CU(cudaStreamSynchronize(st));
        return comms[(size_t)r]->c->barrier();       // nobody's buffer is overwritten (next call) while a peer still reads it
    }

    template <typename F> int run(F&& per_rank)
    {
        const int G = (int)ctxs.size();
        std::vector<int> rc((size_t)G, 0);
        std::vector<std::string> msg((size_t)G);
        auto body = [&](int g) {
            rc[(size_t)g] = per_rank(g);
            if (rc[(size_t)g]) { msg[(size_t)g] = b200sa_last_error(); comms[(size_t)g]->c->raise_error(); }
        };
        std::vector<std::thread> th;
        for (int g = 1; g < G; ++g) th.emplace_back(body, g);
        body(0);
        for (auto& t : th) t.join();
        int first = 0;
        // report the root cause, not the ECOMM of the ranks that were waiting for the failed one
        for (int g = 0; g < G; ++g)
            if (rc[(size_t)g] && rc[(size_t)g] != B200SA_ECOMM && !first) { first = rc[(size_t)g]; b200sa::set_error(first, "%s", msg[(size_t)g].c_str()); }
        for (int g = 0; g < G; ++g)
            if (rc[(size_t)g] && !first) { first = rc[(size_t)g]; b200sa::set_error(first, "%s", msg[(size_t)g].c_str()); }
        if (first) {
            // leave the comm usable for the next call: all threads are gone, so the barrier words can be reset
            b200sa::CommShared* sh = comms[0]->c->sh;
            sh->error.store(0u); sh->arrive.store(0u);
            for (auto* c : comms) c->c->seq = 0;
        }
        return first;
    }
};

extern "C" {

int b200sa_group_create(b200sa_group** out, const int* devices, int count)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    *out = nullptr;
    if (count < 1 || count > b200sa::kMaxPeers || !devices) return b200sa::set_error(B200SA_EINVAL, "between 1 and %d devices", b200sa::kMaxPeers);
    b200sa_group* g = new (std::nothrow) b200sa_group();
    if (!g) return b200sa::set_error(B200SA_ENOMEM, "out of host memory");
    for (int i = 0; i < count; ++i) {
        b200sa_ctx* c = nullptr;
        const int rc = b200sa_create(&c, devices[i]);
        if (rc != 0) { for (auto* q : g->ctxs) b200sa_destroy(q); delete g; return rc; }
        g->ctxs.push_back(c);
    }
#ifndef B200SA_EMU
    // the contexts read each other's HBM (shared uploads, the sharded ISA): enable the NVLink peer mappings once
    for (int i = 0; i < count; ++i)
        for (int j = 0; j < count; ++j)
            if (devices[i] != devices[j]) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                if (can && cudaSetDevice(devices[i]) == cudaSuccess) cudaDeviceEnablePeerAccess(devices[j], 0);
                cudaGetLastError();  // already enabled is fine
            }
#endif
    g->comms.resize((size_t)count);
    const int rc = b200sa_comm_create_local(g->comms.data(), count);
    if (rc != 0) { for (auto* q : g->ctxs) b200sa_destroy(q); delete g; return rc; }
    *out = g;
    return 0;
}

void b200sa_group_destroy(b200sa_group* g)
{
    if (!g) return;
    for (auto* c : g->ctxs) b200sa_destroy(c);
    for (auto* c : g->comms) b200sa_comm_destroy(c);
    delete g;
}

int b200sa_group_size(b200sa_group* g) { return g ? (int)g->ctxs.size() : 0; }

b200sa_ctx* b200sa_group_context(b200sa_group* g, int rank)
{
    return (g && rank >= 0 && rank < (int)g->ctxs.size()) ? g->ctxs[(size_t)rank] : nullptr;
}

int b200sa_group_suffix_array_bwt(b200sa_group* g, const uint8_t* text, int64_t n, int32_t* sa_out, uint8_t* bwt_out, int32_t* sentinel_index_out)
{
    if (!g) return b200sa::set_error(B200SA_EINVAL, "null group");
    if (n < 0 || n > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n);
    if (n > 0 && !text) return b200sa::set_error(B200SA_EINVAL, "null text");
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->ctxs.size();
    // texts too small to give every GPU work (or a single-GPU group) take the single-GPU path on the first context
    if (G == 1 || n < (int64_t)G * 4096) return b200sa_suffix_array_bwt(g->ctxs[0], text, n, sa_out, bwt_out, sentinel_index_out);
    const bool want_bwt = bwt_out != nullptr || sentinel_index_out != nullptr;
    // is the result of the last sharded sort still resident on every context, for a text of this size?
    bool had_resident = want_bwt && g->resident.size() == (size_t)G;
    for (auto* c : g->ctxs) had_resident = had_resident && c->eng.sa_cache.valid && c->eng.sa_cache.sharded && c->eng.sa_cache.n == (u64)n;
    std::vector<Engine::ShardInfo> infos((size_t)G);
    const int rc = g->run([&](int r) -> int {
        Engine& e = g->ctxs[(size_t)r]->eng;
        e.sa_cache.valid = false;
        B200SA_CU(cudaSetDevice(e.device));
        cudaStream_t st = e.own_stream;
        B200SA_TRY(e.text_ws.ensure((size_t)n + 64));
        B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
        if (want_bwt) B200SA_TRY(e.bwt_ws.ensure((size_t)n + 64));
        Engine::ShardInfo& info = infos[(size_t)r];
        bool reuse = false;
        if (had_resident) {
            // the text is uploaded next to the resident one (every GPU one slice, the rest from its peers) and compared on the
            // device; all ranks must agree before any of them reuses its rows
            B200SA_TRY(e.keys[1].ensure((size_t)n + 64));
            B200SA_TRY(e.misc.ensure(8192));
            B200SA_TRY(g->upload_shared(r, text, n, [](Engine& x) -> b200sa::DevBuf& { return x.keys[1]; }));
            u32* d_diff = e.misc.as<u32>() + 548;
            B200SA_CU(cudaMemsetAsync(d_diff, 0, 4, st));
            B200SA_LAUNCH(b200sa::k_bytes_differ, (u32)(e.num_sms * 8), 256, 0, st, (const u8*)e.text_ws.as<u8>(), (const u8*)e.keys[1].as<u8>(), (u64)n, d_diff);
            e.count_launch(B200SA_PH_ALPHABET);
            B200SA_CU(cudaMemcpyAsync(e.h_pinned + 25, d_diff, 4, cudaMemcpyDeviceToHost, st));
            B200SA_CU(cudaStreamSynchronize(st));
            i64 differing = 0;
            B200SA_TRY(g->comms[(size_t)r]->c->allreduce_sum((i64)(e.h_pinned[25] != 0), &differing));
            reuse = differing == 0;
            if (reuse) {
                info = g->resident[(size_t)r];
                if (info.out_end > info.out_begin)
                    B200SA_TRY(e.bwt_rows(e.text_ws.as<u8>(), (u32)n, e.sa_ws.as<i32>(), (u32)info.out_begin, (u32)info.out_end, e.bwt_ws.as<u8>(), st));
            } else {
                // another text of the same size: it becomes the resident one (keys[1] is sort workspace; the peers' pulls from it
                // ended inside upload_shared) and is sorted below
                B200SA_CU(cudaMemcpyAsync(e.text_ws.p, e.keys[1].p, (size_t)n, cudaMemcpyDeviceToDevice, st));
            }
        } else {
            // every GPU uploads one slice of the text and pulls the rest from its peers; the results leave as disjoint slices
            B200SA_TRY(g->upload_shared(r, text, n, [](Engine& x) -> b200sa::DevBuf& { return x.text_ws; }));
        }
        if (!reuse)
            B200SA_TRY(e.sharded_sort(*g->comms[(size_t)r]->c, e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), want_bwt ? e.bwt_ws.as<u8>() : nullptr, &info, st));
        if (sa_out && info.row_end > info.row_begin)
            B200SA_TRY(e.copy_out(sa_out + info.row_begin, e.sa_ws.as<i32>() + info.row_begin, (size_t)(info.row_end - info.row_begin) * 4, st));
        // in-place callers pass bwt_out == text: every rank has finished reading the text (the sort's barriers) before any
        // rank gets here
        if (bwt_out && info.out_end > info.out_begin)
            B200SA_TRY(e.copy_out(bwt_out + info.out_begin, e.bwt_ws.as<u8>() + info.out_begin, (size_t)(info.out_end - info.out_begin), st));
        B200SA_CU(cudaStreamSynchronize(st));
        return 0;
    });
    if (rc == 0) {
        if (sentinel_index_out) *sentinel_index_out = (int32_t)infos[0].sentinel;
        // the text and every rank's rows stay resident for a following call with the same bytes
        g->resident = infos;
        for (auto* c : g->ctxs) { c->eng.sa_cache.valid = true; c->eng.sa_cache.sharded = true; c->eng.sa_cache.n = (u64)n; c->eng.sa_cache.sentinel = infos[0].sentinel; }
    } else {
        g->resident.clear();
    }
    return rc;
}

int b200sa_group_suffix_array(b200sa_group* g, const uint8_t* text, int64_t n, int32_t* sa_out)
{
    if (!sa_out) return b200sa::set_error(B200SA_EINVAL, "null sa_out");
    return b200sa_group_suffix_array_bwt(g, text, n, sa_out, nullptr, nullptr);
}

int b200sa_group_bwt(b200sa_group* g, uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out)
{
    if (!sentinel_index_out) return b200sa::set_error(B200SA_EINVAL, "null sentinel_index_out");
    return b200sa_group_suffix_array_bwt(g, text_inout, n, nullptr, text_inout, sentinel_index_out);
}

int b200sa_group_unbwt(b200sa_group* g, uint8_t* bwt_inout, int64_t n, int32_t sentinel_index)
{
    if (!g) return b200sa::set_error(B200SA_EINVAL, "null group");
    if (n < 0 || n > B200SA_MAX_N_INT32) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, 2^31-2]", (long long)n);
    if (n == 0) return 0;
    if (!bwt_inout) return b200sa::set_error(B200SA_EINVAL, "null buffer");
    if (sentinel_index < 1 || (int64_t)sentinel_index > n) return b200sa::set_error(B200SA_EINVAL, "sentinel index %d outside [1, n]", sentinel_index);
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->ctxs.size();
    if (G == 1 || n < (int64_t)G * 4096) return b200sa_unbwt(g->ctxs[0], bwt_inout, n, sentinel_index);
    // phase 1: every rank decodes; nothing is written to the caller's buffer until all ranks have accepted the input
    std::vector<int64_t> lo((size_t)G, 0), hi((size_t)G, 0);
    int rc = g->run([&](int r) -> int {
        Engine& e = g->ctxs[(size_t)r]->eng;
        e.sa_cache.valid = false;
        B200SA_CU(cudaSetDevice(e.device));
        cudaStream_t st = e.own_stream;
        B200SA_TRY(e.bwt_ws.ensure((size_t)n + 64));
        B200SA_TRY(g->upload_shared(r, bwt_inout, n, [](Engine& x) -> b200sa::DevBuf& { return x.bwt_ws; }));
        return e.sharded_unbwt(*g->comms[(size_t)r]->c, e.bwt_ws.as<u8>(), n, sentinel_index, nullptr, false, &lo[(size_t)r], &hi[(size_t)r], st);
    });
    if (rc) return rc;
    // phase 2: the slices of the text leave over all PCIe links
    return g->run([&](int r) -> int {
        Engine& e = g->ctxs[(size_t)r]->eng;
        B200SA_CU(cudaSetDevice(e.device));
        cudaStream_t st = e.own_stream;
        if (hi[(size_t)r] > lo[(size_t)r])
            B200SA_TRY(e.copy_out(bwt_inout + lo[(size_t)r], e.peer_out.as<u8>() + lo[(size_t)r], (size_t)(hi[(size_t)r] - lo[(size_t)r]), st));
        B200SA_CU(cudaStreamSynchronize(st));
        return 0;
    });
}

}  // extern "C"

// ---- batches of independent blocks over the GPUs of a group ----------------------------------------
// Blocks are independent, so nothing is exchanged: the packed batch is cut into G contiguous runs of blocks of about equal
// size, every GPU transforms its run with the single-GPU batch path (one launch sequence per GPU, engine_batch.inl) and the
// results land at the blocks' own places in the caller's buffers.  Each run must fit the 32-bit batch limit, so a group takes
// batches up to G times as large as one context.

static void split_blocks(const int64_t* offsets, int64_t count, int G, std::vector<int64_t>& cut)
{
    cut.assign((size_t)G + 1, count);
    cut[0] = 0;
    const unsigned __int128 total = (unsigned __int128)(offsets[count] + count);  // one separator slot per block, as the sort sees it
    int64_t b = 0;
    for (int g = 1; g < G; ++g) {
        const int64_t target = (int64_t)(total * (unsigned)g / (unsigned)G);
        while (b < count && offsets[b] + b < target) ++b;
        cut[(size_t)g] = b;
    }
}

static int check_block_table(const int64_t* offsets, int64_t count)
{
    if (count < 0 || (count > 0 && !offsets)) return b200sa::set_error(B200SA_EINVAL, "bad block table");
    if (count > 0 && offsets[0] != 0) return b200sa::set_error(B200SA_EINVAL, "offsets[0] must be 0");
    for (int64_t b = 0; b < count; ++b)
        if (offsets[b + 1] < offsets[b]) return b200sa::set_error(B200SA_EINVAL, "offsets must not decrease (block %lld)", (long long)b);
    return 0;
}

// the block table of run [b0, b1) with offsets relative to the run's first byte
static std::vector<int64_t> run_offsets(const int64_t* offsets, int64_t b0, int64_t b1)
{
    std::vector<int64_t> sub((size_t)(b1 - b0) + 1);
    for (int64_t b = b0; b <= b1; ++b) sub[(size_t)(b - b0)] = offsets[b] - offsets[b0];
    return sub;
}

static int group_batch(b200sa_group* g, const uint8_t* blocks, const int64_t* offsets, int64_t count, uint8_t* bwt_out, int32_t* sa_out,
                       int32_t* sentinel_index_out)
{
    if (!g) return b200sa::set_error(B200SA_EINVAL, "null group");
    B200SA_TRY(check_block_table(offsets, count));
    if (count == 0) return 0;
    if (offsets[count] > 0 && !blocks) return b200sa::set_error(B200SA_EINVAL, "null blocks");
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->ctxs.size();
    if (G == 1 || count < (int64_t)G || offsets[count] < (int64_t)G * 4096)
        return batch_host(g->ctxs[0], blocks, offsets, count, bwt_out, sa_out, sentinel_index_out);
    std::vector<int64_t> cut;
    split_blocks(offsets, count, G, cut);
    return g->run([&](int r) -> int {
        const int64_t b0 = cut[(size_t)r], b1 = cut[(size_t)r + 1];
        g->ctxs[(size_t)r]->eng.sa_cache.valid = false;
        if (b1 <= b0) return 0;
        const std::vector<int64_t> sub = run_offsets(offsets, b0, b1);
        const int64_t base = offsets[b0];
        return batch_host(g->ctxs[(size_t)r], blocks ? blocks + base : nullptr, sub.data(), b1 - b0, bwt_out ? bwt_out + base : nullptr,
                          sa_out ? sa_out + base + b0 : nullptr, sentinel_index_out ? sentinel_index_out + b0 : nullptr);
    });
}

extern "C" {

int b200sa_group_suffix_array_batch(b200sa_group* g, const uint8_t* blocks, const int64_t* offsets, int64_t count, int32_t* sa_out)
{
    if (count > 0 && !sa_out) return b200sa::set_error(B200SA_EINVAL, "null sa_out");
    return group_batch(g, blocks, offsets, count, nullptr, sa_out, nullptr);
}

int b200sa_group_bwt_batch(b200sa_group* g, uint8_t* blocks_inout, const int64_t* offsets, int64_t count, int32_t* sentinel_index_out)
{
    if (count > 0 && !sentinel_index_out) return b200sa::set_error(B200SA_EINVAL, "null sentinel_index_out");
    return group_batch(g, blocks_inout, offsets, count, blocks_inout, nullptr, sentinel_index_out);
}

int b200sa_group_unbwt_batch(b200sa_group* g, uint8_t* blocks_inout, const int64_t* offsets, int64_t count, const int32_t* sentinel_index)
{
    if (!g) return b200sa::set_error(B200SA_EINVAL, "null group");
    B200SA_TRY(check_block_table(offsets, count));
    if (count == 0) return 0;
    if (!sentinel_index) return b200sa::set_error(B200SA_EINVAL, "null sentinel_index");
    std::lock_guard<std::mutex> lk(g->mu);
    const int G = (int)g->ctxs.size();
    if (G == 1 || count < (int64_t)G || offsets[count] < (int64_t)G * 4096)
        return b200sa_unbwt_batch(g->ctxs[0], blocks_inout, offsets, count, sentinel_index);
    std::vector<int64_t> cut;
    split_blocks(offsets, count, G, cut);
    // step 1: every GPU decodes its run; nothing is written to the caller's buffer until all runs have been accepted
    int rc = g->run([&](int r) -> int {
        const int64_t b0 = cut[(size_t)r], b1 = cut[(size_t)r + 1];
        g->ctxs[(size_t)r]->eng.sa_cache.valid = false;
        if (b1 <= b0) return 0;
        const std::vector<int64_t> sub = run_offsets(offsets, b0, b1);
        return unbwt_batch_decode(g->ctxs[(size_t)r], blocks_inout ? blocks_inout + offsets[b0] : nullptr, sub.data(), b1 - b0, sentinel_index + b0,
                                  /*synchronise=*/true);
    });
    if (rc) return rc;
    // step 2: the decoded runs leave over all PCIe links
    return g->run([&](int r) -> int {
        const int64_t b0 = cut[(size_t)r], b1 = cut[(size_t)r + 1];
        if (b1 <= b0) return 0;
        return unbwt_batch_copy_out(g->ctxs[(size_t)r], blocks_inout + offsets[b0], offsets[b1] - offsets[b0]);
    });
}

}  // extern "C"

// ---- context-free calls with a GPU count (the shape SURVEY.md §8(b) proposed for the C ABI) --------------------
// b200sa_*_gpus(..., num_gpus): GPUs 0 .. num_gpus-1 (num_gpus <= 0: all GPUs present); the group behind each count is
// created on first use and kept for the life of the process.

// returns 0 and the group, or the status of what failed (the error text is already set)
static int shared_group(int num_gpus, b200sa_group** out)
{
    static std::mutex mu;
    static std::map<int, b200sa_group*> groups;
    *out = nullptr;
    const int present = b200sa_device_count();
    if (present <= 0) return b200sa::set_error(B200SA_ENODEVICE, "no CUDA device available; this library has no CPU fallback");
    if (num_gpus <= 0 || num_gpus > present) num_gpus = present;
    if (num_gpus > b200sa::kMaxPeers) num_gpus = b200sa::kMaxPeers;
    std::lock_guard<std::mutex> lk(mu);
    auto it = groups.find(num_gpus);
    if (it == groups.end()) {
        std::vector<int> devices;
        for (int g = 0; g < num_gpus; ++g) devices.push_back(g);
        b200sa_group* grp = nullptr;
        B200SA_TRY(b200sa_group_create(&grp, devices.data(), num_gpus));
        it = groups.emplace(num_gpus, grp).first;
    }
    *out = it->second;
    return 0;
}

extern "C" {

int b200sa_suffix_array_gpus(const uint8_t* text, int64_t n, int32_t* sa_out, int num_gpus)
{
    b200sa_group* g = nullptr;
    B200SA_TRY(shared_group(num_gpus, &g));
    return b200sa_group_suffix_array(g, text, n, sa_out);
}

int b200sa_bwt_gpus(uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out, int num_gpus)
{
    b200sa_group* g = nullptr;
    B200SA_TRY(shared_group(num_gpus, &g));
    return b200sa_group_bwt(g, text_inout, n, sentinel_index_out);
}

int b200sa_unbwt_gpus(uint8_t* bwt_inout, int64_t n, int32_t sentinel_index, int num_gpus)
{
    b200sa_group* g = nullptr;
    B200SA_TRY(shared_group(num_gpus, &g));
    return b200sa_group_unbwt(g, bwt_inout, n, sentinel_index);
}

}  // extern "C"
