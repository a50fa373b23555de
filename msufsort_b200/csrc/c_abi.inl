// c_abi.inl — the C ABI of include/b200sa.h: thin wrappers over Engine, the host-buffer entry points, the batch pipeline.
// Included by b200sa.cu (one translation unit: the kernels are templates / static functions in the .cuh headers).

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

// every entry point drops the resident-suffix-array cache (Engine::sa_cache); b200sa_suffix_array_bwt re-establishes it
#define B200SA_NEED_CTX(ctx) \
    if (!(ctx)) return b200sa::set_error(B200SA_EINVAL, "null context (call b200sa_create first)"); \
    (ctx)->eng.sa_cache.valid = false
// instrumentation getters touch no workspace
#define B200SA_NEED_CTX_KEEP(ctx) \
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

int b200sa_unbwt_u32_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n, int64_t sentinel_index, uint8_t* d_text_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    if (n > 0 && d_bwt == d_text_out) return b200sa::set_error(B200SA_EINVAL, "d_text_out must not alias d_bwt");
    return ctx->eng.unbwt_dev(d_bwt, n, sentinel_index, d_text_out, ctx->eng.pick(stream), B200SA_MAX_N_UINT32);
}

int b200sa_lcp_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const uint32_t* d_sa, uint32_t* d_lcp_out, void* stream)
{
    B200SA_NEED_CTX(ctx);
    return ctx->eng.lcp_dev(d_text, n, (const i32*)d_sa, (i32*)d_lcp_out, ctx->eng.pick(stream), B200SA_MAX_N_UINT32);
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

// SA and / or BWT of a host text, int32 or uint32 indices.  The text and its suffix array stay resident afterwards
// (Engine::sa_cache); when the next call brings the same bytes only the missing result is computed.
static int sa_bwt_host(b200sa_ctx* ctx, const uint8_t* text, int64_t n, void* sa_out, uint8_t* bwt_out, int64_t* sentinel_out, int64_t max_n)
{
    const bool had_cache = ctx && ctx->eng.sa_cache.valid && !ctx->eng.sa_cache.sharded && ctx->eng.sa_cache.n == (u64)n;
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > max_n) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n, (long long)max_n);
    if (n > 0 && !text) return b200sa::set_error(B200SA_EINVAL, "null text");
    if (n == 0) {
        if (b200sa_device_count() <= 0) return b200sa::set_error(B200SA_ENODEVICE, "no CUDA device available; this library has no CPU fallback");
        if (sa_out) ((int32_t*)sa_out)[0] = 0;
        if (sentinel_out) *sentinel_out = 0;
        return 0;
    }
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n + 64));
    B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
    const bool want_bwt = bwt_out != nullptr || sentinel_out != nullptr;
    if (want_bwt) B200SA_TRY(e.bwt_ws.ensure((size_t)n + 64));
    bool reuse = false, bwt_done = false;
    if (had_cache && want_bwt) {
        // The text is uploaded next to the resident one and compared on the device (copy stream).  Meanwhile the BWT pass
        // already runs from the resident text and suffix array (compute stream): if the bytes turn out to be the same — the
        // drop-in sequence make_suffix_array, forward_burrows_wheeler_transform — its result is the answer; if not it is
        // simply discarded.
        B200SA_TRY(e.misc.ensure(8192));
        B200SA_TRY(e.keys[1].ensure((size_t)n + 64));
        u8* d_new = e.keys[1].as<u8>();
        u32* d_diff = e.misc.as<u32>() + 548;
        B200SA_CU(cudaStreamSynchronize(st));
        B200SA_TRY(e.bwt_rows(e.text_ws.as<u8>(), (u32)n, e.sa_ws.as<i32>(), 0, (u32)n, e.bwt_ws.as<u8>(), st, /*defer_sync=*/true));
        cudaStream_t cs = e.copy_stream;
        B200SA_CU(cudaMemsetAsync(d_diff, 0, 4, cs));
        B200SA_TRY(e.copy_in(d_new, text, (size_t)n, cs));
        B200SA_LAUNCH(b200sa::k_bytes_differ, (u32)(e.num_sms * 8), 256, 0, cs, (const u8*)e.text_ws.as<u8>(), (const u8*)d_new, (u64)n, d_diff);
        e.count_launch(B200SA_PH_ALPHABET);
        B200SA_CU(cudaMemcpyAsync(e.h_pinned + 25, d_diff, 4, cudaMemcpyDeviceToHost, cs));
        B200SA_CU(cudaStreamSynchronize(cs));
        reuse = e.h_pinned[25] == 0;
        bwt_done = reuse;
        if (!reuse) {
            B200SA_CU(cudaStreamSynchronize(st));  // the speculative pass still reads text_ws
            B200SA_CU(cudaMemcpyAsync(e.text_ws.p, d_new, (size_t)n, cudaMemcpyDeviceToDevice, st));
        }
    } else {
        B200SA_TRY(e.copy_in(e.text_ws.p, text, (size_t)n, st));
    }
    i64 sentinel = e.sa_cache.sentinel;
    if (!reuse) {
        B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st, max_n));
        B200SA_CU(cudaMemcpyAsync(e.h_pinned + 26, e.rank.p, 4, cudaMemcpyDeviceToHost, st));  // rank[0]: the row of suffix 0
        B200SA_CU(cudaStreamSynchronize(st));
        sentinel = (i64)e.h_pinned[26];
    }
    if (want_bwt && !bwt_done) {
        // the gather runs while the suffix array travels to the host (pageable destination: the staging threads' streams;
        // pinned: the copy stream)
        B200SA_TRY(e.bwt_rows(e.text_ws.as<u8>(), (u32)n, e.sa_ws.as<i32>(), 0, (u32)n, e.bwt_ws.as<u8>(), st, /*defer_sync=*/true));
    }
    if (sa_out) B200SA_TRY(e.copy_out(sa_out, e.sa_ws.p, ((size_t)n + 1) * 4, st, /*independent=*/true));
    if (bwt_out) B200SA_TRY(e.copy_out(bwt_out, e.bwt_ws.p, (size_t)n, st));
    B200SA_CU(cudaStreamSynchronize(st));
    B200SA_CU(cudaStreamSynchronize(e.copy_stream));
    if (sentinel_out) *sentinel_out = sentinel;
    if (e.profiling) B200SA_TRY(e.collect_profile());
    e.sa_cache.valid = true;
    e.sa_cache.sharded = false;
    e.sa_cache.n = (u64)n;
    e.sa_cache.sentinel = sentinel;
    return 0;
}

int b200sa_suffix_array_bwt(b200sa_ctx* ctx, const uint8_t* text, int64_t n, int32_t* sa_out, uint8_t* bwt_out,
                            int32_t* sentinel_index_out)
{
    int64_t s64 = 0;
    B200SA_TRY(sa_bwt_host(ctx, text, n, sa_out, bwt_out, sentinel_index_out || bwt_out ? &s64 : nullptr, B200SA_MAX_N_INT32));
    if (sentinel_index_out) *sentinel_index_out = (int32_t)s64;
    return 0;
}

int b200sa_suffix_array_bwt_u32(b200sa_ctx* ctx, const uint8_t* text, int64_t n, uint32_t* sa_out, uint8_t* bwt_out,
                                int64_t* sentinel_index_out)
{
    int64_t s64 = 0;
    B200SA_TRY(sa_bwt_host(ctx, text, n, sa_out, bwt_out, sentinel_index_out || bwt_out ? &s64 : nullptr, B200SA_MAX_N_UINT32));
    if (sentinel_index_out) *sentinel_index_out = s64;
    return 0;
}

int b200sa_check_suffix_array(b200sa_ctx* ctx, const uint8_t* text, int64_t n, const int32_t* sa, int64_t* bad_rows_out)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > B200SA_MAX_N_INT32 || !sa || !bad_rows_out || (n > 0 && !text)) return b200sa::set_error(B200SA_EINVAL, "bad argument");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n + 64));
    B200SA_TRY(e.sa_ws.ensure(((size_t)n + 1) * 4));
    if (n) B200SA_TRY(e.copy_in(e.text_ws.p, text, (size_t)n, st));
    B200SA_TRY(e.copy_in(e.sa_ws.p, sa, ((size_t)n + 1) * 4, st));
    return e.check_sa_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), bad_rows_out, st);
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

static int unbwt_host(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int64_t sentinel_index, int64_t max_n)
{
    B200SA_NEED_CTX(ctx);
    Engine& e = ctx->eng;
    if (n < 0 || n > max_n) return b200sa::set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n, (long long)max_n);
    if (n == 0) return 0;
    if (!bwt_inout) return b200sa::set_error(B200SA_EINVAL, "null buffer");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.text_ws.ensure((size_t)n));
    B200SA_TRY(e.bwt_ws.ensure((size_t)n));
    B200SA_TRY(e.copy_in(e.bwt_ws.p, bwt_inout, (size_t)n, st));
    // an input that is not a BWT fails here with B200SA_EINVAL; the caller's buffer is then left as it was
    B200SA_TRY(e.unbwt_dev(e.bwt_ws.as<u8>(), n, sentinel_index, e.text_ws.as<u8>(), st, max_n));
    B200SA_TRY(e.copy_out(bwt_inout, e.text_ws.p, (size_t)n, st));
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

int b200sa_unbwt(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int32_t sentinel_index)
{
    return unbwt_host(ctx, bwt_inout, n, sentinel_index, B200SA_MAX_N_INT32);
}

int b200sa_unbwt_u32(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int64_t sentinel_index)
{
    return unbwt_host(ctx, bwt_inout, n, sentinel_index, B200SA_MAX_N_UINT32);
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
    B200SA_TRY(e.copy_in(e.text_ws.p, text, (size_t)n, st));
    if (sa) {
        // a suffix array from outside (another tool, a file): the phi scatter and the text probes index with its entries, so it
        // is validated first (O(n), the validator of b200sa_check_suffix_array)
        B200SA_TRY(e.copy_in(e.sa_ws.p, sa, ((size_t)n + 1) * 4, st));
        i64 bad = 0;
        B200SA_TRY(e.check_sa_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), &bad, st));
        if (bad != 0) return b200sa::set_error(B200SA_EINVAL, "sa is not the suffix array of text (%lld offending rows)", (long long)bad);
    } else {
        B200SA_TRY(e.suffix_array_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), st));
    }
    if (sa_out) B200SA_TRY(e.copy_out(sa_out, e.sa_ws.p, ((size_t)n + 1) * 4, st, /*independent=*/true));
    B200SA_TRY(e.lcp_dev(e.text_ws.as<u8>(), n, e.sa_ws.as<i32>(), e.keys[1].as<i32>(), st));
    B200SA_TRY(e.copy_out(lcp_out, e.keys[1].p, ((size_t)n + 1) * 4, st));
    B200SA_CU(cudaStreamSynchronize(st));
    B200SA_CU(cudaStreamSynchronize(e.copy_stream));
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
    if (total) B200SA_TRY(e.copy_in(e.text_ws.p, blocks, (size_t)total, st));
    u8* d_bwt = nullptr;
    i32* d_sa = nullptr;
    if (bwt_out) { B200SA_TRY(e.bwt_ws.ensure((size_t)total + 64)); d_bwt = e.bwt_ws.as<u8>(); }
    if (sa_out) { B200SA_TRY(e.batch_out.ensure(((size_t)total + (size_t)count) * 4 + 64)); d_sa = e.batch_out.as<i32>(); }
    // on failure the stream is drained before returning: no copy into the caller's buffers is still in flight
    int rc = e.batch_dev(e.text_ws.as<u8>(), offsets, count, d_bwt, d_sa, sentinel_index_out, st);
    if (rc == 0 && bwt_out && total) rc = e.copy_out(bwt_out, d_bwt, (size_t)total, st);
    if (rc == 0 && sa_out) rc = e.copy_out(sa_out, d_sa, ((size_t)total + (size_t)count) * 4, st);
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == 0) rc = b200sa::set_error(B200SA_ECUDA, "stream synchronisation failed");
    return rc;
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

// Inverse batch in two steps: decode into the context's text_ws (nothing of the caller's is written), then copy out.  The group
// form (c_abi_group.inl) runs the first step on every GPU before any GPU starts the second.
static int unbwt_batch_decode(b200sa_ctx* ctx, const uint8_t* blocks, const int64_t* offsets, int64_t count, const int32_t* sentinel_index,
                              bool synchronise)
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
    if (total > 0 && !blocks) return b200sa::set_error(B200SA_EINVAL, "null blocks");
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    B200SA_TRY(e.bwt_ws.ensure((size_t)total + 64));
    B200SA_TRY(e.text_ws.ensure((size_t)total + 64));
    int rc = total ? e.copy_in(e.bwt_ws.p, blocks, (size_t)total, st) : 0;
    if (rc == 0) rc = e.unbwt_batch_dev(e.bwt_ws.as<u8>(), offsets, count, sentinel_index, e.text_ws.as<u8>(), st);
    if ((rc != 0 || synchronise) && cudaStreamSynchronize(st) != cudaSuccess && rc == 0) rc = b200sa::set_error(B200SA_ECUDA, "stream synchronisation failed");
    return rc;
}

static int unbwt_batch_copy_out(b200sa_ctx* ctx, uint8_t* blocks_out, int64_t total)
{
    Engine& e = ctx->eng;
    B200SA_CU(cudaSetDevice(e.device));
    cudaStream_t st = e.own_stream;
    int rc = total > 0 ? e.copy_out(blocks_out, e.text_ws.p, (size_t)total, st) : 0;
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == 0) rc = b200sa::set_error(B200SA_ECUDA, "stream synchronisation failed");
    return rc;
}

int b200sa_unbwt_batch(b200sa_ctx* ctx, uint8_t* blocks_inout, const int64_t* offsets, int64_t count, const int32_t* sentinel_index)
{
    B200SA_TRY(unbwt_batch_decode(ctx, blocks_inout, offsets, count, sentinel_index, /*synchronise=*/false));
    return count > 0 ? unbwt_batch_copy_out(ctx, blocks_inout, offsets[count]) : 0;
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

// `depth` contexts on every listed device behind ONE queue: with several devices the stream of batches spreads over all
// GPUs and all PCIe links (batches are independent: nothing is exchanged)
int b200sa_pipeline_create_devices(b200sa_pipeline** out, const int* devices, int count, int depth)
{
    if (!out) return b200sa::set_error(B200SA_EINVAL, "null out pointer");
    *out = nullptr;
    if (!devices || count < 1 || count > b200sa::kMaxPeers) return b200sa::set_error(B200SA_EINVAL, "between 1 and %d devices", b200sa::kMaxPeers);
    if (depth < 1 || depth > 8) return b200sa::set_error(B200SA_EINVAL, "pipeline depth %d outside [1, 8]", depth);
    b200sa_pipeline* p = new (std::nothrow) b200sa_pipeline();
    if (!p) return b200sa::set_error(B200SA_ENOMEM, "out of host memory");
    for (int i = 0; i < depth; ++i)
        for (int d = 0; d < count; ++d) {
            b200sa_ctx* c = nullptr;
            const int rc = b200sa_create(&c, devices[d]);
            if (rc != 0) {
                for (auto* q : p->ctxs) b200sa_destroy(q);
                delete p;
                return rc;
            }
            p->ctxs.push_back(c);
        }
    for (size_t i = 0; i < p->ctxs.size(); ++i) p->workers.emplace_back([p, i] { p->run(i); });
    *out = p;
    return 0;
}

int b200sa_pipeline_create(b200sa_pipeline** out, int device, int depth)
{
    return b200sa_pipeline_create_devices(out, &device, 1, depth);
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
    // also wakes when somebody else collected the ticket meanwhile (a second waiter, or b200sa_pipeline_drain)
    p->cv_done.wait(lk, [&] { return p->done.count(ticket) != 0 || p->open_tickets.count(ticket) == 0; });
    auto it = p->done.find(ticket);
    if (it == p->done.end())
        return b200sa::set_error(B200SA_EINVAL, "ticket %lld was collected by another waiter or by b200sa_pipeline_drain", (long long)ticket);
    const int rc = it->second.first;
    if (rc) b200sa::set_error(rc, "%s", it->second.second.c_str());
    p->done.erase(it);
    p->open_tickets.erase(ticket);
    lk.unlock();
    p->cv_done.notify_all();
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
    lk.unlock();
    p->cv_done.notify_all();  // waiters of tickets collected here wake up and report EINVAL
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

// Partition (multi-split, not stable) of (key, value) pairs by (key >> shift) & 255 — the routing step of every all-to-all of the
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
    auto kp = b200sa::k_onesweep_pass<u32, true, false>;  // multi-split: order inside a bucket is irrelevant
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
    B200SA_NEED_CTX_KEEP(ctx);
    ctx->eng.profiling = enabled != 0;
    return 0;
}

int b200sa_profile_reset(b200sa_ctx* ctx)
{
    B200SA_NEED_CTX_KEEP(ctx);
    memset(&ctx->eng.prof, 0, sizeof(ctx->eng.prof));
    return 0;
}

int b200sa_profile_get(b200sa_ctx* ctx, b200sa_profile* out)
{
    B200SA_NEED_CTX_KEEP(ctx);
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
