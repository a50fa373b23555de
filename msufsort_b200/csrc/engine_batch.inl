// engine_batch.inl — Engine methods of the batched forward / inverse transforms (inside namespace b200sa).
// Included by b200sa.cu (one translation unit: the kernels are templates / static functions in the .cuh headers).

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
    B200SA_TRY(misc.ensure(8192));
    u32* d_bad = misc.as<u32>() + kUnbwtBadWord;
    B200SA_CU(cudaMemsetAsync(d_bad, 0, 4, st));
    prof.memsets++;
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
                  (const i32*)d_sent, count, nreg, D, N, W, (const u32*)ds[cur], (const u32*)nx[cur], (const u32*)idx[0].as<u32>(), (const u32*)ovf,
                  (const u8*)keys[0].as<u8>(), cap, d_out, d_bad);
    count_launch(B200SA_PH_UNBWT_WALK);
    B200SA_TRY(phase_end(st));
    prof.alg_bytes[B200SA_PH_UNBWT_WALK] += (u64)total * 9;
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaMemcpyAsync(h_pinned + 24, d_bad, 4, cudaMemcpyDeviceToHost, st));
    B200SA_CU(cudaStreamSynchronize(st));
    if (profiling) B200SA_TRY(collect_profile());
    if (h_pinned[24] != 0)
        return set_error(B200SA_EINVAL, "a block of the batch is not a Burrows-Wheeler transform (its LF mapping does not form one cycle through "
                                        "the sentinel row); the output buffer holds no valid text for that block");
    return 0;
}

