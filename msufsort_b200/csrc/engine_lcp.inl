// engine_lcp.inl — Engine::lcp_dev, the LCP array driver (inside namespace b200sa).
// Included by b200sa.cu (one translation unit: the kernels are templates / static functions in the .cuh headers).

// ---------------------------------------------------------------------------------------------
// LCP array (lcp_kernels.cuh): phi scatter, hierarchical PLCP levels, gather through the SA

int Engine::lcp_dev(const u8* d_text, i64 n64, const i32* d_sa, i32* d_lcp, cudaStream_t st, i64 max_n)
{
    if (n64 < 0 || n64 > max_n) return set_error(B200SA_EINVAL, "n = %lld outside [0, %lld]", (long long)n64, (long long)max_n);
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

