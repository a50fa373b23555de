// engine_shard.inl — one text sharded over the GPUs of one box, driven from C++ (inside namespace b200sa).
// Included by b200sa.cu (one translation unit).  Control plane: comm.cuh.  Data plane: engine_peer.inl (peer memory).
//
// Replaces, in the reference, the static split of the text over its worker threads (msufsort.cpp:1576-1586, 1635-1643):
// there every thread takes n/T positions of the first-stage count and the B* buckets are sorted in parallel; here every
// GPU takes one key range of the suffixes (whole groups, so all later sorts are local), and the only thing that travels
// between GPUs is the inverse suffix array.

static int owner_shift(u64 n, int G)
{
    const u64 per = (n + (u64)G - 1) / (u64)G;
    return per <= 1 ? 0 : bit_length_u64(per - 1);  // B = 2^shift >= ceil(n/G): owner(p) = p >> shift < G for p < n
}

static int peer_attach_comm(Engine& e, Comm& cm, u64 n, int shift, bool isa)
{
    Engine::PeerDesc mine, all[kMaxPeers];
    static_assert(sizeof(Engine::PeerDesc) <= (size_t)kCommSlotBytes, "peer descriptor fits a comm slot");
    B200SA_TRY(e.peer_describe(n, cm.shm, isa, &mine));
    B200SA_TRY(cm.allgather(&mine, sizeof(mine), all));
    return e.peer_attach_desc(cm.rank, cm.nranks, shift, n, isa, all);
}

int Engine::sharded_sort(Comm& cm, const u8* d_text, i64 n64, i32* d_sa, u8* d_bwt, ShardInfo* info, cudaStream_t st)
{
    if (n64 <= 0 || n64 > B200SA_MAX_N_INT32 || !d_text || !d_sa || !info) return set_error(B200SA_EINVAL, "bad argument");
    const int G = cm.nranks, me = cm.rank;
    memset(info, 0, sizeof(*info));
    B200SA_CU(cudaSetDevice(device));
    const u32 n = (u32)n64;
    if (G == 1) {
        i64 s64 = 0;
        if (d_bwt) B200SA_TRY(bwt_dev(d_text, n64, d_bwt, d_sa, &s64, st));
        else {
            B200SA_TRY(suffix_array_dev(d_text, n64, d_sa, st));
            B200SA_CU(cudaMemcpyAsync(h_pinned + 26, rank.p, 4, cudaMemcpyDeviceToHost, st));
            B200SA_CU(cudaStreamSynchronize(st));
            s64 = h_pinned[26];
        }
        info->row_begin = 0; info->row_end = n64 + 1; info->out_begin = 0; info->out_end = n64; info->sentinel = s64;
        info->rounds = (i64)prof.rounds; info->n_local = n64;
        return 0;
    }
    struct Guard {  // a rank that fails must not leave its peers waiting
        Comm& cm; Engine& e; bool ok = false;
        ~Guard() { e.shard_comm = nullptr; if (!ok) cm.raise_error(); }
    } guard{cm, *this};
    B200SA_TRY(peer_attach_comm(*this, cm, n, owner_shift(n, G), true));
    shard_comm = &cm;
    u32 n_local = 0, m = 0;
    B200SA_TRY(sort_begin(d_text, n, d_sa, me, G, &n_local, st));
    i64 counts[kMaxPeers], mine = n_local, total = 0;
    B200SA_TRY(cm.allgather(&mine, sizeof(mine), counts));
    u64 slot_base = 0;
    for (int g = 0; g < G; ++g) { total += counts[g]; if (g < me) slot_base += (u64)counts[g]; }
    if (total != n64) return set_error(B200SA_EINTERNAL, "key-range parts cover %lld of %lld suffixes", (long long)total, (long long)n64);
    B200SA_TRY(peer_layout(counts, G));
    B200SA_TRY(sort_round0((u32)slot_base, &m, st));
    i64 rounds = 1, sent = 0;
    for (;;) {
        // a rank gets here when its own round is complete.  New ranks go to the owners' INBOXES (nobody reads those during a
        // round), so no barrier is needed before the sends ...
        sent += (i64)8 * ss.upd_count * (G - 1) / G;
        B200SA_TRY(peer_scatter(st));                    // returns when this rank's stores have landed
        i64 active = 0;
        B200SA_TRY(cm.allreduce_sum((i64)m, &active));   // ... all sends have landed (and the termination test) ...
        B200SA_TRY(peer_apply(st));                      // ... every owner updates its ISA shard ...
        B200SA_TRY(cm.barrier());                        // ... and all shards are current before anyone reads again
        if (active == 0) break;
        B200SA_TRY(sort_round(&m, st));
        ++rounds;
    }
    info->row_begin = me == 0 ? 0 : (i64)slot_base + 1;  // row 0 (the empty suffix) belongs to rank 0
    info->row_end = (i64)slot_base + n_local + 1;
    info->rounds = rounds;
    info->sent_bytes = sent;
    info->n_local = n_local;
    // sentinel row s = rank[0], held by the owner of position 0
    B200SA_CU(cudaMemcpyAsync(h_pinned + 9, peer.view.base[0], 4, cudaMemcpyDefault, st));
    B200SA_CU(cudaStreamSynchronize(st));
    const i64 s = h_pinned[9];
    info->sentinel = s;
    info->out_begin = info->row_begin - (info->row_begin > s ? 1 : 0);
    info->out_end = info->row_end - (info->row_end > s ? 1 : 0);
    if (d_bwt && info->out_end > info->out_begin) B200SA_TRY(bwt_rows(d_text, n, d_sa, (u32)info->out_begin, (u32)info->out_end, d_bwt, st));
    if (profiling) B200SA_TRY(collect_profile());
    guard.ok = true;
    return 0;
}

// Inverse BWT, walkers split over the ranks.  The psi table is built on every GPU (one histogram + one sweep over the n
// bytes: replicated, it is the cheap half); every rank walks the segments of its slice of the walkers, the (length,
// successor) entries are broadcast into the peers' inboxes (a few MB), every rank ranks the whole segment list and then
// stores the bytes of ITS segments straight into the owner of their text position: text byte p lives on GPU p >> shift, in
// that GPU's peer-mapped output buffer, written over NVLink in runs of a whole segment.  Round 1 summed G zero-padded
// n-byte buffers with an NCCL all-reduce instead.
int Engine::sharded_unbwt(Comm& cm, const u8* d_bwt, i64 n64, i64 sentinel, u8* d_out, bool gather_all, i64* slice_begin, i64* slice_end,
                          cudaStream_t st)
{
    if (n64 <= 0 || n64 > B200SA_MAX_N_INT32 || !d_bwt || !slice_begin || !slice_end) return set_error(B200SA_EINVAL, "bad argument");
    if (sentinel < 1 || sentinel > n64) return set_error(B200SA_EINVAL, "sentinel index %lld outside [1, n]", (long long)sentinel);
    const int G = cm.nranks, me = cm.rank;
    B200SA_CU(cudaSetDevice(device));
    if (G == 1) {
        if (!d_out) return set_error(B200SA_EINVAL, "null output");
        *slice_begin = 0; *slice_end = n64;
        return unbwt_dev(d_bwt, n64, sentinel, d_out, st);
    }
    struct Guard {
        Comm& cm; bool ok = false;
        ~Guard() { if (!ok) cm.raise_error(); }
    } guard{cm};
    const u32 n = (u32)n64;
    const int shift = owner_shift(n, G);
    B200SA_TRY(peer_out.ensure((size_t)n + 64));
    B200SA_TRY(peer_attach_comm(*this, cm, n, shift, false));
    u32 W = 0;
    B200SA_TRY(unbwt_build(d_bwt, n, (u32)sentinel, &W, st));
    const u32 per = (W + (u32)G - 1) / (u32)G;
    const u32 wb = (u32)me * per < W ? (u32)me * per : W, we = (u32)(me + 1) * per < W ? (u32)(me + 1) * per : W;
    B200SA_TRY(unbwt_measure(wb, we, st));
    // ---- my (length, successor) entries into everybody's inbox: lengths at word 64 + w, successors at word 64 + W + w
    {
        PeerBcast pb;
        for (int g = 0; g < kMaxPeers; ++g) pb.dst[g] = g < G && g != me ? (u32*)(peer.inbox[g] + kInboxHeader) : nullptr;
        pb.nparts = G;
        const u32 cnt = we - wb;
        if (cnt) {
            const u32 want = (u32)div_up_u64(cnt, 256);
            B200SA_LAUNCH(k_peer_bcast_segments, want < (u32)(num_sms * 4) ? want : (u32)(num_sms * 4), 256, 0, st,
                          (const u32*)(walk.as<u32>() + 2 * (size_t)W), (const u32*)walk.as<u32>(), wb, we, W, pb);
            count_launch(B200SA_PH_UNBWT_WALK);
        }
        B200SA_CU(cudaGetLastError());
        B200SA_CU(cudaStreamSynchronize(st));
    }
    B200SA_TRY(cm.barrier());
    for (int g = 0; g < G; ++g) {
        if (g == me) continue;
        const u32 b = (u32)g * per < W ? (u32)g * per : W, e = (u32)(g + 1) * per < W ? (u32)(g + 1) * per : W;
        if (e <= b) continue;
        const u32* in = (const u32*)(peer_inbox.as<u8>() + kInboxHeader);
        B200SA_CU(cudaMemcpyAsync(walk.as<u32>() + 2 * (size_t)W + b, in + b, (size_t)(e - b) * 4, cudaMemcpyDeviceToDevice, st));
        B200SA_CU(cudaMemcpyAsync(walk.as<u32>() + b, in + W + b, (size_t)(e - b) * 4, cudaMemcpyDeviceToDevice, st));
    }
    ShardedOut so;
    for (int g = 0; g < kMaxPeers; ++g) so.base[g] = g < G ? peer.out[g] : nullptr;
    so.shift = shift;
    const int rc_finish = unbwt_finish(wb, we, nullptr, st, &so);
    i64 bad = rc_finish != 0 ? 1 : 0, any_bad = 0;
    B200SA_TRY(cm.allreduce_sum(bad, &any_bad));  // also: every rank's stores have landed
    if (any_bad) {
        guard.ok = true;  // an orderly failure: all ranks return together
        return rc_finish ? rc_finish
                         : set_error(B200SA_EINVAL, "the input is not a Burrows-Wheeler transform (detected by a peer rank); the output holds no valid text");
    }
    const u64 lo = (u64)me << shift, hi = ((u64)(me + 1) << shift) < n ? ((u64)(me + 1) << shift) : (u64)n;
    *slice_begin = lo < n ? (i64)lo : n64;
    *slice_end = lo < n ? (i64)hi : n64;
    if (gather_all) {
        for (int g = 0; g < G; ++g) {
            if (g == me) continue;
            const u64 glo = (u64)g << shift, ghi = ((u64)(g + 1) << shift) < n ? ((u64)(g + 1) << shift) : (u64)n;
            if (glo < ghi) B200SA_CU(cudaMemcpyAsync(peer_out.as<u8>() + glo, peer.out[g] + glo, (size_t)(ghi - glo), cudaMemcpyDefault, st));
        }
        if (d_out) B200SA_CU(cudaMemcpyAsync(d_out, peer_out.p, (size_t)n, cudaMemcpyDeviceToDevice, st));
    } else if (d_out && *slice_end > *slice_begin) {
        B200SA_CU(cudaMemcpyAsync(d_out + *slice_begin, peer_out.as<u8>() + *slice_begin, (size_t)(*slice_end - *slice_begin), cudaMemcpyDeviceToDevice, st));
    }
    B200SA_CU(cudaStreamSynchronize(st));
    B200SA_TRY(cm.barrier());  // nobody overwrites a peer's output buffer (next call) while it is still being read
    if (profiling) B200SA_TRY(collect_profile());
    guard.ok = true;
    return 0;
}
