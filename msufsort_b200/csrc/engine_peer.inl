// engine_peer.inl — Engine methods of the sharded ISA in NVLink peer memory (inside namespace b200sa).
// Included by b200sa.cu (one translation unit: the kernels are templates / static functions in the .cuh headers).

// ---------------------------------------------------------------------------------------------
// ISA sharded over the GPUs of one box, accessed through peer memory (CUDA IPC; msufsort_b200/sharded.py isa="peer").
// Two allocations per GPU are mapped by all peers: the ISA array (peers LOAD rank[suffix + h] from it) and an inbox
// (peers STORE the new ranks of the suffixes this GPU owns into it, in bulk; the owner then applies them locally).

static const size_t kInboxHeader = 32768;  // pair counts and bucket offsets written by the sources (sa_kernels.cuh: kInboxCountWords, kInboxOffsStride), padded

int Engine::peer_describe(u64 n, bool with_ipc, bool isa, PeerDesc* out)
{
    if (n == 0 || n > (u64)B200SA_MAX_N_INT32 || !out) return set_error(B200SA_EINVAL, "bad argument");
    B200SA_CU(cudaSetDevice(device));
    if (isa) {
        B200SA_TRY(ensure_sa_workspace(n));  // the ISA array keeps its address as long as n does not grow
        B200SA_TRY(peer_inbox.ensure(kInboxHeader + (size_t)n * 8 + 64));
    } else {
        B200SA_TRY(peer_inbox.ensure(kInboxHeader + (size_t)8 * (((size_t)1 << 21) + 64)));  // (length, successor) of <= 2^21 + 2 walkers
        B200SA_TRY(peer_out.ensure((size_t)n + 64));
    }
    memset(out, 0, sizeof(*out));
    out->pid = (u64)getpid();
    out->device = device;
    out->rank_ptr = isa ? rank.p : nullptr;
    out->inbox_ptr = peer_inbox.p;
    out->out_ptr = isa ? nullptr : peer_out.p;
    if (with_ipc) {
        cudaIpcMemHandle_t h[3];
        static_assert(sizeof(h[0]) == 64, "IPC handle size");
        memset(h, 0, sizeof(h));
        if (isa) B200SA_CU(cudaIpcGetMemHandle(&h[0], rank.p));
        B200SA_CU(cudaIpcGetMemHandle(&h[1], peer_inbox.p));
        if (!isa) B200SA_CU(cudaIpcGetMemHandle(&h[2], peer_out.p));
        memcpy(out->ipc, h, 192);
    }
    return 0;
}

int Engine::peer_export(u64 n, unsigned char* handles_out)
{
    if (!handles_out) return set_error(B200SA_EINVAL, "bad argument");
    PeerDesc d;
    B200SA_TRY(peer_describe(n, true, true, &d));
    memcpy(handles_out, d.ipc, 128);
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
    if (nparts < 2 || nparts > kMaxPeers || !handles) return set_error(B200SA_EINVAL, "bad argument (at most %d GPUs)", kMaxPeers);
    PeerDesc descs[kMaxPeers];
    memset(descs, 0, sizeof(descs));
    for (int g = 0; g < nparts; ++g) {
        descs[g].pid = ~(u64)0;  // handles only: every peer is treated as another process
        memcpy(descs[g].ipc, handles + (size_t)g * 128, 128);
    }
    return peer_attach_desc(part, nparts, shift, n, true, descs);
}

int Engine::peer_attach_desc(int part, int nparts, int shift, u64 n, bool isa, const PeerDesc* descs)
{
    if (nparts < 2 || nparts > kMaxPeers || part < 0 || part >= nparts || shift < 0 || shift > 31 || !descs || n == 0 ||
        n > (u64)B200SA_MAX_N_INT32)
        return set_error(B200SA_EINVAL, "bad argument (at most %d GPUs)", kMaxPeers);
    if (((n - 1) >> shift) >= (u64)nparts) return set_error(B200SA_EINVAL, "shift %d does not spread %llu positions over %d GPUs", shift, (unsigned long long)n, nparts);
    B200SA_CU(cudaSetDevice(device));
    {
        PeerDesc self;  // sizes the shared buffers exactly as peer_describe did on the peers
        B200SA_TRY(peer_describe(n, false, isa, &self));
    }
    const u64 mypid = (u64)getpid();
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
        if (g == part) { next.view.base[g] = rank.as<u32>(); next.inbox[g] = peer_inbox.as<u8>(); next.out[g] = peer_out.as<u8>(); continue; }
        void *pr = nullptr, *pi = nullptr, *po = nullptr;
        if (descs[g].pid == mypid) {
            // a context of this process: its pointers are valid here; across devices the hardware path is the same NVLink
            // peer mapping, enabled once per device pair
            pr = descs[g].rank_ptr;
            pi = descs[g].inbox_ptr;
            po = descs[g].out_ptr;
#ifndef B200SA_EMU
            if (descs[g].device != device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, device, descs[g].device);
                if (!can) rc = set_error(B200SA_ECOMM, "GPU %d cannot access the memory of GPU %d", device, descs[g].device);
                else {
                    cudaError_t e = cudaDeviceEnablePeerAccess(descs[g].device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = set_error(B200SA_ECOMM, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                    cudaGetLastError();
                }
            }
#endif
        } else {
            if (isa) rc = open_one(descs[g].ipc, &pr);
            if (rc == 0) rc = open_one(descs[g].ipc + 64, &pi);
            if (rc == 0 && !isa) rc = open_one(descs[g].ipc + 128, &po);
        }
        next.view.base[g] = (u32*)pr;
        next.inbox[g] = (u8*)pi;
        next.out[g] = (u8*)po;
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
    next.has_isa = isa;
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
        ps.offs_slot[g] = g < G ? (u32*)peer.inbox[g] + kInboxCountWords + me * kInboxOffsStride : nullptr;
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
        auto kp = k_onesweep_pass<u32, true, false>;  // multi-split: order inside a bucket is irrelevant
        B200SA_LAUNCH(kp, tiles, RS_THREADS, rs_pass_smem_bytes<u32>(), st, ss.upd_idx, bk_key, ss.upd_rank, bk_val, count, bshift,
                      0xffffffffu, (const u32*)ghist, status, counters);
        count_launch(B200SA_PH_ISA);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_ISA] += (u64)count * (4 + 16);
        B200SA_TRY(phase_begin(B200SA_PH_PEER_SEND, st));
        const u32 want = (u32)div_up_u64(count, 256 * 4);
        const u32 grid = want < (u32)(num_sms * 16) ? want : (u32)(num_sms * 16);
        B200SA_LAUNCH(k_peer_send, grid, 256, 0, st, (const u32*)bk_key, (const u32*)bk_val, count, (const u32*)ghist, ps);
        count_launch(B200SA_PH_PEER_SEND);
        prof.alg_bytes[B200SA_PH_PEER_SEND] += (u64)count * 16;
    } else {
        // nothing to send this round: the counts the owners read must still be reset (ghist is all zero)
        B200SA_LAUNCH(k_peer_send, 1, 256, 0, st, (const u32*)nullptr, (const u32*)nullptr, 0u, (const u32*)ghist, ps);
        count_launch(B200SA_PH_PEER_SEND);
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
    InboxArgs ia;
    memset(&ia, 0, sizeof(ia));
    u64 pairs = 0;
    for (int s = 0; s < G; ++s) {
        const u32 cnt = h_pinned[400 + s];
        if (cnt > peer.region_cap[s]) return set_error(B200SA_EINTERNAL, "GPU %d announced %u pairs for a region of %u", s, cnt, peer.region_cap[s]);
        ia.keys[s] = (const u32*)(peer_inbox.as<u8>() + peer.region_off[s]);
        ia.vals[s] = ia.keys[s] + peer.region_cap[s];
        pairs += cnt;
    }
    if (pairs) {
        // the same bucket geometry the senders used (peer_scatter)
        const int nbits = bit_length_u64((u64)ss.n - 1);
        int bshift = nbits > RS_RADIX_BITS ? nbits - RS_RADIX_BITS : 0;
        if (bshift > peer.view.shift) bshift = peer.view.shift;
        ia.header = (const u32*)peer_inbox.p;
        ia.nparts = G;
        ia.buckets = 1 << (peer.view.shift - bshift);
        if (ia.buckets * G > kInboxMaxEntries) return set_error(B200SA_EINTERNAL, "%d buckets x %d GPUs exceed the inbox plan", ia.buckets, G);
        B200SA_TRY(misc.ensure(16384));
        u32* plan = misc.as<u32>() + 1100;  // 3 * kInboxMaxEntries + 2 words
        const u32 tiles_bound = (u32)div_up_u64(pairs, SP_THREADS * SP_IPT) + (u32)(ia.buckets * G);
        B200SA_TRY(phase_begin(B200SA_PH_PEER_APPLY, st));
        B200SA_LAUNCH(k_inbox_plan, 1, kInboxMaxEntries, 0, st, ia, plan);
        count_launch(B200SA_PH_PEER_APPLY);
        B200SA_LAUNCH(k_scatter_plan, tiles_bound, SP_THREADS, 0, st, ia, (const u32*)plan, rank.as<u32>());
        count_launch(B200SA_PH_PEER_APPLY);
        B200SA_TRY(phase_end(st));
        prof.alg_bytes[B200SA_PH_PEER_APPLY] += pairs * 12;
    }
    B200SA_CU(cudaGetLastError());
    B200SA_CU(cudaStreamSynchronize(st));
    return 0;
}

