// engine.cuh — host-side orchestration of the suffix-array / BWT / inverse-BWT kernels.
//
// One Engine = one CUDA device + one stream + a reusable device workspace (the analogue of one
// maniscalco::msufsort object, msufsort.h:50-75, whose members hold one job's state).
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "sa_kernels.cuh"
#include "bwt_kernels.cuh"
#include "lcp_kernels.cuh"
#include "batch_kernels.cuh"
#include "../../include/b200sa.h"
#include "comm.cuh"

#include <string>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace b200sa {

int set_error(int code, const char* fmt, ...);

#define B200SA_CU(expr)                                                                                          \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            return b200sa::set_error(B200SA_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define B200SA_TRY(expr)           \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != 0) return rc__; \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T* as() const { return (T*)p; }
};

struct AlphabetPlan {
    u8 code[256];
    int sigma;     // distinct byte values
    int bits;      // bits per dense symbol
    int k;         // symbols packed into the initial key
    int len_bits;  // bits of the clamped-length field
    // mixed-radix packing (experimental, B200SA_PACK_RADIX=1): key = sum digit_j * radix^(k-1-j), digit = symbol + 1, 0 past
    // the end of the text — no length field, and no bits wasted on alphabets that are not a power of two (27 letters:
    // 13 symbols per key instead of 12).  radix == 0: the bit-packed layout above.
    u64 radix;
    u64 pow[66];   // radix^j
    int key_bits;  // significant bits of the symbol part of the key (both layouts)
};

// reserved_bits: key bits kept free above the symbols (the block number of a batched sort)
AlphabetPlan plan_alphabet(const u32* hist256, int reserved_bits = 0, int max_key_bits = 64, bool allow_radix = false);
static inline int bit_length_u64(u64 x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

struct Engine {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    int num_sms = kNumSMs;

    // workspace (sized for the largest n seen; see DESIGN.md "data layout in HBM")
    DevBuf keys[2], idx[2], slot[2], gid, gstart, glist, rank, sa_ws, sortmeta, agg_cnt, agg_max, misc, text_ws, bwt_ws, walk;
    DevBuf batch_text, batch_meta, batch_out;  // batched sort: expanded text, block tables, staging of packed results
    u32* h_pinned = nullptr;  // 512 words of pinned host memory for small read-backs (offsets: see the uses in b200sa.cu)

    // rank[] arrays up to this size are updated by direct scatter (they stay resident in the 126 MB
    // L2); larger ones by the bucketed update when a round has at least isa_min_updates tuples
    size_t isa_direct_bytes = (size_t)48 << 20;
    u32 isa_min_updates = 1u << 20;

    // rounds >= 1 sort every group where it lies when the average group has at most this many members
    // (0 disables the path); tiny groups go to one thread, medium ones to one CTA
    u32 groupsort_max_avg = 16;
    u32 groupsort_tiny = GS_TINY;
    u32 groupsort_medium = GS_MEDIUM;

    // forward BWT: the gather runs in passes over text windows of at most this many bytes, so that the gathered window stays in
    // L2 (B200SA_BWT_WINDOW_BYTES; 0 = always one pass; tests set a few KB to drive the windowed kernel at small n).  Measured
    // on B200 at 256 MiB (profiles/r02_bwt_window.txt): one pass 4.35 ms; windows of 96 / 64 / 48 / 32 MiB 3.22 / 3.23 / 3.75 /
    // 4.84 ms.  Texts that would need more than bwt_max_passes windows take the single pass.
    size_t bwt_window_bytes = (size_t)96 << 20;
    u32 bwt_max_passes = 4;

    // sharded radix rounds pull the peers' ISA shards in bulk when this GPU reads more than n / isa_pull_fraction ranks
    // (B200SA_ISA_PULL_FRACTION; 0 = always load remotely, 1000000000 = always pull: tests)
    u32 isa_pull_fraction = 12;

    // inverse BWT: bytes of decode window per walker = unbwt_cap_mult * D (D = mean segment length)
    u32 unbwt_cap_mult = 4;

    // width of the round-0 key (symbols + length field + block bits): fewer bits = fewer radix sweeps in round 0 but
    // more suffixes left for the doubling rounds (B200SA_MAX_KEY_BITS; measured in profiles/)
    int max_key_bits = 64;
    // measured on B200, 256 MiB Markov text (profiles/r02_knobs.txt): budgeted row-wise LCP comparison before the PLCP route
    // 5.7 ms against 17.6 ms; mixed-radix round-0 keys (13 instead of 12 symbols) 31.4 ms per SA+BWT step against 33.9 ms
    bool lcp_direct = true;   // B200SA_LCP_DIRECT=0: always take the PLCP route
    bool pack_radix = true;   // B200SA_PACK_RADIX=0: bit-packed round-0 keys (symbols << len_bits | clamped length) only

    // instrumentation
    bool profiling = false;
    b200sa_profile prof;
    uint64_t total_launches = 0;
    struct Span { cudaEvent_t a, b; int phase; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    int open_phase = -1;
    cudaEvent_t open_event = nullptr;

    int init(int dev);
    void shutdown();
    int release_workspace();

    // Host <-> device transfers of the host-buffer entry points.  Pinned (or registered) host memory: one asynchronous copy on
    // `st`.  Pageable memory (what a std::vector handed to the reference-shaped facade is): `copy_threads` host threads move
    // chunks through their own pinned staging buffers and streams, so the PCIe copy of one chunk overlaps the host memcpy of
    // the next (the driver's own pageable path is a single-threaded staging loop).  copy_in: work enqueued on `st` afterwards
    // sees the data.  copy_out: ordered after the work already enqueued on `st`; complete on return for pageable memory.
    struct HostStage {
        static const int kMaxThreads = 8;
        void* buf[kMaxThreads][2] = {};
        cudaStream_t stream[kMaxThreads] = {};
        cudaEvent_t done[kMaxThreads][2] = {};
        cudaEvent_t fence = nullptr;
        size_t chunk = (size_t)8 << 20;
        int threads = 0;  // > 0 once allocated
    } stage;
    int copy_threads = 4;                       // B200SA_COPY_THREADS (0 = always the driver's own path)
    size_t copy_staged_min = (size_t)4 << 20;   // smaller pageable transfers are left to the driver
    // The host entry points keep the last text and its suffix array resident (text_ws, sa_ws).  A following call that is
    // handed the same bytes (compared on the device after the upload) reuses the sort: make_suffix_array followed by
    // forward_burrows_wheeler_transform on one msufsort object costs one sort, not two.  Every other entry point drops it.
    // sharded: the resident result is this context's part of a group's sharded sort (b200sa_group_*): the whole text, but only
    // this rank's rows of the suffix array — only the group may reuse it.
    struct SaCache { bool valid = false; bool sharded = false; u64 n = 0; i64 sentinel = 0; } sa_cache;
    int stage_ensure();
    int copy_in(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
    // independent: the source is complete already (the stream was synchronised after its producer) — the copy need not
    // queue behind later work on `st`; pinned destinations then use copy_stream, which the caller synchronises
    int copy_out(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st, bool independent = false);
    cudaStream_t copy_stream = nullptr;

    cudaStream_t pick(void* s) const { return s ? (cudaStream_t)s : own_stream; }

    // phases / profiling
    int phase_begin(int phase, cudaStream_t st);
    int phase_end(cudaStream_t st);
    int collect_profile();
    void count_launch(int phase) { prof.launches[phase]++; total_launches++; }

    // building blocks
    int radix_sort_pairs(u64* keys2[2], u32* vals2[2], bool gen_vals, u32 m, int begin_bit, int end_bit,
                         int* result_side, cudaStream_t st);
    // target == nullptr: the engine's rank[] (ISA); the LCP path scatters phi[] with the same machinery
    int isa_update(const u32* d_idx, const u32* d_val, u32 count, u32 n, u32* bk_key, u32* bk_val, bool all_suffixes, cudaStream_t st,
                   u32* target = nullptr);
    int rerank(const u64* keys_sorted, const u32* idx_sorted, const u32* slot_in, u32 slot_base, u32 m, u32 n, i32* d_sa,
               u32* idx_out, u32* slot_out, u64* free_keys, int mode, u32* next_m, u32* next_groups, cudaStream_t st);

    // suffix sort as resumable steps (shared by the single-GPU entry point and the sharded driver)
    struct SortState {
        int stage = 0;  // 0 idle, 1 first sort done, 2 ranking rounds, 3 finished
        const u8* d_text = nullptr;
        u32 n = 0;
        i32* d_sa = nullptr;
        int part = 0, nparts = 1;
        AlphabetPlan plan;
        u32 n_local = 0, m = 0, groups = 0;
        int sorted_side = 0, act = 0, cur_slot = 0, rank_bits = 0, guard = 0;
        u64 h = 0;
        const u32* upd_idx = nullptr;   // ISA updates produced by the last step (sharded runs)
        const u32* upd_rank = nullptr;
        u32 upd_count = 0;
        // batched sort (batch_kernels.cuh): separator positions of the blocks in expanded coordinates
        const u32* batch_ends = nullptr;
        u32 batch_count = 0;
        int batch_bits = 0;
    } ss;
    // ISA sharded over the GPUs of one box and read / written through peer memory (isa="peer")
    struct PeerState {
        bool active = false;
        int part = 0, nparts = 1;
        RankView view{};
        bool has_isa = false;               // view.base[] is mapped (suffix sort); otherwise out[] is (inverse BWT)
        u8* inbox[kMaxPeers] = {};          // every GPU's inbox (own one included)
        u8* out[kMaxPeers] = {};            // every GPU's inverse-BWT output buffer
        u64 region_off[kMaxPeers] = {};     // region of source s inside any inbox
        u32 region_cap[kMaxPeers] = {};
        bool laid_out = false;
        std::vector<std::pair<std::string, void*>> opened;  // IPC handle bytes -> mapped pointer
    } peer;
    DevBuf peer_inbox, peer_out;
    // how one peer's ISA array and inbox are reached: a CUDA IPC handle pair (another process) or plain pointers (a context
    // of this process, possibly on another device: peer access is enabled on attach)
    struct PeerDesc {
        u64 pid;
        int device;
        int pad;
        void* rank_ptr;
        void* inbox_ptr;
        void* out_ptr;
        unsigned char ipc[192];  // handles of: ISA array, inbox, inverse-BWT output buffer
    };
    // isa = true: the ISA array and the inbox are shared (suffix sort); false: the inbox and the output buffer (inverse BWT)
    int peer_describe(u64 n, bool with_ipc, bool isa, PeerDesc* out);
    int peer_attach_desc(int part, int nparts, int shift, u64 n, bool isa, const PeerDesc* descs);
    int peer_export(u64 n, unsigned char* handles_out /*128*/);
    int peer_layout(const i64* counts, int nparts);
    int peer_apply(cudaStream_t st);
    int peer_attach(int part, int nparts, int shift, u64 n, const unsigned char* handles);
    int peer_scatter(cudaStream_t st);
    int peer_detach();
    struct BatchDesc { const u32* d_ends = nullptr; u32 count = 0; } next_batch;  // consumed by the next sort_begin
    int sort_begin(const u8* d_text, u32 n, i32* d_sa, int part, int nparts, u32* n_local, cudaStream_t st);
    // the whole sharded sort of one rank, control plane through `cm` (comm.cuh): attach the peers' memory, begin, round 0,
    // doubling rounds with their two barriers each, this rank's rows of the BWT (d_bwt may be null)
    struct ShardInfo { i64 row_begin, row_end, out_begin, out_end, sentinel, rounds, sent_bytes, n_local; };
    Comm* shard_comm = nullptr;  // set while sharded_sort runs: sort_begin then histograms only this rank's slice of the text
    int sharded_sort(Comm& cm, const u8* d_text, i64 n, i32* d_sa, u8* d_bwt, ShardInfo* info, cudaStream_t st);
    // inverse BWT with the walkers split over the ranks; every rank leaves its slice of the text in d_out and, with
    // gather_all, copies the other slices from the peers' buffers (peer memory) so that d_out holds the whole text
    int sharded_unbwt(Comm& cm, const u8* d_bwt, i64 n, i64 sentinel, u8* d_out, bool gather_all, i64* slice_begin, i64* slice_end,
                      cudaStream_t st);
    int sort_round0(u32 slot_base, u32* m_local, cudaStream_t st);
    int sort_round(u32* m_local, cudaStream_t st);

    // entry points
    int ensure_sa_workspace(u64 n);
    // max_n: B200SA_MAX_N_INT32 for the reference-shaped int32 entry points, B200SA_MAX_N_UINT32 for the wide ones
    int suffix_array_dev(const u8* d_text, i64 n, i32* d_sa, cudaStream_t st, i64 max_n = B200SA_MAX_N_INT32);
    // defer_sync: only enqueue the kernel (the caller knows the sentinel row already and synchronises later)
    int bwt_rows(const u8* d_text, u32 n, const i32* d_sa, u32 o_begin, u32 o_end, u8* d_bwt, cudaStream_t st, bool defer_sync = false);
    int bwt_dev(const u8* d_text, i64 n, u8* d_bwt, i32* d_sa_or_null, i64* sentinel_host, cudaStream_t st, i64 max_n = B200SA_MAX_N_INT32);
    struct UnbwtState { int stage = 0, dshift = 6; u32 n = 0, s = 0, D = 0, nreg = 0, nwalkers = 0, cap = 0; } us;
    int unbwt_build(const u8* d_bwt, u32 n, u32 s, u32* nwalkers_out, cudaStream_t st);
    int unbwt_measure(u32 w_begin, u32 w_end, cudaStream_t st);
    // so != nullptr: the bytes go to the owners of their text positions through peer memory instead of d_out
    int unbwt_finish(u32 w_begin, u32 w_end, u8* d_out, cudaStream_t st, const ShardedOut* so = nullptr);
    int unbwt_dev(const u8* d_bwt, i64 n, i64 sentinel, u8* d_out, cudaStream_t st, i64 max_n = B200SA_MAX_N_INT32);
    int check_sa_dev(const u8* d_text, i64 n, const i32* d_sa, i64* bad_rows, cudaStream_t st, i64 max_n = B200SA_MAX_N_INT32);
    int lcp_dev(const u8* d_text, i64 n, const i32* d_sa, i32* d_lcp, cudaStream_t st, i64 max_n = B200SA_MAX_N_INT32);
    // batch of independent blocks, packed back to back at offsets[0..count]; any of the outputs may be null
    int batch_dev(const u8* d_packed, const i64* offsets, i64 count, u8* d_bwt_out, i32* d_sa_out, i32* sentinels_host, cudaStream_t st);
    int batch_tables(const i64* offsets, u32 count, const i32* sentinels_or_null, u32** d_ends, u32** d_offs, i32** d_sent, cudaStream_t st);
    int unbwt_batch_dev(const u8* d_bwt_packed, const i64* offsets, i64 count, const i32* sentinels_host, u8* d_out_packed, cudaStream_t st);
};

}  // namespace b200sa
