/*
 * b200sa.h — C ABI of the B200-native suffix-array / Burrows-Wheeler engine.
 *
 * This is the drop-in boundary for the hot path of michaelmaniscalco/msufsort.
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference tree, src/library/msufsort/...).  The C++ facade in
 * src/library/msufsort/msufsort.h (same public signatures as the reference's
 * header) is a thin marshalling layer over these functions; so is the ctypes
 * mirror in msufsort_b200/api.py.
 *
 * Conventions (identical to the reference, see msufsort.cpp:1720,1755-1766,
 * 1811-1816,1891):
 *   - the suffix array of an n-byte text has n+1 int32 entries, SA[0] = n (the
 *     empty suffix / virtual sentinel, smaller than every byte including 0x00),
 *     SA[1..n] = suffixes in lexicographic order, a proper prefix sorts first;
 *   - BWT[i] = T[SA[i]-1] over the n+1 rows with the single row whose SA[i]==0
 *     removed; that row number is the returned sentinel index (1..n); n bytes;
 *   - the inverse transform takes (n bytes, sentinel index) and returns T.
 *
 * All functions return 0 on success and a non-zero B200SA_E* code on failure;
 * b200sa_last_error() gives the message for the calling thread.  There is no
 * CPU fallback: without a CUDA device every compute entry point fails with
 * B200SA_ENODEVICE.
 *
 * Pointers named d_* are device pointers.  `stream` is a cudaStream_t passed
 * as void* (NULL = the context's own stream).  Plain pointers and sizes only:
 * no C++/torch types cross this boundary.
 */
#ifndef B200SA_H
#define B200SA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200SA_API
#else
#define B200SA_API __attribute__((visibility("default")))
#endif

enum {
    B200SA_OK = 0,
    B200SA_EINVAL = 1,     /* bad argument (null pointer, n < 0, n too large for the index type, bad sentinel) */
    B200SA_ENODEVICE = 2,  /* no usable CUDA device / driver */
    B200SA_ECUDA = 3,      /* a CUDA runtime call or kernel failed */
    B200SA_ENOMEM = 4,     /* device or pinned-host allocation failed */
    B200SA_EINTERNAL = 5,  /* invariant violated (bug) */
    B200SA_ECOMM = 6       /* NCCL / multi-GPU communication failure */
};

/* Largest n the int32 entry points accept (SA values must fit int32: n <= 2^31-2).
 * The reference itself is only correct up to 2^30-2 (flag bits, msufsort.h:84-93). */
#define B200SA_MAX_N_INT32 ((int64_t)2147483646)

/* Largest n of the wide (uint32) entry points: tile arithmetic keeps 8 KiB of headroom below 2^32.  One GPU holds
 * about 53 n bytes of workspace, i.e. n up to ~3.3e9 on a 180 GB B200. */
#define B200SA_MAX_N_UINT32 ((int64_t)4294959102)

typedef struct b200sa_ctx b200sa_ctx;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces: msufsort::msufsort(int32 numThreads) (msufsort.cpp:39-60) — the reference spawns
 * a host worker pool; here a context binds one CUDA device, owns one stream, the device
 * workspace (grown on demand, reused across calls) and pinned staging buffers.  One context
 * = one job at a time (same as one msufsort object, msufsort.h:281-309). */
B200SA_API int b200sa_create(b200sa_ctx** out, int device);
/* Replaces: msufsort::~msufsort() (msufsort.cpp:64-68). */
B200SA_API void b200sa_destroy(b200sa_ctx* ctx);
/* Frees the device workspace but keeps the context usable. */
B200SA_API int b200sa_release_workspace(b200sa_ctx* ctx);

B200SA_API const char* b200sa_last_error(void);
B200SA_API int b200sa_version(void);
/* Number of CUDA devices visible (0 when there is none / no driver). */
B200SA_API int b200sa_device_count(void);

/* ---- host-buffer entry points (what the reference-facing facade binds) ----------------- */

/* Replaces: msufsort::make_suffix_array(uint8_t const*, uint8_t const*) (msufsort.cpp:1730-1767;
 * template msufsort.h:432-445).  text: n bytes of host memory (pageable or pinned), not modified.
 * sa_out: n+1 int32.  n == 0 yields sa_out[0] = 0. */
B200SA_API int b200sa_suffix_array(b200sa_ctx* ctx, const uint8_t* text, int64_t n, int32_t* sa_out);

/* Replaces: msufsort::forward_burrows_wheeler_transform(uint8_t*, uint8_t*) (msufsort.cpp:1771-1817;
 * template msufsort.h:449-462).  In place over n host bytes; *sentinel_index_out in [1,n]
 * (0 when n == 0). */
B200SA_API int b200sa_bwt(b200sa_ctx* ctx, uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out);

/* Replaces: msufsort::reverse_burrows_wheeler_transform(uint8_t*, uint8_t*, int32 sentinelIndex,
 * int32 numThreads) (msufsort.cpp:1821-2096; template msufsort.h:466-476).  In place.
 * Untrusted input: bytes + sentinel index that are not the BWT of any text (a corrupted block) make the call fail
 * with B200SA_EINVAL and leave the caller's buffer unchanged; no access leaves the buffers (the reference returns
 * garbage of the right length in that case).  The same holds for the _dev, _u32 and batch forms; the device
 * output buffer of a failed call holds no valid text. */
B200SA_API int b200sa_unbwt(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int32_t sentinel_index);

/* Superset of the reference API: one suffix sort, both results (the reference needs its two
 * public calls, i.e. two sorts, for this).  sa_out and/or bwt_out may be NULL.
 * The host-buffer calls keep the last text and its suffix array resident in the context: when b200sa_bwt (or this
 * call) is handed, right after b200sa_suffix_array, the same bytes — compared on the device after the upload — it
 * reuses that sort.  Pageable host buffers are moved by several threads through pinned staging buffers
 * (B200SA_COPY_THREADS, default: half the host cores, 2..8); pinned / registered buffers are copied directly. */
B200SA_API int b200sa_suffix_array_bwt(b200sa_ctx* ctx, const uint8_t* text, int64_t n,
                                       int32_t* sa_out, uint8_t* bwt_out, int32_t* sentinel_index_out);

/* ---- device-resident entry points (kernel-only timing, pipelines, multi-GPU callers) ---- */

/* d_text: n bytes; d_sa_out: n+1 int32.  Work is enqueued on `stream`; the call returns after
 * the last doubling round has been observed complete (the round loop reads two counters per
 * round), so results are ready on return. */
B200SA_API int b200sa_suffix_array_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n,
                                       int32_t* d_sa_out, void* stream);

/* SA + BWT from one sort.  d_bwt_out: n bytes (must not alias d_text).  d_sa_out may be NULL
 * (the engine then keeps the SA in its workspace).  *sentinel_index_out is host memory. */
B200SA_API int b200sa_bwt_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n,
                              uint8_t* d_bwt_out, int32_t* d_sa_out,
                              int32_t* sentinel_index_out, void* stream);

/* Inverse BWT.  d_text_out: n bytes (must not alias d_bwt). */
B200SA_API int b200sa_unbwt_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n,
                                int32_t sentinel_index, uint8_t* d_text_out, void* stream);

/* O(n) validator (the role of validate_suffix_array, main.cpp:236-270, without its O(n*LCP)
 * byte compare): checks SA[0]==n, that SA is a permutation of [0,n], and for every pair of
 * neighbouring rows that T[SA[i]] <= T[SA[i+1]] with ties decided by ISA[SA[i]+1] < ISA[SA[i+1]+1].
 * *bad_rows_out receives the number of offending rows (0 = the SA is correct, hence — the SA
 * being unique — bit-identical to the reference's). */
B200SA_API int b200sa_check_suffix_array_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n,
                                             const int32_t* d_sa, int64_t* bad_rows_out, void* stream);

/* ---- wide-index superset -----------------------------------------------------------
 * The reference's suffix_index is int32 (msufsort.h:47) and its flag bits corrupt results above 2^30-2 bytes
 * (msufsort.h:84-93).  These entry points return the same suffix array with uint32 entries and the sentinel row as
 * int64, for texts up to B200SA_MAX_N_UINT32 bytes (e.g. the 2 GiB = 2^31-byte configuration); results for
 * n <= 2^31-2 are bit-identical to the int32 calls.  So do the inverse transform (its walkers are seeded at rows that are
 * multiples of a power of two, so the psi table needs no mark bit and all 32 bits of an entry are row number) and the LCP
 * array (uint32 values).  Batches keep block-local int32 indices and a combined size below 2^31-2; sharded runs stay at the
 * int32 limit. */
B200SA_API int b200sa_suffix_array_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, uint32_t* d_sa_out, void* stream);
B200SA_API int b200sa_bwt_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, uint8_t* d_bwt_out, uint32_t* d_sa_out,
                                  int64_t* sentinel_index_out, void* stream);
B200SA_API int b200sa_check_suffix_array_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const uint32_t* d_sa,
                                                 int64_t* bad_rows_out, void* stream);
/* Replaces reverse_burrows_wheeler_transform (msufsort.cpp:1821-2096) beyond its int32 sentinelIndex. */
B200SA_API int b200sa_unbwt_u32_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n, int64_t sentinel_index, uint8_t* d_text_out,
                                    void* stream);
B200SA_API int b200sa_unbwt_u32(b200sa_ctx* ctx, uint8_t* bwt_inout, int64_t n, int64_t sentinel_index);
/* LCP array with uint32 entries (the demo's lcp_multithreaded, main.cpp:16-105, beyond int32). */
B200SA_API int b200sa_lcp_u32_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const uint32_t* d_sa, uint32_t* d_lcp_out,
                                  void* stream);
/* Host buffers; sa_out (n+1 uint32) and / or bwt_out (n bytes, may alias text) and sentinel_index_out may be NULL. */
B200SA_API int b200sa_suffix_array_bwt_u32(b200sa_ctx* ctx, const uint8_t* text, int64_t n, uint32_t* sa_out, uint8_t* bwt_out,
                                           int64_t* sentinel_index_out);

/* LCP array.  Replaces the demo's LCP construction (src/executable/msufsort/main.cpp:16-105:
 * match_length, lcp, lcp_multithreaded), which is the only LCP code the reference ships.  Convention:
 * n+1 int32 entries aligned with the suffix array, lcp[0] = 0 and lcp[r] = length of the longest common
 * prefix of suffixes SA[r-1] and SA[r] (so lcp[1] = 0: SA[0] is the empty suffix).  The reference's
 * output[i] (main.cpp:150-153, i = 0..n-2) is lcp[i+2]; its last entry reads past its suffix array.
 * d_sa must be the correct suffix array of d_text (e.g. from b200sa_suffix_array_dev); work is
 * O(n log n) text probes in the worst case and independent of the text's repetitiveness. */
B200SA_API int b200sa_lcp_dev(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, const int32_t* d_sa,
                              int32_t* d_lcp_out, void* stream);
/* Host buffers.  sa == NULL: the suffix array is computed first (and returned through sa_out when that
 * is not NULL); otherwise sa must be the n+1-entry suffix array of text.  lcp_out: n+1 int32. */
B200SA_API int b200sa_lcp(b200sa_ctx* ctx, const uint8_t* text, int64_t n, const int32_t* sa,
                          int32_t* sa_out, int32_t* lcp_out);

/* ---- batches of independent blocks (north_star "batches of independent blocks") ----
 *
 * The reference transforms a batch by calling forward_/reverse_burrows_wheeler_transform once per block
 * (msufsort.h:63-75; the demo's round trip main.cpp:466-487).  Here `count` blocks are handed over packed back
 * to back: block b is bytes [offsets[b], offsets[b+1]) of `blocks`, offsets[0] = 0, empty blocks allowed,
 * offsets[count] + count <= 2^31-2.  All blocks are suffix-sorted TOGETHER by one launch sequence (block
 * number as the most significant key part, one separator slot per block), so thousands of small blocks cost
 * the same as one text of their total size; results are bit-identical to per-block calls.
 *   - suffix arrays: sa_out has offsets[count] + count entries; block b's n_b+1 entries (SA_b[0] = n_b)
 *     start at offsets[b] + b, values are block-local;
 *   - BWT: in place, packed like the input; sentinel_index_out[b] in [1, n_b] (0 for an empty block);
 *   - inverse: in place from (BWT bytes, sentinel indices); one sort of (block, byte) pairs builds the LF
 *     tables of all blocks, one walk decodes all blocks.                                                */
B200SA_API int b200sa_suffix_array_batch(b200sa_ctx* ctx, const uint8_t* blocks, const int64_t* offsets, int64_t count,
                                         int32_t* sa_out);
B200SA_API int b200sa_bwt_batch(b200sa_ctx* ctx, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                int32_t* sentinel_index_out);
B200SA_API int b200sa_unbwt_batch(b200sa_ctx* ctx, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                  const int32_t* sentinel_index);
/* Device-resident form: d_blocks packed as above (device), offsets and sentinel_index_out in host memory.
 * d_bwt_out (offsets[count] bytes, must not alias d_blocks), d_sa_out (offsets[count] + count int32) and
 * sentinel_index_out may each be NULL. */
B200SA_API int b200sa_batch_dev(b200sa_ctx* ctx, const uint8_t* d_blocks, const int64_t* offsets, int64_t count,
                                uint8_t* d_bwt_out, int32_t* d_sa_out, int32_t* sentinel_index_out, void* stream);

B200SA_API int b200sa_unbwt_batch_dev(b200sa_ctx* ctx, const uint8_t* d_bwt, const int64_t* offsets, int64_t count,
                                      const int32_t* sentinel_index, uint8_t* d_text_out, void* stream);

/* Streaming pipeline over batches ("pinned-buffer pipeline overlapping H2D, sort, D2H").
 * `depth` contexts with their own streams, workspaces and worker threads serve one queue of submitted batches,
 * so the upload of batch i+1 and the download of batch i-1 overlap the sort of batch i.  submit returns at once
 * with a ticket; the host buffers must stay valid (and untouched) until b200sa_pipeline_wait(ticket) returns
 * the job's status.  Pinned host memory makes the copies asynchronous.  Results are identical to the
 * b200sa_*_batch calls.  The reference has no counterpart (its callers loop over blocks, main.cpp:466-487). */
typedef struct b200sa_pipeline b200sa_pipeline;
B200SA_API int b200sa_pipeline_create(b200sa_pipeline** out, int device, int depth);
/* The same with `depth` contexts on EVERY listed device behind the one queue: the stream of batches spreads over all GPUs
 * and their PCIe links (nothing is exchanged between GPUs). */
B200SA_API int b200sa_pipeline_create_devices(b200sa_pipeline** out, const int* devices, int count, int depth);
B200SA_API void b200sa_pipeline_destroy(b200sa_pipeline* p);
B200SA_API int b200sa_pipeline_submit_bwt(b200sa_pipeline* p, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                          int32_t* sentinel_index_out, int64_t* ticket_out);
B200SA_API int b200sa_pipeline_submit_unbwt(b200sa_pipeline* p, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                            const int32_t* sentinel_index, int64_t* ticket_out);
B200SA_API int b200sa_pipeline_submit_suffix_array(b200sa_pipeline* p, const uint8_t* blocks, const int64_t* offsets,
                                                   int64_t count, int32_t* sa_out, int64_t* ticket_out);
B200SA_API int b200sa_pipeline_wait(b200sa_pipeline* p, int64_t ticket);
/* Waits for everything submitted so far; returns the first failure among tickets nobody waited for. */
B200SA_API int b200sa_pipeline_drain(b200sa_pipeline* p);

/* Host-buffer form of the O(n) validator (the role of validate_suffix_array, main.cpp:236-270, whose byte compares
 * are O(n * LCP)): *bad_rows_out = 0 iff sa is the suffix array of text. */
B200SA_API int b200sa_check_suffix_array(b200sa_ctx* ctx, const uint8_t* text, int64_t n, const int32_t* sa, int64_t* bad_rows_out);

/* ---- sharded (multi-GPU) building blocks --------------------------------------------------
 *
 * One text, G GPUs, one process and one context per GPU (msufsort_b200/sharded.py drives these over
 * torch.distributed / NCCL).  Every GPU holds the whole text and a replica of the ISA.  Round 0 is
 * partitioned by key range: each context packs all n initial keys, derives the same G-1 splitters
 * from a sorted regular sample (no communication) and sorts only the suffixes of its own range.
 * A range consists of whole groups, so every later doubling round sorts locally; what travels
 * between GPUs after each round is the list of (suffix, new rank) ISA updates (all-gather), which
 * every context applies to its replica.  The reference has no counterpart (single address space);
 * this is north_star item (4).                                                                   */

/* Alphabet, keys, splitters, key-range filter and first sort for part `part` of `nparts`.
 * d_sa: n+1 int32 (only this part's rows are written).  *n_local_out = suffixes owned. */
B200SA_API int b200sa_shard_begin(b200sa_ctx* ctx, const uint8_t* d_text, int64_t n, int32_t* d_sa,
                                  int part, int nparts, int64_t* n_local_out, void* stream);
/* Round-0 ranking.  slot_base = number of suffixes owned by lower parts (exclusive scan of the
 * n_local values, exchanged by the caller).  *m_local_out = suffixes of this part still active. */
B200SA_API int b200sa_shard_round0(b200sa_ctx* ctx, int64_t slot_base, int64_t* m_local_out, void* stream);
/* One doubling round on this part's active suffixes (a no-op when none are left). */
B200SA_API int b200sa_shard_round(b200sa_ctx* ctx, int64_t* m_local_out, void* stream);
/* ISA updates produced by the last round0/round call: device arrays valid until the next step. */
B200SA_API int b200sa_shard_updates(b200sa_ctx* ctx, const uint32_t** d_idx_out, const uint32_t** d_rank_out,
                                    int64_t* count_out);
/* Copies those updates into caller-owned device buffers (e.g. the send buffer of an all-gather). */
B200SA_API int b200sa_shard_copy_updates(b200sa_ctx* ctx, uint32_t* d_idx_dst, uint32_t* d_rank_dst,
                                         int64_t capacity, void* stream);
/* rank[d_idx[j]] = d_rank[j] on this context's ISA replica (own and peers' updates alike). */
B200SA_API int b200sa_shard_apply_updates(b200sa_ctx* ctx, const uint32_t* d_idx, const uint32_t* d_rank,
                                          int64_t count, void* stream);
/* Owner-sharded ISA (the scalable variant): GPU g is the authority for rank[] of the text positions
 * [g*B, (g+1)*B), B a power of two.  After a round the producer of a new rank sends it to the owner of
 * that suffix (all-to-all), and before the next round every GPU asks the owners for the ranks it is
 * about to read (all-to-all of positions, all-to-all of values) and drops the replies into its own
 * rank[] as a cache.  shard_partition routes pairs by owner (bucket = (key >> shift) & 255; a multi-split, not stable),
 * shard_requests lists the positions the next round reads, shard_gather_ranks serves a request list. */
B200SA_API int b200sa_shard_partition(b200sa_ctx* ctx, const uint32_t* d_keys, const uint32_t* d_vals, int64_t count,
                                      int shift, uint32_t* d_keys_out, uint32_t* d_vals_out, uint32_t* counts_out /*256, host*/,
                                      void* stream);
B200SA_API int b200sa_shard_requests(b200sa_ctx* ctx, uint32_t* d_pos_out, int64_t capacity, int64_t* count_out, void* stream);
B200SA_API int b200sa_shard_gather_ranks(b200sa_ctx* ctx, const uint32_t* d_pos, int64_t count, uint32_t* d_out, void* stream);
/* ISA in peer memory (the NVLink-native variant): GPU g owns rank[] of the positions [g << shift, (g+1) << shift) and
 * every GPU maps two allocations of each peer (CUDA IPC): its ISA array and its inbox.  Doubling rounds LOAD
 * rank[suffix + h] straight from the owner's HBM; new ranks are routed by owner with one radix sweep and STORED in
 * bulk into the owners' inboxes (coalesced stores over NVLink), then every owner applies its inbox locally.  No NCCL
 * all-to-all, no count exchange, no request/reply lookups: the caller only separates the phases of a round with
 * tiny all-reduces (done reading | sends landed | shards updated).
 *   export : allocates ISA array + inbox for an n-byte text, returns their two 64-byte IPC handles (128 bytes);
 *   attach : handles = nparts x 128 bytes (all-gathered by the caller; the own slot is ignored); mappings persist
 *            across sorts while the peers keep exporting the same allocations;
 *   layout : counts[g] = suffixes owned by GPU g in this sort (the n_local values): fixes the inbox regions;
 *   scatter: after shard_round0 / shard_round — route + send that step's (suffix, rank) pairs;
 *   apply  : after a barrier — scatter what arrived in the own inbox into the own ISA shard.
 * While a peer ISA is attached, shard_round and shard_bwt read ranks through it.                              */
B200SA_API int b200sa_shard_peer_export(b200sa_ctx* ctx, int64_t n, uint8_t* handles_out /*128*/);
B200SA_API int b200sa_shard_peer_attach(b200sa_ctx* ctx, int part, int nparts, int shift, int64_t n, const uint8_t* handles);
B200SA_API int b200sa_shard_peer_layout(b200sa_ctx* ctx, const int64_t* counts, int nparts);
B200SA_API int b200sa_shard_peer_scatter(b200sa_ctx* ctx, void* stream);
B200SA_API int b200sa_shard_peer_apply(b200sa_ctx* ctx, void* stream);
B200SA_API int b200sa_shard_peer_detach(b200sa_ctx* ctx);
/* BWT bytes of SA rows [row_begin,row_end) into d_bwt (an n-byte buffer; bytes
 * [*out_begin,*out_end) are written). */
B200SA_API int b200sa_shard_bwt(b200sa_ctx* ctx, int64_t row_begin, int64_t row_end, uint8_t* d_bwt,
                                int64_t* out_begin, int64_t* out_end, int32_t* sentinel_index_out, void* stream);

/* Inverse BWT over several GPUs: the psi table is built on every GPU (replicated), the walkers are
 * split.  build -> measure(my walker slice) -> exchange the (length, successor) entries of all slices
 * (segments, direction 0 = export mine, 1 = import a peer's) -> finish(my slice) writes my segments'
 * bytes into d_text_out (caller zero-fills it; the slices of all GPUs are disjoint, a sum all-reduce
 * assembles the text).  Single GPU: b200sa_unbwt_dev does build + measure(all) + finish(all).        */
B200SA_API int b200sa_unbwt_shard_build(b200sa_ctx* ctx, const uint8_t* d_bwt, int64_t n, int32_t sentinel_index,
                                        int64_t* nwalkers_out, void* stream);
B200SA_API int b200sa_unbwt_shard_measure(b200sa_ctx* ctx, int64_t w_begin, int64_t w_end, void* stream);
B200SA_API int b200sa_unbwt_shard_segments(b200sa_ctx* ctx, int direction, int64_t w_begin, int64_t w_end,
                                           uint32_t* d_len, uint32_t* d_next, void* stream);
B200SA_API int b200sa_unbwt_shard_finish(b200sa_ctx* ctx, int64_t w_begin, int64_t w_end, uint8_t* d_text_out, void* stream);

/* ---- one text over the GPUs of one box, driven from C++ (replaces the reference's static split of the
 * text over its worker threads, msufsort.cpp:1576-1586, 1635-1643, and is what its numThreads knob, msufsort.h:50-53,
 * 432-445, maps to here) ------------------------------------------------------------------------------------------
 * Every GPU holds the text and sorts one key range of the suffixes (whole groups, so all later sorts are local); the
 * inverse suffix array is sharded by text position and read / written through NVLink peer memory (CUDA IPC between
 * processes, plain peer pointers inside one process).  The control plane — two barriers and one sum per doubling round
 * — is a b200sa_comm: words in memory all ranks map, no collective library.
 *
 * b200sa_group_*: the whole thing behind one call with host buffers, one host thread + one context per listed device
 * (a device may be listed several times: that is how the single-GPU test tier drives the sharded path).  The results
 * leave the GPUs as disjoint slices over all PCIe links.  Texts shorter than 4096 bytes per GPU take one GPU. */
typedef struct b200sa_group b200sa_group;
B200SA_API int b200sa_group_create(b200sa_group** out, const int* devices, int count);
B200SA_API void b200sa_group_destroy(b200sa_group* g);
B200SA_API int b200sa_group_size(b200sa_group* g);
B200SA_API b200sa_ctx* b200sa_group_context(b200sa_group* g, int rank);
/* Replace make_suffix_array / forward_ / reverse_burrows_wheeler_transform exactly like the single-context calls above. */
B200SA_API int b200sa_group_suffix_array(b200sa_group* g, const uint8_t* text, int64_t n, int32_t* sa_out);
B200SA_API int b200sa_group_bwt(b200sa_group* g, uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out);
B200SA_API int b200sa_group_suffix_array_bwt(b200sa_group* g, const uint8_t* text, int64_t n, int32_t* sa_out, uint8_t* bwt_out,
                                             int32_t* sentinel_index_out);
B200SA_API int b200sa_group_unbwt(b200sa_group* g, uint8_t* bwt_inout, int64_t n, int32_t sentinel_index);
/* Batches of independent blocks over the GPUs of the group ("one group of blocks per GPU; no
 * communication; results concatenated on host").  Arguments and results exactly as b200sa_*_batch above; the packed batch
 * is cut into one contiguous run of blocks per GPU (about equal bytes), every GPU transforms its run by the single-GPU
 * batch path and writes the results at the blocks' own places in the caller's buffers.  Each RUN must satisfy the 32-bit
 * batch limit, so a group takes batches up to G times as large as one context.  The inverse writes nothing unless every
 * GPU has accepted its blocks.  Replaces a caller's loop over forward_/reverse_burrows_wheeler_transform
 * (msufsort.h:63-75, main.cpp:466-487) spread over the worker threads numThreads selects. */
B200SA_API int b200sa_group_suffix_array_batch(b200sa_group* g, const uint8_t* blocks, const int64_t* offsets, int64_t count,
                                               int32_t* sa_out);
B200SA_API int b200sa_group_bwt_batch(b200sa_group* g, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                      int32_t* sentinel_index_out);
B200SA_API int b200sa_group_unbwt_batch(b200sa_group* g, uint8_t* blocks_inout, const int64_t* offsets, int64_t count,
                                        const int32_t* sentinel_index);

/* The three reference calls without a context and with a GPU count — the C ABI SURVEY.md §8(b) sketched: GPUs
 * 0 .. num_gpus-1 (num_gpus <= 0: every GPU present); the group behind a count is created on first use and kept. */
B200SA_API int b200sa_suffix_array_gpus(const uint8_t* text, int64_t n, int32_t* sa_out, int num_gpus);
B200SA_API int b200sa_bwt_gpus(uint8_t* text_inout, int64_t n, int32_t* sentinel_index_out, int num_gpus);
B200SA_API int b200sa_unbwt_gpus(uint8_t* bwt_inout, int64_t n, int32_t sentinel_index, int num_gpus);

/* The same with the caller owning the ranks (one process per GPU under torchrun: msufsort_b200/sharded.py, bench.py).
 * b200sa_comm_create_local: nranks handles for the threads of one process.  b200sa_comm_create_shm: rank 0 creates the
 * POSIX shared-memory segment `name` ("/..."; pick a fresh name per job), the other processes of the node attach to it.
 * Collectives have a deadline (B200SA_COMM_TIMEOUT_MS, default 120 s) and fail with B200SA_ECOMM once any rank failed. */
typedef struct b200sa_comm b200sa_comm;
B200SA_API int b200sa_comm_create_local(b200sa_comm** out /* [nranks] */, int nranks);
B200SA_API int b200sa_comm_create_shm(b200sa_comm** out, const char* name, int rank, int nranks);
B200SA_API void b200sa_comm_destroy(b200sa_comm* comm);
B200SA_API int b200sa_comm_set_timeout_ms(b200sa_comm* comm, int timeout_ms);   /* deadline of this rank's waits from now on */
B200SA_API int b200sa_comm_barrier(b200sa_comm* comm);
B200SA_API int b200sa_comm_allreduce_sum(b200sa_comm* comm, int64_t value, int64_t* sum_out);
/* Collective over the ranks of `comm` (each with its own context and the same text in its HBM).  d_sa: n+1 int32, d_bwt
 * (may be NULL): n bytes — both full-size on every rank; on return this rank holds rows [info[0], info[1]) of the suffix
 * array and bytes [info[2], info[3]) of the BWT.  info_out: 8 words = row_begin, row_end, out_begin, out_end, sentinel
 * index, doubling rounds, bytes this rank stored into peer memory, suffixes owned. */
B200SA_API int b200sa_shard_sort(b200sa_ctx* ctx, b200sa_comm* comm, const uint8_t* d_text, int64_t n, int32_t* d_sa, uint8_t* d_bwt,
                                 int64_t* info_out, void* stream);
/* Collective inverse BWT: walkers split over the ranks, bytes stored into the owner of their text position over NVLink.
 * This rank owns text bytes [*slice_begin_out, *slice_end_out); they are copied to d_text_out (may be NULL) at their text
 * offsets.  gather_all != 0: the other slices are pulled from the peers as well, d_text_out holds the whole text. */
B200SA_API int b200sa_shard_unbwt(b200sa_ctx* ctx, b200sa_comm* comm, const uint8_t* d_bwt, int64_t n, int32_t sentinel_index,
                                  uint8_t* d_text_out, int gather_all, int64_t* slice_begin_out, int64_t* slice_end_out, void* stream);

/* ---- instrumentation --------------------------------------------------------------------- */

enum {
    B200SA_PH_ALPHABET = 0,   /* byte histogram of the text                                   */
    B200SA_PH_PACK = 1,       /* initial k-symbol key packing                                 */
    B200SA_PH_SORT_HIST = 2,  /* radix digit histograms                                       */
    B200SA_PH_SORT_PASS = 3,  /* radix scatter passes (decoupled look-back)  — dominant       */
    B200SA_PH_BUILD = 4,      /* (group, rank[i+h]) key build                                 */
    B200SA_PH_RERANK = 5,     /* head flags, segmented scan, ISA scatter, compaction          */
    B200SA_PH_BWT = 6,        /* BWT gather                                                   */
    B200SA_PH_UNBWT_BUILD = 7,/* unBWT: histogram + psi table                                 */
    B200SA_PH_UNBWT_WALK = 8, /* unBWT: walkers, list ranking, emit                           */
    B200SA_PH_CHECK = 9,      /* validator                                                    */
    B200SA_PH_SEGSORT = 10,   /* in-shared-memory sort of small groups (doubling rounds)      */
    B200SA_PH_ISA = 11,       /* bucketed ISA update: one radix sweep by suffix index + scatter */
    B200SA_PH_LCP = 12,       /* LCP array: PLCP levels + gather (the phi scatter is counted under ISA) */
    B200SA_PH_PEER_SEND = 13, /* sharded ISA: bulk stores of routed (suffix, rank) pairs into the owners' inboxes over NVLink */
    B200SA_PH_PEER_APPLY = 14,/* sharded ISA: the owner scatters its inbox into its ISA shard           */
    B200SA_PH_COUNT = 16
};

typedef struct b200sa_profile {
    /* accumulated since b200sa_profile_reset(); times are CUDA-event device milliseconds on the
     * launching stream and are only collected while profiling is enabled */
    double   ms[B200SA_PH_COUNT];
    uint64_t launches[B200SA_PH_COUNT];      /* kernel launches per phase (always counted)     */
    uint64_t alg_bytes[B200SA_PH_COUNT];     /* algorithmic bytes moved per phase  */
    uint64_t rounds;                         /* doubling rounds executed (round 0 included)    */
    uint64_t sort_passes;                    /* radix scatter passes executed                  */
    uint64_t sorted_tuples;                  /* sum over passes of tuples scattered            */
    uint64_t active_tuples;                  /* sum over rounds of active tuples               */
    uint64_t memsets;                        /* cudaMemsetAsync calls (not kernels of ours)    */
} b200sa_profile;

B200SA_API int b200sa_set_profiling(b200sa_ctx* ctx, int enabled);
B200SA_API int b200sa_profile_reset(b200sa_ctx* ctx);
B200SA_API int b200sa_profile_get(b200sa_ctx* ctx, b200sa_profile* out);
/* Total kernels of this library launched by this context since creation. */
B200SA_API uint64_t b200sa_launch_count(b200sa_ctx* ctx);

/* ---- building blocks exported for tests and benches -------------------------------------- */

/* Stable LSD radix sort of (u64 key, u32 value) pairs on bits [begin_bit, end_bit).
 * d_keys/d_vals hold the input and are clobbered; on return *result_in_alt is 0 when the sorted
 * data is in d_keys/d_vals and 1 when it is in d_keys_alt/d_vals_alt.  d_vals == NULL means
 * "values are the element indices 0..m-1". */
B200SA_API int b200sa_radix_sort_pairs_dev(b200sa_ctx* ctx, uint64_t* d_keys, uint64_t* d_keys_alt,
                                           uint32_t* d_vals, uint32_t* d_vals_alt, int64_t m,
                                           int begin_bit, int end_bit, int* result_in_alt, void* stream);

/* Test hook: packs and unpacks the rerank look-back descriptors for one triple; out5 = {kept, kept heads,
 * 1 + last head slot, flag bits of A, flag bits of B}. */
B200SA_API int b200sa_debug_rerank_descriptor(uint32_t kept, uint32_t kheads, uint32_t last_head1, uint64_t* out5);

#ifdef __cplusplus
}
#endif
#endif /* B200SA_H */
