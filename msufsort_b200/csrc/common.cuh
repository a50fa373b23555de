// common.cuh — shared definitions for the B200 (sm_100a) suffix-array engine.
//
// The same sources build two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> msufsort_b200/lib/libb200sa.so (the product)
//   * g++ -DB200SA_EMU (tests/emu/cuda_emu.h)       -> tests/emu/libb200sa_emu.so (CPU-only logic
//     tests; never loaded by the product package)
#pragma once

#include <stdint.h>
#include <stddef.h>

#ifdef B200SA_EMU
#include "cuda_emu.h"
#define B200SA_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define B200SA_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#else
#include <cuda_runtime.h>
#define B200SA_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200SA_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#define B200SA_FULL_MASK 0xffffffffu

namespace b200sa {

static const int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ __forceinline__ u64 div_up_u64(u64 a, u64 b) { return (a + b - 1) / b; }

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// Relaxed, GPU-scope accesses for the single-word (flag|value) look-back descriptors: the flag and
// the payload travel in one 64-bit word, so no fence is needed between them.
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p)
{
#ifdef B200SA_EMU
    return *(const volatile u64*)p;
#else
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v)
{
#ifdef B200SA_EMU
    *(volatile u64*)p = v;
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// Streaming (evict-first) accesses for data that is read or written exactly once per kernel.
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p)
{
#ifdef B200SA_EMU
    return *p;
#else
    return __ldcs(p);
#endif
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v)
{
#ifdef B200SA_EMU
    *p = v;
#else
    __stcs(p, v);
#endif
}

// Lanes of the warp holding the same 8-bit digit as the caller, built from eight ballots:
// peers = AND_b ( bit_b(d) ? ballot(bit_b) : ~ballot(bit_b) ).  MATCH.ANY does the same in one
// instruction but measured ~59 SM-cycles per warp-instruction on B200 (ncu: the consumer of its
// result was the top stall of both the histogram and the scatter kernel); eight VOTEs are ~6x cheaper.
__device__ __forceinline__ u32 warp_peers_digit8(u32 d)
{
    u32 peers = B200SA_FULL_MASK;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
#ifdef B200SA_EMU
        const u32 bit = (d >> b) & 1u;
        const u32 vote = __ballot_sync(B200SA_FULL_MASK, bit);
        peers &= vote ^ (bit - 1u);  // bit ? vote : ~vote
#else
        // Written in PTX so that the bit test feeds the vote as a predicate: ptxas then loads seven predicates
        // with one R2P and uses predicated NOTs — 31 SASS instructions per key instead of 53 for the C form
        // above (which tests the bit, rebuilds a 0/1 word, compares it again and subtracts).
        u32 vote, flip;
        asm volatile(
            "{ .reg .pred p; .reg .u32 t;\n\t"
            "and.b32 t, %2, %3;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
            "selp.u32 %1, 0, 0xffffffff, p; }"
            : "=r"(vote), "=r"(flip)
            : "r"(d), "r"(1u << b));
        peers &= vote ^ flip;
#endif
    }
    return peers;
}

// Where a doubling round reads rank[p] (the inverse suffix array) from.  LocalRank: this GPU's array.  PeerRank:
// the array is sharded over the GPUs of the box — text position p belongs to GPU p >> shift, every GPU maps the
// arrays of all peers (CUDA IPC over NVLink / NVSwitch) and kernels load and store peer memory directly; no
// collective moves ranks around (msufsort_b200/sharded.py, isa="peer").
static const int kMaxPeers = 16;
struct RankView {
    u32* base[kMaxPeers];
    int shift;
    u32 n;
};
struct LocalRank {
    const u32* r;
    u32 n;
    __device__ __forceinline__ u32 operator()(u64 p) const { return r[p < n ? p : n]; }  // r[n] = 0: the empty suffix
};
struct PeerRank {
    RankView v;
    __device__ __forceinline__ u32 operator()(u64 p) const { return p < v.n ? v.base[(u32)p >> v.shift][(u32)p] : 0u; }
};

// Inclusive warp scan (sum) over 32 lanes.
__device__ __forceinline__ u32 warp_incl_scan_u32(u32 v)
{
    const u32 lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(B200SA_FULL_MASK, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_scan_u64(u64 v)
{
    const u32 lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(B200SA_FULL_MASK, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u32 warp_incl_scan_max_u32(u32 v)
{
    const u32 lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(B200SA_FULL_MASK, v, d);
        if (lane >= (u32)d) v = v > t ? v : t;
    }
    return v;
}

}  // namespace b200sa
