/*
 * cuda_emu.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small single-threaded SIMT emulator so that the kernel LOGIC of msufsort_b200/csrc/*.cu(h)
 * (tile indexing, warp-level ranking, look-back chaining, scans, walkers) can be exercised by the
 * CPU-only test tier (`pytest -m "not gpu"`) in a container without a GPU.  The sources are
 * compiled as plain C++ with -DB200SA_EMU into tests/emu/libb200sa_emu.so.  That library is only
 * ever loaded by tests/; the product package loads msufsort_b200/lib/libb200sa.so (nvcc, sm_100a)
 * and fails loudly when it or a CUDA device is missing.  The emulator proves nothing about races,
 * memory ordering or performance — the `-m gpu` tier and compute-sanitizer do that.
 *
 * Model: blocks run one after another on the calling thread; every CUDA thread of a block is a
 * fiber (hand-rolled x86-64 context switch); __syncthreads and the *_sync warp collectives are
 * rendezvous points.  Tile ids handed out by atomicAdd therefore complete in order, which is what
 * decoupled look-back needs to make progress.
 */
#pragma once
#ifndef B200SA_EMU
#error "cuda_emu.h is only for -DB200SA_EMU builds"
#endif
#if !defined(__x86_64__)
#error "the emulator's context switch is x86-64 only"
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <mutex>
#include <vector>

/* ---- qualifiers ------------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(x) __attribute__((aligned(x)))

/* ---- vector types ----------------------------------------------------------------------- */
// (alignments as in CUDA's vector_types.h: a misaligned 64- / 128-bit access faults on the GPU, and the UBSan build reports it)
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
struct __attribute__((aligned(4))) uchar4 { unsigned char x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {

enum { ST_READY = 0, ST_WAIT_WARP = 1, ST_WAIT_BLOCK = 2, ST_DONE = 3 };
enum { OP_SHFL_IDX, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL, OP_MATCH_ANY, OP_SYNCWARP };

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    int state = ST_DONE;
    unsigned tid = 0;
};
struct Warp {
    uint32_t arrived = 0;
    int op = -1;
    uint64_t val[32];
    int param[32];
    int width[32];
    uint64_t res[32];
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned nthreads = 0;
    unsigned live = 0;
    unsigned bar_arrived = 0;
    Fiber* cur = nullptr;
    void* sched_sp = nullptr;
    const std::function<void()>* body = nullptr;
    std::vector<unsigned char> dyn_smem;
    uint64_t launches = 0;
};

inline State& S() { static State s; return s; }

extern "C" void b200sa_emu_switch(void** from_sp, void* to_sp);

inline uint3& tid_ref() { static uint3 v; return v; }
inline uint3& bid_ref() { static uint3 v; return v; }
inline dim3& bdim_ref() { static dim3 v; return v; }
inline dim3& gdim_ref() { static dim3 v; return v; }

static const size_t kStackBytes = 256 * 1024;

inline void yield_to_sched()
{
    State& s = S();
    Fiber* f = s.cur;
    b200sa_emu_switch(&f->sp, s.sched_sp);
}

inline void fiber_main()
{
    State& s = S();
    (*s.body)();
    s.cur->state = ST_DONE;
    s.live--;
    /* a thread that exits may complete a pending block barrier */
    if (s.live > 0 && s.bar_arrived == s.live) {
        s.bar_arrived = 0;
        for (unsigned t = 0; t < s.nthreads; ++t)
            if (s.fibers[t].state == ST_WAIT_BLOCK) s.fibers[t].state = ST_READY;
    }
    yield_to_sched();
    fprintf(stderr, "cuda_emu: resumed a finished fiber\n");
    abort();
}
extern "C" inline void b200sa_emu_entry() { fiber_main(); }

inline void prepare_fiber(Fiber& f, unsigned tid)
{
    if (!f.stack) f.stack = (char*)aligned_alloc(64, kStackBytes);
    f.tid = tid;
    f.state = ST_READY;
    uintptr_t top = ((uintptr_t)(f.stack + kStackBytes)) & ~(uintptr_t)15;
    uint64_t* p = (uint64_t*)top;
    *--p = 0;                                  /* fake return address of the entry function      */
    *--p = (uint64_t)(void*)&b200sa_emu_entry; /* 'ret' target of the first switch                 */
    for (int i = 0; i < 6; ++i) *--p = 0;      /* rbp rbx r12 r13 r14 r15                          */
    f.sp = (void*)p;
}

inline void run_block(unsigned nthreads)
{
    State& s = S();
    if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
    unsigned nwarps = (nthreads + 31) / 32;
    if (s.warps.size() < nwarps) s.warps.resize(nwarps);
    for (unsigned w = 0; w < nwarps; ++w) { s.warps[w].arrived = 0; s.warps[w].op = -1; }
    s.nthreads = nthreads;
    s.live = nthreads;
    s.bar_arrived = 0;
    for (unsigned t = 0; t < nthreads; ++t) prepare_fiber(s.fibers[t], t);
    for (;;) {
        bool progressed = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            if (f.state != ST_READY) continue;
            progressed = true;
            s.cur = &f;
            tid_ref().x = t; tid_ref().y = 0; tid_ref().z = 0;
            b200sa_emu_switch(&s.sched_sp, f.sp);
        }
        if (s.live == 0) break;
        if (!progressed) {
            fprintf(stderr, "cuda_emu: DEADLOCK in block (%u,%u): live=%u bar_arrived=%u\n", bid_ref().x, bid_ref().y, s.live, s.bar_arrived);
            for (unsigned t = 0; t < nthreads && t < 64; ++t) fprintf(stderr, " t%u:%d", t, s.fibers[t].state);
            fprintf(stderr, "\n");
            abort();
        }
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    /* one kernel at a time, whatever host thread launches it (b200sa_group_* runs one host thread per context) */
    static std::mutex launch_mu;
    std::lock_guard<std::mutex> launch_lock(launch_mu);
    State& s = S();
    if (s.cur && s.cur->state != ST_DONE && s.live) { fprintf(stderr, "cuda_emu: nested launch\n"); abort(); }
    s.body = &body;
    s.launches++;
    if (s.dyn_smem.size() < smem + 16) s.dyn_smem.resize(smem + 16);
    bdim_ref() = block;
    gdim_ref() = grid;
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            bid_ref().x = bx; bid_ref().y = by; bid_ref().z = 0;
            run_block(nthreads);
        }
    s.cur = nullptr;
}

inline void syncthreads()
{
    State& s = S();
    s.bar_arrived++;
    if (s.bar_arrived == s.live) {
        s.bar_arrived = 0;
        for (unsigned t = 0; t < s.nthreads; ++t)
            if (s.fibers[t].state == ST_WAIT_BLOCK) s.fibers[t].state = ST_READY;
        return;
    }
    s.cur->state = ST_WAIT_BLOCK;
    yield_to_sched();
}

inline uint64_t warp_collective(int op, unsigned mask, uint64_t v, int param, int width)
{
    State& s = S();
    unsigned tid = s.cur->tid;
    unsigned lane = tid & 31, wid = tid >> 5;
    Warp& w = s.warps[wid];
    /* lanes beyond the block size do not exist */
    unsigned lanes_here = s.nthreads - wid * 32 >= 32 ? 32 : s.nthreads - wid * 32;
    unsigned exist = lanes_here == 32 ? 0xffffffffu : ((1u << lanes_here) - 1u);
    mask &= exist;
    if (!(mask & (1u << lane))) { fprintf(stderr, "cuda_emu: lane %u not in its own mask %08x\n", lane, mask); abort(); }
    if (w.arrived != 0 && w.op != op) { fprintf(stderr, "cuda_emu: divergent collectives (op %d vs %d) in warp %u\n", w.op, op, wid); abort(); }
    w.op = op;
    w.val[lane] = v;
    w.param[lane] = param;
    w.width[lane] = width;
    w.arrived |= 1u << lane;
    if (w.arrived == mask) {
        uint64_t ballot = 0;
        for (unsigned l = 0; l < 32; ++l) if ((mask >> l) & 1) if (w.val[l]) ballot |= 1ull << l;
        for (unsigned l = 0; l < 32; ++l) {
            if (!((mask >> l) & 1)) continue;
            int wd = w.width[l] ? w.width[l] : 32;
            int base = (int)l & ~(wd - 1);
            uint64_t r = w.val[l];
            switch (op) {
            case OP_SHFL_IDX: { int src = base | (w.param[l] & (wd - 1)); if ((mask >> src) & 1) r = w.val[src]; break; }
            case OP_SHFL_UP: { int src = (int)l - w.param[l]; if (src >= base && ((mask >> src) & 1)) r = w.val[src]; break; }
            case OP_SHFL_DOWN: { int src = (int)l + w.param[l]; if (src <= (base | (wd - 1)) && src < 32 && ((mask >> src) & 1)) r = w.val[src]; break; }
            case OP_SHFL_XOR: { int src = (int)l ^ w.param[l]; if (src < 32 && ((mask >> src) & 1)) r = w.val[src]; break; }
            case OP_BALLOT: r = ballot; break;
            case OP_ANY: r = ballot != 0; break;
            case OP_ALL: r = (ballot == (uint64_t)mask); break;
            case OP_MATCH_ANY: { uint64_t m = 0; for (unsigned k = 0; k < 32; ++k) if (((mask >> k) & 1) && w.val[k] == w.val[l]) m |= 1ull << k; r = m; break; }
            default: r = 0; break;
            }
            w.res[l] = r;
        }
        w.arrived = 0;
        w.op = -1;
        for (unsigned l = 0; l < 32; ++l)
            if (((mask >> l) & 1) && l != lane) {
                Fiber& f = s.fibers[wid * 32 + l];
                if (f.state == ST_WAIT_WARP) f.state = ST_READY;
            }
        return w.res[lane];
    }
    s.cur->state = ST_WAIT_WARP;
    yield_to_sched();
    return s.warps[wid].res[lane];
}

inline unsigned char* dyn_smem() { return (unsigned char*)(((uintptr_t)S().dyn_smem.data() + 15) & ~(uintptr_t)15); }

}  // namespace emu

/* the context switch: save callee-saved registers + stack pointer, load the other side's */
__asm__(
    ".text\n"
    ".globl b200sa_emu_switch\n"
    ".type b200sa_emu_switch,@function\n"
    "b200sa_emu_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size b200sa_emu_switch,.-b200sa_emu_switch\n");

#define threadIdx (emu::tid_ref())
#define blockIdx (emu::bid_ref())
#define blockDim (emu::bdim_ref())
#define gridDim (emu::gdim_ref())
#define warpSize 32

/* ---- synchronisation and warp collectives ------------------------------------------------ */
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(emu::OP_SYNCWARP, mask, 0, 0, 32); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline T emu_from_u64(uint64_t v) { T t; memcpy(&t, &v, sizeof(T)); return t; }
template <typename T> static inline uint64_t emu_to_u64(T t) { uint64_t v = 0; static_assert(sizeof(T) <= 8, "shfl type"); memcpy(&v, &t, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) { return emu_from_u64<T>(emu::warp_collective(emu::OP_SHFL_IDX, mask, emu_to_u64(v), src, width)); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) { return emu_from_u64<T>(emu::warp_collective(emu::OP_SHFL_UP, mask, emu_to_u64(v), (int)d, width)); }
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) { return emu_from_u64<T>(emu::warp_collective(emu::OP_SHFL_DOWN, mask, emu_to_u64(v), (int)d, width)); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) { return emu_from_u64<T>(emu::warp_collective(emu::OP_SHFL_XOR, mask, emu_to_u64(v), m, width)); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)emu::warp_collective(emu::OP_BALLOT, mask, pred ? 1 : 0, 0, 32); }
static inline int __any_sync(unsigned mask, int pred) { return (int)emu::warp_collective(emu::OP_ANY, mask, pred ? 1 : 0, 0, 32); }
static inline int __all_sync(unsigned mask, int pred) { return (int)emu::warp_collective(emu::OP_ALL, mask, pred ? 1 : 0, 0, 32); }
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) { return (unsigned)emu::warp_collective(emu::OP_MATCH_ANY, mask, emu_to_u64(v), 0, 32); }

/* ---- integer intrinsics ------------------------------------------------------------------ */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t src = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((src >> (8 * sel)) & 0xff) << (8 * i); }
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (sh & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (sh & 31)) >> 32); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

/* ---- atomics (single host thread: plain read-modify-write) ------------------------------- */
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <typename T> static inline T atomicSub(T* p, T v) { T o = *p; *p = (T)(o - v); return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = (T)(o | v); return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = (T)(o & v); return o; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

/* ---- runtime API subset ------------------------------------------------------------------ */
typedef int cudaError_t;
typedef struct emu_stream_s* cudaStream_t;
typedef struct emu_event_s { double t_ms; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int multiProcessorCount; int major; int minor; size_t totalGlobalMem; char name[256]; size_t sharedMemPerBlockOptin; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { memset(p, 0, sizeof(*p)); p->multiProcessorCount = 148; p->major = 10; p->minor = 0; p->totalGlobalMem = (size_t)8 << 30; p->sharedMemPerBlockOptin = 227 * 1024; strcpy(p->name, "cuda_emu"); return 0; }
#ifdef B200SA_EMU_ASAN
// AddressSanitizer build (make emu-asan): exact sizes, so that a kernel reading or writing one element past a buffer is reported
// (rounded up to whole 32-bit words: device allocations never end inside an aligned word, which the word-wise text loads rely on)
static inline cudaError_t cudaMalloc(void** p, size_t n) { n = (n + 3) & ~(size_t)3; *p = nullptr; if (posix_memalign(p, 256, n ? n : 4) != 0) *p = nullptr; if (*p) memset(*p, 0xA5, n); return *p ? 0 : cudaErrorMemoryAllocation; }
#else
static inline cudaError_t cudaMalloc(void** p, size_t n) { size_t r = (n + 255) & ~(size_t)255; *p = aligned_alloc(256, r); if (*p) memset(*p, 0xA5, r); /* poison: device memory is never zero-initialised */ return *p ? 0 : cudaErrorMemoryAllocation; }
#endif
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
// "IPC" inside one process: the handle carries the pointer (tests drive several contexts of one process in lock step)
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return 0; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return 0; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(8); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(sizeof(emu_event_s)); (*e)->t_ms = 0; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t_ms = emu_now_ms(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)8 << 30; *t = (size_t)8 << 30; return 0; }
