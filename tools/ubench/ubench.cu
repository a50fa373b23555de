// ubench.cu — throughput of the warp-level primitives a radix-sort ranking step can be built from,
// measured on the target GPU (B200).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
// Output: SM-cycles per warp-instruction at full occupancy (lower is better).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

__global__ void k_vote(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc ^= __ballot_sync(0xffffffffu, (x >> u) & 1u); x += acc; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_vote_indep(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
        uint32_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ballot_sync(0xffffffffu, (x >> u) & 1u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u];
        x += acc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_match(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
        uint32_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __match_any_sync(0xffffffffu, (x >> (3 * u)) & 255u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u];
        x += acc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_shfl(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
        uint32_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __shfl_xor_sync(0xffffffffu, x + u, 1 + u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u];
        x += acc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_redux(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
        uint32_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __reduce_or_sync(0xffffffffu, x + u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u];
        x += acc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// shared memory: conflict-free LDS, random-bank LDS+STS read-modify-write on private words, atomics
__global__ void k_lds(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    uint32_t x = threadIdx.x, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += sm[(x + u * 32) & 4095];
        x += 256;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_rmw_u16_private(uint32_t* out, uint32_t seed) {
    // every lane owns a 256-entry u16 table (stride 258 halfwords: equal digits hit different banks)
    extern __shared__ uint16_t tab[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* mine = tab + (warp * 32 + lane) * 258;
    for (int i = 0; i < 258; ++i) mine[i] = 0;
    uint32_t x = threadIdx.x * 2654435761u + seed;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { x = x * 1664525u + 1013904223u; mine[x >> 24] += 1; }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 256; ++i) acc += mine[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_atoms(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[8][256];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) (&sm[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + seed;
    uint32_t* mine = sm[threadIdx.x >> 5 & 7];
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { x = x * 1664525u + 1013904223u; atomicAdd(&mine[x >> 24], 1u); }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[0][threadIdx.x & 255];
}
// atomics with `distinct` different addresses per warp instruction (1 = all lanes on one word)
template <int DISTINCT>
__global__ void k_atoms_skew(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[8][256];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) (&sm[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + seed;
    uint32_t* mine = sm[threadIdx.x >> 5 & 7];
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { x = x * 1664525u + 1013904223u; atomicAdd(&mine[((x >> 24) % DISTINCT) * 37 & 255], 1u); }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[0][threadIdx.x & 255];
}
// peer discovery through shared memory: atomicOr my lane bit, read the mask back, leader clears
__global__ void k_atomor_peers(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[8][256];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) (&sm[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    const uint32_t lane = threadIdx.x & 31, bit = 1u << lane, lt = bit - 1;
    uint32_t* mine = sm[threadIdx.x >> 5 & 7];
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint32_t d = x >> 24;
            atomicOr(&mine[d], bit);
            __syncwarp();
            const uint32_t peers = mine[d];
            __syncwarp();
            if ((peers & lt) == 0) mine[d] = 0;
            __syncwarp();
            acc += __popc(peers & lt);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_ballot_peers(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint32_t d = x >> 24;
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) { const uint32_t bt = (d >> b) & 1u; peers &= __ballot_sync(0xffffffffu, bt) ^ (bt - 1u); }
            acc += __popc(peers & lt);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_popc_lop(uint32_t* out, uint32_t seed) {
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    for (int i = 0; i < ITERS / UNROLL; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc += __popc(x ^ (acc + u)); x = x * 5 + 1; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename K>
static void run(const char* name, K kern, int threads, size_t smem, uint32_t* d_out, int sms, double clock_ghz, int blocks_per_sm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * blocks_per_sm;
    kern<<<grid, threads, smem>>>(d_out, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    kern<<<grid, threads, smem>>>(d_out, 2u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double warp_instr_per_sm = (double)blocks_per_sm * (threads / 32) * ITERS;
    const double cycles = ms * 1e-3 * clock_ghz * 1e9;
    printf("%-22s %8.3f ms  %7.2f SM-cycles per warp-instruction  (%d warps/SM)%s\n", name, ms, cycles / warp_instr_per_sm,
           blocks_per_sm * threads / 32, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("%s, %d SMs, %.3f GHz (max); numbers assume the max clock\n", p.name, p.multiProcessorCount, ghz);
    uint32_t* d_out; cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 8 * 1024 * 4);
    const int sms = p.multiProcessorCount;
    run("vote.ballot (dependent)", k_vote, 256, 0, d_out, sms, ghz, 8);
    run("vote.ballot (indep x8)", k_vote_indep, 256, 0, d_out, sms, ghz, 8);
    run("match.any", k_match, 256, 0, d_out, sms, ghz, 8);
    run("shfl.xor", k_shfl, 256, 0, d_out, sms, ghz, 8);
    run("redux.or", k_redux, 256, 0, d_out, sms, ghz, 8);
    run("lds (conflict-free)", k_lds, 256, 0, d_out, sms, ghz, 8);
    run("lds+sts u16 private rmw", k_rmw_u16_private, 128, 128 * 258 * 2, d_out, sms, ghz, 3);
    run("atoms.add (8 tables)", k_atoms, 256, 0, d_out, sms, ghz, 8);
    run("atoms.add 1 address", k_atoms_skew<1>, 256, 0, d_out, sms, ghz, 8);
    run("atoms.add 2 addresses", k_atoms_skew<2>, 256, 0, d_out, sms, ghz, 8);
    run("atoms.add 4 addresses", k_atoms_skew<4>, 256, 0, d_out, sms, ghz, 8);
    run("atoms.add 16 addresses", k_atoms_skew<16>, 256, 0, d_out, sms, ghz, 8);
    run("peers: atomicOr+lds+clr", k_atomor_peers, 256, 0, d_out, sms, ghz, 8);
    run("peers: 8 ballots", k_ballot_peers, 256, 0, d_out, sms, ghz, 8);
    run("popc+lop+imad", k_popc_lop, 256, 0, d_out, sms, ghz, 8);
    return 0;
}
