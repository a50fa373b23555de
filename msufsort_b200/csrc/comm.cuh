// comm.cuh — the control plane between the ranks that shard ONE text over the GPUs of one box.
//
// The data plane of a sharded sort is NVLink peer memory (engine_peer.inl: every GPU loads rank[suffix + h] from, and
// stores new ranks in bulk into, the owner's HBM).  What is left for the hosts is a few words per doubling round:
// "all my sends have landed", "all shards are current", the number of suffixes still active.  Round 1 drove that from
// Python with three NCCL all-reduces + .item() per round (2.6 of 10.7 ms per step on eight GPUs); here it is a
// sense-reversing barrier and a table of per-rank slots in memory all ranks map: the heap when the ranks are threads
// of one process (b200sa_group_*, the facade's MSUFSORT_NUM_GPUS), a POSIX shared-memory segment when they are the
// processes of one torchrun launch (one process per GPU, one node).  No collective library is involved.
//
// Every wait has a deadline and watches a shared error word: a rank that fails raises it, its peers leave their
// barriers with B200SA_ECOMM instead of hanging.
#pragma once
#include "common.cuh"
#include "../../include/b200sa.h"

#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <thread>
#include <cstring>
#include <cstdlib>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace b200sa {

int set_error(int code, const char* fmt, ...);

static const int kCommSlotBytes = 2048;
static const uint32_t kCommMagic = 0xB2005A01u;

struct CommShared {
    std::atomic<uint32_t> magic;
    uint32_t nranks;
    std::atomic<uint32_t> arrive;
    std::atomic<uint32_t> generation;
    std::atomic<uint32_t> error;
    alignas(64) unsigned char slots[2][kMaxPeers][kCommSlotBytes];  // double-buffered by the parity of the collective's number
};

struct Comm {
    CommShared* sh = nullptr;
    int rank = 0, nranks = 1;
    bool shm = false;                          // ranks are processes (CUDA IPC for peer memory) / threads (plain pointers)
    std::shared_ptr<CommShared> local_owner;   // thread mode
    size_t map_bytes = 0;                      // shm mode
    uint32_t seq = 0;                          // collectives issued so far (same on every rank)
    int timeout_ms = 120000;

    void raise_error() { if (sh) sh->error.store(1u, std::memory_order_release); }

    int barrier()
    {
        if (nranks <= 1) return 0;
        if (sh->error.load(std::memory_order_acquire)) return set_error(B200SA_ECOMM, "a peer rank failed");
        const uint32_t gen = sh->generation.load(std::memory_order_acquire);
        if (sh->arrive.fetch_add(1u, std::memory_order_acq_rel) + 1u == (uint32_t)nranks) {
            sh->arrive.store(0u, std::memory_order_relaxed);
            sh->generation.store(gen + 1u, std::memory_order_release);
            return 0;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t spins = 1;; ++spins) {
            if (sh->generation.load(std::memory_order_acquire) != gen) return 0;
            if ((spins & 255u) == 0) {
                if (sh->error.load(std::memory_order_acquire)) return set_error(B200SA_ECOMM, "a peer rank failed");
                if ((spins & 0xffffu) == 0 &&
                    std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) {
                    raise_error();
                    return set_error(B200SA_ECOMM, "rank %d waited more than %d ms for its peers at a barrier", rank, timeout_ms);
                }
                if (spins > 20000u) std::this_thread::yield();
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }

    // every rank contributes `bytes` (<= kCommSlotBytes); all receives nranks * bytes in rank order.  One barrier: the slots of
    // collective k are not written again before collective k + 2, which no rank enters before all have left collective k + 1.
    int allgather(const void* mine, size_t bytes, void* all)
    {
        if (bytes > (size_t)kCommSlotBytes) return set_error(B200SA_EINTERNAL, "comm payload too large");
        if (nranks <= 1) { memcpy(all, mine, bytes); return 0; }
        const int par = (int)(seq++ & 1u);
        memcpy(sh->slots[par][rank], mine, bytes);
        const int rc = barrier();
        if (rc) return rc;
        for (int r = 0; r < nranks; ++r) memcpy((char*)all + (size_t)r * bytes, sh->slots[par][r], bytes);
        return 0;
    }

    int allreduce_sum(i64 v, i64* out)
    {
        i64 all[kMaxPeers];
        const int rc = allgather(&v, sizeof(v), all);
        if (rc) return rc;
        i64 s = 0;
        for (int r = 0; r < nranks; ++r) s += all[r];
        *out = s;
        return 0;
    }
};

// ranks = threads of this process: `out` receives nranks handles on one shared block
static inline int comm_create_local(Comm** out, int nranks)
{
    if (!out || nranks < 1 || nranks > kMaxPeers) return set_error(B200SA_EINVAL, "between 1 and %d ranks", kMaxPeers);
    std::shared_ptr<CommShared> sh(new (std::nothrow) CommShared());
    if (!sh) return set_error(B200SA_ENOMEM, "out of host memory");
    memset((void*)sh.get(), 0, sizeof(CommShared));
    sh->nranks = (uint32_t)nranks;
    sh->magic.store(kCommMagic);
    for (int r = 0; r < nranks; ++r) {
        Comm* c = new (std::nothrow) Comm();
        if (!c) { for (int q = 0; q < r; ++q) delete out[q]; return set_error(B200SA_ENOMEM, "out of host memory"); }
        c->sh = sh.get(); c->local_owner = sh; c->rank = r; c->nranks = nranks;
        out[r] = c;
    }
    return 0;
}

// ranks = processes of one node: rank 0 creates the segment `name` ("/something"), the others wait for it
static inline int comm_create_shm(Comm** out, const char* name, int rank, int nranks)
{
    if (!out || !name || name[0] != '/' || nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks)
        return set_error(B200SA_EINVAL, "bad argument (name must start with '/', at most %d ranks)", kMaxPeers);
    *out = nullptr;
    const size_t bytes = sizeof(CommShared);
    int fd = -1;
    const auto t0 = std::chrono::steady_clock::now();
    auto waited_ms = [&] { return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count(); };
    if (rank == 0) {
        shm_unlink(name);  // a stale segment of a crashed run
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { if (fd >= 0) close(fd); return set_error(B200SA_ECOMM, "cannot create shared memory segment %s", name); }
    } else {
        for (;;) {
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && (size_t)sb.st_size >= bytes) break;
                close(fd);
                fd = -1;
            }
            if (waited_ms() > 120000) return set_error(B200SA_ECOMM, "rank %d: shared memory segment %s did not appear", rank, name);
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return set_error(B200SA_ECOMM, "mmap of %s failed", name);
    CommShared* sh = (CommShared*)p;
    if (rank == 0) {
        sh->nranks = (uint32_t)nranks;  // fresh segments are zero-filled
        sh->magic.store(kCommMagic, std::memory_order_release);
    } else {
        while (sh->magic.load(std::memory_order_acquire) != kCommMagic) {
            if (waited_ms() > 120000) { munmap(p, bytes); return set_error(B200SA_ECOMM, "rank %d: segment %s was never initialised", rank, name); }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        if (sh->nranks != (uint32_t)nranks) { munmap(p, bytes); return set_error(B200SA_ECOMM, "segment %s was created for %u ranks", name, sh->nranks); }
    }
    Comm* c = new (std::nothrow) Comm();
    if (!c) { munmap(p, bytes); return set_error(B200SA_ENOMEM, "out of host memory"); }
    c->sh = sh; c->rank = rank; c->nranks = nranks; c->shm = true; c->map_bytes = bytes;
    if (const char* e = getenv("B200SA_COMM_TIMEOUT_MS")) { const int v = atoi(e); if (v > 0) c->timeout_ms = v; }
    const int rc = c->barrier();       // everybody has mapped it ...
    if (rank == 0) shm_unlink(name);   // ... so the name can go; the mappings stay
    if (rc) { munmap(p, bytes); delete c; return rc; }
    *out = c;
    return 0;
}

static inline void comm_destroy(Comm* c)
{
    if (!c) return;
    if (c->shm && c->sh) munmap((void*)c->sh, c->map_bytes);
    delete c;
}

}  // namespace b200sa
