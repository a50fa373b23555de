// maniscalco::msufsort — marshalling layer from the reference's public C++ interface to the C ABI
// (include/b200sa.h).  No algorithmic code lives here; see msufsort_b200/csrc for the kernels.
#include "./msufsort.h"

#include <b200sa.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#include <unistd.h>
#endif

namespace
{
    // The suffix array is returned by value as a std::vector<int32_t>(n + 1) (the reference's interface, msufsort.h:48, 57-61),
    // so every call allocates and zero-fills 4 (n + 1) bytes of fresh memory: 1 GiB for a 256 MiB text, one page fault per
    // 4 KiB — measured 390 ms, six times the GPU work.  Asking for transparent huge pages on the reserved storage before it
    // is touched turns that into one fault per 2 MiB where the kernel allows it (a hint: no effect otherwise).
    void advise_huge_pages(void * data, std::size_t bytes)
    {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
        if (bytes < (std::size_t(8) << 20))
            return;
        std::uintptr_t const page = 4096;
        std::uintptr_t const begin = (reinterpret_cast<std::uintptr_t>(data) + page - 1) & ~(page - 1);
        std::uintptr_t const end = (reinterpret_cast<std::uintptr_t>(data) + bytes) & ~(page - 1);
        if (end > begin)
            (void)madvise(reinterpret_cast<void *>(begin), end - begin, MADV_HUGEPAGE);
#else
        (void)data;
        (void)bytes;
#endif
    }

    // What one msufsort object computes with: one GPU context, or a group of contexts that shards every text over several
    // GPUs (MSUFSORT_NUM_GPUS).  The reference's free templates build a short-lived msufsort object per call
    // (msufsort.h:432-476: a fresh worker pool every time); here such objects borrow a backend from a process-wide pool, so
    // the device workspace — and the suffix array kept resident for a following forward_burrows_wheeler_transform of the
    // same bytes — outlive them.  The pool is emptied at process exit.
    struct backend
    {
        b200sa_ctx * context = nullptr;
        b200sa_group * group = nullptr;
        std::vector<int> devices;
    };

    int selected_device()
    {
        char const * env = std::getenv("MSUFSORT_DEVICE");
        return env ? std::atoi(env) : 0;
    }

    // Which GPUs an msufsort(numThreads) object uses:
    //   MSUFSORT_DEVICES=a,b,c   exactly these (a device may be listed more than once: several shards on one GPU)
    //   MSUFSORT_NUM_GPUS=N      GPUs 0..N-1; "auto": as many as the caller asked threads for (the reference's one
    //                            parallelism knob, msufsort.h:50-53), at most all of them
    //   neither                  one GPU, MSUFSORT_DEVICE (default 0)
    std::vector<int> selected_devices(std::int32_t numThreads)
    {
        std::vector<int> devices;
        if (char const * list = std::getenv("MSUFSORT_DEVICES"))
        {
            for (char const * p = list; *p;)
            {
                char * end = nullptr;
                long v = std::strtol(p, &end, 10);
                if (end == p)
                    break;
                devices.push_back(static_cast<int>(v));
                p = (*end == ',') ? end + 1 : end;
            }
            if (devices.size() > 16)
                devices.resize(16);
            if (!devices.empty())
                return devices;
        }
        char const * env = std::getenv("MSUFSORT_NUM_GPUS");
        if (env && *env)
        {
            int available = b200sa_device_count();
            int wanted = (std::strcmp(env, "auto") == 0) ? numThreads : std::atoi(env);
            if (wanted > available)
                wanted = available;
            if (wanted > 16)
                wanted = 16;
            for (int g = 0; g < wanted; ++g)
                devices.push_back(g);
        }
        if (devices.empty())
            devices.push_back(selected_device());
        return devices;
    }

    [[noreturn]] void fail(char const * what, int code)
    {
        throw std::runtime_error(std::string("msufsort (b200): ") + what + " failed with status " + std::to_string(code) + ": " + b200sa_last_error());
    }

    class backend_pool
    {
    public:
        ~backend_pool()
        {
            for (auto * b : idle_)
                destroy(b);
        }

        backend * acquire(std::vector<int> const & devices)
        {
            {
                std::lock_guard<std::mutex> guard(mutex_);
                for (std::size_t i = 0; i < idle_.size(); ++i)
                    if (idle_[i]->devices == devices)
                    {
                        backend * b = idle_[i];
                        idle_.erase(idle_.begin() + static_cast<std::ptrdiff_t>(i));
                        return b;
                    }
            }
            backend * b = new backend;
            b->devices = devices;
            bool const sharded = devices.size() > 1;
            int status = sharded ? b200sa_group_create(&b->group, devices.data(), static_cast<int>(devices.size()))
                                 : b200sa_create(&b->context, devices[0]);
            if (status != B200SA_OK)
            {
                delete b;
                fail(sharded ? "b200sa_group_create" : "b200sa_create", status);
            }
            return b;
        }

        void release(backend * b)
        {
            {
                std::lock_guard<std::mutex> guard(mutex_);
                if (idle_.size() < max_idle)
                {
                    idle_.push_back(b);
                    return;
                }
            }
            destroy(b);
        }

    private:
        static void destroy(backend * b)
        {
            if (b->group)
                b200sa_group_destroy(b->group);
            if (b->context)
                b200sa_destroy(b->context);
            delete b;
        }

        static constexpr std::size_t max_idle = 2;
        std::mutex mutex_;
        std::vector<backend *> idle_;
    };

    backend_pool & pool()
    {
        static backend_pool instance;
        return instance;
    }
}


//==============================================================================
maniscalco::msufsort::msufsort
(
    std::int32_t numThreads // host threads are not used; see selected_devices for what the knob can select
):
    backend_(pool().acquire(selected_devices(numThreads)))
{
}


//==============================================================================
maniscalco::msufsort::~msufsort()
{
    pool().release(static_cast<backend *>(backend_));
}


//==============================================================================
auto maniscalco::msufsort::make_suffix_array
(
    std::uint8_t const * inputBegin,
    std::uint8_t const * inputEnd
) -> suffix_array
{
    backend * b = static_cast<backend *>(backend_);
    std::int64_t inputSize = inputEnd - inputBegin;
    suffix_array suffixArray;
    suffixArray.reserve(static_cast<std::size_t>(inputSize) + 1);
    advise_huge_pages(suffixArray.data(), suffixArray.capacity() * sizeof(suffix_index));
    suffixArray.resize(static_cast<std::size_t>(inputSize) + 1);
    int status = b->group ? b200sa_group_suffix_array(b->group, inputBegin, inputSize, suffixArray.data())
                          : b200sa_suffix_array(b->context, inputBegin, inputSize, suffixArray.data());
    if (status != B200SA_OK)
        fail("make_suffix_array", status);
    return suffixArray;
}


//==============================================================================
int32_t maniscalco::msufsort::forward_burrows_wheeler_transform
(
    std::uint8_t * inputBegin,
    std::uint8_t * inputEnd
)
{
    backend * b = static_cast<backend *>(backend_);
    std::int32_t sentinelIndex = 0;
    int status = b->group ? b200sa_group_bwt(b->group, inputBegin, inputEnd - inputBegin, &sentinelIndex)
                          : b200sa_bwt(b->context, inputBegin, inputEnd - inputBegin, &sentinelIndex);
    if (status != B200SA_OK)
        fail("forward_burrows_wheeler_transform", status);
    return sentinelIndex;
}


//==============================================================================
void maniscalco::msufsort::reverse_burrows_wheeler_transform
(
    std::uint8_t * inputBegin,
    std::uint8_t * inputEnd,
    std::int32_t sentinelIndex,
    std::int32_t numThreads
)
{
    // static in the reference too (it spawns fresh std::threads per call, msufsort.cpp:1858-1879): a backend is borrowed
    // for the call, so concurrent callers do not serialise on one context
    backend * b = pool().acquire(selected_devices(numThreads));
    int status = b->group ? b200sa_group_unbwt(b->group, inputBegin, inputEnd - inputBegin, sentinelIndex)
                          : b200sa_unbwt(b->context, inputBegin, inputEnd - inputBegin, sentinelIndex);
    std::string message = status != B200SA_OK ? b200sa_last_error() : "";
    pool().release(b);
    if (status != B200SA_OK)
        throw std::runtime_error("msufsort (b200): reverse_burrows_wheeler_transform failed with status " + std::to_string(status) + ": " + message);
}
