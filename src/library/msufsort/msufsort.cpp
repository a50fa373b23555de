// maniscalco::msufsort — marshalling layer from the reference's public C++ interface to the C ABI
// (include/b200sa.h).  No algorithmic code lives here; see msufsort_b200/csrc for the kernels.
#include "./msufsort.h"

#include <b200sa.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace
{
    int selected_device()
    {
        char const * env = std::getenv("MSUFSORT_DEVICE");
        return env ? std::atoi(env) : 0;
    }

    [[noreturn]] void fail(char const * what, int code)
    {
        throw std::runtime_error(std::string("msufsort (b200): ") + what + " failed with status " + std::to_string(code) + ": " + b200sa_last_error());
    }

    b200sa_ctx * create_context()
    {
        b200sa_ctx * context = nullptr;
        int status = b200sa_create(&context, selected_device());
        if (status != B200SA_OK)
            fail("b200sa_create", status);
        return context;
    }

    // the static reverse transform has no object to hang a context on: one lazily created,
    // mutex-guarded context per process serves it (the reference spawns fresh std::threads per
    // call there, msufsort.cpp:1858-1879).
    std::mutex sharedContextMutex;
    b200sa_ctx * sharedContext = nullptr;
}


//==============================================================================
maniscalco::msufsort::msufsort
(
    std::int32_t /* numThreads: host threads are not used; kept for source compatibility */
):
    context_(create_context())
{
}


//==============================================================================
maniscalco::msufsort::~msufsort()
{
    b200sa_destroy(context_);
}


//==============================================================================
auto maniscalco::msufsort::make_suffix_array
(
    std::uint8_t const * inputBegin,
    std::uint8_t const * inputEnd
) -> suffix_array
{
    std::int64_t inputSize = inputEnd - inputBegin;
    suffix_array suffixArray(static_cast<std::size_t>(inputSize) + 1);
    int status = b200sa_suffix_array(context_, inputBegin, inputSize, suffixArray.data());
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
    std::int32_t sentinelIndex = 0;
    int status = b200sa_bwt(context_, inputBegin, inputEnd - inputBegin, &sentinelIndex);
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
    std::int32_t /* numThreads */
)
{
    std::lock_guard<std::mutex> guard(sharedContextMutex);
    if (sharedContext == nullptr)
        sharedContext = create_context();
    int status = b200sa_unbwt(sharedContext, inputBegin, inputEnd - inputBegin, sentinelIndex);
    if (status != B200SA_OK)
        fail("reverse_burrows_wheeler_transform", status);
}
