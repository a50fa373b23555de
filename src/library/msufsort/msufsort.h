#pragma once
// maniscalco::msufsort — reference-shaped C++ facade over the B200 engine.
//
// Public surface identical to the reference's header (/root/reference/src/library/msufsort/
// msufsort.h:42-75 class, :403-426 free templates): same namespace, type aliases, member and
// template signatures, ownership (SA returned by value, BWT / inverse BWT in place) — so a caller
// recompiles against this header and links libmsufsort + libb200sa instead of the reference's
// static library.  Everything private in the reference (the multikey-quicksort / induction engine
// and its spinning worker pool, msufsort.h:77-398) is gone: the class holds one opaque GPU
// context and every method is one call into the C ABI of include/b200sa.h.
//
// Differences a caller can observe:
//   * numThreads never changes a result (it did not in the reference either).  No host threads are spawned
//     for it.  What the reference's one parallelism knob can select here is the number of GPUs: with the environment
//     variable MSUFSORT_NUM_GPUS=auto a msufsort(numThreads) object shards every text over min(numThreads, GPUs
//     present) GPUs; MSUFSORT_NUM_GPUS=N fixes the count; MSUFSORT_DEVICES=a,b,c names them; unset = one GPU.
//   * make_suffix_array followed by forward_burrows_wheeler_transform of the same bytes (what the reference's demo and
//     its users do) costs ONE suffix sort: the suffix array stays resident on the GPU and the second call recognises the
//     text after uploading it.  This also holds for the free templates below (their short-lived objects borrow their GPU
//     context from a process-wide pool).
//   * a buffer that is not the Burrows-Wheeler transform of any text (corrupted data) makes
//     reverse_burrows_wheeler_transform throw and leaves the buffer unchanged; the reference returns garbage.
//   * failures (no CUDA device, out of device memory, bad sentinel index) throw
//     std::runtime_error with the b200sa_last_error() text; the reference had no error path.
//   * n == 0 is defined (SA = {0}, BWT returns 0); inputs up to 2^31-2 bytes are handled, the
//     reference silently corrupts above 2^30-2 (its int32 flag bits, msufsort.h:84-93).
//   * the GPU of the single-GPU mode is chosen with the MSUFSORT_DEVICE environment variable (default 0).

#include <cstdint>
#include <stdint.h>
#include <vector>

namespace maniscalco
{

    class msufsort
    {
    public:

        using suffix_index = std::int32_t;
        using suffix_array = std::vector<suffix_index>;

        msufsort
        (
            std::int32_t numThreads = 1
        );

        ~msufsort();

        msufsort(msufsort const &) = delete;
        msufsort & operator = (msufsort const &) = delete;

        suffix_array make_suffix_array
        (
            std::uint8_t const * inputBegin,
            std::uint8_t const * inputEnd
        );

        int32_t forward_burrows_wheeler_transform
        (
            std::uint8_t * inputBegin,
            std::uint8_t * inputEnd
        );

        static void reverse_burrows_wheeler_transform
        (
            std::uint8_t * inputBegin,
            std::uint8_t * inputEnd,
            std::int32_t sentinelIndex,
            std::int32_t numThreads
        );

    private:

        void * backend_;   // one GPU context or a group of them, borrowed from the process-wide pool (msufsort.cpp)

    }; // class msufsort


    template <typename input_iter>
    msufsort::suffix_array make_suffix_array
    (
        input_iter begin,
        input_iter end,
        int32_t numThreads = 1
    )
    {
        // contiguous 1-byte iterators, as in the reference ((uint8_t const *)&*begin, msufsort.h:444)
        return msufsort(numThreads).make_suffix_array((std::uint8_t const *)&*begin, (std::uint8_t const *)&*begin + (end - begin));
    }


    template <typename input_iter>
    int32_t forward_burrows_wheeler_transform
    (
        input_iter begin,
        input_iter end,
        int32_t numThreads = 1
    )
    {
        return msufsort(numThreads).forward_burrows_wheeler_transform((std::uint8_t *)&*begin, (std::uint8_t *)&*begin + (end - begin));
    }


    template <typename input_iter>
    static void reverse_burrows_wheeler_transform
    (
        input_iter begin,
        input_iter end,
        int32_t sentinelIndex,
        int32_t numThreads = 1
    )
    {
        msufsort::reverse_burrows_wheeler_transform((std::uint8_t *)&*begin, (std::uint8_t *)&*begin + (end - begin), sentinelIndex, numThreads);
    }

} // namespace maniscalco
