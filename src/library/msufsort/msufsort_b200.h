#pragma once
// maniscalco::b200 — C++ conveniences for what the B200 engine offers BEYOND the reference's three calls
// (the reference-shaped surface itself is ./msufsort.h and is kept identical to the reference's header).
//
//   make_lcp_array(begin, end)                       suffix array + LCP array (the reference demo's "l" mode, main.cpp:455-464)
//   forward_burrows_wheeler_transform(blocks)        a whole batch of blocks in ONE launch sequence
//   reverse_burrows_wheeler_transform(blocks, idx)   (what a block-sorting compressor loops over, main.cpp:466-487)
//   make_suffix_arrays(blocks)
//   the same three with a gpu_group instead of a context: the batch spread over several GPUs
//
// Header-only over the C ABI (include/b200sa.h); link libb200sa.  Failures throw std::runtime_error with the
// b200sa_last_error() text.  The GPU is chosen with MSUFSORT_DEVICE (default 0), as in msufsort.cpp.

#include <b200sa.h>

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace maniscalco
{
    namespace b200
    {
        // one GPU context; create once and reuse (it owns the device workspace)
        class context
        {
        public:
            context()
            {
                char const * env = std::getenv("MSUFSORT_DEVICE");
                check(b200sa_create(&context_, env ? std::atoi(env) : 0), "b200sa_create");
            }
            ~context() { b200sa_destroy(context_); }
            context(context const &) = delete;
            context & operator = (context const &) = delete;
            b200sa_ctx * get() const { return context_; }

            static void check(int status, char const * what)
            {
                if (status != B200SA_OK)
                    throw std::runtime_error(std::string("msufsort (b200): ") + what + " failed with status " + std::to_string(status) + ": " + b200sa_last_error());
            }

        private:
            b200sa_ctx * context_ = nullptr;
        };

        struct lcp_result
        {
            std::vector<std::int32_t> suffixArray;   // n + 1 entries, [0] = n
            std::vector<std::int32_t> lcpArray;      // n + 1 entries aligned with the suffix array, [0] = [1] = 0
        };

        // contiguous 1-byte iterators, as everywhere in the reference's interface
        template <typename input_iter>
        lcp_result make_lcp_array(context & gpu, input_iter begin, input_iter end)
        {
            std::int64_t const n = end - begin;
            lcp_result result;
            result.suffixArray.resize(static_cast<std::size_t>(n) + 1);
            result.lcpArray.resize(static_cast<std::size_t>(n) + 1);
            context::check(b200sa_lcp(gpu.get(), n ? (std::uint8_t const *)&*begin : nullptr, n, nullptr, result.suffixArray.data(),
                                      result.lcpArray.data()), "make_lcp_array");
            return result;
        }

        // a batch of blocks packed back to back: block b is bytes [offsets[b], offsets[b + 1])
        struct packed_blocks
        {
            std::vector<std::uint8_t> bytes;
            std::vector<std::int64_t> offsets{0};

            template <typename input_iter>
            void push_back(input_iter begin, input_iter end)
            {
                bytes.insert(bytes.end(), (std::uint8_t const *)&*begin, (std::uint8_t const *)&*begin + (end - begin));
                offsets.push_back(static_cast<std::int64_t>(bytes.size()));
            }
            std::int64_t size() const { return static_cast<std::int64_t>(offsets.size()) - 1; }
            std::pair<std::uint8_t const *, std::uint8_t const *> block(std::int64_t b) const
            {
                return {bytes.data() + offsets[b], bytes.data() + offsets[b + 1]};
            }
        };

        // in place over blocks.bytes; returns one sentinel index per block (0 for an empty block)
        inline std::vector<std::int32_t> forward_burrows_wheeler_transform(context & gpu, packed_blocks & blocks)
        {
            std::vector<std::int32_t> sentinelIndices(static_cast<std::size_t>(blocks.size()));
            context::check(b200sa_bwt_batch(gpu.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), sentinelIndices.data()),
                           "forward_burrows_wheeler_transform (batch)");
            return sentinelIndices;
        }

        inline void reverse_burrows_wheeler_transform(context & gpu, packed_blocks & blocks, std::vector<std::int32_t> const & sentinelIndices)
        {
            if (static_cast<std::int64_t>(sentinelIndices.size()) != blocks.size())
                throw std::invalid_argument("one sentinel index per block");
            context::check(b200sa_unbwt_batch(gpu.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), sentinelIndices.data()),
                           "reverse_burrows_wheeler_transform (batch)");
        }

        // block b's suffix array (n_b + 1 entries, block-local values) starts at offsets[b] + b of the result
        inline std::vector<std::int32_t> make_suffix_arrays(context & gpu, packed_blocks const & blocks)
        {
            std::vector<std::int32_t> suffixArrays(blocks.bytes.size() + static_cast<std::size_t>(blocks.size()));
            context::check(b200sa_suffix_array_batch(gpu.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), suffixArrays.data()),
                           "make_suffix_arrays (batch)");
            return suffixArrays;
        }

        // Several GPUs behind the same batch calls: the blocks are cut into one contiguous run per GPU (b200sa_group_*_batch).
        // devices: CUDA device numbers, e.g. {0, 1, 2, 3}; default: MSUFSORT_NUM_GPUS (or every GPU present)
        class gpu_group
        {
        public:
            explicit gpu_group(std::vector<int> devices = {})
            {
                if (devices.empty())
                {
                    char const * env = std::getenv("MSUFSORT_NUM_GPUS");
                    int count = env ? std::atoi(env) : b200sa_device_count();
                    if (count < 1)
                        count = 1;
                    for (int d = 0; d < count; ++d)
                        devices.push_back(d);
                }
                context::check(b200sa_group_create(&group_, devices.data(), static_cast<int>(devices.size())), "b200sa_group_create");
            }
            ~gpu_group() { b200sa_group_destroy(group_); }
            gpu_group(gpu_group const &) = delete;
            gpu_group & operator = (gpu_group const &) = delete;
            b200sa_group * get() const { return group_; }

        private:
            b200sa_group * group_ = nullptr;
        };

        inline std::vector<std::int32_t> forward_burrows_wheeler_transform(gpu_group & gpus, packed_blocks & blocks)
        {
            std::vector<std::int32_t> sentinelIndices(static_cast<std::size_t>(blocks.size()));
            context::check(b200sa_group_bwt_batch(gpus.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), sentinelIndices.data()),
                           "forward_burrows_wheeler_transform (batch, several GPUs)");
            return sentinelIndices;
        }

        inline void reverse_burrows_wheeler_transform(gpu_group & gpus, packed_blocks & blocks, std::vector<std::int32_t> const & sentinelIndices)
        {
            if (static_cast<std::int64_t>(sentinelIndices.size()) != blocks.size())
                throw std::invalid_argument("one sentinel index per block");
            context::check(b200sa_group_unbwt_batch(gpus.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), sentinelIndices.data()),
                           "reverse_burrows_wheeler_transform (batch, several GPUs)");
        }

        inline std::vector<std::int32_t> make_suffix_arrays(gpu_group & gpus, packed_blocks const & blocks)
        {
            std::vector<std::int32_t> suffixArrays(blocks.bytes.size() + static_cast<std::size_t>(blocks.size()));
            context::check(b200sa_group_suffix_array_batch(gpus.get(), blocks.bytes.data(), blocks.offsets.data(), blocks.size(), suffixArrays.data()),
                           "make_suffix_arrays (batch, several GPUs)");
            return suffixArrays;
        }
    } // namespace b200
} // namespace maniscalco
