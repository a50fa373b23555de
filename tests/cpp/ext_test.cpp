// ext_test.cpp — caller of the C++ extension header (src/library/msufsort/msufsort_b200.h): LCP and batched transforms.
// Prints digests that tests/test_cpp_ext.py compares with the oracle.
#include <library/msufsort/msufsort_b200.h>

#include <cstdio>
#include <fstream>
#include <iterator>

static std::uint64_t fnv1a64(void const * data, std::size_t bytes)
{
    auto p = static_cast<std::uint8_t const *>(data);
    std::uint64_t h = 0xcbf29ce484222325ull;
    for (std::size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

int main(int argc, char ** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: ext_test <input file> <block size>\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<std::int8_t> input((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::size_t const blockSize = std::strtoull(argv[2], nullptr, 10);
    try
    {
        maniscalco::b200::context gpu;
        auto r = maniscalco::b200::make_lcp_array(gpu, input.begin(), input.end());
        std::printf("LCP %016llx %016llx\n", (unsigned long long)fnv1a64(r.suffixArray.data(), r.suffixArray.size() * 4),
                    (unsigned long long)fnv1a64(r.lcpArray.data(), r.lcpArray.size() * 4));
        maniscalco::b200::packed_blocks blocks;
        for (std::size_t at = 0; at < input.size(); at += blockSize)
            blocks.push_back(input.begin() + at, input.begin() + std::min(input.size(), at + blockSize));
        blocks.push_back(input.begin(), input.begin());  // an empty block
        auto original = blocks.bytes;
        auto suffixArrays = maniscalco::b200::make_suffix_arrays(gpu, blocks);
        auto sentinels = maniscalco::b200::forward_burrows_wheeler_transform(gpu, blocks);
        for (std::int64_t b = 0; b < blocks.size(); ++b)
        {
            auto range = blocks.block(b);
            std::size_t const n = range.second - range.first;
            std::printf("BLOCK %lld %zu %016llx %016llx %d\n", (long long)b, n,
                        (unsigned long long)fnv1a64(suffixArrays.data() + blocks.offsets[b] + b, (n + 1) * 4),
                        (unsigned long long)fnv1a64(range.first, n), sentinels[b]);
        }
        maniscalco::b200::reverse_burrows_wheeler_transform(gpu, blocks, sentinels);
        std::printf("ROUNDTRIP %s\n", blocks.bytes == original ? "ok" : "MISMATCH");
    }
    catch (std::exception const & e)
    {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    return 0;
}
