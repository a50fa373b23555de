// facade_test.cpp — a caller written against the REFERENCE's public header surface
// (#include <library/msufsort.h>, namespace maniscalco, iterator templates, class msufsort), in the
// shape of the reference demo's own call sites (src/executable/msufsort/main.cpp:403-477), built
// against this repository's drop-in header and libraries.  Prints "OK <fnv of SA> <sentinel>" lines
// that tests/test_facade_gpu.py compares with the oracle.
#include <library/msufsort.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

static std::uint64_t fnv1a64(void const * data, std::size_t bytes)
{
    auto p = static_cast<std::uint8_t const *>(data);
    std::uint64_t h = 0xcbf29ce484222325ull;
    for (std::size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

int main(int argc, char ** argv)
{
    if (argc < 2) { std::cerr << "usage: facade_test <input file>\n"; return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<std::int8_t> input((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());  // the demo uses int8_t (main.cpp:163)
    try
    {
        // free templates, as in main.cpp:440 / :470 / :477
        auto suffixArray = maniscalco::make_suffix_array(input.begin(), input.end(), 4);
        std::printf("SA %016llx %zu\n", (unsigned long long)fnv1a64(suffixArray.data(), suffixArray.size() * sizeof(std::int32_t)), suffixArray.size());
        auto copy = input;
        std::int32_t sentinelIndex = maniscalco::forward_burrows_wheeler_transform(copy.begin(), copy.end(), 4);
        std::printf("BWT %016llx %d\n", (unsigned long long)fnv1a64(copy.data(), copy.size()), sentinelIndex);
        maniscalco::reverse_burrows_wheeler_transform(copy.begin(), copy.end(), sentinelIndex, 4);
        std::printf("UNBWT %s\n", copy == input ? "roundtrip-ok" : "MISMATCH");
        // class interface, as the templates use it internally (msufsort.h:444, :461)
        maniscalco::msufsort sorter(2);
        auto again = sorter.make_suffix_array((std::uint8_t const *)input.data(), (std::uint8_t const *)input.data() + input.size());
        std::printf("CLASS %s\n", again == suffixArray ? "same" : "DIFFERENT");
    }
    catch (std::exception const & e)
    {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    return 0;
}
