// ext_group_test.cpp — caller of the multi-GPU part of the C++ extension header (maniscalco::b200::gpu_group): a batch of
// blocks spread over the listed devices must give exactly what one context gives.  tests/test_cpp_ext.py (emulator) and
// tests/test_zz_group_batch_gpu.py (GPU) build and run it.
//   ext_group_test <input file> <block size> <device,device,...>
#include <library/msufsort/msufsort_b200.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

int main(int argc, char ** argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: ext_group_test <input file> <block size> <device,device,...>\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<std::int8_t> input((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::size_t const blockSize = std::strtoull(argv[2], nullptr, 10);
    std::vector<int> devices;
    for (char * tok = std::strtok(argv[3], ","); tok; tok = std::strtok(nullptr, ",")) devices.push_back(std::atoi(tok));
    try
    {
        maniscalco::b200::packed_blocks blocks;
        for (std::size_t at = 0; at < input.size(); at += blockSize)
            blocks.push_back(input.begin() + at, input.begin() + std::min(input.size(), at + blockSize));
        blocks.push_back(input.begin(), input.begin());  // an empty block
        auto one = blocks, many = blocks;
        maniscalco::b200::context gpu;
        maniscalco::b200::gpu_group gpus(devices);
        std::printf("GROUP %d BLOCKS %lld\n", b200sa_group_size(gpus.get()), (long long)blocks.size());
        bool same = maniscalco::b200::make_suffix_arrays(gpu, one) == maniscalco::b200::make_suffix_arrays(gpus, many);
        std::printf("SA %s\n", same ? "same" : "DIFFERENT");
        auto s1 = maniscalco::b200::forward_burrows_wheeler_transform(gpu, one);
        auto s2 = maniscalco::b200::forward_burrows_wheeler_transform(gpus, many);
        std::printf("BWT %s\n", s1 == s2 && one.bytes == many.bytes && one.bytes != blocks.bytes ? "same" : "DIFFERENT");
        maniscalco::b200::reverse_burrows_wheeler_transform(gpus, many, s2);
        std::printf("ROUNDTRIP %s\n", many.bytes == blocks.bytes ? "ok" : "MISMATCH");
    }
    catch (std::exception const & e)
    {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    return 0;
}
