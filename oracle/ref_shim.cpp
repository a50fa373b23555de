// ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" wrapper around the UNMODIFIED reference library so that tests and bench.py's
// cpu_baseline / --impl reference legs can call it through ctypes.  It is compiled together
// with /root/reference/src/library/msufsort/msufsort.cpp where that file lies (see Makefile);
// no reference source is copied into this repository.  Output: oracle/_ref/libmsufsort_ref.so
// (git-ignored, travels to the GPU box with the snapshot).
#include <library/msufsort.h>   // resolved with -I/root/reference/src
#include <cstdint>
#include <cstring>
#include <thread>

// The reference's LCP construction lives in the demo executable's anonymous namespace
// (src/executable/msufsort/main.cpp:16-105), not in the library: the unmodified file is compiled
// into this translation unit where it lies, with its main() renamed out of the way.
#define main msufsort_reference_demo_main
#include <executable/msufsort/main.cpp>
#undef main

extern "C" {

// maniscalco::make_suffix_array template (msufsort.h:432-445); sa_out has n+1 entries.
int ref_make_suffix_array(const uint8_t* text, int64_t n, int32_t* sa_out, int32_t threads)
{
    if (n <= 0) return -1;  // n == 0 is undefined behaviour in the reference (msufsort.cpp:1588)
    auto sa = maniscalco::make_suffix_array(text, text + n, threads);
    std::memcpy(sa_out, sa.data(), sizeof(int32_t) * (size_t)(n + 1));
    return 0;
}

// maniscalco::forward_burrows_wheeler_transform template (msufsort.h:449-462); in place.
int32_t ref_forward_bwt(uint8_t* text_inout, int64_t n, int32_t threads)
{
    if (n <= 0) return -1;
    return maniscalco::forward_burrows_wheeler_transform(text_inout, text_inout + n, threads);
}

// maniscalco::reverse_burrows_wheeler_transform template (msufsort.h:466-476); in place.
// The reference does not clamp numThreads here (it sizes a VLA with it, msufsort.cpp:1844).
int ref_reverse_bwt(uint8_t* bwt_inout, int64_t n, int32_t sentinel_index, int32_t threads)
{
    if (n <= 0 || threads <= 0) return -1;
    maniscalco::reverse_burrows_wheeler_transform(bwt_inout, bwt_inout + n, sentinel_index, threads);
    return 0;
}

// lcp_multithreaded (main.cpp:68-105) over SA[1..n] exactly as make_lcp_array (main.cpp:141-159) sets it up,
// but with size n-1 instead of n: the demo's last entry compares against one element past its vector.
// out[i] = lcp(SA[i+1], SA[i+2]) for i = 0..n-2.  Needs n >= 2.
int ref_lcp(const uint8_t* text, int64_t n, const int32_t* sa, int32_t* out, int32_t threads)
{
    if (n < 2 || threads < 1) return -1;
    std::vector<int32_t> buf(sa + 1, sa + n + 1);
    if ((int64_t)threads > n - 1) threads = 1;
    lcp_multithreaded((const int8_t*)text, (const int8_t*)text + n, buf.data(), (int32_t)(n - 1), threads);
    std::memcpy(out, buf.data(), sizeof(int32_t) * (size_t)(n - 1));
    return 0;
}

int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
