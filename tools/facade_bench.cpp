// facade_bench.cpp — end-to-end timing of the DROP-IN path exactly as a user of the reference calls it: the free templates
// of <library/msufsort.h> on pageable std::vector storage (the reference demo's call sites, src/executable/msufsort/
// main.cpp:440, :470, :477).  One step = make_suffix_array + forward_burrows_wheeler_transform of the text; the inverse
// transform of the result is timed separately.  Everything inside the timed region is what the caller pays: allocation
// and zero-fill of the returned vector, host<->device copies, the sort.  Prints one JSON object; bench.py embeds it as
// "e2e_facade".
//
//   facade_bench <text file> <steps> <warmup>
#include <library/msufsort.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char ** argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: facade_bench <text file> <steps> <warmup>\n"); return 2; }
    int const steps = std::atoi(argv[2]), warmup = std::atoi(argv[3]);
    std::ifstream in(argv[1], std::ios::binary | std::ios::ate);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::size_t const n = static_cast<std::size_t>(in.tellg());
    in.seekg(0);
    std::vector<std::uint8_t> text(n);
    in.read(reinterpret_cast<char *>(text.data()), static_cast<std::streamsize>(n));
    try
    {
        // what the return type alone costs: a value-initialised std::vector<int32_t>(n + 1) is allocated, zero-filled (first touch
        // of every page) and freed in every make_suffix_array call — in the reference as well (msufsort.cpp:1754-1758)
        double vector_alloc = 0;
        for (int it = 0; it < 2; ++it)
        {
            double const t0 = now_s();
            {
                std::vector<std::int32_t> v(n + 1);
                asm volatile("" : : "r"(v.data()) : "memory");
            }
            vector_alloc = now_s() - t0;
        }
        double forward = 0, inverse = 0, sa_only = 0;
        std::int32_t sentinel = 0;
        bool ok = true;
        for (int it = 0; it < warmup + steps; ++it)
        {
            std::vector<std::uint8_t> work = text;   // the in-place transform needs its own copy: outside the timed region
            double const t0 = now_s();
            auto suffixArray = maniscalco::make_suffix_array(text.begin(), text.end(), 16);
            double const t1 = now_s();
            sentinel = maniscalco::forward_burrows_wheeler_transform(work.begin(), work.end(), 16);
            double const t2 = now_s();
            maniscalco::reverse_burrows_wheeler_transform(work.begin(), work.end(), sentinel, 16);
            double const t3 = now_s();
            ok = ok && work == text && suffixArray.size() == n + 1 && static_cast<std::size_t>(suffixArray[0]) == n && suffixArray[sentinel] == 0;
            if (it >= warmup) { forward += t2 - t0; sa_only += t1 - t0; inverse += t3 - t2; }
        }
        std::printf("{\"n_bytes\": %zu, \"steps\": %d, \"warmup\": %d, \"sa_bwt_ms_per_step\": %.3f, \"sa_ms\": %.3f, \"bwt_ms\": %.3f, "
                    "\"unbwt_ms_per_step\": %.3f, \"sa_bwt_MBps\": %.1f, \"unbwt_MBps\": %.1f, \"roundtrip_ok\": %s, \"vector_alloc_zero_free_ms\": %.3f, "
                    "\"path\": \"maniscalco::make_suffix_array + forward_burrows_wheeler_transform (free templates, pageable std::vector), then reverse_burrows_wheeler_transform\"}\n",
                    n, steps, warmup, 1e3 * forward / steps, 1e3 * sa_only / steps, 1e3 * (forward - sa_only) / steps, 1e3 * inverse / steps,
                    n * steps / forward / 1e6, n * steps / inverse / 1e6, ok ? "true" : "false", 1e3 * vector_alloc);
        return ok ? 0 : 1;
    }
    catch (std::exception const & e)
    {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
}
