// comm_stress.cpp — TEST INFRASTRUCTURE.  Stress of the sharded runs' control plane (msufsort_b200/csrc/comm.cuh) between the
// threads of one process: thousands of barriers, all-gathers and sums in a row, a payload hand-off that only the barrier
// orders, and a rank that fails while its peers wait.  tests/test_sharded_cpu.py builds it twice: plain, and with
// -fsanitize=thread (data races, missing acquire / release pairs).
//   g++ -std=c++17 -O1 -g [-fsanitize=thread] -DB200SA_EMU -Itests/emu -Imsufsort_b200/csrc tests/cpp/comm_stress.cpp -pthread
#include "comm.cuh"

#include <cstdarg>
#include <cstdio>
#include <thread>
#include <vector>

namespace b200sa {
static thread_local char g_err[256];
int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace b200sa

using b200sa::Comm;

static int run_rank(Comm* c, int rounds, std::vector<long long>* board, int fail_round, int* failed_out)
{
    const int G = c->nranks, r = c->rank;
    for (int k = 0; k < rounds; ++k) {
        if (r == G - 1 && k == fail_round) { c->raise_error(); *failed_out = 1; return 0; }
        // plain (non-atomic) stores that the barrier alone publishes: what the round loop does with its pinned counters
        (*board)[(size_t)r] = (long long)k * 1000 + r;
        if (int rc = c->barrier()) return rc;
        long long want = 0, got = 0;
        for (int q = 0; q < G; ++q) { want += (long long)k * 1000 + q; got += (*board)[(size_t)q]; }
        if (want != got) { fprintf(stderr, "rank %d round %d: board sum %lld, expected %lld\n", r, k, got, want); return -1; }
        if (int rc = c->barrier()) return rc;  // nobody overwrites the board while a peer still reads it
        i64 s = 0;
        if (int rc = c->allreduce_sum((long long)(r + 1) * (k + 1), &s)) return rc;
        if (s != (i64)(k + 1) * G * (G + 1) / 2) { fprintf(stderr, "rank %d round %d: sum %lld\n", r, k, (long long)s); return -2; }
        unsigned char mine[24], all[24 * b200sa::kMaxPeers];
        for (int i = 0; i < 24; ++i) mine[i] = (unsigned char)(r * 31 + k + i);
        if (int rc = c->allgather(mine, sizeof(mine), all)) return rc;
        for (int q = 0; q < G; ++q)
            for (int i = 0; i < 24; ++i)
                if (all[q * 24 + i] != (unsigned char)(q * 31 + k + i)) { fprintf(stderr, "rank %d round %d: gather mismatch\n", r, k); return -3; }
    }
    return 0;
}

static int scenario(int G, int rounds, int fail_round)
{
    std::vector<Comm*> cs((size_t)G, nullptr);
    if (b200sa::comm_create_local(cs.data(), G) != 0) return 1;
    for (auto* c : cs) c->timeout_ms = 20000;
    std::vector<long long> board((size_t)G, 0);
    std::vector<int> rc((size_t)G, 0), failed((size_t)G, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < G; ++r) th.emplace_back([&, r] { rc[(size_t)r] = run_rank(cs[(size_t)r], rounds, &board, fail_round, &failed[(size_t)r]); });
    for (auto& t : th) t.join();
    int bad = 0;
    for (int r = 0; r < G; ++r) {
        if (fail_round < 0) bad |= rc[(size_t)r] != 0;
        else if (!failed[(size_t)r]) bad |= rc[(size_t)r] != B200SA_ECOMM;  // the peers of a failed rank leave with ECOMM, not a hang
    }
    for (auto* c : cs) b200sa::comm_destroy(c);
    return bad;
}

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    int bad = 0;
    for (int G : {2, 3, 8}) bad |= scenario(G, rounds, -1);
    bad |= scenario(4, rounds, rounds / 2);
    printf(bad ? "comm_stress: FAILED\n" : "comm_stress: ok\n");
    return bad;
}
