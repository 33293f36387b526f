// paramq_main.cpp -- the C ABI's lock-free parameter hand-off (cuda-audio_b200/csrc/param_queue.h: bounded
// multi-producer / single-consumer ring behind ca_set_params / ca_set_glide) under ThreadSanitizer: W producer threads
// (MIDI threads, UI, the real-time thread) push numbered commands while one consumer (the processing thread) drains.
// Checks: nothing lost, nothing duplicated, every producer's commands arrive in its own order, payloads are intact,
// a full ring refuses instead of overwriting.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../cuda-audio_b200/csrc/param_queue.h"

int main(int argc, char **argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 6;
    const uint32_t N = argc > 2 ? (uint32_t)atoi(argv[2]) : 200000;
    int failures = 0;

    {  // a full ring refuses
        ParamQueue q;
        q.resize(1);  // rounds up to the minimum ring
        ParamCmd c;
        uint64_t pushed = 0;
        while (q.push(c)) pushed++;
        if (pushed < (1u << 13) || q.push(c)) { fprintf(stderr, "FAIL: full ring accepted a command (%llu)\n", (unsigned long long)pushed); failures++; }
        ParamCmd out;
        uint64_t popped = 0;
        while (q.pop(&out)) popped++;
        if (popped != pushed) { fprintf(stderr, "FAIL: popped %llu of %llu\n", (unsigned long long)popped, (unsigned long long)pushed); failures++; }
        if (!q.push(c)) { fprintf(stderr, "FAIL: drained ring refuses\n"); failures++; }
    }

    ParamQueue q;
    std::atomic<bool> go{false};
    std::atomic<uint64_t> refused{0};
    std::vector<std::thread> producers;
    for (int w = 0; w < W; w++)
        producers.emplace_back([&, w] {
            while (!go.load()) std::this_thread::yield();
            for (uint32_t n = 0; n < N; n++) {
                ParamCmd c;
                c.item = (uint32_t)w;
                c.kind = n & 1;
                c.p.select = n;
                c.p.predelay = n ^ 0x5a5a5a5au;
                c.p.wet = (float)w;
                c.glide = (float)n;
                while (!q.push(c)) { refused.fetch_add(1); std::this_thread::yield(); }  // ring full: the caller sees CA_ERR_STATE and retries
            }
        });
    std::vector<uint32_t> next((size_t)W, 0);
    uint64_t got = 0;
    go.store(true);
    while (got < (uint64_t)W * N) {
        ParamCmd c;
        if (!q.pop(&c)) { std::this_thread::yield(); continue; }
        got++;
        if (c.item >= (uint32_t)W) { fprintf(stderr, "FAIL: item %u\n", c.item); failures++; break; }
        const uint32_t n = next[c.item]++;
        if (c.p.select != n || c.p.predelay != (n ^ 0x5a5a5a5au) || c.kind != (n & 1) || c.p.wet != (float)c.item || c.glide != (float)n) {
            if (failures++ < 5) fprintf(stderr, "FAIL: producer %u: expected command %u, got select %u predelay %#x kind %u\n", c.item, n, c.p.select, c.p.predelay, c.kind);
        }
    }
    for (auto &t : producers) t.join();
    ParamCmd c;
    if (q.pop(&c)) { fprintf(stderr, "FAIL: extra command\n"); failures++; }
    fprintf(stderr, "%d producers x %u commands, ring refused %llu pushes\n", W, N, (unsigned long long)refused.load());
    printf("PARAMQ %s failures=%d\n", failures ? "FAIL" : "OK", failures);
    return failures ? 1 : 0;
}
