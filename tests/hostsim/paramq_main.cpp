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

    {  // ParamMirror (what ca_get_params answers): several writers of the SAME item, readers never see a torn block
        ParamMirror m;
        m.resize(4);
        auto block = [](uint32_t n) { ca_params p; p.select = n; p.predelay = n + 1; p.speed = n + 2; p.vsteps = -1; p.dry = (float)(n & 0xffff); p.wet = p.dry + 1.f; p.panDry = p.dry + 2.f; p.panWet = p.dry + 3.f; p.level = p.dry + 4.f; return p; };
        m.set(2, block(0));
        std::atomic<bool> stop{false};
        std::atomic<uint64_t> torn{0}, reads{0};
        std::vector<std::thread> writers, readers;
        for (int w = 0; w < 3; w++) writers.emplace_back([&, w] { for (uint32_t n = 0; n < N; n++) m.set(2, block(n * 3 + (uint32_t)w)); });
        for (int r = 0; r < 2; r++)
            readers.emplace_back([&] {
                while (!stop.load()) {
                    ca_params p;
                    m.get(2, &p);
                    const float d = (float)(p.select & 0xffff);
                    if (p.predelay != p.select + 1 || p.speed != p.select + 2 || p.dry != d || p.wet != d + 1.f || p.panDry != d + 2.f || p.panWet != d + 3.f || p.level != d + 4.f) torn.fetch_add(1);
                    reads.fetch_add(1);
                }
            });
        for (auto &t : writers) t.join();
        stop.store(true);
        for (auto &t : readers) t.join();
        ca_params other;
        m.get(1, &other);
        if (torn.load() || other.select != 0 || other.level != 0.f) { fprintf(stderr, "FAIL: %llu torn reads of %llu\n", (unsigned long long)torn.load(), (unsigned long long)reads.load()); failures++; }
        fprintf(stderr, "mirror: 3 writers x %u sets of one item, %llu reads, %llu torn\n", N, (unsigned long long)reads.load(), (unsigned long long)torn.load());
    }
    printf("PARAMQ %s failures=%d\n", failures ? "FAIL" : "OK", failures);
    return failures ? 1 : 0;
}
