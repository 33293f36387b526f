// sim_main.cpp -- the host mirror's threading, exercised without a GPU (see fake_engine.cpp): K Convolution objects, one
// thread each like one JACK client each (main.cu:31-39), through the in-process JACK stand-in (headless_jack.cpp).
// Every scenario checks: no contract violation in the fake engine, every period's output is either the exact expected
// block or silence that skippedPeriods() accounts for, and the scenario's own expectations.  Built with
// -fsanitize=thread and -fsanitize=address by tests/hostsim/Makefile, run by tests/test_hostsim_cpu.py.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "convolution.h"
#include "headless_jack.h"
#include "settings_file.h"
#include "shared_engine.h"
#include "fake_engine.h"

namespace {
constexpr size_t B = 64;
int g_failures = 0;
std::atomic<int> g_paceUs{0};  // > 0: every member's cycles start at least this far apart (jackd's period clock)
#define CHECK(cond, ...) do { if (!(cond)) { g_failures++; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

float tapOf(int member, size_t ir) { return 10.f * (member + 1) + (float)ir; }

void prepareIR(Convolution &c, int member, size_t idx, size_t frames = 300)
{
    std::vector<float> l(frames, 0.f), r(frames, 0.f);
    l[0] = tapOf(member, idx);
    r[0] = -l[0];
    WavFile w(l.data(), r.data(), frames);
    c.prepare(idx, w, B);
}

struct Member {
    std::unique_ptr<Convolution> c;
    int k = 0;
    std::vector<float> in1, in2, L, R;
    uint64_t good = 0, silent = 0, wrong = 0;
    std::atomic<bool> ir2Ready{false};
    std::atomic<bool> pause{false}, quit{false};
    std::atomic<uint64_t> done{0};

    void open(int member)
    {
        k = member;
        c.reset(new Convolution("sim" + std::to_string(member), 1 << 15));
        in1.assign(B, 0); in2.assign(B, 0); L.assign(B, 0); R.assign(B, 0);
    }
    void start()
    {
        c->start();
        hj_port_set_buffer(c->capture[0], in1.data());
        hj_port_set_buffer(c->capture[1], in2.data());
        hj_port_set_buffer(c->playback[0], L.data());
        hj_port_set_buffer(c->playback[1], R.data());
    }
    // one JACK cycle with a known input; the expected output follows from the select values and the input of THIS period
    // (engine.shared_latency 1: of the previous call)
    int latency = 0;
    std::vector<float> pin1, pin2;
    size_t ps0 = 0, ps1 = 0;
    bool havePrev = false;
    void cycle(uint64_t p)
    {
        size_t s0 = (p / 40) % 2, s1 = (p / 64) % 2;
        if (ir2Ready.load(std::memory_order_acquire) && (p / 100) % 2) s0 = 2;
        c->cc[0].value.select = s0;
        c->cc[1].value.select = s1;
        for (size_t i = 0; i < B; i++) { in1[i] = 0.25f + 0.001f * (float)((p + i + k) % 97); in2[i] = -0.5f + 0.002f * (float)((p * 3 + i) % 89); }
        std::fill(L.begin(), L.end(), 777.f);  // sentinel: the callback must write every sample
        std::fill(R.begin(), R.end(), 777.f);
        const uint64_t skippedBefore = c->skippedPeriods();
        hj_cycle(c->handle, (jack_nframes_t)B);
        const bool skipped = c->skippedPeriods() != skippedBefore;
        const std::vector<float> &e1 = latency ? pin1 : in1, &e2 = latency ? pin2 : in2;
        const size_t es0 = latency ? ps0 : s0, es1 = latency ? ps1 : s1;
        bool isSilent = true, isExact = !latency || havePrev;
        for (size_t i = 0; i < B; i++) {
            isSilent = isSilent && L[i] == 0.f && R[i] == 0.f;
            isExact = isExact && L[i] == tapOf(k, es0) * e1[i] && R[i] == tapOf(k, es1) * e2[i];
        }
        if (isExact && !skipped) good++;
        else if (isSilent) silent++;
        else { wrong++; if (wrong < 4) fprintf(stderr, "member %d period %llu: L[0] = %g, expected %g or silence (skipped %d)\n", k, (unsigned long long)p, L[0], tapOf(k, es0) * (e1.empty() ? 0.f : e1[0]), (int)skipped); }
        if (latency) { pin1 = in1; pin2 = in2; ps0 = s0; ps1 = s1; havePrev = true; }
        done.fetch_add(1, std::memory_order_release);
    }
};

using Members = std::vector<std::unique_ptr<Member>>;

Members makeGroup(int K, uint32_t timeoutMs, bool prebuild, bool periodKnown = true, int latency = 0)
{
    EngineOptions o;
    o.shared = (uint32_t)K;
    o.sharedLatency = (uint32_t)latency;
    o.period = periodKnown ? (uint32_t)B : 0;  // 0: nobody knows the period before the first callback (plain harness)
    o.sharedTimeoutMs = timeoutMs;
    o.flags = CA_FLAG_STREAMING;
    Convolution::setDefaultOptions(o);
    Members m;
    for (int k = 0; k < K; k++) {
        m.emplace_back(new Member());
        m.back()->open(k);
        m.back()->latency = latency;
        prepareIR(*m.back()->c, k, 0);
        prepareIR(*m.back()->c, k, 1);
    }
    if (prebuild) CHECK(m[0]->c->buildNow(B), "buildNow failed");
    for (auto &x : m) x->start();
    return m;
}

void runThreads(Members &m, uint64_t periods, const std::function<void()> &control = nullptr)
{
    std::vector<std::thread> th;
    for (auto &x : m)
        th.emplace_back([&, mp = x.get()] {
            auto next = std::chrono::steady_clock::now();
            for (uint64_t p = 0; p < periods && !mp->quit.load(); p++) {
                while (mp->pause.load()) std::this_thread::sleep_for(std::chrono::milliseconds(1));
                if (int us = g_paceUs.load()) { std::this_thread::sleep_until(next); next = std::max(next, std::chrono::steady_clock::now() - std::chrono::milliseconds(1)) + std::chrono::microseconds(us); }
                mp->cycle(p);
            }
            mp->c->stop();  // like a JACK client that is closed: the others must not wait for it
        });
    std::thread ctl;
    if (control) ctl = std::thread(control);
    for (auto &t : th) t.join();
    if (ctl.joinable()) ctl.join();
}

uint64_t sum(const Members &m, uint64_t Member::*f) { uint64_t s = 0; for (auto &x : m) s += (*x).*f; return s; }

// ---- scenarios --------------------------------------------------------------------------------------------------
void steady(int K, uint64_t P)
{
    fprintf(stderr, "== steady: %d members, %llu periods\n", K, (unsigned long long)P);
    const uint64_t v0 = fake_violations(), c0 = fake_engines_created();
    Members m = makeGroup(K, 5000, true);
    runThreads(m, P);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sum(m, &Member::wrong) == 0, "%llu wrong periods", (unsigned long long)sum(m, &Member::wrong));
    // a member may lose the one period in which it joins a cycle that is already being decided
    // every member is expected from join() on: nobody runs ahead, nothing is lost
    for (auto &x : m) CHECK(x->silent == 0 && x->good == P, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
    CHECK(m[0]->c->sharedGroup()->batches() == P + 0, "%llu batches for %llu periods", (unsigned long long)m[0]->c->sharedGroup()->batches(), (unsigned long long)P);
    CHECK(fake_engines_created() - c0 == 1, "engine built %llu times, expected once", (unsigned long long)(fake_engines_created() - c0));
}

// jackd (and the reference arm's ref_bench, oracle/ref_harness/harness.cu): every member is called once per cycle and
// the next cycle starts when all of them have returned
void lockstepHost(int K, uint64_t P)
{
    fprintf(stderr, "== host with a cycle barrier: %d members, %llu cycles\n", K, (unsigned long long)P);
    const uint64_t v0 = fake_violations();
    Members m = makeGroup(K, 5000, true);
    std::atomic<int> waiting{0};
    std::atomic<uint64_t> cycle{0};
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (auto &x : m)
        th.emplace_back([&, mp = x.get()] {
            for (uint64_t p = 0; p < P; p++) {
                mp->cycle(p);
                if (waiting.fetch_add(1) + 1 == K) { waiting.store(0); cycle.fetch_add(1); }
                else while (cycle.load() <= p) std::this_thread::yield();
            }
            mp->c->stop();
        });
    for (auto &t : th) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sec < 3.0, "%.2f s: somebody waited for a timeout", sec);
    CHECK(m[0]->c->sharedGroup()->dropped() == 0, "a member was set aside");
    for (auto &x : m) CHECK(x->silent == 0 && x->good == P && x->wrong == 0, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
}

// A driver without a clock (an offline render: one loop per member, as fast as it goes) whose engine is only built at
// the first rendezvous: the members wait for that build instead of running through their input
void unpacedDriverFirstBuild(int K, uint64_t P)
{
    fprintf(stderr, "== unpaced driver, engine built at the first rendezvous: %d members\n", K);
    const uint64_t v0 = fake_violations();
    hj_set_buffer_size(0);
    fake_set_create_delay_us(30000);
    Members m = makeGroup(K, 5000, false, false);
    runThreads(m, P);
    fake_set_create_delay_us(0);
    CHECK(fake_violations() == v0, "contract violations");
    for (auto &x : m) CHECK(x->wrong == 0 && x->silent <= 2 && x->good + x->silent == P, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
    auto g = m[0]->c->sharedGroup();
    fprintf(stderr, "   rebuilds %llu batches %llu dropped %llu\n", (unsigned long long)g->rebuilds(), (unsigned long long)g->batches(), (unsigned long long)g->dropped());
}

// engine.shared_latency 1: hand in / take out one period later, nobody waits inside a cycle.  The host that needs it
// calls its clients one after the other on ONE thread (jack1); the same mode must also hold with one thread per
// member, paced or not, and with prepare() against the running batch.
void pipelinedMode(int K, uint64_t P)
{
    fprintf(stderr, "== engine.shared_latency 1: sequential host, threads, prepare(): %d members\n", K);
    const uint64_t v0 = fake_violations();
    {
        Members m = makeGroup(K, 5000, true, true, 1);
        const auto t0 = std::chrono::steady_clock::now();
        for (uint64_t p = 0; p < P; p++)
            for (auto &x : m) x->cycle(p);  // one thread, one client after the other
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (auto &x : m) x->c->stop();
        auto g = m[0]->c->sharedGroup();
        CHECK(sec < 3.0 && g->dropped() == 0, "sequential host: %.2f s, %llu members set aside", sec, (unsigned long long)g->dropped());
        CHECK(g->batches() == P, "sequential host: %llu batches for %llu cycles", (unsigned long long)g->batches(), (unsigned long long)P);
        for (auto &x : m) CHECK(x->wrong == 0 && x->silent == 1 && x->good == P - 1, "sequential host, member %d: good %llu silent %llu wrong %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent, (unsigned long long)x->wrong);
    }
    {
        Members m = makeGroup(K, 5000, true, true, 1);
        runThreads(m, P);  // unpaced threads: a member that runs ahead waits for the generation it already handed a block to
        for (auto &x : m) CHECK(x->wrong == 0 && x->silent <= 2 && x->good + x->silent == P, "threads, member %d: good %llu silent %llu wrong %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent, (unsigned long long)x->wrong);
    }
    {
        fake_set_create_delay_us(2000);
        g_paceUs.store(150);
        Members m = makeGroup(K, 5000, false, true, 1);
        runThreads(m, P, [&] {
            for (int round = 0; round < 9; round++) {
                std::this_thread::sleep_for(std::chrono::milliseconds(8));
                Member &x = *m[round % K];
                prepareIR(*x.c, x.k, round % 2 ? 2 : 1, 300 + 50 * round);
                if (round % 3 == 2) CHECK(x.c->buildNow(B), "buildNow on a live pipelined group failed");
                if (round % 2) x.ir2Ready.store(true, std::memory_order_release);
            }
        });
        fake_set_create_delay_us(0);
        g_paceUs.store(0);
        CHECK(sum(m, &Member::wrong) == 0, "prepare on a pipelined group: %llu wrong periods", (unsigned long long)sum(m, &Member::wrong));
        CHECK(sum(m, &Member::good) > sum(m, &Member::silent), "prepare on a pipelined group: more silence than sound");
        for (auto &x : m) CHECK(x->good + x->silent == P, "member %d lost periods", x->k);
    }
    CHECK(fake_violations() == v0, "contract violations");
}

void prepareOnLiveGroup(int K, uint64_t P)
{
    fprintf(stderr, "== prepare() and buildNow() against a running group: %d members, %llu periods\n", K, (unsigned long long)P);
    const uint64_t v0 = fake_violations(), c0 = fake_engines_created();
    fake_set_create_delay_us(2000);
    g_paceUs.store(150);  // paced like JACK cycles: 2000 periods outlast the control thread below
    fake_set_process_delay_us(30);  // a batch is in flight most of the time
    Members m = makeGroup(K, 5000, false);  // no buildNow: the first rendezvous builds
    runThreads(m, P, [&] {
        for (int round = 0; round < 12; round++) {
            std::this_thread::sleep_for(std::chrono::milliseconds(8));
            Member &x = *m[round % K];
            prepareIR(*x.c, x.k, round % 2 ? 2 : 1, 300 + 50 * round);  // a new slot / a longer IR for an old one: both rebuild
            // every third round races buildNow() against the rebuild the rendezvous is about to do on its own
            if (round % 3 == 2) CHECK(x.c->buildNow(B), "buildNow on a live group failed");
            if (round % 2) x.ir2Ready.store(true, std::memory_order_release);
        }
    });
    fake_set_create_delay_us(0);
    fake_set_process_delay_us(0);
    g_paceUs.store(0);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sum(m, &Member::wrong) == 0, "%llu wrong periods", (unsigned long long)sum(m, &Member::wrong));
    CHECK(sum(m, &Member::good) > sum(m, &Member::silent), "more silence than sound");
    CHECK(fake_engines_created() - c0 >= 2, "no rebuild happened");
    for (auto &x : m) CHECK(x->good + x->silent == P, "member %d lost periods", x->k);
    fprintf(stderr, "   good %llu silent %llu rebuilds %llu\n", (unsigned long long)sum(m, &Member::good), (unsigned long long)sum(m, &Member::silent), (unsigned long long)(fake_engines_created() - c0));
}

void memberStopsAndResumes(int K, uint64_t P)
{
    fprintf(stderr, "== a member stops calling, is set aside, and comes back: %d members\n", K);
    const uint64_t v0 = fake_violations();
    fake_set_process_delay_us(50);  // paces the cycles so the test has a duration
    Members m = makeGroup(K, 30, true);
    std::shared_ptr<SharedEngine> none;
    runThreads(m, P, [&] {
        while (m[1]->done.load() < 200) std::this_thread::yield();
        m[1]->pause.store(true);
        const uint64_t before = m[0]->done.load();
        std::this_thread::sleep_for(std::chrono::milliseconds(300));
        const uint64_t during = m[0]->done.load() - before;
        CHECK(during > 100, "the others stalled with member 1 (%llu periods in 300 ms)", (unsigned long long)during);
        m[1]->pause.store(false);
    });
    fake_set_process_delay_us(0);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sum(m, &Member::wrong) == 0, "wrong periods");
    for (auto &x : m) CHECK(x->good + x->silent == P && x->silent <= 4, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
}

void memberDestroyedMidRun(int K, uint64_t P)
{
    fprintf(stderr, "== a member is destroyed while the others run: %d members\n", K);
    const uint64_t v0 = fake_violations();
    fake_set_process_delay_us(20);
    Members m = makeGroup(K, 5000, true);
    m[2]->quit.store(false);
    std::vector<std::thread> th;
    for (auto &x : m)
        th.emplace_back([&, mp = x.get()] {
            const uint64_t mine = mp->k == 2 ? P / 4 : P;
            for (uint64_t p = 0; p < mine; p++) mp->cycle(p);
            if (mp->k == 2) mp->c.reset();  // no stop(): ~Convolution -> leave() alone must release the others (timeout is 5 s)
            else mp->c->stop();
        });
    const auto t0 = std::chrono::steady_clock::now();
    for (auto &t : th) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fake_set_process_delay_us(0);
    auto g = m[0]->c->sharedGroup();
    fprintf(stderr, "   %.2f s, dropped %llu, taking part %d, batches %llu\n", sec, (unsigned long long)g->dropped(), g->taking_part(), (unsigned long long)g->batches());
    CHECK(g->dropped() == 0, "a member was set aside by timeout instead of by leave()");
    CHECK(sec < 4.0, "the others waited for the destroyed member (%.1f s)", sec);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sum(m, &Member::wrong) == 0, "wrong periods");
    for (auto &x : m) if (x->k != 2) CHECK(x->good + x->silent == P && x->silent <= 2, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
}

void buildFailureThenRecovery(int K, uint64_t P)
{
    fprintf(stderr, "== engine creation fails twice, then works: %d members\n", K);
    const uint64_t v0 = fake_violations();
    fake_fail_next_creates(2);
    EngineOptions o;
    o.shared = (uint32_t)K; o.period = (uint32_t)B; o.sharedTimeoutMs = 5000; o.flags = CA_FLAG_STREAMING;
    Convolution::setDefaultOptions(o);
    Members m;
    for (int k = 0; k < K; k++) { m.emplace_back(new Member()); m.back()->open(k); prepareIR(*m.back()->c, k, 0); prepareIR(*m.back()->c, k, 1); }
    CHECK(!m[0]->c->buildNow(B), "buildNow should have failed");
    for (auto &x : m) x->start();
    g_paceUs.store(100);  // unpaced members would burn through their periods while the rebuild runs
    runThreads(m, P);
    g_paceUs.store(0);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(sum(m, &Member::wrong) == 0, "wrong periods");
    for (auto &x : m) CHECK(x->good >= P - 20 && x->good + x->silent == P, "member %d: good %llu silent %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent);
}

// engine.gpus 2 + engine.shared 3: six objects become two batches of three, one per GPU; engine.ir_split 2: one object
// whose engine is a group over two GPUs, with prepare() against its running callback
void multiGpuOptions(uint64_t P)
{
    fprintf(stderr, "== engine.gpus / engine.ir_split\n");
    const uint64_t v0 = fake_violations(), d0 = fake_engines_on_device(0), d1 = fake_engines_on_device(1);
    {
        EngineOptions o;
        o.shared = 3; o.gpus = 2; o.period = (uint32_t)B; o.sharedTimeoutMs = 5000; o.flags = CA_FLAG_STREAMING;
        Convolution::setDefaultOptions(o);
        Members m;
        for (int k = 0; k < 6; k++) { m.emplace_back(new Member()); m.back()->open(k); prepareIR(*m.back()->c, k, 0); prepareIR(*m.back()->c, k, 1); }
        for (auto &x : m) x->start();
        CHECK(m[0]->c->options().device == 0 && m[1]->c->options().device == 1 && m[4]->c->options().device == 0, "objects are not dealt round-robin onto the GPUs");
        CHECK(m[0]->c->sharedGroup() == m[2]->c->sharedGroup() && m[0]->c->sharedGroup() == m[4]->c->sharedGroup() && m[1]->c->sharedGroup() == m[5]->c->sharedGroup() &&
              m[0]->c->sharedGroup() != m[1]->c->sharedGroup(), "batches do not form per GPU");
        runThreads(m, P);
        for (auto &x : m) CHECK(x->wrong == 0 && x->silent == 0 && x->good == P, "member %d: good %llu silent %llu wrong %llu", x->k, (unsigned long long)x->good, (unsigned long long)x->silent, (unsigned long long)x->wrong);
        CHECK(fake_engines_on_device(0) - d0 == 1 && fake_engines_on_device(1) - d1 == 1, "expected one batched engine per GPU");
    }
    {
        const uint64_t g0 = fake_groups_created(), e0 = fake_engines_on_device(2), e1 = fake_engines_on_device(3);
        EngineOptions o;
        o.irSplit = 2; o.device = 2; o.period = (uint32_t)B;
        Convolution::setDefaultOptions(o);
        Members m;
        m.emplace_back(new Member());
        m[0]->open(0);
        prepareIR(*m[0]->c, 0, 0);
        prepareIR(*m[0]->c, 0, 1);
        m[0]->start();
        CHECK(m[0]->c->group() != nullptr && m[0]->c->engine() == nullptr, "engine.ir_split did not build a group");
        fake_set_process_delay_us(20);
        runThreads(m, P, [&] {
            for (int round = 0; round < 6; round++) {
                std::this_thread::sleep_for(std::chrono::milliseconds(5));
                prepareIR(*m[0]->c, 0, round % 2 ? 2 : 1, 300 + 40 * round);
                if (round % 2) m[0]->ir2Ready.store(true, std::memory_order_release);
            }
        });
        fake_set_process_delay_us(0);
        CHECK(m[0]->wrong == 0 && m[0]->good > P / 2 && m[0]->good + m[0]->silent == P, "ir_split: good %llu silent %llu wrong %llu", (unsigned long long)m[0]->good, (unsigned long long)m[0]->silent, (unsigned long long)m[0]->wrong);
        CHECK(fake_groups_created() > g0 && fake_engines_on_device(2) > e0 && fake_engines_on_device(3) > e1, "the group does not span devices 2 and 3");
    }
    CHECK(fake_violations() == v0, "contract violations");
}

// the engine.* keys of the settings file (host/convolution.h) and their environment twins reach EngineOptions
void optionsFromSettingsAndEnvironment()
{
    fprintf(stderr, "== engine.* settings keys / CA_ENGINE_* environment\n");
    Settings st;
    st.parse("# engine keys\nengine.device 2\nengine.tiers auto\nengine.tier_growth 4\nengine.tier_max_block 4096\nengine.period 128\n"
             "engine.shared 6\nengine.shared_timeout_ms 50\nengine.shared_latency 1\nengine.gpus 4\nengine.async_tiers true\n"
             "engine.l2_persist yes\nengine.ref_quirks true\nengine.graph true\n");
    EngineOptions o = EngineOptions::fromSettings(st);
    CHECK(o.device == 2 && o.autoTiers && o.tierGrowth == 4 && o.tierMaxBlock == 4096 && o.period == 128, "device / tiers / period keys");
    CHECK(o.shared == 6 && o.sharedTimeoutMs == 50 && o.sharedLatency == 1 && o.gpus == 4, "shared / gpus keys");
    CHECK((o.flags & CA_FLAG_ASYNC_TIERS) && (o.flags & CA_FLAG_L2_PERSIST) && (o.flags & CA_FLAG_REF_QUIRKS), "flag keys");
    CHECK(!(o.flags & CA_FLAG_GRAPH) && (o.flags & CA_FLAG_STREAMING), "a shared batch runs host-driven launches, whatever engine.graph says");
    Settings st2;
    st2.parse("engine.ir_split 2\nengine.exchange nccl\nengine.shared 8\nengine.graph false\n");
    o = EngineOptions::fromSettings(st2);
    CHECK(o.irSplit == 2 && o.exchange == CA_EXCHANGE_NCCL && o.shared == 0 && !(o.flags & CA_FLAG_GRAPH), "ir_split keys (one object is the whole group: no shared batch)");
    Settings none;
    none.parse("conv.count 2\n");
    o = EngineOptions::fromSettings(none);
    CHECK(o.device == 0 && !o.autoTiers && o.shared == 0 && o.gpus == 1 && o.irSplit == 0 && (o.flags & CA_FLAG_GRAPH), "defaults: one private uniform engine per object, one CUDA graph per period");
    setenv("CA_ENGINE_SHARED", "3", 1); setenv("CA_ENGINE_TIERS", "auto", 1); setenv("CA_ENGINE_GPUS", "2", 1); setenv("CA_ENGINE_SHARED_LATENCY", "1", 1);
    o = EngineOptions::fromEnv();
    CHECK(o.shared == 3 && o.autoTiers && o.gpus == 2 && o.sharedLatency == 1 && (o.flags & CA_FLAG_STREAMING), "environment twins");
    unsetenv("CA_ENGINE_SHARED"); unsetenv("CA_ENGINE_TIERS"); unsetenv("CA_ENGINE_GPUS"); unsetenv("CA_ENGINE_SHARED_LATENCY");
}

void singleObjectPrepareWhileRunning(uint64_t P)
{
    fprintf(stderr, "== one Convolution of its own: prepare() from a control thread against the running callback\n");
    const uint64_t v0 = fake_violations();
    EngineOptions o;  // not shared
    o.period = (uint32_t)B;
    Convolution::setDefaultOptions(o);
    hj_set_buffer_size((unsigned)B);
    Members m;
    m.emplace_back(new Member());
    m[0]->open(0);
    prepareIR(*m[0]->c, 0, 0);
    prepareIR(*m[0]->c, 0, 1);
    m[0]->start();
    fake_set_process_delay_us(20);
    runThreads(m, P, [&] {
        for (int round = 0; round < 10; round++) {
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
            prepareIR(*m[0]->c, 0, round % 2 ? 2 : 1, 300 + 40 * round);
            if (round % 2) m[0]->ir2Ready.store(true, std::memory_order_release);
        }
    });
    fake_set_process_delay_us(0);
    CHECK(fake_violations() == v0, "contract violations");
    CHECK(m[0]->wrong == 0 && m[0]->good > P / 2 && m[0]->good + m[0]->silent == P, "good %llu silent %llu wrong %llu", (unsigned long long)m[0]->good, (unsigned long long)m[0]->silent, (unsigned long long)m[0]->wrong);
}
}  // namespace

int main(int argc, char **argv)
{
    const uint64_t P = argc > 1 ? (uint64_t)atoll(argv[1]) : 3000;
    const int K = argc > 2 ? atoi(argv[2]) : 6;
    hj_set_sample_rate(48000);
    const std::string only = argc > 3 ? argv[3] : "";  // run one scenario only (debugging)
    auto want = [&](const char *name) { return only.empty() || only == name; };
    if (want("options")) optionsFromSettingsAndEnvironment();
    if (want("steady")) steady(K, P);
    if (want("lockstep")) lockstepHost(K, P);
    if (want("unpaced")) unpacedDriverFirstBuild(3, P);
    if (want("prepare")) prepareOnLiveGroup(K, P);
    if (want("stops")) memberStopsAndResumes(4, P);
    if (want("destroyed")) memberDestroyedMidRun(4, P);
    if (want("failure")) buildFailureThenRecovery(3, P / 2);
    if (want("single")) singleObjectPrepareWhileRunning(P);
    if (want("multigpu")) multiGpuOptions(P);
    if (want("pipelined")) pipelinedMode(K, P);
    fprintf(stderr, "engines created %llu destroyed %llu, batches %llu, IR loads %llu, violations %llu\n", (unsigned long long)fake_engines_created(),
            (unsigned long long)fake_engines_destroyed(), (unsigned long long)fake_periods_processed(), (unsigned long long)fake_ir_loads(), (unsigned long long)fake_violations());
    printf("HOSTSIM %s failures=%d violations=%llu\n", g_failures || fake_violations() ? "FAIL" : "OK", g_failures, (unsigned long long)fake_violations());
    return g_failures || fake_violations() ? 1 : 0;
}
