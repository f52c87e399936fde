// Worker threads of the nte dispatcher: one per extra lane, bound to the lane's GPU.
//
// replaces: the futex-woken worker threads of CMathWorker (src/atop/threads.cpp:72-191,
// src/atop/threads.h:1079-1168: WorkMain wakes the workers once, works as "core 0" itself and waits once).
//
// A resident 128 MiB shard takes ~55 us to stream on a B200; a condition-variable hand-off costs 30-60 us per
// worker each way.  So a worker SPINS on its job word for a while after each job (kSpinNs: back-to-back NumPy
// calls find it awake, the hand-off is one cache-line transfer) and only then goes to sleep in a futex; the
// poster wakes it only if it is asleep.  The caller spins for completion -- a job is bounded by GPU work.
#pragma once
#include <linux/futex.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include <thread>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define NTE_PAUSE() _mm_pause()
#else
#define NTE_PAUSE() do { } while (0)
#endif

namespace nte {

inline uint64_t now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

struct Worker {
    typedef void (*Fn)(void* ctx, int lane);
    std::thread th;
    alignas(64) std::atomic<uint32_t> posted{0};    // job sequence number (futex word)
    std::atomic<uint32_t> asleep{0};
    alignas(64) std::atomic<uint32_t> finished{0};  // == posted once the job has run
    Fn fn = nullptr;
    void* ctx = nullptr;
    int lane = 0;
    bool quit = false;
    uint64_t spin_ns = 200000;

    static long futex(std::atomic<uint32_t>* addr, int op, uint32_t val) {
        return syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), op, val, nullptr, nullptr, 0);
    }

    void main() {
        uint32_t seen = 0;
        for (;;) {
            // spin first, then sleep
            const uint64_t t0 = now_ns();
            uint32_t cur;
            unsigned polls = 0;
            while ((cur = posted.load(std::memory_order_acquire)) == seen) {
                NTE_PAUSE();
                if ((++polls & 0xff) == 0 && now_ns() - t0 > spin_ns) {
                    asleep.store(1, std::memory_order_seq_cst);
                    while ((cur = posted.load(std::memory_order_seq_cst)) == seen) futex(&posted, FUTEX_WAIT_PRIVATE, seen);
                    asleep.store(0, std::memory_order_relaxed);
                    break;
                }
            }
            seen = cur;
            if (quit) return;
            fn(ctx, lane);
            finished.store(seen, std::memory_order_release);
        }
    }

    void start(int lane_index, uint64_t spin) {
        lane = lane_index;
        spin_ns = spin;
        th = std::thread([this] { main(); });
    }
    // called by ONE poster at a time (the dispatcher's lock)
    void post(Fn f, void* c) {
        fn = f;
        ctx = c;
        posted.fetch_add(1, std::memory_order_seq_cst);
        if (asleep.load(std::memory_order_seq_cst)) futex(&posted, FUTEX_WAKE_PRIVATE, 1);
    }
    void wait() {
        const uint32_t want = posted.load(std::memory_order_relaxed);
        unsigned polls = 0;
        while (finished.load(std::memory_order_acquire) != want) {
            NTE_PAUSE();
            if ((++polls & 0xffff) == 0) std::this_thread::yield();
        }
    }
    void stop() {
        quit = true;
        posted.fetch_add(1, std::memory_order_seq_cst);
        futex(&posted, FUTEX_WAKE_PRIVATE, 1);
        if (th.joinable()) th.join();
    }
};

}  // namespace nte
