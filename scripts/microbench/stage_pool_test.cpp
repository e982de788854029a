#include "../../boxmot_b200/csrc/host_stage.h"
#include <chrono>
#include <cstdio>
#include <vector>
#include <random>
int main() {
    bmb::StagePool pool(3);
    std::mt19937 rng(1);
    const size_t N = 2764800;
    std::vector<std::vector<unsigned char>> src(32, std::vector<unsigned char>(N + 100));
    for (auto& v : src) for (auto& b : v) b = (unsigned char)rng();
    std::vector<unsigned char> dst(N + 100);
    // correctness over odd sizes / offsets
    for (int it = 0; it < 2000; ++it) {
        size_t n = (it % 7 == 0) ? rng() % 1000 : 200000 + rng() % (N - 200000);
        size_t so = rng() % 64, d0 = rng() % 64;
        std::fill(dst.begin(), dst.end(), 0);
        size_t covered = 0, last_end = 0; bool ordered = true;
        pool.copy(dst.data() + d0, src[it % 32].data() + so, n, [&](size_t off, size_t len) { ordered &= off == last_end; last_end = off + len; covered += len; });
        if (covered != n || !ordered || memcmp(dst.data() + d0, src[it % 32].data() + so, n)) { printf("FAIL it %d n %zu\n", it, n); return 1; }
    }
    // exception path
    try { pool.copy(dst.data(), src[0].data(), N, [&](size_t off, size_t) { if (off) throw 5; }); printf("no throw\n"); return 1; } catch (int) {}
    pool.copy(dst.data(), src[1].data(), N);
    if (memcmp(dst.data(), src[1].data(), N)) { printf("FAIL after exception\n"); return 1; }
    for (int w : {0, 1, 3}) {
        bmb::StagePool p2(w);
        auto t0 = std::chrono::steady_clock::now();
        const int R = 320;
        for (int i = 0; i < R; ++i) { p2.copy(dst.data(), src[i % 32].data(), N); std::this_thread::sleep_for(std::chrono::microseconds(w == 1 && 0 ? 0 : 0)); }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / R;
        printf("workers %d: %.3f ms per frame copy (%.1f GB/s)\n", w, ms, N / ms / 1e6);
    }
    // with idle gaps (workers go back to sleep between frames, as in a per-frame loop)
    for (int w : {0, 3}) {
        bmb::StagePool p2(w);
        double tot = 0;
        for (int i = 0; i < 100; ++i) {
            std::this_thread::sleep_for(std::chrono::microseconds(1500));
            auto t0 = std::chrono::steady_clock::now();
            p2.copy(dst.data(), src[i % 32].data(), N);
            tot += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        printf("workers %d with 1.5 ms gaps: %.3f ms per frame copy\n", w, tot / 100);
    }
    printf("OK\n");
}
