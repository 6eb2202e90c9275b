// CPU stress of agx::CopyCrew (agilex-ntt_b200/csrc/agx_copycrew.h): many copies of awkward sizes from two crews used by two
// submitting threads at once (the calling thread stages in, the drain helper copies out), every byte checked; run() hands out
// every part exactly once.  Built and run by tests/test_abi.py::test_copycrew_stress.
#include <atomic>
#include <cstdio>
#include <numeric>
#include "agx_copycrew.h"

static int check_copies(agx::CopyCrew &crew, uint64_t seed) {
    const size_t sizes[] = {0, 1, 4095, 1u << 20, (1u << 20) + 1, (3u << 20) + 7, (8u << 20) - 3, 33u << 20};
    std::vector<unsigned char> src(34u << 20), dst(34u << 20);
    for (size_t i = 0; i < src.size(); i++) src[i] = (unsigned char)((i * 2654435761u + seed) >> 13);
    for (int rep = 0; rep < 6; rep++)
        for (size_t n : sizes) {
            std::fill(dst.begin(), dst.end(), 0xEE);
            const size_t off = (rep * 4099u) % 8192;
            crew.copy(dst.data() + off, src.data() + 3 * off, n);
            if (memcmp(dst.data() + off, src.data() + 3 * off, n)) return 1;
            if (dst[off + n] != 0xEE || (off && dst[off - 1] != 0xEE)) return 2;      // nothing outside the range
        }
    for (size_t parts : {1u, 2u, 3u, 7u, 64u, 1000u}) {                               // run(): each part exactly once
        std::vector<std::atomic<int>> hit(parts);
        for (auto &h : hit) h = 0;
        crew.run(parts, [&](size_t i) { hit[i]++; });
        for (auto &h : hit) if (h != 1) return 3;
    }
    return 0;
}

int main() {
    for (int threads : {1, 2, 4, 6}) {
        agx::CopyCrew in(threads), out(threads);
        if (in.threads() != threads) { printf("threads() = %d, wanted %d\n", in.threads(), threads); return 1; }
        int ra = -1, rb = -1;
        std::thread t([&] { rb = check_copies(out, 99); });
        ra = check_copies(in, 7);
        t.join();
        if (ra || rb) { printf("threads=%d: FAILED (%d, %d)\n", threads, ra, rb); return 1; }
    }
    setenv("AGX_HOST_COPY_THREADS", "3", 1);
    if (agx::CopyCrew::default_threads() != 3) { printf("env knob ignored\n"); return 1; }
    unsetenv("AGX_HOST_COPY_THREADS");
    const int d = agx::CopyCrew::default_threads();
    if (d < 1 || d > 4) { printf("default_threads() = %d\n", d); return 1; }
    printf("copycrew ok\n");
    return 0;
}
