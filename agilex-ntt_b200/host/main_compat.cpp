// main_compat.cpp -- host driver with the call sequence of the reference's src/main.cpp (buffers -> host
// accessors -> ntt_input_kernel -> fwd_ntt_kernel<0> -> ntt_output_kernel -> q.wait() -> print), written against
// host/kernel/ntt.h.  Unlike main.cpp it can also run a REAL NTT instance and check it:
//
//   agx_main_compat                       main.cpp's dummy data (N=16384, modulus 65537, tables i+2 / i+3)
//   agx_main_compat --seal N [frames]     q = 1053818881, psi = minimal 2N-th root, ramp input; prints outputs
//   agx_main_compat ... --quiet           print only a checksum line
//
// Output format follows main.cpp:79-84: one decimal value per line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "kernel/ntt.h"
#include <dpc_common.hpp>

namespace {
typedef unsigned __int128 u128;
uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }
uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1;
    for (a %= q; e; e >>= 1, a = mulmod(a, a, q)) if (e & 1) r = mulmod(r, a, q);
    return r;
}
uint32_t bitrev(uint32_t x, int bits) { uint32_t r = 0; for (int i = 0; i < bits; i++, x >>= 1) r = (r << 1) | (x & 1); return r; }
uint64_t min_psi(uint64_t n, uint64_t q) {
    uint64_t g = 0;
    for (uint64_t x = 2; !g; x++) { uint64_t c = powmod(x, (q - 1) / (2 * n), q); if (powmod(c, n, q) == q - 1) g = c; }
    uint64_t g2 = mulmod(g, g, q), cur = g, best = g;
    for (uint64_t k = 0; k < n; k++) { if (cur < best) best = cur; cur = mulmod(cur, g2, q); }
    return best;
}
}  // namespace

int main(int argc, char** argv) {
    size_t N = 16384;
    unsigned numFrames = 1;
    bool seal = false, quiet = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--seal") && i + 1 < argc) {
            seal = true; N = strtoul(argv[++i], nullptr, 10);
            if (i + 1 < argc && argv[i + 1][0] != '-') numFrames = (unsigned)strtoul(argv[++i], nullptr, 10);
        } else if (!strcmp(argv[i], "--quiet")) quiet = true;
    }
    try {
        sycl::ext::intel::fpga_selector selector;
        sycl::queue q(selector, dpc_common::exception_handler);
        const size_t words = N * numFrames;
        buffer<uint64_t, 1> in(words), in2(words), modulus(1), tw(N), pre(N), out(words);
        {
            host_accessor a_in(in, write_only), a_in2(in2, write_only), a_mod(modulus, write_only);
            host_accessor a_tw(tw, write_only), a_pre(pre, write_only);
            if (!seal) {                       // the dummy fill of main.cpp:49-55
                for (size_t i = 0; i < N; i++) { a_in[i] = i; a_in2[i] = i + 1; a_tw[i] = i + 2; a_pre[i] = i + 3; }
                a_mod[0] = 65537;
            } else {
                const uint64_t qq = 1053818881ull;
                int logn = 0; while ((size_t(1) << logn) < N) logn++;
                const uint64_t psi = min_psi(N, qq);
                std::vector<uint64_t> pw(N); pw[0] = 1;
                for (size_t i = 1; i < N; i++) pw[i] = mulmod(pw[i - 1], psi, qq);
                for (size_t k = 0; k < N; k++) {
                    a_tw[k] = pw[bitrev((uint32_t)k, logn)];
                    a_pre[k] = (uint64_t)(((u128)a_tw[k] << 64) / qq);
                }
                for (size_t i = 0; i < words; i++) a_in[i] = a_in2[i] = i % qq;
                a_mod[0] = qq;
            }
        }
        ntt_input_kernel(in, in2, modulus, tw, pre, numFrames, q);
        fwd_ntt_kernel<0>(q);
        ntt_output_kernel(out, (int)numFrames, q);
        q.wait();
        host_accessor a_out(out, read_only);
        uint64_t sum = 0;
        for (size_t i = 0; i < words; i++) {
            sum = sum * 1099511628211ull + a_out[i];
            if (!quiet) std::cout << a_out[i] << "\n";
        }
        if (quiet) std::cout << "fnv-ish checksum " << sum << " over " << words << " words\n";
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
