// agx_tables.h -- host-side parameter validation and twiddle-table generation for libagxntt.
//
// Replaces the "step before the path" of the reference: main.cpp:46-55 fills the twiddle / precon / modulus
// buffers on the host and ntt_input_kernel broadcasts them (ntt.cpp:544-571).  Here the library derives them
// from (n, q): psi = minimal primitive 2n-th root, roots[k] = psi^bitrev(k) in the order
// ntt.cpp:298-300 consumes, Shoup companions floor(w * 2^32 / q), inverse tables, n^-1.
// Independent of oracle/ (the tests cross-check the two).
#pragma once
#include <cstdint>
#include <vector>

namespace agx {

uint64_t mulmod_u64(uint64_t a, uint64_t b, uint64_t q);
uint64_t powmod_u64(uint64_t a, uint64_t e, uint64_t q);
bool is_prime_u64(uint64_t q);
uint32_t bit_reverse(uint32_t x, uint32_t bits);

// 0 when q is not a usable NTT prime for n (not prime, >= 2^30, q != 1 mod 2n)
uint32_t minimal_psi(uint32_t n, uint32_t q);

struct NaturalTables {            // entry k = (w, floor(w * 2^32 / q)), w = base^bitrev(k)
    std::vector<uint32_t> w, wp;
};

NaturalTables natural_tables(uint32_t n, uint32_t q, uint32_t psi, bool inverse);

uint32_t shoup_companion(uint32_t w, uint32_t q);

}  // namespace agx
