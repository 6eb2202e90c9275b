// agx_diag.cuh -- the measured integer roofline: the butterfly's instruction stream with nothing else around it.
//
// bench.py's `roofline` block needs the rate at which THIS GPU retires the butterfly arithmetic when no memory
// operation is in the way ("the 64/clk figure must be microbenchmarked on the box first").  One CTA per
// SM, every thread runs a long unrolled stream of butterflies on 8 independent register chains:
//   kind 0: the u32 Harvey/Shoup butterfly of the batched kernels (agx_arith.cuh ct_bfly; ntt.cpp:331-369 at u32)
//   kind 1: the u64 butterfly of the reference-shaped path (ref_bfly_u64; ntt.cpp:331-369 verbatim widths)
//   kind 2, 3: kind 0 with one / two extra non-multiply instructions per butterfly (an independent LOP3 chain), i.e. the
//           issue-slot load of a real kernel, whose loads, stores, transposes and final reduction add 0.2-0.4 instructions
//           per butterfly-instruction: how far the multiply pipe can still be filled when the issue port is shared
//   kind 4: kind 2 whose extra instruction reads ONE register (immediate operands) instead of three
//   kind 5: kind 0 with the twiddle taken from the kernel-parameter constant bank (two register reads fewer per butterfly)
//   kind 6: kind 0 with a different twiddle register pair per chain (no operand reuse between neighbouring butterflies)
// Cycles are clock64() deltas of the SM (median over CTAs), so the result -- butterflies per clock per SM -- does not
// depend on the clock the GPU happens to run at; the implied clock (cycles / event time) is returned beside it.
#pragma once
#include "agx_ntt_kernels.cuh"

namespace agx {

constexpr int kDiagChains = 8, kDiagUnroll = 16, kDiagIters = 256;

template <int KIND>
__global__ void __launch_bounds__(1024, 1) diag_bfly_kernel(uint32_t *out, long long *cycles, uint32_t seed, LimbConst lc, uint64_t q64,
                                                            uint2 wc) {
    uint32_t a[kDiagChains], b[kDiagChains], e[kDiagChains];
    uint64_t a64[kDiagChains], b64[kDiagChains];
#pragma unroll
    for (int i = 0; i < kDiagChains; i++) {
        a[i] = seed + threadIdx.x * 977u + i * 131u;
        b[i] = (seed ^ 0x9e3779b9u) + i * 7919u + threadIdx.x;
        e[i] = seed * (i + 3u) + threadIdx.x;
        a64[i] = ((uint64_t)a[i] << 20) ^ b[i];
        b64[i] = ((uint64_t)b[i] << 21) ^ a[i];
    }
    const uint2 w = make_uint2(seed | 1u, seed * 3u + 5u);
    uint2 wd[kDiagChains];
#pragma unroll
    for (int i = 0; i < kDiagChains; i++) wd[i] = make_uint2((seed + i * 7919u) | 1u, seed * 3u + i * 104729u + threadIdx.x);
    const uint64_t W = ((uint64_t)w.x << 17) | 1u, Wp = ((uint64_t)w.y << 31) | 7u, twice = q64 << 1;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kDiagIters; it++) {
#pragma unroll
        for (int u = 0; u < kDiagUnroll; u++) {
#pragma unroll
            for (int i = 0; i < kDiagChains; i++) {
                if (KIND == 1) ref_bfly_u64(a64[i], b64[i], W, Wp, q64, twice);
                else if (KIND == 5) ct_bfly(a[i], b[i], wc, lc);
                else if (KIND == 6) ct_bfly(a[i], b[i], wd[i], lc);
                else ct_bfly(a[i], b[i], w, lc);
                if (KIND == 4) asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(e[i]));
                if (KIND == 2 || KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[i]) : "r"(e[(i + 1) % kDiagChains]), "r"(w.y));
                if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[(i + 3) % kDiagChains]) : "r"(e[(i + 5) % kDiagChains]), "r"(w.x));
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kDiagChains; i++) acc ^= a[i] ^ b[i] ^ e[i] ^ wd[i].x ^ (uint32_t)a64[i] ^ (uint32_t)(b64[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace agx
