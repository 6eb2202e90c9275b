// agx_diag.cuh -- the measured integer roofline: the butterfly's instruction stream with nothing else around it.
//
// bench.py's `roofline` block needs the rate at which THIS GPU retires the butterfly arithmetic when no memory
// operation is in the way ("the 64/clk figure must be microbenchmarked on the box first").  One CTA per
// SM, every thread runs a long unrolled stream of butterflies on 8 independent register chains:
//   kind 0: the u32 Harvey/Shoup butterfly of the batched kernels (agx_arith.cuh ct_bfly; ntt.cpp:331-369 at u32)
//   kind 1: the u64 butterfly of the reference-shaped path (ref_bfly_u64; ntt.cpp:331-369 verbatim widths)
// Cycles are clock64() deltas of the SM (median over CTAs), so the result -- butterflies per clock per SM -- does not
// depend on the clock the GPU happens to run at; the implied clock (cycles / event time) is returned beside it.
#pragma once
#include "agx_ntt_kernels.cuh"

namespace agx {

constexpr int kDiagChains = 8, kDiagUnroll = 16, kDiagIters = 256;

template <int KIND>
__global__ void __launch_bounds__(1024, 1) diag_bfly_kernel(uint32_t *out, long long *cycles, uint32_t seed, LimbConst lc, uint64_t q64) {
    uint32_t a[kDiagChains], b[kDiagChains];
    uint64_t a64[kDiagChains], b64[kDiagChains];
#pragma unroll
    for (int i = 0; i < kDiagChains; i++) {
        a[i] = seed + threadIdx.x * 977u + i * 131u;
        b[i] = (seed ^ 0x9e3779b9u) + i * 7919u + threadIdx.x;
        a64[i] = ((uint64_t)a[i] << 20) ^ b[i];
        b64[i] = ((uint64_t)b[i] << 21) ^ a[i];
    }
    const uint2 w = make_uint2(seed | 1u, seed * 3u + 5u);
    const uint64_t W = ((uint64_t)w.x << 17) | 1u, Wp = ((uint64_t)w.y << 31) | 7u, twice = q64 << 1;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kDiagIters; it++) {
#pragma unroll
        for (int u = 0; u < kDiagUnroll; u++) {
#pragma unroll
            for (int i = 0; i < kDiagChains; i++) {
                if (KIND == 0) ct_bfly(a[i], b[i], w, lc);
                else ref_bfly_u64(a64[i], b64[i], W, Wp, q64, twice);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kDiagChains; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)a64[i] ^ (uint32_t)(b64[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace agx
