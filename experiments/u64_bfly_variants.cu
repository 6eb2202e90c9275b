// u64_bfly_variants.cu -- instruction-count study of the 64-bit Harvey/Shoup butterfly (ntt.cpp:331-369 at the reference's
// own word width).  Each kernel runs the same isolated stream as agx_diag.cuh kind 1 on a different formulation of the
// butterfly; `cuobjdump -sass` on the cubin gives the instruction mix, the binary prints butterflies/clk/SM and checks that
// every variant produces the same bits as variant 0 (the shipped round-2 form).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bin/u64_bfly_variants u64_bfly_variants.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Q64 {
    uint64_t q, twice, negq, zero;
};

// variant 0: the shipped form
__device__ __forceinline__ void bfly_v0(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y - c1 * c.q;
    x = tx + Q;
    y = tx + c.twice - Q;
}

// variant 1: -q precomputed (no negation of c1, no final subtraction), three-input add keeps x on the ALU pipe
__device__ __forceinline__ void bfly_v1(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q + c.zero;
    y = tx + c.twice - Q;
}

// variant 3: -q precomputed only
__device__ __forceinline__ void bfly_v3(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = tx + c.twice - Q;
}

// variant 4: -q precomputed, conditional subtraction as an unsigned 64-bit minimum
__device__ __forceinline__ void bfly_v4(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    const uint64_t d = x - c.twice;
    const uint64_t tx = d < x ? d : x;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = tx + c.twice - Q;
}

__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// a*b + c with a 64-bit result (never overflows for c < 2^32 ... any c such that a*b + c < 2^64)
__device__ __forceinline__ uint64_t madwide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint32_t madlo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// variant 2: everything spelled out in 32-bit pieces.
//   mulhi64: hi32(y0*p0) by mul.hi; the two cross products accumulate it; the top product accumulates their high words
//   Q = W*y + c1*(-q) mod 2^64: one wide product per term for the low words, cross terms accumulated into the high word
//   conditional subtraction through the borrow of the 64-bit subtraction
__device__ __forceinline__ void bfly_v2(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    // tx = x >= 2q ? x - 2q : x
    uint32_t d0, d1, m;
    asm("{\n\t"
        "sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, %6;\n\t"
        "subc.u32 %2, 0, 0;\n\t"
        "}" : "=r"(d0), "=r"(d1), "=r"(m) : "r"(lo32(x)), "r"(hi32(x)), "r"(lo32(c.twice)), "r"(hi32(c.twice)));
    const uint32_t t0 = m ? lo32(x) : d0, t1 = m ? hi32(x) : d1;
    const uint64_t tx = pack(t0, t1);
    const uint32_t y0 = lo32(y), y1 = hi32(y), p0 = lo32(Wp), p1 = hi32(Wp);
    const uint32_t ll = __umulhi(y0, p0);
    const uint64_t m1 = madwide(y0, p1, (uint64_t)ll);
    const uint64_t m2 = madwide(y1, p0, (uint64_t)lo32(m1));
    const uint64_t c1 = madwide(y1, p1, (uint64_t)hi32(m1)) + hi32(m2);
    const uint32_t w0 = lo32(W), w1 = hi32(W), n0 = lo32(c.negq), n1 = hi32(c.negq), c10 = lo32(c1), c11 = hi32(c1);
    uint64_t A = madwide(w0, y0, 0ull);
    uint32_t ah = hi32(A);
    ah = madlo(w0, y1, ah);
    ah = madlo(w1, y0, ah);
    uint64_t Qv = madwide(c10, n0, pack(lo32(A), ah));
    uint32_t qh = hi32(Qv);
    qh = madlo(c10, n1, qh);
    qh = madlo(c11, n0, qh);
    const uint64_t Q = pack(lo32(Qv), qh);
    x = tx + Q + c.zero;
    y = tx + c.twice - Q;
}

constexpr int CH = 8, UN = 16, ITERS = 128;

template <int V>
__global__ void __launch_bounds__(512, 1) stream_kernel(uint64_t *out, long long *cycles, uint64_t seed, Q64 c, uint64_t W, uint64_t Wp) {
    uint64_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {      // [0, 2q): small enough for every modulus, lazy enough to exercise the correction
        a[i] = (seed * (2 * i + 1) + threadIdx.x * 0x9e3779b97f4a7c15ull) & (c.q - 1);
        b[i] = ((seed ^ 0x5851f42d4c957f2dull) * (2 * i + 3) + threadIdx.x * 0xda942042e4dd58b5ull) & (c.q - 1);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (V == 0) bfly_v0(a[i], b[i], W, Wp, c);
                else if (V == 1) bfly_v1(a[i], b[i], W, Wp, c);
                else if (V == 2) bfly_v2(a[i], b[i], W, Wp, c);
                else if (V == 3) bfly_v3(a[i], b[i], W, Wp, c);
                else bfly_v4(a[i], b[i], W, Wp, c);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i] * 3 + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
int run(int sms, uint64_t *d_out, long long *d_cyc, const Q64 &c, std::vector<uint64_t> &result, const char *name) {
    const int threads = 512;
    const uint64_t W = (777ull * 0x9e3779b97f4a7c15ull) % c.q;
    const uint64_t Wp = (uint64_t)(((unsigned __int128)W << 64) / c.q);
    stream_kernel<V><<<sms, threads>>>(d_out, d_cyc, 777, c, W, Wp);
    CK(cudaDeviceSynchronize());
    stream_kernel<V><<<sms, threads>>>(d_out, d_cyc, 777, c, W, Wp);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    result.resize((size_t)sms * threads);
    CK(cudaMemcpy(result.data(), d_out, sizeof(uint64_t) * result.size(), cudaMemcpyDeviceToHost));
    const double med = (double)cyc[sms / 2];
    printf("{\"variant\": \"%s\", \"q_bits\": %d, \"warps_per_sched\": %d, \"butterflies_per_clk_per_sm\": %.3f}\n", name,
           64 - __builtin_clzll(c.q), threads / 128, (double)ITERS * UN * CH * threads / med);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint64_t *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint64_t) * sms * 512));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    const uint64_t primes[] = {1053818881ull, 1152921504606584833ull /* 60-bit */, 9223372036854677505ull /* 63-bit, wraps */};
    for (uint64_t q : primes) {
        Q64 c{q, q << 1, 0 - q, 0};
        std::vector<uint64_t> r0, r1, r2, r3, r4;
        if (run<0>(sms, d_out, d_cyc, c, r0, "v0_shipped")) return 1;
        if (run<1>(sms, d_out, d_cyc, c, r1, "v1_negq_add3")) return 1;
        if (run<2>(sms, d_out, d_cyc, c, r2, "v2_pieces")) return 1;
        if (run<3>(sms, d_out, d_cyc, c, r3, "v3_negq")) return 1;
        if (run<4>(sms, d_out, d_cyc, c, r4, "v4_negq_min")) return 1;
        printf("{\"v4_equals_v0\": %s}\n", r4 == r0 ? "true" : "false");
        printf("{\"q_bits\": %d, \"v1_equals_v0\": %s, \"v2_equals_v0\": %s, \"v3_equals_v0\": %s}\n", 64 - __builtin_clzll(q),
               r1 == r0 ? "true" : "false", r2 == r0 ? "true" : "false", r3 == r0 ? "true" : "false");
    }
    return 0;
}
