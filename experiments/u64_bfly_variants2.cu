// u64_bfly_variants2.cu -- second instruction-mix study of the 64-bit Harvey/Shoup butterfly (ntt.cpp:331-369): where do the
// non-multiply instructions of the butterfly execute?  IMAD.WIDE issues at half rate on this GPU, so the ~2.8 IMAD.X /
// IMAD.IADD / IMAD.MOV per butterfly that ptxas places on the multiply pipe to "balance" it against the ALU cost 15 % of the
// pipe the butterfly is bound by.  Each variant is checked bit for bit against the shipped form (variant 3 of
// u64_bfly_variants.cu) on a 30-, 60- and 63-bit modulus.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bin/u64_bfly_variants2 u64_bfly_variants2.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Q64 { uint64_t q, twice, negq, zero; };

__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

// shipped form
__device__ __forceinline__ void bfly_ref(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = tx + c.twice - Q;
}

// conditional subtraction through the borrow of the 64-bit subtraction (no 64-bit compare)
__device__ __forceinline__ uint64_t csub64_borrow(uint64_t x, uint64_t m) {
    uint32_t d0, d1, b;
    asm("{\n\t"
        "sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, %6;\n\t"
        "subc.u32 %2, 0, 0;\n\t"
        "}" : "=r"(d0), "=r"(d1), "=r"(b) : "r"(lo32(x)), "r"(hi32(x)), "r"(lo32(m)), "r"(hi32(m)));
    return pack(b ? lo32(x) : d0, b ? hi32(x) : d1);
}

// V5: the sum output rides on the products' addend (x' = W*y + tx + c1*(-q)); the difference is (2tx + 2q) - x'
__device__ __forceinline__ void bfly_v5(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t xs = W * y + tx + c1 * c.negq;
    y = tx + tx + c.twice - xs;
    x = xs;
}

// V6: V5 with the borrow form of the conditional subtraction
__device__ __forceinline__ void bfly_v6(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    const uint64_t tx = csub64_borrow(x, c.twice);
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t xs = W * y + tx + c1 * c.negq;
    y = tx + tx + c.twice - xs;
    x = xs;
}

// V7: shipped form with the borrow form of the conditional subtraction
__device__ __forceinline__ void bfly_v7(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    const uint64_t tx = csub64_borrow(x, c.twice);
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = tx + c.twice - Q;
}

// V8: mulhi64 as a carry chain of 32-bit multiply-adds (mad.lo.cc / madc.hi.cc): no 64-bit addends to assemble
__device__ __forceinline__ uint64_t mulhi64_chain(uint64_t a, uint64_t b) {
    const uint32_t a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    uint32_t r1, r2, r3;
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "mul.hi.u32 %0, %3, %5;\n\t"            // r1 = hi(a0 b0)
        "mad.lo.cc.u32 %0, %3, %6, %0;\n\t"     // r1 += lo(a0 b1)
        "madc.hi.u32 %1, %3, %6, 0;\n\t"        // r2 = hi(a0 b1) + c
        "mad.lo.cc.u32 %0, %4, %5, %0;\n\t"     // r1 += lo(a1 b0)
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"    // r2 += hi(a1 b0) + c
        "addc.u32 %2, 0, 0;\n\t"                // r3 = c
        "mad.lo.cc.u32 %1, %4, %6, %1;\n\t"     // r2 += lo(a1 b1)
        "madc.hi.u32 %2, %4, %6, %2;\n\t"       // r3 += hi(a1 b1) + c
        "}" : "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(r2, r3);
}
__device__ __forceinline__ void bfly_v8(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = mulhi64_chain(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = tx + c.twice - Q;
}

// V9: three-input 64-bit adds with a run-time zero on both outputs (adds that cannot become IMAD.X / IMAD.IADD)
__device__ __forceinline__ void bfly_v9(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint64_t tx = x;
    if (tx >= c.twice) tx -= c.twice;
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q + c.zero;
    y = tx + c.twice - Q;
}


// V10: V7 with a run-time zero as third addend of the sum output
__device__ __forceinline__ void bfly_v10(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    const uint64_t tx = csub64_borrow(x, c.twice);
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q + c.zero;
    y = tx + c.twice - Q;
}

// 64-bit a + b - s as two three-input adds with their carries (add.cc / addc chains)
__device__ __forceinline__ uint64_t add_sub64(uint64_t a, uint64_t b, uint64_t s) {
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 t0, t1;\n\t"
        "add.cc.u32 t0, %2, %4;\n\t"
        "addc.u32 t1, %3, %5;\n\t"
        "sub.cc.u32 %0, t0, %6;\n\t"
        "subc.u32 %1, t1, %7;\n\t"
        "}" : "=r"(r0), "=r"(r1) : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)), "r"(lo32(s)), "r"(hi32(s)));
    return pack(r0, r1);
}
__device__ __forceinline__ uint64_t add64(uint64_t a, uint64_t b) {
    uint32_t r0, r1;
    asm("{\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.u32 %1, %3, %5;\n\t"
        "}" : "=r"(r0), "=r"(r1) : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
    return pack(r0, r1);
}
// V11: V7 with the output adds spelled as carry chains
__device__ __forceinline__ void bfly_v11(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    const uint64_t tx = csub64_borrow(x, c.twice);
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = add64(tx, Q);
    y = add_sub64(tx, c.twice, Q);
}

// V12: V7, and tx + 2q hoisted out of the difference as its own conditional value: (x >= 2q ? x : x + 2q) - Q
__device__ __forceinline__ void bfly_v12(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    uint32_t d0, d1, b;
    asm("{\n\t"
        "sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, %6;\n\t"
        "subc.u32 %2, 0, 0;\n\t"
        "}" : "=r"(d0), "=r"(d1), "=r"(b) : "r"(lo32(x)), "r"(hi32(x)), "r"(lo32(c.twice)), "r"(hi32(c.twice)));
    const uint64_t tx = pack(b ? lo32(x) : d0, b ? hi32(x) : d1);
    const uint64_t up = x + c.twice;
    const uint64_t t2 = pack(b ? lo32(up) : lo32(x), b ? hi32(up) : hi32(x));      // tx + 2q
    const uint64_t c1 = __umul64hi(y, Wp);
    const uint64_t Q = W * y + c1 * c.negq;
    x = tx + Q;
    y = t2 - Q;
}

constexpr int CH = 8, UN = 16, ITERS = 128;

template <int V>
__device__ __forceinline__ void bfly(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, const Q64 &c) {
    if (V == 0) bfly_ref(x, y, W, Wp, c);
    else if (V == 5) bfly_v5(x, y, W, Wp, c);
    else if (V == 6) bfly_v6(x, y, W, Wp, c);
    else if (V == 7) bfly_v7(x, y, W, Wp, c);
    else if (V == 8) bfly_v8(x, y, W, Wp, c);
    else if (V == 9) bfly_v9(x, y, W, Wp, c);
    else if (V == 10) bfly_v10(x, y, W, Wp, c);
    else if (V == 11) bfly_v11(x, y, W, Wp, c);
    else bfly_v12(x, y, W, Wp, c);
}

template <int V>
__global__ void __launch_bounds__(512, 1) stream_kernel(uint64_t *out, long long *cycles, uint64_t seed, Q64 c, uint64_t W, uint64_t Wp) {
    uint64_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        a[i] = (seed * (2 * i + 1) + threadIdx.x * 0x9e3779b97f4a7c15ull) & (c.q - 1);
        b[i] = ((seed ^ 0x5851f42d4c957f2dull) * (2 * i + 3) + threadIdx.x * 0xda942042e4dd58b5ull) & (c.q - 1);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) bfly<V>(a[i], b[i], W, Wp, c);
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
int run(int sms, uint64_t *d_out, long long *d_cyc, const Q64 &c, std::vector<uint64_t> &result, const char *name,
        const std::vector<uint64_t> *want) {
    const int threads = 512;
    const uint64_t W = (777ull * 0x9e3779b97f4a7c15ull) % c.q;
    const uint64_t Wp = (uint64_t)(((unsigned __int128)W << 64) / c.q);
    for (int rep = 0; rep < 2; rep++) {
        stream_kernel<V><<<sms, threads>>>(d_out, d_cyc, 777, c, W, Wp);
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    result.resize((size_t)sms * threads);
    CK(cudaMemcpy(result.data(), d_out, sizeof(uint64_t) * result.size(), cudaMemcpyDeviceToHost));
    printf("{\"variant\": \"%s\", \"q_bits\": %d, \"warps_per_sched\": %d, \"butterflies_per_clk_per_sm\": %.3f, \"equals_shipped\": %s}\n",
           name, 64 - __builtin_clzll(c.q), threads / 128, (double)ITERS * UN * CH * threads / (double)cyc[sms / 2],
           !want ? "null" : (result == *want ? "true" : "false"));
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
        std::vector<uint64_t> r0, r;
        if (run<0>(sms, d_out, d_cyc, c, r0, "shipped", nullptr)) return 1;
        if (run<5>(sms, d_out, d_cyc, c, r, "v5_sum_on_addend", &r0)) return 1;
        if (run<6>(sms, d_out, d_cyc, c, r, "v6_sum_on_addend_borrow", &r0)) return 1;
        if (run<7>(sms, d_out, d_cyc, c, r, "v7_borrow", &r0)) return 1;
        if (run<8>(sms, d_out, d_cyc, c, r, "v8_mulhi_chain", &r0)) return 1;
        if (run<9>(sms, d_out, d_cyc, c, r, "v9_add3", &r0)) return 1;
        if (run<10>(sms, d_out, d_cyc, c, r, "v10_borrow_add3", &r0)) return 1;
        if (run<11>(sms, d_out, d_cyc, c, r, "v11_borrow_chains", &r0)) return 1;
        if (run<12>(sms, d_out, d_cyc, c, r, "v12_borrow_t2", &r0)) return 1;
    }
    return 0;
}
