// agx_microbench.cu -- integer-pipe microbenchmarks that turn the "integer roofline" of the NTT butterfly into a
// MEASURED number on the B200 in front of us ("the 64/clk figure must be
// microbenchmarked on the box first").  Stand-alone binary: prints one JSON object per line.
//
// Each test runs one CTA of 1024 threads per SM (8 warps per scheduler), every thread executing a long unrolled
// stream of the instruction under test on 8 independent register chains; cycles are SM clock64() deltas, so the
// result is lane-operations per clock per SM, independent of DVFS.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "agx_arith.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CH = 8;        // independent chains per thread
constexpr int UN = 16;       // unrolled repetitions of the chain set per loop iteration
constexpr int ITERS = 256;

enum Test { T_IMAD, T_IMADHI, T_IMADWIDE, T_VIADDMNMX, T_IADD3, T_LOP3, T_MIX_IMAD_IADD3, T_BFLY_CT, T_BFLY_GS, T_SHFL, T_DFMA, T_BFLY_CT_FP64, T_BFLY_CT_HYBRID, T_COUNT };
static const char *kNames[] = {"imad_lo", "imad_hi", "imad_wide", "viaddmnmx_u32", "iadd3", "lop3", "mix_imad+iadd3",
                               "ct_butterfly", "gs_butterfly", "shfl_xor", "dfma", "ct_butterfly_fp64q", "ct_butterfly_hybrid"};
// lane-instructions issued per inner step per chain (for instr/clk) and "units" (butterflies) per step
static const int kInstrPerStep[] = {1, 1, 1, 1, 1, 1, 2, 6, 6, 1, 1, 7, 6};

// Butterfly whose Shoup quotient comes from the FP64 pipe instead of IMAD.HI (which runs at half rate):
// L = low word of fma.rm((2^52 + y), sigma, 2^52 - 2^52*sigma), sigma = fl_down(1 + w/q); then
// Q = y*(w+q) + L*(-q) mod 2^32 lies in [0,2q) exactly like the Shoup form (checked exhaustively-at-random on the host).
__device__ __forceinline__ void ct_bfly_fp64q(uint32_t &x, uint32_t &y, uint32_t wt, double sigma, double cc,
                                              const agx::LimbConst &c) {
    const uint32_t tx = agx::csub(x, c.neg2q);
    const double dy = __hiloint2double(0x43300000, (int)y);
    const double r = __fma_rd(dy, sigma, cc);
    const uint32_t L = (uint32_t)__double2loint(r);
    const uint32_t Q = y * wt + L * c.negq;
    x = tx + Q + c.zero;
    y = tx + c.twoq - Q;
}

template <int TEST>
__global__ void __launch_bounds__(1024, 1) bench_kernel(uint32_t *out, long long *cycles, uint32_t seed, agx::LimbConst lc) {
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 977u + i * 131u; b[i] = (seed ^ 0x9e3779b9u) + i * 7919u + threadIdx.x; }
    const uint2 w = make_uint2(seed | 1u, seed * 3u + 5u);
    const double sigma = 1.0 + (double)(seed & 0xffff) / 65536.0, cc = 4503599627370496.0 * (1.0 - sigma);
    double dacc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) dacc[i] = 1.0 + i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (TEST == T_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(w.x));
                else if (TEST == T_IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                else if (TEST == T_IMADWIDE) {
                    unsigned long long r;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a[i]), "r"(b[i]));
                    a[i] = (uint32_t)r; b[i] ^= (uint32_t)(r >> 32);
                } else if (TEST == T_VIADDMNMX) a[i] = __viaddmin_u32(a[i], lc.neg2q, b[i]);
                else if (TEST == T_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                else if (TEST == T_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(w.y));
                else if (TEST == T_MIX_IMAD_IADD3) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(w.x), "r"(w.y));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(w.x));
                } else if (TEST == T_BFLY_CT) agx::ct_bfly(a[i], b[i], w, lc);
                else if (TEST == T_BFLY_GS) agx::gs_bfly(a[i], b[i], w, lc);
                else if (TEST == T_SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
                else if (TEST == T_DFMA) dacc[i] = __fma_rd(dacc[i], sigma, cc);
                else if (TEST == T_BFLY_CT_FP64) ct_bfly_fp64q(a[i], b[i], w.x, sigma, cc, lc);
                else if (TEST == T_BFLY_CT_HYBRID) {
                    if (i & 1) ct_bfly_fp64q(a[i], b[i], w.x, sigma, cc, lc);
                    else agx::ct_bfly(a[i], b[i], w, lc);
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)__double2loint(dacc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int TEST>
int run(int sms, uint32_t *d_out, long long *d_cyc, const agx::LimbConst &lc, cudaStream_t s, int threads = 1024) {
    bench_kernel<TEST><<<sms, threads, 0, s>>>(d_out, d_cyc, 12345u, lc);   // warm-up
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, s));
    bench_kernel<TEST><<<sms, threads, 0, s>>>(d_out, d_cyc, 12345u, lc);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[sms / 2];
    const double steps = (double)ITERS * UN * CH * (double)threads;  // per SM
    const double lane_instr = steps * kInstrPerStep[TEST];
    printf("{\"test\": \"%s\", \"warps_per_sched\": %d, \"lane_instr_per_clk_per_sm\": %.2f, \"units_per_clk_per_sm\": %.3f, \"median_cycles\": %.0f, "
           "\"event_ms\": %.4f, \"implied_sm_mhz\": %.0f}\n",
           kNames[TEST], threads / 128, lane_instr / med, steps / med, med, ms, med / (ms * 1e3));
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Instruction-supply test: a straight-line body of K instructions (alternating IMAD / IADD3 on 8 chains, which the
// mix test above shows can issue at 1 warp-instruction/clk/scheduler) executed `reps` times by W warps per
// scheduler.  desync != 0 delays each CTA of an SM by a different amount so the warps sit at different program
// counters -- the situation of independent polynomial teams in the NTT kernels.  Reports warp-instr/clk/scheduler.
template <int K>
__global__ void __launch_bounds__(128) icache_kernel(uint32_t *out, long long *cycles, int reps, int desync,
                                                     unsigned *sm_slot) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 31u + i;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    __shared__ unsigned slot;
    if (threadIdx.x == 0) slot = atomicAdd(&sm_slot[smid], 1u);
    __syncthreads();
    if (desync) {
        const long long until = clock64() + (long long)slot * desync;
        while (clock64() < until) {}
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int u = 0; u < K / 16; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 4) & 7]) : "r"(a[(i + 5) & 7]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K>
int run_icache(int sms, uint32_t *d_out, long long *d_cyc, unsigned *d_slot, cudaStream_t s) {
    const long long total_instr = 1 << 21;                 // per warp
    const int reps = (int)(total_instr / K);
    for (int wps = 1; wps <= 4; wps *= 2) {                // warps per scheduler = CTAs per SM (4 warps per CTA)
        for (int desync = 0; desync <= 1; desync++) {
            const int ctas = sms * wps;
            CK(cudaMemsetAsync(d_slot, 0, sizeof(unsigned) * 256, s));
            icache_kernel<K><<<ctas, 128, 0, s>>>(d_out, d_cyc, reps, desync ? 3000 + K : 0, d_slot);
            CK(cudaStreamSynchronize(s));
            CK(cudaGetLastError());
            std::vector<long long> cyc(ctas);
            CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
            std::sort(cyc.begin(), cyc.end());
            const double med = (double)cyc[ctas / 2];
            printf("{\"test\": \"icache\", \"body_instr\": %d, \"body_kb\": %.1f, \"warps_per_sched\": %d, \"desync\": %d, "
                   "\"warp_instr_per_clk_per_sched\": %.3f}\n",
                   K, K * 16 / 1024.0, wps, desync, (double)reps * K * wps / med);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Register-file pressure test: the NTT kernels' inner structure without any memory traffic.  Each thread keeps 64
// coefficients in registers and runs passes of 6 stages (stage j pairs x[k] with x[k + (32 >> j)], as the kernels do).
// TW = 1: one twiddle for everything (operand-reuse friendly); TW = 2: one twiddle per stage group, i.e. 2^j distinct
// twiddles in stage j, held in registers -- the kernels' situation.  512 threads/SM at <= 128 registers = 4 warps per
// scheduler, the kernels' occupancy.  Reports butterflies/clk/SM.
template <int TW>
__global__ void __launch_bounds__(512, 1) regfile_kernel(uint32_t *out, long long *cycles, uint32_t seed, agx::LimbConst lc, int reps) {
    uint32_t x[64];
    uint2 w[32];
#pragma unroll
    for (int i = 0; i < 64; i++) x[i] = seed * (i + 1) + threadIdx.x * 977u;
#pragma unroll
    for (int i = 0; i < 32; i++) w[i] = make_uint2((seed + i * 7919u) | 1u, seed * 3u + i * 104729u);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const int half = 32 >> j;
#pragma unroll
            for (int g = 0; g < (1 << j); g++)
#pragma unroll
                for (int i = 0; i < half; i++)
                    agx::ct_bfly(x[g * 2 * half + i], x[g * 2 * half + i + half], TW == 1 ? w[0] : w[g], lc);
        }
        if (TW == 2) {      // keep the twiddle registers live and varying without loads
#pragma unroll
            for (int i = 0; i < 32; i++) w[i].x += x[i] & 2u;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int TW>
int run_regfile(int sms, uint32_t *d_out, long long *d_cyc, const agx::LimbConst &lc, cudaStream_t s) {
    const int reps = 64;
    for (int thr = 128; thr <= 512; thr *= 2) {
        regfile_kernel<TW><<<sms, thr, 0, s>>>(d_out, d_cyc, 777u, lc, reps);
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        std::vector<long long> cyc(sms);
        CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        std::sort(cyc.begin(), cyc.end());
        const double med = (double)cyc[sms / 2];
        printf("{\"test\": \"regfile_64coeff\", \"twiddles\": \"%s\", \"warps_per_sched\": %d, \"butterflies_per_clk_per_sm\": %.3f}\n",
               TW == 1 ? "one" : "per_group", thr / 128, (double)reps * 192.0 * thr / med);
    }
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", prop.name, sms, prop.major, prop.minor);
    uint32_t *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
    const uint32_t q = 1053818881u;
    agx::LimbConst lc{q, 2 * q, 0u - q, 0u - 2 * q, 0, 29, 0, 0};
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    if (run<T_IMAD>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_IMADHI>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_IMADWIDE>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_VIADDMNMX>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_IADD3>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_LOP3>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_MIX_IMAD_IADD3>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_BFLY_CT>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_BFLY_GS>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run<T_SHFL>(sms, d_out, d_cyc, lc, s)) return 1;
    for (int thr = 128; thr <= 1024; thr *= 2) {      // the butterfly at 1, 2, 4, 8 warps per scheduler
        if (run<T_BFLY_CT>(sms, d_out, d_cyc, lc, s, thr)) return 1;
        if (run<T_MIX_IMAD_IADD3>(sms, d_out, d_cyc, lc, s, thr)) return 1;
    }
    if (run<T_DFMA>(sms, d_out, d_cyc, lc, s)) return 1;
    for (int thr = 256; thr <= 1024; thr *= 2) {
        if (run<T_BFLY_CT_FP64>(sms, d_out, d_cyc, lc, s, thr)) return 1;
        if (run<T_BFLY_CT_HYBRID>(sms, d_out, d_cyc, lc, s, thr)) return 1;
    }
    if (run_regfile<1>(sms, d_out, d_cyc, lc, s)) return 1;
    if (run_regfile<2>(sms, d_out, d_cyc, lc, s)) return 1;
    if (getenv("AGX_MB_SKIP_ICACHE")) return 0;
    unsigned *d_slot;
    CK(cudaMalloc(&d_slot, sizeof(unsigned) * 256));
    if (run_icache<128>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<256>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<512>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<1024>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<2048>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<4096>(sms, d_out, d_cyc, d_slot, s)) return 1;
    if (run_icache<8192>(sms, d_out, d_cyc, d_slot, s)) return 1;
    return 0;
}
