// agx_tmem_microbench.cu -- can Blackwell tensor memory serve as a per-thread coefficient store for the NTT kernels?
//
// The u32 kernels are bound by dependent IMAD / IMAD.HI chains at 4 warps per scheduler: 64 coefficients + twiddles
// need 128 registers per thread, which caps an SM at 512 threads.  TMEM is as large as the register file (512 columns
// x 128 lanes x 32 bit = 64 K words per SM), unused by this workload, and tcgen05.ld/st.32x32b give every thread of a
// warp private columns of "its" lane.  If a thread kept its 64 coefficients there and pulled 16 at a time into
// registers (radix-16 sub-passes: 4 stages per TMEM round trip), ~56 registers would do and 1024 threads (8 warps per
// scheduler) would fit.  This binary measures whether the TMEM round trips are cheap enough for that:
//   tmem_ld16 / tmem_st16 : raw tcgen05.ld / st .32x32b.x16 rate, 8 CTAs x 128 threads per SM
//   subpass_tmem          : ld16 -> 4 CT stages (32 butterflies, 15 twiddles in registers) -> st16, over 4 column blocks
//   subpass_regs          : the same butterflies on 16 registers without TMEM, same occupancy
// Prints one JSON object per line; butterflies/clk/SM compare with agx_microbench's ct_butterfly / regfile_64coeff.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "agx_arith.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

#define R16(r) "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}"
#define OUT16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
#define IO16(r) "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), \
                "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
#define IN16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 " R16(r) ", [%16];" : OUT16(r) : "r"(taddr) : "memory");
}
// the loaded registers may be used only after the wait: tie them to it so the compiler cannot hoist their uses
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : IO16(r) : : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], " R16(r) ";" : : IN16(r), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int COLS = 64;   // TMEM columns per CTA (128 lanes x 64 columns: 64 words per thread of a 128-thread CTA)

__device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t *slot) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free_cta(uint32_t base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// 4 CT stages on 16 registers: stage j pairs x[k] with x[k + (8 >> j)], 2^j twiddles (w[2^j - 1 + g])
__device__ __forceinline__ void radix16(uint32_t (&x)[16], const uint2 (&w)[15], const agx::LimbConst &c) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int half = 8 >> j;
#pragma unroll
        for (int g = 0; g < (1 << j); g++)
#pragma unroll
            for (int i = 0; i < half; i++) agx::ct_bfly(x[g * 2 * half + i], x[g * 2 * half + i + half], w[(1 << j) - 1 + g], c);
    }
}

enum Mode { M_LD, M_ST, M_SUBPASS_TMEM, M_SUBPASS_REGS };

template <int MODE>
__global__ void __launch_bounds__(128, 8) tmem_kernel(uint32_t *out, long long *cycles, uint32_t seed, agx::LimbConst lc, int reps) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_cta(&slot);
    const uint32_t taddr = base + (((threadIdx.x >> 5) & 3) << 21);   // lane = 32 * (warp % 4) in bits 16+
    uint32_t x[16];
    uint2 w[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = seed * (i + 1) + threadIdx.x * 977u;
#pragma unroll
    for (int i = 0; i < 15; i++) w[i] = make_uint2((seed + i * 7919u) | 1u, seed * 3u + i * 104729u);
#pragma unroll
    for (int b = 0; b < COLS / 16; b++) tmem_st16(taddr + 16 * b, x);
    tmem_st_wait();
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int b = 0; b < COLS / 16; b++) {
            if (MODE == M_LD) {
                uint32_t y[16];
                tmem_ld16(taddr + 16 * b, y);
                tmem_ld_wait(y);
#pragma unroll
                for (int i = 0; i < 16; i++) acc ^= y[i];
            } else if (MODE == M_ST) {
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] += acc + i;
                tmem_st16(taddr + 16 * b, x);
                if (b == COLS / 16 - 1) tmem_st_wait();
            } else if (MODE == M_SUBPASS_TMEM) {
                if (b == 0) tmem_st_wait();          // the previous round's stores of these columns have landed
                tmem_ld16(taddr + 16 * b, x);
                tmem_ld_wait(x);
                radix16(x, w, lc);
                tmem_st16(taddr + 16 * b, x);
            } else {
                radix16(x, w, lc);
            }
        }
#pragma unroll
        for (int i = 0; i < 15; i++) w[i].x += x[i] & 2u;   // keep the twiddle registers live and varying
    }
    tmem_st_wait();
    __syncthreads();
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    tmem_free_cta(base);
}

template <int MODE>
int run(const char *name, int sms, uint32_t *d_out, long long *d_cyc, const agx::LimbConst &lc, int ctas_per_sm) {
    const int reps = 256, grid = sms * ctas_per_sm;
    tmem_kernel<MODE><<<grid, 128>>>(d_out, d_cyc, 777u, lc, reps);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    tmem_kernel<MODE><<<grid, 128>>>(d_out, d_cyc, 777u, lc, reps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[grid / 2];
    const double threads_per_sm = 128.0 * ctas_per_sm;
    if (MODE == M_LD || MODE == M_ST)
        printf("{\"test\": \"%s\", \"warps_per_sched\": %d, \"bytes_per_clk_per_sm\": %.1f, \"clk_per_warp_x16\": %.1f}\n", name,
               ctas_per_sm, reps * 4.0 * 16 * 4 * threads_per_sm / med, med / (reps * 4.0) / (threads_per_sm / 32 / 4));
    else
        printf("{\"test\": \"%s\", \"warps_per_sched\": %d, \"butterflies_per_clk_per_sm\": %.3f, \"event_ms\": %.4f, "
               "\"cta_cycles_min\": %lld, \"cta_cycles_median\": %.0f, \"cta_cycles_max\": %lld, "
               "\"butterflies_per_clk_per_sm_from_event_at_1965MHz\": %.3f}\n", name, ctas_per_sm,
               reps * 4.0 * 32 * threads_per_sm / med, ms, cyc.front(), med, cyc.back(),
               reps * 4.0 * 32 * threads_per_sm / (ms * 1e-3 * 1.965e9));
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    uint32_t *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
    const uint32_t q = 1053818881u;
    agx::LimbConst lc{q, 2 * q, 0u - q, 0u - 2 * q, 0, 29, 0, 0};
    for (int c : {1, 2, 4, 8}) {
        if (run<M_LD>("tmem_ld16", sms, d_out, d_cyc, lc, c)) return 1;
        if (run<M_ST>("tmem_st16", sms, d_out, d_cyc, lc, c)) return 1;
    }
    for (int c : {2, 4, 6, 8}) {
        if (run<M_SUBPASS_REGS>("subpass_regs", sms, d_out, d_cyc, lc, c)) return 1;
        if (run<M_SUBPASS_TMEM>("subpass_tmem", sms, d_out, d_cyc, lc, c)) return 1;
    }
    return 0;
}
