// agx_dispatch_microbench.cu -- which instructions share the dispatch slot of the integer multiply?
//
// Round-2 finding (profiles/r02_experiments.md): the butterfly stream loses ~2 cycles per butterfly for EVERY extra
// LOP3, whether it reads one register or three -- as if each "half-rate" (64 lanes/clk/SM) instruction, multiply or not,
// took two cycles of one shared dispatch slot per scheduler, with only full-rate instructions (IADD3) slipping in
// between.  This tool measures, for a list of candidate instructions X, the rate of the pair stream (IMAD ; X) on
// independent registers: ~64 pair-lanes/clk/SM means X rides along with the IMAD for free, ~32 means X costs a slot of
// its own.  One CTA of 1024 threads per SM, cycles from clock64(), median over CTAs.  Prints one JSON line per X.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "agx_arith.cuh"

constexpr int CH = 8, UN = 16, ITERS = 256;
enum X { X_NONE, X_IADD3, X_IADD3_CC, X_LOP3, X_VIADDMNMX, X_ISETP_SEL, X_IMNMX, X_SHF, X_LEA, X_FADD, X_FMNMX, X_FFMA, X_PRMT, X_IMAD2,
         X_IADD3_PRED, X_POPC, X_COUNT };
static const char *kNames[] = {"none (IMAD alone)", "IADD3", "IADD3 carry pair (add.cc + addc)", "LOP3", "VIADDMNMX", "ISETP+SEL", "IMNMX",
                               "SHF", "LEA", "FADD", "FMNMX", "FFMA", "PRMT", "second IMAD", "ISETP + predicated IADD3", "POPC"};
static const int kExtra[] = {0, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1};

template <int T>
__global__ void __launch_bounds__(1024, 1) k(uint32_t *out, long long *cycles, uint32_t seed, uint32_t m) {
    uint32_t a[CH], b[CH], e[CH], f[CH];
    float g[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        a[i] = seed + threadIdx.x * 977u + i * 131u; b[i] = (seed ^ 0x9e3779b9u) + i * 7919u + threadIdx.x;
        e[i] = seed * (i + 3u) + threadIdx.x; f[i] = e[i] ^ 0x55aa55aau; g[i] = 1.0f + i + threadIdx.x;
    }
    const uint32_t w = seed | 1u;
    const float gw = 1.0f + (seed & 255) * 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(w));
                if (T == X_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(f[i]));
                if (T == X_IADD3_CC) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %2;" : "+r"(e[i]), "+r"(f[i]) : "r"(m));
                if (T == X_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[i]) : "r"(f[i]), "r"(w));
                if (T == X_VIADDMNMX) e[i] = __viaddmin_u32(e[i], m, f[i]);
                if (T == X_ISETP_SEL) asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\tselp.u32 %0, %1, %2, p;\n\t}" : "+r"(e[i]) : "r"(f[i]), "r"(m));
                if (T == X_IMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(f[i]));
                if (T == X_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(e[i]) : "r"(f[i]));
                if (T == X_LEA) e[i] = (e[i] << 3) + f[i];
                if (T == X_FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(gw));
                if (T == X_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(gw));
                if (T == X_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(g[i]) : "f"(gw));
                if (T == X_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3021;" : "+r"(e[i]) : "r"(f[i]));
                if (T == X_IMAD2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e[i]) : "r"(f[i]), "r"(w));
                if (T == X_IADD3_PRED) asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p add.u32 %0, %0, %2;\n\t}" : "+r"(e[i]) : "r"(f[i]), "r"(m));
                if (T == X_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(e[i]));
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i] ^ e[i] ^ f[i] ^ __float_as_uint(g[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int T>
int run(int sms, uint32_t *d_out, long long *d_cyc) {
    for (int rep = 0; rep < 2; rep++) k<T><<<sms, 1024>>>(d_out, d_cyc, 12345u, 0xC12F4A81u);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"launch\"}\n"); return 1; }
    std::vector<long long> cyc(sms);
    cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[sms / 2], steps = (double)ITERS * UN * CH * 1024.0;
    printf("{\"with\": \"%s\", \"pair_lanes_per_clk_per_sm\": %.1f, \"cycles_per_warp_pair_per_scheduler\": %.2f, \"extra_instructions\": %d}\n",
           kNames[T], steps / med, med / ((double)ITERS * UN * CH * 8.0), kExtra[T]);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Part 2: the butterfly stream itself (agx_arith.cuh's arithmetic) with (a) the conditional subtract written two ways
// and (b) one extra instruction of a given kind every `EVERY` butterflies -- what a real kernel's loads, stores and
// final reduction cost the multiply pipe.  Reports butterflies/clk/SM at 4 and 8 warps per scheduler.
enum CS { CS_VIADDMNMX, CS_ADD_MIN };
enum EX { EX_NONE, EX_LOP3, EX_LDS, EX_VIADDMNMX, EX_IADD3, EX_VIMNMX, EX_STS };
static const char *kCs[] = {"VIADDMNMX", "IADD3 + VIMNMX3"};
static const char *kEx[] = {"none", "LOP3", "LDS.32", "VIADDMNMX", "IADD3", "VIMNMX3", "STS.32"};

template <int CSV, int EXV, int EVERY>
__global__ void __launch_bounds__(1024, 1) bf(uint32_t *out, long long *cycles, uint32_t seed, agx::LimbConst c) {
    __shared__ uint32_t sh[1024 + 64];
    uint32_t a[CH], b[CH], e[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 977u + i * 131u; b[i] = (seed ^ 0x9e3779b9u) + i * 7919u + threadIdx.x; e[i] = seed * (i + 3u) + threadIdx.x; }
    sh[threadIdx.x] = seed;
    const uint2 w = make_uint2(seed | 1u, seed * 3u + 5u);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (CSV == CS_VIADDMNMX) agx::ct_bfly(a[i], b[i], w, c);
                else {
                    uint32_t t, tx;
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(c.neg2q));      // volatile: keep ptxas from re-fusing
                    asm volatile("min.u32 %0, %1, %2;" : "=r"(tx) : "r"(t), "r"(a[i]));
                    const uint32_t Q = agx::shoup_mul(b[i], w, c.negq);
                    a[i] = tx + Q + c.zero;
                    b[i] = tx + c.twoq - Q;
                }
                if ((u * CH + i) % EVERY == 0) {
                    if (EXV == EX_LOP3) asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(e[i]));
                    if (EXV == EX_LDS) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e[i]) : "r"((uint32_t)__cvta_generic_to_shared(sh + threadIdx.x + i)));
                    if (EXV == EX_STS) asm volatile("st.shared.u32 [%1], %0;" :: "r"(e[i]), "r"((uint32_t)__cvta_generic_to_shared(sh + threadIdx.x + i)) : "memory");
                    if (EXV == EX_VIADDMNMX) e[i] = __viaddmin_u32(e[i], c.neg2q, e[(i + 1) % CH]);
                    if (EXV == EX_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(e[(i + 1) % CH]));
                    if (EXV == EX_VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(e[(i + 1) % CH]));
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i] ^ b[i] ^ e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CSV, int EXV, int EVERY>
int run_bf(int sms, uint32_t *d_out, long long *d_cyc) {
    double r[2];
    int k = 0;
    const uint32_t q = 1053818881u;
    const agx::LimbConst lc{q, 2 * q, 0u - q, 0u - 2 * q, 0, 29, 0, 0};
    for (int thr : {512, 1024}) {
        for (int rep = 0; rep < 2; rep++) bf<CSV, EXV, EVERY><<<sms, thr>>>(d_out, d_cyc, 12345u, lc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"launch\"}\n"); return 1; }
        std::vector<long long> cyc(sms);
        cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        std::sort(cyc.begin(), cyc.end());
        r[k++] = (double)ITERS * UN * CH * thr / (double)cyc[sms / 2];
    }
    printf("{\"csub\": \"%s\", \"extra\": \"%s\", \"extra_per_butterfly\": %.3f, \"butterflies_per_clk_per_sm_4w\": %.2f, \"8w\": %.2f}\n",
           kCs[CSV], kEx[EXV], EXV == EX_NONE ? 0.0 : 1.0 / EVERY, r[0], r[1]);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    int rc = 0;
    rc |= run<X_NONE>(sms, d_out, d_cyc); rc |= run<X_IADD3>(sms, d_out, d_cyc); rc |= run<X_IADD3_CC>(sms, d_out, d_cyc);
    rc |= run<X_LOP3>(sms, d_out, d_cyc); rc |= run<X_VIADDMNMX>(sms, d_out, d_cyc); rc |= run<X_ISETP_SEL>(sms, d_out, d_cyc);
    rc |= run<X_IMNMX>(sms, d_out, d_cyc); rc |= run<X_SHF>(sms, d_out, d_cyc); rc |= run<X_LEA>(sms, d_out, d_cyc);
    rc |= run<X_FADD>(sms, d_out, d_cyc); rc |= run<X_FMNMX>(sms, d_out, d_cyc); rc |= run<X_FFMA>(sms, d_out, d_cyc);
    rc |= run<X_PRMT>(sms, d_out, d_cyc); rc |= run<X_IMAD2>(sms, d_out, d_cyc); rc |= run<X_IADD3_PRED>(sms, d_out, d_cyc);
    rc |= run<X_POPC>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_NONE, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_ADD_MIN, EX_NONE, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LOP3, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LOP3, 2>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LOP3, 4>(sms, d_out, d_cyc);
    rc |= run_bf<CS_ADD_MIN, EX_LOP3, 2>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LDS, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LDS, 2>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_LDS, 4>(sms, d_out, d_cyc);
    rc |= run_bf<CS_ADD_MIN, EX_LDS, 2>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_STS, 2>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_VIADDMNMX, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_VIADDMNMX, 3>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_IADD3, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_IADD3, 3>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_VIMNMX, 1>(sms, d_out, d_cyc);
    rc |= run_bf<CS_VIADDMNMX, EX_VIMNMX, 3>(sms, d_out, d_cyc);
    return rc;
}
