// agx_ntt_r16.cuh -- radix-16 three-pass negacyclic NTT kernels for n = 4096 (sm_100a).
//
// Same job as agx_ntt_kernels.cuh's two-pass kernels (one CTA per polynomial, loader + compute + drain of the reference,
// ntt.cpp:508-640, fused), different geometry: 16 coefficients per thread instead of 64, so a CTA is 256 threads, a
// thread needs ~48 registers instead of 128, and an SM holds 40 warps instead of 16.  The log2(n) = 12 butterfly stages
// (ntt.cpp:146-159 loop nest, :292-300 twiddle index m + i, :331-369 butterfly) run as three register-resident passes
// of four stages with two shared-memory transposes between them:
//
//   forward  pass 0: thread holds x[k*256 + c(tid)], k < 16   -> stages 0..3   twiddles are the same for every thread:
//                                                                              kernel-parameter constants (c[0x0][..]
//                                                                              operands, no loads, no registers)
//            pass 1: thread (hi4, lo4) holds x[hi4*256 + k*16 + lo4] -> stages 4..7   15 twiddles per hi4 (L1 broadcast)
//            pass 2: thread T holds x[T*16 + k]                -> stages 8..11  15 twiddles per thread, coalesced 16-byte
//                                                                              loads from a table in kernel order
//            reduction to [0,q), rows into a dense 128-byte-swizzled tile, ONE TMA tensor store of the 16 KB.
//   inverse  is the mirror image (Gentleman-Sande): TMA tensor load, passes over bits 3..0, 7..4, 11..8, n^-1 folded into
//            the constant twiddles of the last pass (only 1 of a thread's 8 last-stage pairs pays an extra multiply).
//
// Each pass is its own straight-line code (3 x ~3 KB): everything stays in the instruction cache, so -- unlike the
// 64-coefficient kernels, whose two passes must share one copy of the stage code -- every pass gets the twiddle
// addressing that suits it.
//
// Shared-memory image of a polynomial (words): A(idx) = (idx & 15) + 20*((idx >> 4) & 7) + 168*((idx >> 7) & 1)
// + 336*(idx >> 8).  A row of 16 consecutive coefficients is 64 contiguous bytes (LDS.128 / STS.128 in pass 2); rows
// are 80 bytes apart so that eight consecutive rows start in eight different 16-byte bank groups; the bank index
// contributed by idx bit 6 is 16 and by bit 8 is 16, which makes the 32-bit accesses of pass 0 (lanes = idx bits
// 0,1,2,3,6) and pass 1 (lanes = idx bits 0,1,2,3,8) conflict-free too.  Every access is a per-thread base plus a
// compile-time offset.
#pragma once
#include "agx_ntt_kernels.cuh"

namespace agx {

constexpr int kR16UFwd = 16;     // uniform (pass-0) forward twiddles: entry k = roots[k], k < 16 (entry 0 unused)
constexpr int kR16UInv = 24;     // uniform (last-pass) inverse twiddles with n^-1 folded in, layout below

// Inverse uniform-twiddle layout (iroots[k] = psi^-bitrev(k); s = n^-1):
//   [0..8)   local stage 3, group g: iroots[8+g]*s                      (every pair: first scaling)
//   [8..12)  local stage 2, group g: iroots[4+g]*s    [12..16) iroots[4+g]   (scaled / unscaled variant)
//   [16..18) local stage 1, group g: iroots[2+g]*s    [18..20) iroots[2+g]
//   [20] iroots[1]   [21] s   [22] iroots[1]*s
struct R16Params {
    const uint2 *tw_fwd;     // [L][n]  natural order for k < 256, kernel order (tw_pos<12,4>) above
    const uint2 *tw_inv;
    const uint2 *u_fwd;      // [L][kR16UFwd]  multi-limb launches read the uniform twiddles from here
    const uint2 *u_inv;      // [L][kR16UInv]
    const LimbConst *lc;     // [L]
    uint32_t L;
    LimbConst c0;            // limb 0 by value (single-limb launches: constant-bank operands)
    uint2 u[kR16UInv];       // limb 0's uniform twiddles of THIS launch's direction, by value
};

template <int LOGN>
struct R16 {
    static_assert(LOGN == 12, "the radix-16 three-pass geometry is laid out for n = 4096");
    static constexpr int N = 1 << LOGN;
    static constexpr int TPP = N / 16;                     // threads per polynomial = CTA size
    static constexpr int LT = LOGN - 4;
    static constexpr int ROWS = N / 32;                    // 128-byte rows of the dense tile
    static constexpr int IMG_WORDS = 15 + 20 * 7 + 168 + 336 * 15 + 1;   // 5364
    static constexpr int SMEM_BYTES = ((IMG_WORDS * 4 + 15) / 16) * 16;     // 21 456 B: the padded image; the dense 16 KB tile aliases it
    static_assert(SMEM_BYTES >= N * 4, "the dense tile must fit the image buffer");
};

__host__ __device__ constexpr uint32_t r16_img(uint32_t idx) {
    return (idx & 15u) + 20u * ((idx >> 4) & 7u) + 168u * ((idx >> 7) & 1u) + 336u * (idx >> 8);
}

// kernel-order position of natural table entry k (n = 2^logn, 16 coefficients per thread)
__host__ __device__ inline uint32_t r16_tw_pos(uint32_t k, uint32_t logn) {
    const uint32_t lt = logn - 4;
    if (k < (1u << lt)) return k;
    uint32_t s = 31;
    while (!((k >> s) & 1u)) s--;
    const uint32_t c = 1u << (s - lt), rr = k - (1u << s), tpp = 1u << lt;
    const uint32_t T = rr / c, kk = rr % c;
    return c == 1 ? (1u << s) + T : (1u << s) + ((kk >> 1) * tpp + T) * 2 + (kk & 1);
}

// natural-order table -> r16 kernel order (one thread per entry; the inverse table's entry 0 keeps whatever nat has)
__global__ void __launch_bounds__(256) r16_relayout_kernel(uint2 *__restrict__ dst, const uint2 *__restrict__ nat, uint32_t logn,
                                                           uint32_t entries) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= entries) return;
    const uint32_t n = 1u << logn, k = gid & (n - 1);
    dst[(gid - k) + r16_tw_pos(k, logn)] = nat[gid];
}

// ----------------------------------------------------------------------------------------------- twiddle access
template <bool CL>
__device__ __forceinline__ uint2 r16_u(const R16Params &p, const uint2 *ug, int i) {
    if constexpr (CL) return p.u[i]; else return __ldg(ug + i);
}

// stage J (0..3) of a pass over 16 registers: pairs x[g*2h + i], x[g*2h + i + h], h = 8 >> J, twiddle w[g]
template <int J>
__device__ __forceinline__ void r16_ct_stage(uint32_t (&x)[16], const uint2 (&w)[1 << J], const LimbConst &c) {
    constexpr int h = 8 >> J;
#pragma unroll
    for (int g = 0; g < (1 << J); g++)
#pragma unroll
        for (int i = 0; i < h; i++) ct_bfly(x[g * 2 * h + i], x[g * 2 * h + i + h], w[g], c);
}
template <int J>
__device__ __forceinline__ void r16_gs_stage(uint32_t (&x)[16], const uint2 (&w)[1 << J], const LimbConst &c) {
    constexpr int h = 8 >> J;
#pragma unroll
    for (int g = 0; g < (1 << J); g++)
#pragma unroll
        for (int i = 0; i < h; i++) gs_bfly(x[g * 2 * h + i], x[g * 2 * h + i + h], w[g], c);
}

// twiddles of stage J of a per-thread pass: 2^J consecutive table entries starting at `first` (16-byte aligned for J >= 1)
template <int J>
__device__ __forceinline__ void r16_load_tw_run(uint2 (&w)[1 << J], const uint2 *first) {
    if constexpr (J == 0) {
        w[0] = ld_twiddle(first);
    } else {
        const uint4 *f4 = reinterpret_cast<const uint4 *>(first);
#pragma unroll
        for (int h = 0; h < (1 << (J - 1)); h++) {
            const uint4 v = ld_twiddle(f4 + h);
            w[2 * h] = make_uint2(v.x, v.y);
            w[2 * h + 1] = make_uint2(v.z, v.w);
        }
    }
}
// ... of the last forward / first inverse pass: kernel order, thread T's h-th 16-byte load at uint4 index 2^(LT+J-1) + h*TPP + T
template <int LOGN, int J>
__device__ __forceinline__ void r16_load_tw_kernel_order(uint2 (&w)[1 << J], const uint2 *tw, uint32_t T) {
    using G = R16<LOGN>;
    if constexpr (J == 0) {
        w[0] = ld_twiddle(tw + (1u << G::LT) + T);
    } else {
        const uint4 *b4 = reinterpret_cast<const uint4 *>(tw) + T;
#pragma unroll
        for (int h = 0; h < (1 << (J - 1)); h++) {
            const uint4 v = ld_twiddle(b4 + ((1 << (G::LT + J - 1)) + h * G::TPP));
            w[2 * h] = make_uint2(v.x, v.y);
            w[2 * h + 1] = make_uint2(v.z, v.w);
        }
    }
}

// pass-0 position of a thread inside a 256-block: lanes cover idx bits 0,1,2,3,6; warps bits 4,5,7
__device__ __forceinline__ uint32_t r16_c0(uint32_t tid) {
    const uint32_t l = tid & 31u, w = tid >> 5;
    return (l & 15u) | ((l >> 4) << 6) | ((w & 3u) << 4) | ((w >> 2) << 7);
}

// dense tile of the polynomial in the TMA engine's 128-byte swizzle: thread T's 16 consecutive coefficients are half of
// 128-byte row T>>1; 16-byte chunk cc of a row lives at chunk position cc ^ (row & 7)
__device__ __forceinline__ uint32_t r16_swz_off(uint32_t T, int c) {   // byte offset of chunk c (0..3) of thread T's run
    const uint32_t row = T >> 1, cc = 4u * (T & 1u) + (uint32_t)c;
    return row * 128u + ((cc ^ (row & 7u)) << 4);
}

// ------------------------------------------------------------------------------------------------------ forward
// dst may equal src (in place).  MUL: spectrum multiplied pointwise by `mul` (same layout; may equal dst) and left in
// [0,2q) -- the middle step of a three-launch polynomial product.
template <int LOGN, bool MUL, bool CL>
__global__ void __launch_bounds__(R16<LOGN>::TPP, 5)
r16_fwd_kernel(uint32_t *dst, const uint32_t *src, const uint32_t *mul, R16Params p, uint32_t T,
               const __grid_constant__ CUtensorMap tmap) {
    using G = R16<LOGN>;
    __shared__ __align__(1024) uint32_t sm[G::SMEM_BYTES / 4];
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_fwd + (size_t)limb * G::N;
    const uint2 *ug = p.u_fwd + (size_t)limb * kR16UFwd;
    const uint32_t *gs = src + (size_t)poly * G::N;

    uint32_t x[16];
    {   // ---- pass 0: stages 0..3, uniform twiddles
        const uint32_t c0 = r16_c0(tid);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = ld_stream(gs + c0 + 256 * k);
        prefetch_ahead<LOGN, G::TPP>(gs, poly, T, tid);
        { const uint2 w[1] = {r16_u<CL>(p, ug, 1)}; r16_ct_stage<0>(x, w, c); }
        { const uint2 w[2] = {r16_u<CL>(p, ug, 2), r16_u<CL>(p, ug, 3)}; r16_ct_stage<1>(x, w, c); }
        { const uint2 w[4] = {r16_u<CL>(p, ug, 4), r16_u<CL>(p, ug, 5), r16_u<CL>(p, ug, 6), r16_u<CL>(p, ug, 7)}; r16_ct_stage<2>(x, w, c); }
        { const uint2 w[8] = {r16_u<CL>(p, ug, 8), r16_u<CL>(p, ug, 9), r16_u<CL>(p, ug, 10), r16_u<CL>(p, ug, 11),
                              r16_u<CL>(p, ug, 12), r16_u<CL>(p, ug, 13), r16_u<CL>(p, ug, 14), r16_u<CL>(p, ug, 15)};
          r16_ct_stage<3>(x, w, c); }
        uint32_t *s0 = sm + r16_img(c0);
#pragma unroll
        for (int k = 0; k < 16; k++) s0[336 * k] = x[k];
    }
    __syncthreads();
    {   // ---- pass 1: stages 4..7, thread (hi4, lo4): lanes cover idx bits 0,1,2,3,8
        const uint32_t lo4 = tid & 15u, hi4 = ((tid >> 4) & 1u) | ((tid >> 5) << 1);
        uint32_t *s1 = sm + 336 * hi4 + lo4;
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = s1[20 * (k & 7) + 168 * (k >> 3)];
        { uint2 w[1]; r16_load_tw_run<0>(w, tw + 16 + hi4); r16_ct_stage<0>(x, w, c); }
        { uint2 w[2]; r16_load_tw_run<1>(w, tw + 32 + 2 * hi4); r16_ct_stage<1>(x, w, c); }
        { uint2 w[4]; r16_load_tw_run<2>(w, tw + 64 + 4 * hi4); r16_ct_stage<2>(x, w, c); }
        { uint2 w[8]; r16_load_tw_run<3>(w, tw + 128 + 8 * hi4); r16_ct_stage<3>(x, w, c); }
#pragma unroll
        for (int k = 0; k < 16; k++) s1[20 * (k & 7) + 168 * (k >> 3)] = x[k];      // in place: own elements only
    }
    __syncthreads();
    {   // ---- pass 2: stages 8..11, thread T holds 16 consecutive coefficients
        const uint4 *s2 = reinterpret_cast<const uint4 *>(sm + 20 * (tid & 7u) + 168 * ((tid >> 3) & 1u) + 336 * (tid >> 4));
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++) {
            const uint4 v = s2[cidx];
            x[4 * cidx] = v.x; x[4 * cidx + 1] = v.y; x[4 * cidx + 2] = v.z; x[4 * cidx + 3] = v.w;
        }
        { uint2 w[1]; r16_load_tw_kernel_order<LOGN, 0>(w, tw, tid); r16_ct_stage<0>(x, w, c); }
        { uint2 w[2]; r16_load_tw_kernel_order<LOGN, 1>(w, tw, tid); r16_ct_stage<1>(x, w, c); }
        { uint2 w[4]; r16_load_tw_kernel_order<LOGN, 2>(w, tw, tid); r16_ct_stage<2>(x, w, c); }
        { uint2 w[8]; r16_load_tw_kernel_order<LOGN, 3>(w, tw, tid); r16_ct_stage<3>(x, w, c); }
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = reduce4q(x[j], c);
    }
    __syncthreads();                                   // everybody has read its row: the image may be overwritten
    if constexpr (MUL) {
        // the other spectrum arrives through the (now free) buffer as a dense swizzled tile by TMA load? -- no: plain
        // coalesced 16-byte loads into the dense tile layout, then each thread reads its own run
        const uint4 *m4 = reinterpret_cast<const uint4 *>(mul + (size_t)poly * G::N);
        char *tile = reinterpret_cast<char *>(sm);
#pragma unroll
        for (int i = 0; i < 4; i++) {                   // chunk i*256 + tid of the polynomial: row-major dense, unswizzled
            const uint4 v = ld_stream(m4 + i * G::TPP + tid);
            *reinterpret_cast<uint4 *>(tile + (i * G::TPP + tid) * 16) = v;
        }
        __syncthreads();
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++) {
            const uint4 mv = *reinterpret_cast<const uint4 *>(tile + tid * 64 + cidx * 16);
            x[4 * cidx + 0] = csub(barrett_mul_lazy(mv.x, x[4 * cidx + 0], c), c.neg2q);
            x[4 * cidx + 1] = csub(barrett_mul_lazy(mv.y, x[4 * cidx + 1], c), c.neg2q);
            x[4 * cidx + 2] = csub(barrett_mul_lazy(mv.z, x[4 * cidx + 2], c), c.neg2q);
            x[4 * cidx + 3] = csub(barrett_mul_lazy(mv.w, x[4 * cidx + 3], c), c.neg2q);
        }
        __syncthreads();
    }
    {   // ---- results: dense 128-byte-swizzled tile, one tensor store
        char *tile = reinterpret_cast<char *>(sm);
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++)
            *reinterpret_cast<uint4 *>(tile + r16_swz_off(tid, cidx)) =
                make_uint4(x[4 * cidx], x[4 * cidx + 1], x[4 * cidx + 2], x[4 * cidx + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap),
                         "r"(smem_u32(tile)), "r"(0), "r"(poly * G::ROWS) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------------------ inverse
template <int LOGN, bool CL>
__global__ void __launch_bounds__(R16<LOGN>::TPP, 5)
r16_inv_kernel(uint32_t *data, R16Params p, uint32_t T, const __grid_constant__ CUtensorMap tmap) {
    using G = R16<LOGN>;
    __shared__ __align__(1024) uint32_t sm[G::SMEM_BYTES / 4];
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_inv + (size_t)limb * G::N;
    const uint2 *ug = p.u_inv + (size_t)limb * kR16UInv;
    uint32_t *g = data + (size_t)poly * G::N;
    char *tile = reinterpret_cast<char *>(sm);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&mbar, G::N * 4);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(&mbar)), "r"(0), "r"(poly * G::ROWS) : "memory");
    }
    prefetch_ahead<LOGN, G::TPP>(g, poly, T, tid);
    __syncthreads();                                   // the barrier object is initialised for everybody
    mbar_wait(&mbar, 0);

    uint32_t x[16];
    {   // ---- pass 0': stages 11..8 on 16 consecutive coefficients
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(tile + r16_swz_off(tid, cidx));
            x[4 * cidx] = v.x; x[4 * cidx + 1] = v.y; x[4 * cidx + 2] = v.z; x[4 * cidx + 3] = v.w;
        }
        { uint2 w[8]; r16_load_tw_kernel_order<LOGN, 3>(w, tw, tid); r16_gs_stage<3>(x, w, c); }
        { uint2 w[4]; r16_load_tw_kernel_order<LOGN, 2>(w, tw, tid); r16_gs_stage<2>(x, w, c); }
        { uint2 w[2]; r16_load_tw_kernel_order<LOGN, 1>(w, tw, tid); r16_gs_stage<1>(x, w, c); }
        { uint2 w[1]; r16_load_tw_kernel_order<LOGN, 0>(w, tw, tid); r16_gs_stage<0>(x, w, c); }
    }
    __syncthreads();                                   // the dense tile has been read by everybody: build the image over it
    {
        uint4 *s2 = reinterpret_cast<uint4 *>(sm + 20 * (tid & 7u) + 168 * ((tid >> 3) & 1u) + 336 * (tid >> 4));
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++) s2[cidx] = make_uint4(x[4 * cidx], x[4 * cidx + 1], x[4 * cidx + 2], x[4 * cidx + 3]);
    }
    __syncthreads();
    {   // ---- pass 1': stages 7..4
        const uint32_t lo4 = tid & 15u, hi4 = ((tid >> 4) & 1u) | ((tid >> 5) << 1);
        uint32_t *s1 = sm + 336 * hi4 + lo4;
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = s1[20 * (k & 7) + 168 * (k >> 3)];
        { uint2 w[8]; r16_load_tw_run<3>(w, tw + 128 + 8 * hi4); r16_gs_stage<3>(x, w, c); }
        { uint2 w[4]; r16_load_tw_run<2>(w, tw + 64 + 4 * hi4); r16_gs_stage<2>(x, w, c); }
        { uint2 w[2]; r16_load_tw_run<1>(w, tw + 32 + 2 * hi4); r16_gs_stage<1>(x, w, c); }
        { uint2 w[1]; r16_load_tw_run<0>(w, tw + 16 + hi4); r16_gs_stage<0>(x, w, c); }
#pragma unroll
        for (int k = 0; k < 16; k++) s1[20 * (k & 7) + 168 * (k >> 3)] = x[k];
    }
    __syncthreads();
    {   // ---- pass 2': stages 3..0, uniform twiddles with n^-1 folded in (layout at kR16UInv)
        const uint32_t c0 = r16_c0(tid);
        const uint32_t *s0 = sm + r16_img(c0);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = s0[336 * k];
        // local stage 3 (pairs at distance 1): every twiddle scaled -> odd registers now carry n^-1
#pragma unroll
        for (int gq = 0; gq < 8; gq++) gs_bfly(x[2 * gq], x[2 * gq + 1], r16_u<CL>(p, ug, gq), c);
        // local stage 2 (distance 2): pair index bit 0 set -> both inputs scaled -> unscaled twiddle
#pragma unroll
        for (int gq = 0; gq < 4; gq++)
#pragma unroll
            for (int i = 0; i < 2; i++) gs_bfly(x[4 * gq + i], x[4 * gq + i + 2], r16_u<CL>(p, ug, (i & 1) ? 12 + gq : 8 + gq), c);
        // local stage 1 (distance 4): bits 0,1
#pragma unroll
        for (int gq = 0; gq < 2; gq++)
#pragma unroll
            for (int i = 0; i < 4; i++) gs_bfly(x[8 * gq + i], x[8 * gq + i + 4], r16_u<CL>(p, ug, (i & 3) ? 18 + gq : 16 + gq), c);
        // last stage (distance 8): only pair 0 is still unscaled
        gs_bfly_last(x[0], x[8], r16_u<CL>(p, ug, 21), r16_u<CL>(p, ug, 22), c);
#pragma unroll
        for (int i = 1; i < 8; i++) gs_bfly_last_prescaled(x[i], x[i + 8], r16_u<CL>(p, ug, 20), c);
#pragma unroll
        for (int k = 0; k < 16; k++) st_stream(g + c0 + 256 * k, x[k]);
    }
}

}  // namespace agx
