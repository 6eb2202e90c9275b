// agx_ntt_kernels.cuh -- register-resident two-pass negacyclic NTT kernels for sm_100a.
//
// Replaces the reference's loader + compute + drain single_tasks (ntt_input_kernel ntt.cpp:508-607,
// fwd_ntt_kernel ntt.cpp:86-506, ntt_output_kernel ntt.cpp:610-640) with ONE kernel per direction: a CTA of
// TPP = n/E threads owns one polynomial (one (batch, limb) row of the [B][L][n] array), each thread keeps
// E = 2^LE coefficients in registers and the log2(n) butterfly stages run as two register-resident passes with
// a single shared-memory transpose between them:
//
//   forward  pass A: thread t holds x[t + TPP*k]   -> stages 0..LE-1     (twiddles identical for all threads)
//            transpose through a padded smem image (STS.32 columns -> LDS.128 rows, both conflict-free)
//            pass B: thread T holds x[E*T .. E*T+E) -> stages LE..logn-1 (per-thread twiddles, coalesced 16-B loads
//                                                      from a table stored in kernel order)
//            final reduction to [0,q); rows into hardware-swizzled smem tiles, written back by TMA tensor stores
//            (or, without a tensor-map encoder, staged through the padded image and coalesced 16-B stores).
//   inverse  is the mirror image (Gentleman-Sande), n^-1 folded into the twiddles of the last four stages (kInvFold).
//   polymul  n = 1024: forward(a), parked in smem, forward(b), pointwise Barrett, inverse -- one launch;
//            n >= 2048: three launches (NTT(a); NTT(b) .* it; INTT), see agx_api.cu.
// Also here: the generic any-n kernel, the limb-wise element-wise ops, the bit-reversal adapter, the reference-shaped
// u64 pass kernels, the device-side table generator, synthetic data and checksum.
//
// The FPGA's X/X2/Xm double-buffer and reorder muxes (ntt.cpp:90-98, 160-289, 397-496) exist to dodge BRAM port
// conflicts and have no arithmetic effect; they have no counterpart here.
//
// Twiddle table (per limb, per direction) `tw[n]` of (w, w') pairs:
//   entries [1, E)      : natural order, entry k = psi^bitrev(k)  (ntt.cpp:298-300 indexing, roots[m + i])
//   entries [2^s, 2^s+1): for stages s >= LE the 2^s entries of the stage are permuted so that thread T's
//                         c = 2^(s-LT) twiddles are reached by coalesced loads:
//                         natural 2^s + T*c + kk  ->  2^s + ((kk>>1)*TPP + T)*2 + (kk&1)   (c >= 2)
//   inverse table: entry 0 = (n^-1, .), entry 1 = (psi^-bitrev(1) * n^-1, .)
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "agx_arith.cuh"

namespace agx {

struct KParams {
    const uint2 *tw_fwd;     // [L][n] kernel order
    const uint2 *tw_inv;     // [L][n] kernel order
    const uint2 *twc_fwd;    // [L][n] column-pass copy: stage j < LE twiddles at the offsets the row pass uses
    const uint2 *twc_inv;
    const LimbConst *lc;     // [L]
    uint32_t L;
    LimbConst c0;                // limb 0's constants by value: kernel parameters live in the constant bank, so the
                                 // butterfly's q / 2q / -q / -2q / 0 operands cost no register-file read when L == 1
};

// Per-limb constants of a polynomial: kernels instantiated with CL = true (single-limb batches) read them from the
// parameter block, i.e. as constant-bank operands; CL = false loads the limb's entry of the table in global memory.
#define AGX_LIMB_CONSTS(CL, p, limb)            \
    LimbConst limb_consts_;                     \
    if constexpr (!(CL)) limb_consts_ = (p).lc[limb]; \
    const LimbConst &c = (CL) ? (p).c0 : limb_consts_

template <int LOGN, int LE>
struct Geo {
    static constexpr int N = 1 << LOGN;
    static constexpr int E = 1 << LE;            // coefficients per thread
    static constexpr int LT = LOGN - LE;
    static constexpr int TPP = 1 << LT;          // threads per polynomial
    static constexpr int CPR = E / 4;            // 16-byte chunks of data per smem row
    static constexpr int PITCH4 = CPR + 1;       // row pitch in 16-byte chunks: one chunk of padding per row
    static constexpr int PITCH = 4 * PITCH4;     // ... in words
    static constexpr int SMEM_CHUNKS = TPP * PITCH4;
    static_assert(LT >= 5 && LT <= LE, "need 32 <= TPP <= E");
};

// kernel-order position of the twiddle for (stage s >= LE, thread T, local index kk)
template <int LOGN, int LE>
__host__ __device__ constexpr uint32_t tw_pos(int s, uint32_t T, uint32_t kk) {
    constexpr int LT = LOGN - LE;
    constexpr uint32_t TPP = 1u << LT;
    const uint32_t c = 1u << (s - LT);
    return c == 1 ? (1u << s) + T : (1u << s) + ((kk >> 1) * TPP + T) * 2 + (kk & 1);
}

// L2 prefetch of the polynomial a CTA slot will work on about half a wave later.  The CTA that owns it then finds its 16 KB in
// L2 (~250 clk) instead of paying a loaded-HBM round trip (~1.8 us measured as the load wait in the phase trace).
#ifndef AGX_PREFETCH_DIST
#define AGX_PREFETCH_DIST (4 * 148)   // half a wave of n = 4096 CTAs ahead: 0.3-0.4 % better than a full wave in two A/B runs
#endif
template <int LOGN, int TPP>
__device__ __forceinline__ void prefetch_ahead(const uint32_t *g_this, uint32_t poly, uint32_t T, uint32_t tid) {
    constexpr int LINES = (4 << LOGN) / 128;                 // 128-byte lines per polynomial
    if (AGX_PREFETCH_DIST > 0 && poly + AGX_PREFETCH_DIST < T) {
        const char *nxt = reinterpret_cast<const char *>(g_this) + (size_t)AGX_PREFETCH_DIST * (4u << LOGN);
#pragma unroll
        for (int i = 0; i < (LINES + TPP - 1) / TPP; i++)
            if (i * TPP + tid < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)(i * TPP + tid) * 128));
    }
}


// Streaming accesses to the polynomial rows: read once / written once per launch, so they must not displace the
// twiddle table from L1 (AGX_STREAM_POLICY 0: ld/st.global.cs; 1: L1::no_allocate loads; experiments in profiles/).
#ifndef AGX_STREAM_POLICY
#define AGX_STREAM_POLICY 0
#endif
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
#if AGX_STREAM_POLICY == 1
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
#if AGX_STREAM_POLICY == 1
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldcs(p);
#endif
}
// Result stores (AGX_STORE_POLICY 0: st.global.cs streaming; 1: default write-back; 2: .cg; 3: .wt; experiments in profiles/)
#ifndef AGX_STORE_POLICY
#define AGX_STORE_POLICY 0
#endif
template <typename V>
__device__ __forceinline__ void st_stream(V *p, V v) {
#if AGX_STORE_POLICY == 1
    *p = v;
#elif AGX_STORE_POLICY == 2
    __stcg(p, v);
#elif AGX_STORE_POLICY == 3
    __stwt(p, v);
#else
    __stcs(p, v);
#endif
}
__device__ __forceinline__ uint2 ld_twiddle(const uint2 *p) {
#if AGX_STREAM_POLICY == 1
    uint2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint4 ld_twiddle(const uint4 *p) {
#if AGX_STREAM_POLICY == 1
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

template <int TPP>
__device__ __forceinline__ void poly_sync() {
    if constexpr (TPP == 32) __syncwarp(); else __syncthreads();
}

// ---------------------------------------------------------------------------------------------- smem transposes

// Shared-memory image of one polynomial: TPP rows of E words, each row padded by 16 bytes (pitch 4E+16 bytes).
// With that pitch every access pattern of the kernels is conflict-free AND a compile-time offset from a per-thread
// base, so no address arithmetic is left in the unrolled code:
//   columns (STS.32/LDS.32): element tid + TPP*k -> row*PITCH + col, lanes on consecutive words;
//   rows (LDS.128/STS.128):  thread T, chunk c -> T*PITCH4 + c; eight consecutive T start 16 bytes apart mod 128;
//   copy in/out (LDS.128/STS.128): chunk index i*TPP + tid -> consecutive chunks of one row per quarter-warp.
// (An earlier XOR swizzle was also conflict-free but cost one LOP3 per access.)
template <int LOGN, int LE>
__device__ __forceinline__ constexpr uint32_t col_off(int k) {
    using G = Geo<LOGN, LE>;
    const int ebase = G::TPP * k;
    return (ebase >> LE) * G::PITCH + (ebase & (G::E - 1));
}

template <int LOGN, int LE>
__device__ __forceinline__ void sts_columns(uint32_t *sw, const uint32_t (&x)[1 << LE], uint32_t tid) {
#pragma unroll
    for (int k = 0; k < (1 << LE); k++) sw[tid + col_off<LOGN, LE>(k)] = x[k];
}

template <int LOGN, int LE>
__device__ __forceinline__ void lds_columns(const uint32_t *sw, uint32_t (&x)[1 << LE], uint32_t tid) {
#pragma unroll
    for (int k = 0; k < (1 << LE); k++) x[k] = sw[tid + col_off<LOGN, LE>(k)];
}

template <int LOGN, int LE>
__device__ __forceinline__ void lds_row(const uint4 *sm, uint32_t (&x)[1 << LE], uint32_t tid) {
    using G = Geo<LOGN, LE>;
    const uint4 *r = sm + tid * G::PITCH4;
#pragma unroll
    for (int c = 0; c < G::CPR; c++) {
        const uint4 v = r[c];
        x[4 * c + 0] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
    }
}

template <int LOGN, int LE>
__device__ __forceinline__ void sts_row(uint4 *sm, const uint32_t (&x)[1 << LE], uint32_t tid) {
    using G = Geo<LOGN, LE>;
    uint4 *r = sm + tid * G::PITCH4;
#pragma unroll
    for (int c = 0; c < G::CPR; c++) r[c] = make_uint4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
}

// coalesced 16-byte copies between the padded smem image and the polynomial's row in global memory
template <int LOGN, int LE>
__device__ __forceinline__ void smem_to_global(const uint4 *sm, uint32_t *g, uint32_t tid) {
    using G = Geo<LOGN, LE>;
    uint4 *g4 = reinterpret_cast<uint4 *>(g) + tid;
    const uint4 *s = sm + (tid / G::CPR) * G::PITCH4 + (tid % G::CPR);
#pragma unroll
    for (int i = 0; i < G::N / 4 / G::TPP; i++)      // chunk i*TPP + tid: row advances by TPP/CPR per step
        st_stream(g4 + i * G::TPP, s[i * (G::TPP / G::CPR) * G::PITCH4]);
}

template <int LOGN, int LE>
__device__ __forceinline__ void global_to_smem(uint4 *sm, const uint32_t *g, uint32_t tid) {
    using G = Geo<LOGN, LE>;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(g) + tid;
    uint4 *s = sm + (tid / G::CPR) * G::PITCH4 + (tid % G::CPR);
    uint4 v[G::N / 4 / G::TPP];
#pragma unroll
    for (int i = 0; i < G::N / 4 / G::TPP; i++) v[i] = ld_stream(g4 + i * G::TPP);
#pragma unroll
    for (int i = 0; i < G::N / 4 / G::TPP; i++) s[i * (G::TPP / G::CPR) * G::PITCH4] = v[i];
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Inverse-kernel input.  Shipped (AGX_INV_INPUT = 3): cp.async (LDGSTS.128: L2 -> shared memory without passing registers
// or allocating in L1), and each warp stages exactly the rows its own threads will read (warp w: rows 32w .. 32w+31,
// 8 KB contiguous in global memory), so a __syncwarp replaces the CTA barrier: inverse 0.4983 -> 0.4943 ms in one call
// (profiles/r02_experiments.md).  A/B values: 0 = LDG.128 -> registers -> STS.128 (global_to_smem above) + CTA barrier
// (round 1), 1 = cp.async + CTA barrier (0.4979-0.4983), 2 = warp-local rows with LDG/STS (0.5011).
#ifndef AGX_INV_INPUT
#define AGX_INV_INPUT 3
#endif
template <int LOGN, int LE>
__device__ __forceinline__ void inv_stage_input(uint4 *sm, const uint32_t *g, uint32_t tid) {
    using G = Geo<LOGN, LE>;
    constexpr int NCH = G::N / 4 / G::TPP;               // 16-byte chunks per thread
    constexpr bool WARP_LOCAL = (AGX_INV_INPUT & 2) != 0 && G::TPP > 32;
    constexpr bool ASYNC = (AGX_INV_INPUT & 1) != 0;
    // thread's chunk i is polynomial chunk c0 + i*STEP (row c / CPR, column c % CPR of the image); STEP is a multiple of
    // CPR, so source and destination are a per-thread base plus a compile-time offset
    constexpr int STEP = WARP_LOCAL ? 32 : G::TPP;
    static_assert(STEP % G::CPR == 0, "a step must cover whole image rows");
    const uint32_t c0 = WARP_LOCAL ? (tid >> 5) * (32 * NCH) + (tid & 31) : tid;
    const uint4 *src = reinterpret_cast<const uint4 *>(g) + c0;
    uint4 *dst = sm + (c0 / G::CPR) * G::PITCH4 + (c0 % G::CPR);
    if constexpr (ASYNC) {
        const uint32_t d32 = smem_u32(dst);
#pragma unroll
        for (int i = 0; i < NCH; i++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + i * (STEP / G::CPR) * G::PITCH4 * 16), "l"(src + i * STEP)
                         : "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        uint4 v[NCH];
#pragma unroll
        for (int i = 0; i < NCH; i++) v[i] = ld_stream(src + i * STEP);
#pragma unroll
        for (int i = 0; i < NCH; i++) dst[i * (STEP / G::CPR) * G::PITCH4] = v[i];
    }
    if constexpr (WARP_LOCAL) __syncwarp(); else poly_sync<G::TPP>();
}

// ---------------------------------------------------------------------------------- results by TMA tensor store
// Forward results leave the SM as E/32 asynchronous tensor stores issued by ONE thread instead of 16 x (LDS.128 +
// STG.128) per thread.  Every thread writes its row into DENSE shared-memory tiles -- one per 32-word half of the rows,
// TPP rows x 128 bytes each -- in the TMA engine's 128-byte swizzle (16-byte chunk index XOR (row & 7): conflict-free
// STS.128 at a pitch of 128 bytes); after a proxy fence and a barrier one thread issues cp.async.bulk.tensor (3-D map
// over the destination batch: {32 words, E/32 halves, T*TPP rows}, box {32, 1, TPP}) and waits only until the tiles have
// been READ.  The tiles must start on a 1024-byte boundary (the swizzle is a function of address bits 7..9).
// Measured and not kept (profiles/r01_experiments.md): the same tiles for the transpose as well, tensor LOADS for the
// inverse kernel's input, tensor stores for its output.

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "AGX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra AGX_DONE;\n"
        "bra AGX_WAIT;\n"
        "AGX_DONE:\n"
        "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int LOGN, int LE>
struct Swz {
    using G = Geo<LOGN, LE>;
    static constexpr int HALVES = G::E / 32;
    static constexpr int HALF_BYTES = G::TPP * 128;
    static constexpr int HALF_WORDS = G::TPP * 32;
    static constexpr int BYTES = HALVES * HALF_BYTES;    // = 4n
};

template <int LOGN, int LE>
__device__ __forceinline__ void sts_row_swz(uint4 *sm, const uint32_t (&x)[1 << LE], uint32_t tid) {
    using S = Swz<LOGN, LE>;
    char *base = reinterpret_cast<char *>(sm) + tid * 128;
    const uint32_t swz = (tid & 7) << 4;
#pragma unroll
    for (int cc = 0; cc < 8; cc++) {
        char *at = base + ((cc << 4) ^ swz);
#pragma unroll
        for (int h = 0; h < S::HALVES; h++) {
            const int j = 4 * (8 * h + cc);
            *reinterpret_cast<uint4 *>(at + h * S::HALF_BYTES) = make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        }
    }
}

// one thread: the CTA's polynomial, shared-memory tiles -> global, asynchronously; returns when the tiles have been read
template <int LOGN, int LE>
__device__ __forceinline__ void tma_store_poly(const uint4 *sm, const CUtensorMap *map, uint32_t poly) {
    using G = Geo<LOGN, LE>;
    using S = Swz<LOGN, LE>;
#pragma unroll
    for (int h = 0; h < S::HALVES; h++)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                     "r"(smem_u32(reinterpret_cast<const char *>(sm) + h * S::HALF_BYTES)), "r"(0), "r"(h), "r"(poly * G::TPP)
                     : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the engine's reads
}

// minimum resident CTAs per SM requested from ptxas (sets the register cap): 64 coefficients/thread -> 512 threads
// per SM (128 registers each), 32 coefficients/thread -> 768 threads (85 registers).
#ifndef AGX_THREADS_E64
#define AGX_THREADS_E64 512
#endif
#ifndef AGX_MINB
#define AGX_MINB(LOGN, LE) ((1 << (LE)) >= 64 ? (AGX_THREADS_E64 >> ((LOGN) - (LE))) : (768 >> ((LOGN) - (LE))))
#endif
// A/B knob: an explicit register cap for the 64-coefficient kernels instead of a minimum CTA count
// (-DAGX_MAXNREG_E64=112 -> 9 CTAs of 64 threads per SM)
#ifdef AGX_MAXNREG_E64
#define AGX_KERNEL_BOUNDS(LOGN, LE) __launch_bounds__(1 << ((LOGN) - (LE))) __maxnreg__((1 << (LE)) >= 64 ? AGX_MAXNREG_E64 : 80)
#else
#define AGX_KERNEL_BOUNDS(LOGN, LE) __launch_bounds__(1 << ((LOGN) - (LE)), AGX_MINB(LOGN, LE))
#endif

// ------------------------------------------------------------------------------ looped two-pass kernels
// The fully unrolled passes above are ~47 KB of SASS per kernel; with 16 warps per SM each at a different point
// of that straight-line code the instruction caches thrash (ncu: stall_no_instruction 2.7 per issue, the top
// stall).  Both passes of a transform apply the SAME register pairing pattern to x[0..E): stage j pairs
// x[k] with x[k + (E/2 >> j)].  Only the twiddle addressing differs, and that is affine in three per-pass
// values (sbase, tstride, toff).  So the kernels below run ONE copy of the stage code inside a 2-trip loop,
// which halves the code footprint and lets it stay resident in the SM's instruction cache.
//   pass over columns (uniform twiddles): sbase = 0,  tstride = 1,   toff = 0
//   pass over rows (per-thread twiddles): sbase = LT, tstride = TPP, toff = tid
//   twiddle (local stage j >= 1, group g) lives at uint4 index (1 << (sbase+j-1)) + (g>>1)*tstride + toff,
//   twiddle (j = 0) at uint2 index (1 << sbase) + toff.
// When LT < LE (n = 2048) the row pass has only LE-1... LT stages: local stage 0 is skipped there (j0 = LE-LT).

// Per-pass twiddle base pointers.  Every twiddle load of the shared stage code is `base + compile-time offset`:
//   row pass   : table = tw (kernel order), b4 = (uint4*)tw + tid,  b2 = tw + (1<<LT) + tid
//   column pass: table = twc (the same offsets, populated only where a thread with tid = 0 would look), b4, b2 alike
// so the loop body contains no address arithmetic (an earlier version computed base + h*stride at run time, which
// ptxas turned into IMAD.WIDE -- on the pipe this kernel is bound by).
struct PassAddr {
    const uint4 *b4;
    const uint2 *b2;
    const uint4 *b4b;    // inverse only: the "already scaled" twiddle variant (see kInvFold); == b4 in the row pass
};

template <int LOGN, int LE>
__device__ __forceinline__ PassAddr pass_addr(const uint2 *__restrict__ table, uint32_t toff, uint32_t variant_b = 0u) {
    const uint4 *b4 = reinterpret_cast<const uint4 *>(table) + toff;
    return PassAddr{b4, table + (1u << (LOGN - LE)) + toff, b4 + variant_b};
}

// Inverse transform, where n^-1 goes.  Every stage multiplies only its difference outputs, so n^-1 cannot ride on one
// stage's twiddles alone.  It is folded into the column pass (the last LE stages) like this: local stage kInvFold
// stores ALL its twiddles pre-multiplied by n^-1, which scales every coefficient whose register index has bit
// half(kInvFold) set.  In the later stages J < kInvFold a pair whose index has any of the bits half(J+1..kInvFold) set
// is already scaled and takes the unscaled twiddle (variant B, stored one 16-byte slot after variant A in the column
// table); a pair with none of them set is still unscaled: its sum stays so, its difference takes the scaled twiddle
// (variant A).  Only the E / 2^(kInvFold+1) last-stage pairs that were sums all the way need the two explicit scaling
// multiplies.  The row pass runs the same stage code with b4b == b4 (both variants are the plain twiddle).
#ifndef AGX_INV_FOLD
#define AGX_INV_FOLD 3
#endif
constexpr int kInvFold = AGX_INV_FOLD;
template <int LE>
__host__ __device__ constexpr int inv_fold_mask(int J) {   // index bits that mark a pair as already scaled at local stage J
    int m = 0;
    for (int jj = J + 1; jj <= kInvFold; jj++) m += (1 << LE) >> (jj + 1);
    return J < kInvFold ? m : 0;
}

// AGX_TW_INTERLEAVE: the two twiddles of a 16-byte table slot are stored word-interleaved, (w0, w1, w0', w1') instead of
// (w0, w0', w1, w1').  A 16-byte load fills four consecutive registers, and the register file has two banks by register
// parity: with (w, w') in an even/odd pair one of the butterfly's two multiplies y*w, hi(y*w') always reads both of its
// operands from y's bank; interleaved, both words of a twiddle share a parity and ptxas can keep y in the other bank.
#ifndef AGX_TW_INTERLEAVE
#define AGX_TW_INTERLEAVE 1
#endif
template <int LOGN, int LE, int J>
__device__ __forceinline__ void load_tw_generic(uint2 (&w)[1 << J], const PassAddr &a) {
    constexpr int LT = LOGN - LE, TPP = 1 << LT;
    if constexpr (J == 0) {
        w[0] = ld_twiddle(a.b2);
    } else {
#pragma unroll
        for (int h = 0; h < (1 << (J - 1)); h++) {
            const uint4 v = ld_twiddle(a.b4 + ((1 << (LT + J - 1)) + h * TPP));
#if AGX_TW_INTERLEAVE
            w[2 * h] = make_uint2(v.x, v.z);
            w[2 * h + 1] = make_uint2(v.y, v.w);
#else
            w[2 * h] = make_uint2(v.x, v.y);
            w[2 * h + 1] = make_uint2(v.z, v.w);
#endif
        }
    }
}

template <int LOGN, int LE, int J>
__device__ __forceinline__ void ct_stage(uint32_t (&x)[1 << LE], const PassAddr &a, const LimbConst &c) {
    constexpr int half = (1 << LE) >> (J + 1);
    uint2 w[1 << J];
    load_tw_generic<LOGN, LE, J>(w, a);
#pragma unroll
    for (int g = 0; g < (1 << J); g++)
#pragma unroll
        for (int i = 0; i < half; i++) ct_bfly(x[g * 2 * half + i], x[g * 2 * half + i + half], w[g], c);
}

template <int LOGN, int LE, int J>
__device__ __forceinline__ void gs_stage(uint32_t (&x)[1 << LE], const PassAddr &a, const LimbConst &c) {
    constexpr int half = (1 << LE) >> (J + 1);
    constexpr int mask = J >= 1 ? inv_fold_mask<LE>(J) : 0;
    uint2 w[1 << J];
    load_tw_generic<LOGN, LE, J>(w, a);
    if constexpr (mask != 0) {
        uint2 wb[1 << J];
        PassAddr ab = a;
        ab.b4 = a.b4b;
        load_tw_generic<LOGN, LE, J>(wb, ab);
#pragma unroll
        for (int g = 0; g < (1 << J); g++)
#pragma unroll
            for (int i = 0; i < half; i++)
                gs_bfly(x[g * 2 * half + i], x[g * 2 * half + i + half], (i & mask) ? wb[g] : w[g], c);
    } else {
#pragma unroll
        for (int g = 0; g < (1 << J); g++)
#pragma unroll
            for (int i = 0; i < half; i++) gs_bfly(x[g * 2 * half + i], x[g * 2 * half + i + half], w[g], c);
    }
}

template <int LOGN, int LE, int J>
__device__ __forceinline__ void ct_stages_from(uint32_t (&x)[1 << LE], const PassAddr &a, const LimbConst &c) {
    ct_stage<LOGN, LE, J>(x, a, c);
    if constexpr (J + 1 < LE) ct_stages_from<LOGN, LE, J + 1>(x, a, c);
}

template <int LOGN, int LE, int J>
__device__ __forceinline__ void gs_stages_down_to1(uint32_t (&x)[1 << LE], const PassAddr &a, const LimbConst &c) {
    gs_stage<LOGN, LE, J>(x, a, c);
    if constexpr (J > 1) gs_stages_down_to1<LOGN, LE, J - 1>(x, a, c);
}

// Last inverse stage (global stage 0).  With the n^-1 folding described at kInvFold only the pairs whose index has none
// of the fold bits set are still unscaled and take the two scaling multiplies (tw[0] = (n^-1, .), tw[1] =
// (iroot1 * n^-1, .)); all others arrive scaled and take the plain butterfly with the unscaled twiddle twc[TPP] = iroot1.
// Saves (E/2)(1 - 2^-kInvFold) Shoup multiplies per thread against scaling every last-stage pair.
template <int LOGN, int LE>
__device__ __forceinline__ void inv_last_stage(uint32_t (&x)[1 << LE], const uint2 *__restrict__ tw,
                                               const uint2 *__restrict__ twc, const LimbConst &c) {
    using G = Geo<LOGN, LE>;
    const uint2 wn = __ldg(tw), w1n = __ldg(tw + 1), w1 = __ldg(twc + G::TPP);
    constexpr int mask = inv_fold_mask<LE>(0);
#pragma unroll
    for (int j = 0; j < G::E / 2; j++) {
        if ((j & mask) == 0) gs_bfly_last(x[j], x[j + G::E / 2], wn, w1n, c);
        else gs_bfly_last_prescaled(x[j], x[j + G::E / 2], w1, c);
    }
}

// dst may equal src (in place: agx_ntt_fwd).  MUL: the three-launch polynomial product's middle step -- the spectrum
// is multiplied pointwise by `mul` (the other operand's spectrum, same layout; may equal dst) before it is stored,
// and left in [0,2q), which is what the inverse kernel accepts.
// TMA: results leave through tma_store_rows (dst is then described by `tmap`); otherwise through the padded staging
// image and coalesced 16-byte stores.
template <int LOGN, int LE, bool MUL, bool CL, bool TMA = false>
__global__ void AGX_KERNEL_BOUNDS(LOGN, LE)
ntt_fwd_loop_kernel(uint32_t *dst, const uint32_t *src, const uint32_t *mul, KParams p, uint32_t T,
                    const __grid_constant__ CUtensorMap tmap) {
    using G = Geo<LOGN, LE>;
    __shared__ __align__(1024) uint4 sm[G::SMEM_CHUNKS];
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_fwd + (size_t)limb * G::N;
    const uint2 *twc = p.twc_fwd + (size_t)limb * G::N;
    const uint32_t *gs = src + (size_t)poly * G::N;
    uint32_t *g = dst + (size_t)poly * G::N;

    uint32_t x[G::E];
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        PassAddr a;
        if (pass == 0) {                             // column pass: x[k] = poly[tid + TPP*k], stages 0..LE-1
            a = pass_addr<LOGN, LE>(twc, 0u);
#pragma unroll
            for (int k = 0; k < G::E; k++) x[k] = ld_stream(gs + tid + G::TPP * k);
            prefetch_ahead<LOGN, G::TPP>(gs, poly, T, tid);
        } else {                                     // row pass: x[j] = poly[E*tid + j], stages LE..logn-1
            a = pass_addr<LOGN, LE>(tw, tid);
            lds_row<LOGN, LE>(sm, x, tid);
        }
        if (G::LT == LE || pass == 0) ct_stage<LOGN, LE, 0>(x, a, c);
        ct_stages_from<LOGN, LE, 1>(x, a, c);
        if (pass == 0) {
            sts_columns<LOGN, LE>(reinterpret_cast<uint32_t *>(sm), x, tid);
            poly_sync<G::TPP>();
        }
    }
#pragma unroll
    for (int j = 0; j < G::E; j++) x[j] = reduce4q(x[j], c);
    if constexpr (MUL) {
        poly_sync<G::TPP>();                         // every thread has read its row of the transpose buffer
        global_to_smem<LOGN, LE>(sm, mul + (size_t)poly * G::N, tid);
        poly_sync<G::TPP>();
#pragma unroll
        for (int cc = 0; cc < G::CPR; cc++) {
            const uint4 mv = sm[tid * G::PITCH4 + cc];
            x[4 * cc + 0] = csub(barrett_mul_lazy(mv.x, x[4 * cc + 0], c), c.neg2q);
            x[4 * cc + 1] = csub(barrett_mul_lazy(mv.y, x[4 * cc + 1], c), c.neg2q);
            x[4 * cc + 2] = csub(barrett_mul_lazy(mv.z, x[4 * cc + 2], c), c.neg2q);
            x[4 * cc + 3] = csub(barrett_mul_lazy(mv.w, x[4 * cc + 3], c), c.neg2q);
        }
    }
    if constexpr (TMA) {
        poly_sync<G::TPP>();                         // the dense tiles overlap other threads' rows of the padded image
        sts_row_swz<LOGN, LE>(sm, x, tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
        poly_sync<G::TPP>();
        if (tid == 0) tma_store_poly<LOGN, LE>(sm, &tmap, poly);
    } else {
        sts_row<LOGN, LE>(sm, x, tid);               // own row only: no barrier needed before
        poly_sync<G::TPP>();
        smem_to_global<LOGN, LE>(sm, g, tid);
    }
}

template <int LOGN, int LE, bool CL>
__global__ void AGX_KERNEL_BOUNDS(LOGN, LE)
ntt_inv_loop_kernel(uint32_t *__restrict__ data, KParams p, uint32_t T) {
    using G = Geo<LOGN, LE>;
    __shared__ uint4 sm[G::SMEM_CHUNKS];
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_inv + (size_t)limb * G::N;
    const uint2 *twc = p.twc_inv + (size_t)limb * G::N;
    uint32_t *g = data + (size_t)poly * G::N;

    uint32_t x[G::E];
#if AGX_INV_INPUT == 0
    global_to_smem<LOGN, LE>(sm, g, tid);
    prefetch_ahead<LOGN, G::TPP>(g, poly, T, tid);
    poly_sync<G::TPP>();
#else
    prefetch_ahead<LOGN, G::TPP>(g, poly, T, tid);
    inv_stage_input<LOGN, LE>(sm, g, tid);
#endif
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        PassAddr a;
        if (pass == 0) {                             // row pass: stages logn-1 .. LE
            a = pass_addr<LOGN, LE>(tw, tid);
            lds_row<LOGN, LE>(sm, x, tid);
        } else {                                     // column pass: stages LE-1 .. 1 (stage 0 below)
            a = pass_addr<LOGN, LE>(twc, 0u, 1u);
            lds_columns<LOGN, LE>(reinterpret_cast<const uint32_t *>(sm), x, tid);
        }
        gs_stages_down_to1<LOGN, LE, LE - 1>(x, a, c);
        if (pass == 0) {
            if (G::LT == LE) gs_stage<LOGN, LE, 0>(x, a, c);
            sts_row<LOGN, LE>(sm, x, tid);
            poly_sync<G::TPP>();
        }
    }
    inv_last_stage<LOGN, LE>(x, tw, twc, c);
#pragma unroll
    for (int k = 0; k < G::E; k++) st_stream(g + tid + G::TPP * k, x[k]);
}

// Fused negacyclic product c = a * b mod (X^n + 1, q): forward(a), forward(b), pointwise, inverse in ONE launch
// (BASELINE.json config 4).  The forward stage code runs 4 times (2 operands x 2 passes) and the inverse stage code
// twice, each from a single copy; NTT(a) waits in a second shared-memory buffer (own row per thread) while b is
// transformed.  HBM traffic: 3 * n * 4 bytes per product.
template <int LOGN, int LE, bool CL>
__global__ void __launch_bounds__(1 << (LOGN - LE), AGX_MINB(LOGN, LE))
polymul_loop_kernel(uint32_t *out, const uint32_t *a, const uint32_t *b, KParams p) {   // out may alias a and/or b
    using G = Geo<LOGN, LE>;
    __shared__ uint4 sm[G::SMEM_CHUNKS];
    __shared__ uint4 park[G::SMEM_CHUNKS];
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *twf = p.tw_fwd + (size_t)limb * G::N, *twfc = p.twc_fwd + (size_t)limb * G::N;
    const uint2 *twi = p.tw_inv + (size_t)limb * G::N, *twic = p.twc_inv + (size_t)limb * G::N;
    const size_t off = (size_t)poly * G::N;

    uint32_t x[G::E];
#pragma unroll 1
    for (int opnd = 0; opnd < 2; opnd++) {
        const uint32_t *src = (opnd == 0 ? a : b) + off;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            PassAddr ad;
            if (pass == 0) {
                ad = pass_addr<LOGN, LE>(twfc, 0u);
#pragma unroll
                for (int k = 0; k < G::E; k++) x[k] = ld_stream(src + tid + G::TPP * k);
            } else {
                ad = pass_addr<LOGN, LE>(twf, tid);
                lds_row<LOGN, LE>(sm, x, tid);
            }
            if (G::LT == LE || pass == 0) ct_stage<LOGN, LE, 0>(x, ad, c);
            ct_stages_from<LOGN, LE, 1>(x, ad, c);
            if (pass == 0) {
                poly_sync<G::TPP>();                 // all rows of the previous operand have been read
                sts_columns<LOGN, LE>(reinterpret_cast<uint32_t *>(sm), x, tid);
                poly_sync<G::TPP>();
            }
        }
#pragma unroll
        for (int j = 0; j < G::E; j++) x[j] = reduce4q(x[j], c);
        if (opnd == 0) sts_row<LOGN, LE>(park, x, tid);   // NTT(a): written and later read by this thread only
    }
#pragma unroll
    for (int cc = 0; cc < G::CPR; cc++) {            // pointwise: NTT(a) .* NTT(b), lazily reduced to [0,2q)
        const uint4 av = park[tid * G::PITCH4 + cc];
        x[4 * cc + 0] = csub(barrett_mul_lazy(av.x, x[4 * cc + 0], c), c.neg2q);
        x[4 * cc + 1] = csub(barrett_mul_lazy(av.y, x[4 * cc + 1], c), c.neg2q);
        x[4 * cc + 2] = csub(barrett_mul_lazy(av.z, x[4 * cc + 2], c), c.neg2q);
        x[4 * cc + 3] = csub(barrett_mul_lazy(av.w, x[4 * cc + 3], c), c.neg2q);
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        PassAddr ad;
        if (pass == 0) {
            ad = pass_addr<LOGN, LE>(twi, tid);
        } else {
            ad = pass_addr<LOGN, LE>(twic, 0u, 1u);
            lds_columns<LOGN, LE>(reinterpret_cast<const uint32_t *>(sm), x, tid);
        }
        gs_stages_down_to1<LOGN, LE, LE - 1>(x, ad, c);
        if (pass == 0) {
            if (G::LT == LE) gs_stage<LOGN, LE, 0>(x, ad, c);
            sts_row<LOGN, LE>(sm, x, tid);           // own row of `sm`: last read by this thread (lds_row of b)
            poly_sync<G::TPP>();
        }
    }
    inv_last_stage<LOGN, LE>(x, twi, twic, c);
#pragma unroll
    for (int k = 0; k < G::E; k++) st_stream(out + off + tid + G::TPP * k, x[k]);
}

// ---------------------------------------------------------------------------------- generic (any n) u32 kernels
// One CTA per polynomial, whole polynomial in dynamic shared memory, radix-2 stages with a barrier between
// them.  Serves sizes outside {1024, 2048, 4096}; natural-order tables (ntt.cpp:298-300 indexing).
template <bool INVERSE>
__global__ void __launch_bounds__(256) ntt_generic_kernel(uint32_t *__restrict__ data, const uint2 *__restrict__ tw_nat,
                                                          const LimbConst *__restrict__ lc, uint32_t L, uint32_t logn) {
    extern __shared__ uint32_t s[];
    const uint32_t n = 1u << logn;
    const uint32_t poly = blockIdx.x, limb = poly % L;
    const LimbConst c = lc[limb];
    const uint2 *tw = tw_nat + (size_t)limb * n;
    uint32_t *g = data + (size_t)poly * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s[i] = g[i];
    __syncthreads();
    if (!INVERSE) {
        uint32_t tlog = logn - 1;
        for (uint32_t m = 1; m < n; m <<= 1, tlog--) {
            for (uint32_t bf = threadIdx.x; bf < n / 2; bf += blockDim.x) {
                const uint32_t i = bf >> tlog, j = bf & ((1u << tlog) - 1);   // ntt.cpp:292-297
                const uint32_t a0 = (i << (tlog + 1)) + j;
                uint32_t x = s[a0], y = s[a0 + (1u << tlog)];
                ct_bfly(x, y, tw[m + i], c);
                s[a0] = x; s[a0 + (1u << tlog)] = y;
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = reduce4q(s[i], c);
    } else {
        uint32_t tlog = 0;
        for (uint32_t h = n >> 1; h >= 1; h >>= 1, tlog++) {
            for (uint32_t bf = threadIdx.x; bf < n / 2; bf += blockDim.x) {
                const uint32_t i = bf >> tlog, j = bf & ((1u << tlog) - 1);
                const uint32_t a0 = (i << (tlog + 1)) + j;
                uint32_t u = s[a0], v = s[a0 + (1u << tlog)];
                gs_bfly(u, v, tw[h + i], c);
                s[a0] = u; s[a0 + (1u << tlog)] = v;
            }
            __syncthreads();
        }
        const uint2 wn = tw[0];   // natural inverse table keeps (n^-1, .) in the unused entry 0
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = csub(shoup_mul(s[i], wn, c.negq), c.negq);
    }
}

// ------------------------------------------------------------------- large sizes (n = 8192, 16384, 32768), u32
// Above the two-pass kernels' reach a polynomial (32-128 KB) no longer fits the register files of one CTA's threads, but it
// fits ONE CTA's shared memory, so it is still read from HBM once and written once.  The log2 n stages (ntt.cpp:146-159 loop
// nest, :292-300 twiddle index m + i) run as passes of up to four stages on 16 coefficients per (virtual) thread -- the
// scheme of the u64 frame kernel below, with the u32 butterflies of agx_arith.cuh and natural-order tables:
//   pass 0 works on index bits [logn-4, logn), the next ones on the four bits below, ... (the last of those may be partial:
//   it then holds bits [4,8) and runs only the stages of the bits not yet done), a final pass on bits 3..0.
// Pass p is described by (lo, jf, sf): element k of virtual thread vt = (t_hi, t_lo) is index (t_hi << (lo+4)) + t_lo + (k << lo),
// local stages jf..3 are global stages sf.., stage j pairs x[g*2h+i] with x[g*2h+i+h], h = 8 >> j, under twiddle
// tw[2^(sf+j-jf) + (t_hi << j) + g].  The inverse runs the same passes backwards with Gentleman-Sande butterflies on the same
// indices of the inverse table and multiplies by n^-1 (entry 0 of that table) on the way out.
// Image in shared memory: one word of padding per 32 (A(idx) = idx + idx/32): conflict-free when lanes walk along
// consecutive indices (lo >= 5) and when every lane reads its own 16 consecutive words (lo = 0).
__device__ __forceinline__ uint32_t big_img(uint32_t idx) { return idx + (idx >> 5); }

template <bool INVERSE, int JF>
__device__ __forceinline__ void big_pass16(uint32_t (&x)[16], const uint2 *__restrict__ tw, uint32_t sf, uint32_t t_hi,
                                           const LimbConst &c) {
#pragma unroll
    for (int jj = JF; jj < 4; jj++) {
        const int j = INVERSE ? 3 + JF - jj : jj;              // forward JF..3, inverse 3..JF
        const int h = 8 >> j;
        const uint32_t tbase = (1u << (sf + j - JF)) + (t_hi << j);
        uint2 w[8];
        if (j == 0) {
            w[0] = __ldg(tw + tbase);
        } else {                                             // 2^j consecutive entries from a 2^j-aligned index: 16-byte pairs
            const uint4 *t4 = reinterpret_cast<const uint4 *>(tw + tbase);
#pragma unroll
            for (int g2 = 0; g2 < (1 << j) / 2; g2++) {
                const uint4 v = __ldg(t4 + g2);
                w[2 * g2] = make_uint2(v.x, v.y);
                w[2 * g2 + 1] = make_uint2(v.z, v.w);
            }
        }
#pragma unroll
        for (int g = 0; g < (1 << j); g++)
#pragma unroll
            for (int i = 0; i < h; i++) {
                if (INVERSE) gs_bfly(x[g * 2 * h + i], x[g * 2 * h + i + h], w[g], c);
                else ct_bfly(x[g * 2 * h + i], x[g * 2 * h + i + h], w[g], c);
            }
    }
}

template <bool INVERSE>
__device__ __forceinline__ void big_pass16_jf(uint32_t (&x)[16], const uint2 *__restrict__ tw, uint32_t jf, uint32_t sf,
                                              uint32_t t_hi, const LimbConst &c) {
    if (jf == 0) big_pass16<INVERSE, 0>(x, tw, sf, t_hi, c);
    else if (jf == 1) big_pass16<INVERSE, 1>(x, tw, sf, t_hi, c);
    else if (jf == 2) big_pass16<INVERSE, 2>(x, tw, sf, t_hi, c);
    else big_pass16<INVERSE, 3>(x, tw, sf, t_hi, c);
}

// LOGN is a template parameter so that every pass's bit position, stride and stage count is a compile-time constant (image
// accesses are base + immediate, the pass loop unrolls into exactly the passes a size runs), as in ref_u64_frame_kernel.
template <bool INVERSE, int LOGN>
__global__ void __launch_bounds__(1024, 1) ntt_big_kernel(uint32_t *data, const uint2 *__restrict__ tw_nat,
                                                           const LimbConst *__restrict__ lc, uint32_t L) {
    extern __shared__ uint32_t bimg[];
    constexpr uint32_t logn = LOGN;
    constexpr uint32_t n = 1u << logn, vthreads = n >> 4;
    const uint32_t tid = threadIdx.x, NT = blockDim.x;
    const uint32_t poly = blockIdx.x, limb = poly % L;
    const LimbConst c = lc[limb];
    const uint2 *tw = tw_nat + (size_t)limb * n;
    uint32_t *g = data + (size_t)poly * n;
    constexpr int np = 2 + (int)((logn - 8 + 3) / 4);          // pass 0, the passes on the bits between, the final pass
    uint32_t x[16];

#pragma unroll
    for (int pp = 0; pp < np; pp++) {
        const int p = INVERSE ? np - 1 - pp : pp;
        uint32_t lo, jf, sf;
        if (p == 0) { lo = logn - 4; jf = 0; sf = 0; }
        else if (p == np - 1) { lo = 0; jf = 0; sf = logn - 4; }
        else {
            const uint32_t rem = logn - 8 - 4 * (uint32_t)(p - 1);
            lo = rem >= 4 ? rem : 4; jf = rem >= 4 ? 0 : 4 - rem; sf = 4 + 4 * (uint32_t)(p - 1);
        }
        // the first pass of either direction reads global memory, the last one writes it: pass 0 by coalesced 4-byte
        // accesses (lanes on consecutive indices), the final pass by four 16-byte accesses of a thread's own 64 bytes
        const bool from_global = pp == 0, to_global = pp == np - 1;
        // Image address of element k: A(idx0 + (k << lo)) = A(idx0) + k * stride + fix(k), all of it a per-thread base plus
        // uniform or compile-time terms: lo >= 5: k << lo is a multiple of 32, stride = 2^lo + 2^(lo-5), fix = 0;
        // lo == 4 (idx0 % 256 < 16): stride = 16, fix = k >> 1;  lo == 0 (idx0 % 16 == 0): stride = 1, fix = 0.
        const uint32_t stride = lo >= 5 ? (1u << lo) + (1u << (lo - 5)) : (1u << lo);
#pragma unroll 1
        for (uint32_t vt = tid; vt < vthreads; vt += NT) {
            const uint32_t t_lo = vt & ((1u << lo) - 1), t_hi = vt >> lo;
            const uint32_t idx0 = (t_hi << (lo + 4)) + t_lo;
            uint32_t *sp = bimg + big_img(idx0);
            if (from_global && INVERSE) {                      // final pass first: lo == 0
                const uint4 *g4 = reinterpret_cast<const uint4 *>(g + idx0);
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                    const uint4 v = g4[k4];
                    x[4 * k4] = v.x; x[4 * k4 + 1] = v.y; x[4 * k4 + 2] = v.z; x[4 * k4 + 3] = v.w;
                }
            } else if (from_global) {
                const uint32_t *gp = g + idx0;
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = __ldcs(gp + ((uint32_t)k << lo));
            } else if (lo == 4) {
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sp[16 * k + (k >> 1)];
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sp[k * stride];
            }
            big_pass16_jf<INVERSE>(x, tw, jf, sf, t_hi, c);
            if (!INVERSE && p == np - 1) {                     // ntt.cpp:377-393
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = reduce4q(x[k], c);
            }
            if (to_global && INVERSE) {                        // pass 0 last: scale by n^-1, coalesced 4-byte stores
                const uint2 wn = __ldg(tw);                    // the inverse table keeps (n^-1, .) in its unused entry 0
                uint32_t *gp = g + idx0;
#pragma unroll
                for (int k = 0; k < 16; k++) __stcs(gp + ((uint32_t)k << lo), csub(shoup_mul(x[k], wn, c.negq), c.negq));
            } else if (to_global) {                            // final pass last: lo == 0
                uint4 *g4 = reinterpret_cast<uint4 *>(g + idx0);
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) g4[k4] = make_uint4(x[4 * k4], x[4 * k4 + 1], x[4 * k4 + 2], x[4 * k4 + 3]);
            } else if (lo == 4) {
#pragma unroll
                for (int k = 0; k < 16; k++) sp[16 * k + (k >> 1)] = x[k];
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) sp[k * stride] = x[k];
            }
        }
        __syncthreads();
    }
}

template <int DUMMY = 0>
__global__ void __launch_bounds__(256) pointwise_generic_kernel(uint32_t *a, const uint32_t *b,   // a may equal b (squaring)
                                                                const LimbConst *__restrict__ lc, uint32_t L,
                                                                uint32_t logn, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const LimbConst c = lc[(i >> logn) % L];
    a[i] = csub(barrett_mul_lazy(a[i], b[i], c), c.neg2q);   // [0,2q): valid inverse input
}

// ------------------------------------------------------------------- limb-wise (RNS) element-wise arithmetic
// The operations a caller performs around the transforms on [B][L][n] data: add, subtract,
// multiply and multiply-accumulate modulo each limb's prime, e.g. on spectra between agx_ntt_fwd and agx_ntt_inv.
// Pure streaming kernels: 16-byte vector accesses, grid-stride, 3 (add/sub/mul) or 4 (mac) u32 streams of HBM traffic.
enum EwOp { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_MAC = 3 };

template <int OP>
__device__ __forceinline__ uint32_t ew_apply(uint32_t cv, uint32_t av, uint32_t bv, const LimbConst &c) {
    if (OP == EW_ADD) return csub(av + bv, c.negq);
    if (OP == EW_SUB) return csub(av + c.q - bv, c.negq);
    const uint32_t m = csub(csub(barrett_mul_lazy(av, bv, c), c.neg2q), c.negq);
    if (OP == EW_MUL) return m;
    return csub(cv + m, c.negq);
}

template <int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(uint4 *c4, const uint4 *a4, const uint4 *b4,
                                                          const LimbConst *__restrict__ lc, uint32_t L, uint32_t logn,
                                                          size_t total4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const LimbConst c = lc[((i * 4) >> logn) % L];
        const uint4 a = a4[i], b = b4[i];
        uint4 r = make_uint4(0, 0, 0, 0);
        if (OP == EW_MAC) r = c4[i];
        r.x = ew_apply<OP>(r.x, a.x, b.x, c);
        r.y = ew_apply<OP>(r.y, a.y, b.y, c);
        r.z = ew_apply<OP>(r.z, a.z, b.z, c);
        r.w = ew_apply<OP>(r.w, a.w, b.w, c);
        c4[i] = r;
    }
}

// ------------------------------------------------------------------------- reference-shaped u64 forward kernel
// Arithmetic of fwd_ntt_kernel (ntt.cpp:146-159, 292-300, 331-332, 344-363, 368-369, 377-393), including
// wrap-around mod 2^64 when the tables are not Shoup pairs (main.cpp:49-55 feeds such data).  One CTA per frame
// (ntt.cpp:579-595 frame layout: low half from `in`, high half from `in2`); work array in dynamic smem when
// N*8 bytes fits, else in the output buffer itself.
__global__ void __launch_bounds__(1024) ref_fwd_u64_kernel(const uint64_t *__restrict__ in, const uint64_t *__restrict__ in2,
                                                           uint64_t *__restrict__ out, const uint64_t *__restrict__ roots,
                                                           const uint64_t *__restrict__ precons, uint64_t q, uint32_t logn,
                                                           int use_smem) {
    extern __shared__ uint64_t s64[];
    const uint32_t N = 1u << logn;
    const size_t base = (size_t)blockIdx.x * N;
    uint64_t *X = use_smem ? s64 : out + base;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) X[i] = i < N / 2 ? in[base + i] : in2[base + i];
    __syncthreads();
    const uint64_t twice = q << 1;
    uint32_t tlog = logn - 1;
    for (uint32_t m = 1; m < N; m <<= 1, tlog--) {
        const bool last = (m == N / 2);
        for (uint32_t bf = threadIdx.x; bf < N / 2; bf += blockDim.x) {
            const uint32_t i = bf >> tlog, j = bf & ((1u << tlog) - 1);
            const uint32_t a0 = (i << (tlog + 1)) + j, a1 = a0 + (1u << tlog);
            const uint64_t W = roots[m + i], Wp = precons[m + i];
            uint64_t tx = X[a0];
            if (tx >= twice) tx -= twice;
            const uint64_t a = X[a1];
            const uint64_t c1 = __umul64hi(a, Wp);
            const uint64_t Q = W * a - c1 * q;
            uint64_t o0 = tx + Q, o1 = tx + twice - Q;
            if (last) {
                if (o0 >= twice) o0 -= twice;
                if (o0 >= q) o0 -= q;
                if (o1 >= twice) o1 -= twice;
                if (o1 >= q) o1 -= q;
            }
            X[a0] = o0; X[a1] = o1;
        }
        __syncthreads();
    }
    if (use_smem)
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) out[base + i] = X[i];
}

// ------------------------------------------------------------------------------------ coefficient-order adapter
// The transforms keep the reference's orders (ntt.cpp: natural in, bit-reversed out; the inverse the other way round).
// Callers that hold or want spectra in natural order (textbook / NTL-style code) permute with this kernel: row i of
// each n-coefficient polynomial moves to bitrev(i).  An involution; in place; one CTA per polynomial, through shared
// memory so that both the global reads and the global writes are coalesced 16-byte accesses.
__global__ void __launch_bounds__(256) bitrev_rows_kernel(uint32_t *__restrict__ data, uint32_t logn) {
    extern __shared__ uint32_t s[];
    const uint32_t n = 1u << logn;
    uint32_t *g = data + ((size_t)blockIdx.x << logn);
    for (uint32_t i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(g + i);
        s[__brev(i) >> (32 - logn)] = v.x;
        s[__brev(i + 1) >> (32 - logn)] = v.y;
        s[__brev(i + 2) >> (32 - logn)] = v.z;
        s[__brev(i + 3) >> (32 - logn)] = v.w;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<uint4 *>(g + i) = make_uint4(s[i], s[i + 1], s[i + 2], s[i + 3]);
}

// ------------------------------------------------------------- reference-shaped u64 forward: register-radix passes
// The reference's native sizes (ntt.h:11-20: 1024, 8192, 16384, 32768) as a sequence of launches over a chunk of
// frames that stays resident in L2: every pass keeps 2^LS coefficients of one butterfly network in registers and
// runs LS stages on them (ntt.cpp:146-159 loop nest, :292-300 twiddle index m + i, :331-369 butterfly), so a frame
// makes logN/4 round trips instead of logN through shared memory, and N = 32768 (256 KB, more than an SM's shared
// memory) needs no special case.  Arithmetic is the u64 butterfly of ref_fwd_u64_kernel, mod 2^64 exactly.
//   strided passes : thread holds p = p_hi * 2^(logN-s0) + k * 2^(logN-s0-LS) + p_lo, k < 2^LS; consecutive threads
//                    take consecutive p_lo (coalesced 8-byte accesses); twiddles are uniform over p_lo.
//   last pass      : the final 4 stages act on 16 consecutive coefficients (128 B) per thread; a warp stages its
//                    512 consecutive coefficients through a private padded shared-memory tile so that global
//                    accesses stay coalesced; includes the final reduction to [0,q) (ntt.cpp:377-393).
// A/B knobs of the 64-bit butterfly and its twiddle loads (profiles/r02_experiments.md).  AGX_U64_VECTW: 0 = 8-byte twiddle
// loads everywhere, 1 = 16-byte pairs in the frame kernel's 512-thread instantiation only (N >= 16384: +2-6 %; the smaller
// CTAs lose up to 7 % to the 32 bytes of spills the pairs cost at 128 registers), 2 = pairs everywhere.
#ifndef AGX_U64_NEGQ
#define AGX_U64_NEGQ 1
#endif
#ifndef AGX_U64_VECTW
#define AGX_U64_VECTW 1
#endif
// AGX_U64_BORROW: the x-correction (ntt.cpp:331-332) decided by the borrow of the 64-bit subtraction it needs anyway
// instead of a separate 64-bit compare: two ISETP less per butterfly, and fewer of the remaining adds placed on the multiply
// pipe by ptxas -- the isolated butterfly stream goes from 3.00-3.05 to 3.36-3.38 butterflies/clk/SM, the frame kernel gains
// 4 % (profiles/r02_u64_borrow_ab.txt).
#ifndef AGX_U64_BORROW
#define AGX_U64_BORROW 1
#endif
__device__ __forceinline__ uint64_t ref_csub_u64(uint64_t x, uint64_t m) {
#if AGX_U64_BORROW
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
    uint32_t d0, d1, b;
    asm("{\n\t"
        "sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, %6;\n\t"
        "subc.u32 %2, 0, 0;\n\t"
        "}" : "=r"(d0), "=r"(d1), "=r"(b) : "r"(x0), "r"(x1), "r"((uint32_t)m), "r"((uint32_t)(m >> 32)));
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(b ? x0 : d0), "r"(b ? x1 : d1));
    return r;
#else
    return x >= m ? x - m : x;
#endif
}

__device__ __forceinline__ void ref_bfly_u64(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, uint64_t q, uint64_t twice) {
    const uint64_t tx = ref_csub_u64(x, twice);   // ntt.cpp:331-332
    const uint64_t c1 = __umul64hi(y, Wp);        // ntt.cpp:344-358
#if AGX_U64_NEGQ
    uint64_t negq;                                // -q through an opaque (pure, hence hoisted and shared) instruction:
    asm("neg.s64 %0, %1;" : "=l"(negq) : "l"(q)); // left as 0 - q the compiler folds the sum below back into a difference
    const uint64_t Q = W * y + c1 * negq;         // ntt.cpp:363  W*y - c1*q mod 2^64; written as a sum the second product
                                                  // accumulates onto the first (ptxas otherwise negates c1 per butterfly:
                                                  // 25.5 instead of 27.5 SASS instructions, profiles/r02_experiments.md)
#else
    const uint64_t Q = W * y - c1 * q;            // ntt.cpp:363
#endif
    x = tx + Q;                                   // ntt.cpp:368
    y = tx + twice - Q;                           // ntt.cpp:369
}

// The 2^j twiddles (and precons) a thread needs at local stage j of a register pass are consecutive table entries starting
// at an index that is a multiple of 2^j (ntt.cpp:298-300: m + i with m = 2^s >= 2^j groups and i = (group of the pass) << j),
// so for j >= 1 they come in as 16-byte pairs: half the load instructions of one LDG.64 per entry.  The tables must be
// 16-byte aligned (cudaMalloc'd ones are; agx_ref_fwd_dev checks the caller's).
template <int CNT, bool VEC>
__device__ __forceinline__ void ref_load_tw_u64(uint64_t (&W)[CNT], uint64_t (&Wp)[CNT], const uint64_t *__restrict__ roots,
                                                const uint64_t *__restrict__ precons, uint32_t tbase) {
    if constexpr (CNT == 1 || !VEC) {
#pragma unroll
        for (int g = 0; g < CNT; g++) {
            W[g] = __ldg(roots + tbase + g);
            Wp[g] = __ldg(precons + tbase + g);
        }
    } else {
        const ulonglong2 *r2 = reinterpret_cast<const ulonglong2 *>(roots + tbase);
        const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(precons + tbase);
#pragma unroll
        for (int g2 = 0; g2 < CNT / 2; g2++) {
            const ulonglong2 a = __ldg(r2 + g2), b = __ldg(p2 + g2);
            W[2 * g2] = a.x; W[2 * g2 + 1] = a.y;
            Wp[2 * g2] = b.x; Wp[2 * g2 + 1] = b.y;
        }
    }
}

// one stage of a register pass over E = 2^LS coefficients: local stage J pairs x[g*2h + i] with x[g*2h + i + h], h = E >> (J+1)
template <int LS, int J, bool VEC>
__device__ __forceinline__ void ref_stage_u64(uint64_t (&x)[1 << LS], const uint64_t *__restrict__ roots,
                                              const uint64_t *__restrict__ precons, uint32_t tbase, uint64_t q, uint64_t twice) {
    constexpr int half = (1 << LS) >> (J + 1);
    uint64_t W[1 << J], Wp[1 << J];
    ref_load_tw_u64<(1 << J), VEC>(W, Wp, roots, precons, tbase);
#pragma unroll
    for (int g = 0; g < (1 << J); g++)
#pragma unroll
        for (int i = 0; i < half; i++) ref_bfly_u64(x[g * 2 * half + i], x[g * 2 * half + i + half], W[g], Wp[g], q, twice);
}

template <int LS, int J = 0>
__device__ __forceinline__ void ref_stages_u64(uint64_t (&x)[1 << LS], const uint64_t *__restrict__ roots,
                                               const uint64_t *__restrict__ precons, uint32_t s0, uint32_t p_hi, uint64_t q) {
    // roots[m + i]: m = 2^s groups, i = p >> (logN - s)
    ref_stage_u64<LS, J, (AGX_U64_VECTW >= 2)>(x, roots, precons, (1u << (s0 + J)) + (p_hi << J), q, q << 1);
    if constexpr (J + 1 < LS) ref_stages_u64<LS, J + 1>(x, roots, precons, s0, p_hi, q);
}

template <int LS>
__global__ void __launch_bounds__(256) ref_u64_strided_pass_kernel(const uint64_t *src, const uint64_t *src2, uint64_t *dst,
                                                                   const uint64_t *__restrict__ roots,
                                                                   const uint64_t *__restrict__ precons, uint64_t q,
                                                                   uint32_t logn, uint32_t s0, uint32_t frames) {
    constexpr int E = 1 << LS;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t frame = gid >> (logn - LS), tidx = gid & ((1u << (logn - LS)) - 1);
    if (frame >= frames) return;
    const uint32_t lo_bits = logn - s0 - LS;
    const uint32_t p_lo = tidx & ((1u << lo_bits) - 1), p_hi = tidx >> lo_bits;
    const uint32_t pos = (p_hi << (logn - s0)) + p_lo;                    // position of element k = 0 inside the frame
    const size_t base = ((size_t)frame << logn) + pos;
    uint64_t x[E];
#pragma unroll
    for (int k = 0; k < E; k++) {                                         // high half of a frame comes from src2 (ntt.cpp:587-589)
        const uint64_t *s = ((pos + ((uint32_t)k << lo_bits)) >> (logn - 1)) ? src2 : src;
        x[k] = s[base + ((size_t)k << lo_bits)];
    }
    ref_stages_u64<LS>(x, roots, precons, s0, p_hi, q);
#pragma unroll
    for (int k = 0; k < E; k++) dst[base + ((size_t)k << lo_bits)] = x[k];
}

__global__ void __launch_bounds__(128) ref_u64_last_pass_kernel(uint64_t *data, const uint64_t *__restrict__ roots,
                                                                const uint64_t *__restrict__ precons, uint64_t q,
                                                                uint32_t logn, uint32_t frames) {
    constexpr int LS = 4, E = 16, PITCH4 = 9;                 // 8 chunks of 16 bytes per thread row + 1 of padding
    __shared__ uint4 tile[4][32 * PITCH4];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t frame = gid >> (logn - LS), tidx = gid & ((1u << (logn - LS)) - 1);
    if (frame >= frames) return;                              // whole warps: a frame has N/16 >= 64 threads
    uint4 *t = tile[wib];
    uint4 *g4 = reinterpret_cast<uint4 *>(data + ((size_t)(gid - lane) << LS));   // the warp's 512 coefficients
#pragma unroll
    for (int i = 0; i < 8; i++) {                             // chunk c = i*32 + lane: row c/8, column c%8
        const uint32_t c = i * 32 + lane;
        t[(c >> 3) * PITCH4 + (c & 7)] = g4[c];
    }
    __syncwarp();
    uint64_t x[E];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 v = t[lane * PITCH4 + c];
        x[2 * c] = (uint64_t)v.x | ((uint64_t)v.y << 32);
        x[2 * c + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
    }
    ref_stages_u64<LS>(x, roots, precons, logn - LS, tidx, q);
    const uint64_t twice = q << 1;
#pragma unroll
    for (int k = 0; k < E; k++) {                             // ntt.cpp:377-393
        x[k] = ref_csub_u64(ref_csub_u64(x[k], twice), q);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; c++)
        t[lane * PITCH4 + c] = make_uint4((uint32_t)x[2 * c], (uint32_t)(x[2 * c] >> 32), (uint32_t)x[2 * c + 1],
                                          (uint32_t)(x[2 * c + 1] >> 32));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t c = i * 32 + lane;
        g4[c] = t[(c >> 3) * PITCH4 + (c & 7)];
    }
}

// ---------------------------------------------------------------- reference-shaped u64 forward: one launch per round
// The whole frame lives in ONE CTA's shared memory (N <= 16384: 128 KB + padding; a B200 SM offers 227 KB), so a frame is
// read from HBM once and written once instead of making logN/4 round trips: loader, compute unit and drain of the
// reference (ntt.cpp:508-607, 86-506, 610-640) in one kernel.  N = 32768 (256 KB) takes two CTAs per frame: stage 0
// pairs x[i] with x[i + N/2] (ntt.cpp:292-300 with m = 1), after which the halves are independent 16384-point
// sub-transforms; each CTA reads both halves, keeps its own output of stage 0 and carries on alone (costs one extra
// stage of arithmetic in 15, no exchange between CTAs).
// 512 threads per SM at <= 128 registers (the 64-bit butterfly with its twiddle pairs in flight does not fit 64, measured:
// 570 bytes of spills), so the 1024 "virtual threads" of a 16384-point frame are two rounds of a 512-thread CTA.
// Every "virtual thread" keeps 16 coefficients in registers per pass (index bits [lo, lo+4) vary inside the thread) and
// runs up to 4 stages on them; passes go from the top index bits down to bit 4 -- the last of those may be partial --
// and a final pass works on bits 3..0, i.e. on 16 consecutive coefficients.  Image in shared memory: a row of 16
// coefficients every 17 words of 8 bytes (A(idx) = idx + idx/16), which makes every pass's 64-bit accesses
// conflict-free: lanes walk along a row when lo >= 4, and down the rows at a 136-byte pitch when lo = 0.
// Arithmetic: ref_bfly_u64 (ntt.cpp:331-369 mod 2^64, any tables), final reduction ntt.cpp:377-393.
template <bool VEC, int JFIRST, int J = JFIRST>
__device__ __forceinline__ void ref_pass16_u64(uint64_t (&x)[16], const uint64_t *__restrict__ roots,
                                               const uint64_t *__restrict__ precons, uint32_t s_first, uint32_t t_hi, uint64_t q,
                                               uint64_t twice) {
    // local stage J (JFIRST..3) is global stage s_first + (J - JFIRST); pairs x[g*2h + i], x[g*2h + i + h], h = 8 >> J;
    // twiddle index m + i of ntt.cpp:298-300 = 2^s + (t_hi << J) + g  (s_first >= JFIRST in every caller, so the index of
    // g = 0 is a multiple of 2^J as ref_load_tw_u64 needs)
    ref_stage_u64<4, J, VEC>(x, roots, precons, (1u << (s_first + J - JFIRST)) + (t_hi << J), q, twice);
    if constexpr (J < 3) ref_pass16_u64<VEC, JFIRST, J + 1>(x, roots, precons, s_first, t_hi, q, twice);
}

template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT) ref_u64_frame_kernel(const uint64_t *in, const uint64_t *in2,   // may alias each other and out
                                                              uint64_t *out, const uint64_t *__restrict__ roots,
                                                              const uint64_t *__restrict__ precons, uint64_t q, uint32_t logn,
                                                              uint32_t split) {
    // split = 0: one CTA per frame, logn <= 14.  split = 1: two CTAs per frame (logn = 15), CTA parity = which half.
    extern __shared__ uint64_t img[];
    constexpr bool VEC = AGX_U64_VECTW >= 2 || (AGX_U64_VECTW == 1 && NT >= 512);   // 16-byte twiddle pairs
    const uint32_t tid = threadIdx.x;
    const uint32_t frame = split ? blockIdx.x >> 1 : blockIdx.x, half_id = split ? blockIdx.x & 1u : 0u;
    // log2 of the coefficients this CTA owns: always log2(NT) + 5 (a CTA is N/32 threads, agx_api.cu ref_launch), so it is a
    // compile-time constant and with it every pass's bit position, stride and stage count: shared-memory accesses become
    // base + immediate and the pass loop unrolls into exactly the passes this size runs
    constexpr uint32_t lg = (NT == 512 ? 14u : NT == 256 ? 13u : NT == 128 ? 12u : NT == 64 ? 11u : 10u);
    constexpr uint32_t M = 1u << lg, vthreads = M >> 4;
    if (logn - split != lg) return;                                // launch contract (never taken)
    const uint64_t twice = q << 1;
    const size_t fbase = (size_t)frame << logn;
    const uint64_t *lo_src = in + fbase, *hi_src = in2 + fbase;    // low / high half of the frame (ntt.cpp:587-589)
    uint64_t x[16];

    // ---- pass 0: stages (split ..) on index bits [lg-4, lg), coefficients straight from global memory
    {
        constexpr uint32_t lo = lg - 4;
#pragma unroll 1
        for (uint32_t vt = tid; vt < vthreads; vt += NT) {
            if (split) {
                // stage 0 of the 32768-point frame, this CTA keeping only its half (ntt.cpp:331-369 with roots[1])
                const uint64_t W = __ldg(roots + 1), Wp = __ldg(precons + 1);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t i = (k << lo) + vt;
                    uint64_t a = lo_src[i], b = hi_src[M + i];
                    ref_bfly_u64(a, b, W, Wp, q, twice);
                    x[k] = half_id ? b : a;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t i = (k << lo) + vt;
                    x[k] = (k < 8 ? lo_src : hi_src)[i];
                }
            }
            // this CTA's sub-transform: stage s of it is global stage s + split, group offset half_id << s
            // (roots index = 2^(s+split) + (half_id << s) + local group)
            ref_pass16_u64<VEC, 0>(x, roots, precons, split, half_id, q, twice);
            uint64_t *s = img + vt + (vt >> 4);
            constexpr uint32_t stride = (1u << lo) + (1u << (lo - 4));
#pragma unroll
            for (int k = 0; k < 16; k++) s[k * stride] = x[k];
        }
    }
    __syncthreads();
    // ---- middle passes: index bits from lg-5 down to 4, four at a time (the last one may cover fewer)
    uint32_t s_done = 4;                                           // stages of the sub-transform finished so far
#pragma unroll
    for (int rem = (int)lg - 8; rem > 0; rem -= 4) {
        const uint32_t lo = rem >= 4 ? (uint32_t)rem : 4u;
        const uint32_t jfirst = rem >= 4 ? 0u : (uint32_t)(4 - rem);
        const uint32_t stride = (1u << lo) + (1u << (lo - 4));
#pragma unroll 1
        for (uint32_t vt = tid; vt < vthreads; vt += NT) {
            const uint32_t t_lo = vt & ((1u << lo) - 1), t_hi = vt >> lo;
            const uint32_t idx0 = (t_hi << (lo + 4)) + t_lo;
            uint64_t *s = img + idx0 + (idx0 >> 4);
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = s[k * stride];
            // twiddle index: 2^(s+split) + ((half_id << s) + group) with group = (t_hi << j) + g at local stage j
            const uint32_t sg = s_done + split;                    // global stage of the first active local stage
            const uint32_t th = t_hi + (half_id << (s_done - jfirst));   // the CTA's half as the top bit of the group index
            if (jfirst == 0) ref_pass16_u64<VEC, 0>(x, roots, precons, sg, th, q, twice);
            else if (jfirst == 1) ref_pass16_u64<VEC, 1>(x, roots, precons, sg, th, q, twice);
            else if (jfirst == 2) ref_pass16_u64<VEC, 2>(x, roots, precons, sg, th, q, twice);
            else ref_pass16_u64<VEC, 3>(x, roots, precons, sg, th, q, twice);
#pragma unroll
            for (int k = 0; k < 16; k++) s[k * stride] = x[k];
        }
        s_done += 4 - jfirst;
        __syncthreads();
    }
    // ---- final pass: bits 3..0 (16 consecutive coefficients), reduction to [0,q) (ntt.cpp:377-393)
#pragma unroll 1
    for (uint32_t vt = tid; vt < vthreads; vt += NT) {
        uint64_t *s = img + vt * 17;
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = s[k];
        ref_pass16_u64<VEC, 0>(x, roots, precons, s_done + split, vt + (half_id << s_done), q, twice);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            s[k] = ref_csub_u64(ref_csub_u64(x[k], twice), q);
        }
    }
    __syncthreads();
    uint64_t *dst = out + fbase + (size_t)half_id * M;
    for (uint32_t i = tid; i < M; i += NT) dst[i] = img[i + (i >> 4)];
}

// ------------------------------------------------------------------------------------- twiddle tables on device
// The step before the path: the reference fills its root / precon buffers on the host (main.cpp:46-55) and the loader
// broadcasts them (ntt.cpp:544-571).  Here one thread per (limb, direction, k) computes
//   w = base^bitrev(k) mod q   (base = psi forward, psi^-1 inverse; the table order ntt.cpp:298-300 consumes),
//   w' = floor(w * 2^32 / q)   (Shoup companion),
// and writes it to the natural-order table (introspection, generic kernel), to its kernel-order slot (tw_pos) and, for
// k < E, to the column pass's slot.  Inverse kernel-order entries 0 and 1 carry n^-1 (see ntt_inv_loop_kernel).
struct LimbGen {
    uint32_t q, psi, psi_inv, n_inv;
};

__device__ __forceinline__ uint32_t mulmod_dev(uint32_t a, uint32_t b, uint32_t q) {
    return (uint32_t)(((uint64_t)a * b) % q);
}

// Places entry k (value r = base^bitrev(k)) of one limb's table into every layout the kernels read:
// natural order, kernel order (tw_pos) and -- for k < E -- the column pass's slots; inverse tables get n^-1 folded in.
// le == 0: generic sizes, kernel order == natural order and no column tables.
struct TableSet {
    uint2 *nat, *tw, *twc;      // one limb, one direction
};

__device__ __forceinline__ uint2 shoup_pair(uint32_t w, uint32_t q) { return make_uint2(w, (uint32_t)(((uint64_t)w << 32) / q)); }

// entry `pos` of a kernel-order table; `vec4`: the slot is fetched by the kernels' 16-byte loads (see AGX_TW_INTERLEAVE)
__device__ __forceinline__ void store_tw(uint2 *tab, uint32_t pos, uint2 v, bool vec4) {
#if AGX_TW_INTERLEAVE
    if (vec4) {
        uint32_t *w = reinterpret_cast<uint32_t *>(tab) + (pos >> 1) * 4 + (pos & 1);
        w[0] = v.x;
        w[2] = v.y;
        return;
    }
#endif
    tab[pos] = v;
}

__device__ __forceinline__ void place_table_entry(const TableSet &t, uint32_t k, uint32_t r, bool inverse, uint32_t q,
                                                  uint32_t n_inv, uint32_t logn, int le) {
    const uint2 natural = shoup_pair(r, q);
    t.nat[k] = natural;
    uint32_t w = r;
    if (inverse && k == 0) w = n_inv;                                    // unused slot carries n^-1
    if (inverse && k == 1 && le) w = mulmod_dev(r, n_inv, q);            // last GS stage folds n^-1 (two-pass kernels)
    uint32_t pos = k;
    if (le && k >= (1u << le)) {
        const int s = 31 - __clz(k), lt = (int)logn - le;
        const uint32_t c = 1u << (s - lt), rr = k - (1u << s), tpp = 1u << lt;
        const uint32_t T = rr / c, kk = rr % c;
        pos = c == 1 ? (1u << s) + T : (1u << s) + ((kk >> 1) * tpp + T) * 2 + (kk & 1);   // = tw_pos<LOGN,LE>(s, T, kk)
    }
    store_tw(t.tw, pos, shoup_pair(w, q), le && pos >= (2u << ((int)logn - le)));   // stages read by 16-byte loads
    if (le && k >= 1 && k < (1u << le)) {   // column pass: local stage j, group gq sits where the row pass of thread 0 looks
        const int j = 31 - __clz(k), lt = (int)logn - le;
        const uint32_t gq = k - (1u << j), tpp = 1u << lt;
        const uint32_t cpos = j == 0 ? tpp : 2 * ((1u << (lt + j - 1)) + (gq >> 1) * tpp) + (gq & 1);
        uint2 cval = natural;
        if (inverse && j >= 1 && j <= kInvFold) cval = shoup_pair(mulmod_dev(r, n_inv, q), q);   // variant A carries n^-1
        store_tw(t.twc, cpos, cval, j >= 1);
        if (inverse && j >= 1) store_tw(t.twc, cpos + 2, natural, true);  // variant B: one 16-byte slot further
    }
}

// Tables from psi: one thread per (direction, limb, k).
__global__ void __launch_bounds__(256) gen_tables_kernel(uint2 *__restrict__ nat_fwd, uint2 *__restrict__ nat_inv,
                                                         uint2 *__restrict__ tw_fwd, uint2 *__restrict__ tw_inv,
                                                         uint2 *__restrict__ twc_fwd, uint2 *__restrict__ twc_inv,
                                                         const LimbGen *__restrict__ lg, uint32_t L, uint32_t logn, int le) {
    const uint32_t n = 1u << logn;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 2u * L * n) return;
    const uint32_t k = gid & (n - 1), limb = (gid >> logn) % L;
    const bool inverse = gid >= L * n;
    const LimbGen g = lg[limb];
    const uint32_t e = __brev(k) >> (32 - logn);
    uint32_t r = 1, b = inverse ? g.psi_inv : g.psi;
    for (uint32_t i = 0; i < logn; i++) {
        if ((e >> i) & 1) r = mulmod_dev(r, b, g.q);
        b = mulmod_dev(b, b, g.q);
    }
    const size_t base = (size_t)limb * n;
    const TableSet t{(inverse ? nat_inv : nat_fwd) + base, (inverse ? tw_inv : tw_fwd) + base,
                     le ? (inverse ? twc_inv : twc_fwd) + base : nullptr};
    place_table_entry(t, k, r, inverse, g.q, g.n_inv, logn, le);
}

// Tables handed in by the caller in the reference's order (ntt.cpp:298-300: entry m + i = twiddle of group i in the
// stage with m groups; include/kernel/ntt.h:35-41, main.cpp:36-37,46-55 hand such buffers to the loader): checked for
// consistency and re-laid-out for the kernels.  One thread per k.  A table is consistent when, with I = roots[1],
//   I^2 = -1,  roots[2k]^2 = roots[k],  roots[2k+1] = roots[2k] * I   (k >= 1),  every entry < q,
// which is equivalent to roots[k] = psi^bitrev(k) for the primitive 2n-th root psi = roots[n/2]; precons (optional)
// must be the Shoup companions floor(roots[k] * 2^32 / q).  Violations are counted in *bad.
__global__ void __launch_bounds__(256) relayout_tables_kernel(TableSet t, const uint32_t *__restrict__ roots,
                                                              const uint32_t *__restrict__ precons, bool inverse, uint32_t q,
                                                              uint32_t n_inv, uint32_t logn, int le, unsigned *bad) {
    const uint32_t n = 1u << logn;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t r = k ? roots[k] : 1u;
    bool ok = r < q && r != 0;
    if (ok && k >= 1) {
        if (precons && precons[k] != shoup_pair(r, q).y) ok = false;
        const uint32_t I = roots[1];
        if (k == 1 && mulmod_dev(I, I, q) != q - 1) ok = false;
        if (2 * k < n) {
            const uint32_t e = roots[2 * k], o = roots[2 * k + 1];
            if (e >= q || o >= q || mulmod_dev(e, e, q) != r || mulmod_dev(e, I % q, q) != o) ok = false;
        }
    }
    if (!ok) { atomicAdd(bad, 1u); return; }
    place_table_entry(t, k, r, inverse, q, n_inv, logn, le);
}

// ---------------------------------------------------------------------------------- synthetic data + checksum
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) fill_synthetic_kernel(uint32_t *__restrict__ data, size_t total, uint32_t logn,
                                                             uint32_t L, const LimbConst *__restrict__ lc, uint64_t seed,
                                                             uint64_t first_elem) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = lc[((first_elem + i) >> logn) % L].q;
        data[i] = (uint32_t)(splitmix64(seed + first_elem + i) % q);
    }
}

__global__ void __launch_bounds__(256) checksum_kernel(const uint32_t *__restrict__ data, size_t count, uint64_t first_index,
                                                       unsigned long long *__restrict__ sum) {
    uint64_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        acc += splitmix64((first_index + i) * 0xD6E8FEB86659FD93ULL + data[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, (unsigned long long)acc);
}

}  // namespace agx
