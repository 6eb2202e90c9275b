// agx_ntt_pers.cuh -- persistent forward / inverse kernels with asynchronous staging of the NEXT polynomial.
//
// Same arithmetic, passes, tables and shared-memory transposes as agx_ntt_kernels.cuh (the functions on the path are
// the reference's loader / compute / drain trio, ntt.cpp:508-607, 86-506, 610-640).  What changes is who waits for
// HBM.  In the one-CTA-per-polynomial kernels a CTA's first ~25 % of life is the wait for its own 16 KB (phase
// trace in profiles/r01_experiments.md), and with 8 CTAs of 128 registers per SM there is no spare warp to cover
// it.  Here a CTA stays resident and walks polynomials blockIdx.x, +gridDim.x, ...; the polynomial it will
// work on NEXT is copied into its shared-memory buffer by hardware while it computes the current one:
//
//   forward : one elected thread issues a 1-D TMA bulk copy (cp.async.bulk, completion on an mbarrier) of the whole
//             row into the transpose buffer, right after the current polynomial's transpose has been read back --
//             the buffer is idle from then until the next polynomial starts.  The column pass then reads its
//             strided coefficients from shared memory (LDS.32, conflict-free) instead of from global memory.
//             Results leave through a half-row staging buffer (two rounds), because the main buffer is busy
//             receiving: 16.6 KB + 9.2 KB of shared memory per CTA keeps 8 CTAs (n = 4096) on an SM.
//   inverse : needs its rows in the padded (conflict-free LDS.128) layout, which a linear bulk copy cannot
//             produce, so every thread issues 16-byte cp.async (LDGSTS) copies into the padded image -- issued
//             after the rows->columns transpose has been read back, waited for (cp.async.wait_group) at the top
//             of the next iteration.  Results leave straight from registers as coalesced STG.32, as before.
//
// Which polynomial comes next is decided by the hardware work queue (Blackwell cluster launch control): the grid
// still has one CTA per polynomial, but a running CTA asks the scheduler to CANCEL a not-yet-launched CTA and takes
// over its blockIdx (clusterlaunchcontrol.try_cancel, answer delivered asynchronously into shared memory on an
// mbarrier).  The request is issued at the top of an iteration and read half an iteration later, when the staging
// copy is issued.  A static stride (poly += resident CTAs) loses 15 % instead of gaining: the warps of an SM do not
// progress at equal rates (11.5k .. 20k clk per polynomial inside one SM; round-1 trace of the persistent build, not tracked -- the shipped
// kernels' phase trace is profiles/r01_phase_trace.txt), so
// equal shares finish far apart; with work stealing every CTA index is processed exactly once, whether the
// hardware launches it or a resident CTA steals it, so correctness does not depend on any request succeeding.
// AGX_PERS_SCHED=0 builds the static-stride variant (grid = resident CTAs, a multiple of L) for comparison.
#pragma once
#include "agx_ntt_kernels.cuh"

namespace agx {

// ------------------------------------------------------------------------------------- async-copy primitives
// global -> shared bulk copy by the TMA engine; bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *b) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to dst are ordered first
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src_gmem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldgsts16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(__cvta_generic_to_global(src_gmem)) : "memory");
}
__device__ __forceinline__ void ldgsts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ldgsts_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// polynomials further ahead than the one being staged are pulled into L2 (0 = off)
#ifndef AGX_PERS_L2_AHEAD
#define AGX_PERS_L2_AHEAD 0
#endif

// -------------------------------------------------------------------------------- half-row output staging (fwd)
// TPP rows of E/2 words, each padded by 16 bytes: STS.128 of eight consecutive rows and LDS.128 of one row's
// chunks are both conflict-free (pitch in chunks is odd); every address is a per-thread base + immediate.
template <int LOGN, int LE>
struct HalfGeo {
    using G = Geo<LOGN, LE>;
    static constexpr int HC = G::CPR / 2;          // 16-byte chunks per half row
    static constexpr int HP4 = HC + 1;             // pitch in chunks
    static constexpr int CHUNKS = G::TPP * HP4;
    static constexpr int ROWS_PER_STEP = G::TPP / HC;   // rows covered by one copy-out step of the CTA
    static constexpr int STEPS = G::TPP / ROWS_PER_STEP;
};

template <int LOGN, int LE>
__device__ __forceinline__ void sts_half_row(uint4 *ob, const uint32_t (&x)[1 << LE], uint32_t tid, int h) {
    using H = HalfGeo<LOGN, LE>;
    uint4 *r = ob + tid * H::HP4;
#pragma unroll
    for (int c = 0; c < H::HC; c++) {
        const int j = 4 * (h * H::HC + c);
        r[c] = make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    }
}

template <int LOGN, int LE>
__device__ __forceinline__ void half_to_global(const uint4 *ob, uint32_t *g, uint32_t tid, int h) {
    using G = Geo<LOGN, LE>;
    using H = HalfGeo<LOGN, LE>;
    const uint32_t row = tid / H::HC, cc = tid % H::HC;
    const uint4 *s = ob + row * H::HP4 + cc;
    uint4 *g4 = reinterpret_cast<uint4 *>(g) + row * G::CPR + h * H::HC + cc;
#pragma unroll
    for (int i = 0; i < H::STEPS; i++)
        __stcs(g4 + i * H::ROWS_PER_STEP * G::CPR, s[i * H::ROWS_PER_STEP * H::HP4]);
}

// padded-image cp.async of one polynomial (the asynchronous twin of global_to_smem)
template <int LOGN, int LE>
__device__ __forceinline__ void global_to_smem_async(uint4 *sm, const uint32_t *g, uint32_t tid) {
    using G = Geo<LOGN, LE>;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(g) + tid;
    uint4 *s = sm + (tid / G::CPR) * G::PITCH4 + (tid % G::CPR);
#pragma unroll
    for (int i = 0; i < G::N / 4 / G::TPP; i++) ldgsts16(s + i * (G::TPP / G::CPR) * G::PITCH4, g4 + i * G::TPP);
    ldgsts_commit();
}


#ifndef AGX_PERS_SCHED
#define AGX_PERS_SCHED 1
#endif

// ------------------------------------------------------------------------------------------ next-work queries
__device__ __forceinline__ void clc_try_cancel(uint4 *resp, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous answer has been read (generic proxy)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(
                     smem_u32(resp)), "r"(smem_u32(bar)) : "memory");
}
// returns true and the stolen blockIdx.x if the request cancelled a pending CTA
__device__ __forceinline__ bool clc_answer(const uint4 *resp, uint32_t &ctaid_x) {
    const uint4 r = *resp;
    uint32_t ok, x;
    asm volatile(
        "{\n"
        ".reg .b128 R;\n"
        ".reg .pred P;\n"
        ".reg .b32 y, z, w;\n"
        "mov.b128 R, {%2, %3};\n"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 P, R;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "mov.b32 %1, 0;\n"
        "@P clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%1, y, z, w}, R;\n"
        "}" : "=r"(ok), "=r"(x)
        : "l"((unsigned long long)r.x | ((unsigned long long)r.y << 32)), "l"((unsigned long long)r.z | ((unsigned long long)r.w << 32)));
    ctaid_x = x;
    return ok != 0;
}

// Per-CTA work cursor.  begin(): once, by every thread, before the first barrier.  request(): thread 0, top of an
// iteration (after a barrier that follows the previous answer()).  answer(): every thread, at the staging point.
struct WorkCursor {
    uint4 *resp;
    uint64_t *bar;
    uint32_t parity;
    bool more;
    __device__ __forceinline__ void begin(uint4 *r, uint64_t *b, uint32_t tid) {
        resp = r; bar = b; parity = 0; more = true;
#if AGX_PERS_SCHED
        if (tid == 0) mbar_init(bar, 1);
#endif
    }
    __device__ __forceinline__ void request(uint32_t tid) {
#if AGX_PERS_SCHED
        if (tid == 0 && more) clc_try_cancel(resp, bar);
#endif
    }
    __device__ __forceinline__ bool answer(uint32_t poly, uint32_t T, uint32_t &next) {
#if AGX_PERS_SCHED
        if (more) {
            mbar_wait(bar, parity);
            parity ^= 1;
            more = clc_answer(resp, next);       // after a refusal no further request may be made
        }
        return more;
#else
        next = poly + gridDim.x;
        return next < T;
#endif
    }
};

// per-limb state of a CTA; reloaded only when the batch has several limbs and the polynomial's limb changed
template <int N>
struct LimbState {
    LimbConst c;
    const uint2 *tw, *twc;
    uint32_t limb;
    __device__ __forceinline__ void load(const KParams &p, const uint2 *tw_all, const uint2 *twc_all, uint32_t poly) {
        limb = p.L == 1 ? 0 : poly % p.L;
        c = p.lc[limb];
        tw = tw_all + (size_t)limb * N;
        twc = twc_all + (size_t)limb * N;
    }
    __device__ __forceinline__ void update(const KParams &p, const uint2 *tw_all, const uint2 *twc_all, uint32_t poly) {
        if (p.L != 1 && poly % p.L != limb) load(p, tw_all, twc_all, poly);
    }
};

// ----------------------------------------------------------------------------------------------------- forward
template <int LOGN, int LE>
__global__ void __launch_bounds__(1 << (LOGN - LE), AGX_MINB(LOGN, LE))
ntt_fwd_pers_kernel(uint32_t *dst, const uint32_t *src, KParams p, uint32_t T) {
    using G = Geo<LOGN, LE>;
    using H = HalfGeo<LOGN, LE>;
    __shared__ __align__(128) uint4 sm[G::SMEM_CHUNKS];   // landing zone (linear image) / transpose buffer (padded image)
    __shared__ __align__(16) uint4 ob[H::CHUNKS];         // output staging, half a row per thread
    __shared__ __align__(16) uint4 clc_resp;
    __shared__ __align__(8) uint64_t mbar, clc_bar;
    const uint32_t tid = threadIdx.x;
    uint32_t poly = blockIdx.x;
    LimbState<G::N> ls;
    ls.load(p, p.tw_fwd, p.twc_fwd, poly);
    constexpr uint32_t BYTES = 4u << LOGN;
    WorkCursor wc;
    wc.begin(&clc_resp, &clc_bar, tid);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, BYTES);
        bulk_g2s(sm, src + (size_t)poly * G::N, BYTES, &mbar);
    }
    poly_sync<G::TPP>();
    uint32_t parity = 0;
    const uint32_t *swl = reinterpret_cast<const uint32_t *>(sm);
#if AGX_TRACE
    unsigned long long tr_gt0, tr_c0 = clock64(), tr_wait = 0, tr_out = 0, tr_n = 0, tr_t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_gt0));
#endif

    for (;;) {
        uint32_t next = 0;
        bool have_next = false;
        uint32_t x[G::E];
        wc.request(tid);
#if AGX_TRACE
        tr_t = clock64();
#endif
        mbar_wait(&mbar, parity);
        parity ^= 1;
#if AGX_TRACE
        tr_wait += clock64() - tr_t; tr_n++;
#endif
        const LimbConst c = ls.c;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            PassAddr a;
            if (pass == 0) {                             // column pass: x[k] = poly[tid + TPP*k], stages 0..LE-1
                a = pass_addr<LOGN, LE>(ls.twc, 0u);
#pragma unroll
                for (int k = 0; k < G::E; k++) x[k] = swl[tid + G::TPP * k];
            } else {                                     // row pass: x[j] = poly[E*tid + j], stages LE..logn-1
                a = pass_addr<LOGN, LE>(ls.tw, tid);
                lds_row<LOGN, LE>(sm, x, tid);
                poly_sync<G::TPP>();                     // the buffer is idle from here on: stage the next polynomial
                have_next = wc.answer(poly, T, next);
                if (tid == 0 && have_next) {
                    mbar_expect_tx(&mbar, BYTES);
                    bulk_g2s(sm, src + (size_t)next * G::N, BYTES, &mbar);
                }
            }
            if (G::LT == LE || pass == 0) ct_stage<LOGN, LE, 0>(x, a, c);
            ct_stages_from<LOGN, LE, 1>(x, a, c);
            if (pass == 0) {
                poly_sync<G::TPP>();                     // every thread has taken its columns out of the linear image
                sts_columns<LOGN, LE>(reinterpret_cast<uint32_t *>(sm), x, tid);
                poly_sync<G::TPP>();
            }
        }
#pragma unroll
        for (int j = 0; j < G::E; j++) x[j] = reduce4q(x[j], c);
        uint32_t *g = dst + (size_t)poly * G::N;
#if AGX_TRACE
        tr_t = clock64();
#endif
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (h) poly_sync<G::TPP>();                  // first half has been copied out
            sts_half_row<LOGN, LE>(ob, x, tid, h);
            poly_sync<G::TPP>();
            half_to_global<LOGN, LE>(ob, g, tid, h);
        }
#if AGX_TRACE
        tr_out += clock64() - tr_t;
        if (!have_next && tid == 0 && blockIdx.x < 8192) {
            unsigned long long gt1; uint32_t smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long *tr = p.trace + (size_t)blockIdx.x * 8;
            tr[0] = smid; tr[1] = tr_gt0; tr[2] = gt1; tr[3] = tr_n; tr[4] = clock64() - tr_c0; tr[5] = tr_wait; tr[6] = tr_out;
            tr[7] = threadIdx.x;
        }
#endif
        if (!have_next) break;
        poly = next;
        ls.update(p, p.tw_fwd, p.twc_fwd, poly);
        // `ob` and the work-queue answer are next written after further barriers of this loop; `sm` is owned by
        // the bulk copy until its mbarrier flips: no barrier needed here.
    }
}

// One-shot forward kernel whose input arrives by ONE TMA bulk copy into the (still empty) transpose buffer instead of
// 64 coalesced LDG.32 per thread: experiment AGX_FWD_TMA=1 (profiles/r01_experiments.md).
template <int LOGN, int LE, bool CL>
__global__ void __launch_bounds__(1 << (LOGN - LE), AGX_MINB(LOGN, LE))
ntt_fwd_tma_kernel(uint32_t *dst, const uint32_t *src, KParams p, uint32_t T) {
    using G = Geo<LOGN, LE>;
    __shared__ __align__(128) uint4 sm[G::SMEM_CHUNKS];
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t tid = threadIdx.x;
    const uint32_t poly = blockIdx.x;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_fwd + (size_t)limb * G::N;
    const uint2 *twc = p.twc_fwd + (size_t)limb * G::N;
    const uint32_t *gs = src + (size_t)poly * G::N;
    constexpr uint32_t BYTES = 4u << LOGN;
    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, BYTES);
        bulk_g2s(sm, gs, BYTES, &mbar);
    }
    prefetch_ahead<LOGN, G::TPP>(gs, poly, T, tid);
    poly_sync<G::TPP>();
    mbar_wait(&mbar, 0);
    uint32_t x[G::E];
    const uint32_t *swl = reinterpret_cast<const uint32_t *>(sm);
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        PassAddr a;
        if (pass == 0) {
            a = pass_addr<LOGN, LE>(twc, 0u);
#pragma unroll
            for (int k = 0; k < G::E; k++) x[k] = swl[tid + G::TPP * k];
        } else {
            a = pass_addr<LOGN, LE>(tw, tid);
            lds_row<LOGN, LE>(sm, x, tid);
        }
        if (G::LT == LE || pass == 0) ct_stage<LOGN, LE, 0>(x, a, c);
        ct_stages_from<LOGN, LE, 1>(x, a, c);
        if (pass == 0) {
            poly_sync<G::TPP>();                         // every thread has taken its columns out of the linear image
            sts_columns<LOGN, LE>(reinterpret_cast<uint32_t *>(sm), x, tid);
            poly_sync<G::TPP>();
        }
    }
#pragma unroll
    for (int j = 0; j < G::E; j++) x[j] = reduce4q(x[j], c);
    sts_row<LOGN, LE>(sm, x, tid);
    poly_sync<G::TPP>();
    smem_to_global<LOGN, LE>(sm, dst + (size_t)poly * G::N, tid);
}

// ----------------------------------------------------------------------------------------------------- inverse
template <int LOGN, int LE>
__global__ void __launch_bounds__(1 << (LOGN - LE), AGX_MINB(LOGN, LE))
ntt_inv_pers_kernel(uint32_t *__restrict__ data, KParams p, uint32_t T) {
    using G = Geo<LOGN, LE>;
    __shared__ __align__(16) uint4 sm[G::SMEM_CHUNKS];
    __shared__ __align__(16) uint4 clc_resp;
    __shared__ __align__(8) uint64_t clc_bar;
    const uint32_t tid = threadIdx.x;
    uint32_t poly = blockIdx.x;
    LimbState<G::N> ls;
    ls.load(p, p.tw_inv, p.twc_inv, poly);
    WorkCursor wc;
    wc.begin(&clc_resp, &clc_bar, tid);

    global_to_smem_async<LOGN, LE>(sm, data + (size_t)poly * G::N, tid);
    for (;;) {
        uint32_t next = 0;
        bool have_next = false;
        uint32_t x[G::E];
        ldgsts_wait_all();
        poly_sync<G::TPP>();
        wc.request(tid);
        const LimbConst c = ls.c;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            PassAddr a;
            if (pass == 0) {                             // row pass: stages logn-1 .. LE
                a = pass_addr<LOGN, LE>(ls.tw, tid);
                lds_row<LOGN, LE>(sm, x, tid);
            } else {                                     // column pass: stages LE-1 .. 1 (stage 0 below)
                a = pass_addr<LOGN, LE>(ls.twc, 0u, 1u);
                lds_columns<LOGN, LE>(reinterpret_cast<const uint32_t *>(sm), x, tid);
                poly_sync<G::TPP>();                     // transpose read back: the buffer can receive the next polynomial
                have_next = wc.answer(poly, T, next);
                if (have_next) global_to_smem_async<LOGN, LE>(sm, data + (size_t)next * G::N, tid);
            }
            gs_stages_down_to1<LOGN, LE, LE - 1>(x, a, c);
            if (pass == 0) {
                if (G::LT == LE) gs_stage<LOGN, LE, 0>(x, a, c);
                sts_row<LOGN, LE>(sm, x, tid);           // own row only
                poly_sync<G::TPP>();
            }
        }
        inv_last_stage<LOGN, LE>(x, ls.tw, ls.twc, c);
        uint32_t *g = data + (size_t)poly * G::N;
#pragma unroll
        for (int k = 0; k < G::E; k++) __stcs(g + tid + G::TPP * k, x[k]);
        if (!have_next) break;
        poly = next;
        ls.update(p, p.tw_inv, p.twc_inv, poly);
    }
}

}  // namespace agx
