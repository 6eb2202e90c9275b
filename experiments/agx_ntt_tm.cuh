// agx_ntt_tm.cuh -- n = 4096 forward kernel that keeps the coefficients in TENSOR MEMORY.
//
// EXPERIMENT (-DAGX_TMEM=1), bit-exact but slower than the shipped kernel (0.53 vs 0.49 ms; profiles/r01_experiments.md).
//
// Idea: the register-resident kernels (agx_ntt_kernels.cuh) need 64 coefficients + twiddles = 128 registers per thread,
// which caps an SM at 512 threads = 4 warps per scheduler.  Blackwell's tensor memory is as large as the register file
// (512 columns x 128 lanes x 32 bit per SM), unused by this workload, and tcgen05.ld/st.32x32b give every thread of a
// warp private columns of "its" lane -- a second register file with a ~12-cycle access.  So a thread parks its 64
// coefficients in 64 TMEM columns and works on 16 at a time.  What the measurement said: event-timed, the butterfly
// stream runs 13.2 butterflies/clk/SM at 4 warps per scheduler and 14.1 at 8 (the multiply pipe is 90 % busy), so
// doubling the resident warps is worth 7 %, less than the TMEM round trips (6 %) and the 15 % extra instructions cost.
//
//   pass over 64 coefficients x[0..64) (6 stages, stage j pairs x[k] with x[k + (32 >> j)]) =
//     part 1: stages 0,1 on the 16 radix-4 groups {kl + 16*kh, kh < 4}, four groups (16 registers) per batch;
//             batch b holds registers [kh][e] = x[16*kh + 4*b + e] and leaves them as four x4 stores at columns 16*kh + 4*b;
//     part 2: stages 2..5 on the four runs x[16*kh .. 16*kh+16), one x16 load each.
//
// With ~60 registers per thread 1024 threads (8 warps per scheduler) fit an SM.  A CTA is 128 threads = two
// polynomials (warps 0-1 and 2-3: a warp reaches only the TMEM lanes 32*(warp%4).., so four warps use all 128 lanes
// of the CTA's 64-column allocation; 8 CTAs x 64 columns = the whole TMEM).  Shared memory would not hold 16 full
// transpose images per SM, and it does not have to: with the data parked in TMEM the transpose runs in two rounds
// through a 9 KB buffer per polynomial.  In round r warp w writes its coefficient chunk w^r (32 image rows, its own 32
// columns) and reads back, from its own rows, the 32 words that warp w^r wrote -- the XOR makes the chunk a thread
// receives land exactly in the TMEM columns it has just freed (TMEM addresses are warp-uniform, hence whole warps).
// Results leave in two rounds of 32 full rows (512 contiguous bytes per store instruction).
//
// Arithmetic, twiddle tables (kernel order, tw_pos) and results are those of the register-resident kernels; the
// reference lines are the same (ntt.cpp:146-159 loop nest, :292-300 twiddle index, :331-369 butterfly, :377-393 final
// reduction).
#pragma once
#include "agx_ntt_kernels.cuh"

namespace agx {

// --------------------------------------------------------------------------------------------- TMEM primitives
#define AGX_R4 "{%0, %1, %2, %3}"
#define AGX_R16 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}"
#define AGX_OUT16(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), \
    "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), \
    "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])
#define AGX_IO16(r) "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), \
    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
#define AGX_IN16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])

__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 " AGX_R16 ", [%16];" : AGX_OUT16(r, 0) : "r"(taddr) : "memory");
}
template <int O>
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 " AGX_R4 ", [%4];"
                 : "=r"(r[O]), "=r"(r[O + 1]), "=r"(r[O + 2]), "=r"(r[O + 3]) : "r"(taddr) : "memory");
}
// registers written by tcgen05.ld are valid only after the wait: tie them to it so their uses cannot be hoisted above
__device__ __forceinline__ void tm_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : AGX_IO16(r) : : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], " AGX_R16 ";" : : AGX_IN16(r), "r"(taddr) : "memory");
}
template <int O>
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], " AGX_R4 ";"
                 : : "r"(r[O]), "r"(r[O + 1]), "r"(r[O + 2]), "r"(r[O + 3]), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kTmCols = 64;   // TMEM columns per CTA: 64 words for each of the 128 threads

__device__ __forceinline__ uint32_t tm_alloc(uint32_t *slot) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(slot)), "n"(kTmCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tm_free(uint32_t base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmCols) : "memory");
}

// barrier ids are immediates: with run-time ids ptxas reserves all 16 hardware barriers per CTA, which limits an SM to 4 CTAs
template <int ID, int COUNT> __device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void named_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
__device__ __forceinline__ void team_sync(uint32_t team) { if (team == 0) named_sync<1, 64>(); else named_sync<2, 64>(); }

// ------------------------------------------------------------------------------------ the two parts of a pass
// Twiddle of local stage J, group g (run-time g): same table slots as load_tw_generic (agx_ntt_kernels.cuh).
template <int LOGN, int LE, int J>
__device__ __forceinline__ const uint4 *tw_row4(const PassAddr &a, uint32_t h) {   // uint4 holding groups 2h, 2h+1 (J >= 1)
    constexpr int LT = LOGN - LE, TPP = 1 << LT;
    return a.b4 + ((1 << (LT + J - 1)) + h * TPP);
}

// part 1, forward: registers [kh][e] = x[16*kh + 4*b + e]; stage 0 pairs kh with kh+2 (one twiddle), stage 1 pairs
// kh = 0,1 (group 0) and kh = 2,3 (group 1)
__device__ __forceinline__ void ct_part1(uint32_t (&x)[16], uint2 w0, uint2 w10, uint2 w11, const LimbConst &c) {
#pragma unroll
    for (int e = 0; e < 4; e++) { ct_bfly(x[e], x[8 + e], w0, c); ct_bfly(x[4 + e], x[12 + e], w0, c); }
#pragma unroll
    for (int e = 0; e < 4; e++) { ct_bfly(x[e], x[4 + e], w10, c); ct_bfly(x[8 + e], x[12 + e], w11, c); }
}

// part 2, forward: registers x[i] = pass element 16*kh + i; local stage J = 2..5 has groups kh*2^(J-2) + gl
template <int LOGN, int LE>
__device__ __forceinline__ void ct_part2(uint32_t (&x)[16], const PassAddr &a, uint32_t kh, const LimbConst &c) {
    {   // stage 2: group kh, pairs (i, i+8)
        const uint2 w = ld_twiddle(reinterpret_cast<const uint2 *>(tw_row4<LOGN, LE, 2>(a, kh >> 1)) + (kh & 1));
#pragma unroll
        for (int i = 0; i < 8; i++) ct_bfly(x[i], x[i + 8], w, c);
    }
    {   // stage 3: groups 2kh, 2kh+1, pairs (i, i+4) in each run of 8
        const uint4 v = ld_twiddle(tw_row4<LOGN, LE, 3>(a, kh));
#pragma unroll
        for (int i = 0; i < 4; i++) { ct_bfly(x[i], x[i + 4], make_uint2(v.x, v.y), c); ct_bfly(x[8 + i], x[12 + i], make_uint2(v.z, v.w), c); }
    }
    {   // stage 4: groups 4kh .. 4kh+3
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 v = ld_twiddle(tw_row4<LOGN, LE, 4>(a, 2 * kh + h));
#pragma unroll
            for (int i = 0; i < 2; i++) {
                ct_bfly(x[8 * h + i], x[8 * h + i + 2], make_uint2(v.x, v.y), c);
                ct_bfly(x[8 * h + 4 + i], x[8 * h + 6 + i], make_uint2(v.z, v.w), c);
            }
        }
    }
    {   // stage 5: groups 8kh .. 8kh+7
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const uint4 v = ld_twiddle(tw_row4<LOGN, LE, 5>(a, 4 * kh + h));
            ct_bfly(x[4 * h], x[4 * h + 1], make_uint2(v.x, v.y), c);
            ct_bfly(x[4 * h + 2], x[4 * h + 3], make_uint2(v.z, v.w), c);
        }
    }
}

// ---------------------------------------------------------------------------------- transpose / output buffer
// Per polynomial team (64 threads): max(2 blocks of 32 rows x 32 words for a transpose round, 32 rows x 64 words for an
// output round), rows padded by 16 bytes.
struct TmBuf {
    static constexpr int TP = 36;                  // transpose round: row pitch in words (32 + 4)
    static constexpr int TB = 32 * TP;             // block of 32 rows
    static constexpr int OP4 = 17;                 // output round: row pitch in 16-byte chunks (16 + 1)
    static constexpr int WORDS = 2 * TB > 32 * OP4 * 4 ? 2 * TB : 32 * OP4 * 4;
};

// ----------------------------------------------------------------------------------------------------- forward
// grid = ceil(T / 2) CTAs of 128 threads; polynomial 2*blockIdx.x + (threadIdx.x >> 6).  An odd tail repeats the last
// polynomial in the second half of the CTA without storing it.  Named barriers 1 and 2 = the two teams.
template <int LOGN, int LE, bool CL>
__global__ void __launch_bounds__(128, 8)
ntt_fwd_tm_kernel(uint32_t *dst, const uint32_t *src, KParams p, uint32_t T) {
    static_assert(LOGN == 12 && LE == 6, "TMEM-resident kernels are written for n = 4096 (64 threads x 64 coefficients)");
    using G = Geo<LOGN, LE>;
    __shared__ __align__(16) uint32_t buf_all[2][TmBuf::WORDS];
    __shared__ uint32_t tm_slot;
    const uint32_t team = threadIdx.x >> 6, tid = threadIdx.x & 63;
    uint32_t poly = 2 * blockIdx.x + team;
    const bool store = poly < T;
    if (!store) poly = T - 1;
    const uint32_t limb = p.L == 1 ? 0 : poly % p.L;
    AGX_LIMB_CONSTS(CL, p, limb);
    const uint2 *tw = p.tw_fwd + (size_t)limb * G::N;
    const uint2 *twc = p.twc_fwd + (size_t)limb * G::N;
    const uint32_t tm_base = tm_alloc(&tm_slot);
    const uint32_t taddr = tm_base + (((threadIdx.x >> 5) & 3) << 21);              // lane 32 * (warp % 4)
    uint32_t *buf = buf_all[team];
    uint32_t x[16];

    {   // ---- global -> TMEM: x[k] = poly[tid + 64k] at column k; two register sets keep 32 loads in flight
        const uint32_t *gs = src + (size_t)poly * G::N + tid;
        uint32_t y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = ld_stream(gs + G::TPP * i);
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = ld_stream(gs + G::TPP * (16 + i));
        prefetch_ahead<LOGN, G::TPP>(gs - tid, poly, T, tid);
        tm_st16(taddr, x);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = ld_stream(gs + G::TPP * (32 + i));
        tm_st16(taddr + 16, y);
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = ld_stream(gs + G::TPP * (48 + i));
        tm_st16(taddr + 32, x);
        tm_st16(taddr + 48, y);
        tm_st_wait();
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const PassAddr a = pass == 0 ? pass_addr<LOGN, LE>(twc, 0u) : pass_addr<LOGN, LE>(tw, tid);
        {   // ---- part 1: stages 0,1 on batches [kh][e] = x[16*kh + 4*b + e]
            const uint2 w0 = ld_twiddle(a.b2);
            const uint4 w1 = ld_twiddle(tw_row4<LOGN, LE, 1>(a, 0));
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                tm_ld4<0>(taddr + 4 * b, x); tm_ld4<4>(taddr + 16 + 4 * b, x); tm_ld4<8>(taddr + 32 + 4 * b, x); tm_ld4<12>(taddr + 48 + 4 * b, x);
                tm_ld_wait(x);
                ct_part1(x, w0, make_uint2(w1.x, w1.y), make_uint2(w1.z, w1.w), c);
                tm_st4<0>(taddr + 4 * b, x); tm_st4<4>(taddr + 16 + 4 * b, x); tm_st4<8>(taddr + 32 + 4 * b, x); tm_st4<12>(taddr + 48 + 4 * b, x);
            }
            tm_st_wait();
        }
        // ---- part 2: stages 2..5 on the runs x[16*kh .. 16*kh+16)
#pragma unroll
        for (int kh = 0; kh < 4; kh++) {
            tm_ld16(taddr + 16 * kh, x);
            tm_ld_wait(x);
            ct_part2<LOGN, LE>(x, a, kh, c);
            tm_st16(taddr + 16 * kh, x);
        }
        tm_st_wait();
        if (pass == 0) {
            // ---- transpose: column layout (thread t holds poly[t + 64k]) -> row layout (thread T holds poly[64T + j])
            const uint32_t wrp = tid >> 5, lane = tid & 31;
            uint32_t y[16];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint32_t ch = wrp ^ r;                           // 32-coefficient chunk written, and received, this round
                tm_ld16(taddr + 32 * ch, x);
                tm_ld16(taddr + 32 * ch + 16, y);
                tm_ld_wait(x);
                tm_ld_wait(y);
                if (r) team_sync(team);                                // the previous round has been read
                uint32_t *wr = buf + ch * TmBuf::TB + lane;            // rows k = 32*ch + i, this thread's column
#pragma unroll
                for (int i = 0; i < 16; i++) { wr[i * TmBuf::TP] = x[i]; wr[(16 + i) * TmBuf::TP] = y[i]; }
                team_sync(team);
                const uint4 *rd = reinterpret_cast<const uint4 *>(buf + wrp * TmBuf::TB + lane * TmBuf::TP);   // own row, 32 words
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
                    const uint4 v = rd[cc], u = rd[4 + cc];
                    x[4 * cc] = v.x; x[4 * cc + 1] = v.y; x[4 * cc + 2] = v.z; x[4 * cc + 3] = v.w;
                    y[4 * cc] = u.x; y[4 * cc + 1] = u.y; y[4 * cc + 2] = u.z; y[4 * cc + 3] = u.w;
                }
                tm_st16(taddr + 32 * ch, x);                           // words 32*ch .. 32*ch+31 of the row: their natural columns
                tm_st16(taddr + 32 * ch + 16, y);
            }
            tm_st_wait();
        }
    }
    // ---- results: two rounds of 32 full rows through the buffer, coalesced 16-byte stores
    {
        const uint32_t wrp = tid >> 5, lane = tid & 31;
        uint4 *b4 = reinterpret_cast<uint4 *>(buf);
        uint4 *g4 = reinterpret_cast<uint4 *>(dst + (size_t)poly * G::N);
#pragma unroll 1
        for (int r = 0; r < 2; r++) {
            team_sync(team);                                           // transpose reads / previous round's copy-out are done
            if (wrp == (uint32_t)r) {                                  // whole warp: tcgen05.ld is warp-collective
                uint4 *row = b4 + lane * TmBuf::OP4;
#pragma unroll 1
                for (int q = 0; q < 4; q++) {
                    tm_ld16(taddr + 16 * q, x);
                    tm_ld_wait(x);
#pragma unroll
                    for (int i = 0; i < 16; i++) x[i] = reduce4q(x[i], c);
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) row[4 * q + cc] = make_uint4(x[4 * cc], x[4 * cc + 1], x[4 * cc + 2], x[4 * cc + 3]);
                }
            }
            team_sync(team);
            if (store) {
#pragma unroll
                for (int i = 0; i < 8; i++) {                          // chunk i*64 + tid of the round's 512
                    const uint32_t idx = i * 64 + tid;
                    st_stream(g4 + r * 512 + idx, b4[(idx >> 4) * TmBuf::OP4 + (idx & 15)]);
                }
            }
        }
    }
    tm_free(tm_base);
}

}  // namespace agx
