// agx_arith.cuh -- modular butterflies for the u32 datapath (sm_100a).
//
// The arithmetic is the reference's Harvey lazy butterfly with a Shoup-precomputed twiddle
// (/root/reference/src/kernel/ntt.cpp:331-332 x-correction, :344-363 mulhi + Q, :368-369 outputs,
// :377-393 final reduction) at half the word width: operands are u32, q < 2^30, so [0,4q) fits a register and
// the 64x64->hi64 product ntt.cpp builds from four 32x32 partial products becomes one IMAD.HI.
//
// Instruction budget per butterfly (what the integer roofline in DESIGN.md counts): 3 IMAD-class
// (mul.hi, mul.lo, mad.lo) + 3 ALU-class (VIADDMNMX conditional subtract, 2 adds) = 6.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace agx {

// Per-limb constants, 32 bytes, read once per CTA.
struct LimbConst {
    uint32_t q;
    uint32_t twoq;
    uint32_t negq;     // 2^32 - q
    uint32_t neg2q;    // 2^32 - 2q
    uint32_t bar_mu;   // floor(2^(2k)/q) << (31-k), k = bit length of q  (pointwise Barrett)
    uint32_t bar_sh;   // k - 1
    uint32_t psi;
    uint32_t zero;     // always 0, but only known at run time: a third IADD3 operand that keeps two-input adds off
                       // the multiply pipe (ptxas otherwise turns half of them into IMAD.IADD, and IMAD is the
                       // binding pipe of this kernel -- see DESIGN.md "integer roofline")
};

// min(x - m, x) as unsigned: x - m if x >= m else x.  One VIADDMNMX.U32 on sm_90+ (DPX), `negm` = 2^32 - m.
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t negm) { return __viaddmin_u32(x, negm, x); }

// x * w mod q, lazily: result in [0,2q) for ANY 32-bit x, given w' = floor(w * 2^32 / q)  (ntt.cpp:344-363).
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint2 w, uint32_t negq) {
    return x * w.x + __umulhi(x, w.y) * negq;
}

// Cooley-Tukey (forward) butterfly: x,y in [0,4q) -> x + w*y, x - w*y, both in [0,4q).
__device__ __forceinline__ void ct_bfly(uint32_t &x, uint32_t &y, uint2 w, const LimbConst &c) {
    const uint32_t tx = csub(x, c.neg2q);          // ntt.cpp:331-332
    const uint32_t Q = shoup_mul(y, w, c.negq);    // ntt.cpp:344-363
    x = tx + Q + c.zero;                           // ntt.cpp:368  (three-input form: stays an IADD3)
    y = tx + c.twoq - Q;                           // ntt.cpp:369
}

// Gentleman-Sande (inverse) butterfly: u,v in [0,2q) -> u + v, (u - v)*w, both in [0,2q).
__device__ __forceinline__ void gs_bfly(uint32_t &u, uint32_t &v, uint2 w, const LimbConst &c) {
    const uint32_t s = u + v + c.zero;
    const uint32_t d = u + c.twoq - v;
    u = csub(s, c.neg2q);
    v = shoup_mul(d, w, c.negq);
}

// Last inverse stage with n^-1 folded in: wn = (n^-1, .), w1n = (iroot[1] * n^-1, .); outputs in [0,q).
__device__ __forceinline__ void gs_bfly_last(uint32_t &u, uint32_t &v, uint2 wn, uint2 w1n, const LimbConst &c) {
    const uint32_t s = u + v + c.zero;
    const uint32_t d = u + c.twoq - v;
    u = csub(shoup_mul(s, wn, c.negq), c.negq);
    v = csub(shoup_mul(d, w1n, c.negq), c.negq);
}

// Last inverse stage for a pair whose inputs already carry n^-1 (they are difference outputs of the stage before, whose
// column-pass twiddles are stored pre-multiplied by n^-1): the plain butterfly with the unscaled twiddle w1; no multiply
// on the sum.  u,v in [0,2q) -> outputs in [0,q).
__device__ __forceinline__ void gs_bfly_last_prescaled(uint32_t &u, uint32_t &v, uint2 w1, const LimbConst &c) {
    const uint32_t s = u + v + c.zero;
    const uint32_t d = u + c.twoq - v;
    u = csub(csub(s, c.neg2q), c.negq);
    v = csub(shoup_mul(d, w1, c.negq), c.negq);
}

// [0,4q) -> [0,q)  (ntt.cpp:377-393)
__device__ __forceinline__ uint32_t reduce4q(uint32_t v, const LimbConst &c) {
    return csub(csub(v, c.neg2q), c.negq);
}

// a*b mod q for a,b < 2^k (k = bit length of q): Barrett with a 32-bit quotient estimate; result in [0,3q).
__device__ __forceinline__ uint32_t barrett_mul_lazy(uint32_t a, uint32_t b, const LimbConst &c) {
    const uint32_t lo = a * b;
    const uint32_t hi = __umulhi(a, b);
    const uint32_t ph = __funnelshift_r(lo, hi, c.bar_sh);   // (a*b) >> (k-1), < 2^(k+1)
    const uint32_t est = __umulhi(ph, c.bar_mu);             // floor(ph * mu / 2^(k+1))
    return lo + est * c.negq;
}

}  // namespace agx
