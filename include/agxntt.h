/*
 * agxntt.h -- C ABI of the B200-native Agilex-NTT hot path (libagxntt.so).
 *
 * Drop-in boundary for joekurina/Agilex-NTT's kernel API (reference = /root/reference):
 *
 *   reference interface                                        replaced by
 *   ---------------------------------------------------------  -------------------------------------------
 *   ntt_input_kernel(in, in2, modulus, twiddles, precons,      agx_ref_input()      include/kernel/ntt.h:35-41,
 *                    numFrames, q)                                                   src/kernel/ntt.cpp:508-607
 *   fwd_ntt_kernel<id>(q)                                      agx_ref_fwd()        include/kernel/ntt.h:32-33,
 *                                                                                    src/kernel/ntt.cpp:86-506
 *   ntt_output_kernel(out, numFrames, q)                       agx_ref_output()     include/kernel/ntt.h:43-45,
 *                                                                                    src/kernel/ntt.cpp:610-640
 *   sycl::queue::wait()   (src/main.cpp:74)                    agx_wait()
 *   compile-time FPGA_NTT_SIZE / modulus buffer                agx_parms + agx_create()   (the reference has no
 *   (ntt.h:7-23, main.cpp:34)                                  parameter struct)
 *   caller-filled twiddle / precon buffers                     agx_create_tables() (caller's psi), agx_set_tables()
 *   (ntt.h:38-39, main.cpp:36-37,46-55, ntt.cpp:122-141)       (caller's tables in the reference's order)
 *
 * and the batched u32 entry points BASELINE.json's north_star adds (no reference counterpart):
 *   agx_ntt_fwd / agx_ntt_inv / agx_polymul on device pointers, *_host variants on host pointers.
 *
 * Conventions
 *   - Plain C: pointers and sizes only.  Every function returns 0 (AGX_OK), a negative AGX_E_* code, or a
 *     positive cudaError_t value.  Nothing throws or aborts across this boundary.  There is NO CPU fallback:
 *     without a CUDA device agx_create() fails with the cudaError_t.
 *   - Transform definition (matches ntt.cpp): forward is negacyclic Cooley-Tukey, natural
 *     order in, BIT-REVERSED order out, outputs fully reduced to [0,q); inverse is Gentleman-Sande,
 *     bit-reversed in, natural out, n^-1 folded in.  psi = the minimal primitive 2n-th root of unity mod q unless the
 *     caller supplies its own root (agx_create_tables) or its own tables (agx_set_tables).
 *   - Batched layout: uint32_t data[B][L][n] (B polynomials x L RNS limbs), row-major, in place.
 *     Inputs must be < 2q for the u32 entry points (uniform-mod-q data is < q).  Device pointers must be 16-byte
 *     aligned (anything from cudaMalloc is).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer calls only
 *     enqueue; completion follows normal stream semantics.  *_host calls return after the results are in
 *     host memory.
 *   - A context is bound to one device and is not thread-safe; use one context per host thread / GPU.  Every call
 *     makes the context's device current for its own duration and restores the caller's current device on return.
 */
#ifndef AGXNTT_H
#define AGXNTT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_OK 0
#define AGX_E_INVALID (-1)     /* bad argument (null pointer, n not a supported power of two, bad prime ...) */
#define AGX_E_UNSUPPORTED (-2) /* valid request this build cannot serve */
#define AGX_E_NOMEM (-3)       /* host allocation failed */
#define AGX_E_STATE (-4)       /* reference-shaped calls used out of protocol */

typedef struct agx_ctx agx_ctx;

/* Run-time replacement for the reference's compile-time configuration (FPGA_NTT_SIZE ntt.h:7-23, one modulus
 * buffer main.cpp:34).  n = 2^logn in [8, 32768]; every q[i] prime, < 2^30, q = 1 (mod 2n).  The tuned
 * register-resident kernels cover n in {1024, 2048, 4096}; other sizes run a generic shared-memory kernel. */
typedef struct {
    uint32_t n;
    uint32_t logn;
    uint32_t nlimbs;
    const uint32_t *q; /* [nlimbs] */
} agx_parms;

/* parms == NULL creates a table-less context usable only with the agx_ref_* calls. */
int agx_create(agx_ctx **out, const agx_parms *parms, int device);
int agx_destroy(agx_ctx *ctx);

/* The reference's contract is "the caller hands in the roots" (include/kernel/ntt.h:38-39; main.cpp:36-37,46-55 fill
 * the buffers, ntt.cpp:122-141 receives them).  Two ways to do that on the u32 entry points:
 *
 * agx_create_tables: like agx_create, with the caller's primitive 2n-th root per limb, psi[nlimbs] (NULL = the
 *   minimal root, i.e. agx_create).  AGX_E_INVALID unless psi[i]^n = -1 (mod q[i]).  The library derives the forward
 *   and inverse tables from it on the device.
 *
 * agx_set_tables: replaces one limb's forward (inverse == 0) or inverse (inverse != 0) table by the caller's, given in
 *   the reference's order (ntt.cpp:298-300: entry m + i is the twiddle of group i in the stage with m groups, i.e.
 *   roots[k] = psi^bitrev(k) resp. psi^-bitrev(k); entry 0 is unused) as HOST arrays of n words.  precons[k] =
 *   floor(roots[k] * 2^32 / q) -- the 32-bit counterpart of the reference's barrettTwiddleFactors buffer -- or NULL to
 *   have them computed.  The table is checked on the device (roots[1]^2 = -1, roots[2k]^2 = roots[k], roots[2k+1] =
 *   roots[2k] * roots[1], entries < q, precons exact) and re-laid-out for the kernels (the inverse one with n^-1
 *   folded in); AGX_E_INVALID, with the previous table still in place, when it is not consistent.  A caller that
 *   replaces the forward table replaces the inverse one too (the library does not derive one from the other). */
int agx_create_tables(agx_ctx **out, const agx_parms *parms, const uint32_t *psi, int device);
int agx_set_tables(agx_ctx *ctx, uint32_t limb, int inverse, const uint32_t *roots, const uint32_t *precons);

/* Introspection (parity tests compare these with the oracle's tables). roots/precons in the reference's table
 * order, ntt.cpp:298-300: entry k = psi^bitrev(k) and floor(entry * 2^32 / q); inverse != 0 -> psi^-1. */
int agx_get_psi(const agx_ctx *ctx, uint32_t limb, uint32_t *psi);
int agx_get_tables(const agx_ctx *ctx, uint32_t limb, int inverse, uint32_t *roots, uint32_t *precons);

/* ---- batched u32 transforms on DEVICE pointers, in place, asynchronous on `stream` ---- */
int agx_ntt_fwd(agx_ctx *ctx, uint32_t *d_data, size_t B, void *stream);
int agx_ntt_inv(agx_ctx *ctx, uint32_t *d_data, size_t B, void *stream);
/* c = a * b mod (X^n + 1, q_limb) = INTT(NTT(a) .* NTT(b)).  c may alias a and/or b, and a may equal b (squaring), at
 * every size.  n = 1024: forward(a), forward(b), pointwise product and inverse run in ONE launch (3 streams of HBM
 * traffic); n = 2048 and 4096: three launches (c = NTT(a); c = NTT(b) .* c; c = INTT(c) -- 7 streams, no scratch
 * memory; the product is bound by the integer-multiply pipe, not by that traffic: DESIGN.md s.4); other sizes: the
 * transforms of that size (n >= 8192: shared-memory-resident register passes; n <= 512: the generic kernel) around a
 * pointwise launch, with a stream-ordered scratch buffer unless a == b. */
int agx_polymul(agx_ctx *ctx, uint32_t *d_c, const uint32_t *d_a, const uint32_t *d_b, size_t B, void *stream);
/* The product with one operand already in evaluation form -- the shape of an RLWE encryption or key switch, where the
 * key is transformed once and multiplied many times (no reference counterpart):
 * c = INTT(NTT(a) .* b_hat), b_hat = agx_ntt_fwd(b) (reduced, bit-reversed order as the forward transform leaves it).
 * Two launches at n = 1024, 2048, 4096 (c = NTT(a) .* b_hat inside the forward kernel; c = INTT(c): 5 streams of HBM
 * traffic, two transforms' worth of multiplies instead of agx_polymul's three); the generic transforms around a
 * pointwise launch at other sizes.  c may alias a and/or b_hat (aliasing b_hat at the generic sizes costs a scratch
 * copy).  Equals agx_polymul(c, a, b) bit for bit. */
int agx_polymul_by_spectrum(agx_ctx *ctx, uint32_t *d_c, const uint32_t *d_a, const uint32_t *d_b_hat, size_t B,
                            void *stream);

/* ---- limb-wise element-wise arithmetic on [B][L][n] DEVICE data (the operations callers run around the transforms,
 * e.g. spectrum multiply-accumulate between agx_ntt_fwd and agx_ntt_inv; no reference
 * counterpart).  Operands must be reduced (< q_limb); results are in [0, q_limb).  d_c may alias d_a and/or d_b. ---- */
#define AGX_EW_ADD 0 /* c = a + b     */
#define AGX_EW_SUB 1 /* c = a - b     */
#define AGX_EW_MUL 2 /* c = a * b     */
#define AGX_EW_MAC 3 /* c = c + a * b */
int agx_elementwise(agx_ctx *ctx, int op, uint32_t *d_c, const uint32_t *d_a, const uint32_t *d_b, size_t B,
                    void *stream);

/* ---- coefficient-order adapter: permutes every n-coefficient row of [B][L][n] DEVICE data by bit reversal, in place
 * (an involution).  The transforms keep the reference's orders (ntt.cpp:292-300: natural in, bit-reversed out; the
 * inverse takes bit-reversed input); callers holding natural-order spectra convert with this call.  No reference
 * counterpart. ---- */
int agx_bitrev(agx_ctx *ctx, uint32_t *d_data, size_t B, void *stream);

/* ---- the same on HOST pointers: chunked H2D / kernel / D2H pipeline over the context's own streams.
 * Pinned memory (agx_host_alloc, cudaHostAlloc, cudaHostRegister) is copied directly; pageable memory is staged
 * through internal pinned buffers.  h_out may equal h_in. ---- */
int agx_ntt_fwd_host(agx_ctx *ctx, const uint32_t *h_in, uint32_t *h_out, size_t B);
int agx_ntt_inv_host(agx_ctx *ctx, const uint32_t *h_in, uint32_t *h_out, size_t B);
int agx_polymul_host(agx_ctx *ctx, uint32_t *h_c, const uint32_t *h_a, const uint32_t *h_b, size_t B);
int agx_polymul_by_spectrum_host(agx_ctx *ctx, uint32_t *h_c, const uint32_t *h_a, const uint32_t *h_b_hat, size_t B);
int agx_host_alloc(void **p, size_t bytes);
int agx_host_free(void *p);

/* ---- synthetic inputs and checksums on the device ----
 * element g of [B][L][n], counted from polynomial `first_poly`, = splitmix64(seed + g) mod q_limb.
 * checksum = sum_i splitmix64((first_index + i) * 0xD6E8FEB86659FD93 + data[i]) mod 2^64 (shard sums add up). */
int agx_fill_synthetic(agx_ctx *ctx, uint32_t *d_data, size_t B, uint64_t seed, size_t first_poly, void *stream);
int agx_checksum(agx_ctx *ctx, const uint32_t *d_data, size_t count, size_t first_index, uint64_t *h_sum,
                 void *stream);

/* ---- reference-shaped u64 forward pipeline (host pointers; mirrors main.cpp:60-74) ----
 * N in {2^2 .. 2^15} (the reference builds 32, 1024, 8192, 16384, 32768: ntt.h:11-20).  Arithmetic is the
 * reference's u64 Harvey butterfly including wrap-around mod 2^64 for tables that are not Shoup pairs -- and for
 * moduli of 62 bits and more, whose lazy range [0,4q) no longer fits 64 bits (the reference wraps identically; parity
 * is pinned on 50-, 60-, 62- and 63-bit primes against the reference's own code).
 * The three calls only record/enqueue, in any order (the reference's three kernels run concurrently); the work
 * runs once all three were made; agx_wait() blocks until `out` is filled (= q.wait()).  `in`, `in2` hold
 * numFrames*N words (frame b: low half from in[b*N..], high half from in2[b*N + N/2..], ntt.cpp:582-591);
 * twiddles/precon_twiddles hold N words, modulus 1 word.  Host buffers must stay valid until agx_wait(). */
int agx_ref_input(agx_ctx *ctx, uint32_t N, const uint64_t *in, const uint64_t *in2, const uint64_t *modulus,
                  const uint64_t *twiddles, const uint64_t *precon_twiddles, uint32_t numFrames);
int agx_ref_fwd(agx_ctx *ctx, uint32_t compute_unit_id);
int agx_ref_output(agx_ctx *ctx, uint64_t *out, int32_t numFrames);
int agx_wait(agx_ctx *ctx);

/* The same transform on DEVICE pointers, asynchronous on `stream` (what the three calls above run per chunk; exposes
 * the kernels' rate without PCIe): frame b reads d_in[b*N .. b*N + N/2) and d_in2[b*N + N/2 .. (b+1)*N)
 * (ntt.cpp:582-591) and writes d_out[b*N .. (b+1)*N) (ntt.cpp:626-633).  d_twiddles / d_precon_twiddles: N words each
 * on the device.  All five device pointers must be 16-byte aligned (AGX_E_INVALID otherwise; cudaMalloc'd buffers are).
 * d_out may be the same buffer as d_in AND d_in2 (in place) but must not overlap just one of them.
 * Any 64-bit modulus: the arithmetic is the reference's, mod 2^64 (ntt.cpp:147-148, 331-369), so results are the
 * reference's for every modulus it accepts; they are the NTT when q < 2^62 (lazy range [0,4q) inside 64 bits). */
int agx_ref_fwd_dev(agx_ctx *ctx, uint32_t N, const uint64_t *d_in, const uint64_t *d_in2, uint64_t *d_out,
                    uint64_t modulus, const uint64_t *d_twiddles, const uint64_t *d_precon_twiddles, uint32_t numFrames,
                    void *stream);

/* ---- diagnostics ---- */
const char *agx_error_string(int code);
/* kernels this library launched on ctx since creation (bench.py's gpu_launches claim) */
int agx_launch_count(const agx_ctx *ctx, uint64_t *count);
/* name of the kernel variant serving (n): "ntt2p<LOGN,LE>" or "generic" */
int agx_variant(const agx_ctx *ctx, char *buf, size_t buflen);
/* The measured integer roofline of this GPU: the butterfly's instruction stream alone (no memory operations), one CTA
 * of `threads_per_sm` (128..1024, multiple of 128) threads per SM, timed in SM clocks.  kind 0: the u32 Harvey/Shoup
 * butterfly of the batched kernels; kind 1: the u64 butterfly of the reference-shaped path (ntt.cpp:331-369); kind 2 / 3:
 * kind 0 with one / two extra non-multiply instructions per butterfly (the issue-slot load of a real kernel); kinds 4-6:
 * further variants of the stream used by profiles/diag_issue_pressure.py (see csrc/agx_diag.cuh).
 * *per_clk_per_sm = butterflies retired per clock per SM; *sm_mhz (may be NULL) = the clock the run implied. */
int agx_measure_butterfly_peak(agx_ctx *ctx, int kind, int threads_per_sm, double *per_clk_per_sm, double *sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* AGXNTT_H */
