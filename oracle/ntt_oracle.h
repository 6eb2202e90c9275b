/*
 * ntt_oracle.h -- CPU oracle for the Agilex-NTT hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (agilex-ntt_b200/, include/) may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
 *
 * What it restates (reference = /root/reference, joekurina/Agilex-NTT):
 *   - orc_ref_fwd_u64      : the arithmetic of fwd_ntt_kernel, src/kernel/ntt.cpp:146-159 (stage/loop nest),
 *                            :292-300 (group / twiddle index), :331-332 (lazy x correction), :344-363 (Shoup
 *                            mulhi + Q), :368-369 (outputs), :377-393 (final reduction), :499-500 (t halves),
 *                            with the batch layout of ntt_input_kernel (:579-595) / ntt_output_kernel (:622-637).
 *   - orc_*_u32_*          : the same transform on the u32 / 30-bit-prime datapath BASELINE.json names
 *                            (SEAL-Embedded-style Barrett and Harvey/Shoup-lazy variants; SEAL-Embedded is named
 *                            by README.md:13 but no line of it is in the reference tree, so these follow the
 *                            published algorithm and are pinned by the textbook definition in oracle.py).
 *   - inverse / polymul    : no reference counterpart exists; pinned mathematically.
 *
 * Parity pinning: the reference holds no golden vectors or tests (include/test.h is 0 bytes).  The oracle is
 * pinned (tests/test_oracle.py) against (1) SURVEY.md App. A known-answer hashes, (2) the O(n^2) big-int
 * definition, and (3) the reference's own ntt.cpp compiled against a host SYCL stand-in (oracle/_ref, built by
 * oracle/Makefile from the sources where they lie) at the sizes ntt.h supports.
 */
#ifndef NTT_ORACLE_H
#define NTT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar helpers ---- */
uint64_t orc_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t orc_powmod(uint64_t a, uint64_t e, uint64_t q);
uint64_t orc_invmod(uint64_t a, uint64_t q); /* q prime */
int      orc_is_prime(uint64_t q);
uint32_t orc_bitrev(uint32_t x, uint32_t bits);
uint64_t orc_splitmix64(uint64_t x);

/* minimal primitive 2n-th root of unity mod q (SEAL convention); 0 if none */
uint64_t orc_min_psi(uint64_t n, uint64_t q);

/* ---- tables: roots[i] = psi^bitrev(i, log2 n), i in [0,n); ntt.cpp:298-300 uses roots[m+i] ---- */
void orc_tables_u64(uint32_t n, uint64_t q, uint64_t psi, int inverse, uint64_t *roots, uint64_t *precons);
void orc_tables_u32(uint32_t n, uint32_t q, uint32_t psi, int inverse, uint32_t *roots, uint32_t *precons);

/* ---- reference-shaped u64 forward transform (wraps mod 2^64 exactly like ntt.cpp) ---- */
void orc_ref_fwd_u64(uint32_t N, const uint64_t *in, const uint64_t *in2, uint64_t modulus,
                     const uint64_t *roots, const uint64_t *precons, uint32_t numFrames, uint64_t *out);

/* ---- u32 datapath, one polynomial, in place.  fwd: natural -> bit-reversed; inv: bit-reversed -> natural ---- */
void orc_fwd_u32_barrett(uint32_t n, uint32_t q, const uint32_t *roots, uint32_t *x);
void orc_inv_u32_barrett(uint32_t n, uint32_t q, const uint32_t *iroots, uint32_t *x);
void orc_fwd_u32_shoup(uint32_t n, uint32_t q, const uint32_t *roots, const uint32_t *precons, uint32_t *x);
void orc_inv_u32_shoup(uint32_t n, uint32_t q, const uint32_t *iroots, const uint32_t *iprecons, uint32_t *x);

/* ---- batched [B][L][n] drivers (OpenMP over transforms when threads > 1); variant: 0 barrett, 1 shoup ---- */
typedef struct {
    uint32_t n, nlimbs;
    uint32_t *q;        /* [L] */
    uint32_t *psi;      /* [L] */
    uint32_t *roots;    /* [L][n] */
    uint32_t *precons;  /* [L][n] */
    uint32_t *iroots;   /* [L][n] */
    uint32_t *iprecons; /* [L][n] */
} orc_plan;

orc_plan *orc_plan_create(uint32_t n, uint32_t nlimbs, const uint32_t *q);
void      orc_plan_destroy(orc_plan *p);
void orc_batch_fwd_u32(const orc_plan *p, uint32_t *data, size_t B, int variant, int threads);
void orc_batch_inv_u32(const orc_plan *p, uint32_t *data, size_t B, int variant, int threads);
void orc_batch_polymul_u32(const orc_plan *p, uint32_t *c, const uint32_t *a, const uint32_t *b, size_t B,
                           int threads);
void orc_batch_ref_fwd_u64(uint32_t N, uint64_t *data, uint64_t modulus, const uint64_t *roots,
                           const uint64_t *precons, size_t numFrames, int threads);

/* exact negacyclic schoolbook product mod (X^n + 1, q): the NTL ZZ_pX stand-in (README.md:9-10) */
void orc_polymul_schoolbook(uint32_t n, uint32_t q, const uint32_t *a, const uint32_t *b, uint32_t *c);

/* synthetic data, SURVEY.md s.8(d): element g of [B][L][n] = splitmix64(seed + g) mod q_limb */
void orc_fill_synthetic(uint32_t *data, size_t B, uint32_t nlimbs, uint32_t n, const uint32_t *q, uint64_t seed,
                        size_t first_poly);
/* order-sensitive 64-bit checksum of a u32 array (sum of splitmix64(index ^ value<<32)) */
uint64_t orc_checksum_u32(const uint32_t *data, size_t count, size_t first_index);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
