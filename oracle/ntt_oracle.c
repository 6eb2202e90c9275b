/*
 * ntt_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see ntt_oracle.h for the rules and the pinning story).
 *
 * Every function that follows reference code cites the file:line it restates.  Reference root: /root/reference.
 */
#include "ntt_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------------ scalars */

uint64_t orc_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

uint64_t orc_powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = orc_mulmod(r, a, q);
        a = orc_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

uint64_t orc_invmod(uint64_t a, uint64_t q) { return orc_powmod(a, q - 2, q); }

int orc_is_prime(uint64_t q) {
    if (q < 2) return 0;
    static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (size_t i = 0; i < sizeof small / sizeof *small; i++) {
        if (q == small[i]) return 1;
        if (q % small[i] == 0) return 0;
    }
    uint64_t d = q - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (size_t i = 0; i < sizeof small / sizeof *small; i++) { /* deterministic for 64-bit */
        uint64_t x = orc_powmod(small[i], d, q);
        if (x == 1 || x == q - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) {
            x = orc_mulmod(x, x, q);
            if (x == q - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

uint32_t orc_bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

uint64_t orc_splitmix64(uint64_t x) { /* SURVEY.md s.8(d) */
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static uint32_t ilog2u(uint32_t n) { uint32_t l = 0; while ((1u << l) < n) l++; return l; }

/* Minimal primitive 2n-th root of unity: psi^n = -1 (n a power of two makes that sufficient).  All primitive
 * 2n-th roots are the odd powers of any one of them; take the smallest (SEAL convention). */
uint64_t orc_min_psi(uint64_t n, uint64_t q) {
    if (n == 0 || (n & (n - 1)) || (q - 1) % (2 * n) != 0) return 0;
    uint64_t g = 0;
    for (uint64_t x = 2; x < q; x++) {
        uint64_t c = orc_powmod(x, (q - 1) / (2 * n), q);
        if (orc_powmod(c, n, q) == q - 1) { g = c; break; }
    }
    if (!g) return 0;
    uint64_t g2 = orc_mulmod(g, g, q), cur = g, best = g;
    for (uint64_t k = 1; k < 2 * n; k += 2) {
        if (cur < best) best = cur;
        cur = orc_mulmod(cur, g2, q);
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------- tables */
/* Table order consumed by ntt.cpp:298-300: stage with m groups reads roots[m + i] => roots[k] = psi^bitrev(k).
 * precons[k] = floor(roots[k] * 2^W / q): the Shoup companion ntt.cpp:344-363 multiplies y by (named
 * "barrettTwiddleFactors" there, ntt.h:39).  inverse != 0 builds the same table for psi^-1. */
void orc_tables_u64(uint32_t n, uint64_t q, uint64_t psi, int inverse, uint64_t *roots, uint64_t *precons) {
    uint32_t logn = ilog2u(n);
    uint64_t base = inverse ? orc_invmod(psi, q) : psi;
    uint64_t *pw = (uint64_t *)malloc(sizeof(uint64_t) * n);
    pw[0] = 1;
    for (uint32_t i = 1; i < n; i++) pw[i] = orc_mulmod(pw[i - 1], base, q);
    for (uint32_t k = 0; k < n; k++) {
        uint64_t w = pw[orc_bitrev(k, logn)];
        roots[k] = w;
        if (precons) precons[k] = (uint64_t)((((u128)w) << 64) / q);
    }
    free(pw);
}

void orc_tables_u32(uint32_t n, uint32_t q, uint32_t psi, int inverse, uint32_t *roots, uint32_t *precons) {
    uint32_t logn = ilog2u(n);
    uint64_t base = inverse ? orc_invmod(psi, q) : psi;
    uint64_t *pw = (uint64_t *)malloc(sizeof(uint64_t) * n);
    pw[0] = 1;
    for (uint32_t i = 1; i < n; i++) pw[i] = orc_mulmod(pw[i - 1], base, q);
    for (uint32_t k = 0; k < n; k++) {
        uint64_t w = pw[orc_bitrev(k, logn)];
        roots[k] = (uint32_t)w;
        if (precons) precons[k] = (uint32_t)((w << 32) / q);
    }
    free(pw);
}

/* -------------------------------------------------------------------- reference-shaped u64 forward (C1) */

/* ntt.cpp:344-362: 64x64 -> high 64 built from four 32x32 partial products (macros LOW/HIGH, ntt.cpp:26-30).
 * Restated with the same partial-product structure so wrap-around behaviour is identical for ANY operands. */
static inline uint64_t ref_mulhi64(uint64_t a, uint64_t b) {
    uint64_t a0 = a & 0xFFFFFFFFu, a1 = a >> 32, b0 = b & 0xFFFFFFFFu, b1 = b >> 32;
    uint64_t p00 = a0 * b0, p01 = a0 * b1, p10 = a1 * b0, p11 = a1 * b1;
    uint64_t mid = (p00 >> 32) + (p10 & 0xFFFFFFFFu) + (p01 & 0xFFFFFFFFu);
    return p11 + (p10 >> 32) + (p01 >> 32) + (mid >> 32);
}

static void ref_fwd_one_u64(uint32_t N, uint64_t *X, uint64_t q, const uint64_t *roots, const uint64_t *precons) {
    const uint64_t twice = q << 1;                         /* ntt.cpp:148 */
    uint32_t t = N >> 1;                                   /* ntt.cpp:149 */
    for (uint32_t m = 1; m < N; m <<= 1) {                 /* ntt.cpp:155 */
        for (uint32_t bf = 0; bf < N / 2; bf++) {          /* ntt.cpp:157-159 with k*VEC+n flattened */
            uint32_t i = bf / t, j = bf % t;               /* ntt.cpp:292-297 */
            uint32_t j1 = i * 2 * t;
            uint64_t W = roots[m + i], Wp = precons[m + i]; /* ntt.cpp:298-300 */
            uint64_t tx = X[j1 + j];
            if (tx >= twice) tx -= twice;                  /* ntt.cpp:331-332 */
            uint64_t a = X[j1 + j + t];
            uint64_t c1 = ref_mulhi64(a, Wp);              /* ntt.cpp:344-362 */
            uint64_t Q = W * a - c1 * q;                   /* ntt.cpp:363 (mod 2^64) */
            uint64_t o0 = tx + Q, o1 = tx + twice - Q;     /* ntt.cpp:368-369 */
            if (m == N / 2) {                              /* ntt.cpp:377-393 */
                if (o0 >= twice) o0 -= twice;
                if (o0 >= q) o0 -= q;
                if (o1 >= twice) o1 -= twice;
                if (o1 >= q) o1 -= q;
            }
            X[j1 + j] = o0;
            X[j1 + j + t] = o1;
        }
        t >>= 1;                                           /* ntt.cpp:499 */
    }
}

/* Frame b: low half from in[b*N + ...], high half from in2[b*N + N/2 + ...] (ntt.cpp:582-591); output row-major
 * [numFrames][N] (ntt.cpp:626-633), butterfly bf of the last stage emitting (out[2bf], out[2bf+1]) (:385,393). */
void orc_ref_fwd_u64(uint32_t N, const uint64_t *in, const uint64_t *in2, uint64_t modulus, const uint64_t *roots,
                     const uint64_t *precons, uint32_t numFrames, uint64_t *out) {
    for (uint32_t b = 0; b < numFrames; b++) {
        uint64_t *X = out + (size_t)b * N;
        const uint64_t *lo = in + (size_t)b * N, *hi = in2 + (size_t)b * N;
        uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * N);
        memcpy(tmp, lo, sizeof(uint64_t) * (N / 2));
        memcpy(tmp + N / 2, hi + N / 2, sizeof(uint64_t) * (N / 2));
        ref_fwd_one_u64(N, tmp, modulus, roots, precons);
        memcpy(X, tmp, sizeof(uint64_t) * N);
        free(tmp);
    }
}

void orc_batch_ref_fwd_u64(uint32_t N, uint64_t *data, uint64_t modulus, const uint64_t *roots,
                           const uint64_t *precons, size_t numFrames, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long long b = 0; b < (long long)numFrames; b++)
        ref_fwd_one_u64(N, data + (size_t)b * N, modulus, roots, precons);
}

/* ------------------------------------------------------------------------------------------ u32 datapath */

/* SEAL-Embedded-style non-lazy butterflies (published algorithm; see header).  Every value stays in [0,q). */
void orc_fwd_u32_barrett(uint32_t n, uint32_t q, const uint32_t *roots, uint32_t *x) {
    uint32_t t = n >> 1;
    for (uint32_t m = 1; m < n; m <<= 1, t >>= 1)
        for (uint32_t i = 0; i < m; i++) {
            uint64_t w = roots[m + i];
            uint32_t *p = x + 2 * i * t;
            for (uint32_t j = 0; j < t; j++) {
                uint32_t u = p[j], v = (uint32_t)((w * p[j + t]) % q);
                uint32_t s = u + v;
                p[j] = s >= q ? s - q : s;
                p[j + t] = u >= v ? u - v : u + q - v;
            }
        }
}

/* Gentleman-Sande inverse: (u,v) -> (u+v, (u-v)*w), then *n^-1. */
void orc_inv_u32_barrett(uint32_t n, uint32_t q, const uint32_t *iroots, uint32_t *x) {
    uint32_t t = 1;
    for (uint32_t h = n >> 1; h >= 1; h >>= 1, t <<= 1)
        for (uint32_t i = 0; i < h; i++) {
            uint64_t w = iroots[h + i];
            uint32_t *p = x + 2 * i * t;
            for (uint32_t j = 0; j < t; j++) {
                uint32_t u = p[j], v = p[j + t];
                uint32_t s = u + v;
                p[j] = s >= q ? s - q : s;
                uint32_t d = u >= v ? u - v : u + q - v;
                p[j + t] = (uint32_t)((w * d) % q);
            }
        }
    uint64_t ninv = orc_invmod(n, q);
    for (uint32_t j = 0; j < n; j++) x[j] = (uint32_t)((ninv * x[j]) % q);
}

/* Harvey lazy butterflies on [0,4q) with (w, floor(w*2^32/q)) -- the same arithmetic as ntt.cpp:331-393 at
 * half the word width.  Requires q < 2^30. */
void orc_fwd_u32_shoup(uint32_t n, uint32_t q, const uint32_t *roots, const uint32_t *precons, uint32_t *x) {
    const uint32_t twoq = q << 1;
    uint32_t t = n >> 1;
    for (uint32_t m = 1; m < n; m <<= 1, t >>= 1)
        for (uint32_t i = 0; i < m; i++) {
            uint32_t w = roots[m + i], wp = precons[m + i];
            uint32_t *p = x + 2 * i * t;
            for (uint32_t j = 0; j < t; j++) {
                uint32_t tx = p[j];
                if (tx >= twoq) tx -= twoq;
                uint32_t y = p[j + t];
                uint32_t c1 = (uint32_t)(((uint64_t)y * wp) >> 32);
                uint32_t Q = w * y - c1 * q;
                p[j] = tx + Q;
                p[j + t] = tx + twoq - Q;
            }
        }
    for (uint32_t j = 0; j < n; j++) {
        uint32_t v = x[j];
        if (v >= twoq) v -= twoq;
        if (v >= q) v -= q;
        x[j] = v;
    }
}

void orc_inv_u32_shoup(uint32_t n, uint32_t q, const uint32_t *iroots, const uint32_t *iprecons, uint32_t *x) {
    const uint32_t twoq = q << 1;
    uint32_t t = 1;
    for (uint32_t h = n >> 1; h >= 1; h >>= 1, t <<= 1)
        for (uint32_t i = 0; i < h; i++) {
            uint32_t w = iroots[h + i], wp = iprecons[h + i];
            uint32_t *p = x + 2 * i * t;
            for (uint32_t j = 0; j < t; j++) {
                uint32_t u = p[j], v = p[j + t];
                uint32_t s = u + v;
                if (s >= twoq) s -= twoq;
                uint32_t d = u + twoq - v;
                uint32_t c1 = (uint32_t)(((uint64_t)d * wp) >> 32);
                p[j] = s;
                p[j + t] = w * d - c1 * q;
            }
        }
    uint32_t ninv = (uint32_t)orc_invmod(n, q);
    uint32_t ninvp = (uint32_t)(((uint64_t)ninv << 32) / q);
    for (uint32_t j = 0; j < n; j++) {
        uint32_t v = x[j];
        uint32_t c1 = (uint32_t)(((uint64_t)v * ninvp) >> 32);
        uint32_t r = ninv * v - c1 * q;
        if (r >= q) r -= q;
        x[j] = r;
    }
}

/* ------------------------------------------------------------------------------------------------ plans */

orc_plan *orc_plan_create(uint32_t n, uint32_t nlimbs, const uint32_t *q) {
    orc_plan *p = (orc_plan *)calloc(1, sizeof *p);
    p->n = n;
    p->nlimbs = nlimbs;
    p->q = (uint32_t *)malloc(4 * nlimbs);
    p->psi = (uint32_t *)malloc(4 * nlimbs);
    p->roots = (uint32_t *)malloc((size_t)4 * nlimbs * n);
    p->precons = (uint32_t *)malloc((size_t)4 * nlimbs * n);
    p->iroots = (uint32_t *)malloc((size_t)4 * nlimbs * n);
    p->iprecons = (uint32_t *)malloc((size_t)4 * nlimbs * n);
    for (uint32_t l = 0; l < nlimbs; l++) {
        p->q[l] = q[l];
        p->psi[l] = (uint32_t)orc_min_psi(n, q[l]);
        if (!p->psi[l] || q[l] >= (1u << 30) || !orc_is_prime(q[l])) { orc_plan_destroy(p); return NULL; }
        orc_tables_u32(n, q[l], p->psi[l], 0, p->roots + (size_t)l * n, p->precons + (size_t)l * n);
        orc_tables_u32(n, q[l], p->psi[l], 1, p->iroots + (size_t)l * n, p->iprecons + (size_t)l * n);
    }
    return p;
}

void orc_plan_destroy(orc_plan *p) {
    if (!p) return;
    free(p->q); free(p->psi); free(p->roots); free(p->precons); free(p->iroots); free(p->iprecons);
    free(p);
}

void orc_batch_fwd_u32(const orc_plan *p, uint32_t *data, size_t B, int variant, int threads) {
    const size_t T = B * p->nlimbs;
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long long k = 0; k < (long long)T; k++) {
        uint32_t l = (uint32_t)(k % p->nlimbs);
        uint32_t *x = data + (size_t)k * p->n;
        if (variant == 0) orc_fwd_u32_barrett(p->n, p->q[l], p->roots + (size_t)l * p->n, x);
        else orc_fwd_u32_shoup(p->n, p->q[l], p->roots + (size_t)l * p->n, p->precons + (size_t)l * p->n, x);
    }
}

void orc_batch_inv_u32(const orc_plan *p, uint32_t *data, size_t B, int variant, int threads) {
    const size_t T = B * p->nlimbs;
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long long k = 0; k < (long long)T; k++) {
        uint32_t l = (uint32_t)(k % p->nlimbs);
        uint32_t *x = data + (size_t)k * p->n;
        if (variant == 0) orc_inv_u32_barrett(p->n, p->q[l], p->iroots + (size_t)l * p->n, x);
        else orc_inv_u32_shoup(p->n, p->q[l], p->iroots + (size_t)l * p->n, p->iprecons + (size_t)l * p->n, x);
    }
}

/* c = INTT(NTT(a) .* NTT(b)) per (polynomial, limb): BASELINE.json config 4 on the CPU NTT path. */
void orc_batch_polymul_u32(const orc_plan *p, uint32_t *c, const uint32_t *a, const uint32_t *b, size_t B,
                           int threads) {
    const size_t T = B * p->nlimbs;
    const uint32_t n = p->n;
    (void)threads;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        uint32_t *ta = (uint32_t *)malloc(4 * (size_t)n), *tb = (uint32_t *)malloc(4 * (size_t)n);
#pragma omp for schedule(static)
        for (long long k = 0; k < (long long)T; k++) {
            uint32_t l = (uint32_t)(k % p->nlimbs);
            uint32_t q = p->q[l];
            memcpy(ta, a + (size_t)k * n, 4 * (size_t)n);
            memcpy(tb, b + (size_t)k * n, 4 * (size_t)n);
            orc_fwd_u32_shoup(n, q, p->roots + (size_t)l * n, p->precons + (size_t)l * n, ta);
            orc_fwd_u32_shoup(n, q, p->roots + (size_t)l * n, p->precons + (size_t)l * n, tb);
            for (uint32_t j = 0; j < n; j++) ta[j] = (uint32_t)(((uint64_t)ta[j] * tb[j]) % q);
            orc_inv_u32_shoup(n, q, p->iroots + (size_t)l * n, p->iprecons + (size_t)l * n, ta);
            memcpy(c + (size_t)k * n, ta, 4 * (size_t)n);
        }
        free(ta); free(tb);
    }
}

void orc_polymul_schoolbook(uint32_t n, uint32_t q, const uint32_t *a, const uint32_t *b, uint32_t *c) {
    /* c[k] = sum_{i+j=k} a_i b_j - sum_{i+j=k+n} a_i b_j  (mod q): exact, accumulators reduced every term */
    for (uint32_t k = 0; k < n; k++) {
        uint64_t pos = 0, neg = 0;
        for (uint32_t i = 0; i <= k; i++) pos = (pos + (uint64_t)a[i] * b[k - i]) % q;
        for (uint32_t i = k + 1; i < n; i++) neg = (neg + (uint64_t)a[i] * b[n + k - i]) % q;
        c[k] = (uint32_t)((pos + q - neg) % q);
    }
}

/* ------------------------------------------------------------------------------------- synthetic + checks */

void orc_fill_synthetic(uint32_t *data, size_t B, uint32_t nlimbs, uint32_t n, const uint32_t *q, uint64_t seed,
                        size_t first_poly) {
    const size_t T = B * nlimbs;
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)T; k++) {
        uint32_t l = (uint32_t)(k % nlimbs);
        size_t g0 = ((size_t)first_poly * nlimbs + (size_t)k) * n;
        for (uint32_t j = 0; j < n; j++) data[(size_t)k * n + j] = (uint32_t)(orc_splitmix64(seed + g0 + j) % q[l]);
    }
}

uint64_t orc_checksum_u32(const uint32_t *data, size_t count, size_t first_index) {
    uint64_t s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long long i = 0; i < (long long)count; i++)
        s += orc_splitmix64(((uint64_t)first_index + (uint64_t)i) * 0xD6E8FEB86659FD93ULL + data[i]);
    return s;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
