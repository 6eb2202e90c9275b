/* c_example.c -- the C ABI from plain C99 (no CUDA headers, no C++): forward + inverse round trip and a negacyclic
 * product through the host-pointer entry points.  Build: gcc -std=c99 -I include c_example.c -L lib -lagxntt */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "agxntt.h"

int main(void) {
    const uint32_t n = 4096, primes[3] = {1053818881u, 1054015489u, 1054212097u};
    const size_t B = 8, words = B * 3 * n;
    agx_parms parms;
    agx_ctx *ctx = NULL;
    uint32_t *x = malloc(words * 4), *y = malloc(words * 4), *z = malloc(words * 4);
    size_t i;
    int rc;
    parms.n = n; parms.logn = 12; parms.nlimbs = 3; parms.q = primes;
    rc = agx_create(&ctx, &parms, 0);
    if (rc) { fprintf(stderr, "agx_create: %s\n", agx_error_string(rc)); return 1; }
    for (i = 0; i < words; i++) x[i] = (uint32_t)(i * 2654435761u) % primes[(i / n) % 3];
    rc = agx_ntt_fwd_host(ctx, x, y, B);
    if (!rc) rc = agx_ntt_inv_host(ctx, y, y, B);
    if (rc) { fprintf(stderr, "transform: %s\n", agx_error_string(rc)); return 1; }
    if (memcmp(x, y, words * 4)) { fprintf(stderr, "round trip mismatch\n"); return 1; }
    /* (X^(n-1)) * X = -1 mod (X^n + 1) in every limb */
    memset(x, 0, words * 4); memset(y, 0, words * 4);
    for (i = 0; i < B * 3; i++) { x[i * n + n - 1] = 1; y[i * n + 1] = 1; }
    rc = agx_polymul_host(ctx, z, x, y, B);
    if (rc) { fprintf(stderr, "polymul: %s\n", agx_error_string(rc)); return 1; }
    for (i = 0; i < B * 3; i++)
        if (z[i * n] != primes[i % 3] - 1 || z[i * n + 1] != 0) { fprintf(stderr, "polymul mismatch\n"); return 1; }
    agx_destroy(ctx);
    free(x); free(y); free(z);
    printf("c_example ok\n");
    return 0;
}
