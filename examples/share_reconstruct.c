/* examples/share_reconstruct.c -- the C ABI of libsclgpu.so used from plain C (no CUDA, no C++ headers):
 * share N secrets over Fp<61> (n = 5, t = 2), corrupt one share of one sharing, reconstruct with error
 * detection and with error correction.  Build (see tests/cpp/Makefile):
 *   gcc -std=c11 -Iinclude examples/share_reconstruct.c -Lsecure-computation-library_b200/csrc -lsclgpu
 * Every call below names the SCL function it batches (include/sclgpu.h). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sclgpu.h"

#define CHECK(call)                                                                          \
  do {                                                                                       \
    int rc_ = (call);                                                                        \
    if (rc_ != SCLGPU_OK) {                                                                  \
      fprintf(stderr, "%s -> %s (%s)\n", #call, sclgpu_strerror(rc_), sclgpu_last_error(ctx)); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

int main(void) {
  enum { N = 1000, n = 7, t = 2 };
  sclgpu_ctx* ctx = NULL;
  if (sclgpu_init(0, &ctx) != SCLGPU_OK) {
    fprintf(stderr, "no usable sm_100 device (there is no CPU fallback)\n");
    return 2;
  }
  uint8_t seed[16] = {0};
  memcpy(seed, "c example", 9); /* PRG::create("c example"): zero padded to 16 bytes (prg.cc:88-101) */

  uint64_t* secrets = malloc(sizeof(uint64_t) * N);
  uint64_t* shares = malloc(sizeof(uint64_t) * N * n); /* [N][n]: row j = shamirSecretShare(secret_j, t, n, prg) */
  uint64_t* out = malloc(sizeof(uint64_t) * N);
  uint8_t* flags = malloc(N);
  CHECK(sclgpu_fp61_random(ctx, seed, 0, N, secrets));                       /* Vector<Fp61>::random(N, prg) */
  const uint64_t consumed = (N * 8 + 15) / 16;                                /* blocks that draw consumed   */
  CHECK(sclgpu_fp61_shamir_share(ctx, secrets, N, t, n, seed, consumed, shares));
  CHECK(sclgpu_fp61_recover_p(ctx, shares, N, n, NULL, NULL, out));           /* shamirRecoverP(shares)       */
  if (memcmp(out, secrets, sizeof(uint64_t) * N) != 0) return 3;

  shares[17 * n + 3] ^= 1;                                                    /* one bad share in sharing 17  */
  uint64_t n_bad = 0;
  int rc = sclgpu_fp61_recover_d(ctx, shares, N, n, t, NULL, 0, 0, NULL, out, flags, &n_bad); /* shamirRecoverD(shares, t) */
  if (rc != SCLGPU_EDETECT || n_bad != 1 || !flags[17]) return 4;
  printf("recoverD: \"%s\" for %llu sharing(s), first flagged = 17\n", sclgpu_last_error(ctx), (unsigned long long)n_bad);

  /* shamirRecoverC: t = (n-1)/3 = 2 errors per sharing can be corrected; the secret is f[j][0] */
  const uint32_t np = 3 * ((n - 1) / 3) + 1;
  uint64_t* f = malloc(sizeof(uint64_t) * N * np);
  uint64_t* err = malloc(sizeof(uint64_t) * N * ((n - 1) / 3 + 1));
  uint64_t n_failed = 0;
  CHECK(sclgpu_fp61_recover_c(ctx, shares, N, n, NULL, f, err, flags, &n_failed));
  for (int j = 0; j < N; ++j)
    if (f[(size_t)j * np] != secrets[j]) return 5;
  printf("recoverC: all %d secrets recovered, error locator of sharing 17 has root x = %llu\n", N,
         (unsigned long long)((0x1FFFFFFFFFFFFFFFull - err[17 * 3]) % 0x1FFFFFFFFFFFFFFFull)); /* err = x - a: root a = -err[0] */
  printf("C_EXAMPLE_OK launches=%llu\n", (unsigned long long)sclgpu_launch_count(ctx));
  free(secrets); free(shares); free(out); free(flags); free(f); free(err);
  sclgpu_destroy(ctx);
  return 0;
}
