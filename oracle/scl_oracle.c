/* oracle/scl_oracle.c -- TEST INFRASTRUCTURE ONLY (see scl_oracle.h).
 *
 * Plain-C restatement of the reference's (SCL 0.1.0) hot path:
 *   util::PRG          src/scl/util/prg.cc:34-146, include/scl/util/prg.h:34-43
 *   ff::Mersenne61     src/scl/math/fields/mersenne61.cc:33-95
 *   ff::Mersenne127    src/scl/math/fields/mersenne127.cc:33-123
 *   modAdd/Sub/Neg/Inv src/scl/math/fields/small_ff.h:28-92
 *   Vector / Matrix / Polynomial / Lagrange / Shamir: see scl_oracle_field.inc
 *
 * The reference does AES-128 with AES-NI intrinsics (prg.cc:34-80).  AES-128 is
 * FIPS-197; this file restates it portably (byte-wise S-box rounds) and is
 * pinned against the reference's own keystream in tests/golden.
 *
 * Parity status: PINNED (tests/test_oracle_golden.py).
 */
#define _POSIX_C_SOURCE 200809L
#include "scl_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ AES-128 */

static uint8_t g_sbox[256];
static uint32_t g_te0[256]; /* MixColumns(SubBytes(x)) column for row 0 */
static pthread_once_t g_aes_once = PTHREAD_ONCE_INIT;

static uint8_t xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }

static void aes_init_tables(void) {
  /* S-box from its definition: multiplicative inverse in GF(2^8) followed by
   * the affine map (FIPS-197 section 5.1.1). */
  uint8_t p = 1, q = 1;
  do {
    p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));
    q ^= (uint8_t)(q << 1);
    q ^= (uint8_t)(q << 2);
    q ^= (uint8_t)(q << 4);
    if (q & 0x80) q ^= 0x09;
    uint8_t x = (uint8_t)(q ^ (q << 1 | q >> 7) ^ (q << 2 | q >> 6) ^
                          (q << 3 | q >> 5) ^ (q << 4 | q >> 4));
    g_sbox[p] = (uint8_t)(x ^ 0x63);
  } while (p != 1);
  g_sbox[0] = 0x63;
  for (int i = 0; i < 256; ++i) {
    const uint8_t s = g_sbox[i], s2 = xtime(s), s3 = (uint8_t)(s2 ^ s);
    /* little-endian column word: byte0 = 2s, byte1 = s, byte2 = s, byte3 = 3s */
    g_te0[i] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) |
               ((uint32_t)s3 << 24);
  }
}

static inline uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }

/* Key schedule, FIPS-197 5.2 == aes128LoadKey (prg.cc:64-80).  Round keys as
 * four little-endian column words each. */
static void aes128_expand(const uint8_t key[16], uint32_t rk[44]) {
  static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10,
                                   0x20, 0x40, 0x80, 0x1b, 0x36};
  for (int i = 0; i < 4; ++i) memcpy(&rk[i], key + 4 * i, 4);
  for (int i = 4; i < 44; ++i) {
    uint32_t t = rk[i - 1];
    if ((i & 3) == 0) {
      /* RotWord (bytes b0 b1 b2 b3 -> b1 b2 b3 b0) then SubWord then Rcon */
      t = (t >> 8) | (t << 24);
      t = (uint32_t)g_sbox[t & 0xff] | ((uint32_t)g_sbox[(t >> 8) & 0xff] << 8) |
          ((uint32_t)g_sbox[(t >> 16) & 0xff] << 16) |
          ((uint32_t)g_sbox[t >> 24] << 24);
      t ^= rcon[i / 4 - 1];
    }
    rk[i] = rk[i - 4] ^ t;
  }
}

/* One block, FIPS-197 5.1 == DO_ENC_BLOCK (prg.cc:36-49).  State = 4 LE column
 * words; out column j of a round = XOR over rows r of Te_r[ byte r of column
 * (j + r) mod 4 ]. */
static void aes128_encrypt(const uint32_t rk[44], const uint32_t in[4],
                           uint32_t out[4]) {
  uint32_t s0 = in[0] ^ rk[0], s1 = in[1] ^ rk[1], s2 = in[2] ^ rk[2],
           s3 = in[3] ^ rk[3];
  for (int r = 1; r < 10; ++r) {
    const uint32_t t0 = g_te0[s0 & 0xff] ^ rotl32(g_te0[(s1 >> 8) & 0xff], 8) ^
                        rotl32(g_te0[(s2 >> 16) & 0xff], 16) ^
                        rotl32(g_te0[s3 >> 24], 24) ^ rk[4 * r];
    const uint32_t t1 = g_te0[s1 & 0xff] ^ rotl32(g_te0[(s2 >> 8) & 0xff], 8) ^
                        rotl32(g_te0[(s3 >> 16) & 0xff], 16) ^
                        rotl32(g_te0[s0 >> 24], 24) ^ rk[4 * r + 1];
    const uint32_t t2 = g_te0[s2 & 0xff] ^ rotl32(g_te0[(s3 >> 8) & 0xff], 8) ^
                        rotl32(g_te0[(s0 >> 16) & 0xff], 16) ^
                        rotl32(g_te0[s1 >> 24], 24) ^ rk[4 * r + 2];
    const uint32_t t3 = g_te0[s3 & 0xff] ^ rotl32(g_te0[(s0 >> 8) & 0xff], 8) ^
                        rotl32(g_te0[(s1 >> 16) & 0xff], 16) ^
                        rotl32(g_te0[s2 >> 24], 24) ^ rk[4 * r + 3];
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
  }
#define SB(w, sh) ((uint32_t)g_sbox[((w) >> (sh)) & 0xff])
  out[0] = (SB(s0, 0) | SB(s1, 8) << 8 | SB(s2, 16) << 16 | SB(s3, 24) << 24) ^ rk[40];
  out[1] = (SB(s1, 0) | SB(s2, 8) << 8 | SB(s3, 16) << 16 | SB(s0, 24) << 24) ^ rk[41];
  out[2] = (SB(s2, 0) | SB(s3, 8) << 8 | SB(s0, 16) << 16 | SB(s1, 24) << 24) ^ rk[42];
  out[3] = (SB(s3, 0) | SB(s0, 8) << 8 | SB(s1, 16) << 16 | SB(s2, 24) << 24) ^ rk[43];
#undef SB
}

/* PRG_NONCE / PRG_INITIAL_COUNTER defaults, prg.h:34-43 */
#define SCLO_PRG_NONCE 0x0123456789ABCDEFULL

/* Keystream block i = AES_seed( _mm_set_epi64x(PRG_NONCE, i) ) (prg.cc:82-84):
 * plaintext bytes = LE64(i) || LE64(nonce).  next(buf, n) emits ceil(n/16)
 * blocks and copies the first n bytes (prg.cc:124-146). */
void sclo_prg_next(const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes,
                   uint8_t* out) {
  pthread_once(&g_aes_once, aes_init_tables);
  uint32_t rk[44];
  aes128_expand(seed, rk);
  const uint64_t nblocks = (n_bytes + 15) / 16;
  for (uint64_t i = 0; i < nblocks; ++i) {
    const uint64_t ctr = first_block + i;
    const uint32_t in[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32),
                            (uint32_t)SCLO_PRG_NONCE,
                            (uint32_t)(SCLO_PRG_NONCE >> 32)};
    uint32_t ct[4];
    aes128_encrypt(rk, in, ct);
    const uint64_t left = n_bytes - i * 16;
    memcpy(out + i * 16, ct, left < 16 ? left : 16);
  }
}

/* ------------------------------------------------------------ field multiply */

#define SCLO_P61 0x1FFFFFFFFFFFFFFFULL
#define SCLO_P127 ((((sclo_u128)0x7FFFFFFFFFFFFFFFULL) << 64) | 0xFFFFFFFFFFFFFFFFULL)

/* ff::multiply<Mersenne61>, mersenne61.cc:59-69 */
static inline uint64_t sclo_fp61_mul1(uint64_t x, uint64_t y) {
  const sclo_u128 z = (sclo_u128)x * y;
  uint64_t a = (uint64_t)(z >> 61);
  uint64_t b = (uint64_t)z;
  a |= b >> 61;
  b &= SCLO_P61;
  a += b;
  if (a >= SCLO_P61) a -= SCLO_P61;
  return a;
}

/* ff::multiply<Mersenne127>, mersenne127.cc:60-97: schoolbook 64x64 partial
 * products (multiplyFull :66-83), hi<<1 | lo>>127, lo & p, modAdd */
static inline sclo_u128 sclo_fp127_mul1(sclo_u128 x, sclo_u128 y) {
  const uint64_t a = (uint64_t)(x >> 64), b = (uint64_t)x;
  const uint64_t c = (uint64_t)(y >> 64), d = (uint64_t)y;
  const sclo_u128 ac = (sclo_u128)a * c, ad = (sclo_u128)a * d,
                  bc = (sclo_u128)b * c, bd = (sclo_u128)b * d;
  const sclo_u128 carry = (sclo_u128)(uint64_t)ad + (sclo_u128)(uint64_t)bc + (bd >> 64);
  const sclo_u128 high = ac + (ad >> 64) + (bc >> 64) + (carry >> 64);
  const sclo_u128 low = (ad << 64) + (bc << 64) + bd;
  sclo_u128 out = high << 1;
  sclo_u128 lo = low;
  out |= lo >> 127;
  lo &= SCLO_P127;
  out = out + lo;
  if (out >= SCLO_P127) out -= SCLO_P127;
  return out;
}

/* ------------------------------------------------ field-generic instantiation */

#define FE uint64_t
#define SFE int64_t
#define FP SCLO_P61
#define FBYTES 8
#define FMUL sclo_fp61_mul1
#define FN(x) sclo_fp61_##x
#include "scl_oracle_field.inc"
#undef FE
#undef SFE
#undef FP
#undef FBYTES
#undef FMUL
#undef FN

#define FE sclo_u128
#define SFE __int128
#define FP SCLO_P127
#define FBYTES 16
#define FMUL sclo_fp127_mul1
#define FN(x) sclo_fp127_##x
#include "scl_oracle_field.inc"
