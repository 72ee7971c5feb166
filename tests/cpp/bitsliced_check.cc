// tests/cpp/bitsliced_check.cc -- TEST INFRASTRUCTURE (CPU only): the bitsliced AES-128-CTR of
// secure-computation-library_b200/csrc/aes_bitsliced.cuh, compiled for the host, against the oracle's PRG
// (oracle/scl_oracle.c, util::PRG of src/scl/util/prg.cc:82-84, 124-146) on several seeds and counter ranges.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../oracle/scl_oracle.h"
#include "../../secure-computation-library_b200/csrc/aes_bitsliced.cuh"

// FIPS-197 key expansion into 44 little-endian column words, S-box from its definition (inverse + affine map)
static uint8_t g_sbox[256];
static void make_sbox() {
  uint8_t p = 1, q = 1;
  do {
    p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));
    q ^= (uint8_t)(q << 1);
    q ^= (uint8_t)(q << 2);
    q ^= (uint8_t)(q << 4);
    if (q & 0x80) q ^= 0x09;
    const uint8_t x = (uint8_t)(q ^ (uint8_t)(q << 1 | q >> 7) ^ (uint8_t)(q << 2 | q >> 6) ^ (uint8_t)(q << 3 | q >> 5) ^
                                (uint8_t)(q << 4 | q >> 4));
    g_sbox[p] = (uint8_t)(x ^ 0x63);
  } while (p != 1);
  g_sbox[0] = 0x63;
}
static void expand(const uint8_t seed[16], uint32_t rk[44]) {
  std::memcpy(rk, seed, 16);
  uint32_t rcon = 1;
  for (int i = 4; i < 44; ++i) {
    uint32_t t = rk[i - 1];
    if ((i & 3) == 0) {
      t = (t >> 8) | (t << 24);
      t = (uint32_t)g_sbox[t & 0xff] | ((uint32_t)g_sbox[(t >> 8) & 0xff] << 8) | ((uint32_t)g_sbox[(t >> 16) & 0xff] << 16) |
          ((uint32_t)g_sbox[t >> 24] << 24);
      t ^= rcon;
      rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11b : 0);
    }
    rk[i] = rk[i - 4] ^ t;
  }
}

int main() {
  make_sbox();
  // the S-box circuit alone, all 256 inputs (bit l of plane k = bit k of input 32 * round + l)
  for (int base = 0; base < 256; base += 32) {
    uint32_t u[8] = {0};
    for (int l = 0; l < 32; ++l)
      for (int k = 0; k < 8; ++k) u[k] |= (uint32_t)(((base + l) >> k) & 1) << l;
    sclgpu::bs_sbox(u);
    for (int l = 0; l < 32; ++l) {
      int y = 0;
      for (int k = 0; k < 8; ++k) y |= (int)((u[k] >> l) & 1) << k;
      if (y != g_sbox[base + l]) {
        std::printf("S-box circuit: input %d gives %d, want %d\n", base + l, y, g_sbox[base + l]);
        return 1;
      }
    }
  }
  const char* seeds[] = {"", "shamir bench", "prg bench", "0123456789abcdefXYZ"};
  const uint64_t groups[] = {0, 1, 7, (1ull << 32) / 32 - 1, (1ull << 32) / 32, (1ull << 40) / 32 + 5, (1ull << 58)};
  long checked = 0;
  for (const char* sd : seeds) {
    uint8_t seed[16] = {0};
    std::memcpy(seed, sd, std::strlen(sd) < 16 ? std::strlen(sd) : 16);
    uint32_t rk[44];
    expand(seed, rk);
    std::vector<uint32_t> km(sclgpu::kBsKeyWords);
    for (uint32_t i = 0; i < sclgpu::kBsKeyWords; ++i) km[i] = sclgpu::bs_key_mask(rk, i);
    for (uint64_t g : groups) {
      uint32_t s[128];
      sclgpu::bs_aes_ctr32(km.data(), g << 5, s);
      uint8_t want[32 * 16];
      sclo_prg_next(seed, g << 5, sizeof(want), want);
      for (int l = 0; l < 32; ++l) {
        uint32_t w[4] = {s[l], s[32 + l], s[64 + l], s[96 + l]};
        if (std::memcmp(w, want + 16 * l, 16) != 0) {
          std::printf("seed '%s' group %llu block %d differs\n", sd, (unsigned long long)g, l);
          return 1;
        }
        ++checked;
      }
    }
  }
  std::printf("BITSLICED_CHECK PASSED %ld blocks\n", checked);
  return 0;
}
