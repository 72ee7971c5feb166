// field.cuh -- Mersenne-61 / Mersenne-127 arithmetic for sm_100a.
//
// Semantics follow SCL's scl::math::Fp (values are the unique canonical residue
// in [0,p), so any correct evaluation order is bit-identical to the reference):
//   add/sub/neg : src/scl/math/fields/small_ff.h:28-56 (modAdd/modSub/modNeg)
//   mul  Fp61   : src/scl/math/fields/mersenne61.cc:59-69
//   mul  Fp127  : src/scl/math/fields/mersenne127.cc:60-97
//   from_raw    : FF::read = LE word "% p"  (mersenne61.cc:87-90, mersenne127.cc:115-118)
//   inverse     : a^(p-2) here (the reference uses extended Euclid,
//                 small_ff.h:61-92; the inverse is unique, so equal)
//
// Elements in HBM are SCL's FF::write bytes: u64 for Fp61, 16-byte little-endian
// (lo, hi) for Fp127.  The Mersenne reductions are shifts and adds on the
// 64x64->128 products; lazy 128-bit accumulation is used for dot products.
#pragma once
#include <cstdint>

namespace sclgpu {

#define SCLGPU_HD __host__ __device__ __forceinline__
#define SCLGPU_D __device__ __forceinline__

typedef unsigned __int128 u128;

// 64x64 -> 128 as (lo, hi)
SCLGPU_HD void mul64wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  const u128 z = (u128)a * b;
  lo = (uint64_t)z;
  hi = (uint64_t)(z >> 64);
#endif
}

// ============================================================== Mersenne-61
struct F61 {
  typedef uint64_t E;
  static constexpr int BYTES = 8;
  static constexpr uint64_t P = 0x1FFFFFFFFFFFFFFFULL;

  static SCLGPU_HD E zero() { return 0; }
  static SCLGPU_HD E one() { return 1; }
  static SCLGPU_HD E from_u32(uint32_t v) { return v; }  // FF(int), v >= 0 (mersenne61.cc:38-40)
  static SCLGPU_HD bool eq(E a, E b) { return a == b; }
  static SCLGPU_HD bool is_zero(E a) { return a == 0; }

  // any 64-bit word -> canonical ("% p")
  static SCLGPU_HD E from_raw(uint64_t w) {
    uint64_t r = (w & P) + (w >> 61);  // <= p + 7
    return r >= P ? r - P : r;
  }
  static SCLGPU_HD E add(E a, E b) {
    uint64_t r = a + b;
    return r >= P ? r - P : r;
  }
  static SCLGPU_HD E sub(E a, E b) { return b > a ? a + P - b : a - b; }
  static SCLGPU_HD E neg(E a) { return a ? P - a : 0; }
  static SCLGPU_HD E mul(E a, E b) {
    uint64_t lo, hi;
    mul64wide(a, b, lo, hi);
    uint64_t r = (lo & P) + ((hi << 3) | (lo >> 61));  // < 2p for canonical inputs
    return r >= P ? r - P : r;
  }

  // lazy dot-product accumulator: sum of <= 32 products of canonical values
  // fits 128 bits (32 * 2^122 < 2^128); fold() brings it back below 2^62.
  struct Acc {
    uint64_t lo, hi;
  };
  static constexpr int ACC_TERMS = 32;
  static SCLGPU_HD Acc acc_zero() { return Acc{0, 0}; }
  static SCLGPU_HD void mac(Acc& s, E a, E b) {
    uint64_t lo, hi;
    mul64wide(a, b, lo, hi);
    s.lo += lo;
    s.hi += hi + (s.lo < lo);
  }
  // value of the accumulator mod p, canonical
  static SCLGPU_HD E acc_reduce(const Acc& s) {
    // 2^64 = 8 (mod p); 8*hi = ((hi & (2^58-1)) << 3) + (hi >> 58) * 2^61
    uint64_t r = (s.lo & P) + (s.lo >> 61) + ((s.hi & 0x03FFFFFFFFFFFFFFULL) << 3) + (s.hi >> 58);
    return from_raw(r);  // r < 2^63
  }
  static SCLGPU_HD void acc_fold(Acc& s) {
    s.lo = acc_reduce(s);
    s.hi = 0;
  }
  static SCLGPU_HD void acc_merge(Acc& s, const Acc& o) {  // both folded
    s.lo += o.lo;
    s.hi += o.hi + (s.lo < o.lo);
  }

  static SCLGPU_HD E pow(E a, uint64_t e_lo, uint64_t e_hi) {
    (void)e_hi;
    E r = 1;
    for (int i = 60; i >= 0; --i) {
      r = mul(r, r);
      if ((e_lo >> i) & 1) r = mul(r, a);
    }
    return r;
  }
  // a != 0
  static SCLGPU_HD E inv(E a) { return pow(a, P - 2, 0); }
};

// ============================================================= Mersenne-127
struct alignas(16) E127 {
  uint64_t lo, hi;
};

struct F127 {
  typedef E127 E;
  static constexpr int BYTES = 16;
  static constexpr uint64_t PHI = 0x7FFFFFFFFFFFFFFFULL;  // p = (PHI << 64) | ~0

  static SCLGPU_HD E zero() { return E{0, 0}; }
  static SCLGPU_HD E one() { return E{1, 0}; }
  static SCLGPU_HD E from_u32(uint32_t v) { return E{v, 0}; }
  static SCLGPU_HD bool eq(E a, E b) { return a.lo == b.lo && a.hi == b.hi; }
  static SCLGPU_HD bool is_zero(E a) { return (a.lo | a.hi) == 0; }
  static SCLGPU_HD bool ge_p(E a) { return a.hi > PHI || (a.hi == PHI && a.lo == ~0ULL); }
  static SCLGPU_HD E sub_p(E a) {  // a - p = a + 1 - 2^127
    E r;
    r.lo = a.lo + 1;
    r.hi = a.hi + (r.lo == 0) - 0x8000000000000000ULL;
    return r;
  }
  static SCLGPU_HD E from_raw(E w) {  // any 128-bit word -> canonical
    E r;
    const uint64_t top = w.hi >> 63;
    r.lo = w.lo + top;
    r.hi = (w.hi & PHI) + (r.lo < top);  // <= p + 1
    return ge_p(r) ? sub_p(r) : r;
  }
  static SCLGPU_HD E add(E a, E b) {
    E r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo);  // < 2^128: no overflow for canonical inputs
    return ge_p(r) ? sub_p(r) : r;
  }
  static SCLGPU_HD bool gt(E a, E b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }
  static SCLGPU_HD E sub(E a, E b) {
    E r;
    r.lo = a.lo - b.lo;
    r.hi = a.hi - b.hi - (a.lo < b.lo);
    if (gt(b, a)) {  // + p = + 2^127 - 1
      const uint64_t borrow = (r.lo == 0);
      r.lo -= 1;
      r.hi = r.hi + 0x8000000000000000ULL - borrow;
    }
    return r;
  }
  static SCLGPU_HD E neg(E a) { return is_zero(a) ? a : sub(E{~0ULL, PHI}, a); }

  // 128x128 -> 256, then hi*2 + lo (2^127 = 1 mod p)
  static SCLGPU_HD E mul(E a, E b) {
    uint64_t p00l, p00h, p01l, p01h, p10l, p10h, p11l, p11h;
    mul64wide(a.lo, b.lo, p00l, p00h);
    mul64wide(a.lo, b.hi, p01l, p01h);
    mul64wide(a.hi, b.lo, p10l, p10h);
    mul64wide(a.hi, b.hi, p11l, p11h);
    // mid = p01 + p10 < 2^128 (a.hi, b.hi < 2^63)
    uint64_t ml = p01l + p10l;
    uint64_t mh = p01h + p10h + (ml < p01l);
    // 256-bit product words w0..w3
    uint64_t w0 = p00l;
    uint64_t w1 = p00h + ml;
    uint64_t c1 = (w1 < ml);
    uint64_t w2 = p11l + mh;
    uint64_t c2 = (w2 < mh);
    w2 += c1;
    c2 += (w2 < c1);
    uint64_t w3 = p11h + c2;
    // lo127 = (w1 & PHI, w0); hi = product >> 127 = (w3:w2:w1) >> 63   (< 2^127)
    E l{w0, w1 & PHI};
    E h{(w1 >> 63) | (w2 << 1), (w2 >> 63) | (w3 << 1)};
    E r;
    r.lo = l.lo + h.lo;
    r.hi = l.hi + h.hi + (r.lo < l.lo);
    return ge_p(r) ? sub_p(r) : r;
  }

  // dot products: accumulate canonical products with modular adds (simple form)
  struct Acc {
    E v;
  };
  static constexpr int ACC_TERMS = 1 << 30;
  static SCLGPU_HD Acc acc_zero() { return Acc{E{0, 0}}; }
  static SCLGPU_HD void mac(Acc& s, E a, E b) { s.v = add(s.v, mul(a, b)); }
  static SCLGPU_HD E acc_reduce(const Acc& s) { return s.v; }
  static SCLGPU_HD void acc_fold(Acc&) {}
  static SCLGPU_HD void acc_merge(Acc& s, const Acc& o) { s.v = add(s.v, o.v); }

  static SCLGPU_HD E inv(E a) {  // a^(p-2), p-2 = 2^127 - 3 = 0b111...1101 (127 bits)
    E r = one();
    for (int i = 126; i >= 0; --i) {
      r = mul(r, r);
      if (i != 1) r = mul(r, a);  // every exponent bit is 1 except bit 1
    }
    return r;
  }
};

}  // namespace sclgpu
