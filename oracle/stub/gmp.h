/* Stub <gmp.h> for building the reference's Fp/Shamir/PRG translation units
 * WITHOUT libgmp (test infrastructure only; never linked into the product).
 *
 * scl/math/fields/ff_ops.h:26 includes scl/math/number.h, whose only contact
 * with GMP on this path is the data member `mpz_t m_value` (number.h:447).
 * No mpz_* function is ever called by the Mersenne-61/127, Vector, Matrix,
 * Polynomial, Lagrange, Shamir or PRG code, so a layout-compatible typedef is
 * enough for the compiler. */
#ifndef SCLGPU_ORACLE_STUB_GMP_H
#define SCLGPU_ORACLE_STUB_GMP_H
typedef struct {
  int _mp_alloc;
  int _mp_size;
  unsigned long* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
#endif
