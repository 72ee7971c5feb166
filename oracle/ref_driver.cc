// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never part of the product).
//
// A thin extern "C" driver around the UNMODIFIED reference (SCL 0.1.0) sources.
// It is compiled by oracle/Makefile together with, and only with, these
// reference translation units taken where they lie under /root/reference:
//     src/scl/math/fields/mersenne61.cc   src/scl/math/fields/mersenne127.cc
//     src/scl/util/prg.cc                 src/scl/util/str.cc
// and the reference's headers (include/scl/...), plus oracle/stub/gmp.h.
// The output goes to oracle/_ref/libsclref.so (git-ignored, travels to the GPU
// box).  Everything below calls the reference's OWN functions:
//     util::PRG::next                    (src/scl/util/prg.cc:124-146)
//     math::Vector<T>::random            (include/scl/math/vector.h:508-519)
//     math::FF<F>::random                (include/scl/math/ff.h:72-76)
//     ss::shamirSecretShare              (include/scl/ss/shamir.h:52-68)
//     ss::shamirRecoverP / shamirRecoverD(include/scl/ss/shamir.h:82-155)
//     ss::additiveShare                  (include/scl/ss/additive.h:42-53)
//     ss::shamirRecoverC                 (include/scl/ss/shamir.h:203-258)
//     math::computeLagrangeBasis         (include/scl/math/lagrange.h:55-71)
//     math::Matrix<T>::multiply(Vector)  (include/scl/math/matrix.h:498-513)
//     math::Matrix<T>::multiply(Matrix)  (include/scl/math/matrix.h:476-495)
//     math::Matrix<T>::hyperInvertible   (include/scl/math/matrix.h:462-475)
//     ss::shamirSecretShare on math::Array<T, W> (the sharing step of
//         ss::pedersenSecretShare, include/scl/ss/pedersen.h:137-138)
//     Vector add/subtract/multiplyEntryWise/scalarMultiply/dot/sum
// All element buffers are the reference's FF::write bytes (little-endian
// canonical residues, 8 B for Fp<61>, 16 B for Fp<127>).

#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "scl/math/array.h"
#include "scl/math/fp.h"
#include "scl/math/lagrange.h"
#include "scl/math/matrix.h"
#include "scl/math/poly.h"
#include "scl/math/vector.h"
#include "scl/ss/additive.h"
#include "scl/ss/shamir.h"
#include "scl/util/prg.h"

namespace {

using scl::util::PRG;
using Fp61 = scl::math::Fp<61>;
using Fp127 = scl::math::Fp<127>;

// PRG has no seek API (prg.h:64-173): fast-forward by drawing and discarding.
void skipBlocks(PRG& prg, uint64_t blocks) {
  std::vector<unsigned char> sink(1 << 20);
  uint64_t bytes_left = blocks * 16;
  while (bytes_left > 0) {
    const std::size_t n =
        bytes_left > sink.size() ? sink.size() : (std::size_t)bytes_left;
    prg.next(sink.data(), n);  // n is a multiple of 16: no tail is dropped
    bytes_left -= n;
  }
}

PRG makePrg(const unsigned char* seed, uint64_t seed_len, uint64_t skip) {
  PRG prg = PRG::create(seed, seed_len);
  if (skip) {
    skipBlocks(prg, skip);
  }
  return prg;
}

template <typename T>
scl::math::Vector<T> readVec(const unsigned char* src, std::size_t n) {
  std::vector<T> v;
  v.reserve(n);
  for (std::size_t i = 0; i < n; ++i) {
    v.emplace_back(T::read(src + i * T::byteSize()));
  }
  return scl::math::Vector<T>(v);
}

template <typename T>
void writeVec(const scl::math::Vector<T>& v, unsigned char* dst) {
  for (std::size_t i = 0; i < v.size(); ++i) {
    v[i].write(dst + i * T::byteSize());
  }
}

template <typename T>
void vectorRandom(const unsigned char* seed, uint64_t seed_len, uint64_t skip,
                  uint64_t n, unsigned char* out) {
  PRG prg = makePrg(seed, seed_len, skip);
  // chunked so that a 2^28 draw does not need 3 copies in RAM; chunk sizes are
  // multiples of 16 bytes so the concatenation equals one big Vector::random.
  const uint64_t chunk = 1 << 20;
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = (n - off) < chunk ? (n - off) : chunk;
    auto v = scl::math::Vector<T>::random(m, prg);
    writeVec(v, out + off * T::byteSize());
  }
}

template <typename T>
void ffRandom(const unsigned char* seed, uint64_t seed_len, uint64_t skip,
              uint64_t n, unsigned char* out) {
  PRG prg = makePrg(seed, seed_len, skip);
  for (uint64_t i = 0; i < n; ++i) {
    T::random(prg).write(out + i * T::byteSize());
  }
}

template <typename T>
void shamirShare(const unsigned char* secrets, uint64_t N, uint64_t t,
                 uint64_t n, const unsigned char* seed, uint64_t seed_len,
                 uint64_t skip, unsigned char* shares) {
  PRG prg = makePrg(seed, seed_len, skip);
  const std::size_t bs = T::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    const T secret = T::read(secrets + j * bs);
    const auto sh = scl::ss::shamirSecretShare(secret, t, n, prg);
    writeVec(sh, shares + j * n * bs);
  }
}

// ss::shamirSecretShare(math::Array<T, W>, t, n, prg) per array-valued secret
// on one PRG: the sharing step of pedersenSecretShare (pedersen.h:137-138 uses
// W = 2: {secret, randomness}).  secrets: N x W elements, shares: N x n x W.
template <typename T, std::size_t W>
void shamirShareArrayW(const unsigned char* secrets, uint64_t N, uint64_t t,
                       uint64_t n, PRG& prg, unsigned char* shares) {
  using A = scl::math::Array<T, W>;
  const std::size_t bs = A::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    const A secret = A::read(secrets + j * bs);
    const auto sh = scl::ss::shamirSecretShare(secret, t, n, prg);
    for (uint64_t i = 0; i < n; ++i) sh[i].write(shares + (j * n + i) * bs);
  }
}

template <typename T>
int shamirShareArray(const unsigned char* secrets, uint64_t N, uint64_t W,
                     uint64_t t, uint64_t n, const unsigned char* seed,
                     uint64_t seed_len, uint64_t skip, unsigned char* shares) {
  PRG prg = makePrg(seed, seed_len, skip);
  switch (W) {
    case 1: shamirShareArrayW<T, 1>(secrets, N, t, n, prg, shares); return 0;
    case 2: shamirShareArrayW<T, 2>(secrets, N, t, n, prg, shares); return 0;
    case 3: shamirShareArrayW<T, 3>(secrets, N, t, n, prg, shares); return 0;
    case 4: shamirShareArrayW<T, 4>(secrets, N, t, n, prg, shares); return 0;
    case 5: shamirShareArrayW<T, 5>(secrets, N, t, n, prg, shares); return 0;
    default: return -1;  // widths instantiated in this driver: 1..5
  }
}

// ss::shamirRecoverP on a Vector<Array<T, W>> (shamir.h:100-104 -> :82-87, the
// Lagrange basis being Array-valued with equal components).
template <typename T, std::size_t W>
void recoverPArrayW(const unsigned char* shares, uint64_t N, uint64_t n,
                    unsigned char* out) {
  using A = scl::math::Array<T, W>;
  const std::size_t bs = A::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    std::vector<A> v;
    v.reserve(n);
    for (uint64_t i = 0; i < n; ++i) v.emplace_back(A::read(shares + (j * n + i) * bs));
    const A s = scl::ss::shamirRecoverP(scl::math::Vector<A>(v));
    s.write(out + j * bs);
  }
}

template <typename T>
int recoverPArray(const unsigned char* shares, uint64_t N, uint64_t W,
                  uint64_t n, unsigned char* out) {
  switch (W) {
    case 1: recoverPArrayW<T, 1>(shares, N, n, out); return 0;
    case 2: recoverPArrayW<T, 2>(shares, N, n, out); return 0;
    case 3: recoverPArrayW<T, 3>(shares, N, n, out); return 0;
    case 4: recoverPArrayW<T, 4>(shares, N, n, out); return 0;
    case 5: recoverPArrayW<T, 5>(shares, N, n, out); return 0;
    default: return -1;
  }
}

// math::Matrix<T>::hyperInvertible(n, m) (matrix.h:462-475), row-major.
template <typename T>
void hyperInvertible(uint64_t n, uint64_t m, unsigned char* out) {
  const auto him = scl::math::Matrix<T>::hyperInvertible(n, m);
  const std::size_t bs = T::byteSize();
  for (uint64_t i = 0; i < n; ++i)
    for (uint64_t j = 0; j < m; ++j) him(i, j).write(out + (i * m + j) * bs);
}

// ss::additiveShare(secret, n, prg) per secret on one PRG; reconstruction is
// shares.sum() (additive.h:38-39).
template <typename T>
void additiveShare(const unsigned char* secrets, uint64_t N, uint64_t n,
                   const unsigned char* seed, uint64_t seed_len, uint64_t skip,
                   unsigned char* shares) {
  PRG prg = makePrg(seed, seed_len, skip);
  const std::size_t bs = T::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    const T secret = T::read(secrets + j * bs);
    const auto sh = scl::ss::additiveShare(secret, n, prg);
    writeVec(sh, shares + j * n * bs);
  }
}

template <typename T>
void additiveRecover(const unsigned char* shares, uint64_t N, uint64_t n,
                     unsigned char* out) {
  const std::size_t bs = T::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    readVec<T>(shares + j * n * bs, n).sum().write(out + j * bs);
  }
}

// ss::shamirRecoverC(shares[, alphas]) per sharing (Berlekamp-Welch, shamir.h:203-258).
// t = (n-1)/3, np = 3t+1.  f_out: [N][np] coefficients of the recovered polynomial, zero padded;
// e_out: [N][t+1] coefficients of the (monic) error polynomial, zero padded; status[j] = 1 where the
// reference throws std::logic_error("could not correct shares").  Returns the number of such j.
template <typename T>
int64_t recoverC(const unsigned char* shares, uint64_t N, uint64_t n,
                 const unsigned char* alphas, unsigned char* f_out,
                 unsigned char* e_out, unsigned char* status) {
  const std::size_t bs = T::byteSize();
  const std::size_t t = (n - 1) / 3, np = 3 * t + 1;
  int64_t failed = 0;
  for (uint64_t j = 0; j < N; ++j) {
    const auto sh = readVec<T>(shares + j * n * bs, n);
    std::memset(f_out + j * np * bs, 0, np * bs);
    std::memset(e_out + j * (t + 1) * bs, 0, (t + 1) * bs);
    try {
      const auto r = alphas == nullptr
                         ? scl::ss::shamirRecoverC(sh)
                         : scl::ss::shamirRecoverC(sh, readVec<T>(alphas, n));
      for (std::size_t k = 0; k <= r.f.degree() && k < np; ++k) {
        r.f[k].write(f_out + (j * np + k) * bs);
      }
      for (std::size_t k = 0; k <= r.err.degree() && k <= t; ++k) {
        r.err[k].write(e_out + (j * (t + 1) + k) * bs);
      }
      status[j] = 0;
    } catch (const std::logic_error&) {
      status[j] = 1;
      ++failed;
    }
  }
  return failed;
}

template <typename T>
void recoverP(const unsigned char* shares, uint64_t N, uint64_t n,
              const unsigned char* alphas, const unsigned char* x,
              unsigned char* out) {
  const std::size_t bs = T::byteSize();
  for (uint64_t j = 0; j < N; ++j) {
    const auto sh = readVec<T>(shares + j * n * bs, n);
    T r;
    if (alphas == nullptr) {
      r = scl::ss::shamirRecoverP(sh);
    } else {
      r = scl::ss::shamirRecoverP(sh, readVec<T>(alphas, n), T::read(x));
    }
    r.write(out + j * bs);
  }
}

// returns number of secrets for which the reference threw
// "error detected during recovery"; -1 if it threw "not enough shares...".
template <typename T>
int64_t recoverD(const unsigned char* shares, uint64_t N, uint64_t n_given,
                 uint64_t t, int use_custom, const unsigned char* alphas,
                 uint64_t n_alphas, uint64_t d, const unsigned char* x,
                 unsigned char* out, unsigned char* err) {
  const std::size_t bs = T::byteSize();
  int64_t n_err = 0;
  for (uint64_t j = 0; j < N; ++j) {
    const auto sh = readVec<T>(shares + j * n_given * bs, n_given);
    err[j] = 0;
    try {
      T r;
      if (!use_custom) {
        r = scl::ss::shamirRecoverD(sh, t);
      } else {
        r = scl::ss::shamirRecoverD(sh, readVec<T>(alphas, n_alphas), t, d,
                                    T::read(x));
      }
      r.write(out + j * bs);
    } catch (const std::logic_error& e) {
      if (std::string(e.what()) == "not enough shares provided to detect errors") {
        return -1;
      }
      err[j] = 1;
      std::memset(out + j * bs, 0, bs);
      n_err++;
    }
  }
  return n_err;
}

template <typename T>
void lagrange(const unsigned char* nodes, uint64_t n, const unsigned char* x,
              unsigned char* out) {
  const auto lb =
      scl::math::computeLagrangeBasis(readVec<T>(nodes, n), T::read(x));
  writeVec(lb, out);
}

// op: 0 add, 1 subtract, 2 multiplyEntryWise, 3 scalarMultiply (b[0] is the
// scalar), 4 dot (out = 1 element), 5 sum (out = 1 element; b unused)
template <typename T>
int vecOp(int op, const unsigned char* a, const unsigned char* b, uint64_t n,
          unsigned char* out) {
  const auto va = readVec<T>(a, n);
  switch (op) {
    case 0:
      writeVec(va.add(readVec<T>(b, n)), out);
      return 0;
    case 1:
      writeVec(va.subtract(readVec<T>(b, n)), out);
      return 0;
    case 2:
      writeVec(va.multiplyEntryWise(readVec<T>(b, n)), out);
      return 0;
    case 3:
      writeVec(va.scalarMultiply(T::read(b)), out);
      return 0;
    case 4:
      va.dot(readVec<T>(b, n)).write(out);
      return 0;
    case 5:
      va.sum().write(out);
      return 0;
    default:
      return -1;
  }
}

// Beaver-style multiply-add as the reference's Vector API spells it:
// z = e.multiplyEntryWise(b).add(d.multiplyEntryWise(a)).add(c).add(e.multiplyEntryWise(d))
// (scalar form: test/scl/protocol/beaver.h:57-61)
template <typename T>
void beaver(const unsigned char* e, const unsigned char* b,
            const unsigned char* d, const unsigned char* a,
            const unsigned char* c, uint64_t n, unsigned char* z) {
  const auto ve = readVec<T>(e, n), vb = readVec<T>(b, n), vd = readVec<T>(d, n),
             va = readVec<T>(a, n), vc = readVec<T>(c, n);
  writeVec(ve.multiplyEntryWise(vb)
               .add(vd.multiplyEntryWise(va))
               .add(vc)
               .add(ve.multiplyEntryWise(vd)),
           z);
}

template <typename T>
void matvec(const unsigned char* A, uint64_t rows, uint64_t cols,
            const unsigned char* x, unsigned char* y) {
  const auto av = readVec<T>(A, rows * cols);
  const auto m =
      scl::math::Matrix<T>::fromVector(rows, cols, av.toStlVector());
  writeVec(m.multiply(readVec<T>(x, cols)), y);
}

// Matrix::multiply(Matrix), matrix.h:476-495
template <typename T>
void matmul(const unsigned char* A, uint64_t rows, uint64_t inner,
            const unsigned char* B, uint64_t cols, unsigned char* C) {
  const auto a = scl::math::Matrix<T>::fromVector(
      rows, inner, readVec<T>(A, rows * inner).toStlVector());
  const auto b = scl::math::Matrix<T>::fromVector(
      inner, cols, readVec<T>(B, inner * cols).toStlVector());
  const auto c = a.multiply(b);
  for (uint64_t i = 0; i < rows; ++i) {
    for (uint64_t j = 0; j < cols; ++j) {
      c(i, j).write(C + (i * cols + j) * T::byteSize());
    }
  }
}

template <typename T>
void vandermonde(uint64_t n, uint64_t m, unsigned char* out) {
  const auto v = scl::math::Matrix<T>::vandermonde(n, m);
  for (uint64_t i = 0; i < n; ++i) {
    for (uint64_t j = 0; j < m; ++j) {
      v(i, j).write(out + (i * m + j) * T::byteSize());
    }
  }
}

// Matrix::vandermonde(n, m, xs) (matrix.h:445-460); returns 1 where the reference throws "|xs| != number of rows"
template <typename T>
int vandermondeXs(uint64_t n, uint64_t m, const unsigned char* xs, uint64_t n_xs, unsigned char* out) {
  try {
    const auto v = scl::math::Matrix<T>::vandermonde(n, m, readVec<T>(xs, n_xs));
    for (uint64_t i = 0; i < n; ++i) {
      for (uint64_t j = 0; j < m; ++j) {
        v(i, j).write(out + (i * m + j) * T::byteSize());
      }
    }
  } catch (const std::invalid_argument&) {
    return 1;
  }
  return 0;
}

// Polynomial::evaluate (poly.h:56-64): N polynomials (coeffs [N][m], constant term first) at n points -> [N][n]
template <typename T>
void polyEvaluate(const unsigned char* coeffs, uint64_t N, uint64_t m, const unsigned char* xs, uint64_t n,
                  unsigned char* out) {
  const auto pts = readVec<T>(xs, n);
  for (uint64_t j = 0; j < N; ++j) {
    const auto p = scl::math::Polynomial<T>::create(readVec<T>(coeffs + j * m * T::byteSize(), m));
    for (uint64_t i = 0; i < n; ++i) {
      p.evaluate(pts[i]).write(out + (j * n + i) * T::byteSize());
    }
  }
}

// Matrix::transpose (matrix.h:344-355)
template <typename T>
void transposeM(const unsigned char* in, uint64_t rows, uint64_t cols, unsigned char* out) {
  const scl::math::Matrix<T> a = scl::math::Matrix<T>::fromVector(rows, cols, readVec<T>(in, rows * cols).toStlVector());
  const auto tr = a.transpose();
  for (uint64_t i = 0; i < cols; ++i) {
    for (uint64_t j = 0; j < rows; ++j) {
      tr(i, j).write(out + (i * rows + j) * T::byteSize());
    }
  }
}

// op: 0 add 1 sub 2 mul 3 negate(a) 4 inverse(a) 5 divide
template <typename T>
int scalarOp(int op, const unsigned char* a, const unsigned char* b,
             unsigned char* out) {
  const T x = T::read(a);
  const T y = b ? T::read(b) : T{};
  try {
    switch (op) {
      case 0: (x + y).write(out); return 0;
      case 1: (x - y).write(out); return 0;
      case 2: (x * y).write(out); return 0;
      case 3: x.negated().write(out); return 0;
      case 4: x.inverse().write(out); return 0;
      case 5: (x / y).write(out); return 0;
      default: return -1;
    }
  } catch (const std::logic_error&) {
    return -2;  // "0 not invertible modulo prime"
  }
}

// CPU baseline: the reference's verbatim call sequence, per secret:
//   shares = shamirSecretShare(secret, t, n, prg);  out = shamirRecoverP(shares)
//   (or shamirRecoverD(shares, t) when detect != 0)
// split over `threads` contiguous chunks, one PRG per thread fast-forwarded to
// its chunk (SCL itself is single-threaded; threads > 1 is "all host cores").
// Returns elapsed seconds (wall clock around the worker threads, PRG skip
// excluded); xor_out receives an xor of all outputs so nothing is elided.
template <typename T>
double benchShareRecover(uint64_t N, uint64_t t, uint64_t n, int detect,
                         int threads, uint64_t blocks_per_secret,
                         unsigned char* check_out /* T::byteSize() bytes */) {
  std::vector<std::thread> pool;
  std::vector<T> acc(threads);
  std::vector<PRG> prgs;
  auto sprg = PRG::create("secrets");
  const auto secrets = scl::math::Vector<T>::random(N, sprg);
  for (int w = 0; w < threads; ++w) {
    const uint64_t lo = N * w / threads;
    prgs.emplace_back(makePrg((const unsigned char*)"shamir bench", 12,
                              lo * blocks_per_secret));
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (int w = 0; w < threads; ++w) {
    pool.emplace_back([&, w]() {
      const uint64_t lo = N * w / threads, hi = N * (w + 1) / threads;
      T a;
      for (uint64_t j = lo; j < hi; ++j) {
        const auto sh = scl::ss::shamirSecretShare(secrets[j], t, n, prgs[w]);
        a += detect ? scl::ss::shamirRecoverD(sh, t)
                    : scl::ss::shamirRecoverP(sh);
      }
      acc[w] = a;
    });
  }
  for (auto& th : pool) {
    th.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  T total;
  for (const auto& a : acc) {
    total += a;
  }
  // sum of recovered secrets must equal sum of the inputs
  if (!(total == secrets.sum())) {
    return -1.0;
  }
  total.write(check_out);
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

#define SCLREF_INSTANTIATE(SUF, T)                                             \
  void sclref_##SUF##_vector_random(const unsigned char* seed,                 \
                                    uint64_t seed_len, uint64_t skip,          \
                                    uint64_t n, unsigned char* out) {          \
    vectorRandom<T>(seed, seed_len, skip, n, out);                             \
  }                                                                            \
  void sclref_##SUF##_ff_random(const unsigned char* seed, uint64_t seed_len,  \
                                uint64_t skip, uint64_t n,                     \
                                unsigned char* out) {                          \
    ffRandom<T>(seed, seed_len, skip, n, out);                                 \
  }                                                                            \
  void sclref_##SUF##_shamir_share(                                            \
      const unsigned char* secrets, uint64_t N, uint64_t t, uint64_t n,        \
      const unsigned char* seed, uint64_t seed_len, uint64_t skip,             \
      unsigned char* shares) {                                                 \
    shamirShare<T>(secrets, N, t, n, seed, seed_len, skip, shares);            \
  }                                                                            \
  int sclref_##SUF##_shamir_share_array(                                       \
      const unsigned char* secrets, uint64_t N, uint64_t W, uint64_t t,        \
      uint64_t n, const unsigned char* seed, uint64_t seed_len, uint64_t skip, \
      unsigned char* shares) {                                                 \
    return shamirShareArray<T>(secrets, N, W, t, n, seed, seed_len, skip,      \
                               shares);                                        \
  }                                                                            \
  int sclref_##SUF##_recover_p_array(const unsigned char* shares, uint64_t N,  \
                                     uint64_t W, uint64_t n,                   \
                                     unsigned char* out) {                     \
    return recoverPArray<T>(shares, N, W, n, out);                             \
  }                                                                            \
  void sclref_##SUF##_hyper_invertible(uint64_t n, uint64_t m,                 \
                                       unsigned char* out) {                   \
    hyperInvertible<T>(n, m, out);                                             \
  }                                                                            \
  int64_t sclref_##SUF##_recover_c(const unsigned char* shares, uint64_t N,    \
                                   uint64_t n, const unsigned char* alphas,    \
                                   unsigned char* f_out, unsigned char* e_out, \
                                   unsigned char* status) {                    \
    return recoverC<T>(shares, N, n, alphas, f_out, e_out, status);            \
  }                                                                            \
  void sclref_##SUF##_additive_share(                                          \
      const unsigned char* secrets, uint64_t N, uint64_t n,                    \
      const unsigned char* seed, uint64_t seed_len, uint64_t skip,             \
      unsigned char* shares) {                                                 \
    additiveShare<T>(secrets, N, n, seed, seed_len, skip, shares);             \
  }                                                                            \
  void sclref_##SUF##_additive_recover(const unsigned char* shares,            \
                                       uint64_t N, uint64_t n,                 \
                                       unsigned char* out) {                   \
    additiveRecover<T>(shares, N, n, out);                                     \
  }                                                                            \
  void sclref_##SUF##_recover_p(const unsigned char* shares, uint64_t N,       \
                                uint64_t n, const unsigned char* alphas,       \
                                const unsigned char* x, unsigned char* out) {  \
    recoverP<T>(shares, N, n, alphas, x, out);                                 \
  }                                                                            \
  int64_t sclref_##SUF##_recover_d(                                            \
      const unsigned char* shares, uint64_t N, uint64_t n_given, uint64_t t,   \
      int use_custom, const unsigned char* alphas, uint64_t n_alphas,          \
      uint64_t d, const unsigned char* x, unsigned char* out,                  \
      unsigned char* err) {                                                    \
    return recoverD<T>(shares, N, n_given, t, use_custom, alphas, n_alphas, d, \
                       x, out, err);                                           \
  }                                                                            \
  void sclref_##SUF##_lagrange(const unsigned char* nodes, uint64_t n,         \
                               const unsigned char* x, unsigned char* out) {   \
    lagrange<T>(nodes, n, x, out);                                             \
  }                                                                            \
  int sclref_##SUF##_vec_op(int op, const unsigned char* a,                    \
                            const unsigned char* b, uint64_t n,                \
                            unsigned char* out) {                              \
    return vecOp<T>(op, a, b, n, out);                                         \
  }                                                                            \
  void sclref_##SUF##_beaver(const unsigned char* e, const unsigned char* b,   \
                             const unsigned char* d, const unsigned char* a,   \
                             const unsigned char* c, uint64_t n,               \
                             unsigned char* z) {                               \
    beaver<T>(e, b, d, a, c, n, z);                                            \
  }                                                                            \
  void sclref_##SUF##_matvec(const unsigned char* A, uint64_t rows,            \
                             uint64_t cols, const unsigned char* x,            \
                             unsigned char* y) {                               \
    matvec<T>(A, rows, cols, x, y);                                            \
  }                                                                            \
  void sclref_##SUF##_matmul(const unsigned char* A, uint64_t rows,            \
                             uint64_t inner, const unsigned char* B,           \
                             uint64_t cols, unsigned char* C) {                \
    matmul<T>(A, rows, inner, B, cols, C);                                     \
  }                                                                            \
  void sclref_##SUF##_vandermonde(uint64_t n, uint64_t m,                      \
                                  unsigned char* out) {                        \
    vandermonde<T>(n, m, out);                                                 \
  }                                                                            \
  int sclref_##SUF##_vandermonde_xs(uint64_t n, uint64_t m,                    \
                                    const unsigned char* xs, uint64_t n_xs,    \
                                    unsigned char* out) {                      \
    return vandermondeXs<T>(n, m, xs, n_xs, out);                              \
  }                                                                            \
  void sclref_##SUF##_poly_evaluate(const unsigned char* coeffs, uint64_t N,   \
                                    uint64_t m, const unsigned char* xs,       \
                                    uint64_t n, unsigned char* out) {          \
    polyEvaluate<T>(coeffs, N, m, xs, n, out);                                 \
  }                                                                            \
  void sclref_##SUF##_transpose(const unsigned char* in, uint64_t rows,        \
                                uint64_t cols, unsigned char* out) {           \
    transposeM<T>(in, rows, cols, out);                                        \
  }                                                                            \
  int sclref_##SUF##_scalar_op(int op, const unsigned char* a,                 \
                               const unsigned char* b, unsigned char* out) {   \
    return scalarOp<T>(op, a, b, out);                                         \
  }                                                                            \
  double sclref_##SUF##_bench_share_recover(                                   \
      uint64_t N, uint64_t t, uint64_t n, int detect, int threads,             \
      uint64_t blocks_per_secret, unsigned char* check_out) {                  \
    return benchShareRecover<T>(N, t, n, detect, threads, blocks_per_secret,   \
                                check_out);                                    \
  }

extern "C" {

// raw keystream: util::PRG::next after discarding `skip` blocks
void sclref_prg_next(const unsigned char* seed, uint64_t seed_len,
                     uint64_t skip, uint64_t n, unsigned char* out) {
  PRG prg = makePrg(seed, seed_len, skip);
  prg.next(out, n);
}

SCLREF_INSTANTIATE(fp61, Fp61)
SCLREF_INSTANTIATE(fp127, Fp127)

const char* sclref_version() {
  return "scl-0.1.0-unmodified";
}

}  // extern "C"
