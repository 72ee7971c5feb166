// tests/cpp/test_shim.cc -- TEST: the C++ host mirror (include/sclgpu_scl.hpp) used
// exactly as an SCL user would, on SCL's own types, checked in-process against the
// reference's own CPU functions (the unmodified reference sources are compiled into
// this test binary, like oracle/_ref; see tests/cpp/Makefile).  Bit-exact or abort.
//
// The cases follow the reference's own tests: test/scl/ss/test_shamir.cc:34-109
// (share -> recoverP, recoverD ok / tampered -> "error detected during recovery",
// custom alphas), test/scl/math/test_vector.cc, test_matrix.cc:175-222 (mat-vec),
// test/scl/util/test_prg.cc (determinism of the stream after a draw).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "scl/math/fp.h"
#include "scl/math/matrix.h"
#include "scl/math/vector.h"
#include "scl/ss/additive.h"
#include "scl/ss/shamir.h"
#include "scl/util/prg.h"
#include "sclgpu_scl.hpp"

using scl::util::PRG;
namespace ss = scl::ss;
namespace math = scl::math;

static int g_checks = 0;
#define REQUIRE(cond)                                                          \
  do {                                                                         \
    ++g_checks;                                                                \
    if (!(cond)) {                                                             \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
      std::exit(1);                                                            \
    }                                                                          \
  } while (0)

template <class F>
static bool throwsLogic(F&& f, const std::string& what) {
  try {
    f();
  } catch (const std::logic_error& e) {
    return what == e.what();
  } catch (...) {
    return false;
  }
  return false;
}
template <class F>
static bool throwsInvalid(F&& f, const std::string& what) {
  try {
    f();
  } catch (const std::invalid_argument& e) {
    return what == e.what();
  } catch (...) {
    return false;
  }
  return false;
}

int main() {
  using Fp61 = math::Fp<61>;
  using Fp127 = math::Fp<127>;
  sclgpu::Context ctx(0);  // throws without a B200: no CPU fallback

  // ---- test_shamir.cc:34-40 in batch form: share(123, t=3, n=4) -> recoverP
  {
    PRG a = PRG::create("shamir passive"), b = PRG::create("shamir passive");
    math::Vector<Fp61> secret = {Fp61(123), Fp61(124)};
    const auto m = sclgpu::shamirSecretShare(ctx, secret, 2, 5, b);
    const auto s0 = ss::shamirSecretShare(Fp61(123), 2, 5, a);
    const auto s1 = ss::shamirSecretShare(Fp61(124), 2, 5, a);
    for (int i = 0; i < 5; ++i) REQUIRE(m(0, i) == s0[i] && m(1, i) == s1[i]);
    REQUIRE(m(0, 0).toString() == "bc6378a9608f2f4");  // SURVEY 8c golden vector
    REQUIRE(sclgpu::shamirRecoverP(ctx, m)[0] == Fp61(123));
  }

  // ---- batches, both fields, PRG state, recoverP
  for (auto [N, t, n] : {std::tuple<std::size_t, std::size_t, std::size_t>{3000, 2, 5}, {1500, 15, 32}, {1, 0, 1}, {257, 7, 16}}) {
    PRG sprg = PRG::create("secrets");
    const auto secrets = math::Vector<Fp61>::random(N, sprg);
    PRG cpu = PRG::create("shamir bench"), gpu = PRG::create("shamir bench");
    (void)cpu.next(48);
    (void)gpu.next(48);
    const auto got = sclgpu::shamirSecretShare(ctx, secrets, t, n, gpu);
    REQUIRE(got.rows() == N && got.cols() == n);
    for (std::size_t j = 0; j < N; ++j) {
      const auto want = ss::shamirSecretShare(secrets[j], t, n, cpu);
      for (std::size_t i = 0; i < n; ++i) REQUIRE(got(j, i) == want[i]);
    }
    REQUIRE(cpu.next(40) == gpu.next(40));
    const auto rec = sclgpu::shamirRecoverP(ctx, got);
    for (std::size_t j = 0; j < N; ++j) REQUIRE(rec[j] == secrets[j]);
    // test_shamir.cc:42-66: reconstruct at x = 27 from the same nodes
    const auto alphas = math::Vector<Fp61>::range(1, n + 1);
    const auto at27 = sclgpu::shamirRecoverP(ctx, got, alphas, Fp61(27));
    for (std::size_t j = 0; j < N; j += 97) {
      std::vector<Fp61> row(n);
      for (std::size_t i = 0; i < n; ++i) row[i] = got(j, i);
      REQUIRE(at27[j] == ss::shamirRecoverP(math::Vector<Fp61>(row), alphas, Fp61(27)));
    }
  }
  {
    const std::size_t N = 600, t = 7, n = 16;
    PRG sprg = PRG::create("secrets127");
    const auto secrets = math::Vector<Fp127>::random(N, sprg);
    PRG cpu = PRG::create("m127"), gpu = PRG::create("m127");
    auto got = sclgpu::shamirSecretShare(ctx, secrets, t, n, gpu);
    for (std::size_t j = 0; j < N; ++j) {
      const auto want = ss::shamirSecretShare(secrets[j], t, n, cpu);
      for (std::size_t i = 0; i < n; ++i) REQUIRE(got(j, i) == want[i]);
    }
    REQUIRE(cpu.next(16) == gpu.next(16));
    // ---- test_shamir.cc:68-79: recoverD ok, then a tampered share -> throws
    const auto ok = sclgpu::shamirRecoverD(ctx, got, t);
    for (std::size_t j = 0; j < N; ++j) REQUIRE(ok[j] == secrets[j]);
    got(5, 14) = got(5, 14) + Fp127(1);   // index 2t: never checked by the reference (shamir.h:129)
    got(6, 15) = Fp127(4);                // beyond 2t: never read
    const auto still_ok = sclgpu::shamirRecoverD(ctx, got, t);
    REQUIRE(still_ok[5] == secrets[5] && still_ok[6] == secrets[6]);
    got(9, 2) = Fp127(4);                 // test_shamir.cc:76 "shares[2] = 4"
    got(11, 13) = got(11, 13) + Fp127(1);
    REQUIRE(throwsLogic([&] { (void)sclgpu::shamirRecoverD(ctx, got, t); }, "error detected during recovery"));
    std::vector<std::uint8_t> flags;
    const auto out = sclgpu::shamirRecoverD(ctx, got, t, &flags);
    for (std::size_t j = 0; j < N; ++j) {
      std::vector<Fp127> row(n);
      for (std::size_t i = 0; i < n; ++i) row[i] = got(j, i);
      bool threw = false;
      Fp127 want;
      try {
        want = ss::shamirRecoverD(math::Vector<Fp127>(row), t);
      } catch (const std::logic_error&) {
        threw = true;
      }
      REQUIRE(threw == (flags[j] != 0));
      REQUIRE(threw == (j == 9 || j == 11));
      if (!threw) REQUIRE(out[j] == want);
    }
    // not enough shares (shamir.h:122-124)
    math::Matrix<Fp127> few(3, 2 * t - 1);
    REQUIRE(throwsLogic([&] { (void)sclgpu::shamirRecoverD(ctx, few, t); },
                        "not enough shares provided to detect errors"));
  }

  // ---- test_shamir.cc:111-142: shamirRecoverC corrects <= t errors; beyond that it throws
  {
    const std::size_t N = 300, n = 10, t = 3;
    PRG sprg = PRG::create("secrets");
    const auto secrets = math::Vector<Fp61>::random(N, sprg);
    PRG gpu = PRG::create("rc");
    auto shares = sclgpu::shamirSecretShare(ctx, secrets, t, n, gpu);
    for (std::size_t j = 0; j < N; ++j)
      for (std::size_t k = 0; k < j % (t + 2); ++k) shares(j, (3 * k + j) % n) = shares(j, (3 * k + j) % n) + Fp61(j + 1);
    std::vector<std::uint8_t> status;
    const auto got = sclgpu::shamirRecoverC(ctx, shares, (const math::Vector<Fp61>*)nullptr, &status);
    bool any_failed = false;
    for (std::size_t j = 0; j < N; ++j) {
      std::vector<Fp61> row(n);
      for (std::size_t i = 0; i < n; ++i) row[i] = shares(j, i);
      bool threw = false;
      ss::ErrorCorrectedSecret<Fp61> want;
      try {
        want = ss::shamirRecoverC(math::Vector<Fp61>(row));
      } catch (const std::logic_error& e) {
        threw = std::string(e.what()) == "could not correct shares";
      }
      REQUIRE(threw == (status[j] != 0));
      any_failed = any_failed || threw;
      if (!threw) {
        REQUIRE(got[j].f.coefficients().equals(want.f.coefficients()));
        REQUIRE(got[j].err.coefficients().equals(want.err.coefficients()));
        if (j % (t + 2) <= t) REQUIRE(got[j].f.evaluate(Fp61(0)) == secrets[j]);
      }
    }
    REQUIRE(any_failed);
    REQUIRE(throwsLogic([&] { (void)sclgpu::shamirRecoverC(ctx, shares); }, "could not correct shares"));
  }

  // ---- per-party packets: what the dealer would build with Packet::write(Vector) on the CPU
  {
    const std::size_t N = 2500, t = 3, n = 7;
    PRG sprg = PRG::create("secrets");
    const auto secrets = math::Vector<Fp61>::random(N, sprg);
    PRG cpu = PRG::create("packets"), gpu = PRG::create("packets");
    auto packets = sclgpu::shamirSharePackets(ctx, secrets, t, n, gpu);
    REQUIRE(packets.size() == n);
    std::vector<std::vector<Fp61>> cols(n, std::vector<Fp61>(N));
    for (std::size_t j = 0; j < N; ++j) {
      const auto sh = ss::shamirSecretShare(secrets[j], t, n, cpu);
      for (std::size_t i = 0; i < n; ++i) cols[i][j] = sh[i];
    }
    REQUIRE(cpu.next(16) == gpu.next(16));
    for (std::size_t i = 0; i < n; ++i) {
      scl::net::Packet want;
      want.write(math::Vector<Fp61>(cols[i]));                 // SCL's own serializer
      REQUIRE(want == packets[i]);                             // byte-identical packet
    }
    REQUIRE(sclgpu::shamirRecoverP<Fp61>(ctx, packets).equals(secrets));
    REQUIRE(packets[2].read<math::Vector<Fp61>>().equals(math::Vector<Fp61>(cols[2])));  // and SCL can read it

    // ---- packets from OTHER parties may carry non-canonical words (any 8 bytes is a legal wire element:
    // Serializer<Vector<FF>>::read -> FF::read -> `% p`, vector.h:623-626, ff.h:63-67).  Overwrite words with values in
    // [p, 2^64) and compare with what SCL itself reconstructs from the same bytes.
    {
      const std::uint64_t p61 = 0x1FFFFFFFFFFFFFFFull;
      std::vector<scl::net::Packet> wire;
      for (std::size_t i = 0; i < n; ++i) {
        scl::net::Packet pk;
        pk.write(math::Vector<Fp61>(cols[i]));
        unsigned char* bytes = pk.get();
        for (std::size_t j = i; j < N; j += 5) {   // every fifth word of every packet, staggered
          std::uint64_t w;
          std::memcpy(&w, bytes + 4 + 8 * j, 8);
          std::uint64_t bad = (j % 3 == 0) ? w + p61 : (j % 3 == 1 ? w + 7 * p61 : ~0ull - (j % 97));
          if (j == i) bad = ~0ull;
          std::memcpy(bytes + 4 + 8 * j, &bad, 8);
        }
        wire.push_back(pk);
      }
      std::vector<math::Vector<Fp61>> read_back;
      for (std::size_t i = 0; i < n; ++i) {
        scl::net::Packet copy = wire[i];
        read_back.push_back(copy.read<math::Vector<Fp61>>());   // SCL's deserializer canonicalises
      }
      const auto got = sclgpu::shamirRecoverP<Fp61>(ctx, wire);
      for (std::size_t j = 0; j < N; ++j) {
        std::vector<Fp61> sh(n);
        for (std::size_t i = 0; i < n; ++i) sh[i] = read_back[i][j];
        REQUIRE(got[j] == ss::shamirRecoverP(math::Vector<Fp61>(sh)));
      }
      // a packet that announces more elements than it holds is refused before any byte of it is read
      scl::net::Packet short_pk;
      short_pk.write((std::uint32_t)N);
      std::vector<scl::net::Packet> bad_set = wire;
      bad_set[1] = short_pk;
      REQUIRE(throwsInvalid([&] { (void)sclgpu::shamirRecoverP<Fp61>(ctx, bad_set); }, "packet shorter than the Vec it announces"));
    }
  }

  // ---- Matrix::vandermonde(n, m, xs), scalarMultiply, transpose, Polynomial::evaluate at caller points
  {
    PRG prg = PRG::create("small surface");
    const auto xs = math::Vector<Fp61>::random(24, prg);
    REQUIRE(sclgpu::vandermonde<Fp61>(ctx, 24, 9, xs).equals(math::Matrix<Fp61>::vandermonde(24, 9, xs)));
    REQUIRE(throwsInvalid([&] { (void)sclgpu::vandermonde<Fp61>(ctx, 25, 9, xs); }, "|xs| != number of rows"));
    const auto A = math::Matrix<Fp61>::random(37, 19, prg);
    const Fp61 s = Fp61::random(prg);
    REQUIRE(sclgpu::scalarMultiply(ctx, A, s).equals(A.scalarMultiply(s)));
    REQUIRE(sclgpu::transpose(ctx, A).equals(A.transpose()));
    std::vector<math::Polynomial<Fp61>> polys;
    for (std::size_t j = 0; j < 300; ++j) polys.push_back(math::Polynomial<Fp61>::create(math::Vector<Fp61>::random(1 + j % 17, prg)));
    polys.push_back(math::Polynomial<Fp61>());
    const auto ys = sclgpu::evaluate(ctx, polys, xs);
    for (std::size_t j = 0; j < polys.size(); ++j)
      for (std::size_t i = 0; i < xs.size(); ++i) REQUIRE(ys(j, i) == polys[j].evaluate(xs[i]));
    const auto x127 = math::Vector<Fp127>::random(5, prg);
    REQUIRE(sclgpu::vandermonde<Fp127>(ctx, 5, 6, x127).equals(math::Matrix<Fp127>::vandermonde(5, 6, x127)));
  }

  // ---- additiveShare (test/scl/ss/test_additive.cc: shares sum to the secret), both fields, PRG state
  {
    for (std::size_t n : {1, 2, 5, 33}) {
      PRG sprg = PRG::create("secrets");
      const auto secrets = math::Vector<Fp61>::random(700, sprg);
      PRG cpu = PRG::create("additive"), gpu = PRG::create("additive");
      (void)cpu.next(3);
      (void)gpu.next(3);
      const auto got = sclgpu::additiveShare(ctx, secrets, n, gpu);
      for (std::size_t j = 0; j < secrets.size(); ++j) {
        const auto want = ss::additiveShare(secrets[j], n, cpu);
        for (std::size_t i = 0; i < n; ++i) REQUIRE(got(j, i) == want[i]);
        REQUIRE(want.sum() == secrets[j]);
      }
      REQUIRE(cpu.next(24) == gpu.next(24));
      REQUIRE(sclgpu::additiveReconstruct(ctx, got).equals(secrets));
    }
    PRG sprg = PRG::create("secrets127");
    const auto s127 = math::Vector<Fp127>::random(99, sprg);
    PRG cpu = PRG::create("a127"), gpu = PRG::create("a127");
    const auto got = sclgpu::additiveShare(ctx, s127, 4, gpu);
    for (std::size_t j = 0; j < 99; ++j) {
      const auto want = ss::additiveShare(s127[j], 4, cpu);
      for (std::size_t i = 0; i < 4; ++i) REQUIRE(got(j, i) == want[i]);
    }
    REQUIRE(cpu.next(16) == gpu.next(16));
    REQUIRE(sclgpu::additiveReconstruct(ctx, got).equals(s127));
  }

  // ---- Vector::random, entrywise ops, dot, sum, Beaver combination, mat-vec
  {
    const std::size_t n = 10001;
    PRG a = PRG::create("vec"), b = PRG::create("vec");
    (void)a.next(5);  // a 5-byte draw still consumes one whole block (prg.cc:129-133)
    (void)b.next(5);
    const auto v = sclgpu::randomVector<Fp61>(ctx, n, b);
    REQUIRE(v.equals(math::Vector<Fp61>::random(n, a)));
    const auto w = sclgpu::randomVector<Fp61>(ctx, n, b);
    REQUIRE(w.equals(math::Vector<Fp61>::random(n, a)));
    const auto v127 = sclgpu::randomVector<Fp127>(ctx, 333, b);
    REQUIRE(v127.equals(math::Vector<Fp127>::random(333, a)));
    REQUIRE(sclgpu::add(ctx, v, w).equals(v.add(w)));
    REQUIRE(sclgpu::subtract(ctx, v, w).equals(v.subtract(w)));
    REQUIRE(sclgpu::multiplyEntryWise(ctx, v, w).equals(v.multiplyEntryWise(w)));
    REQUIRE(sclgpu::scalarMultiply(ctx, v, w[3]).equals(v.scalarMultiply(w[3])));
    REQUIRE(sclgpu::dot(ctx, v, w) == v.dot(w));
    REQUIRE(sclgpu::sum(ctx, v) == v.sum());
    REQUIRE(sclgpu::equals(ctx, v, v) && !sclgpu::equals(ctx, v, w) && !sclgpu::equals(ctx, v, math::Vector<Fp61>(3)));
    const auto e = v, bb = w, d = sclgpu::add(ctx, v, v), aa = sclgpu::multiplyEntryWise(ctx, w, w), c = sclgpu::subtract(ctx, w, v);
    const auto z = e.multiplyEntryWise(bb).add(d.multiplyEntryWise(aa)).add(c).add(e.multiplyEntryWise(d));
    REQUIRE(sclgpu::beaverCombine(ctx, e, bb, d, aa, c).equals(z));
    REQUIRE(throwsInvalid([&] { (void)sclgpu::add(ctx, v, v127.size() ? math::Vector<Fp61>(3) : v); }, "Vec sizes mismatch"));

    PRG ma = PRG::create("mat A"), xa = PRG::create("vec x");
    const auto A = math::Matrix<Fp61>::random(64, 64, ma);
    const auto x = math::Vector<Fp61>::random(64, xa);
    const auto y = sclgpu::multiply(ctx, A, x);
    REQUIRE(y.equals(A.multiply(x)));
    REQUIRE(y[0].toString() == "1172bf06cc5d2e8b");  // SURVEY 8c golden vector
    REQUIRE(throwsInvalid([&] { (void)sclgpu::multiply(ctx, A, v); }, "matmul: this->cols() != vec.size()"));
    // test_matrix.cc:175-222 (mat-mul) on a size that takes the tensor-core path, and a small Fp127 one
    PRG mb = PRG::create("mat B");
    const auto A2 = math::Matrix<Fp61>::random(130, 96, ma);
    const auto B2 = math::Matrix<Fp61>::random(96, 70, mb);
    REQUIRE(sclgpu::multiply(ctx, A2, B2).equals(A2.multiply(B2)));
    const auto A3 = math::Matrix<Fp127>::random(9, 5, ma);
    const auto B3 = math::Matrix<Fp127>::random(5, 11, mb);
    REQUIRE(sclgpu::multiply(ctx, A3, B3).equals(A3.multiply(B3)));
    REQUIRE(throwsInvalid([&] { (void)sclgpu::multiply(ctx, A2, A2); }, "matmul: this->cols() != that->rows()"));
    // test_matrix.cc:397-406 (HIM) -- here against the reference's matrix itself
    REQUIRE(sclgpu::hyperInvertible<Fp61>(ctx, 4, 5).equals(math::Matrix<Fp61>::hyperInvertible(4, 5)));
    REQUIRE(sclgpu::hyperInvertible<Fp127>(ctx, 7, 3).equals(math::Matrix<Fp127>::hyperInvertible(7, 3)));
    REQUIRE(throwsInvalid([&] { (void)sclgpu::hyperInvertible<Fp61>(ctx, 0, 5); }, "n or m cannot be 0"));
  }
  {  // array-valued secrets: the sharing step of pedersenSecretShare (pedersen.h:137-138), W = 2 and 3
    PRG a = PRG::create("pedersen"), b = PRG::create("pedersen");
    using A2 = math::Array<Fp61, 2>;
    using A3 = math::Array<Fp127, 3>;
    std::vector<A2> s2;
    for (int j = 0; j < 300; ++j) s2.push_back(A2{{Fp61(1000 + j), Fp61::random(a)}});
    for (int j = 0; j < 300; ++j) (void)Fp61::random(b);  // keep both PRGs at the same counter
    const auto g2 = sclgpu::shamirSecretShare(ctx, s2, 4, 9, b);
    REQUIRE(g2.size() == s2.size());
    bool same = true;
    for (std::size_t j = 0; j < s2.size(); ++j) {
      const auto want = ss::shamirSecretShare(s2[j], 4, 9, a);
      same = same && want.equals(g2[j]);
    }
    REQUIRE(same);
    REQUIRE(math::Vector<Fp61>::random(5, a).equals(math::Vector<Fp61>::random(5, b)));  // PRG state after the batch
    const auto r2 = sclgpu::shamirRecoverP(ctx, g2);
    same = true;
    for (std::size_t j = 0; j < s2.size(); ++j) same = same && r2[j] == s2[j] && r2[j] == ss::shamirRecoverP(g2[j]);
    REQUIRE(same);
    std::vector<A3> s3;
    for (int j = 0; j < 77; ++j) s3.push_back(A3::random(a));
    for (int j = 0; j < 77; ++j) (void)A3::random(b);
    const auto g3 = sclgpu::shamirSecretShare(ctx, s3, 7, 16, b);
    same = true;
    for (std::size_t j = 0; j < s3.size(); ++j) same = same && ss::shamirSecretShare(s3[j], 7, 16, a).equals(g3[j]);
    REQUIRE(same);
    const auto r3 = sclgpu::shamirRecoverP(ctx, g3);
    same = true;
    for (std::size_t j = 0; j < s3.size(); ++j) same = same && r3[j] == s3[j];
    REQUIRE(same);
    REQUIRE(math::Vector<Fp127>::random(3, a).equals(math::Vector<Fp127>::random(3, b)));
  }
  std::printf("SHIM_OK checks=%d launches=%llu\n", g_checks, (unsigned long long)ctx.launches());
  return 0;
}
