// include/sclgpu_scl.hpp -- the host side above the C ABI, in the reference's own
// language: batched overloads of SCL's hot-path functions on SCL's own types.
//
// Header-only C++20.  It is compiled in the USER's SCL build (it includes the
// reference's headers, which are not part of this repository) and calls
// libsclgpu.so through include/sclgpu.h only.  Each function names the SCL
// function it batches; semantics, argument meaning, exception types and what()
// strings are the reference's.  "Batched" always means: exactly what a loop of N
// calls of the SCL function on the same util::PRG would produce, including the
// state the PRG is left in.
//
//   scl::ss::shamirSecretShare(secret, t, n, prg)        shamir.h:52-68
//     -> sclgpu::shamirSecretShare(ctx, secrets, t, n, prg) : Matrix (N x n), row j = SCL's result for secret j
//   scl::ss::shamirRecoverP(shares[, alphas, x])         shamir.h:82-104
//     -> sclgpu::shamirRecoverP(ctx, shares[, alphas, x])   : Vector (N)
//   scl::ss::shamirRecoverD(shares, t)                   shamir.h:117-155
//     -> sclgpu::shamirRecoverD(ctx, shares, t[, flags])    : Vector (N); throws as SCL unless flags != nullptr
//   scl::ss::shamirRecoverC(shares[, alphas])            shamir.h:203-258
//     -> sclgpu::shamirRecoverC(ctx, shares[, alphas][, status]) : std::vector<ErrorCorrectedSecret<FF>>
//   packet.write(Vector of party i's shares)             net/packet.h:145-149, vector.h:596-629
//     -> sclgpu::shamirSharePackets(ctx, secrets, t, n, prg) : n net::Packet, ready for channel->send
//        sclgpu::shamirRecoverP(ctx, packets)                : Vector (N) from the n packets received
//   scl::ss::additiveShare(secret, n, prg)               additive.h:42-53
//     -> sclgpu::additiveShare(ctx, secrets, n, prg)        : Matrix (N x n); sclgpu::additiveReconstruct = row sums
//   scl::math::Vector<FF>::random(n, prg)                vector.h:508-519
//     -> sclgpu::randomVector<FF>(ctx, n, prg)
//   Vector add / subtract / multiplyEntryWise / scalarMultiply / dot / sum   vector.h:192-301
//   Matrix::multiply(Vector)                             matrix.h:498-513
//   Matrix::multiply(Matrix)                             matrix.h:476-495
//   Matrix::hyperInvertible(n, m)                        matrix.h:462-475   -> sclgpu::hyperInvertible<FF>(ctx, n, m)
//   Matrix::vandermonde(n, m, xs)                        matrix.h:445-460   -> sclgpu::vandermonde<FF>(ctx, n, m, xs)
//   Matrix::scalarMultiply / transpose                   matrix.h:325-355   -> sclgpu::scalarMultiply / transpose(ctx, M)
//   Polynomial::evaluate, N polynomials at n points      poly.h:56-64       -> sclgpu::evaluate(ctx, polys, xs)
//   scl::ss::shamirSecretShare(math::Array<FF, W>, t, n, prg)   (pedersen.h:137-138: W = 2, {secret, randomness})
//     -> sclgpu::shamirSecretShare(ctx, std::vector<math::Array<FF, W>>, t, n, prg) : N Vectors of n Arrays
//        sclgpu::shamirRecoverP(ctx, std::vector<math::Vector<math::Array<FF, W>>>)  : N Arrays
//   Beaver combination e*b + d*a + c + e*d               test/scl/protocol/beaver.h:57-61
//
// There is no CPU fallback: Context's constructor throws when no B200 is usable.
#ifndef SCLGPU_SCL_HPP
#define SCLGPU_SCL_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "scl/math/array.h"
#include "scl/math/fp.h"
#include "scl/math/matrix.h"
#include "scl/math/vector.h"
#include "scl/math/poly.h"
#include "scl/net/packet.h"
#include "scl/ss/shamir.h"
#include "scl/util/prg.h"
#include "sclgpu.h"

namespace sclgpu {

namespace detail {

// util::PRG keeps its block counter private and offers no seek (prg.h:167-169).
// A batched draw must leave the PRG where N sequential SCL calls would, so the
// counter is reached through an explicit template instantiation (access checks do
// not apply to explicit-instantiation arguments) -- the reference stays unmodified.
template <auto Member>
struct PrgCounterAccess {
  friend long& prgCounter(scl::util::PRG& prg) { return prg.*Member; }
};
long& prgCounter(scl::util::PRG& prg);
template struct PrgCounterAccess<&scl::util::PRG::m_counter>;

template <class FF>
struct Abi;  // maps an SCL field type onto the fp61 / fp127 entry points

#define SCLGPU_SCL_ABI(FIELD, SUF, BYTES_)                                                                    \
  template <>                                                                                                 \
  struct Abi<scl::math::FF<FIELD>> {                                                                          \
    static constexpr std::size_t BYTES = BYTES_;                                                              \
    static constexpr auto random = &sclgpu_##SUF##_random;                                                    \
    static constexpr auto share = &sclgpu_##SUF##_shamir_share;                                               \
    static constexpr auto recover_p = &sclgpu_##SUF##_recover_p;                                              \
    static constexpr auto recover_c = &sclgpu_##SUF##_recover_c;                                              \
    static constexpr auto share_packets = &sclgpu_##SUF##_shamir_share_packets;                               \
    static constexpr auto recover_p_packets = &sclgpu_##SUF##_recover_p_packets;                              \
    static constexpr auto additive_share = &sclgpu_##SUF##_additive_share;                                    \
    static constexpr auto additive_recover = &sclgpu_##SUF##_additive_recover;                                \
    static constexpr auto recover_d = &sclgpu_##SUF##_recover_d;                                              \
    static constexpr auto vec_add = &sclgpu_##SUF##_vec_add;                                                  \
    static constexpr auto vec_sub = &sclgpu_##SUF##_vec_sub;                                                  \
    static constexpr auto vec_mul = &sclgpu_##SUF##_vec_mul;                                                  \
    static constexpr auto vec_scale = &sclgpu_##SUF##_vec_scale;                                              \
    static constexpr auto vec_muladd = &sclgpu_##SUF##_vec_muladd;                                            \
    static constexpr auto vec_equal = &sclgpu_##SUF##_vec_equal;                                              \
    static constexpr auto dot = &sclgpu_##SUF##_dot;                                                          \
    static constexpr auto sum = &sclgpu_##SUF##_sum;                                                          \
    static constexpr auto matvec = &sclgpu_##SUF##_matvec;                                                    \
    static constexpr auto matmul = &sclgpu_##SUF##_matmul;                                                    \
    static constexpr auto share_array = &sclgpu_##SUF##_shamir_share_array;                                   \
    static constexpr auto recover_p_array = &sclgpu_##SUF##_recover_p_array;                                  \
    static constexpr auto hyper_invertible = &sclgpu_##SUF##_hyper_invertible;                                \
    static constexpr auto vandermonde_xs = &sclgpu_##SUF##_vandermonde_xs;                                    \
    static constexpr auto poly_evaluate = &sclgpu_##SUF##_poly_evaluate;                                      \
    static constexpr auto transpose = &sclgpu_##SUF##_transpose;                                              \
  }
SCLGPU_SCL_ABI(scl::math::ff::Mersenne61, fp61, 8);
SCLGPU_SCL_ABI(scl::math::ff::Mersenne127, fp127, 16);
#undef SCLGPU_SCL_ABI

// SCL elements are trivially copyable wrappers around one uint64_t / __uint128_t
// holding the canonical residue (ff.h:313-314), i.e. exactly FF::write's bytes.
template <class FF, class P>
auto* raw(P* p) {
  static_assert(sizeof(FF) == Abi<FF>::BYTES, "unexpected SCL element layout");
  if constexpr (Abi<FF>::BYTES == 8) {
    if constexpr (std::is_const_v<P>) return reinterpret_cast<const std::uint64_t*>(p);
    else return reinterpret_cast<std::uint64_t*>(p);
  } else {
    if constexpr (std::is_const_v<P>) return reinterpret_cast<const void*>(p);
    else return reinterpret_cast<void*>(p);
  }
}

inline std::uint64_t shareBlocks(std::size_t bytes, std::size_t t) { return ((t + 1) * bytes + 15) / 16; }

}  // namespace detail

// One context per device per process (calls are serialised by the caller, as SCL
// itself is single-threaded).
class Context {
 public:
  explicit Context(int device = 0) {
    if (sclgpu_init(device, &m_ctx) != SCLGPU_OK) {
      throw std::runtime_error("sclgpu: no usable sm_100 device (there is no CPU fallback)");
    }
  }
  ~Context() { sclgpu_destroy(m_ctx); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  sclgpu_ctx* get() const { return m_ctx; }
  std::uint64_t launches() const { return sclgpu_launch_count(m_ctx); }

  // SCLGPU_EINVAL -> std::invalid_argument, ELOGIC/EDETECT/ECORRECT -> std::logic_error,
  // with the reference's own message (sclgpu_last_error).
  void check(int rc) const {
    if (rc == SCLGPU_OK) return;
    const std::string msg = sclgpu_last_error(m_ctx);
    if (rc == SCLGPU_EINVAL) throw std::invalid_argument(msg);
    if (rc == SCLGPU_ELOGIC || rc == SCLGPU_EDETECT || rc == SCLGPU_ECORRECT) throw std::logic_error(msg);
    throw std::runtime_error(std::string("sclgpu: ") + sclgpu_strerror(rc) + ": " + msg);
  }

 private:
  sclgpu_ctx* m_ctx = nullptr;
};

// ---- Vector<FF>::random(n, prg), vector.h:508-519 (ONE next() of n*byteSize bytes)
template <class FF>
scl::math::Vector<FF> randomVector(Context& ctx, std::size_t n, scl::util::PRG& prg) {
  using A = detail::Abi<FF>;
  std::vector<FF> v(n);
  long& ctr = detail::prgCounter(prg);
  const auto seed = prg.Seed();
  ctx.check(A::random(ctx.get(), seed.data(), (std::uint64_t)ctr, n, detail::raw<FF>(v.data())));
  ctr += (long)((n * A::BYTES + 15) / 16);
  return scl::math::Vector<FF>(std::move(v));
}

// ---- shamirSecretShare on every element of `secrets`, shamir.h:52-68
template <class FF>
scl::math::Matrix<FF> shamirSecretShare(Context& ctx, const scl::math::Vector<FF>& secrets, std::size_t t,
                                        std::size_t n, scl::util::PRG& prg) {
  using A = detail::Abi<FF>;
  const std::size_t N = secrets.size();
  if (N == 0 || n == 0) {
    detail::prgCounter(prg) += (long)(N * detail::shareBlocks(A::BYTES, t));
    return scl::math::Matrix<FF>();
  }
  scl::math::Matrix<FF> shares(N, n);
  long& ctr = detail::prgCounter(prg);
  const auto seed = prg.Seed();
  ctx.check(A::share(ctx.get(), detail::raw<FF>(secrets.toStlVector().data()), N, (std::uint32_t)t,
                     (std::uint32_t)n, seed.data(), (std::uint64_t)ctr, detail::raw<FF>(&shares(0, 0))));
  ctr += (long)(N * detail::shareBlocks(A::BYTES, t));
  return shares;
}

// ---- shamirSecretShare straight into per-party packets: packets[i] holds what
// `Packet p; p.write(Vector<FF>(shares of party i))` holds (u32 count + FF::write bytes), so
// `co_await channel_i->send(packets[i])` is the next line of the dealer's protocol.
template <class FF>
std::vector<scl::net::Packet> shamirSharePackets(Context& ctx, const scl::math::Vector<FF>& secrets, std::size_t t,
                                                 std::size_t n, scl::util::PRG& prg) {
  using A = detail::Abi<FF>;
  const std::size_t N = secrets.size();
  const std::size_t bytes = (std::size_t)sclgpu_packet_bytes((std::uint32_t)A::BYTES, N);
  std::vector<scl::net::Packet> packets;
  packets.reserve(n);
  std::vector<std::uint8_t*> bufs(n);
  for (std::size_t i = 0; i < n; ++i) {
    packets.emplace_back(bytes);
    bufs[i] = packets.back().get();
  }
  long& ctr = detail::prgCounter(prg);
  const auto seed = prg.Seed();
  ctx.check(A::share_packets(ctx.get(), detail::raw<FF>(secrets.toStlVector().data()), N, (std::uint32_t)t,
                             (std::uint32_t)n, seed.data(), (std::uint64_t)ctr, bufs.data()));
  ctr += (long)(N * detail::shareBlocks(A::BYTES, t));
  for (auto& p : packets) p.setWritePtr((std::ptrdiff_t)bytes);
  return packets;
}

// ---- shamirRecoverP(shares) for all N secrets from the n packets received (packet i = the Vector
// party i sent); the packets' read pointers are not moved.
template <class FF>
scl::math::Vector<FF> shamirRecoverP(Context& ctx, const std::vector<scl::net::Packet>& packets) {
  using A = detail::Abi<FF>;
  if (packets.empty()) return scl::math::Vector<FF>();
  if (packets[0].size() < sizeof(std::uint32_t)) throw std::invalid_argument("packet without an element count");
  std::uint32_t count = 0;
  std::memcpy(&count, packets[0].get(), sizeof(count));
  std::vector<const std::uint8_t*> bufs(packets.size());
  for (std::size_t i = 0; i < packets.size(); ++i) {
    // a packet from another party: it must actually hold the count + count elements it announces before a byte of it
    // is handed to the DMA engine (Serializer<Vector<FF>>::read would run off its end, vector.h:612-629)
    if ((std::size_t)packets[i].size() < sizeof(std::uint32_t) + (std::size_t)count * A::BYTES)
      throw std::invalid_argument("packet shorter than the Vec it announces");
    bufs[i] = packets[i].get();
  }
  std::vector<FF> out(count);
  ctx.check(A::recover_p_packets(ctx.get(), bufs.data(), count, (std::uint32_t)packets.size(), nullptr, nullptr,
                                 detail::raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}

// ---- additiveShare on every element of `secrets`, additive.h:42-53: n-1 FF::random draws per
// secret (one keystream block each) and secret - sum.  Reconstruction (shares.sum(), additive.h:38-39)
// per row: additiveReconstruct.
template <class FF>
scl::math::Matrix<FF> additiveShare(Context& ctx, const scl::math::Vector<FF>& secrets, std::size_t n,
                                    scl::util::PRG& prg) {
  using A = detail::Abi<FF>;
  const std::size_t N = secrets.size();
  if (n == 0) throw std::invalid_argument("additiveShare needs n >= 1");
  if (N == 0) return scl::math::Matrix<FF>();
  scl::math::Matrix<FF> shares(N, n);
  long& ctr = detail::prgCounter(prg);
  const auto seed = prg.Seed();
  ctx.check(A::additive_share(ctx.get(), detail::raw<FF>(secrets.toStlVector().data()), N, (std::uint32_t)n,
                              seed.data(), (std::uint64_t)ctr, detail::raw<FF>(&shares(0, 0))));
  ctr += (long)(N * (n - 1));
  return shares;
}
template <class FF>
scl::math::Vector<FF> additiveReconstruct(Context& ctx, const scl::math::Matrix<FF>& shares) {
  using A = detail::Abi<FF>;
  std::vector<FF> out(shares.rows());
  if (shares.rows() == 0) return scl::math::Vector<FF>(std::move(out));
  ctx.check(A::additive_recover(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(shares)(0, 0)),
                                shares.rows(), (std::uint32_t)shares.cols(), detail::raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}

// ---- shamirRecoverP(shares), shamir.h:100-104: row j of `shares` = one sharing
template <class FF>
scl::math::Vector<FF> shamirRecoverP(Context& ctx, const scl::math::Matrix<FF>& shares) {
  using A = detail::Abi<FF>;
  std::vector<FF> out(shares.rows());
  if (shares.rows() == 0) return scl::math::Vector<FF>(std::move(out));
  ctx.check(A::recover_p(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(shares)(0, 0)),
                         shares.rows(), (std::uint32_t)shares.cols(), nullptr, nullptr,
                         detail::raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}

// ---- shamirRecoverP(shares, alphas, x), shamir.h:82-87
template <class FF>
scl::math::Vector<FF> shamirRecoverP(Context& ctx, const scl::math::Matrix<FF>& shares,
                                     const scl::math::Vector<FF>& alphas, const FF& x) {
  using A = detail::Abi<FF>;
  if (alphas.size() != shares.cols()) throw std::invalid_argument("Vec sizes mismatch");  // vector.h:483
  std::vector<FF> out(shares.rows());
  if (shares.rows() == 0) return scl::math::Vector<FF>(std::move(out));
  ctx.check(A::recover_p(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(shares)(0, 0)),
                         shares.rows(), (std::uint32_t)shares.cols(),
                         detail::raw<FF>(alphas.toStlVector().data()), detail::raw<FF>(&x),
                         detail::raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}

// ---- shamirRecoverD(shares, t), shamir.h:152-155.  With flags == nullptr the call
// throws std::logic_error("error detected during recovery") if ANY sharing is
// inconsistent (what the loop over SCL's function would do at the first one).
// With flags != nullptr nothing is thrown for inconsistent sharings: (*flags)[j] = 1
// and result[j] = 0 for those.
template <class FF>
scl::math::Vector<FF> shamirRecoverD(Context& ctx, const scl::math::Matrix<FF>& shares, std::size_t t,
                                     std::vector<std::uint8_t>* flags = nullptr) {
  using A = detail::Abi<FF>;
  const std::size_t N = shares.rows();
  std::vector<FF> out(N);
  std::vector<std::uint8_t> err(N);
  std::uint64_t n_bad = 0;
  const int rc = A::recover_d(
      ctx.get(), N ? detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(shares)(0, 0)) : nullptr, N,
      (std::uint32_t)shares.cols(), (std::uint32_t)t, nullptr, 0, 0, nullptr, detail::raw<FF>(out.data()),
      err.data(), &n_bad);
  if (rc == SCLGPU_EDETECT && flags != nullptr) {
    *flags = std::move(err);
    return scl::math::Vector<FF>(std::move(out));
  }
  ctx.check(rc);
  if (flags != nullptr) *flags = std::move(err);
  return scl::math::Vector<FF>(std::move(out));
}

// ---- shamirRecoverC(shares[, alphas]), shamir.h:203-258, on every row of `shares`.  With status ==
// nullptr the call throws std::logic_error("could not correct shares") if ANY sharing cannot be
// corrected (what the loop over SCL's function does at the first one); otherwise (*status)[j] = 1 marks
// those rows and their result holds zero polynomials.
template <class FF>
std::vector<scl::ss::ErrorCorrectedSecret<FF>> shamirRecoverC(Context& ctx, const scl::math::Matrix<FF>& shares,
                                                              const scl::math::Vector<FF>* alphas = nullptr,
                                                              std::vector<std::uint8_t>* status = nullptr) {
  using A = detail::Abi<FF>;
  const std::size_t N = shares.rows(), n = shares.cols();
  if (N == 0) return {};
  if (alphas != nullptr && alphas->size() != n) throw std::invalid_argument("Vec sizes mismatch");
  const std::size_t t = (n - 1) / 3, np = 3 * t + 1;
  std::vector<FF> f(N * np), e(N * (t + 1));
  std::vector<std::uint8_t> st(N);
  std::uint64_t n_failed = 0;
  const int rc = A::recover_c(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(shares)(0, 0)), N,
                              (std::uint32_t)n, alphas ? detail::raw<FF>(alphas->toStlVector().data()) : nullptr,
                              detail::raw<FF>(f.data()), detail::raw<FF>(e.data()), st.data(), &n_failed);
  if (!(rc == SCLGPU_ECORRECT && status != nullptr)) ctx.check(rc);
  std::vector<scl::ss::ErrorCorrectedSecret<FF>> out;
  out.reserve(N);
  for (std::size_t j = 0; j < N; ++j) {
    using Vec = scl::math::Vector<FF>;
    out.push_back({scl::math::Polynomial<FF>::create(Vec(f.begin() + j * np, f.begin() + (j + 1) * np)),
                   scl::math::Polynomial<FF>::create(Vec(e.begin() + j * (t + 1), e.begin() + (j + 1) * (t + 1)))});
  }
  if (status != nullptr) *status = std::move(st);
  return out;
}

// ---- Vector entrywise operations, vector.h:192-301 ("Vec sizes mismatch", :481-485)
namespace detail {
template <class FF, class Fn>
scl::math::Vector<FF> binop(Context& ctx, Fn fn, const scl::math::Vector<FF>& a, const scl::math::Vector<FF>& b) {
  if (a.size() != b.size()) throw std::invalid_argument("Vec sizes mismatch");
  std::vector<FF> out(a.size());
  ctx.check(fn(ctx.get(), raw<FF>(a.toStlVector().data()), raw<FF>(b.toStlVector().data()), a.size(),
               raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}
}  // namespace detail

template <class FF>
scl::math::Vector<FF> add(Context& ctx, const scl::math::Vector<FF>& a, const scl::math::Vector<FF>& b) {
  return detail::binop<FF>(ctx, detail::Abi<FF>::vec_add, a, b);
}
template <class FF>
scl::math::Vector<FF> subtract(Context& ctx, const scl::math::Vector<FF>& a, const scl::math::Vector<FF>& b) {
  return detail::binop<FF>(ctx, detail::Abi<FF>::vec_sub, a, b);
}
template <class FF>
scl::math::Vector<FF> multiplyEntryWise(Context& ctx, const scl::math::Vector<FF>& a,
                                        const scl::math::Vector<FF>& b) {
  return detail::binop<FF>(ctx, detail::Abi<FF>::vec_mul, a, b);
}
template <class FF>
scl::math::Vector<FF> scalarMultiply(Context& ctx, const scl::math::Vector<FF>& a, const FF& s) {
  std::vector<FF> out(a.size());
  ctx.check(detail::Abi<FF>::vec_scale(ctx.get(), detail::raw<FF>(a.toStlVector().data()), detail::raw<FF>(&s),
                                       a.size(), detail::raw<FF>(out.data())));
  return scl::math::Vector<FF>(std::move(out));
}
template <class FF>
FF dot(Context& ctx, const scl::math::Vector<FF>& a, const scl::math::Vector<FF>& b) {
  if (a.size() != b.size()) throw std::invalid_argument("Vec sizes mismatch");
  FF out;
  ctx.check(detail::Abi<FF>::dot(ctx.get(), detail::raw<FF>(a.toStlVector().data()),
                                 detail::raw<FF>(b.toStlVector().data()), a.size(), detail::raw<FF>(&out)));
  return out;
}
// Vector::equals, vector.h:358-375
template <class FF>
bool equals(Context& ctx, const scl::math::Vector<FF>& a, const scl::math::Vector<FF>& b) {
  if (a.size() != b.size()) return false;
  int eq = 0;
  ctx.check(detail::Abi<FF>::vec_equal(ctx.get(), detail::raw<FF>(a.toStlVector().data()),
                                       detail::raw<FF>(b.toStlVector().data()), a.size(), &eq));
  return eq != 0;
}
template <class FF>
FF sum(Context& ctx, const scl::math::Vector<FF>& a) {
  FF out;
  ctx.check(detail::Abi<FF>::sum(ctx.get(), detail::raw<FF>(a.toStlVector().data()), a.size(),
                                 detail::raw<FF>(&out)));
  return out;
}
// z = e*b + d*a + c + e*d over Vectors (the Beaver combination, beaver.h:57-61)
template <class FF>
scl::math::Vector<FF> beaverCombine(Context& ctx, const scl::math::Vector<FF>& e, const scl::math::Vector<FF>& b,
                                    const scl::math::Vector<FF>& d, const scl::math::Vector<FF>& a,
                                    const scl::math::Vector<FF>& c) {
  const std::size_t n = e.size();
  if (b.size() != n || d.size() != n || a.size() != n || c.size() != n)
    throw std::invalid_argument("Vec sizes mismatch");
  std::vector<FF> z(n);
  ctx.check(detail::Abi<FF>::vec_muladd(ctx.get(), detail::raw<FF>(e.toStlVector().data()),
                                        detail::raw<FF>(b.toStlVector().data()),
                                        detail::raw<FF>(d.toStlVector().data()),
                                        detail::raw<FF>(a.toStlVector().data()),
                                        detail::raw<FF>(c.toStlVector().data()), n, detail::raw<FF>(z.data())));
  return scl::math::Vector<FF>(std::move(z));
}

// ---- Matrix::multiply(Vector), matrix.h:498-513
template <class FF>
scl::math::Vector<FF> multiply(Context& ctx, const scl::math::Matrix<FF>& A, const scl::math::Vector<FF>& x) {
  if (A.cols() != x.size()) throw std::invalid_argument("matmul: this->cols() != vec.size()");  // matrix.h:500
  std::vector<FF> y(A.rows());
  ctx.check(detail::Abi<FF>::matvec(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(A)(0, 0)),
                                    (std::uint32_t)A.rows(), (std::uint32_t)A.cols(),
                                    detail::raw<FF>(x.toStlVector().data()), detail::raw<FF>(y.data())));
  return scl::math::Vector<FF>(std::move(y));
}

// ---- Matrix::multiply(Matrix), matrix.h:476-495
template <class FF>
scl::math::Matrix<FF> multiply(Context& ctx, const scl::math::Matrix<FF>& A, const scl::math::Matrix<FF>& B) {
  if (A.cols() != B.rows()) throw std::invalid_argument("matmul: this->cols() != that->rows()");  // matrix.h:480
  scl::math::Matrix<FF> C(A.rows(), B.cols());
  ctx.check(detail::Abi<FF>::matmul(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(A)(0, 0)),
                                    (std::uint32_t)A.rows(), (std::uint32_t)A.cols(),
                                    detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(B)(0, 0)),
                                    (std::uint32_t)B.cols(), detail::raw<FF>(&C(0, 0))));
  return C;
}

// ---- Matrix::hyperInvertible(n, m), matrix.h:462-475
template <class FF>
scl::math::Matrix<FF> hyperInvertible(Context& ctx, std::size_t n, std::size_t m) {
  if (n == 0 || m == 0) throw std::invalid_argument("n or m cannot be 0");  // matrix.h:165
  scl::math::Matrix<FF> him(n, m);
  ctx.check(detail::Abi<FF>::hyper_invertible(ctx.get(), (std::uint32_t)n, (std::uint32_t)m, detail::raw<FF>(&him(0, 0))));
  return him;
}

// ---- Matrix::vandermonde(n, m, xs), matrix.h:445-460
template <class FF>
scl::math::Matrix<FF> vandermonde(Context& ctx, std::size_t n, std::size_t m, const scl::math::Vector<FF>& xs) {
  if (xs.size() != n) throw std::invalid_argument("|xs| != number of rows");
  if (n == 0 || m == 0) return scl::math::Matrix<FF>();
  scl::math::Matrix<FF> v(n, m);
  ctx.check(detail::Abi<FF>::vandermonde_xs(ctx.get(), (std::uint32_t)n, (std::uint32_t)m, detail::raw<FF>(xs.toStlVector().data()),
                                            (std::uint32_t)xs.size(), detail::raw<FF>(&v(0, 0))));
  return v;
}

// ---- Matrix::scalarMultiply (matrix.h:325-342) and Matrix::transpose (matrix.h:344-355)
template <class FF>
scl::math::Matrix<FF> scalarMultiply(Context& ctx, const scl::math::Matrix<FF>& a, const FF& s) {
  if (a.rows() == 0 || a.cols() == 0) return a;
  scl::math::Matrix<FF> out(a.rows(), a.cols());
  ctx.check(detail::Abi<FF>::vec_scale(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(a)(0, 0)), detail::raw<FF>(&s), a.rows() * a.cols(),
                                       detail::raw<FF>(&out(0, 0))));
  return out;
}
template <class FF>
scl::math::Matrix<FF> transpose(Context& ctx, const scl::math::Matrix<FF>& a) {
  if (a.rows() == 0 || a.cols() == 0) return a.transpose();
  scl::math::Matrix<FF> out(a.cols(), a.rows());
  ctx.check(detail::Abi<FF>::transpose(ctx.get(), detail::raw<FF>(&const_cast<scl::math::Matrix<FF>&>(a)(0, 0)), a.rows(), a.cols(),
                                       detail::raw<FF>(&out(0, 0))));
  return out;
}

// ---- Polynomial::evaluate (poly.h:56-64) of every polynomial at every point: result(j, i) = polys[j].evaluate(xs[i])
template <class FF>
scl::math::Matrix<FF> evaluate(Context& ctx, const std::vector<scl::math::Polynomial<FF>>& polys, const scl::math::Vector<FF>& xs) {
  const std::size_t N = polys.size(), n = xs.size();
  if (N == 0 || n == 0) return scl::math::Matrix<FF>();
  std::size_t m = 1;
  for (const auto& p : polys) m = std::max<std::size_t>(m, p.degree() + 1);
  std::vector<FF> coeffs(N * m);  // zero padded to the largest degree
  for (std::size_t j = 0; j < N; ++j)
    for (std::size_t k = 0; k <= polys[j].degree(); ++k) coeffs[j * m + k] = polys[j][k];
  scl::math::Matrix<FF> out(N, n);
  ctx.check(detail::Abi<FF>::poly_evaluate(ctx.get(), detail::raw<FF>(coeffs.data()), N, (std::uint32_t)(m - 1),
                                           detail::raw<FF>(xs.toStlVector().data()), (std::uint32_t)n, detail::raw<FF>(&out(0, 0))));
  return out;
}

// ---- shamirSecretShare on array-valued secrets (shamir.h:52-68 with T = math::Array<FF, W>; the sharing
// step of ss::pedersenSecretShare, pedersen.h:137-138).  Element j of the result is what
// scl::ss::shamirSecretShare(secrets[j], t, n, prg) returns, for the calls made in order on `prg`.
// Arrays cross the ABI as their Array::write bytes (array.h:407-411): W x FF::write.
template <class FF, std::size_t W>
std::vector<scl::math::Vector<scl::math::Array<FF, W>>> shamirSecretShare(
    Context& ctx, const std::vector<scl::math::Array<FF, W>>& secrets, std::size_t t, std::size_t n,
    scl::util::PRG& prg) {
  using A = detail::Abi<FF>;
  using Arr = scl::math::Array<FF, W>;
  const std::size_t N = secrets.size();
  const std::uint64_t blocks = sclgpu_share_array_blocks((std::uint32_t)A::BYTES, (std::uint32_t)W, (std::uint32_t)t);
  std::vector<FF> in(N * W), flat(N * n * W);
  for (std::size_t j = 0; j < N; ++j) secrets[j].write(reinterpret_cast<unsigned char*>(in.data() + j * W));
  long& ctr = detail::prgCounter(prg);
  const auto seed = prg.Seed();
  if (N != 0 && n != 0) {
    ctx.check(A::share_array(ctx.get(), detail::raw<FF>(static_cast<const FF*>(in.data())), N, (std::uint32_t)W,
                             (std::uint32_t)t, (std::uint32_t)n, seed.data(), (std::uint64_t)ctr,
                             detail::raw<FF>(flat.data())));
  }
  ctr += (long)(N * blocks);
  std::vector<scl::math::Vector<Arr>> out;
  out.reserve(N);
  for (std::size_t j = 0; j < N; ++j) {
    std::vector<Arr> row;
    row.reserve(n);
    for (std::size_t i = 0; i < n; ++i)
      row.emplace_back(Arr::read(reinterpret_cast<const unsigned char*>(flat.data() + (j * n + i) * W)));
    out.emplace_back(scl::math::Vector<Arr>(std::move(row)));
  }
  return out;
}

// ---- shamirRecoverP(shares) on Vectors of Arrays, shamir.h:100-104 (all sharings of the same size n)
template <class FF, std::size_t W>
std::vector<scl::math::Array<FF, W>> shamirRecoverP(Context& ctx,
                                                    const std::vector<scl::math::Vector<scl::math::Array<FF, W>>>& shares) {
  using A = detail::Abi<FF>;
  using Arr = scl::math::Array<FF, W>;
  const std::size_t N = shares.size();
  std::vector<Arr> out;
  if (N == 0) return out;
  const std::size_t n = shares[0].size();
  std::vector<FF> flat(N * n * W), rec(N * W);
  for (std::size_t j = 0; j < N; ++j) {
    if (shares[j].size() != n) throw std::invalid_argument("Vec sizes mismatch");  // vector.h:483
    for (std::size_t i = 0; i < n; ++i) shares[j][i].write(reinterpret_cast<unsigned char*>(flat.data() + (j * n + i) * W));
  }
  ctx.check(A::recover_p_array(ctx.get(), detail::raw<FF>(static_cast<const FF*>(flat.data())), N, (std::uint32_t)W,
                               (std::uint32_t)n, detail::raw<FF>(rec.data())));
  out.reserve(N);
  for (std::size_t j = 0; j < N; ++j) out.emplace_back(Arr::read(reinterpret_cast<const unsigned char*>(rec.data() + j * W)));
  return out;
}

}  // namespace sclgpu

#endif  // SCLGPU_SCL_HPP
