// recover_c_syndrome.cuh -- shamirRecoverC (include/scl/ss/shamir.h:203-258) for sharings WITH errors, one THREAD per
// sharing, by syndrome decoding instead of the reference's (3t+1)-square eliminations.
//
// The np = 3t+1 shares of a sharing are a word of the generalised Reed-Solomon code {(f(a_0), .., f(a_{np-1})) :
// deg f <= t} (distance 2t+1).  With w_i = 1 / prod_{j != i} (a_i - a_j), the 2t syndromes
//     S_k = sum_i w_i r_i a_i^k,  k = 0 .. 2t-1,
// vanish on codewords, so S_k = sum_{i in Err} (w_i e_i) a_i^k.  Then
//   1. inversion-free Berlekamp-Massey on S_0..S_{2t-1}: shortest recurrence, length L = number of errors (if <= t);
//      reversed, it is a multiple of the locator prod_{i in Err} (x - a_i)  (no 1/a_i anywhere: a node may be 0);
//   2. its roots among the nodes = error positions (exactly L of them, else: not decodable this way);
//   3. error values: with M(x) = prod_{Err} (x - a_i), N_i = M / (x - a_i):  w_i e_i = (sum_k N_i[k] S_k) / N_i(a_i)
//      -- one Fermat inversion per sharing (Montgomery's trick over the L denominators);
//   4. f = the interpolant of the first t+1 CORRECTED shares (the (t+1) x (t+1) matrix k_recover_c_clean uses);
//   5. verification, as in k_recover_c's shortcut: f disagrees with the shares exactly at the L <= t positions found.
//      Unique decoding then makes the reference's answer f with the locator M (its e = L system is the first uniquely
//      solvable one; every e > L system has many solutions and is rejected, matrix.h:741-764).
// Anything else -- more than t errors, a recurrence whose roots are not nodes -- is handed on, compacted, to
// k_recover_c (the reference's own sequence).  Work per sharing: ~2 000 field multiplications in one thread against
// ~30 000 multiply-adds across a warp for one (3t+1)-square elimination.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace sclgpu {

static constexpr uint32_t kSynMaxT = 10;     // 3t+1 <= 31: the range of the warp-per-sharing elimination kernel
static constexpr uint32_t kSynMaxTBig = 55;  // 3t+1 <= 166: the range of the CTA-per-sharing one (recover_c_big.cuh)

// consts: [0, np) nodes a_i | [np, 2np) w_i | [2np, 3np) 1 / w_i | then the (t+1) x (t+1) coefficient matrix
// MAXT bounds the per-thread arrays (local memory: they are indexed dynamically).  MAXT = kSynMaxT serves 3t+1 <= 31,
// MAXT = kSynMaxTBig everything the CTA kernel takes; the host launches the latter with one CTA per SM so that the
// threads' arrays (7 / 15 KiB each for Fp61 / Fp127 at t = 55 / 39) stay in L2.
template <class F, uint32_t MAXT>
__global__ void __launch_bounds__(128)
k_recover_c_syndrome(const typename F::E* __restrict__ in, uint64_t stride_i, uint64_t stride_j, uint32_t t,
                     const typename F::E* __restrict__ consts, typename F::E* __restrict__ f_out,
                     typename F::E* __restrict__ e_out, uint8_t* __restrict__ status, const uint32_t* __restrict__ pending,
                     const unsigned long long* __restrict__ n_pending, uint32_t* __restrict__ pending2,
                     unsigned long long* __restrict__ n_pending2) {
  typedef typename F::E E;
  constexpr uint32_t kSynMaxT = MAXT, kSynMaxPoints = 3 * MAXT + 1;  // shadow the namespace constants: this kernel's bounds
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  E* sm = reinterpret_cast<E*>(dyn_smem);
  const uint32_t m = t + 1u, np = 3u * t + 1u, n_const = 3u * np + m * m;
  for (uint32_t i = threadIdx.x; i < n_const; i += blockDim.x) sm[i] = consts[i];
  __syncthreads();
  const E* A = sm;
  const E* W = sm + np;
  const E* WI = sm + 2u * np;
  const E* COEF = sm + 3u * np;
  const uint64_t n_work = *n_pending;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_work; q += stride) {
    const uint64_t j = pending[q];
    const E* src = in + j * stride_j;
    E r[kSynMaxPoints];
#pragma unroll 1
    for (uint32_t i = 0; i < np; ++i) r[i] = src[(uint64_t)i * stride_i];
    // ---- syndromes: node by node, the running power in a register (S is the only array touched in the inner loop)
    E S[2 * kSynMaxT];
#pragma unroll 1
    for (uint32_t k = 0; k < 2u * t; ++k) S[k] = F::zero();
#pragma unroll 1
    for (uint32_t i = 0; i < np; ++i) {
      E pw = F::mul(W[i], r[i]);
      const E a = A[i];
#pragma unroll 1
      for (uint32_t k = 0; k < 2u * t; ++k) {
        S[k] = F::add(S[k], pw);
        pw = F::mul(pw, a);
      }
    }
    // ---- inversion-free Berlekamp-Massey: C(x) <- b C(x) - d x^mm B(x)
    E C[kSynMaxT + 2], B[kSynMaxT + 2];
#pragma unroll 1
    for (uint32_t i = 0; i < kSynMaxT + 2; ++i) C[i] = B[i] = F::zero();
    C[0] = B[0] = F::one();
    E b = F::one();
    uint32_t L = 0, mm = 1;
    bool ok = true;
#pragma unroll 1
    for (uint32_t nn = 0; nn < 2u * t && ok; ++nn) {
      E d = F::zero();
#pragma unroll 1
      for (uint32_t i = 0; i <= L; ++i) d = F::add(d, F::mul(C[i], S[nn - i]));
      if (F::is_zero(d)) {
        ++mm;
        continue;
      }
      const bool grow = 2u * L <= nn;
      E Tm[kSynMaxT + 2];
      if (grow) {
#pragma unroll 1
        for (uint32_t i = 0; i < kSynMaxT + 2; ++i) Tm[i] = C[i];
      }
#pragma unroll 1
      for (uint32_t i = kSynMaxT + 1; i + 1 > 0; --i) {  // high to low: C[i] = b C[i] - d B[i - mm]
        const E sub = i >= mm ? F::mul(d, B[i - mm]) : F::zero();
        C[i] = F::sub(F::mul(b, C[i]), sub);
      }
      if (grow) {
        L = nn + 1u - L;
        if (L > t) ok = false;  // more than t errors: not this decoder's business
#pragma unroll 1
        for (uint32_t i = 0; i < kSynMaxT + 2; ++i) B[i] = Tm[i];
        b = d;
        mm = 1;
      } else {
        ++mm;
      }
    }
    ok = ok && L >= 1u && L <= t && !F::is_zero(C[0]);
    // ---- roots among the nodes: sigma(x) = sum_m C[L - m] x^m
    uint32_t errpos[kSynMaxT];
    uint32_t n_err = 0;
    if (ok) {
#pragma unroll 1
      for (uint32_t i = 0; i < np; ++i) {
        E y = C[0];
#pragma unroll 1
        for (uint32_t k = 1; k <= L; ++k) y = F::add(F::mul(y, A[i]), C[k]);
        if (F::is_zero(y)) {
          if (n_err < kSynMaxT) errpos[n_err] = i;
          ++n_err;
        }
      }
      ok = n_err == L;
    }
    E fo[kSynMaxT + 1];
    E Mx[kSynMaxT + 1];
    if (ok) {
      // ---- monic locator M(x) = prod (x - a_i), low coefficient first
      Mx[0] = F::one();
#pragma unroll 1
      for (uint32_t u = 0; u < L; ++u) {
        const E a = A[errpos[u]];
        Mx[u + 1] = Mx[u];
#pragma unroll 1
        for (uint32_t k = u; k >= 1; --k) Mx[k] = F::sub(Mx[k - 1], F::mul(a, Mx[k]));
        Mx[0] = F::neg(F::mul(a, Mx[0]));
      }
      // ---- error values: numerators and denominators, then one inversion for all
      E num[kSynMaxT], den[kSynMaxT], pre[kSynMaxT];
#pragma unroll 1
      for (uint32_t u = 0; u < L; ++u) {
        const E a = A[errpos[u]];
        // N = M / (x - a) by synthetic division from the top: N[L-1] = 1, N[k-1] = M[k] + a N[k]
        E nk = F::one(), dotp = S[L - 1], ev = F::one();
#pragma unroll 1
        for (uint32_t k = L - 1; k >= 1; --k) {
          nk = F::add(Mx[k], F::mul(a, nk));       // N[k-1]
          dotp = F::add(dotp, F::mul(nk, S[k - 1]));
          ev = F::add(F::mul(ev, a), nk);           // Horner of N at a, from the top
        }
        num[u] = dotp;
        den[u] = ev;
      }
      E run = F::one();
#pragma unroll 1
      for (uint32_t u = 0; u < L; ++u) {
        pre[u] = run;
        run = F::mul(run, den[u]);
      }
      ok = !F::is_zero(run);
      if (ok) {
        E inv = F::inv(run);
#pragma unroll 1
        for (uint32_t u = L; u-- > 0;) {
          const E dinv = F::mul(inv, pre[u]);
          inv = F::mul(inv, den[u]);
          const uint32_t i = errpos[u];
          const E e = F::mul(F::mul(num[u], dinv), WI[i]);  // (w_i e_i) / w_i
          if (i < m) r[i] = F::sub(r[i], e);                // only the first t+1 shares feed the interpolation
        }
        // ---- f from the first t+1 corrected shares
#pragma unroll 1
        for (uint32_t rr = 0; rr < m; ++rr) {
          const E* row = COEF + rr * m;
          typename F::Acc acc = F::acc_zero();
          uint32_t terms = 0;
#pragma unroll 1
          for (uint32_t k = 0; k < m; ++k) {
            F::mac(acc, r[k], row[k]);
            if (++terms == F::ACC_TERMS - 1) {  // the lazy accumulator holds ACC_TERMS products (m can reach 56 here)
              F::acc_fold(acc);
              terms = 0;
            }
          }
          fo[rr] = F::acc_reduce(acc);
        }
        // ---- verification against the ORIGINAL shares: disagreements exactly at the positions found
        uint32_t u = 0, bad = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < np && ok; ++i) {
          E y = fo[t];
#pragma unroll 1
          for (uint32_t k = t; k-- > 0;) y = F::add(F::mul(y, A[i]), fo[k]);
          const E orig = src[(uint64_t)i * stride_i];
          const bool is_err = u < L && errpos[u] == i;
          if (is_err) ++u;
          const bool differs = !F::eq(y, orig);
          if (differs != is_err) ok = false;
          bad += differs;
        }
        ok = ok && bad == L;
      }
    }
    if (ok) {
      E* fp = f_out + j * np;
#pragma unroll 1
      for (uint32_t k = 0; k < np; ++k) fp[k] = k < m ? fo[k] : F::zero();
      E* ep = e_out + j * (uint64_t)m;
#pragma unroll 1
      for (uint32_t k = 0; k < m; ++k) ep[k] = k <= L ? Mx[k] : F::zero();
      status[j] = 0;
    } else {
      pending2[atomicAdd(n_pending2, 1ull)] = (uint32_t)j;
    }
  }
}

}  // namespace sclgpu
