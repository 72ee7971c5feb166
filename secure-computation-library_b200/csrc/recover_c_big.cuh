// recover_c_big.cuh -- shamirRecoverC (Berlekamp-Welch, include/scl/ss/shamir.h:203-258) for sharings with more than
// 32 interpolation points (3t+1 > 32): the reference takes any size (solveLinearSystem, matrix.h:598-828; division,
// poly.h:262-278), and so does this path.  k_recover_c (kernels.cuh) gives a warp to a sharing, one lane per row of
// the (3t+1) x (3t+2) system; here a sharing gets a CTA of ceil32(3t+1) <= 256 threads, thread i owning row i of the
// system in shared memory, with the warp votes replaced by shared-memory minima.  Same algorithm, same results:
//   quick: ONE rank-revealing, fraction-free Gauss-Jordan elimination of the e = t system (no row exchanges, free
//          unknowns := 0), f = Q / E, then verification -- degree <= t and at most t disagreements with the shares
//          mean, by unique decoding, that the reference's own answer is f with the locator of the disagreeing
//          positions; anything else falls through to
//   full : the reference's sequence e = t, t-1, .., 0, accepting the first uniquely solvable system.
// Error-free sharings never get here: k_recover_c_clean_any settles them with 2t Lagrange checks.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace sclgpu {

static constexpr uint32_t kRecoverCBigMaxPoints = 256;

// shared-memory bytes of one sharing: matrix np x (np + 1), X, R, Qt, bad flags, a few control words
template <class F>
static inline size_t recover_c_big_smem(uint32_t np) {
  return ((size_t)np * (np + 1) + 3 * (size_t)np) * sizeof(typename F::E) + (size_t)np * 4 + 64;
}

// error-free fast path for any t (k_recover_c_clean keeps the first t+1 shares in registers, t <= 10)
template <class F>
__global__ void __launch_bounds__(256)
k_recover_c_clean_any(const typename F::E* __restrict__ in, uint64_t N, uint64_t stride_i, uint64_t stride_j, uint32_t t,
                      const typename F::E* __restrict__ check, const typename F::E* __restrict__ coef,
                      typename F::E* __restrict__ f_out, typename F::E* __restrict__ e_out, uint8_t* __restrict__ status,
                      uint32_t* __restrict__ pending, unsigned long long* __restrict__ n_pending) {
  typedef typename F::E E;
  const uint32_t m = t + 1u, np = 3u * t + 1u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const E* src = in + j * stride_j;
    bool clean = true;
    for (uint32_t r = 0; r < 2u * t && clean; ++r) {
      const E* row = check + (size_t)r * m;
      typename F::Acc acc = F::acc_zero();
      for (uint32_t k = 0; k < m; ++k) {
        if (k && (k % (uint32_t)F::ACC_TERMS) == 0) F::acc_fold(acc);
        F::mac(acc, src[(uint64_t)k * stride_i], row[k]);
      }
      clean = F::eq(F::acc_reduce(acc), src[(uint64_t)(m + r) * stride_i]);
    }
    if (clean) {
      E* fo = f_out + j * np;
      for (uint32_t r = 0; r < m; ++r) {
        const E* row = coef + (size_t)r * m;
        typename F::Acc acc = F::acc_zero();
        for (uint32_t k = 0; k < m; ++k) {
          if (k && (k % (uint32_t)F::ACC_TERMS) == 0) F::acc_fold(acc);
          F::mac(acc, src[(uint64_t)k * stride_i], row[k]);
        }
        fo[r] = F::acc_reduce(acc);
      }
      for (uint32_t k = m; k < np; ++k) fo[k] = F::zero();
      E* eo = e_out + j * (uint64_t)m;
      eo[0] = F::one();
      for (uint32_t k = 1; k < m; ++k) eo[k] = F::zero();
      status[j] = 0;
    } else {
      status[j] = 2;  // kRecoverCPending
      pending[atomicAdd(n_pending, 1ull)] = (uint32_t)j;
    }
  }
}

template <class F>
__global__ void __launch_bounds__(256)
k_recover_c_cta(const typename F::E* __restrict__ in, uint64_t N, uint64_t stride_i, uint64_t stride_j, uint32_t t,
                const typename F::E* __restrict__ alphas, typename F::E* __restrict__ f_out,
                typename F::E* __restrict__ e_out, uint8_t* __restrict__ status, unsigned long long* __restrict__ n_failed,
                const uint32_t* __restrict__ pending, const unsigned long long* __restrict__ n_pending, int quick) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t np = 3u * t + 1u, cols = np + 1u;
  E* M = reinterpret_cast<E*>(dyn_smem);
  E* X = M + (size_t)np * cols;
  E* R = X + np;
  E* Qt = R + np;
  int* bad = reinterpret_cast<int*>(Qt + np);          // per row: disagrees with the decoded polynomial
  int* ctl = bad + np;                                   // [1] flag, [2] degree, [3] count, [4..5] pivot rows
  const uint32_t tid = threadIdx.x;
  const bool row = tid < np;
  const E minus1 = F::neg(F::one());
  const E a_i = row ? alphas[tid] : F::zero();
  unsigned long long local_failed = 0;

  const uint64_t n_work = pending ? *n_pending : N;
  for (uint64_t q = blockIdx.x; q < n_work; q += gridDim.x) {
    const uint64_t j = pending ? pending[q] : q;
    const E s_i = row ? in[(uint64_t)tid * stride_i + j * stride_j] : F::zero();
    E* fo = f_out + j * np;
    E* eo = e_out + j * (uint64_t)(t + 1);
    bool done = false;
    __syncthreads();
    if (quick && t > 0) {
      if (row) {
        E* mr = M + (size_t)tid * cols;
        E v = s_i;
        for (uint32_t c = 0; c < t; ++c) {
          mr[c] = v;
          v = F::mul(v, a_i);
        }
        mr[np] = F::neg(v);
        E u = minus1;
        for (uint32_t c = t; c < np; ++c) {
          mr[c] = u;
          u = F::mul(u, a_i);
        }
        X[tid] = F::zero();
      }
      int my_col = -1;
      for (uint32_t c = 0; c < np; ++c) {
        int* slot = ctl + 4 + (c & 1u);  // alternating pivot slots: resetting the next one never touches the one being read
        if (tid == 0) *slot = 0x7fffffff;
        __syncthreads();
        if (row && my_col < 0 && !F::is_zero(M[(size_t)tid * cols + c])) atomicMin(slot, (int)tid);
        __syncthreads();
        const int piv = *slot;
        if (piv == 0x7fffffff) continue;  // free unknown
        if ((int)tid == piv) my_col = (int)c;
        const E p = M[(size_t)piv * cols + c];
        if (row && (int)tid != piv) {
          E* mr = M + (size_t)tid * cols;
          const E f = mr[c];
          if (!F::is_zero(f)) {
            const E nf = F::neg(f);
            const E* pr = M + (size_t)piv * cols;
            for (uint32_t k = c + 1; k < cols; ++k) {
              typename F::Acc acc = F::acc_zero();
              F::mac(acc, mr[k], p);
              F::mac(acc, nf, pr[k]);
              mr[k] = F::acc_reduce(acc);
            }
            mr[c] = F::zero();
            if (my_col >= 0) mr[my_col] = F::mul(mr[my_col], p);
          }
        }
        __syncthreads();
      }
      // rows without a pivot read 0 = b: consistent iff b == 0
      if (tid == 0) ctl[1] = 0;
      __syncthreads();
      if (row && my_col < 0 && !F::is_zero(M[(size_t)tid * cols + np])) ctl[1] = 1;
      __syncthreads();
      bool good = ctl[1] == 0;
      if (good) {
        if (row && my_col >= 0) X[my_col] = F::mul(M[(size_t)tid * cols + np], F::inv(M[(size_t)tid * cols + my_col]));
        __syncthreads();
        const uint32_t qn = np - t;  // Q = x[t..np-1], E = (x_0..x_{t-1}, 1)
        if (tid < qn) R[tid] = X[t + tid];
        if (tid < np) Qt[tid] = F::zero();
        if (tid == 0) ctl[2] = -1;
        __syncthreads();
        if (tid < qn && !F::is_zero(R[tid])) atomicMax(&ctl[2], (int)tid);
        __syncthreads();
        const int deg = ctl[2];
        if (deg >= (int)t) {
          for (int d = deg; d >= (int)t; --d) {
            const E c = R[d];
            __syncthreads();
            if (tid < t) R[d - t + tid] = F::sub(R[d - t + tid], F::mul(c, X[tid]));
            if (tid == 0) {
              Qt[d - t] = c;
              R[d] = F::zero();
            }
            __syncthreads();
          }
          if (tid == 0) ctl[1] = 0;
          __syncthreads();
          if (tid < t && !F::is_zero(R[tid])) ctl[1] = 1;
          __syncthreads();
          good = ctl[1] == 0;
        } else {
          good = deg < 0;
        }
        // f = Qt has degree <= t by construction; count the disagreements with the shares
        if (tid == 0) ctl[3] = 0;
        __syncthreads();
        if (row) {
          E y = F::zero();
          for (int k = (int)t; k >= 0; --k) y = F::add(F::mul(y, a_i), Qt[k]);
          const int b = F::eq(y, s_i) ? 0 : 1;
          bad[tid] = b;
          if (b) atomicAdd(&ctl[3], 1);
        }
        __syncthreads();
        const uint32_t d_err = (uint32_t)ctl[3];
        if (good && d_err <= t) {
          if (tid < np) fo[tid] = Qt[tid];
          if (tid == 0) {  // locator: prod over the bad positions of (x - a_i), low coefficient first
            R[0] = F::one();
            uint32_t dg = 0;
            for (uint32_t i = 0; i < np; ++i) {
              if (!bad[i]) continue;
              const E a = alphas[i];
              R[dg + 1] = R[dg];
              for (uint32_t k = dg; k >= 1; --k) R[k] = F::sub(R[k - 1], F::mul(a, R[k]));
              R[0] = F::neg(F::mul(a, R[0]));
              ++dg;
            }
            for (uint32_t k = 0; k <= t; ++k) eo[k] = k <= dg ? R[k] : F::zero();
            status[j] = 0;
          }
          done = true;
        }
      }
      __syncthreads();
    }
    if (done) continue;

    // ---- the reference's sequence: e = t .. 0, first uniquely solvable system wins
    int e = (int)t;
    for (;; --e) {
      __syncthreads();
      if (row) {
        E* mr = M + (size_t)tid * cols;
        E v = s_i;
        for (int c = 0; c < e; ++c) {
          mr[c] = v;
          v = F::mul(v, a_i);
        }
        mr[np] = F::neg(v);
        E u = minus1;
        for (uint32_t c = (uint32_t)e; c < np; ++c) {
          mr[c] = u;
          u = F::mul(u, a_i);
        }
      }
      bool singular = false;
      for (uint32_t c = 0; c < np; ++c) {
        int* slot = ctl + 4 + (c & 1u);
        if (tid == 0) *slot = 0x7fffffff;
        __syncthreads();
        if (row && tid >= c && !F::is_zero(M[(size_t)tid * cols + c])) atomicMin(slot, (int)tid);
        __syncthreads();
        const int piv = *slot;
        if (piv == 0x7fffffff) {
          singular = true;
          break;
        }
        if ((uint32_t)piv != c) {
          for (uint32_t k = tid; k < cols; k += blockDim.x) {
            const E tmp = M[(size_t)piv * cols + k];
            M[(size_t)piv * cols + k] = M[(size_t)c * cols + k];
            M[(size_t)c * cols + k] = tmp;
          }
          __syncthreads();
        }
        const E p = M[(size_t)c * cols + c];
        if (row && tid != c) {
          E* mr = M + (size_t)tid * cols;
          const E f = mr[c];
          if (!F::is_zero(f)) {
            const E nf = F::neg(f);
            const E* pr = M + (size_t)c * cols;
            for (uint32_t k = c + 1; k < cols; ++k) {
              typename F::Acc acc = F::acc_zero();
              F::mac(acc, mr[k], p);
              F::mac(acc, nf, pr[k]);
              mr[k] = F::acc_reduce(acc);
            }
            mr[c] = F::zero();
            if (tid < c) mr[tid] = F::mul(mr[tid], p);  // keep the diagonal of finished rows consistent
          }
        }
        __syncthreads();
      }
      if (!singular) break;
      if (e == 0) {  // only with coinciding nodes: the reference then proceeds with x = 0 (shamir.h:213,238-240)
        e = -1;
        break;
      }
    }
    __syncthreads();
    if (row) X[tid] = e < 0 ? F::zero() : F::mul(M[(size_t)tid * cols + np], F::inv(M[(size_t)tid * cols + tid]));
    if (e < 0) e = 0;
    __syncthreads();
    // Q = x[e..np-1] (trailing zeros stripped), E = (x_0..x_{e-1}, 1); f = Q / E
    const uint32_t ue = (uint32_t)e, qn = np - ue;
    if (tid < qn) R[tid] = X[ue + tid];
    if (tid == 0) ctl[2] = -1;
    __syncthreads();
    if (tid < qn && !F::is_zero(R[tid])) atomicMax(&ctl[2], (int)tid);
    if (tid < np) fo[tid] = F::zero();
    if (tid <= t) eo[tid] = F::zero();
    __syncthreads();
    const int deg = ctl[2];
    bool ok;
    if (deg >= (int)ue) {
      for (int d = deg; d >= (int)ue; --d) {
        const E c = R[d];
        __syncthreads();
        if (tid < ue) R[d - ue + tid] = F::sub(R[d - ue + tid], F::mul(c, X[tid]));
        if (tid == 0) {
          fo[d - ue] = c;
          R[d] = F::zero();
        }
        __syncthreads();
      }
      if (tid == 0) ctl[1] = 0;
      __syncthreads();
      if (tid < ue && !F::is_zero(R[tid])) ctl[1] = 1;
      __syncthreads();
      ok = ctl[1] == 0;
    } else {
      ok = deg < 0;  // Q == 0: quotient and remainder are zero
    }
    if (ok) {
      if (tid < ue) eo[tid] = X[tid];
      if (tid == ue) eo[tid] = F::one();
      if (tid == 0) status[j] = 0;
    } else {
      __syncthreads();
      if (tid < np) fo[tid] = F::zero();
      if (tid == 0) {
        status[j] = 1;
        ++local_failed;
      }
    }
  }
  if (local_failed) atomicAdd(n_failed, local_failed);
}

}  // namespace sclgpu
