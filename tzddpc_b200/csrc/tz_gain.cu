// Batched robust gain synthesis: the counterpart of compute_theta (tzddpc/utils.py:58-103), one CTA per data set.
//
//   An, Bn = A0, B0                                                           (utils.py:72-75)
//   repeat                                                                    (utils.py:80-94)
//     K      = stabilising gain of (An, Bn)       reference: LMI feasibility SDP (utils.py:43-56), any feasible point;
//                                                 here: the LQR gain of the DARE (Q = R = I), doubling algorithm
//     An, Bn = argmax ||A + B K||_F over M_Sigma  reference: DCCP + MOSEK from `initial_points` starts (utils.py:13-41),
//              (independent beta_A, beta_B)       here: the same convex-concave iteration in closed form (the linearised
//                                                 problem is maximised by beta = sign(gradient)), starts: centre + Philox
//     lambda_max = max(rho(An + Bn K), rho(A0 + B0 K))
//   until lambda_max < 1 or |lambda_max - previous| < tol or max_iter
//   is_gain_robust (utils.py:105-129): N samples of M_Sigma, all rho(A + B K) < 1
//
// The reference's K is solver-dependent (and needs MOSEK), so this path is opt-in and is checked against the numpy
// restatement oracle/gain.py, not against the reference.
// M_Sigma's generators are -g_k P[j,:] (rank one): A + B K = F0 - sum_k g_k r_k', r_k = sum_j bA_kj P[j,:n] + bB_kj P[j,n:] K;
// P = pinv([X0;U0]) ((T-1) x (n+m), from tz_identify) is staged in shared memory once, the (T-1) generators are spread over
// the threads.  Spectral radii by repeated squaring (Gelfand), each thread its own n x n matrix.
#include "tz_philox.cuh"

namespace tz {

constexpr int kGThreads = 128;
constexpr int kMaxGW = 4;
constexpr int kSquarings = 30;
constexpr int kNN = kMaxN * kMaxN;

template <int N>
__device__ double spectral_radius_n(const double* M) {
  double X[N * N], Y[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) X[i] = M[i];
  double logr = 0.0, w = 1.0;
  for (int it = 0; it <= kSquarings; ++it) {
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < N * N; ++i) s2 = fma(X[i], X[i], s2);
    const double s = sqrt(s2);
    if (!(s < INFINITY)) return INFINITY;
    if (s == 0.0) return 0.0;
    logr = fma(w, log(s), logr);
    if (it == kSquarings) break;
    const double inv = 1.0 / s;
#pragma unroll
    for (int i = 0; i < N * N; ++i) X[i] *= inv;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(X[i * N + k], X[k * N + j], acc);
        Y[i * N + j] = acc;
      }
#pragma unroll
    for (int i = 0; i < N * N; ++i) X[i] = Y[i];
    w *= 0.5;
  }
  return exp(logr);
}

// rho(M) by repeated squaring; the matrix lives in registers (the size is a template parameter: with a run-time n the
// two n x n arrays sat in local memory and their loads were 40 % of the kernel's stall samples, profiles/r1_gain_kernel_ncu.txt)
__device__ double spectral_radius_sq(const double* M, int n) {
  switch (n) {
    case 1: return fabs(M[0]);
    case 2: return spectral_radius_n<2>(M);
    case 3: return spectral_radius_n<3>(M);
    case 4: return spectral_radius_n<4>(M);
    case 5: return spectral_radius_n<5>(M);
    case 6: return spectral_radius_n<6>(M);
    case 7: return spectral_radius_n<7>(M);
    default: return spectral_radius_n<8>(M);
  }
}

// C (p x r) = A (p x q) B (q x r)
__device__ void mm(double* C, const double* A, const double* B, int p, int q, int r) {
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < r; ++j) {
      double acc = 0.0;
      for (int k = 0; k < q; ++k) acc = fma(A[i * q + k], B[k * r + j], acc);
      C[i * r + j] = acc;
    }
}
// C (q x r) = A' B with A (p x q), B (p x r)
__device__ void mtm(double* C, const double* A, const double* B, int p, int q, int r) {
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < r; ++j) {
      double acc = 0.0;
      for (int k = 0; k < p; ++k) acc = fma(A[k * q + i], B[k * r + j], acc);
      C[i * r + j] = acc;
    }
}
// C (p x r) = A B' with A (p x q), B (r x q)
__device__ void mmt(double* C, const double* A, const double* B, int p, int q, int r) {
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < r; ++j) {
      double acc = 0.0;
      for (int k = 0; k < q; ++k) acc = fma(A[i * q + k], B[j * q + k], acc);
      C[i * r + j] = acc;
    }
}
// X (n x r) = A^-1 B by Gaussian elimination with partial pivoting (A is destroyed); false when singular
__device__ bool solve_inplace(double* A, double* B, int n, int r) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double big = fabs(A[c * n + c]);
    for (int i = c + 1; i < n; ++i)
      if (fabs(A[i * n + c]) > big) { big = fabs(A[i * n + c]); piv = i; }
    if (!(big > 0.0) || !(big < INFINITY)) return false;
    if (piv != c) {
      for (int j = 0; j < n; ++j) { const double t = A[c * n + j]; A[c * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
      for (int j = 0; j < r; ++j) { const double t = B[c * r + j]; B[c * r + j] = B[piv * r + j]; B[piv * r + j] = t; }
    }
    const double inv = 1.0 / A[c * n + c];
    for (int i = c + 1; i < n; ++i) {
      const double f = A[i * n + c] * inv;
      if (f == 0.0) continue;
      for (int j = c; j < n; ++j) A[i * n + j] = fma(-f, A[c * n + j], A[i * n + j]);
      for (int j = 0; j < r; ++j) B[i * r + j] = fma(-f, B[c * r + j], B[i * r + j]);
    }
  }
  for (int c = n - 1; c >= 0; --c) {
    const double inv = 1.0 / A[c * n + c];
    for (int j = 0; j < r; ++j) {
      double acc = B[c * r + j];
      for (int k = c + 1; k < n; ++k) acc = fma(-A[c * n + k], B[k * r + j], acc);
      B[c * r + j] = acc * inv;
    }
  }
  return true;
}

// LQR gain of (A, B), Q = R = I: structure-preserving doubling  W = (I + G H)^-1; A <- A W A; G <- G + A W G A';
// H <- H + A' H W A;  H -> P;  K = -(I + B'PB)^-1 B'PA.   (one thread; n <= 8)
__device__ bool lqr_gain_sda(const double* A, const double* B, int n, int m, double* K) {
  double Ak[kNN], G[kNN], H[kNN], W[kNN], T1[kNN], T2[kNN], T3[kNN];
  for (int i = 0; i < n * n; ++i) { Ak[i] = A[i]; H[i] = 0.0; }
  for (int i = 0; i < n; ++i) H[i * n + i] = 1.0;
  mmt(G, B, B, n, m, n);
  bool ok = true;
  for (int it = 0; it < 60; ++it) {
    mm(T1, G, H, n, n, n);                                   // I + G H
    for (int i = 0; i < n; ++i) T1[i * n + i] += 1.0;
    for (int i = 0; i < n * n; ++i) W[i] = 0.0;
    for (int i = 0; i < n; ++i) W[i * n + i] = 1.0;
    if (!solve_inplace(T1, W, n, n)) { ok = false; break; }  // W = (I + G H)^-1
    mm(T1, Ak, W, n, n, n);                                  // AW
    mm(T2, T1, G, n, n, n);                                  // AW G
    mmt(T3, T2, Ak, n, n, n);                                // AW G A'
    for (int i = 0; i < n * n; ++i) G[i] += T3[i];
    mtm(T2, Ak, H, n, n, n);                                 // A' H
    mm(T3, T2, W, n, n, n);                                  // A' H W
    mm(T2, T3, Ak, n, n, n);                                 // A' H W A
    mm(T3, T1, Ak, n, n, n);                                 // A W A
    double dn = 0.0, hn = 0.0;
    for (int i = 0; i < n * n; ++i) {
      dn = fma(T2[i], T2[i], dn);
      H[i] += T2[i];
      hn = fma(H[i], H[i], hn);
      Ak[i] = T3[i];
    }
    if (!(hn < INFINITY)) { ok = false; break; }
    if (sqrt(dn) <= 1e-15 * sqrt(hn)) break;
  }
  // K = -(I + B'PB)^-1 B'PA
  double BtP[kMaxM * kMaxN], S[kMaxM * kMaxM], R[kMaxM * kMaxN];
  mtm(BtP, B, H, n, m, n);
  mm(S, BtP, B, m, n, m);
  for (int i = 0; i < m; ++i) S[i * m + i] += 1.0;
  mm(R, BtP, A, m, n, n);
  if (!solve_inplace(S, R, m, n)) ok = false;
  for (int i = 0; i < m * n; ++i) {
    K[i] = -R[i];
    ok = ok && (fabs(K[i]) < INFINITY);
  }
  return ok;
}

struct GainArgs {
  int T, n, m, gW;
  const double *AB, *Pinv, *WZ;
  const double* K_fixed;          // tz_gain_adversary: the caller's gain (D x m x n); no synthesis, one adversary pass
  double tol;
  int max_iter, num_init, nsamp;
  uint64_t seed;
  int64_t dataset_offset;
  double *K, *dA, *dB, *rho;
  int32_t *robust, *iters, *status;
};

// CTA reduction of `cnt` doubles per thread (cnt <= 2 * kMaxGW * kMaxN); result in out[] (shared), valid after the call
__device__ void cta_reduce(const double* v, int cnt, double (*part)[2 * kMaxGW * kMaxN], double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = 0; i < cnt; ++i) {
    const double s = warp_sum(v[i]);
    if (lane == 0) part[wid][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < cnt) {
    double s = 0.0;
    for (int w = 0; w < kGThreads / 32; ++w) s += part[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kGThreads) gain_kernel(const GainArgs a) {
  extern __shared__ double smg[];                  // P: Tm x d | PBK: Tm x n | signs: 4 x gW*Tm bytes
  __shared__ double A0[kNN], B0[kMaxN * kMaxM], An[kNN], Bn[kMaxN * kMaxM], K[kMaxM * kMaxN], F0[kNN], F[kNN];
  __shared__ double GW[kMaxN * kMaxGW], Y[kMaxGW * kMaxN], red[2 * kMaxGW * kMaxN];
  __shared__ double part[kGThreads / 32][2 * kMaxGW * kMaxN];
  __shared__ double s_best, s_f;
  __shared__ int s_flag, s_stop, s_ok;
  const int n = a.n, m = a.m, d = n + m, Tm = a.T - 1, gW = a.gW, tid = threadIdx.x;
  const int64_t ds = blockIdx.x;
  const int ngen = gW * Tm;
  double* P = smg;
  double* PBK = P + (size_t)Tm * d;
  signed char* bA = reinterpret_cast<signed char*>(PBK + (size_t)Tm * n);
  signed char* bB = bA + ngen;
  signed char* bestA = bB + ngen;
  signed char* bestB = bestA + ngen;
  const double* Pg = a.Pinv + ds * (int64_t)Tm * d;
  for (int i = tid; i < Tm * d; i += kGThreads) P[i] = Pg[i];
  if (tid < n * n) { A0[tid] = a.AB[ds * n * d + (tid / n) * d + tid % n]; An[tid] = A0[tid]; }
  if (tid < n * m) { B0[tid] = a.AB[ds * n * d + (tid / m) * d + n + tid % m]; Bn[tid] = B0[tid]; }
  if (tid < n * gW) GW[tid] = a.WZ[(tid / gW) * (1 + gW) + 1 + tid % gW];
  if (tid == 0) { s_ok = 1; s_stop = 0; }
  __syncthreads();

  double prev = 0.0, rho0 = NAN, rho_adv = NAN;
  int iteration = 0;
  while (true) {
    // ---- K = LQR gain of (An, Bn)
    if (tid == 0) {
      double Kl[kMaxM * kMaxN];
      bool ok = true;
      if (a.K_fixed != nullptr) {
        for (int i = 0; i < m * n; ++i) Kl[i] = a.K_fixed[ds * m * n + i];
      } else {
        ok = lqr_gain_sda(An, Bn, n, m, Kl);
      }
      for (int i = 0; i < m * n; ++i) K[i] = Kl[i];
      if (!ok) s_ok = 0;
      double f0[kNN];
      mm(f0, B0, Kl, n, m, n);
      for (int i = 0; i < n * n; ++i) F0[i] = A0[i] + f0[i];
      s_best = -INFINITY;
    }
    __syncthreads();
    if (!s_ok) break;
    for (int i = tid; i < Tm * n; i += kGThreads) {           // PBK[j,:] = P[j,n:] K
      const int j = i / n, c = i - j * n;
      double acc = 0.0;
      for (int q = 0; q < m; ++q) acc = fma(P[(size_t)j * d + n + q], K[q * n + c], acc);
      PBK[i] = acc;
    }
    __syncthreads();
    // ---- adversary: convex-concave iteration from the centre and from random starts
    for (int start = 0; start < a.num_init; ++start) {
      // r_k = sum_j bA_kj PA_j + bB_kj PBK_j with the STARTING beta (real-valued draws; zero for start 0)
      double acc[kMaxGW * kMaxN];
      for (int i = 0; i < gW * n; ++i) acc[i] = 0.0;
      if (start > 0) {
        for (int g = tid; g < ngen; g += kGThreads) {
          const int k = g / Tm, j = g - k * Tm;
          const double ba = draw(a.seed, (uint64_t)(a.dataset_offset + ds), (uint32_t)start, 4u, g, false);
          const double bb = draw(a.seed, (uint64_t)(a.dataset_offset + ds), (uint32_t)start, 4u, ngen + g, false);
          for (int c = 0; c < n; ++c) acc[k * n + c] += ba * P[(size_t)j * d + c] + bb * PBK[(size_t)j * n + c];
        }
      }
      for (int g = tid; g < ngen; g += kGThreads) { bA[g] = 0; bB[g] = 0; }
      for (int ccp = 0; ccp < 50; ++ccp) {
        cta_reduce(acc, gW * n, part, red);
        if (tid == 0) {                                        // F = F0 - sum_k g_k r_k',  Y[k] = g_k' F
          for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) {
              double v = F0[r * n + c];
              for (int k = 0; k < gW; ++k) v = fma(-GW[r * gW + k], red[k * n + c], v);
              F[r * n + c] = v;
            }
          for (int k = 0; k < gW; ++k)
            for (int c = 0; c < n; ++c) {
              double v = 0.0;
              for (int r = 0; r < n; ++r) v = fma(GW[r * gW + k], F[r * n + c], v);
              Y[k * n + c] = v;
            }
          s_flag = 0;
        }
        __syncthreads();
        int changed = 0;
        for (int i = 0; i < gW * n; ++i) acc[i] = 0.0;
        for (int g = tid; g < ngen; g += kGThreads) {
          const int k = g / Tm, j = g - k * Tm;
          double da = 0.0, db = 0.0;
          for (int c = 0; c < n; ++c) {
            da = fma(Y[k * n + c], P[(size_t)j * d + c], da);
            db = fma(Y[k * n + c], PBK[(size_t)j * n + c], db);
          }
          const signed char na = da > 0.0 ? -1 : 1, nb = db > 0.0 ? -1 : 1;
          changed |= (na != bA[g]) || (nb != bB[g]);
          bA[g] = na; bB[g] = nb;
          for (int c = 0; c < n; ++c) acc[k * n + c] += (double)na * P[(size_t)j * d + c] + (double)nb * PBK[(size_t)j * n + c];
        }
        if (changed) s_flag = 1;
        __syncthreads();
        if (!s_flag) break;
        __syncthreads();
      }
      // objective at the final vertex
      cta_reduce(acc, gW * n, part, red);
      if (tid == 0) {
        double f = 0.0;
        for (int r = 0; r < n; ++r)
          for (int c = 0; c < n; ++c) {
            double v = F0[r * n + c];
            for (int k = 0; k < gW; ++k) v = fma(-GW[r * gW + k], red[k * n + c], v);
            f = fma(v, v, f);
          }
        s_f = f;
        s_flag = f > s_best ? 1 : 0;
        if (s_flag) s_best = f;
      }
      __syncthreads();
      if (s_flag)
        for (int g = tid; g < ngen; g += kGThreads) { bestA[g] = bA[g]; bestB[g] = bB[g]; }
      __syncthreads();
    }
    // ---- An = A0 - sum_k g_k (sum_j bA_kj PA_j)',  Bn = B0 - sum_k g_k (sum_j bB_kj PB_j)'
    {
      double acc[kMaxGW * (kMaxN + kMaxM)];
      for (int i = 0; i < gW * d; ++i) acc[i] = 0.0;
      for (int g = tid; g < ngen; g += kGThreads) {
        const int k = g / Tm, j = g - k * Tm;
        for (int c = 0; c < n; ++c) acc[k * d + c] += (double)bestA[g] * P[(size_t)j * d + c];
        for (int c = n; c < d; ++c) acc[k * d + c] += (double)bestB[g] * P[(size_t)j * d + c];
      }
      cta_reduce(acc, gW * d, part, red);
      if (tid == 0) {
        for (int r = 0; r < n; ++r) {
          for (int c = 0; c < n; ++c) {
            double v = A0[r * n + c];
            for (int k = 0; k < gW; ++k) v = fma(-GW[r * gW + k], red[k * d + c], v);
            An[r * n + c] = v;
          }
          for (int c = 0; c < m; ++c) {
            double v = B0[r * m + c];
            for (int k = 0; k < gW; ++k) v = fma(-GW[r * gW + k], red[k * d + n + c], v);
            Bn[r * m + c] = v;
          }
        }
        double Fa[kNN], t[kNN];
        mm(t, Bn, K, n, m, n);
        for (int i = 0; i < n * n; ++i) Fa[i] = An[i] + t[i];
        rho_adv = spectral_radius_sq(Fa, n);
        rho0 = spectral_radius_sq(F0, n);
        const double lam = fmax(rho_adv, rho0);
        int stop = (fabs(lam - prev) < a.tol || lam < 1.0 || a.K_fixed != nullptr) ? 1 : 0;
        if (!stop) {
          ++iteration;
          prev = lam;
          if (iteration >= a.max_iter) stop = 1;
        }
        s_stop = stop;
      }
      __syncthreads();
    }
    if (s_stop) break;
  }
  // ---- outputs + Monte-Carlo robustness check
  const bool ok = s_ok != 0;
  double worst = 0.0;
  if (ok) {
    // Q_j = PA_j + PBK_j; sample t: F = F0 - sum_k g_k (sum_j beta_kj Q_j)'
    for (int t = tid; t < a.nsamp; t += kGThreads) {
      double r[kMaxGW * kMaxN];
      for (int i = 0; i < gW * n; ++i) r[i] = 0.0;
      for (int g = 0; g < ngen; ++g) {
        const int k = g / Tm, j = g - k * Tm;
        const double b = draw(a.seed, (uint64_t)(a.dataset_offset + ds), (uint32_t)t, 5u, g, false);
        for (int c = 0; c < n; ++c) r[k * n + c] = fma(b, P[(size_t)j * d + c] + PBK[(size_t)j * n + c], r[k * n + c]);
      }
      double Fs[kNN];
      for (int rr = 0; rr < n; ++rr)
        for (int c = 0; c < n; ++c) {
          double v = F0[rr * n + c];
          for (int k = 0; k < gW; ++k) v = fma(-GW[rr * gW + k], r[k * n + c], v);
          Fs[rr * n + c] = v;
        }
      worst = fmax(worst, spectral_radius_sq(Fs, n));
    }
  }
  // max over the CTA
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((tid & 31) == 0) part[tid >> 5][0] = worst;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kGThreads / 32; ++w) worst = fmax(worst, part[w][0]);
    if (a.K != nullptr)
      for (int i = 0; i < m * n; ++i) a.K[ds * m * n + i] = K[i];
    for (int i = 0; i < n * n; ++i) a.dA[ds * n * n + i] = An[i] - A0[i];
    for (int i = 0; i < n * m; ++i) a.dB[ds * n * m + i] = Bn[i] - B0[i];
    a.rho[ds * 3 + 0] = ok ? rho0 : NAN;
    a.rho[ds * 3 + 1] = ok ? rho_adv : NAN;
    a.rho[ds * 3 + 2] = ok ? worst : NAN;
    a.robust[ds] = (ok && worst < 1.0) ? 1 : 0;
    if (a.iters != nullptr) a.iters[ds] = iteration;
    a.status[ds] = ok ? TZ_STATUS_OK : TZ_STATUS_NONFINITE;
  }
}

}  // namespace tz

using namespace tz;

extern "C" int tz_gain_synthesis(int64_t D, int32_t T, int32_t n, int32_t m, int32_t gW, const double* AB, const double* Pinv,
                                 const double* WZ, double tol, int32_t max_iter, int32_t num_init, double accuracy,
                                 double confidence, uint64_t seed, int64_t dataset_offset, double* K, double* dA, double* dB,
                                 double* rho, int32_t* robust, int32_t* iters, int32_t* status, void* stream) {
  TZ_REQUIRE(D >= 0 && T >= 2 && n >= 1 && n <= kMaxN && m >= 1 && m <= kMaxM && gW >= 1 && gW <= kMaxGW, "bad shape");
  TZ_REQUIRE(accuracy > 0 && accuracy < 1 && confidence > 0 && confidence < 1, "accuracy and confidence must be in (0, 1)");
  TZ_REQUIRE(max_iter >= 1 && num_init >= 1 && tol >= 0, "bad iteration options");
  TZ_REQUIRE((int64_t)2 * gW * (T - 1) < 131072, "too many generators for the draw counter");
  if (D == 0) return TZ_OK;
  TZ_REQUIRE(AB && Pinv && WZ && K && dA && dB && rho && robust && iters && status, "null pointer");
  GainArgs a;
  a.K_fixed = nullptr;
  a.T = T; a.n = n; a.m = m; a.gW = gW; a.AB = AB; a.Pinv = Pinv; a.WZ = WZ; a.tol = tol; a.max_iter = max_iter;
  a.num_init = num_init; a.seed = seed; a.dataset_offset = dataset_offset;
  a.nsamp = (int)ceil(log(1.0 / confidence) / log(1.0 / (1.0 - accuracy)));      // utils.py:120
  a.K = K; a.dA = dA; a.dB = dB; a.rho = rho; a.robust = robust; a.iters = iters; a.status = status;
  const size_t Tm = (size_t)T - 1;
  size_t smem = Tm * (size_t)(n + m + n) * sizeof(double) + 4 * (size_t)gW * Tm + 16;
  TZ_REQUIRE(smem <= 200 * 1024, "data set too long for the shared-memory staging (%zu bytes)", smem);
  if (smem > 40 * 1024) TZ_CUDA(cudaFuncSetAttribute(gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gain_kernel<<<(unsigned)D, kGThreads, smem, (cudaStream_t)stream>>>(a);
  TZ_CUDA(cudaGetLastError());
  return TZ_OK;
}

// compute_A_B + is_gain_robust for a GIVEN gain (tzddpc/utils.py:13-41,105-129): the adversarial pair of M_Sigma for K and
// the Monte-Carlo robustness check, without the synthesis loop.
extern "C" int tz_gain_adversary(int64_t D, int32_t T, int32_t n, int32_t m, int32_t gW, const double* AB, const double* Pinv,
                                 const double* WZ, const double* K, int32_t num_init, double accuracy, double confidence,
                                 uint64_t seed, int64_t dataset_offset, double* dA, double* dB, double* rho, int32_t* robust,
                                 int32_t* status, void* stream) {
  TZ_REQUIRE(D >= 0 && T >= 2 && n >= 1 && n <= kMaxN && m >= 1 && m <= kMaxM && gW >= 1 && gW <= kMaxGW, "bad shape");
  TZ_REQUIRE(accuracy > 0 && accuracy < 1 && confidence > 0 && confidence < 1, "accuracy and confidence must be in (0, 1)");
  TZ_REQUIRE(num_init >= 1, "bad iteration options");
  TZ_REQUIRE((int64_t)2 * gW * (T - 1) < 131072, "too many generators for the draw counter");
  if (D == 0) return TZ_OK;
  TZ_REQUIRE(AB && Pinv && WZ && K && dA && dB && rho && robust && status, "null pointer");
  GainArgs a;
  a.K_fixed = K;
  a.T = T; a.n = n; a.m = m; a.gW = gW; a.AB = AB; a.Pinv = Pinv; a.WZ = WZ; a.tol = 0.0; a.max_iter = 1;
  a.num_init = num_init; a.seed = seed; a.dataset_offset = dataset_offset;
  a.nsamp = (int)ceil(log(1.0 / confidence) / log(1.0 / (1.0 - accuracy)));      // utils.py:120
  a.K = nullptr; a.dA = dA; a.dB = dB; a.rho = rho; a.robust = robust; a.iters = nullptr; a.status = status;
  const size_t Tm = (size_t)T - 1;
  size_t smem = Tm * (size_t)(n + m + n) * sizeof(double) + 4 * (size_t)gW * Tm + 16;
  TZ_REQUIRE(smem <= 200 * 1024, "data set too long for the shared-memory staging (%zu bytes)", smem);
  if (smem > 40 * 1024) TZ_CUDA(cudaFuncSetAttribute(gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gain_kernel<<<(unsigned)D, kGThreads, smem, (cudaStream_t)stream>>>(a);
  TZ_CUDA(cudaGetLastError());
  return TZ_OK;
}

extern "C" int32_t tz_gain_robust_samples(double accuracy, double confidence) {
  if (!(accuracy > 0 && accuracy < 1 && confidence > 0 && confidence < 1)) return -1;
  return (int32_t)ceil(log(1.0 / confidence) / log(1.0 / (1.0 - accuracy)));
}
