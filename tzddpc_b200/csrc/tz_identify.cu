// identify: the data-driven model  M_Sigma = (X1 - M_w) pinv([X0; U0])  (tzddpc/tzddpc.py:81-83) and what
// tzddpc/tzddpc.py:119-128 derives from it (M_K = M_Sigma [I; K], M_Delta, order-1 reduction), one CTA per data set.
//
// With D = [X0; U0] ((n+m) x (T-1), full row rank) the pseudo-inverse is P = D'(DD')^{-1}.  Round 1 formed the Gram matrix
// DD' and solved with its Cholesky factor, which squares the condition number: generator matrices drifted to 1e-8 .. 1e-7
// of numpy's SVD-based pinv on the batched data sets (north_star asks for 1e-9).  Now:
//     Householder QR of D' (T-1 rows, n+m columns, in shared memory):   D' = Q R
//     P  = Q [R^-T; 0]                       (the reflections applied in reverse to the (T-1) x (n+m) block [R^-T; 0])
//     AB = ((Q' X1c)[:n+m])' R^-T            (least squares  min |D' AB' - X1c|, X1c = X1 - c_W 1')
// so the error grows with cond(D), not cond(D)^2.  The order-1 boxes are the closed forms of SURVEY.md App. A.6 for the
// rank-one generators -g_k P[j,:]:
//     dAB = (sum_k |g_k|) (sum_j |P[j,:]|)',     dK = (sum_k |g_k|) (sum_j |P[j,:] [I; K]|)'.
// HBM traffic: the data set once (8 (T-1)(2n+m) bytes), P once when the caller asks for it.
#include "tz_common.cuh"

namespace tz {

constexpr int kIdThreads = 128;
constexpr int kIdWarps = kIdThreads / 32;
constexpr int kMaxD = kMaxN + kMaxM;
constexpr int kMaxRed = kMaxD + kMaxN;      // simultaneous CTA-wide sums: the trailing columns of D' and the n right-hand sides

// sums `cnt` per-thread partials over the CTA; every thread gets the totals in out[]
__device__ __forceinline__ void cta_sum(const double* acc, int cnt, double (*part)[kMaxRed], double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < cnt; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) part[wid][k] = v;
  }
  __syncthreads();
  for (int k = 0; k < cnt; ++k) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kIdWarps; ++w) t += part[w][k];
    out[k] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kIdThreads) identify_kernel(int T, int n, int m, int gW, const double* __restrict__ X,
                                                              const double* __restrict__ U, const double* __restrict__ WZ,
                                                              const double* __restrict__ K, double* __restrict__ AB,
                                                              double* __restrict__ dAB, double* __restrict__ dK,
                                                              double* __restrict__ Pinv, int32_t* __restrict__ status) {
  extern __shared__ double smd[];                   // M = D': (T-1) x d | B = X1c: (T-1) x n | Y: (T-1) x d
  __shared__ double part[kIdWarps][kMaxRed];
  __shared__ double Rdiag[kMaxD], beta[kMaxD], Rinv[kMaxD * kMaxD];
  __shared__ double sP[kMaxD], sPK[kMaxN];
  __shared__ double wP[kIdThreads / 32][kMaxD], wPK[kIdThreads / 32][kMaxN];      // per-warp partial sums, combined in a fixed order
  __shared__ int bad;
  const int64_t s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int d = n + m, Tm = T - 1;
  double* M = smd;
  double* B = M + (size_t)Tm * d;
  double* Y = B + (size_t)Tm * n;
  const double* Xs = X + s * (int64_t)T * n;
  const double* Us = U + s * (int64_t)T * m;
  for (int i = tid; i < Tm * n; i += kIdThreads) {
    const int j = i / n, r = i - j * n;
    M[(size_t)j * d + r] = Xs[i];                                    // Xm = x[:-1]   (tzddpc/tzddpc.py:60)
    B[i] = Xs[i + n] - WZ[(int64_t)r * (1 + gW)];                    // Xp = x[1:] minus c_W (:61, App. A.7)
  }
  for (int i = tid; i < Tm * m; i += kIdThreads) {
    const int j = i / m, r = i - j * m;
    M[(size_t)j * d + n + r] = Us[i];                                // Um = u[:-1]   (:62)
  }
  if (tid == 0) bad = (Tm < d) ? 1 : 0;
  for (int i = tid; i < kMaxD; i += kIdThreads) { sP[i] = 0.0; Rdiag[i] = 1.0; beta[i] = 0.0; }
  for (int i = tid; i < kMaxN; i += kIdThreads) sPK[i] = 0.0;
  __syncthreads();
  // scale of the data, for the rank test
  double mscale;
  {
    double acc[1] = {0.0}, tot[1];
    for (int i = tid; i < Tm * d; i += kIdThreads) acc[0] = fma(M[i], M[i], acc[0]);
    cta_sum(acc, 1, part, tot);
    mscale = sqrt(tot[0]);
  }
  // ---- Householder QR of M (and Q' applied to B).  Reflection c: v = M[c:, c] - alpha e_c, H = I - beta v v'.
  const int ncols = d < Tm ? d : Tm;
  for (int c = 0; c < ncols; ++c) {
    double acc[kMaxRed], tot[kMaxRed];
    acc[0] = 0.0;
    for (int j = c + tid; j < Tm; j += kIdThreads) { const double v = M[(size_t)j * d + c]; acc[0] = fma(v, v, acc[0]); }
    cta_sum(acc, 1, part, tot);
    const double sigma = tot[0];
    const double mcc = M[(size_t)c * d + c];
    const double alpha = mcc >= 0.0 ? -sqrt(sigma) : sqrt(sigma);
    const double denom = sigma - mcc * alpha;                         // = v'v / 2 > 0 unless the column is zero
    const bool degenerate = !(sqrt(sigma) > 1e-13 * mscale) || !(denom > 0.0);
    const double bt = degenerate ? 0.0 : 1.0 / denom;
    __syncthreads();                                                  // (everybody has read M[c][c])
    if (tid == 0) {
      if (degenerate) bad = 1;
      M[(size_t)c * d + c] = mcc - alpha;                             // v_c
      Rdiag[c] = degenerate ? 1.0 : alpha;
      beta[c] = bt;
    }
    __syncthreads();
    // s_k = v' X[:, k] for the trailing columns of M and all columns of B
    const int nk = (d - c - 1) + n;
    for (int k = 0; k < nk; ++k) acc[k] = 0.0;
    for (int j = c + tid; j < Tm; j += kIdThreads) {
      const double vj = M[(size_t)j * d + c];
      for (int k = c + 1; k < d; ++k) acc[k - c - 1] = fma(vj, M[(size_t)j * d + k], acc[k - c - 1]);
      for (int r = 0; r < n; ++r) acc[d - c - 1 + r] = fma(vj, B[(size_t)j * n + r], acc[d - c - 1 + r]);
    }
    cta_sum(acc, nk, part, tot);
    for (int j = c + tid; j < Tm; j += kIdThreads) {
      const double bv = bt * M[(size_t)j * d + c];
      for (int k = c + 1; k < d; ++k) M[(size_t)j * d + k] = fma(-bv, tot[k - c - 1], M[(size_t)j * d + k]);
      for (int r = 0; r < n; ++r) B[(size_t)j * n + r] = fma(-bv, tot[d - c - 1 + r], B[(size_t)j * n + r]);
    }
    __syncthreads();
  }
  // ---- R^-1 (upper triangular, d x d), one column per thread:  R Rinv[:, k] = e_k
  if (tid < d) {
    const int k = tid;
    for (int i = d - 1; i >= 0; --i) {
      double t = (i == k) ? 1.0 : 0.0;
      for (int q = i + 1; q <= k; ++q) t = fma(-M[(size_t)i * d + q], Rinv[q * kMaxD + k], t);
      Rinv[i * kMaxD + k] = (i <= k) ? t / Rdiag[i] : 0.0;
    }
  }
  __syncthreads();
  // ---- AB' = R^-1 (Q'B)[:d]  (the least-squares solution of D' AB' = X1c):  AB[r][c] = sum_{q >= c} Rinv[c][q] (Q'B)[q][r]
  for (int i = tid; i < n * d; i += kIdThreads) {
    const int r = i / d, c = i - r * d;
    double t = 0.0;
    for (int q = c; q < d; ++q) t = fma(Rinv[c * kMaxD + q], B[(size_t)q * n + r], t);
    AB[s * (int64_t)n * d + i] = t;
  }
  // ---- P = Q [R^-T; 0]: Y starts as [R^-T; 0], the reflections are applied in reverse order
  for (int i = tid; i < Tm * d; i += kIdThreads) {
    const int j = i / d, k = i - j * d;
    Y[i] = (j < d && k <= j) ? Rinv[k * kMaxD + j] : 0.0;             // (R^-T)[j][k] = Rinv[k][j]
  }
  __syncthreads();
  for (int c = ncols - 1; c >= 0; --c) {
    double acc[kMaxRed], tot[kMaxRed];
    for (int k = 0; k < d; ++k) acc[k] = 0.0;
    for (int j = c + tid; j < Tm; j += kIdThreads) {
      const double vj = M[(size_t)j * d + c];
      for (int k = 0; k < d; ++k) acc[k] = fma(vj, Y[(size_t)j * d + k], acc[k]);
    }
    cta_sum(acc, d, part, tot);
    const double bt = beta[c];
    for (int j = c + tid; j < Tm; j += kIdThreads) {
      const double bv = bt * M[(size_t)j * d + c];
      for (int k = 0; k < d; ++k) Y[(size_t)j * d + k] = fma(-bv, tot[k], Y[(size_t)j * d + k]);
    }
    __syncthreads();
  }
  // ---- rows of the pseudo-inverse and their absolute column sums
  {
    double aP[kMaxD], aPK[kMaxN];
    for (int b = 0; b < kMaxD; ++b) aP[b] = 0.0;
    for (int b = 0; b < kMaxN; ++b) aPK[b] = 0.0;
    const double* Ks = K ? K + s * (int64_t)m * n : nullptr;
    for (int j = tid; j < Tm; j += kIdThreads) {
      const double* v = Y + (size_t)j * d;
      for (int b = 0; b < d; ++b) {
        aP[b] += fabs(v[b]);
        if (Pinv) Pinv[s * (int64_t)Tm * d + (int64_t)j * d + b] = v[b];
      }
      if (Ks)
        for (int c = 0; c < n; ++c) {
          double t = v[c];                                        // P[j,:] [I; K] column c
          for (int k = 0; k < m; ++k) t = fma(v[n + k], Ks[k * n + c], t);
          aPK[c] += fabs(t);
        }
    }
    // (fixed-order combination, no atomics: the box widths -- and with them every program built from this data set -- come
    //  out bit-identical on every run and every GPU, which the N-GPU == 1-GPU check of bench.py relies on)
    const int wid = tid >> 5;
    for (int b = 0; b < d; ++b) {
      const double v = warp_sum(aP[b]);
      if (lane == 0) wP[wid][b] = v;
    }
    for (int c = 0; c < n; ++c) {
      const double v = warp_sum(aPK[c]);
      if (lane == 0) wPK[wid][c] = v;
    }
  }
  __syncthreads();
  if (tid < d) {
    double t = 0.0;
    for (int w = 0; w < kIdThreads / 32; ++w) t += wP[w][tid];
    sP[tid] = t;
  }
  if (tid < n) {
    double t = 0.0;
    for (int w = 0; w < kIdThreads / 32; ++w) t += wPK[w][tid];
    sPK[tid] = t;
  }
  __syncthreads();
  for (int i = tid; i < n * d; i += kIdThreads) {
    const int r = i / d, c = i - r * d;
    double gw = 0.0;
    for (int k = 0; k < gW; ++k) gw += fabs(WZ[(int64_t)r * (1 + gW) + 1 + k]);
    dAB[s * (int64_t)n * d + i] = gw * sP[c];
  }
  if (K && dK)
    for (int i = tid; i < n * n; i += kIdThreads) {
      const int r = i / n, c = i - r * n;
      double gw = 0.0;
      for (int k = 0; k < gW; ++k) gw += fabs(WZ[(int64_t)r * (1 + gW) + 1 + k]);
      dK[s * (int64_t)n * n + i] = gw * sPK[c];
    }
  if (tid == 0 && status) status[s] = bad ? TZ_STATUS_NONFINITE : TZ_STATUS_OK;
}

}  // namespace tz

using namespace tz;

extern "C" int tz_identify(int64_t S, int32_t T, int32_t n, int32_t m, int32_t gW, const double* X, const double* U,
                           const double* WZ, const double* K, double* AB, double* dAB, double* dK, double* Pinv,
                           int32_t* status, void* stream) {
  TZ_REQUIRE(S >= 0 && T >= 2 && n >= 1 && n <= kMaxN && m >= 1 && m <= kMaxM && gW >= 0, "bad shape");
  if (S == 0) return TZ_OK;
  TZ_REQUIRE(X && U && WZ && AB && dAB, "null pointer");
  TZ_REQUIRE(!dK || K, "dK needs K");
  const size_t smem = (size_t)(T - 1) * (3 * n + 2 * m) * sizeof(double);
  TZ_REQUIRE(smem <= 200 * 1024, "dataset too long for shared memory (T=%d)", T);
  if (smem + 8192 > 48 * 1024) TZ_CUDA(cudaFuncSetAttribute(identify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  identify_kernel<<<(unsigned)S, kIdThreads, smem, (cudaStream_t)stream>>>(T, n, m, gW, X, U, WZ, K, AB, dAB, dK, Pinv, status);
  TZ_CUDA(cudaGetLastError());
  return TZ_OK;
}
