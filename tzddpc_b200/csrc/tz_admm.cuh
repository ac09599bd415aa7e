// Batched ADMM QP core: one scenario per thread, the whole KKT system of that scenario in
// registers (NZ x NZ Cholesky factor, z, y) and its bounds in shared memory.
//
// Replaces `problem_full.solve(**kw)` of the reference (tzddpc/tzddpc.py:367), which hands a
// cvxpy-canonicalised cone program to a CPU interior-point solver.  Here the program is the
// parametric QP of tzddpc_b200/program.py in OSQP form, already Ruiz-scaled on the host:
//
//     min 0.5 x'Px + q'x + sum_{i<NKINK} w_i |(Ax)_i - k_i|    s.t.  l <= Ax <= u
//
// Algorithm: OSQP-style ADMM (x-update through the reduced NZ x NZ system
// P + sigma I + A' diag(rho) A, over-relaxation alpha) with
//   * the |.| cost rows handled by their prox (soft threshold) inside the z-update, so an
//     LP cost needs no slack variable,
//   * a per-row penalty rho_i switched between rho*rho_active / rho*rho_inactive by the
//     detected activity of the row, on a geometrically growing schedule (updates stop
//     changing once the active set has settled, after which this is plain ADMM),
//   * OSQP's primal-infeasibility certificate on the dual increments,
//   * a masked augmented-Lagrangian polish on the detected active set.
// The matrices (P, A, ...) are shared by all scenarios of a launch and are read straight
// from the constant bank (the program is a __grid_constant__ kernel parameter), so every
// DFMA of the inner loops takes its matrix operand without a load instruction.
#pragma once
#include "tz_common.cuh"

namespace tz {

template <int NZ_, int NC_, int NPAR_, int NA_, int NCHK_, int NKINK_, int TPB_>
struct Bucket {
  static constexpr int NZ = NZ_, NC = NC_, NPAR = NPAR_, NA = NA_, NCHK = NCHK_, NKINK = NKINK_, TPB = TPB_;
  static constexpr int NCOL = 1 + NPAR + NA;
  static constexpr int NW32 = (NC + 31) / 32;
};

// Scaled program, padded to the bucket (padding rows: A = 0, l = -inf, u = +inf).
template <class BK>
struct QpProg {
  double P[BK::NZ][BK::NZ];
  double A[BK::NC][BK::NZ];
  double l0[BK::NC], u0[BK::NC];
  double kink0[BK::NKINK], wabs[BK::NKINK];
  double R[BK::NC][BK::NCOL];
  double q0[BK::NZ];
  double Qp[BK::NZ][BK::NPAR];
  double Bt[BK::NA][BK::NPAR];
  double gam[BK::NA];
  double Rchk[BK::NCHK][BK::NCOL];
  double cc[BK::NCOL];
  double CC2[BK::NPAR][BK::NPAR];
  double D[BK::NZ];
  double Einv[BK::NC];
  double cinv;            // 1 / cost scaling
  int nz, nc, npar, na, nchk, nkink;
};

struct SolverParams {     // TzSolverOpts, device side
  double rho, rho_act, rho_inact, sigma, alpha, eps_abs, eps_rel;
  int max_iter, check_every, polish, warm;
};

template <int NZ>
__device__ __forceinline__ void chol_factor(double (&K)[NZ][NZ]) {
  // in-place lower Cholesky; the strictly upper part is ignored.  Diagonal stores 1/L_jj.
#pragma unroll
  for (int j = 0; j < NZ; ++j) {
    double d = K[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= K[j][k] * K[j][k];
    d = fmax(d, 1e-300);
    const double inv = rsqrt(d);
    K[j][j] = inv;
#pragma unroll
    for (int i = j + 1; i < NZ; ++i) {
      double s = K[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= K[i][k] * K[j][k];
      K[i][j] = s * inv;
    }
  }
}

template <int NZ>
__device__ __forceinline__ void chol_solve(const double (&L)[NZ][NZ], double (&b)[NZ]) {
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * b[k];
    b[i] = s * L[i][i];
  }
#pragma unroll
  for (int i = NZ - 1; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < NZ; ++k) s -= L[k][i] * b[k];
    b[i] = s * L[i][i];
  }
}

template <class BK>
struct RhoSet {            // the three penalty levels and their inverses
  double base, act, inact, ibase, iact, iinact;
};

template <class BK>
__device__ __forceinline__ double row_rho(const RhoSet<BK>& r, bool switched, bool active) {
  return switched ? (active ? r.act : r.inact) : r.base;
}
template <class BK>
__device__ __forceinline__ double row_irho(const RhoSet<BK>& r, bool switched, bool active) {
  return switched ? (active ? r.iact : r.iinact) : r.ibase;
}

// K = P + sigma I + sum_i rho_i a_i a_i'   (lower triangle), then factor.
template <class BK, class PG>
__device__ __forceinline__ void build_factor(const PG& pg, const RhoSet<BK>& rs, double sigma, bool switched,
                                             const uint32_t (&mask)[BK::NW32], double (&L)[BK::NZ][BK::NZ]) {
  constexpr int NZ = BK::NZ, NC = BK::NC;
#pragma unroll
  for (int a = 0; a < NZ; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) L[a][b] = pg.P[a][b] + (a == b ? sigma : 0.0);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const double rho = row_rho<BK>(rs, switched, (mask[i >> 5] >> (i & 31)) & 1u);
#pragma unroll
    for (int a = 0; a < NZ; ++a) {
      const double ra = rho * pg.A[i][a];
#pragma unroll
      for (int b = 0; b <= a; ++b) L[a][b] = fma(ra, pg.A[i][b], L[a][b]);
    }
  }
  chol_factor<NZ>(L);
}

// One scenario's ADMM solve.  lb/ub: shared memory, element i at [i * BK::TPB].
// Every thread of the warp must call this (warp votes inside); `live` = has a real scenario.
// Returns TZ_STATUS_*.  x is in the scaled space (z_unscaled = D x).
template <class BK, class PG>
__device__ int admm_solve(const PG& pg, const SolverParams& sp, bool live, const double (&q)[BK::NZ],
                          const double* __restrict__ lb, const double* __restrict__ ub,
                          const double (&kink)[BK::NKINK], double (&x)[BK::NZ], double (&z)[BK::NC],
                          double (&y)[BK::NC], uint32_t (&mask)[BK::NW32], bool warm, int& iters_out) {
  constexpr int NZ = BK::NZ, NC = BK::NC, NKINK = BK::NKINK, TPB = BK::TPB;
  RhoSet<BK> rs;
  rs.base = sp.rho; rs.act = sp.rho * sp.rho_act; rs.inact = sp.rho * sp.rho_inact;
  rs.ibase = 1.0 / rs.base; rs.iact = 1.0 / rs.act; rs.iinact = 1.0 / rs.inact;
  const double alpha = sp.alpha, sigma = sp.sigma;

  bool switched = warm;
  if (!warm) {
#pragma unroll
    for (int j = 0; j < NZ; ++j) x[j] = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) y[i] = 0.0;
#pragma unroll
    for (int w = 0; w < BK::NW32; ++w) mask[w] = 0u;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    double ax = 0.0;
#pragma unroll
    for (int j = 0; j < NZ; ++j) ax = fma(pg.A[i][j], x[j], ax);
    z[i] = fmin(fmax(ax, lb[i * TPB]), ub[i * TPB]);
  }
  double qn = 0.0;
#pragma unroll
  for (int j = 0; j < NZ; ++j) qn = fmax(qn, fabs(q[j]));

  double L[NZ][NZ];
  build_factor<BK>(pg, rs, sigma, switched, mask, L);

  int status = TZ_STATUS_MAXITER;
  bool done = !live;
  int iters = 0;
  int next_upd = 2, gap = 2;
  const int check_every = sp.check_every > 0 ? sp.check_every : 1;

  for (int k = 1; k <= sp.max_iter; ++k) {
    // ---- x-update: (P + sigma I + A' diag(rho) A) xt = sigma x - q + A'(rho z - y)
    double xt[NZ];
#pragma unroll
    for (int j = 0; j < NZ; ++j) xt[j] = sigma * x[j] - q[j];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const double rho = row_rho<BK>(rs, switched, (mask[i >> 5] >> (i & 31)) & 1u);
      const double t = rho * z[i] - y[i];
#pragma unroll
      for (int j = 0; j < NZ; ++j) xt[j] = fma(pg.A[i][j], t, xt[j]);
    }
    chol_solve<NZ>(L, xt);
    const bool check = (k % check_every == 0) || (k == next_upd) || (k == sp.max_iter);
    // ---- z / y update (with the prox of the |.| rows), dual increment statistics on check iterations
    double xn[NZ];
#pragma unroll
    for (int j = 0; j < NZ; ++j) xn[j] = alpha * xt[j] + (1.0 - alpha) * x[j];
    double atdy[NZ];
    double dy_norm = 0.0, supp = 0.0, dy_inf = 0.0;
#pragma unroll
    for (int j = 0; j < NZ; ++j) atdy[j] = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const bool act = (mask[i >> 5] >> (i & 31)) & 1u;
      const double rho = row_rho<BK>(rs, switched, act);
      const double irho = row_irho<BK>(rs, switched, act);
      double zt = 0.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j) zt = fma(pg.A[i][j], xt[j], zt);
      const double zr = alpha * zt + (1.0 - alpha) * z[i];
      double xi = fma(y[i], irho, zr);
      if (i < NKINK) {                       // prox of w|z - k| : soft threshold around the kink
        const double d = xi - kink[i < NKINK ? i : 0];
        const double th = pg.wabs[i < NKINK ? i : 0] * irho;
        xi = kink[i < NKINK ? i : 0] + copysign(fmax(fabs(d) - th, 0.0), d);
      }
      const double l = lb[i * TPB], u = ub[i * TPB];
      const double zn = fmin(fmax(xi, l), u);
      const double dy = rho * (zr - zn);
      if (!done) {
        y[i] += dy;
        z[i] = zn;
      }
      if (check) {
        dy_norm = fmax(dy_norm, fabs(dy));
#pragma unroll
        for (int j = 0; j < NZ; ++j) atdy[j] = fma(pg.A[i][j], dy, atdy[j]);
        // support function of [l, u] along dy; a component pushing against an infinite bound
        // must vanish for dy to be a certificate (rounding leaves O(eps) residues there)
        if (dy > 0.0) { if (u < 1e300) supp = fma(u, dy, supp); else dy_inf = fmax(dy_inf, dy); }
        if (dy < 0.0) { if (l > -1e300) supp = fma(l, dy, supp); else dy_inf = fmax(dy_inf, -dy); }
      }
    }
    if (!done) {
#pragma unroll
      for (int j = 0; j < NZ; ++j) x[j] = xn[j];
      iters = k;
    }
    if (check) {
      // ---- residuals (scaled space)
      double rp = 0.0, axn = 0.0, zn_ = 0.0;
      double aty[NZ];
#pragma unroll
      for (int j = 0; j < NZ; ++j) aty[j] = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        double ax = 0.0;
#pragma unroll
        for (int j = 0; j < NZ; ++j) ax = fma(pg.A[i][j], x[j], ax);
        rp = fmax(rp, fabs(ax - z[i]));
        axn = fmax(axn, fabs(ax));
        zn_ = fmax(zn_, fabs(z[i]));
#pragma unroll
        for (int j = 0; j < NZ; ++j) aty[j] = fma(pg.A[i][j], y[i], aty[j]);
      }
      double rd = 0.0, pxn = 0.0, atyn = 0.0, atdyn = 0.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j) {
        double px = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) px = fma(pg.P[j][b], x[b], px);
        rd = fmax(rd, fabs(px + q[j] + aty[j]));
        pxn = fmax(pxn, fabs(px));
        atyn = fmax(atyn, fabs(aty[j]));
        atdyn = fmax(atdyn, fabs(atdy[j]));
      }
      const double ep = sp.eps_abs + sp.eps_rel * fmax(axn, zn_);
      const double ed = sp.eps_abs + sp.eps_rel * fmax(fmax(pxn, atyn), qn);
      if (!done) {
        if (!(rp == rp) || !(rd == rd)) {
          status = TZ_STATUS_NONFINITE;
          done = true;
        } else if (rp <= ep && rd <= ed) {
          status = TZ_STATUS_OK;
          done = true;
        } else if (k >= 10 && dy_norm > 1e-12 && dy_inf <= 1e-6 * dy_norm && atdyn <= 1e-6 * dy_norm &&
                   supp < -1e-6 * dy_norm) {
          status = TZ_STATUS_INFEASIBLE;     // Farkas certificate (OSQP, Banjac et al. 2019)
          done = true;
        }
      }
      if (__all_sync(0xffffffffu, done)) break;
    }
    // ---- activity-driven rho switch on a geometric schedule
    if (k == next_upd) {
      gap = (gap * 3 + 1) / 2;
      next_upd = k + gap;
      uint32_t nm[BK::NW32];
#pragma unroll
      for (int w = 0; w < BK::NW32; ++w) nm[w] = 0u;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        bool act = (z[i] <= lb[i * TPB]) || (z[i] >= ub[i * TPB]);
        if (i < NKINK) act = act || (pg.wabs[i < NKINK ? i : 0] > 0.0 && z[i] == kink[i < NKINK ? i : 0]);
        nm[i >> 5] |= (act ? 1u : 0u) << (i & 31);
      }
      bool changed = !switched;
#pragma unroll
      for (int w = 0; w < BK::NW32; ++w) changed = changed || (nm[w] != mask[w]);
      if (changed && !done) {
#pragma unroll
        for (int w = 0; w < BK::NW32; ++w) mask[w] = nm[w];
        switched = true;
        build_factor<BK>(pg, rs, sigma, switched, mask, L);
      }
    }
  }
  iters_out = iters;
  return live ? status : TZ_STATUS_OK;
}

// Polish: solve the equality-constrained QP on the detected active set with a masked
// augmented-Lagrangian iteration (only NZ x NZ systems, no index compaction, branch-free):
//   K = P + delta I + mu sum_{i active} a_i a_i',  x <- K^{-1}(delta x - qt + sum_act a_i (mu b_i - lam_i)),
//   lam_i <- lam_i + mu (a_i x - b_i).   Accepted only if the result is feasible.
template <class BK, class PG>
__device__ bool admm_polish(const PG& pg, const double (&q)[BK::NZ], const double* __restrict__ lb,
                            const double* __restrict__ ub, const double (&kink)[BK::NKINK], double (&x)[BK::NZ],
                            const double (&z)[BK::NC], double (&y)[BK::NC]) {
  constexpr int NZ = BK::NZ, NC = BK::NC, NKINK = BK::NKINK, TPB = BK::TPB;
  const double delta = 1e-9, mu = 1e6;
  double L[NZ][NZ], qt[NZ], xk[NZ];
  uint32_t act[BK::NW32];
#pragma unroll
  for (int w = 0; w < BK::NW32; ++w) act[w] = 0u;
#pragma unroll
  for (int a = 0; a < NZ; ++a) {
    qt[a] = q[a];
    xk[a] = x[a];
#pragma unroll
    for (int b = 0; b <= a; ++b) L[a][b] = pg.P[a][b] + (a == b ? delta : 0.0);
  }
  // active targets b_i are recomputed on the fly: lower bound, upper bound or kink
  auto target = [&](int i, bool& is_act) -> double {
    const double l = lb[i * TPB], u = ub[i * TPB];
    double b = 0.0;
    is_act = false;
    if (z[i] <= l) { b = l; is_act = true; }
    else if (z[i] >= u) { b = u; is_act = true; }
    else if (i < NKINK) {
      const double kk = kink[i < NKINK ? i : 0];
      if (pg.wabs[i < NKINK ? i : 0] > 0.0 && z[i] == kk) { b = kk; is_act = true; }
    }
    return b;
  };
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    bool ia;
    (void)target(i, ia);
    if (ia) {
      act[i >> 5] |= 1u << (i & 31);
#pragma unroll
      for (int a = 0; a < NZ; ++a) {
        const double ra = mu * pg.A[i][a];
#pragma unroll
        for (int b = 0; b <= a; ++b) L[a][b] = fma(ra, pg.A[i][b], L[a][b]);
      }
    } else {
      y[i] = 0.0;
      if (i < NKINK) {           // |.| row away from its kink: a linear cost term
        const double w = pg.wabs[i < NKINK ? i : 0];
        if (w > 0.0) {
          const double sg = z[i] > kink[i < NKINK ? i : 0] ? w : -w;
          y[i] = sg;             // its multiplier is the subgradient
#pragma unroll
          for (int a = 0; a < NZ; ++a) qt[a] = fma(sg, pg.A[i][a], qt[a]);
        }
      }
    }
  }
  chol_factor<NZ>(L);
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    double rhs[NZ];
#pragma unroll
    for (int a = 0; a < NZ; ++a) rhs[a] = delta * xk[a] - qt[a];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if ((act[i >> 5] >> (i & 31)) & 1u) {
        bool ia;
        const double t = mu * target(i, ia) - y[i];
#pragma unroll
        for (int a = 0; a < NZ; ++a) rhs[a] = fma(pg.A[i][a], t, rhs[a]);
      }
    }
    chol_solve<NZ>(L, rhs);
#pragma unroll
    for (int a = 0; a < NZ; ++a) xk[a] = rhs[a];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if ((act[i >> 5] >> (i & 31)) & 1u) {
        double ax = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ax = fma(pg.A[i][a], xk[a], ax);
        bool ia;
        y[i] = fma(mu, ax - target(i, ia), y[i]);
      }
    }
  }
  // accept only a feasible polished point
  double viol = 0.0, scale = 1.0;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    double ax = 0.0;
#pragma unroll
    for (int a = 0; a < NZ; ++a) ax = fma(pg.A[i][a], xk[a], ax);
    viol = fmax(viol, fmax(lb[i * TPB] - ax, ax - ub[i * TPB]));
    scale = fmax(scale, fabs(ax));
  }
  bool ok = (viol <= 1e-9 * scale);
#pragma unroll
  for (int a = 0; a < NZ; ++a) ok = ok && (xk[a] == xk[a]);
  if (ok) {
#pragma unroll
    for (int a = 0; a < NZ; ++a) x[a] = xk[a];
  }
  return ok;
}

}  // namespace tz
