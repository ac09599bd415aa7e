// Per-scenario parameter evaluation shared by step_kernel (G lanes per scenario) and fast_step_kernel (one thread per
// scenario): both must produce bit-identical q(p), c0(p) and atoms from the same w = [1 | p | |p| | atoms].
#pragma once
#include "tz_admm.cuh"

namespace tz {

// general atoms |Bt p + gam| -> w[1 + 2 NPAR + i]
template <class BK>
__device__ __forceinline__ void eval_atoms(const QpProg<BK>& pg, double (&w)[2 * BK::NCOL2]) {
  constexpr int NPAR = BK::NPAR, NAG = BK::NAG;
#pragma unroll
  for (int i = 0; i < NAG; ++i) w[1 + 2 * NPAR + i] = 0.0;
  if (pg.nag > 0) {
#pragma unroll
    for (int i = 0; i < NAG; ++i) {
      double acc = pg.gam[i];
#pragma unroll
      for (int j = 0; j < NPAR; ++j) acc = fma(pg.Bt[i][j], w[1 + j], acc);
      w[1 + 2 * NPAR + i] = fabs(acc);
    }
  }
}

// q = q0 + Qp p
template <class BK>
__device__ __forceinline__ void eval_q(const QpProg<BK>& pg, const double (&w)[2 * BK::NCOL2], double (&q)[BK::NZ]) {
#pragma unroll
  for (int j = 0; j < BK::NZ; ++j) {
    double acc = pg.q0[j];
    if (pg.has_qp) {
#pragma unroll
      for (int k = 0; k < BK::NPAR; ++k) acc = fma(pg.Qp[j][k], w[1 + k], acc);
    }
    q[j] = acc;
  }
}

// cost constant c0(p) = cc . w + p' CC2 p
template <class BK>
__device__ __forceinline__ double cost_const(const QpProg<BK>& pg, const double (&w)[2 * BK::NCOL2]) {
  constexpr int NPAR = BK::NPAR, NCOL = BK::NCOL;
  double c0 = 0.0;
#pragma unroll
  for (int j = 0; j < NCOL; ++j) c0 = fma(pg.cc[j], w[j], c0);
  if (pg.has_cc2) {
#pragma unroll
    for (int i = 0; i < NPAR; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NPAR; ++j) acc = fma(pg.CC2[i][j], w[1 + j], acc);
      c0 = fma(acc, w[1 + i], c0);
    }
  }
  return c0;
}

// parameter-only feasibility row i:  sum_j Rchk_ij w_j <= chk_tol_i,  chk_tol_i = 1e-9 max(1, |constant of the row|)
// (the oracle's test: violated by more than 1e-9 relative to the bound, oracle/program.py:274-292)
template <class BK>
__device__ __forceinline__ bool param_row_violated(const QpProg<BK>& pg, int i, const double (&w)[2 * BK::NCOL2]) {
  const double2* Rc = reinterpret_cast<const double2*>(&pg.Rchk[i][0]);
  double r = 0.0, r1 = 0.0;
#pragma unroll
  for (int j = 0; j < BK::NCOL2; ++j) {
    const double2 c2 = Rc[j];
    r = fma(c2.x, w[2 * j], r);
    r1 = fma(c2.y, w[2 * j + 1], r1);
  }
  return r + r1 > pg.chk_tol[i];
}

// NR parameter-only rows at once (rows beyond nchk are clamped to the last one): 2 NR fma chains in flight
template <class BK, int NR>
__device__ __forceinline__ bool param_rows_violated(const QpProg<BK>& pg, int i0, const double (&w)[2 * BK::NCOL2]) {
  const double2* Rc[NR];
  double r[NR], r1[NR];
  int rows[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) {
    rows[q] = i0 + q < pg.nchk ? i0 + q : pg.nchk - 1;
    Rc[q] = reinterpret_cast<const double2*>(&pg.Rchk[rows[q]][0]);
    r[q] = 0.0;
    r1[q] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < BK::NCOL2; ++j) {
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      const double2 c2 = Rc[q][j];
      r[q] = fma(c2.x, w[2 * j], r[q]);
      r1[q] = fma(c2.y, w[2 * j + 1], r1[q]);
    }
  }
  bool bad = false;
#pragma unroll
  for (int q = 0; q < NR; ++q) bad = bad || (r[q] + r1[q] > pg.chk_tol[rows[q]]);
  return bad;
}

}  // namespace tz
