// Generic large-program path (tz_big.cu): device view, host handle and entry points shared with tz_fused.cu.
#pragma once
#include "tz_step.cuh"

namespace tz {

struct BigDev {
  int nz, nc, npar, na, ncol, nchk, n, m, N, nv, g1, n_nz, scratch_doubles;
  double cinv;
  // scaled program (zbar = z / D, Abar = E A D, Pbar = c D P D), run-time sizes, row-major
  const double *P, *A, *At, *l0, *u0, *kink0, *wabs, *R, *q0, *Qp, *Bt, *gam, *Rchk, *chk_tol, *cc, *CC2, *D, *sing_inv, *XB, *K, *ze1_val;
  const int32_t *sing_var, *ze1_ptr, *ze1_idx, *tube_ent;
};

struct BigProgram {
  BigDev dev{};
  double* dbl_dev = nullptr;
  int32_t* int_dev = nullptr;
  std::vector<int32_t> tube_ent;
};

bool big_fits(const TzProgramDesc& d);
int big_create(const TzProgramDesc& d, BigProgram** out);
void big_destroy(BigProgram* bp);
int big_launch(const BigProgram* bp, const SolverParams& sp, const StepArgs& a, cudaStream_t st);

}  // namespace tz
