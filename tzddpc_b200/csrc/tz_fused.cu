// Fused per-step kernel of the TZDDPC hot path: bounds of the parametric program from
// (xbar0, e0)  ->  ADMM + polish  ->  nominal trajectory, cost, Ze[1].Z  ->  closed-loop update.
//
// One scenario per thread, scenario-fastest (SoA) global layout so that every load/store of
// a warp is one or two full 128-byte lines.  Replaces, per closed-loop step,
//   tzddpc/tzddpc.py:357-377  (TZDDPC.solve: parameter update + cvxpy solve + Ze[1])
//   examples/2.pulley_sim.py:90-96 (nominal/plant/error update, Zek.Z.value)
// Algorithmic HBM bytes per scenario-step:
//   8*[6n + N*m + (N+1)n + 1 + n(1+g1)] + 4.
#include <new>
#include <vector>

#include "tz_admm.cuh"

namespace tz {

struct Aux {                       // device-resident tables with run-time sizes
  const double* XB;                // (N+1)n x (1+nv+npar)
  const int32_t* ze1_ptr;          // n(1+g1)+1
  const int32_t* ze1_idx;
  const double* ze1_val;
  const double* K;                 // m x n
  int n, m, N, nv, g1, nw;         // nw = 1 + nv + npar
};

struct StepArgs {
  int64_t S;                       // scenarios in this launch
  int64_t ld;                      // leading dimension of every SoA array (>= S)
  const double* xbar0;             // parameters (n x S); aliases xbar/e in closed loop
  const double* e0;
  double* x;                       // closed loop only (NULL = solve only)
  double* xbar;
  double* e;
  const double* noise;
  const double* A_true;
  const double* B_true;
  double* cost;
  double* v;
  double* xbar_traj;
  double* ze1;
  double* u_out;
  int32_t* status;
  int32_t* iters;
  double* warm;
  double* stats;
  // explicit-instance mode (tz_qp_solve): q, l, u given, z / y returned
  const double* q_in;
  const double* l_in;
  const double* u_in;
  double* z_out;
  double* y_out;
};

template <class BK, class PG>
__device__ __forceinline__ void step_body(const PG& pg, const Aux& ax, const SolverParams& sp, const StepArgs& a,
                                          double* smem) {
  constexpr int NZ = BK::NZ, NC = BK::NC, NPAR = BK::NPAR, NA = BK::NA, NCHK = BK::NCHK, NKINK = BK::NKINK,
                NCOL = BK::NCOL, TPB = BK::TPB;
  const int tid = threadIdx.x;
  const int64_t s = (int64_t)blockIdx.x * TPB + tid;
  const int64_t S = a.ld;          // SoA leading dimension: element (i, s) at [i * S + s]
  const bool live = s < a.S;
  double* lb = smem + tid;
  double* ub = smem + (size_t)NC * TPB + tid;
  const int n = ax.n, m = ax.m;
  const bool explicit_qp = a.q_in != nullptr;

  double q[NZ], kink[NKINK];
  double p[NPAR];
  double c0 = 0.0;
  bool param_ok = true, finite = true;

  if (!explicit_qp) {
    // ---- parameters p = [xbar0; e0]
#pragma unroll
    for (int j = 0; j < NPAR; ++j) {
      double val = 0.0;
      if (live && j < 2 * n) val = (j < n) ? a.xbar0[(int64_t)j * S + s] : a.e0[(int64_t)(j - n) * S + s];
      p[j] = val;
      finite = finite && (fabs(val) < 1e300);
    }
    // ---- w = [1; p; alpha(p)], alpha_j = |Bt_j p + gam_j|
    double w[NCOL];
    w[0] = 1.0;
#pragma unroll
    for (int j = 0; j < NPAR; ++j) w[1 + j] = p[j];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double acc = pg.gam[i];
#pragma unroll
      for (int j = 0; j < NPAR; ++j) acc = fma(pg.Bt[i][j], p[j], acc);
      w[1 + NPAR + i] = fabs(acc);
    }
    // ---- bounds l = l0 + R w, u = u0 + R w (scaled), kinks, q, parameter-only checks, cost constant
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) r = fma(pg.R[i][j], w[j], r);
      lb[(size_t)i * TPB] = pg.l0[i] + r;
      ub[(size_t)i * TPB] = pg.u0[i] + r;
      if (i < NKINK) kink[i < NKINK ? i : 0] = pg.kink0[i < NKINK ? i : 0] + r;
    }
#pragma unroll
    for (int j = 0; j < NZ; ++j) {
      double acc = pg.q0[j];
#pragma unroll
      for (int k = 0; k < NPAR; ++k) acc = fma(pg.Qp[j][k], p[k], acc);
      q[j] = acc;
    }
#pragma unroll
    for (int i = 0; i < NCHK; ++i) {
      double r = 0.0, sc = 1.0;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        r = fma(pg.Rchk[i][j], w[j], r);
        sc = fmax(sc, fabs(pg.Rchk[i][j] * w[j]));
      }
      param_ok = param_ok && (r <= 1e-9 * sc);
    }
#pragma unroll
    for (int j = 0; j < NCOL; ++j) c0 = fma(pg.cc[j], w[j], c0);
#pragma unroll
    for (int i = 0; i < NPAR; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NPAR; ++j) acc = fma(pg.CC2[i][j], p[j], acc);
      c0 = fma(acc, p[i], c0);
    }
  } else {
    // explicit instance: scale the caller's q, l, u  (qbar = c D q, lbar = E l)
#pragma unroll
    for (int j = 0; j < NZ; ++j) q[j] = (live && j < pg.nz) ? a.q_in[(int64_t)j * S + s] * pg.D[j] / pg.cinv : 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const bool rr = live && i < pg.nc;
      lb[(size_t)i * TPB] = rr ? a.l_in[(int64_t)i * S + s] / pg.Einv[i] : -INFINITY;
      ub[(size_t)i * TPB] = rr ? a.u_in[(int64_t)i * S + s] / pg.Einv[i] : INFINITY;
      if (i < NKINK) kink[i < NKINK ? i : 0] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NPAR; ++j) p[j] = 0.0;
  }

  // ---- ADMM (+ polish)
  double x[NZ], z[NC], y[NC];
  uint32_t mask[BK::NW32];
  bool warm = false;
  if (sp.warm && a.warm != nullptr && live) {
    // layout: [x (NZ) | y (NC) | mask words | valid flag] x S
    const double flag = a.warm[(int64_t)(NZ + NC + BK::NW32) * S + s];
    warm = (flag == 1.0);
    if (warm) {
#pragma unroll
      for (int j = 0; j < NZ; ++j) x[j] = a.warm[(int64_t)j * S + s];
#pragma unroll
      for (int i = 0; i < NC; ++i) y[i] = a.warm[(int64_t)(NZ + i) * S + s];
#pragma unroll
      for (int wd = 0; wd < BK::NW32; ++wd)
        mask[wd] = (uint32_t)__double_as_longlong(a.warm[(int64_t)(NZ + NC + wd) * S + s]);
    }
  }
  const bool solve_it = live && param_ok && finite;
  int iters = 0;
  int status = admm_solve<BK>(pg, sp, solve_it, q, lb, ub, kink, x, z, y, mask, warm, iters);
  if (live && !finite) status = TZ_STATUS_NONFINITE;
  else if (live && !param_ok) status = TZ_STATUS_INFEASIBLE;
  const bool good = live && (status == TZ_STATUS_OK || status == TZ_STATUS_MAXITER);
  if (good && a.warm != nullptr) {
#pragma unroll
    for (int j = 0; j < NZ; ++j) a.warm[(int64_t)j * S + s] = x[j];
#pragma unroll
    for (int i = 0; i < NC; ++i) a.warm[(int64_t)(NZ + i) * S + s] = y[i];
#pragma unroll
    for (int wd = 0; wd < BK::NW32; ++wd)
      a.warm[(int64_t)(NZ + NC + wd) * S + s] = __longlong_as_double((long long)mask[wd]);
    a.warm[(int64_t)(NZ + NC + BK::NW32) * S + s] = 1.0;
  }
  if (good && sp.polish) (void)admm_polish<BK>(pg, q, lb, ub, kink, x, z, y);

  if (live && a.status) a.status[s] = status;
  if (live && a.iters) a.iters[s] = iters;

  if (explicit_qp) {
    if (live) {
      if (a.z_out) {
#pragma unroll
        for (int j = 0; j < NZ; ++j)
          if (j < pg.nz) a.z_out[(int64_t)j * S + s] = good ? pg.D[j] * x[j] : NAN;
      }
      if (a.y_out) {
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i < pg.nc) a.y_out[(int64_t)i * S + s] = good ? y[i] * pg.cinv / pg.Einv[i] : NAN;
      }
    }
    return;
  }

  // ---- objective value (reference `result`, tzddpc/tzddpc.py:367,377; constant terms included, quirk Q7)
  double cost = NAN;
  if (good) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NZ; ++j) {
      double px = 0.0;
#pragma unroll
      for (int b = 0; b < NZ; ++b) px = fma(pg.P[j][b], x[b], px);
      acc = fma(0.5 * px + q[j], x[j], acc);
    }
#pragma unroll
    for (int i = 0; i < NKINK; ++i) {
      double axv = 0.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j) axv = fma(pg.A[i][j], x[j], axv);
      acc = fma(pg.wabs[i], fabs(axv - kink[i]), acc);
    }
    cost = fma(acc, pg.cinv, c0);
  } else if (live && status == TZ_STATUS_INFEASIBLE) {
    cost = INFINITY;                      // cvxpy returns +inf for an infeasible Minimize (:374)
  }
  if (live && a.cost) a.cost[s] = cost;

  // ---- om = [1; v; p] drives every affine output
  double om[1 + 16 + 2 * kMaxN];
  const int nv = ax.nv, nw = ax.nw;
  om[0] = 1.0;
#pragma unroll
  for (int j = 0; j < NZ; ++j)
    if (j < nv) om[1 + j] = good ? pg.D[j] * x[j] : NAN;
#pragma unroll
  for (int j = 0; j < NPAR; ++j)
    if (j < 2 * n) om[1 + nv + j] = p[j];
  if (live && a.v)
    for (int j = 0; j < nv; ++j) a.v[(int64_t)j * S + s] = om[1 + j];

  // nominal trajectory xbar_0..xbar_N = XB om  (tzddpc/tzddpc.py:166-170)
  double xbar1[kMaxN];
#pragma unroll
  for (int i = 0; i < kMaxN; ++i) xbar1[i] = 0.0;
  const int nrows = (ax.N + 1) * n;
  if (live) {
    for (int i = 0; i < nrows; ++i) {
      const double* row = ax.XB + (size_t)i * nw;
      double acc = 0.0;
      for (int j = 0; j < nw; ++j) acc = fma(__ldg(row + j), om[j], acc);
      if (a.xbar_traj) a.xbar_traj[(int64_t)i * S + s] = acc;
      const int r = i - n;
#pragma unroll
      for (int k = 0; k < kMaxN; ++k)
        if (k == r) xbar1[k] = acc;
    }
  }
  // Ze[1].Z at the optimum (examples/2.pulley_sim.py:96: Zek.Z.value), dense n x (1+g1)
  if (live && a.ze1) {
    const int nent = n * (1 + ax.g1);
    int t0 = __ldg(ax.ze1_ptr);
    for (int en = 0; en < nent; ++en) {
      const int t1 = __ldg(ax.ze1_ptr + en + 1);
      double acc = 0.0;
      for (int t = t0; t < t1; ++t) acc = fma(__ldg(ax.ze1_val + t), om[__ldg(ax.ze1_idx + t)], acc);
      t0 = t1;
      __stcs(a.ze1 + (int64_t)en * S + s, acc);      // streaming store: written once, never re-read here
    }
  }

  // ---- closed-loop update (examples/2.pulley_sim.py:90-94)
  if (a.x != nullptr) {
    double xs[kMaxN], es[kMaxN], us[kMaxM], xn[kMaxN];
    double nrm2 = 0.0;
    bool viol = false;
    if (live) {
#pragma unroll
      for (int i = 0; i < kMaxN; ++i) {
        xs[i] = (i < n) ? a.x[(int64_t)i * S + s] : 0.0;
        es[i] = (i < n) ? om[1 + nv + n + i] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        double acc = 0.0;
        if (j < m) {
          acc = om[1 + j];                                          // v[0]
#pragma unroll
          for (int i = 0; i < kMaxN; ++i)
            if (i < n) acc = fma(__ldg(ax.K + j * n + i), es[i], acc);
        }
        us[j] = acc;                                                // u = K e + v[0]
      }
#pragma unroll
      for (int i = 0; i < kMaxN; ++i) {
        double acc = 0.0;
        if (i < n) {
          acc = a.noise ? a.noise[(int64_t)i * S + s] : 0.0;
#pragma unroll
          for (int k = 0; k < kMaxN; ++k)
            if (k < n) acc = fma(__ldg(a.A_true + i * n + k), xs[k], acc);
#pragma unroll
          for (int k = 0; k < kMaxM; ++k)
            if (k < m) acc = fma(__ldg(a.B_true + i * m + k), us[k], acc);
        }
        xn[i] = acc;                                                // x+ = A x + B u + w
      }
      if (good) {
#pragma unroll
        for (int i = 0; i < kMaxN; ++i)
          if (i < n) {
            a.x[(int64_t)i * S + s] = xn[i];
            a.xbar[(int64_t)i * S + s] = xbar1[i];                 // xbar+ = xbar_traj[1]
            a.e[(int64_t)i * S + s] = xn[i] - xbar1[i];            // e+ = x+ - xbar+
            nrm2 = fma(xn[i], xn[i], nrm2);
          }
        if (a.u_out)
#pragma unroll
          for (int j = 0; j < kMaxM; ++j)
            if (j < m) a.u_out[(int64_t)j * S + s] = us[j];
      }
    }
    (void)viol;
    if (a.stats != nullptr) {
      double st[TZ_NSTATS];
      st[0] = good ? sqrt(nrm2) : 0.0;
      st[1] = good ? nrm2 : 0.0;
      st[2] = good ? cost : 0.0;
      st[3] = (live && status == TZ_STATUS_INFEASIBLE) ? 1.0 : 0.0;
      st[4] = (live && status == TZ_STATUS_MAXITER) ? 1.0 : 0.0;
      st[5] = live ? (double)iters : 0.0;
      st[6] = (live && status == TZ_STATUS_NONFINITE) ? 1.0 : 0.0;
      st[7] = live ? 1.0 : 0.0;
      __syncthreads();                         // smem (bounds) is dead from here on
      double* red = smem;                      // [TZ_NSTATS][TPB/32]
#pragma unroll
      for (int k = 0; k < TZ_NSTATS; ++k) {
        const double v_ = warp_sum(st[k]);
        if ((tid & 31) == 0) red[k * (TPB / 32) + (tid >> 5)] = v_;
      }
      __syncthreads();
      if (tid < TZ_NSTATS) {
        double acc = 0.0;
        for (int wv = 0; wv < TPB / 32; ++wv) acc += red[tid * (TPB / 32) + wv];
        atomicAdd(a.stats + tid, acc);
      }
    }
  }
}

template <class BK>
__global__ void __launch_bounds__(BK::TPB) step_kernel_param(const __grid_constant__ QpProg<BK> pg, const Aux ax,
                                                             const SolverParams sp, const StepArgs a) {
  extern __shared__ double smem[];
  step_body<BK>(pg, ax, sp, a, smem);
}

template <class BK>
__global__ void __launch_bounds__(BK::TPB) step_kernel_global(const QpProg<BK>* __restrict__ pg, const Aux ax,
                                                              const SolverParams sp, const StepArgs a) {
  extern __shared__ double smem[];
  step_body<BK>(*pg, ax, sp, a, smem);
}

// ---- compiled buckets: <NZ, NC, NPAR, NA, NCHK, NKINK, TPB> -------------------------------------
using B0 = Bucket<2, 16, 4, 6, 4, 2, 128>;       // double integrator, N = 2
using B1 = Bucket<2, 24, 8, 10, 8, 2, 128>;      // pulley (n = 4), N = 2
using B2 = Bucket<2, 28, 10, 12, 12, 2, 128>;    // 5-dim, N = 2
using B3 = Bucket<4, 40, 16, 20, 16, 4, 128>;    // generic small
using B4 = Bucket<8, 72, 16, 40, 16, 8, 64>;     // generic medium (N = 3..4)
using B5 = Bucket<16, 128, 16, 72, 16, 8, 32>;   // generic large; program read from global memory

template <class BK>
constexpr bool fits(int nz, int nc, int npar, int na, int nchk, int nkink) {
  return nz <= BK::NZ && nc <= BK::NC && npar <= BK::NPAR && na <= BK::NA && nchk <= BK::NCHK && nkink <= BK::NKINK;
}

template <class BK>
void pack(const TzProgramDesc& d, QpProg<BK>& g) {
  std::memset(&g, 0, sizeof(g));
  const int nz = d.nz, nc = d.nc, npar = d.npar, na = d.na, ncol = 1 + npar + na;
  auto colmap = [&](int j) { return j <= npar ? j : 1 + BK::NPAR + (j - 1 - npar); };   // [1 | p | alpha] -> padded slot
  for (int a = 0; a < BK::NZ; ++a) g.D[a] = 1.0;
  for (int i = 0; i < BK::NC; ++i) { g.l0[i] = -INFINITY; g.u0[i] = INFINITY; g.Einv[i] = 1.0; }
  for (int a = 0; a < nz; ++a) {
    g.D[a] = d.D[a];
    g.q0[a] = d.c * d.D[a] * d.q0[a];
    for (int b = 0; b < nz; ++b) g.P[a][b] = d.c * d.D[a] * d.P[a * nz + b] * d.D[b];
    for (int k = 0; k < npar; ++k) g.Qp[a][k] = d.c * d.D[a] * d.Qp[a * npar + k];
  }
  for (int i = 0; i < nc; ++i) {
    const double E = d.E[i];
    g.Einv[i] = 1.0 / E;
    for (int a = 0; a < nz; ++a) g.A[i][a] = E * d.A[i * nz + a] * d.D[a];
    g.l0[i] = E * d.l0[i];
    g.u0[i] = E * d.u0[i];
    for (int j = 0; j < ncol; ++j) g.R[i][colmap(j)] = E * d.R[i * ncol + j];
    if (i < d.nkink) {
      g.kink0[i] = E * d.kink0[i];
      g.wabs[i] = d.c * d.wabs[i] / E;
    }
  }
  for (int i = 0; i < na; ++i) {
    g.gam[i] = d.gam[i];
    for (int k = 0; k < npar; ++k) g.Bt[i][k] = d.Bt[i * npar + k];
  }
  for (int i = 0; i < d.nchk; ++i)
    for (int j = 0; j < ncol; ++j) g.Rchk[i][colmap(j)] = d.Rchk[i * ncol + j];
  for (int j = 0; j < ncol; ++j) g.cc[colmap(j)] = d.cc[j];
  for (int a = 0; a < npar; ++a)
    for (int b = 0; b < npar; ++b) g.CC2[a][b] = d.CC2[a * npar + b];
  g.cinv = 1.0 / d.c;
  g.nz = nz; g.nc = nc; g.npar = npar; g.na = na; g.nchk = d.nchk; g.nkink = d.nkink;
}

}  // namespace tz

using namespace tz;

struct TzProgram {
  int bucket = -1;
  std::vector<unsigned char> packed;     // host image of QpProg<bucket>
  void* packed_dev = nullptr;            // device image (always kept; used by the global-memory bucket)
  void* aux_dev = nullptr;               // one allocation holding XB | ze1_val | K | ze1_ptr | ze1_idx
  Aux aux{};
  int nz = 0, nc = 0, n = 0, m = 0, N = 0, nv = 0, g1 = 0, npar = 0;
  int nw32 = 0, NZ = 0, NC = 0;
};

template <class BK>
static int create_bucket(const TzProgramDesc& d, TzProgram* p, int id) {
  p->bucket = id;
  p->packed.resize(sizeof(QpProg<BK>));
  pack<BK>(d, *reinterpret_cast<QpProg<BK>*>(p->packed.data()));
  p->nw32 = BK::NW32;
  p->NZ = BK::NZ;
  p->NC = BK::NC;
  TZ_CUDA(cudaMalloc(&p->packed_dev, sizeof(QpProg<BK>)));
  TZ_CUDA(cudaMemcpy(p->packed_dev, p->packed.data(), sizeof(QpProg<BK>), cudaMemcpyHostToDevice));
  return TZ_OK;
}

extern "C" int tz_program_create(const TzProgramDesc* d, TzProgram** out) {
  TZ_REQUIRE(d && out, "null argument");
  TZ_REQUIRE(d->n >= 1 && d->n <= kMaxN && d->m >= 1 && d->m <= kMaxM, "dim_x must be 1..%d and dim_u 1..%d", kMaxN, kMaxM);
  TZ_REQUIRE(d->npar == 2 * d->n, "npar must be 2*dim_x");
  TZ_REQUIRE(d->nv == d->horizon * d->m && d->nv <= 16 && d->nz >= d->nv, "bad nv/nz");
  TZ_REQUIRE(d->nkink >= 0 && d->nkink <= d->nc, "bad nkink");
  TzProgram* p = new (std::nothrow) TzProgram();
  if (!p) return fail(TZ_ENOMEM, "out of host memory");
  int rc = TZ_ERANGE;
#define TZ_TRY(BK, ID)                                                               \
  if (rc == TZ_ERANGE && fits<BK>(d->nz, d->nc, d->npar, d->na, d->nchk, d->nkink)) rc = create_bucket<BK>(*d, p, ID);
  TZ_TRY(B0, 0) TZ_TRY(B1, 1) TZ_TRY(B2, 2) TZ_TRY(B3, 3) TZ_TRY(B4, 4) TZ_TRY(B5, 5)
#undef TZ_TRY
  if (rc != TZ_OK) {
    delete p;
    if (rc == TZ_ERANGE)
      return fail(TZ_ERANGE, "program (nz=%d nc=%d npar=%d na=%d nchk=%d nkink=%d) exceeds every compiled bucket",
                  d->nz, d->nc, d->npar, d->na, d->nchk, d->nkink);
    return rc;
  }
  p->nz = d->nz; p->nc = d->nc; p->n = d->n; p->m = d->m; p->N = d->horizon; p->nv = d->nv; p->g1 = d->g1; p->npar = d->npar;
  const int nw = 1 + d->nv + d->npar;
  const size_t nXB = (size_t)(d->horizon + 1) * d->n * nw, nval = (size_t)d->nterms, nK = (size_t)d->m * d->n;
  const size_t nptr = (size_t)d->n * (1 + d->g1) + 1;
  const size_t bytes = (nXB + nval + nK) * sizeof(double) + (nptr + nval) * sizeof(int32_t);
  std::vector<unsigned char> host(bytes);
  double* hd = reinterpret_cast<double*>(host.data());
  std::memcpy(hd, d->XB, nXB * sizeof(double));
  if (nval) std::memcpy(hd + nXB, d->ze1_val, nval * sizeof(double));
  std::memcpy(hd + nXB + nval, d->K, nK * sizeof(double));
  int32_t* hi = reinterpret_cast<int32_t*>(hd + nXB + nval + nK);
  std::memcpy(hi, d->ze1_ptr, nptr * sizeof(int32_t));
  if (nval) std::memcpy(hi + nptr, d->ze1_idx, nval * sizeof(int32_t));
  for (size_t t = 0; t < nval; ++t)
    if (hi[nptr + t] < 0 || hi[nptr + t] >= nw) {
      tz_program_destroy(p);
      return fail(TZ_EINVAL, "ze1_idx[%zu] out of range", t);
    }
  cudaError_t err = cudaMalloc(&p->aux_dev, bytes);
  if (err == cudaSuccess) err = cudaMemcpy(p->aux_dev, host.data(), bytes, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    tz_program_destroy(p);
    return fail(TZ_ECUDA, "aux upload: %s", cudaGetErrorString(err));
  }
  double* dd = reinterpret_cast<double*>(p->aux_dev);
  int32_t* di = reinterpret_cast<int32_t*>(dd + nXB + nval + nK);
  p->aux = Aux{dd, di, di + nptr, dd + nXB, dd + nXB + nval, d->n, d->m, d->horizon, d->nv, d->g1, nw};
  *out = p;
  return TZ_OK;
}

extern "C" void tz_program_destroy(TzProgram* p) {
  if (!p) return;
  if (p->packed_dev) cudaFree(p->packed_dev);
  if (p->aux_dev) cudaFree(p->aux_dev);
  delete p;
}

extern "C" int tz_program_bucket(const TzProgram* p, char* buf, size_t cap) {
  TZ_REQUIRE(p && buf && cap > 0, "null argument");
  snprintf(buf, cap, "B%d(NZ=%d,NC=%d)", p->bucket, p->NZ, p->NC);
  return p->bucket;
}

extern "C" void tz_solver_opts_default(TzSolverOpts* o) {
  if (!o) return;
  o->rho = 0.1; o->rho_active = 100.0; o->rho_inactive = 0.1; o->sigma = 1e-6; o->alpha = 1.6;
  o->eps_abs = 1e-6; o->eps_rel = 1e-6; o->max_iter = 4000; o->check_every = 4; o->polish = 1; o->warm_start = 0;
}

static SolverParams to_params(const TzSolverOpts* o) {
  TzSolverOpts d;
  tz_solver_opts_default(&d);
  if (o) d = *o;
  return SolverParams{d.rho, d.rho_active, d.rho_inactive, d.sigma, d.alpha, d.eps_abs, d.eps_rel,
                      d.max_iter, d.check_every, d.polish, d.warm_start};
}

template <class BK, bool PARAM>
static int launch_bucket(const TzProgram* p, const SolverParams& sp, const StepArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)2 * BK::NC * BK::TPB * sizeof(double);
  const unsigned grid = (unsigned)((a.S + BK::TPB - 1) / BK::TPB);
  if constexpr (PARAM) {
    static bool configured = false;     // benign race: the attribute is idempotent
    if (!configured) {
      TZ_CUDA(cudaFuncSetAttribute(step_kernel_param<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    step_kernel_param<BK><<<grid, BK::TPB, smem, st>>>(*reinterpret_cast<const QpProg<BK>*>(p->packed.data()), p->aux, sp, a);
  } else {
    static bool configured = false;
    if (!configured) {
      TZ_CUDA(cudaFuncSetAttribute(step_kernel_global<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    step_kernel_global<BK><<<grid, BK::TPB, smem, st>>>(reinterpret_cast<const QpProg<BK>*>(p->packed_dev), p->aux, sp, a);
  }
  TZ_CUDA(cudaGetLastError());
  return TZ_OK;
}

static int launch(const TzProgram* p, const TzSolverOpts* o, const StepArgs& a, void* stream) {
  TZ_REQUIRE(p != nullptr, "null program");
  TZ_REQUIRE(a.S >= 0, "negative batch");
  if (a.S == 0) return TZ_OK;
  const SolverParams sp = to_params(o);
  TZ_REQUIRE(sp.max_iter >= 1 && sp.rho > 0 && sp.rho_act > 0 && sp.rho_inact > 0 && sp.alpha > 0 && sp.alpha < 2,
             "bad solver options");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (p->bucket) {
#ifndef TZ_DEV_ONLY_B2
    case 0: return launch_bucket<B0, true>(p, sp, a, st);
    case 1: return launch_bucket<B1, true>(p, sp, a, st);
#endif
    case 2: return launch_bucket<B2, true>(p, sp, a, st);
#ifndef TZ_DEV_ONLY_B2
    case 3: return launch_bucket<B3, true>(p, sp, a, st);
#endif
  }
  return fail(TZ_EINVAL, "corrupt program handle");
}

extern "C" int tz_solve(const TzProgram* prog, const TzSolverOpts* opts, int64_t S, const double* xbar0,
                        const double* e0, double* cost, double* v, double* xbar_traj, double* ze1, int32_t* status,
                        int32_t* iters, double* warm, void* stream) {
  TZ_REQUIRE(S == 0 || (xbar0 && e0 && status), "xbar0, e0 and status are required");
  StepArgs a{};
  a.S = S; a.ld = S; a.xbar0 = xbar0; a.e0 = e0; a.cost = cost; a.v = v; a.xbar_traj = xbar_traj; a.ze1 = ze1;
  a.status = status; a.iters = iters; a.warm = warm;
  return launch(prog, opts, a, stream);
}

extern "C" int tz_closed_loop_step(const TzProgram* prog, const TzSolverOpts* opts, int64_t S, double* x, double* xbar,
                                   double* e, const double* noise, const double* A_true, const double* B_true,
                                   double* cost, double* v, double* xbar_traj, double* ze1, double* u_out,
                                   int32_t* status, int32_t* iters, double* warm, double* stats, void* stream) {
  TZ_REQUIRE(S == 0 || (x && xbar && e && A_true && B_true && status), "x, xbar, e, A_true, B_true, status are required");
  StepArgs a{};
  a.S = S; a.ld = S; a.xbar0 = xbar; a.e0 = e; a.x = x; a.xbar = xbar; a.e = e; a.noise = noise; a.A_true = A_true; a.B_true = B_true;
  a.cost = cost; a.v = v; a.xbar_traj = xbar_traj; a.ze1 = ze1; a.u_out = u_out; a.status = status; a.iters = iters;
  a.warm = warm; a.stats = stats;
  return launch(prog, opts, a, stream);
}

extern "C" int tz_qp_solve(const TzProgram* prog, const TzSolverOpts* opts, int64_t S, const double* q, const double* l,
                           const double* u, double* z, double* y, int32_t* status, int32_t* iters, void* stream) {
  TZ_REQUIRE(S == 0 || (q && l && u && z && status), "q, l, u, z, status are required");
  StepArgs a{};
  a.S = S; a.ld = S; a.q_in = q; a.l_in = l; a.u_in = u; a.z_out = z; a.y_out = y; a.status = status; a.iters = iters;
  TzSolverOpts o;
  tz_solver_opts_default(&o);
  if (opts) o = *opts;
  o.warm_start = 0;
  return launch(prog, &o, a, stream);
}

// ---- host-buffer variant: chunked H2D -> kernel -> D2H pipeline ---------------------------------
static size_t host_scratch_doubles(const TzProgram* p, int64_t S) {
  const size_t n = p->n, nent = (size_t)p->n * (1 + p->g1), nt = (size_t)(p->N + 1) * p->n;
  // x, xbar, e, noise | cost | v | xbar_traj | ze1 | status (as int32, rounded up) | A, B
  return (size_t)S * (4 * n + 1 + p->nv + nt + nent + 1) + (size_t)(n * n + n * p->m) + 16;
}

extern "C" size_t tz_closed_loop_step_host_scratch_bytes(const TzProgram* prog, int64_t S) {
  if (!prog || S < 0) return 0;
  return host_scratch_doubles(prog, S) * sizeof(double);
}

extern "C" int tz_closed_loop_step_host(const TzProgram* prog, const TzSolverOpts* opts, int64_t S, double* x_host,
                                        double* xbar_host, double* e_host, const double* noise_host,
                                        const double* A_true_host, const double* B_true_host, double* cost_host,
                                        double* v_host, double* xbar_traj_host, double* ze1_host, int32_t* status_host,
                                        void* dev_scratch, int32_t nchunks) {
  TZ_REQUIRE(prog && dev_scratch, "null argument");
  TZ_REQUIRE(S == 0 || (x_host && xbar_host && e_host && noise_host && A_true_host && B_true_host && status_host),
             "x, xbar, e, noise, A_true, B_true, status are required");
  if (S == 0) return TZ_OK;
  const TzProgram* p = prog;
  const int64_t n = p->n, m = p->m, nent = (int64_t)p->n * (1 + p->g1), nt = (int64_t)(p->N + 1) * p->n, nv = p->nv;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > 16) nchunks = 16;
  if ((int64_t)nchunks > S) nchunks = (int)S;
  double* d = reinterpret_cast<double*>(dev_scratch);
  double* dA = d; d += n * n;
  double* dB = d; d += n * m;
  d += (16 - ((n * n + n * m) % 16)) % 16;
  double *dx = d, *dxb = dx + n * S, *de = dxb + n * S, *dw = de + n * S, *dcost = dw + n * S, *dv = dcost + S,
         *dtraj = dv + nv * S, *dze = dtraj + nt * S;
  int32_t* dst = reinterpret_cast<int32_t*>(dze + nent * S);
  cudaStream_t streams[16];
  for (int c = 0; c < nchunks; ++c) TZ_CUDA(cudaStreamCreateWithFlags(&streams[c], cudaStreamNonBlocking));
  int rc = TZ_OK;
  cudaError_t err = cudaMemcpyAsync(dA, A_true_host, n * n * sizeof(double), cudaMemcpyHostToDevice, streams[0]);
  if (err == cudaSuccess) err = cudaMemcpyAsync(dB, B_true_host, n * m * sizeof(double), cudaMemcpyHostToDevice, streams[0]);
  if (err == cudaSuccess) err = cudaStreamSynchronize(streams[0]);
  const int64_t per = (S + nchunks - 1) / nchunks;
  // The device arrays are SoA with leading dimension S; a chunk [s0, s1) of a d x S array is d strided
  // segments, moved with one 2-D copy per array.
  auto h2d = [&](double* dev, const double* host, int64_t rows, int64_t s0, int64_t cnt, cudaStream_t st) {
    return cudaMemcpy2DAsync(dev + s0, S * sizeof(double), host + s0, S * sizeof(double), cnt * sizeof(double), rows,
                             cudaMemcpyHostToDevice, st);
  };
  auto d2h = [&](double* host, const double* dev, int64_t rows, int64_t s0, int64_t cnt, cudaStream_t st) {
    return cudaMemcpy2DAsync(host + s0, S * sizeof(double), dev + s0, S * sizeof(double), cnt * sizeof(double), rows,
                             cudaMemcpyDeviceToHost, st);
  };
  for (int c = 0; c < nchunks && err == cudaSuccess && rc == TZ_OK; ++c) {
    const int64_t s0 = (int64_t)c * per, cnt = (s0 + per <= S ? per : S - s0);
    if (cnt <= 0) break;
    cudaStream_t st = streams[c];
    err = h2d(dx, x_host, n, s0, cnt, st);
    if (err == cudaSuccess) err = h2d(dxb, xbar_host, n, s0, cnt, st);
    if (err == cudaSuccess) err = h2d(de, e_host, n, s0, cnt, st);
    if (err == cudaSuccess) err = h2d(dw, noise_host, n, s0, cnt, st);
    if (err != cudaSuccess) break;
    // the chunk is its own batch of `cnt` scenarios inside arrays of leading dimension S
    StepArgs a{};
    a.S = cnt; a.ld = S;
    a.xbar0 = dxb + s0; a.e0 = de + s0; a.x = dx + s0; a.xbar = dxb + s0; a.e = de + s0; a.noise = dw + s0;
    a.A_true = dA; a.B_true = dB;
    a.cost = cost_host ? dcost + s0 : nullptr;
    a.v = v_host ? dv + s0 : nullptr;
    a.xbar_traj = xbar_traj_host ? dtraj + s0 : nullptr;
    a.ze1 = ze1_host ? dze + s0 : nullptr;
    a.status = dst + s0;
    rc = launch(p, opts, a, st);
    if (rc != TZ_OK) break;
    err = d2h(x_host, dx, n, s0, cnt, st);
    if (err == cudaSuccess) err = d2h(xbar_host, dxb, n, s0, cnt, st);
    if (err == cudaSuccess) err = d2h(e_host, de, n, s0, cnt, st);
    if (err == cudaSuccess && cost_host) err = d2h(cost_host, dcost, 1, s0, cnt, st);
    if (err == cudaSuccess && v_host) err = d2h(v_host, dv, nv, s0, cnt, st);
    if (err == cudaSuccess && xbar_traj_host) err = d2h(xbar_traj_host, dtraj, nt, s0, cnt, st);
    if (err == cudaSuccess && ze1_host) err = d2h(ze1_host, dze, nent, s0, cnt, st);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(status_host + s0, dst + s0, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  }
  for (int c = 0; c < nchunks; ++c) {
    cudaError_t e2 = cudaStreamSynchronize(streams[c]);
    if (err == cudaSuccess) err = e2;
    cudaStreamDestroy(streams[c]);
  }
  if (rc != TZ_OK) return rc;
  if (err != cudaSuccess) return fail(TZ_ECUDA, "closed_loop_step_host: %s", cudaGetErrorString(err));
  return TZ_OK;
}
