// Fused per-step kernel of the TZDDPC hot path: bounds of the parametric program from
// (xbar0, e0)  ->  ADMM + polish  ->  nominal trajectory, cost, Ze[1].Z  ->  closed-loop update.
//
// Persistent kernel: one wave of CTAs, each staging the (scaled, padded) program into shared
// memory once and then looping over tiles of SPB = 128/G scenarios.  Inside a tile
//   * solve phase : G adjacent lanes per scenario (tz_admm.cuh);
//   * output phase: threads are re-mapped to (scenario, slice) so that every warp writes full
//     128-byte lines of the scenario-fastest (SoA) output arrays; Ze[1].Z -- 88 % structural
//     zeros, but dense by contract -- is written from a zero-entry list and a one-term-per-entry
//     table with streaming stores.
// Replaces, per closed-loop step,
//   tzddpc/tzddpc.py:357-377  (TZDDPC.solve: parameter update + cvxpy solve + Ze[1])
//   examples/2.pulley_sim.py:90-96 (nominal/plant/error update, Zek.Z.value)
// Algorithmic HBM bytes per scenario-step:
//   8*[6n + N*m + (N+1)n + 1 + n(1+g1)] + 4.
#include <cmath>
#include <new>
#include <vector>

#include "tz_admm.cuh"

namespace tz {

constexpr int kMaxOm = 1 + 16 + 2 * kMaxN + 2 * kMaxN;   // [1; v; p; centre of Ze[1]; staged x+]

struct Aux {                       // device-resident tables with run-time sizes
  const double* XB;                // (N+1)n x nw : xbar_0..xbar_N = XB [1; v; p]
  const double* CZ;                // n x nw      : centre of Ze[1] = CZ [1; v; p]
  const double* K;                 // m x n
  const double* nz_coef;           // Ze[1] entries with one term: value = coef * om[idx]
  const int32_t* nz_ent;           //   their entry index r*(1+g1)+j
  const int32_t* nz_idx;           //   index into om = [1; v; p; centre]
  const int32_t* zero_ent;         // entries that are structurally zero
  int n_nz, n_zero;
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

// Shared-memory image of one CTA: the program (read-only after staging) and, per warp, the
// exchange buffers between the solve phase and the output phase of a tile.
template <class BK>
struct WarpBuf {
  double om[kMaxOm][BK::SPW];      // [1; v; p; centre of Ze[1]] per scenario of the warp's tile
  double ysave[BK::NCL][32];       // duals at the previous residual check (certificate of infeasibility)
  double cost[BK::SPW];
  int status[BK::SPW];
  int iters[BK::SPW];
};
template <class BK>
struct Smem {
  QpProg<BK> pg;
  WarpBuf<BK> wb[BK::WPB];
};

// Persistent kernel; every WARP loops on its own over tiles of SPW = 32/G scenarios, so there is
// no CTA barrier after the program has been staged (a CTA barrier made fast warps wait for the
// slowest ADMM solve of the CTA: 8 % of the samples in profiles/r1_v2_*).
template <class BK>
__global__ void __launch_bounds__(BK::TPB, BK::MINB) step_kernel(const QpProg<BK>* __restrict__ gpg, const Aux ax,
                                                                const SolverParams sp, const StepArgs a) {
  constexpr int NZ = BK::NZ, NCL = BK::NCL, N2 = BK::N2, NU = BK::NU, NPAR = BK::NPAR, NAG = BK::NAG, NCHL = BK::NCHL,
                NCOL = BK::NCOL, TPB = BK::TPB, G = BK::G, SPW = BK::SPW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<BK>& sm = *reinterpret_cast<Smem<BK>*>(smem_raw);
  const int tid = threadIdx.x;
  {  // stage the program once per CTA (persistent kernel: amortised over all tiles of this CTA)
    const double* src = reinterpret_cast<const double*>(gpg);
    double* dst = reinterpret_cast<double*>(&sm.pg);
    for (int i = tid; i < (int)(sizeof(QpProg<BK>) / sizeof(double)); i += TPB) dst[i] = src[i];
  }
  __syncthreads();
  const QpProg<BK>& pg = sm.pg;
  const int lane = tid & 31, wib = tid >> 5;
  WarpBuf<BK>& wb = sm.wb[wib];
  const int g = lane % G;                 // lane within the scenario's group
  const int sl = lane / G;                // scenario within the warp's tile (solve-phase mapping)
  const int64_t LD = a.ld;
  const int n = ax.n, m = ax.m, nv = ax.nv, nw = ax.nw;
  const bool explicit_qp = a.q_in != nullptr;
  const int64_t ntiles = (a.S + SPW - 1) / SPW;
  const int64_t nwarps = (int64_t)gridDim.x * BK::WPB;
  const double alpha = sp.alpha, inv_alpha = 1.0 / sp.alpha;

  for (int64_t tile = (int64_t)blockIdx.x * BK::WPB + wib; tile < ntiles; tile += nwarps) {
    const int64_t s = tile * SPW + sl;
    const bool live = s < a.S;
    double c0 = 0.0;
    bool param_ok = true, finite = true;
    LaneQp<BK> qp;

    if (!explicit_qp) {
      // ---- parameters p = [xbar0; e0] (every lane of the group loads them: same sectors)
      double w[NCOL];                        // w = [1; p; |p|; general atoms |Bt p + gam|]
      w[0] = 1.0;
#pragma unroll
      for (int j = 0; j < NPAR; ++j) {
        double val = 0.0;
        if (live && j < 2 * n) val = (j < n) ? a.xbar0[(int64_t)j * LD + s] : a.e0[(int64_t)(j - n) * LD + s];
        w[1 + j] = val;
        w[1 + NPAR + j] = fabs(val);
        finite = finite && (fabs(val) < 1e300);
        if (g == 0 && j < 2 * n) wb.om[1 + nv + j][sl] = val;      // kept for the output phase
      }
#pragma unroll
      for (int i = 0; i < NAG; ++i) {
        double acc = pg.gam[i];
#pragma unroll
        for (int j = 0; j < NPAR; ++j) acc = fma(pg.Bt[i][j], w[1 + j], acc);
        w[1 + 2 * NPAR + i] = fabs(acc);
      }
      // ---- this lane's rows of the bounds: l = l0 + R w, u = u0 + R w (scaled), kinks
#pragma unroll
      for (int k = 0; k < NCL; ++k) {
        const int i = k * G + g;
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) r = fma(pg.R[i][j], w[j], r);
        if (k < N2) {
          qp.lo[k < N2 ? k : 0] = pg.l0[i] + r;
          qp.hi[k < N2 ? k : 0] = pg.u0[i] + r;
          qp.kink[k < N2 ? k : 0] = pg.kink0[i % BK::NK] + r;
        } else if (k < N2 + NU) {
          qp.hi[k < N2 + NU ? k : 0] = pg.u0[i] + r;
        } else {
          qp.lo[k - NU] = pg.l0[i] + r;
        }
      }
#pragma unroll
      for (int j = 0; j < NZ; ++j) {
        double acc = pg.q0[j];
#pragma unroll
        for (int k = 0; k < NPAR; ++k) acc = fma(pg.Qp[j][k], w[1 + k], acc);
        qp.q[j] = acc;
      }
      // ---- parameter-only feasibility rows, split over the group
      int bad = 0;
#pragma unroll
      for (int k = 0; k < NCHL; ++k) {
        const int i = k * G + g;
        if (i < BK::NCHK) {
          double r = 0.0, sc = 1.0;
#pragma unroll
          for (int j = 0; j < NCOL; ++j) {
            const double t = pg.Rchk[i % BK::NCHK][j] * w[j];
            r += t;
            sc = fmax(sc, fabs(t));
          }
          bad |= (r > 1e-9 * sc) ? 1 : 0;
        }
      }
      param_ok = gor<G>(bad) == 0;
      // ---- cost constant c0(p)
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
    } else {
      // explicit instance: scale the caller's q, l, u  (qbar = c D q, lbar = E l)
#pragma unroll
      for (int j = 0; j < NZ; ++j)
        qp.q[j] = (live && j < pg.nz) ? a.q_in[(int64_t)j * LD + s] * pg.D[j] / pg.cinv : 0.0;
#pragma unroll
      for (int k = 0; k < NCL; ++k) {
        const int i = k * G + g;
        const int row = pg.row_of_slot[i];
        const bool rr = live && row >= 0;
        const double lv = rr ? a.l_in[(int64_t)row * LD + s] / pg.Einv[i] : -INFINITY;
        const double uv = rr ? a.u_in[(int64_t)row * LD + s] / pg.Einv[i] : INFINITY;
        if (k < N2) {
          qp.lo[k < N2 ? k : 0] = lv;
          qp.hi[k < N2 ? k : 0] = uv;
          qp.kink[k < N2 ? k : 0] = 0.0;
        } else if (k < N2 + NU) {
          qp.hi[k < N2 + NU ? k : 0] = uv;
        } else {
          qp.lo[k - NU] = lv;
        }
      }
    }
    // ---- this lane's rows of alpha*A and P (from shared memory; not kept live across the tile's other phases)
#pragma unroll
    for (int k = 0; k < NCL; ++k)
#pragma unroll
      for (int j = 0; j < NZ; ++j) qp.Aa[k][j] = alpha * pg.A[k * G + g][j];
    qp.P = pg.P;
    qp.wk = &pg.wabs[g];

    // ---- ADMM (+ polish)
    LaneState<BK> st;
    bool warm = false;
    if (sp.warm && a.warm != nullptr && live) {
      // layout: [x (NZ) | y (NC, slot-indexed) | activity words (G) | valid flag] x LD
      warm = (a.warm[(int64_t)(NZ + BK::NC + G) * LD + s] == 1.0);
      if (warm) {
#pragma unroll
        for (int j = 0; j < NZ; ++j) st.x[j] = a.warm[(int64_t)j * LD + s];
#pragma unroll
        for (int k = 0; k < NCL; ++k) st.w[k] = a.warm[(int64_t)(NZ + k * G + g) * LD + s];
        st.act = (uint32_t)__double_as_longlong(a.warm[(int64_t)(NZ + BK::NC + g) * LD + s]);
        st.switched = true;
      }
    }
    const bool solve_it = live && param_ok && finite;
    int iters = 0;
    int status = admm_solve<BK>(qp, sp, solve_it, st, warm, &wb.ysave[0][lane], iters);
    if (live && !finite) status = TZ_STATUS_NONFINITE;
    else if (live && !param_ok) status = TZ_STATUS_INFEASIBLE;
    const bool good = live && (status == TZ_STATUS_OK || status == TZ_STATUS_MAXITER);
    if (good && a.warm != nullptr) {
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < NZ; ++j) a.warm[(int64_t)j * LD + s] = st.x[j];
        a.warm[(int64_t)(NZ + BK::NC + G) * LD + s] = 1.0;
      }
#pragma unroll
      for (int k = 0; k < NCL; ++k) a.warm[(int64_t)(NZ + k * G + g) * LD + s] = st.w[k];
      a.warm[(int64_t)(NZ + BK::NC + g) * LD + s] = __longlong_as_double((long long)st.act);
    }
    if (sp.polish) (void)admm_polish<BK>(qp, inv_alpha, st, good, sp.polish);

    if (explicit_qp) {
      if (live) {
        if (g == 0) {
          a.status[s] = status;
          if (a.iters) a.iters[s] = iters;
          if (a.z_out) {
#pragma unroll
            for (int j = 0; j < NZ; ++j)
              if (j < pg.nz) a.z_out[(int64_t)j * LD + s] = good ? pg.D[j] * st.x[j] : NAN;
          }
        }
        if (a.y_out) {
#pragma unroll
          for (int k = 0; k < NCL; ++k) {
            const int i = k * G + g;
            const int row = pg.row_of_slot[i];
            if (row >= 0) a.y_out[(int64_t)row * LD + s] = good ? st.w[k] * pg.cinv / pg.Einv[i] : NAN;
          }
        }
      }
      continue;
    }

    // ---- objective value (reference `result`, tzddpc/tzddpc.py:367,377; constant terms included, quirk Q7)
    double kcost = 0.0;
#pragma unroll
    for (int k = 0; k < N2; ++k) {
      double axv = 0.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j) axv = fma(qp.Aa[k][j], st.x[j], axv);
      kcost = fma(qp.wk[k * G], fabs(axv * inv_alpha - qp.kink[k]), kcost);
    }
    kcost = gsum<G>(kcost);
    double cost = NAN;
    if (good) {
      double acc = kcost;
#pragma unroll
      for (int j = 0; j < NZ; ++j) {
        double px = 0.0;
#pragma unroll
        for (int b = 0; b < NZ; ++b) px = fma(qp.P[j][b], st.x[b], px);
        acc = fma(0.5 * px + qp.q[j], st.x[j], acc);
      }
      cost = fma(acc, pg.cinv, c0);
    } else if (live && status == TZ_STATUS_INFEASIBLE) {
      cost = INFINITY;                      // cvxpy returns +inf for an infeasible Minimize (:374)
    }
    // ---- hand om = [1; v; p; centre of Ze[1]], cost, status to the output phase (warp-private buffer)
    if (g == 0) {
      wb.om[0][sl] = 1.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j)
        if (j < nv) wb.om[1 + j][sl] = good ? pg.D[j] * st.x[j] : NAN;
      wb.cost[sl] = cost;
      wb.status[sl] = live ? status : -1;
      wb.iters[sl] = iters;
    }
    __syncwarp();
    for (int r = g; r < n; r += G) {        // centre of Ze[1]: rows split over the group
      const double* row = ax.CZ + (size_t)r * nw;
      double acc = 0.0;
      for (int j = 0; j < nw; ++j) acc = fma(__ldg(row + j), wb.om[j][sl], acc);
      wb.om[nw + r][sl] = acc;
    }
    __syncwarp();

    // =========== output phase: lane -> (scenario sc of the tile, slice); one store instruction covers
    // G entries x SPW scenarios = G segments of SPW*8 contiguous bytes ===============================
    {
      const int sc = lane % SPW, slice = lane / SPW;
      const int64_t so = tile * SPW + sc;
      const int stt = wb.status[sc];
      const bool olive = stt >= 0;
      if (olive) {
        // Ze[1].Z at the optimum (examples/2.pulley_sim.py:96: Zek.Z.value), dense n x (1+g1): zero-fill,
        // then overwrite the few non-zero entries (both writes merge in L2 before reaching HBM)
        if (a.ze1) {
          double* base = a.ze1 + so;
          const int nent = n * (1 + ax.g1);
#pragma unroll 4
          for (int en = slice; en < nent; en += G) __stcs(base + (int64_t)en * LD, 0.0);
          __syncwarp();
          for (int i = slice; i < ax.n_nz; i += G)
            __stcs(base + (int64_t)__ldg(ax.nz_ent + i) * LD, __ldg(ax.nz_coef + i) * wb.om[__ldg(ax.nz_idx + i)][sc]);
        }
        // nominal trajectory xbar_0..xbar_N = XB om  (tzddpc/tzddpc.py:166-170)
        if (a.xbar_traj) {
          const int nrows = (ax.N + 1) * n;
          for (int i = slice; i < nrows; i += G) {
            const double* row = ax.XB + (size_t)i * nw;
            double acc = 0.0;
            for (int j = 0; j < nw; ++j) acc = fma(__ldg(row + j), wb.om[j][sc], acc);
            a.xbar_traj[(int64_t)i * LD + so] = acc;
          }
        }
        if (a.v)
          for (int j = slice; j < nv; j += G) a.v[(int64_t)j * LD + so] = wb.om[1 + j][sc];
        if (slice == 0) {
          if (a.status) a.status[so] = stt;
          if (a.iters) a.iters[so] = wb.iters[sc];
          if (a.cost) a.cost[so] = wb.cost[sc];
        }
      }
      // ---- closed-loop update (examples/2.pulley_sim.py:90-94): row i of the update by slice i mod G
      if (a.x != nullptr) {
        const bool ogood = olive && (stt == TZ_STATUS_OK || stt == TZ_STATUS_MAXITER);
        double nrm2 = 0.0;
        if (ogood) {
          double us[kMaxM];
#pragma unroll
          for (int j = 0; j < kMaxM; ++j) {
            double acc = 0.0;
            if (j < m) {
              acc = wb.om[1 + j][sc];                                   // v[0]
              for (int i = 0; i < n; ++i) acc = fma(__ldg(ax.K + j * n + i), wb.om[1 + nv + n + i][sc], acc);
              if (a.u_out && slice == 0) a.u_out[(int64_t)j * LD + so] = acc;
            }
            us[j] = acc;                                                // u = K e + v[0]
          }
          for (int i = slice; i < n; i += G) {
            double acc = a.noise ? a.noise[(int64_t)i * LD + so] : 0.0;
            for (int k = 0; k < n; ++k) acc = fma(__ldg(a.A_true + i * n + k), a.x[(int64_t)k * LD + so], acc);
#pragma unroll
            for (int k = 0; k < kMaxM; ++k)
              if (k < m) acc = fma(__ldg(a.B_true + i * m + k), us[k], acc);
            const double* row = ax.XB + (size_t)(n + i) * nw;           // xbar+ = xbar_traj[1]
            double xb1 = 0.0;
            for (int j = 0; j < nw; ++j) xb1 = fma(__ldg(row + j), wb.om[j][sc], xb1);
            wb.om[nw + n + i][sc] = acc;                                // x+ staged: x is read by the other slices
            a.xbar[(int64_t)i * LD + so] = xb1;
            a.e[(int64_t)i * LD + so] = acc - xb1;                      // e+ = x+ - xbar+
          }
        }
        __syncwarp();
        if (ogood) {
          for (int i = slice; i < n; i += G) a.x[(int64_t)i * LD + so] = wb.om[nw + n + i][sc];   // x+ = A x + B u + w
          if (slice == 0)
            for (int i = 0; i < n; ++i) nrm2 = fma(wb.om[nw + n + i][sc], wb.om[nw + n + i][sc], nrm2);
        }
        if (a.stats != nullptr) {
          const bool cnt = slice == 0;      // one lane per scenario contributes
          double stv[TZ_NSTATS];
          stv[0] = (cnt && ogood) ? sqrt(nrm2) : 0.0;
          stv[1] = (cnt && ogood) ? nrm2 : 0.0;
          stv[2] = (cnt && ogood) ? wb.cost[sc] : 0.0;
          stv[3] = (cnt && olive && stt == TZ_STATUS_INFEASIBLE) ? 1.0 : 0.0;
          stv[4] = (cnt && olive && stt == TZ_STATUS_MAXITER) ? 1.0 : 0.0;
          stv[5] = (cnt && olive) ? (double)wb.iters[sc] : 0.0;
          stv[6] = (cnt && olive && stt == TZ_STATUS_NONFINITE) ? 1.0 : 0.0;
          stv[7] = (cnt && olive) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < TZ_NSTATS; ++k) {
            const double v_ = warp_sum(stv[k]);
            if (lane == 0 && v_ != 0.0) atomicAdd(a.stats + k, v_);
          }
        }
      }
    }
    __syncwarp();     // wb is rewritten by the next tile
  }
}

// ---- compiled buckets: <NZ, N2, NU, NL, G, NPAR, NAG, NCHK, MINB> -------------------------------
#ifndef TZ_B0_MINB
#define TZ_B0_MINB 3
#endif
using B0 = Bucket<2, 1, 3, 3, 4, 10, 2, 12, TZ_B0_MINB>;      // N = 2, m = 1, n <= 5: the three shipped examples (28 row slots)
using B1 = Bucket<4, 2, 4, 4, 8, 16, 8, 16, 3>;      // generic small   (80 row slots)
using B2 = Bucket<8, 3, 5, 5, 8, 16, 24, 16, 2>;     // generic medium  (104 row slots, N = 3..4)

struct RowClasses { int n2 = 0, nu = 0, nl = 0; };

static RowClasses classify(const TzProgramDesc& d, std::vector<int>* cls) {
  RowClasses rc;
  for (int i = 0; i < d.nc; ++i) {
    const bool lf = std::isfinite(d.l0[i]), uf = std::isfinite(d.u0[i]);
    int c = 0;                                       // 0: two-sided / kink / free, 1: upper only, 2: lower only
    if (d.wabs[i] > 0.0 || (lf && uf) || (!lf && !uf)) c = 0;
    else if (uf) c = 1;
    else c = 2;
    if (cls) (*cls)[i] = c;
    (c == 0 ? rc.n2 : c == 1 ? rc.nu : rc.nl)++;
  }
  return rc;
}

// unit atoms |p_c| are implicit columns; everything else is a "general" atom
static int general_atoms(const TzProgramDesc& d, std::vector<int>* colmap) {
  const int npar = d.npar;
  int nag = 0;
  for (int i = 0; i < d.na; ++i) {
    int nnz = 0, where = -1;
    for (int k = 0; k < npar; ++k)
      if (d.Bt[i * npar + k] != 0.0) { ++nnz; where = k; }
    const bool unit = (nnz == 1 && d.gam[i] == 0.0 && std::fabs(std::fabs(d.Bt[i * npar + where]) - 1.0) < 1e-15);
    if (colmap) (*colmap)[i] = unit ? -(where + 1) : nag;      // negative: unit atom of parameter `where`
    if (!unit) ++nag;
  }
  return nag;
}

template <class BK>
bool fits(const TzProgramDesc& d) {
  const RowClasses rc = classify(d, nullptr);
  return d.nz <= BK::NZ && rc.n2 <= BK::N2 * BK::G && rc.nu <= BK::NU * BK::G && rc.nl <= BK::NL * BK::G &&
         d.npar <= BK::NPAR && general_atoms(d, nullptr) <= BK::NAG && d.nchk <= BK::NCHK;
}

template <class BK>
void pack(const TzProgramDesc& d, QpProg<BK>& g) {
  std::memset(&g, 0, sizeof(g));
  const int nz = d.nz, nc = d.nc, npar = d.npar, na = d.na, ncol = 1 + npar + na;
  std::vector<int> cls(nc), amap(na);
  classify(d, &cls);
  const int nag = general_atoms(d, &amap);
  // column j of the caller's [1 | p | alpha] layout -> column of the padded [1 | p | |p| | general] layout
  auto colmap = [&](int j) {
    if (j <= npar) return j;
    const int am = amap[j - 1 - npar];
    return am < 0 ? 1 + BK::NPAR + (-am - 1) : 1 + 2 * BK::NPAR + am;
  };
  for (int a = 0; a < BK::NZ; ++a) g.D[a] = 1.0;
  for (int i = 0; i < BK::NC; ++i) { g.l0[i] = -INFINITY; g.u0[i] = INFINITY; g.Einv[i] = 1.0; g.row_of_slot[i] = -1; }
  for (int a = 0; a < nz; ++a) {
    g.D[a] = d.D[a];
    g.q0[a] = d.c * d.D[a] * d.q0[a];
    for (int b = 0; b < nz; ++b) g.P[a][b] = d.c * d.D[a] * d.P[a * nz + b] * d.D[b];
    for (int k = 0; k < npar; ++k) g.Qp[a][k] = d.c * d.D[a] * d.Qp[a * npar + k];
  }
  int next[3] = {0, BK::N2 * BK::G, (BK::N2 + BK::NU) * BK::G};
  for (int i = 0; i < nc; ++i) {
    const int s = next[cls[i]]++;
    const double E = d.E[i];
    g.row_of_slot[s] = i;
    g.Einv[s] = 1.0 / E;
    for (int a = 0; a < nz; ++a) g.A[s][a] = E * d.A[i * nz + a] * d.D[a];
    g.l0[s] = E * d.l0[i];
    g.u0[s] = E * d.u0[i];
    for (int j = 0; j < ncol; ++j) g.R[s][colmap(j)] += E * d.R[i * ncol + j];
    if (cls[i] == 0) {
      g.kink0[s] = E * d.kink0[i];
      g.wabs[s] = d.wabs[i] > 0.0 ? d.c * d.wabs[i] / E : 0.0;
    }
  }
  for (int i = 0; i < na; ++i) {
    if (amap[i] < 0) continue;
    g.gam[amap[i]] = d.gam[i];
    for (int k = 0; k < npar; ++k) g.Bt[amap[i]][k] = d.Bt[i * npar + k];
  }
  for (int i = 0; i < d.nchk; ++i)
    for (int j = 0; j < ncol; ++j) g.Rchk[i][colmap(j)] += d.Rchk[i * ncol + j];
  for (int j = 0; j < ncol; ++j) g.cc[colmap(j)] += d.cc[j];
  for (int a = 0; a < npar; ++a)
    for (int b = 0; b < npar; ++b) g.CC2[a][b] = d.CC2[a * npar + b];
  g.cinv = 1.0 / d.c;
  g.nz = nz; g.nc = nc; g.npar = npar; g.nag = nag; g.nchk = d.nchk;
  g.has_cc2 = 0;
  for (int a = 0; a < npar * npar; ++a) g.has_cc2 |= (d.CC2[a] != 0.0) ? 1 : 0;
}

}  // namespace tz

using namespace tz;

struct TzProgram {
  int bucket = -1;
  void* packed_dev = nullptr;            // device image of QpProg<bucket>, staged into shared memory by every CTA
  void* aux_dev = nullptr;               // one allocation holding the run-time sized tables
  Aux aux{};
  int nz = 0, nc = 0, n = 0, m = 0, N = 0, nv = 0, g1 = 0, npar = 0;
  int NZ = 0, NC = 0, G = 0;
  int num_sms = 148;
};

template <class BK>
static int create_bucket(const TzProgramDesc& d, TzProgram* p, int id) {
  p->bucket = id;
  std::vector<unsigned char> packed(sizeof(QpProg<BK>));
  pack<BK>(d, *reinterpret_cast<QpProg<BK>*>(packed.data()));
  p->NZ = BK::NZ;
  p->NC = BK::NC;
  p->G = BK::G;
  int dev = 0;
  TZ_CUDA(cudaGetDevice(&dev));
  TZ_CUDA(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev));
  TZ_CUDA(cudaMalloc(&p->packed_dev, sizeof(QpProg<BK>)));
  TZ_CUDA(cudaMemcpy(p->packed_dev, packed.data(), sizeof(QpProg<BK>), cudaMemcpyHostToDevice));
  return TZ_OK;
}

extern "C" int tz_program_create(const TzProgramDesc* d, TzProgram** out) {
  TZ_REQUIRE(d && out, "null argument");
  TZ_REQUIRE(d->n >= 1 && d->n <= kMaxN && d->m >= 1 && d->m <= kMaxM, "dim_x must be 1..%d and dim_u 1..%d", kMaxN, kMaxM);
  TZ_REQUIRE(d->npar == 2 * d->n, "npar must be 2*dim_x");
  TZ_REQUIRE(d->nv == d->horizon * d->m && d->nv <= 16 && d->nz >= d->nv, "bad nv/nz");
  TzProgram* p = new (std::nothrow) TzProgram();
  if (!p) return fail(TZ_ENOMEM, "out of host memory");
  int rc = TZ_ERANGE;
#define TZ_TRY(BK, ID) \
  if (rc == TZ_ERANGE && fits<BK>(*d)) rc = create_bucket<BK>(*d, p, ID);
  TZ_TRY(B0, 0) TZ_TRY(B1, 1) TZ_TRY(B2, 2)
#undef TZ_TRY
  if (rc != TZ_OK) {
    delete p;
    if (rc == TZ_ERANGE) {
      const RowClasses c = classify(*d, nullptr);
      return fail(TZ_ERANGE, "program (nz=%d rows: %d two-sided, %d upper, %d lower; npar=%d general atoms=%d nchk=%d) "
                  "exceeds every compiled bucket", d->nz, c.n2, c.nu, c.nl, d->npar, general_atoms(*d, nullptr), d->nchk);
    }
    return rc;
  }
  p->nz = d->nz; p->nc = d->nc; p->n = d->n; p->m = d->m; p->N = d->horizon; p->nv = d->nv; p->g1 = d->g1; p->npar = d->npar;
  // ---- run-time sized tables: XB, centre map of Ze[1], K, and the Ze[1] entry lists
  const int n = d->n, nw = 1 + d->nv + d->npar, ld1 = 1 + d->g1, nent = n * ld1;
  std::vector<double> CZ((size_t)n * nw, 0.0), coef;
  std::vector<int32_t> ent, idx, zero;
  for (int e = 0; e < nent; ++e) {
    const int t0 = d->ze1_ptr[e], t1 = d->ze1_ptr[e + 1];
    for (int t = t0; t < t1; ++t)
      if (d->ze1_idx[t] < 0 || d->ze1_idx[t] >= nw) { delete p; return fail(TZ_EINVAL, "ze1_idx[%d] out of range", t); }
    const int r = e / ld1, j = e % ld1;
    if (j == 0) {                        // centre column: its (possibly many) terms become om[nw + r]
      for (int t = t0; t < t1; ++t) CZ[(size_t)r * nw + d->ze1_idx[t]] += d->ze1_val[t];
      ent.push_back(e); idx.push_back(nw + r); coef.push_back(1.0);
    } else if (t1 - t0 == 0) {
      zero.push_back(e);
    } else if (t1 - t0 == 1) {
      ent.push_back(e); idx.push_back(d->ze1_idx[t0]); coef.push_back(d->ze1_val[t0]);
    } else {
      delete p;
      return fail(TZ_EINVAL, "generator entry %d of Ze[1] has %d terms: only single-term generator entries are supported "
                  "(boxed M_K / M_Delta)", e, t1 - t0);
    }
  }
  const size_t nXB = (size_t)(d->horizon + 1) * n * nw, nCZ = CZ.size(), nK = (size_t)d->m * n, nco = coef.size();
  const size_t ndbl = nXB + nCZ + nK + nco, nint = ent.size() + idx.size() + zero.size();
  std::vector<unsigned char> host(ndbl * sizeof(double) + nint * sizeof(int32_t) + 16);
  double* hd = reinterpret_cast<double*>(host.data());
  std::memcpy(hd, d->XB, nXB * sizeof(double));
  std::memcpy(hd + nXB, CZ.data(), nCZ * sizeof(double));
  std::memcpy(hd + nXB + nCZ, d->K, nK * sizeof(double));
  if (nco) std::memcpy(hd + nXB + nCZ + nK, coef.data(), nco * sizeof(double));
  int32_t* hi = reinterpret_cast<int32_t*>(hd + ndbl);
  if (!ent.empty()) std::memcpy(hi, ent.data(), ent.size() * sizeof(int32_t));
  if (!idx.empty()) std::memcpy(hi + ent.size(), idx.data(), idx.size() * sizeof(int32_t));
  if (!zero.empty()) std::memcpy(hi + ent.size() + idx.size(), zero.data(), zero.size() * sizeof(int32_t));
  cudaError_t err = cudaMalloc(&p->aux_dev, host.size());
  if (err == cudaSuccess) err = cudaMemcpy(p->aux_dev, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    tz_program_destroy(p);
    return fail(TZ_ECUDA, "aux upload: %s", cudaGetErrorString(err));
  }
  double* dd = reinterpret_cast<double*>(p->aux_dev);
  int32_t* di = reinterpret_cast<int32_t*>(dd + ndbl);
  Aux& ax = p->aux;
  ax.XB = dd; ax.CZ = dd + nXB; ax.K = dd + nXB + nCZ; ax.nz_coef = dd + nXB + nCZ + nK;
  ax.nz_ent = di; ax.nz_idx = di + ent.size(); ax.zero_ent = di + ent.size() + idx.size();
  ax.n_nz = (int)ent.size(); ax.n_zero = (int)zero.size();
  ax.n = n; ax.m = d->m; ax.N = d->horizon; ax.nv = d->nv; ax.g1 = d->g1; ax.nw = nw;
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
  snprintf(buf, cap, "B%d(NZ=%d,NC=%d,G=%d)", p->bucket, p->NZ, p->NC, p->G);
  return p->bucket;
}

extern "C" int tz_program_warm_rows(const TzProgram* p) {
  if (!p) return fail(TZ_EINVAL, "null program");
  return p->NZ + p->NC + p->G + 1;        // x | y | activity words | valid flag
}

extern "C" void tz_solver_opts_default(TzSolverOpts* o) {
  if (!o) return;
  o->rho = 0.1; o->rho_active = 100.0; o->rho_inactive = 0.1; o->sigma = 1e-6; o->alpha = 1.6;
  o->eps_abs = 1e-6; o->eps_rel = 1e-6; o->max_iter = 4000; o->check_every = 8; o->polish = 3; o->warm_start = 0;
}

static SolverParams to_params(const TzSolverOpts* o) {
  TzSolverOpts d;
  tz_solver_opts_default(&d);
  if (o) d = *o;
  return SolverParams{d.rho, d.rho_active, d.rho_inactive, d.sigma, d.alpha, d.eps_abs, d.eps_rel,
                      d.max_iter, d.check_every, d.polish, d.warm_start};
}

template <class BK>
static int launch_bucket(const TzProgram* p, const SolverParams& sp, const StepArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(Smem<BK>);
  static bool configured = false;     // benign race: the attribute is idempotent
  if (!configured) {
    TZ_CUDA(cudaFuncSetAttribute(step_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // persistent grid: one wave of CTAs (MINB per SM); every warp loops over tiles of SPW scenarios
  const int64_t ntiles = (a.S + BK::SPW - 1) / BK::SPW;
  const int64_t need = (ntiles + BK::WPB - 1) / BK::WPB;
  const int64_t wave = (int64_t)p->num_sms * BK::MINB;
  const unsigned grid = (unsigned)(need < wave ? need : wave);
  step_kernel<BK><<<grid, BK::TPB, smem, st>>>(reinterpret_cast<const QpProg<BK>*>(p->packed_dev), p->aux, sp, a);
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
    case 0: return launch_bucket<B0>(p, sp, a, st);
#ifndef TZ_DEV_ONLY_B0
    case 1: return launch_bucket<B1>(p, sp, a, st);
    case 2: return launch_bucket<B2>(p, sp, a, st);
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
