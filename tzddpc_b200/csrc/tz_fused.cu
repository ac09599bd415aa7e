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

// Run-time sized tables of a program, one device blob staged into shared memory by every CTA:
//   doubles: XB ((N+1)n x NW) | CZ (n x NW) | K (m x n) | nz_coef (n_nz)      [+ A_true, B_true appended in smem]
//   int32  : nz_ent (n_nz) | nz_idx (n_nz)
// XB / CZ columns and nz_idx address the per-scenario vector om (layout in Bucket: OM_*).
struct Aux {
  const double* tab;
  int n_dbl, n_int;                // sizes of the two parts
  int o_XB, o_CZ, o_K, o_coef;     // offsets (doubles)
  int o_ent, o_idx;                // offsets (int32, from the start of the int part)
  int n_nz;                        // entries of Ze[1].Z that are not structurally zero (centre column included)
  int n, m, N, nv, g1;
};

struct StepArgs {
  int64_t S;                       // scenarios in this launch
  int64_t ld;                      // leading dimension of every SoA array (>= S)
  int vec2;                        // 1: S, ld even and every array 16-byte aligned -> two scenarios per lane in the output phase
  const double* xbar0;             // parameters (n x S); aliases xbar/e in closed loop
  const double* e0;
  double* x;                       // closed loop only (NULL = solve only)
  double* xbar;
  double* e;
  const double* noise;
  const double* x_restart;         // closed loop only: state an infeasible scenario restarts from (NULL: it keeps its state)
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
struct alignas(16) WarpBuf {
  double pre[2][BK::PRE_ROWS][BK::SPO];  // cp.async double buffer: rows [xbar0 | e0 | x | noise] (n each) of this / the next output tile
  double om[BK::KOM][BK::SPO];     // [1 | v | xbar0 | e0 | centre of Ze[1] | x+] per scenario of the warp's output tile
  double ysave[BK::NCL][32];       // duals at the previous residual check (certificate of infeasibility)
  double cost[BK::SPO];
  double stacc[TZ_NSTATS][BK::SPO];   // closed-loop statistics of this warp's tiles, reduced once at the end of the kernel
  int status[BK::SPO];
  int iters[BK::SPO];
};
template <class BK>
struct alignas(16) Smem {
  double Aa[BK::NC][BK::NZ];       // alpha * A (alpha is a solver option, so this is built when the program is staged)
  QpProg<BK> pg;
  WarpBuf<BK> wb[BK::WPB];
};

// ---- W scenarios per lane (1: scalar accesses, 2: 16-byte accesses) ----------------------------
template <int W> struct Vec;
template <> struct Vec<1> { double a; };
template <> struct Vec<2> { double a, b; };
__device__ __forceinline__ Vec<1> vld(const double* p, Vec<1>*) { return Vec<1>{*p}; }
__device__ __forceinline__ Vec<2> vld(const double* p, Vec<2>*) { const double2 t = *reinterpret_cast<const double2*>(p); return Vec<2>{t.x, t.y}; }
__device__ __forceinline__ void vst(double* p, Vec<1> v) { *p = v.a; }
__device__ __forceinline__ void vst(double* p, Vec<2> v) { *reinterpret_cast<double2*>(p) = make_double2(v.a, v.b); }
#ifndef TZ_ZST
#define TZ_ZST 1
#endif
#if TZ_ZST == 0
__device__ __forceinline__ void vstcs(double* p, Vec<1> v) { *p = v.a; }
__device__ __forceinline__ void vstcs(double* p, Vec<2> v) { *reinterpret_cast<double2*>(p) = make_double2(v.a, v.b); }
#elif TZ_ZST == 1
__device__ __forceinline__ void vstcs(double* p, Vec<1> v) { __stcs(p, v.a); }
__device__ __forceinline__ void vstcs(double* p, Vec<2> v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.a, v.b)); }
#else
__device__ __forceinline__ void vstcs(double* p, Vec<1> v) { __stcg(p, v.a); }
__device__ __forceinline__ void vstcs(double* p, Vec<2> v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.a, v.b)); }
#endif
__device__ __forceinline__ Vec<1> vfma(double c, Vec<1> x, Vec<1> acc) { return Vec<1>{fma(c, x.a, acc.a)}; }
__device__ __forceinline__ Vec<2> vfma(double c, Vec<2> x, Vec<2> acc) { return Vec<2>{fma(c, x.a, acc.a), fma(c, x.b, acc.b)}; }
__device__ __forceinline__ Vec<1> vmul(double c, Vec<1> x) { return Vec<1>{c * x.a}; }
__device__ __forceinline__ Vec<2> vmul(double c, Vec<2> x) { return Vec<2>{c * x.a, c * x.b}; }
__device__ __forceinline__ Vec<1> vsub(Vec<1> x, Vec<1> y) { return Vec<1>{x.a - y.a}; }
__device__ __forceinline__ Vec<2> vsub(Vec<2> x, Vec<2> y) { return Vec<2>{x.a - y.a, x.b - y.b}; }
__device__ __forceinline__ Vec<1> vzero(Vec<1>*) { return Vec<1>{0.0}; }
__device__ __forceinline__ Vec<2> vzero(Vec<2>*) { return Vec<2>{0.0, 0.0}; }
__device__ __forceinline__ Vec<1> vsel(const bool* g, Vec<1> x, double other) { return Vec<1>{g[0] ? x.a : other}; }
__device__ __forceinline__ Vec<2> vsel(const bool* g, Vec<2> x, double other) { return Vec<2>{g[0] ? x.a : other, g[1] ? x.b : other}; }
__device__ __forceinline__ double vget(Vec<1> x, int) { return x.a; }
__device__ __forceinline__ double vget(Vec<2> x, int i) { return i == 0 ? x.a : x.b; }

// ---- cp.async prefetch of the per-scenario inputs of an output tile (hides the DRAM latency of the only loads of the step)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <class BK>
__device__ __forceinline__ void prefetch_inputs(double (*dst)[BK::SPO], const StepArgs& a, const double* hint, int n,
                                                int64_t otile, int lane) {
  constexpr int SPO = BK::SPO;
  const double* src[4] = {a.xbar0, a.e0, a.x, a.noise};
  const int64_t s0 = otile * SPO;
  const int narr = a.x != nullptr ? 4 : 2;
  if (hint != nullptr) {                                 // active-set hint words: G rows behind the 4n input rows
    if (a.vec2) {
      for (int c = lane; c < BK::G * (SPO / 2); c += 32) {
        const int row = c / (SPO / 2), ch = c - row * (SPO / 2);
        const int64_t s = s0 + 2 * ch;
        if (s < a.S) cp_async16(&dst[4 * n + row][2 * ch], hint + (int64_t)row * a.ld + s);
      }
    } else {
      for (int c = lane; c < BK::G * SPO; c += 32) {
        const int row = c / SPO, ch = c - row * SPO;
        const int64_t s = s0 + ch;
        if (s < a.S) cp_async8(&dst[4 * n + row][ch], hint + (int64_t)row * a.ld + s);
      }
    }
  }
  if (a.vec2) {
    constexpr int CPR = SPO / 2;                        // 16-byte chunks per row
    const int total = narr * n * CPR;
    for (int c = lane; c < total; c += 32) {
      const int row = c / CPR, ch = c - row * CPR, arr = row / n, r = row - arr * n;
      const int64_t s = s0 + 2 * ch;
      if (src[arr] != nullptr && s < a.S) cp_async16(&dst[row][2 * ch], src[arr] + (int64_t)r * a.ld + s);
    }
  } else {
    const int total = narr * n * SPO;
    for (int c = lane; c < total; c += 32) {
      const int row = c / SPO, ch = c - row * SPO, arr = row / n, r = row - arr * n;
      const int64_t s = s0 + ch;
      if (src[arr] != nullptr && s < a.S) cp_async8(&dst[row][ch], src[arr] + (int64_t)r * a.ld + s);
    }
  }
  cp_async_commit();
}

// Zero-fill of the dense Ze[1].Z block of one output tile (88 % of its entries are structural zeros, but the reference
// returns the matrix dense, examples/2.pulley_sim.py:96).  It depends on nothing the solve computes, so half of the
// warps issue it BEFORE solving their tile and the other half after: the stores of one half drain while the other half
// computes (all warps in lock step would alternate between an idle DRAM and a saturated one: profiles/r1_v8).
template <class BK, int W>
__device__ __forceinline__ void zero_fill(const Aux& ax, const StepArgs& a, int64_t tile, int lane) {
  constexpr int SPW = BK::SPO, NGRP = SPW / W, NSL = 32 / NGRP;
  using V = Vec<W>;
  V* const vt = nullptr;
  const int pc = lane % NGRP, slice = lane / NGRP;
  const int64_t so = tile * SPW + pc * W;
  if (so >= a.S) return;
  const int64_t LD = a.ld;
  const int nent = ax.n * (1 + ax.g1);
  const int64_t stepb = (int64_t)NSL * LD;
  double* ptr = a.ze1 + so + (int64_t)slice * LD;
  const V z0 = vzero(vt);
  // (a down-counter: with the trip count as loop bound the compiler spilled it and re-loaded it from local memory in
  // every iteration, behind the stores in the same LSU queue -- 20 % of all stall samples in profiles/r1_v6)
#pragma unroll 4
  for (int cnt = (nent - slice + NSL - 1) / NSL; cnt > 0; --cnt, ptr += stepb) vstcs(ptr, z0);
}

// Output phase of one output tile (SPO >= 16 consecutive scenarios = TPO solve tiles): lane -> (W consecutive
// scenarios, slice); one store instruction of the warp covers NSL entries x SPO scenarios = NSL runs of SPO*8 >= 128
// contiguous bytes (full lines) of the scenario-fastest arrays.
template <class BK, int W>
__device__ __forceinline__ void output_phase(WarpBuf<BK>& wb, const double (*pre)[BK::SPO], const Aux& ax, const StepArgs& a,
                                             const double* __restrict__ tabd, const int* __restrict__ tabi, int64_t tile, int lane,
                                             bool zero_done) {
  constexpr int SPW = BK::SPO, NW = BK::NW, NGRP = SPW / W, NSL = 32 / NGRP;
  using V = Vec<W>;
  V* const vt = nullptr;
  const int pc = lane % NGRP, slice = lane / NGRP, sc0 = pc * W;
  const int64_t so = tile * SPW + sc0;
  const int64_t LD = a.ld;
  const int n = ax.n, m = ax.m, nv = ax.nv;
  const double* sXB = tabd + ax.o_XB;
  int stt[W];
  bool olive[W], ogood[W];
  bool any_live = false, all_good = true;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    stt[t] = wb.status[sc0 + t];
    olive[t] = stt[t] >= 0;
    ogood[t] = stt[t] == TZ_STATUS_OK || stt[t] == TZ_STATUS_MAXITER;
    any_live = any_live || olive[t];
    all_good = all_good && ogood[t];
  }
  // (vector path: S is even, so the scenarios of a pair are live together)
  auto om = [&](int j) { return vld(&wb.om[j][sc0], vt); };
  if (any_live) {
    // ---- Ze[1].Z at the optimum (examples/2.pulley_sim.py:96: Zek.Z.value), dense n x (1+g1): zero-fill, then
    // overwrite the ~12 % entries that are not structurally zero (both writes merge in L2 before reaching HBM)
    if (a.ze1) {
      double* base = a.ze1 + so;
      if (!zero_done) zero_fill<BK, W>(ax, a, tile, lane);
      __syncwarp();
      const double* coef = tabd + ax.o_coef;
      const int* ent = tabi + ax.o_ent;
      const int* idx = tabi + ax.o_idx;
#pragma unroll 4
      for (int i = slice; i < ax.n_nz; i += NSL) vstcs(base + (int64_t)ent[i] * LD, vmul(coef[i], om(idx[i])));
    }
    // ---- nominal trajectory xbar_0..xbar_N = XB om  (tzddpc/tzddpc.py:166-170)
    if (a.xbar_traj) {
      const int nrows = (ax.N + 1) * n;
      for (int i = slice; i < nrows; i += NSL) {
        const double* row = sXB + i * NW;
        V acc = vzero(vt);
#pragma unroll
        for (int j = 0; j < NW; ++j) acc = vfma(row[j], om(j), acc);
        vst(a.xbar_traj + (int64_t)i * LD + so, acc);
      }
    }
    if (a.v)
      for (int j = slice; j < nv; j += NSL) vst(a.v + (int64_t)j * LD + so, om(BK::OM_V + j));
    if (slice == 0) {
#pragma unroll
      for (int t = 0; t < W; ++t) {
        if (olive[t]) {
          if (a.status) a.status[so + t] = stt[t];
          if (a.iters) a.iters[so + t] = wb.iters[sc0 + t];
          if (a.cost) a.cost[so + t] = wb.cost[sc0 + t];
        }
      }
    }
  }
  // ---- closed-loop update (examples/2.pulley_sim.py:90-94): row i of the update by slice i
  if (a.x != nullptr) {
    const double* sK = tabd + ax.o_K;
    const double* sA = tabd + ax.n_dbl;
    const double* sB = sA + n * n;
    if (any_live) {
      V us[kMaxM];
#pragma unroll
      for (int j = 0; j < kMaxM; ++j) {
        V acc = vzero(vt);
        if (j < m) {
          acc = om(BK::OM_V + j);                                              // v[0]
          for (int i = 0; i < n; ++i) acc = vfma(sK[j * n + i], om(BK::OM_P + BK::NPAR / 2 + i), acc);
          if (a.u_out && slice == 0) vst(a.u_out + (int64_t)j * LD + so, vsel(ogood, acc, NAN));
        }
        us[j] = acc;                                                           // u = K e + v[0]
      }
      for (int i = slice; i < n; i += NSL) {
        V acc = a.noise ? vld(&pre[3 * n + i][sc0], vt) : vzero(vt);
        for (int k = 0; k < n; ++k) acc = vfma(sA[i * n + k], vld(&pre[2 * n + k][sc0], vt), acc);
#pragma unroll
        for (int k = 0; k < kMaxM; ++k)
          if (k < m) acc = vfma(sB[i * m + k], us[k], acc);
        const double* row = sXB + (n + i) * NW;                                // xbar+ = xbar_traj[1]
        V xb1 = vzero(vt);
#pragma unroll
        for (int j = 0; j < NW; ++j) xb1 = vfma(row[j], om(j), xb1);
        V en = vsub(acc, xb1);                                                 // e+ = x+ - xbar+
        if (!all_good) {
          // a scenario whose step failed keeps its state, or -- the reference raises and the run ends
          // (tzddpc/tzddpc.py:374-375) -- starts a new run from x_restart: x = xbar = x_restart, e = 0
          const V xo = vld(&pre[2 * n + i][sc0], vt);
          const V xbo = om(BK::OM_P + i), eo = om(BK::OM_P + BK::NPAR / 2 + i);
          const V xr = a.x_restart ? vld(a.x_restart + (int64_t)i * LD + so, vt) : xo;
          double xa[W], ba[W], ea[W];
#pragma unroll
          for (int t = 0; t < W; ++t) {
            xa[t] = ogood[t] ? vget(acc, t) : vget(xr, t);
            ba[t] = ogood[t] ? vget(xb1, t) : (a.x_restart ? vget(xr, t) : vget(xbo, t));
            ea[t] = ogood[t] ? vget(en, t) : (a.x_restart ? 0.0 : vget(eo, t));
          }
          if constexpr (W == 1) { acc = V{xa[0]}; xb1 = V{ba[0]}; en = V{ea[0]}; }
          else { acc = V{xa[0], xa[1]}; xb1 = V{ba[0], ba[1]}; en = V{ea[0], ea[1]}; }
        }
        vst(&wb.om[BK::OM_XP + i][sc0], acc);
        vst(a.x + (int64_t)i * LD + so, acc);                                  // x+ = A x + B u + w
        vst(a.xbar + (int64_t)i * LD + so, xb1);
        vst(a.e + (int64_t)i * LD + so, en);
      }
    }
    if (a.stats != nullptr) {       // per-scenario-slot partial sums in shared memory, reduced once at the end of the kernel
      __syncwarp();
      if (slice == 0) {
#pragma unroll
        for (int t = 0; t < W; ++t) {
          if (olive[t]) {
            const int c = sc0 + t;
            double nrm2 = 0.0;
            for (int i = 0; i < n; ++i) { const double xv = wb.om[BK::OM_XP + i][c]; nrm2 = fma(xv, xv, nrm2); }
            if (ogood[t]) { wb.stacc[0][c] += sqrt(nrm2); wb.stacc[1][c] += nrm2; wb.stacc[2][c] += wb.cost[c]; }
            if (stt[t] == TZ_STATUS_INFEASIBLE) wb.stacc[3][c] += 1.0;
            if (stt[t] == TZ_STATUS_MAXITER) wb.stacc[4][c] += 1.0;
            wb.stacc[5][c] += (double)wb.iters[c];
            if (stt[t] == TZ_STATUS_NONFINITE) wb.stacc[6][c] += 1.0;
            wb.stacc[7][c] += 1.0;
          }
        }
      }
    }
  }
}

// Persistent kernel; every WARP loops on its own over tiles of SPW = 32/G scenarios, so there is
// no CTA barrier after the program has been staged (a CTA barrier made fast warps wait for the
// slowest ADMM solve of the CTA: 8 % of the samples in profiles/r1_v2_*).
template <class BK>
__global__ void __launch_bounds__(BK::TPB, BK::MINB) step_kernel(const QpProg<BK>* __restrict__ gpg, const Aux ax,
                                                                const SolverParams sp, const StepArgs a) {
  constexpr int NZ = BK::NZ, NCL = BK::NCL, N2 = BK::N2, NU = BK::NU, NPAR = BK::NPAR, NAG = BK::NAG, NCHL = BK::NCHL,
                NCOL = BK::NCOL, TPB = BK::TPB, G = BK::G, SPW = BK::SPW, NW = BK::NW, HP = BK::NPAR / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<BK>& sm = *reinterpret_cast<Smem<BK>*>(smem_raw);
  double* tabd = reinterpret_cast<double*>(smem_raw + sizeof(Smem<BK>));
  const int tid = threadIdx.x;
  const int n = ax.n, m = ax.m, nv = ax.nv;
  int* tabi = reinterpret_cast<int*>(tabd + ax.n_dbl + n * n + n * m);
  {  // stage the program and its tables once per CTA (persistent kernel: amortised over all tiles of this CTA)
    const double* src = reinterpret_cast<const double*>(gpg);
    double* dst = reinterpret_cast<double*>(&sm.pg);
    for (int i = tid; i < (int)(sizeof(QpProg<BK>) / sizeof(double)); i += TPB) dst[i] = src[i];
    for (int i = tid; i < ax.n_dbl; i += TPB) tabd[i] = ax.tab[i];
    if (a.x != nullptr) {
      for (int i = tid; i < n * n; i += TPB) tabd[ax.n_dbl + i] = a.A_true[i];
      for (int i = tid; i < n * m; i += TPB) tabd[ax.n_dbl + n * n + i] = a.B_true[i];
    }
    const int* gi = reinterpret_cast<const int*>(ax.tab + ax.n_dbl);
    for (int i = tid; i < ax.n_int; i += TPB) tabi[i] = gi[i];
  }
  __syncthreads();
  for (int i = tid; i < BK::NC * NZ; i += TPB) (&sm.Aa[0][0])[i] = sp.alpha * (&sm.pg.A[0][0])[i];
  __syncthreads();
  const QpProg<BK>& pg = sm.pg;
  const int lane = tid & 31, wib = tid >> 5;
  WarpBuf<BK>& wb = sm.wb[wib];
  const int g = lane % G;                 // lane within the scenario's group
  const int sl = lane / G;                // scenario within the warp's tile (solve-phase mapping)
  const int64_t LD = a.ld;
  const bool explicit_qp = a.q_in != nullptr;
  const int64_t ntiles = (a.S + BK::SPO - 1) / BK::SPO;
  const int64_t nwarps = (int64_t)gridDim.x * BK::WPB;
  const double inv_alpha = 1.0 / sp.alpha;
  const double* sCZ = tabd + ax.o_CZ;
  for (int i = lane; i < TZ_NSTATS * BK::SPO; i += 32) (&wb.stacc[0][0])[i] = 0.0;
  __syncwarp();

  int buf = 0;
  const int64_t otile0 = (int64_t)blockIdx.x * BK::WPB + wib;
  const double* hintp = (sp.warm == 2 && !explicit_qp) ? a.warm : nullptr;
  if (!explicit_qp && otile0 < ntiles) prefetch_inputs<BK>(wb.pre[0], a, hintp, n, otile0, lane);
  for (int64_t otile = otile0; otile < ntiles; otile += nwarps, buf ^= 1) {
    if (!explicit_qp) {        // inputs of the NEXT output tile stream in while this one is solved
      if (otile + nwarps < ntiles) {
        prefetch_inputs<BK>(wb.pre[buf ^ 1], a, hintp, n, otile + nwarps, lane);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
    }
    const double (*pre)[BK::SPO] = wb.pre[buf];
    const bool zero_first = !explicit_qp && a.ze1 != nullptr && ((blockIdx.x * BK::WPB + wib) & 1);
    if (zero_first) {
      if (a.vec2) zero_fill<BK, 2>(ax, a, otile, lane);
      else zero_fill<BK, 1>(ax, a, otile, lane);
    }
   #pragma unroll 1
   for (int half = 0; half < BK::TPO; ++half) {
    const int col = half * SPW + sl;        // column of this scenario in the warp's exchange buffers
    const int64_t s = otile * BK::SPO + col;
    const bool live = s < a.S;
    double c0 = 0.0;
    bool param_ok = true, finite = true;
    LaneQp<BK> qp;

    if (!explicit_qp) {
      // ---- parameters p = [xbar0 | e0] (every lane of the group loads them: same sectors)
      double w[BK::NCOLP];                   // w = [1 | p | |p| | general atoms |Bt p + gam|]
      w[0] = 1.0;
      w[BK::NCOLP - 1] = 0.0;
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        double xv = 0.0, ev = 0.0;
        if (live && j < n) { xv = pre[j][col]; ev = pre[n + j][col]; }
        w[1 + j] = xv;
        w[1 + HP + j] = ev;
        w[1 + NPAR + j] = fabs(xv);
        w[1 + NPAR + HP + j] = fabs(ev);
        finite = finite && (fabs(xv) < 1e300) && (fabs(ev) < 1e300);
        if (g == 0) { wb.om[BK::OM_P + j][col] = xv; wb.om[BK::OM_P + HP + j][col] = ev; }   // kept for the output phase
      }
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
      // ---- this lane's rows of the bounds: l = l0 + R w, u = u0 + R w (scaled), kinks
#pragma unroll
      for (int k = 0; k < NCL; ++k) {
        const int i = k * G + g;
        // (16-byte shared loads, two accumulation chains per row)
        const double2* Rr = reinterpret_cast<const double2*>(&pg.R[i][0]);
        double r = 0.0, r1 = 0.0;
#pragma unroll
        for (int j = 0; j < BK::NCOLP / 2; ++j) {
          const double2 c2 = Rr[j];
          r = fma(c2.x, w[2 * j], r);
          r1 = fma(c2.y, w[2 * j + 1], r1);
        }
        r += r1;
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
        if (pg.has_qp) {
#pragma unroll
          for (int k = 0; k < NPAR; ++k) acc = fma(pg.Qp[j][k], w[1 + k], acc);
        }
        qp.q[j] = acc;
      }
      // ---- parameter-only feasibility rows, split over the group:  sum_j R_j w_j <= 1e-9 max(1, sum_j |R_j| |w_j|)
      int bad = 0;
#pragma unroll
      for (int k = 0; k < NCHL; ++k) {
        const int i = k * G + g;
        if (i < pg.nchk) {
          const double2* Rc = reinterpret_cast<const double2*>(&pg.Rchk[i % BK::NCHK][0]);
          double r = 0.0, ra = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < BK::NCOLP / 2; ++j2) {
            const double2 c2 = Rc[j2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = 2 * j2 + h;
              const double c = h == 0 ? c2.x : c2.y;
              r = fma(c, w[j], r);
              ra = fma(fabs(c), (j >= 1 && j <= NPAR) ? w[j + NPAR] : w[j], ra);    // |w_j|: the |p| columns are already there
            }
          }
          bad |= (r > 1e-9 * fmax(1.0, ra)) ? 1 : 0;
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
    qp.Aa.base = &sm.Aa[g][0];
    qp.P = pg.P;
    qp.wk = &pg.wabs[g];
    qp.sinv = &pg.sing_inv[g];
    qp.svar = &pg.sing_var[g];

    // ---- ADMM (+ certificate / polish)
    LaneState<BK> st;
    bool warm = false;
    if (sp.warm == 1 && a.warm != nullptr && live) {
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
    bool certified = false;
    int status = TZ_STATUS_OK;
    // ---- active-set hint (warm_start == 2): the optimal active set of the scenario's previous closed-loop step is
    // tried first; when its KKT certificate holds the step is solved exactly without a single ADMM iteration
    const bool use_hint = sp.warm == 2 && a.warm != nullptr;
    bool hint_ok = false;
    if (use_hint) {
      unsigned long long hint = 0ull;
      if (live) hint = (unsigned long long)__double_as_longlong(pre[4 * n + g][col]);
      const bool valid = solve_it && gor<G>((hint & kCodeValid) ? 0 : 1) == 0;
      if (__any_sync(0xffffffffu, valid)) {
        double lam[NCL], xk[NZ], x0[NZ];
#pragma unroll
        for (int k = 0; k < NCL; ++k) lam[k] = 0.0;
#pragma unroll
        for (int j = 0; j < NZ; ++j) x0[j] = 0.0;
        const unsigned long long code = hint & ~kCodeValid;
        const bool ok = admm_certify<BK>(qp, inv_alpha, sp.polish > 0 ? sp.polish : 3, code, x0, lam, xk);
        if (ok && valid) {
#pragma unroll
          for (int j = 0; j < NZ; ++j) st.x[j] = xk[j];
#pragma unroll
          for (int k = 0; k < NCL; ++k) st.w[k] = lam[k];
          st.code = code;
          hint_ok = true;
        }
      }
    }
    if (!__all_sync(0xffffffffu, hint_ok || !solve_it)) {
      bool cert2 = false;
      const int st2 = admm_solve<BK>(qp, sp, solve_it && !hint_ok, st, warm, &wb.ysave[0][lane], iters, cert2);
      if (!hint_ok) { status = st2; certified = cert2; }
    }
    if (hint_ok) { certified = true; iters = 0; }
    if (live && !finite) status = TZ_STATUS_NONFINITE;
    else if (live && !param_ok) status = TZ_STATUS_INFEASIBLE;
    const bool good = live && (status == TZ_STATUS_OK || status == TZ_STATUS_MAXITER);
    if (use_hint) {
      if (live) a.warm[(int64_t)g * LD + s] = good ? __longlong_as_double((long long)(st.code | kCodeValid)) : 0.0;
    } else if (good && a.warm != nullptr) {
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < NZ; ++j) a.warm[(int64_t)j * LD + s] = st.x[j];
        a.warm[(int64_t)(NZ + BK::NC + G) * LD + s] = 1.0;
      }
#pragma unroll
      for (int k = 0; k < NCL; ++k) a.warm[(int64_t)(NZ + k * G + g) * LD + s] = st.w[k];
      a.warm[(int64_t)(NZ + BK::NC + g) * LD + s] = __longlong_as_double((long long)st.act);
    }
    // residual exits are polished; certified exits already are an exact KKT point
    if (sp.polish && __any_sync(0xffffffffu, good && !certified)) (void)admm_polish<BK>(qp, inv_alpha, st, good && !certified, sp.polish);

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
    // ---- hand om = [1 | v | p | centre of Ze[1]], cost, status to the output phase (warp-private buffer)
    if (g == 0) {
      wb.om[0][col] = 1.0;
#pragma unroll
      for (int j = 0; j < NZ; ++j) wb.om[BK::OM_V + j][col] = (j < nv) ? (good ? pg.D[j] * st.x[j] : NAN) : 0.0;
      wb.cost[col] = cost;
      wb.status[col] = live ? status : -1;
      wb.iters[col] = iters;
    }
    __syncwarp();
    for (int r = g; r < n; r += G) {        // centre of Ze[1]: rows split over the group
      const double* row = sCZ + r * NW;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NW; ++j) acc = fma(row[j], wb.om[j][col], acc);
      wb.om[BK::OM_C + r][col] = acc;
    }
   }     // solve tiles of this output tile
    if (explicit_qp) continue;
    __syncwarp();
    if (a.vec2) output_phase<BK, 2>(wb, pre, ax, a, tabd, tabi, otile, lane, zero_first);
    else output_phase<BK, 1>(wb, pre, ax, a, tabd, tabi, otile, lane, zero_first);
    __syncwarp();     // wb is rewritten by the next tile
  }
  if (a.stats != nullptr && a.x != nullptr) {
    __syncwarp();
    if (lane < TZ_NSTATS) {
      double v_ = 0.0;
#pragma unroll
      for (int c = 0; c < BK::SPO; ++c) v_ += wb.stacc[lane][c];
      if (v_ != 0.0) atomicAdd(a.stats + lane, v_);
    }
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
         2 * d.n <= BK::NPAR && general_atoms(d, nullptr) <= BK::NAG && d.nchk <= BK::NCHK;
}

template <class BK>
void pack(const TzProgramDesc& d, QpProg<BK>& g) {
  std::memset(&g, 0, sizeof(g));
  const int nz = d.nz, nc = d.nc, npar = d.npar, na = d.na, ncol = 1 + npar + na;
  std::vector<int> cls(nc), amap(na);
  classify(d, &cls);
  const int nag = general_atoms(d, &amap);
  // parameter k of the caller's p = [xbar0 (n); e0 (n)] -> slot of the padded p = [xbar0 (NPAR/2) | e0 (NPAR/2)]
  const int nx = d.n;
  auto pk = [&](int k) { return k < nx ? k : BK::NPAR / 2 + (k - nx); };
  // column j of the caller's [1 | p | alpha] layout -> column of the padded [1 | p | |p| | general] layout
  auto colmap = [&](int j) {
    if (j == 0) return 0;
    if (j <= npar) return 1 + pk(j - 1);
    const int am = amap[j - 1 - npar];
    return am < 0 ? 1 + BK::NPAR + pk(-am - 1) : 1 + 2 * BK::NPAR + am;
  };
  for (int a = 0; a < BK::NZ; ++a) g.D[a] = 1.0;
  for (int i = 0; i < BK::NC; ++i) {
    g.l0[i] = -INFINITY; g.u0[i] = INFINITY; g.Einv[i] = 1.0; g.row_of_slot[i] = -1; g.sing_var[i] = -1; g.sing_inv[i] = 0.0;
  }
  for (int a = 0; a < nz; ++a) {
    g.D[a] = d.D[a];
    g.q0[a] = d.c * d.D[a] * d.q0[a];
    for (int b = 0; b < nz; ++b) g.P[a][b] = d.c * d.D[a] * d.P[a * nz + b] * d.D[b];
    for (int k = 0; k < npar; ++k) {
      g.Qp[a][pk(k)] = d.c * d.D[a] * d.Qp[a * npar + k];
      if (d.Qp[a * npar + k] != 0.0) g.has_qp = 1;
    }
  }
  int next[3] = {0, BK::N2 * BK::G, (BK::N2 + BK::NU) * BK::G};
  for (int i = 0; i < nc; ++i) {
    const int s = next[cls[i]]++;
    const double E = d.E[i];
    g.row_of_slot[s] = i;
    g.Einv[s] = 1.0 / E;
    int nnz = 0, where = -1;
    for (int a = 0; a < nz; ++a) {
      g.A[s][a] = E * d.A[i * nz + a] * d.D[a];
      if (g.A[s][a] != 0.0) { ++nnz; where = a; }
    }
    if (nnz == 1) { g.sing_var[s] = where; g.sing_inv[s] = 1.0 / g.A[s][where]; }     // a bound on one variable
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
    for (int k = 0; k < npar; ++k) g.Bt[amap[i]][pk(k)] = d.Bt[i * npar + k];
  }
  for (int i = 0; i < d.nchk; ++i)
    for (int j = 0; j < ncol; ++j) g.Rchk[i][colmap(j)] += d.Rchk[i * ncol + j];
  for (int j = 0; j < ncol; ++j) g.cc[colmap(j)] += d.cc[j];
  for (int a = 0; a < npar; ++a)
    for (int b = 0; b < npar; ++b) g.CC2[pk(a)][pk(b)] = d.CC2[a * npar + b];
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
  int NW = 0, OM_V = 0, OM_P = 0, OM_C = 0, HP = 0;     // om layout of the bucket (Bucket::OM_*)
  size_t smem_tab = 0;                   // bytes of the run-time tables staged behind Smem<bucket>
  int num_sms = 148;
};

constexpr int kMaxTabBytes = 24 * 1024;

template <class BK>
static int create_bucket(const TzProgramDesc& d, TzProgram* p, int id) {
  p->bucket = id;
  std::vector<unsigned char> packed(sizeof(QpProg<BK>));
  pack<BK>(d, *reinterpret_cast<QpProg<BK>*>(packed.data()));
  p->NZ = BK::NZ;
  p->NC = BK::NC;
  p->G = BK::G;
  p->NW = BK::NW; p->OM_V = BK::OM_V; p->OM_P = BK::OM_P; p->OM_C = BK::OM_C; p->HP = BK::NPAR / 2;
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
  // ---- run-time sized tables: XB, centre map of Ze[1], K, and the non-zero entries of Ze[1].Z.  Columns / indices
  // of the caller's w = [1; v (nv); xbar0 (n); e0 (n)] are remapped to the kernel's padded vector om (Bucket::OM_*).
  const int n = d->n, nwc = 1 + d->nv + d->npar, ld1 = 1 + d->g1, nent = n * ld1, NW = p->NW;
  auto omidx = [&](int j) {
    if (j == 0) return 0;
    if (j <= d->nv) return p->OM_V + (j - 1);
    const int k = j - 1 - d->nv;
    return k < n ? p->OM_P + k : p->OM_P + p->HP + (k - n);
  };
  const int nrows = (d->horizon + 1) * n;
  std::vector<double> XB((size_t)nrows * NW, 0.0), CZ((size_t)n * NW, 0.0), coef;
  std::vector<int32_t> ent, idx;
  for (int i = 0; i < nrows; ++i)
    for (int j = 0; j < nwc; ++j) XB[(size_t)i * NW + omidx(j)] = d->XB[(size_t)i * nwc + j];
  for (int e = 0; e < nent; ++e) {
    const int t0 = d->ze1_ptr[e], t1 = d->ze1_ptr[e + 1];
    for (int t = t0; t < t1; ++t)
      if (d->ze1_idx[t] < 0 || d->ze1_idx[t] >= nwc) { delete p; return fail(TZ_EINVAL, "ze1_idx[%d] out of range", t); }
    const int r = e / ld1, j = e % ld1;
    if (j == 0) {                        // centre column: its (possibly many) terms become om[OM_C + r]
      for (int t = t0; t < t1; ++t) CZ[(size_t)r * NW + omidx(d->ze1_idx[t])] += d->ze1_val[t];
      ent.push_back(e); idx.push_back(p->OM_C + r); coef.push_back(1.0);
    } else if (t1 - t0 == 1) {
      ent.push_back(e); idx.push_back(omidx(d->ze1_idx[t0])); coef.push_back(d->ze1_val[t0]);
    } else if (t1 - t0 > 1) {
      delete p;
      return fail(TZ_EINVAL, "generator entry %d of Ze[1] has %d terms: only single-term generator entries are supported "
                  "(boxed M_K / M_Delta)", e, t1 - t0);
    }
  }
  const size_t nXB = XB.size(), nCZ = CZ.size(), nK = (size_t)d->m * n, nco = coef.size();
  const size_t ndbl = nXB + nCZ + nK + nco, nint = ent.size() + idx.size();
  const size_t smem_tab = (ndbl + (size_t)n * n + (size_t)n * d->m) * sizeof(double) + nint * sizeof(int32_t);
  if (smem_tab > kMaxTabBytes) {
    delete p;
    return fail(TZ_ERANGE, "program tables need %zu bytes of shared memory (limit %d)", smem_tab, kMaxTabBytes);
  }
  std::vector<unsigned char> host(ndbl * sizeof(double) + nint * sizeof(int32_t) + 16);
  double* hd = reinterpret_cast<double*>(host.data());
  std::memcpy(hd, XB.data(), nXB * sizeof(double));
  std::memcpy(hd + nXB, CZ.data(), nCZ * sizeof(double));
  std::memcpy(hd + nXB + nCZ, d->K, nK * sizeof(double));
  if (nco) std::memcpy(hd + nXB + nCZ + nK, coef.data(), nco * sizeof(double));
  int32_t* hi = reinterpret_cast<int32_t*>(hd + ndbl);
  if (!ent.empty()) std::memcpy(hi, ent.data(), ent.size() * sizeof(int32_t));
  if (!idx.empty()) std::memcpy(hi + ent.size(), idx.data(), idx.size() * sizeof(int32_t));
  cudaError_t err = cudaMalloc(&p->aux_dev, host.size());
  if (err == cudaSuccess) err = cudaMemcpy(p->aux_dev, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    tz_program_destroy(p);
    return fail(TZ_ECUDA, "aux upload: %s", cudaGetErrorString(err));
  }
  Aux& ax = p->aux;
  ax.tab = reinterpret_cast<const double*>(p->aux_dev);
  ax.n_dbl = (int)ndbl; ax.n_int = (int)nint;
  ax.o_XB = 0; ax.o_CZ = (int)nXB; ax.o_K = (int)(nXB + nCZ); ax.o_coef = (int)(nXB + nCZ + nK);
  ax.o_ent = 0; ax.o_idx = (int)ent.size();
  ax.n_nz = (int)ent.size();
  ax.n = n; ax.m = d->m; ax.N = d->horizon; ax.nv = d->nv; ax.g1 = d->g1;
  p->smem_tab = (smem_tab + 15) & ~(size_t)15;
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
  o->cert_first = 3;
}

static SolverParams to_params(const TzSolverOpts* o) {
  TzSolverOpts d;
  tz_solver_opts_default(&d);
  if (o) d = *o;
  return SolverParams{d.rho, d.rho_active, d.rho_inactive, d.sigma, d.alpha, d.eps_abs, d.eps_rel,
                      d.max_iter, d.check_every, d.polish, d.warm_start, d.cert_first};
}

template <class BK>
static int launch_bucket(const TzProgram* p, const SolverParams& sp, const StepArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(Smem<BK>) + p->smem_tab;
  static bool configured = false;     // benign race: the attribute is idempotent
  if (!configured) {
    TZ_CUDA(cudaFuncSetAttribute(step_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(Smem<BK>) + kMaxTabBytes)));
    configured = true;
  }
  // persistent grid: one wave of CTAs (MINB per SM); every warp loops over tiles of SPW scenarios
  const int64_t ntiles = (a.S + BK::SPO - 1) / BK::SPO;
  const int64_t need = (ntiles + BK::WPB - 1) / BK::WPB;
  const int64_t wave = (int64_t)p->num_sms * BK::MINB;
  const unsigned grid = (unsigned)(need < wave ? need : wave);
  step_kernel<BK><<<grid, BK::TPB, smem, st>>>(reinterpret_cast<const QpProg<BK>*>(p->packed_dev), p->aux, sp, a);
  TZ_CUDA(cudaGetLastError());
  return TZ_OK;
}

static int launch(const TzProgram* p, const TzSolverOpts* o, const StepArgs& a_in, void* stream) {
  TZ_REQUIRE(p != nullptr, "null program");
  TZ_REQUIRE(a_in.S >= 0, "negative batch");
  if (a_in.S == 0) return TZ_OK;
  StepArgs a = a_in;
  {  // two scenarios per lane in the output phase need 16-byte aligned rows: S, ld even and aligned base pointers
    const void* ptrs[] = {a.x, a.xbar, a.e, a.noise, a.x_restart, a.cost, a.v, a.xbar_traj, a.ze1, a.u_out};
    bool ok = (a.S % 2 == 0) && (a.ld % 2 == 0);
    for (const void* q : ptrs) ok = ok && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0);
    ok = ok && ((reinterpret_cast<uintptr_t>(a.status) & 7u) == 0) && ((reinterpret_cast<uintptr_t>(a.iters) & 7u) == 0);
    a.vec2 = ok ? 1 : 0;
  }
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
                                   double* e, const double* noise, const double* x_restart, const double* A_true,
                                   const double* B_true,
                                   double* cost, double* v, double* xbar_traj, double* ze1, double* u_out,
                                   int32_t* status, int32_t* iters, double* warm, double* stats, void* stream) {
  TZ_REQUIRE(S == 0 || (x && xbar && e && A_true && B_true && status), "x, xbar, e, A_true, B_true, status are required");
  StepArgs a{};
  a.S = S; a.ld = S; a.xbar0 = xbar; a.e0 = e; a.x = x; a.xbar = xbar; a.e = e; a.noise = noise; a.x_restart = x_restart; a.A_true = A_true; a.B_true = B_true;
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
  int64_t per = (S + nchunks - 1) / nchunks;
  per = (per + 15) & ~(int64_t)15;          // whole tiles, 16-byte aligned chunk starts
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
