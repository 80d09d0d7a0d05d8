// Host harness: compiles bloodflow_b200/csrc/af_physics.cuh with g++ so the
// per-thread formulas the CUDA kernels run can be unit-tested against the
// oracle in the CPU-only build container (pytest -m "not gpu").
//
// TEST TOOL ONLY.  It is not linked into libarteryfe_b200.so, is not reachable
// from the product API and is not a CPU fallback: the serial loops below stand
// in for the thread/block orchestration of af_kernels.cu (cells -> vertices ->
// block cyclic reduction), which itself only exists on the device.
#include <string.h>

#include <vector>

#include "../../bloodflow_b200/csrc/af_physics.cuh"

using namespace af;

static Vessel make_vessel(const double* p, int Nx) {
  // p = Ru, Rd, L, k1, k2, k3, rho, Re, db, p0   (mirrors k_setup_vessels)
  Vessel v;
  v.Ru = p[0]; v.Rd = p[1]; v.L = p[2];
  v.dx = v.L / Nx;
  v.k1 = p[3]; v.k2 = p[4]; v.k3 = p[5];
  v.lr_over_L = log(v.Rd / v.Ru) / v.L;
  v.rho = p[6];
  v.p0 = p[9];
  v.src_k = -2.0 * kSqrtPi / p[8] / p[7];
  v.rpi_dbRe = kSqrtPi / p[8] / p[7];
  v.tapered = (v.Rd != v.Ru) ? 1 : 0;
  Coef c0 = coef_exact(v, 0.0), cL = coef_exact(v, v.L);
  v.f0 = c0.f; v.A00 = c0.A0; v.fL = cL.f; v.A0L = cL.A0; v.dfdr0 = c0.dfdr;
  return v;
}

extern "C" {

void hh_coef(const double* vp, int Nx, int n, const double* x, double* out) {
  Vessel v = make_vessel(vp, Nx);
  for (int i = 0; i < n; ++i) {
    Coef c = coef_exact(v, x[i]);
    out[4 * i] = c.f; out[4 * i + 1] = c.A0; out[4 * i + 2] = c.dfdr; out[4 * i + 3] = c.drdx;
  }
}

void hh_interp(const double* vp, int Nx, const double* A, const double* q, int n, const double* x, double* out) {
  Vessel v = make_vessel(vp, Nx);
  for (int i = 0; i < n; ++i) interp_state(A, q, Nx, v.dx, x[i], out[2 * i], out[2 * i + 1]);
}

// vp: [3][10]; A, q: [3][Nx+1]
static JunctionCtx make_ctx(const double* vp, int Nx, const double* A, const double* q, double dt,
                            const double* dex) {
  JunctionCtx c;
  for (int v = 0; v < 3; ++v) {
    Vessel ves = make_vessel(vp + 10 * v, Nx);
    junction_prepare_vessel(c, v, ves, A + v * (Nx + 1), q + v * (Nx + 1), Nx, dt, dex[v]);
  }
  return c;
}

double hh_junction_cfl_dex(const double* vp, int Nx, const double* A, const double* q, double dt, double margin) {
  double M = 1e300;
  for (int v = 0; v < 3; ++v) {
    Vessel ves = make_vessel(vp + 10 * v, Nx);
    double m = junction_cfl(ves, v, A + v * (Nx + 1), q + v * (Nx + 1), Nx);
    M = m < M ? m : M;
  }
  return (1.0 + margin) * dt / M;
}

void hh_junction_eval(const double* vp, int Nx, const double* A, const double* q, double dt, const double* dex,
                      const double* x, double* y, double* J) {
  JunctionCtx c = make_ctx(vp, Nx, A, q, dt, dex);
  junction_residual(c, x, y);
  junction_jacobian(c, x, J, 18);
}

// mode 0: hot path (per-vessel rsqrt linearisation + structured solve);
// mode 1: verbatim dense path (reference formulas, 18x18 LU)
int hh_junction_newton(const double* vp, int Nx, const double* A, const double* q, double dt, const double* dex,
                       double* x, int k_max, double tol, unsigned* flags, int mode) {
  uint32_t fl = 0;
  int s;
  if (mode == 1) {
    JunctionCtx c = make_ctx(vp, Nx, A, q, dt, dex);
    s = junction_newton(c, x, k_max, tol, &fl);
  } else {
    JVessel jv[3];
    for (int v = 0; v < 3; ++v) {
      Vessel ves = make_vessel(vp + 10 * v, Nx);
      jvessel_prepare(jv[v], v, ves, A + v * (Nx + 1), q + v * (Nx + 1), Nx, dt, dex[v]);
    }
    s = junction_newton_fast(jv, x, k_max, tol, &fl);
  }
  *flags = fl;
  return s;
}

void hh_lu_solve18(double* J, double* b, int* singular) { *singular = lu_solve18(J, 18, b); }

void hh_windkessel(const double* vp, int Nx, const double* A, const double* q, double dt, double R1, double R2,
                   double CT, int k_max, double tol, double margin, double* out) {
  Vessel v = make_vessel(vp, Nx);
  WkResult r = windkessel(v, A, q, Nx, dt, R1, R2, CT, k_max, tol, margin);
  out[0] = r.A_out; out[1] = r.dex; out[2] = r.its;
}

// Serial stand-in for k_artery: T "threads" of four vertices each, executed
// phase by phase in the order the kernel's barriers impose (cells -> rows ->
// levels 0/1 in "registers" -> reduced level-blocked cyclic reduction ->
// back-substitution).  bc = A_in, q_in, A_out, q_out.  On entry A, q hold Un (= the
// initial guess U); on return they hold U.  If Rout/Jout are given, the first
// assembled residual [np][2] and blocks [3][np][4] (Lo, Dg, Up; Dirichlet rows
// applied) are copied out in vertex order.
int hh_artery_solve(const double* vp, int Nx, double dt, double theta, double* A, double* q, const double* bc,
                    int root, int leaf, double rtol, double atol, int max_it, double* r_last, double* Rout,
                    double* Jout) {
  Vessel v = make_vessel(vp, Nx);
  const int np = Nx + 1, ne = Nx;
  int T = 32;
  while (4 * T < np) T *= 2;
  const int NV = 4 * T;
  std::vector<double> sA(NV, 1.0), sq(NV, 0.0), EA(NV), Eq(NV);
  for (int i = 0; i < np; ++i) { sA[i] = A[i]; sq[i] = q[i]; }
  std::vector<Row> rows(NV);
  std::vector<Elim> ea(T), eb(T), ec(T);
  ArteryCtx c;
  c.ne = ne; c.np = np; c.h = v.dx; c.Kr = -v.src_k; c.dth = dt * theta; c.dte = dt * (1.0 - theta);
  c.c1_0 = v.f0 * sqrt(v.A00); c.c1_L = v.fL * sqrt(v.A0L);
  c.root = root; c.leaf = leaf;
  c.gAin = bc[0]; c.gqin = bc[1]; c.gAout = bc[2]; c.gqout = bc[3];
  auto cell = [&](int i) {
    CellInt ci;
    cell_zero(ci);
    if (i < ne) {
      CellCoef cc;
      if (v.tapered) {
        cc = cell_coef(v, v.dx * i, v.dx * (i + 1));
      } else {
        for (int g = 0; g < 3; ++g) { cc.c1[g] = c.c1_0; cc.t2[g] = 0.0; cc.t3[g] = 0.0; }
      }
      ci = cell_integrals<true>(c.h, c.Kr, cc, sA[i], sA[i + 1], sq[i], sq[i + 1]);
    }
    return ci;
  };
  int it = 0;
  double r_first = 0.0, r = 0.0;
  for (;;) {
    const bool first = (it == 0);
    double part = 0.0;
    for (int i = 0; i < NV; ++i) {
      CellInt left;
      cell_zero(left);
      if (i > 0) left = cell(i - 1);
      CellInt own = cell(i);
      build_row<true>(c, i, first, i > 0 ? sA[i - 1] : 0.0, i > 0 ? sq[i - 1] : 0.0, sA[i], sq[i],
                      i + 1 < (int)sA.size() ? sA[i + 1] : 0.0, i + 1 < (int)sq.size() ? sq[i + 1] : 0.0, left, own,
                      EA[i], Eq[i], rows[i]);
      part += rows[i].r[0] * rows[i].r[0] + rows[i].r[1] * rows[i].r[1];
      if (first && i < np) {
        if (Rout) { Rout[2 * i] = rows[i].r[0]; Rout[2 * i + 1] = rows[i].r[1]; }
        if (Jout)
          for (int k = 0; k < 4; ++k) {
            Jout[(0 * np + i) * 4 + k] = rows[i].lo[k];
            Jout[(1 * np + i) * 4 + k] = rows[i].dg[k];
            Jout[(2 * np + i) * 4 + k] = rows[i].up[k];
          }
      }
    }
    r = sqrt(part);
    if (first) r_first = r;
    if (r / r_first < rtol || r < atol) break;
    if (!(r == r) || it >= max_it) { it = -1 - it; break; }
    for (int t = 0; t < T; ++t) {          // level 0, own part
      Row &ra = rows[4 * t], &rb = rows[4 * t + 1], &rc = rows[4 * t + 2], &rd = rows[4 * t + 3];
      ea[t] = make_elim(ra);
      apply_left(rb, ea[t]);
      ec[t] = make_elim(rc);
      apply_right(rb, ec[t]);
      apply_left(rd, ec[t]);
    }
    for (int t = 0; t < T; ++t) {          // level 0 across threads, level 1 own part
      Row &rb = rows[4 * t + 1], &rd = rows[4 * t + 3];
      if (t + 1 < T) apply_right(rd, ea[t + 1]);
      eb[t] = make_elim(rb);
      apply_left(rd, eb[t]);
    }
    for (int t = 0; t < T; ++t)            // level 1 across threads
      if (t + 1 < T) apply_right(rows[4 * t + 3], eb[t + 1]);
    // reduced system: parallel cyclic reduction, one row per "thread"
    std::vector<Elim> buf(T);
    for (int sdist = 1; sdist < T; sdist *= 2) {
      for (int t = 0; t < T; ++t) buf[t] = make_elim(rows[4 * t + 3]);
      for (int t = 0; t < T; ++t) {
        Row& rd = rows[4 * t + 3];
        if (t >= sdist) apply_left(rd, buf[t - sdist]);
        if (t + sdist < T) apply_right(rd, buf[t + sdist]);
      }
    }
    std::vector<double> X(2 * T);
    for (int t = 0; t < T; ++t) {
      const Row& rd = rows[4 * t + 3];
      double inv[4];
      inv2v(rd.dg, inv);
      X[t] = inv[0] * rd.r[0] + inv[1] * rd.r[1];
      X[T + t] = inv[2] * rd.r[0] + inv[3] * rd.r[1];
    }
    for (int t = 0; t < T; ++t) {
      double xl[2] = {0.0, 0.0}, xd[2] = {X[t], X[T + t]}, xa[2], xb[2], xc[2];
      if (t > 0) { xl[0] = X[t - 1]; xl[1] = X[T + t - 1]; }
      back_sub(eb[t], xl, xd, xb);
      back_sub(ea[t], xl, xb, xa);
      back_sub(ec[t], xb, xd, xc);
      const int i0 = 4 * t;
      sA[i0] -= xa[0]; sq[i0] -= xa[1];
      sA[i0 + 1] -= xb[0]; sq[i0 + 1] -= xb[1];
      sA[i0 + 2] -= xc[0]; sq[i0 + 2] -= xc[1];
      sA[i0 + 3] -= xd[0]; sq[i0 + 3] -= xd[1];
    }
    ++it;
  }
  for (int i = 0; i < np; ++i) { A[i] = sA[i]; q[i] = sq[i]; }
  if (r_last) *r_last = r;
  return it;
}

}  // extern "C"
