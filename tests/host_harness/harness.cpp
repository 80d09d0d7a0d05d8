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

int hh_junction_newton(const double* vp, int Nx, const double* A, const double* q, double dt, const double* dex,
                       double* x, int k_max, double tol, unsigned* flags) {
  JunctionCtx c = make_ctx(vp, Nx, A, q, dt, dex);
  uint32_t fl = 0;
  int s = junction_newton(c, x, k_max, tol, &fl);
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

// Serial stand-in for k_artery (same phases: cell integrals -> vertex rows ->
// level-blocked cyclic reduction).  bc = A_in, q_in, A_out, q_out.  On entry A, q
// hold Un (= the initial guess U); on return they hold U.  If Rout/Jout are given,
// the first assembled residual [np][2] and blocks [3][np][4] (Lo, Dg, Up; Dirichlet
// rows applied, true diagonal blocks) are copied out in vertex order.
int hh_artery_solve(const double* vp, int Nx, double dt, double theta, double* A, double* q, const double* bc,
                    int root, int leaf, double rtol, double atol, int max_it, double* r_last, double* Rout,
                    double* Jout) {
  Vessel v = make_vessel(vp, Nx);
  const int np = Nx + 1, ne = Nx, S = np;
  std::vector<double> EA(S), Eq(S), Lo(4 * S), Dg(4 * S), Up(4 * S), R(2 * S);
  std::vector<CellInt> cells(ne);
  CrLayout lay;
  cr_layout_init(lay, np);
  const double h = v.dx, Kr = -v.src_k, dth = dt * theta, dte = dt * (1.0 - theta);
  const double c1_0 = v.f0 * sqrt(v.A00), c1_L = v.fL * sqrt(v.A0L);
  int it = 0;
  double r_first = 0.0, r = 0.0;
  for (;;) {
    for (int e = 0; e < ne; ++e) {
      CellCoef cc;
      if (v.tapered) {
        cc = cell_coef(v, v.dx * e, v.dx * (e + 1));
      } else {
        for (int g = 0; g < 3; ++g) { cc.c1[g] = c1_0; cc.t2[g] = 0.0; cc.t3[g] = 0.0; }
      }
      cells[e] = cell_integrals<true>(h, Kr, cc, A[e], A[e + 1], q[e], q[e + 1]);
    }
    double part = 0.0;
    for (int i = 0; i < np; ++i) {
      CellInt zero;
      zero.Gl = zero.Gr = 0.0;
      for (int c = 0; c < 4; ++c) zero.JA[c] = zero.Jq[c] = 0.0;
      const CellInt& left = i > 0 ? cells[i - 1] : zero;
      const CellInt& own = i < ne ? cells[i] : zero;
      const bool end = (i == 0 || i == ne);
      double fF2 = 0.0, fdA = 0.0, fdq = 0.0;
      if (end) {
        PointTerms ft = facet_terms<true>(i == 0 ? c1_0 : c1_L, A[i], q[i]);
        fF2 = ft.F2; fdA = ft.dF2dA; fdq = ft.dF2dq;
      }
      VertexX vx = vertex_x(i, ne, h, i > 0 ? A[i - 1] : 0.0, i > 0 ? q[i - 1] : 0.0, A[i], q[i],
                            i < ne ? A[i + 1] : 0.0, i < ne ? q[i + 1] : 0.0, left.Gr, own.Gl, fF2);
      if (it == 0) { EA[i] = -vx.mA + dte * vx.XA; Eq[i] = -vx.mq + dte * vx.Xq; }
      double r0 = vx.mA + dth * vx.XA + EA[i], r1 = vx.mq + dth * vx.Xq + Eq[i];
      VertexRow row = vertex_jacobian(i, ne, h, dth, left.JA, left.Jq, own.JA, own.Jq, fdA, fdq);
      if (end) {
        bool fixA = (i == 0) ? !root : true;
        bool fixq = (i == 0) ? true : !leaf;
        double gA = (i == 0) ? bc[0] : bc[2], gq = (i == 0) ? bc[1] : bc[3];
        if (fixA) { r0 = A[i] - gA; row.dg[0] = 1.0; row.dg[1] = 0.0; row.lo[0] = row.lo[1] = row.up[0] = row.up[1] = 0.0; }
        if (fixq) { r1 = q[i] - gq; row.dg[2] = 0.0; row.dg[3] = 1.0; row.lo[2] = row.lo[3] = row.up[2] = row.up[3] = 0.0; }
      }
      part += r0 * r0 + r1 * r1;
      if (it == 0 && Rout) { Rout[2 * i] = r0; Rout[2 * i + 1] = r1; }
      if (it == 0 && Jout)
        for (int c = 0; c < 4; ++c) {
          Jout[(0 * np + i) * 4 + c] = row.lo[c];
          Jout[(1 * np + i) * 4 + c] = row.dg[c];
          Jout[(2 * np + i) * 4 + c] = row.up[c];
        }
      const int p = cr_pos(lay, i);
      if ((i & 1) == 0) { double di[4]; inv2v(row.dg, di); st2(Dg.data(), S, p, di); }
      else st2(Dg.data(), S, p, row.dg);
      st2(Lo.data(), S, p, row.lo);
      st2(Up.data(), S, p, row.up);
      R[p] = r0; R[S + p] = r1;
    }
    r = sqrt(part);
    if (it == 0) r_first = r;
    if (r / r_first < rtol || r < atol) break;
    if (!(r == r) || it >= max_it) { it = -1 - it; break; }
    for (int l = 0; (np >> (l + 1)) > 0; ++l)
      for (int m = 0; m < (np >> (l + 1)); ++m) cr_forward(lay, l, m, Lo.data(), Dg.data(), Up.data(), R.data(), S);
    for (int l = lay.nlev - 1; l >= 0; --l)
      for (int k = 0; k < lay.cnt[l]; ++k) cr_backward(lay, l, k, Lo.data(), Dg.data(), Up.data(), R.data(), S);
    for (int i = 0; i < np; ++i) {
      const int p = cr_pos(lay, i);
      A[i] -= R[p]; q[i] -= R[S + p];
    }
    ++it;
  }
  if (r_last) *r_last = r;
  return it;
}

}  // extern "C"
