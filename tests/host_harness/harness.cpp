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

// Serial stand-in for k_artery.  bc = A_in, q_in, A_out, q_out.  On return A, q
// hold U.  If Rout/Jout are given, the first assembled residual [np][2] and
// blocks [3][np][4] (Lo, Dg, Up) are copied out (Dirichlet rows applied).
int hh_artery_solve(const double* vp, int Nx, double dt, double theta, double* A, double* q, const double* bc,
                    int root, int leaf, double rtol, double atol, int max_it, double* r_last, double* Rout,
                    double* Jout) {
  Vessel v = make_vessel(vp, Nx);
  const int np = Nx + 1, ne = Nx, S = np;
  std::vector<double> An(A, A + np), qn(q, q + np);
  std::vector<double> Lo(4 * S), Dg(4 * S), Up(4 * S), R(2 * S), D1(4 * S), Rr(2 * S);
  ArteryConst ac;
  ac.h = v.dx; ac.dt = dt; ac.theta = theta; ac.Kr = -v.src_k;
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
      double rl[2], rr[2], b00[4], b01[4], b10[4], b11[4];
      cell_assemble<true>(ac, cc, A[e], A[e + 1], q[e], q[e + 1], An[e], An[e + 1], qn[e], qn[e + 1], rl, rr, b00,
                          b01, b10, b11);
      R[e] = rl[0]; R[S + e] = rl[1];
      Rr[e + 1] = rr[0]; Rr[S + e + 1] = rr[1];
      for (int c = 0; c < 4; ++c) {
        Dg[c * S + e] = b00[c]; Up[c * S + e] = b01[c];
        Lo[c * S + e + 1] = b10[c]; D1[c * S + e + 1] = b11[c];
      }
    }
    double part = 0.0;
    for (int i = 0; i < np; ++i) {
      double r0 = 0.0, r1 = 0.0, d[4] = {0, 0, 0, 0};
      if (i < ne) {
        r0 = R[i]; r1 = R[S + i];
        for (int c = 0; c < 4; ++c) d[c] = Dg[c * S + i];
      } else {
        for (int c = 0; c < 4; ++c) Up[c * S + i] = 0.0;
      }
      if (i > 0) {
        r0 += Rr[i]; r1 += Rr[S + i];
        for (int c = 0; c < 4; ++c) d[c] += D1[c * S + i];
      } else {
        for (int c = 0; c < 4; ++c) Lo[c * S + i] = 0.0;
      }
      if (i == 0 || i == ne) {
        double rq, dA, dq;
        facet_term<true>(ac, i == 0 ? c1_0 : c1_L, A[i], q[i], An[i], qn[i], rq, dA, dq);
        r1 += rq; d[2] += dA; d[3] += dq;
        bool fixA = (i == 0) ? !root : true;
        bool fixq = (i == 0) ? true : !leaf;
        double gA = (i == 0) ? bc[0] : bc[2], gq = (i == 0) ? bc[1] : bc[3];
        if (fixA) {
          r0 = A[i] - gA; d[0] = 1.0; d[1] = 0.0;
          Lo[0 * S + i] = Lo[1 * S + i] = Up[0 * S + i] = Up[1 * S + i] = 0.0;
        }
        if (fixq) {
          r1 = q[i] - gq; d[2] = 0.0; d[3] = 1.0;
          Lo[2 * S + i] = Lo[3 * S + i] = Up[2 * S + i] = Up[3 * S + i] = 0.0;
        }
      }
      R[i] = r0; R[S + i] = r1;
      for (int c = 0; c < 4; ++c) Dg[c * S + i] = d[c];
      part += r0 * r0 + r1 * r1;
    }
    r = sqrt(part);
    if (it == 0 && Rout) {
      for (int i = 0; i < np; ++i) { Rout[2 * i] = R[i]; Rout[2 * i + 1] = R[S + i]; }
      Rout = nullptr;
    }
    if (it == 0 && Jout) {
      const std::vector<double>* m[3] = {&Lo, &Dg, &Up};
      for (int k = 0; k < 3; ++k)
        for (int i = 0; i < np; ++i)
          for (int c = 0; c < 4; ++c) Jout[(k * np + i) * 4 + c] = (*m[k])[c * S + i];
      Jout = nullptr;
    }
    if (it == 0) r_first = r;
    if (r / r_first < rtol || r < atol) break;
    if (!(r == r) || it >= max_it) { it = -1 - it; break; }
    int s = 1;
    for (; 2 * s - 1 < np; s *= 2)
      for (int i = 2 * s - 1; i < np; i += 2 * s) cr_forward(i, s, np, Lo.data(), Dg.data(), Up.data(), R.data(), S);
    for (; s >= 1; s /= 2)
      for (int i = s - 1; i < np; i += 2 * s) cr_backward(i, s, np, Lo.data(), Dg.data(), Up.data(), R.data(), S);
    for (int i = 0; i < np; ++i) { A[i] -= R[i]; q[i] -= R[S + i]; }
    ++it;
  }
  if (r_last) *r_last = r;
  return it;
}

}  // extern "C"
