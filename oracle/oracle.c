/* oracle.c -- plain-C restatement of artery.fe's per-timestep network solve.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for larger
 * parity cases and the timed CPU baseline ("restated reference", never
 * "FEniCS") of bench.py.  It follows oracle/arteryfe_oracle.py, which is pinned
 * against the reference's own code (tests/test_oracle_golden.py) and is itself
 * checked against this file in tests/test_oracle_c.py.  It shares no code with
 * the product (bloodflow_b200/csrc): formulas are written as the reference
 * writes them, the linear solve is a banded LU with partial pivoting (stand-in
 * for MUMPS, artery.py:241), compiled -O2 -ffp-contract=off, no fast-math.
 *
 * Citations: an = arteryfe/artery_network.py, ar = arteryfe/artery.py of the
 * reference; [3P] = FEniCS 2017.2.0 semantics.
 * OpenMP parallelism is over junctions/leaves and over arteries (the two
 * bulk-parallel phases of a step); the reference itself is serial.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DOLFIN_EPS 3.0e-16
static const double GXI[3] = {0.5 - 0.5 * 0.7745966692414834, 0.5, 0.5 + 0.5 * 0.7745966692414834};
static const double GW[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

typedef struct {
  int na, nj, nl, Nx, np, Nt;
  double dt, theta, k1, k2, k3, rho, Re, db, p0;
  int *jp, *j1, *j2, *leaf, *art_leaf;
  double *Ru, *Rd, *L, *dx, *q0, *R1, *R2, *CT, *q_ins;
  double *A, *q;       /* Un  [na][np] */
  double *UA, *Uq;     /* U   [na][np] */
  double *bc;          /* [na][4] A_in q_in A_out q_out */
  double *dex;         /* [na] */
  double *x;           /* [nj][18] */
  double *cf, *cA0, *cdf, *cdr; /* P2 interpolants at Gauss points [na][Nx][3] */
  int *art_its, *j_solves, *p_its;
  long long tot[3];
  int failed;
  /* constants (defaults of the reference) */
  double rtol, atol, jtol, ptol, margin;
  int max_it, jk_max, pk_max;
} orc_net;

/* ---- Expressions (ar:110-118) ------------------------------------------------ */
static double e_r0(const orc_net* n, int a, double x) { return n->Ru[a] * pow(n->Rd[a] / n->Ru[a], x / n->L[a]); }
static double e_A0(const orc_net* n, int a, double x) { return M_PI * pow(e_r0(n, a, x), 2); }
static double e_f(const orc_net* n, int a, double x) { return 4.0 / 3.0 * (n->k1 * exp(n->k2 * e_r0(n, a, x)) + n->k3); }
static double e_dfdr(const orc_net* n, int a, double x) { return 4.0 / 3.0 * n->k1 * n->k2 * exp(n->k2 * e_r0(n, a, x)); }
static double e_drdx(const orc_net* n, int a, double x) { return log(n->Rd[a] / n->Ru[a]) / n->L[a] * e_r0(n, a, x); }

/* [3P] Function.__call__: P1 interpolation of Un */
static void un_at(const orc_net* n, int a, double x, double U[2]) {
  double dx = n->dx[a];
  int j = (int)floor(x / dx);
  if (j < 0) j = 0;
  if (j > n->Nx - 1) j = n->Nx - 1;
  double xi = (x - j * dx) / dx;
  const double* A = n->A + (size_t)a * n->np;
  const double* q = n->q + (size_t)a * n->np;
  U[0] = A[j] * (1.0 - xi) + A[j + 1] * xi;
  U[1] = q[j] * (1.0 - xi) + q[j + 1] * xi;
}

/* ---- boundary physics (an:282-358) -------------------------------------------- */
static void flux(const orc_net* n, int a, const double U[2], double x, double F[2]) {
  F[0] = U[1];
  F[1] = U[1] * U[1] + e_f(n, a, x) * sqrt(e_A0(n, a, x) * U[0]); /* no /A: an:300 */
}

static void source(const orc_net* n, int a, const double U[2], double x, double S[2]) {
  double rpi = sqrt(M_PI);
  S[0] = 0.0;
  S[1] = -2 * rpi / n->db / n->Re * U[1] / sqrt(U[0]) +
         (2 * sqrt(U[0]) * (rpi * e_f(n, a, x) + sqrt(e_A0(n, a, x)) * e_dfdr(n, a, x)) - U[0] * e_dfdr(n, a, x)) *
             e_drdx(n, a, x);
}

static void u_half(const orc_net* n, int a, double x0, double x1, const double U0[2], const double U1[2],
                   double Uh[2]) {
  double F0[2], F1[2], S0[2], S1[2];
  flux(n, a, U0, x0, F0); source(n, a, U0, x0, S0);
  flux(n, a, U1, x1, F1); source(n, a, U1, x1, S1);
  for (int c = 0; c < 2; ++c)
    Uh[c] = (U0[c] + U1[c]) / 2 - n->dt / (x1 - x0) * (F1[c] - F0[c]) + n->dt / 4 * (S0[c] + S1[c]);
}

static double cfl_term(const orc_net* n, int a, double x, double A, double q) { /* ar:304-325 */
  double c = sqrt(e_f(n, a, x) / 2 / n->rho * sqrt(e_A0(n, a, x) / A));
  double p = fabs(q / A + c), m = fabs(q / A - c);
  return 1 / (p > m ? p : m);
}

static double outlet_pressure(const orc_net* n, int a, double A) { /* ar:285-301 */
  double L = n->L[a];
  return n->p0 + e_f(n, a, L) * (1 - sqrt(e_A0(n, a, L) / A));
}

/* windkessel (an:361-422); R1/R2 already carry HEAD's factors */
static double windkessel(orc_net* n, int l) {
  int a = n->leaf[l];
  double L = n->L[a], UL[2];
  un_at(n, a, L, UL);
  n->dex[a] = (1 + n->margin) * n->dt / cfl_term(n, a, L, UL[0], UL[1]);
  double dex = n->dex[a];
  double x2 = L - 2 * dex, x1 = L - dex, x0 = L, x21 = L - 1.5 * dex, x10 = L - 0.5 * dex;
  double Um2[2], Um1[2], Um0[2], Uh21[2], Uh10[2], F21[2], S21[2], F10[2], S10[2];
  un_at(n, a, x2, Um2); un_at(n, a, x1, Um1); un_at(n, a, x0, Um0);
  u_half(n, a, x2, x1, Um2, Um1, Uh21);
  u_half(n, a, x1, x0, Um1, Um0, Uh10);
  flux(n, a, Uh21, x21, F21); source(n, a, Uh21, x21, S21);
  flux(n, a, Uh10, x10, F10); source(n, a, Uh10, x10, S10);
  double qm1 = Um1[1] - n->dt / dex * (F10[1] - F21[1]) + n->dt / 2 * (S10[1] + S21[1]);
  double pn = outlet_pressure(n, a, Um0[0]), p = pn;
  double R1 = n->R1[l], R2 = n->R2[l], CT = n->CT[l], Am0 = Um0[0];
  int k = 0;
  while (k < n->pk_max) {
    double p_old = p;
    double qm0 = Um0[1] + (p - pn) / R1 + n->dt / R1 / R2 / CT * pn - n->dt * (R1 + R2) / R1 / R2 / CT * Um0[1];
    Am0 = Um0[0] - n->dt / dex * (qm0 - qm1);
    p = outlet_pressure(n, a, Am0);
    ++k;
    if (fabs(p - p_old) < n->ptol) break;
  }
  n->p_its[l] = k;
  return Am0;
}

/* ---- bifurcation (an:476-710) ---------------------------------------------------- */
typedef struct { int a[3]; double dex[3]; } jn_t;

static void jn_points(const orc_net* n, const jn_t* J, int v, double* xe, double* xi, double* xg, double* xh,
                      double* xj) {
  double L = n->L[J->a[v]], d = J->dex[v];
  if (v == 0) { *xe = L; *xi = L - d; *xg = L + d / 2; *xh = L - d / 2; *xj = L + d; }
  else { *xe = 0; *xi = d; *xg = -d / 2; *xh = d / 2; *xj = -d; }
}

static void problem_function(const orc_net* n, const jn_t* J, const double* x, double* y) {
  double ph[3], pf[3];
  for (int v = 0; v < 3; ++v) {
    int a = J->a[v];
    double s = v == 0 ? 1.0 : -1.0, xe, xi, xg, xh, xj;
    jn_points(n, J, v, &xe, &xi, &xg, &xh, &xj);
    double Ue[2], Ui[2], Uh[2], Ug[2] = {x[11 + 3 * v], x[2 + 3 * v]}, Fg[2], Sg[2], Fh[2], Sh[2];
    un_at(n, a, xe, Ue); un_at(n, a, xi, Ui);
    if (v == 0) u_half(n, a, xi, xe, Ui, Ue, Uh); else u_half(n, a, xe, xi, Ue, Ui, Uh);
    flux(n, a, Ug, xg, Fg); source(n, a, Ug, xg, Sg);
    flux(n, a, Uh, xh, Fh); source(n, a, Uh, xh, Sh);
    double r = n->dt / J->dex[v];
    y[v] = 2 * x[3 * v + 1] - Uh[1] - x[3 * v + 2];
    y[3 + v] = 2 * x[10 + 3 * v] - Uh[0] - x[11 + 3 * v];
    y[12 + v] = x[3 * v] - Ue[1] + s * r * (Fg[1] - Fh[1]) - n->dt / 2 * (Sg[1] + Sh[1]);
    y[15 + v] = x[9 + 3 * v] - Ue[0] + s * r * (Fg[0] - Fh[0]);
    double fe = e_f(n, a, xe), A0e = e_A0(n, a, xe);
    ph[v] = fe * (1 - sqrt(A0e / x[10 + 3 * v]));
    pf[v] = fe * (1 - sqrt(A0e / x[9 + 3 * v]));
  }
  y[6] = x[0] - x[3] - x[6];
  y[7] = x[1] - x[4] - x[7];
  y[8] = ph[0] - ph[1]; y[9] = ph[0] - ph[2];
  y[10] = pf[0] - pf[1]; y[11] = pf[0] - pf[2];
}

static void jacobian(const orc_net* n, const jn_t* J, const double* x, double* M /* [18][18] */) {
  memset(M, 0, 324 * sizeof(double));
  double rpi = sqrt(M_PI);
  double r_d1 = n->dt / J->dex[1];
  for (int v = 0; v < 3; ++v) {
    int a = J->a[v];
    double s = v == 0 ? 1.0 : -1.0, xe, xi, xg, xh, xj;
    jn_points(n, J, v, &xe, &xi, &xg, &xh, &xj);
    int iq = 3 * v, iqh = iq + 1, iqg = iq + 2, iA = 9 + 3 * v, iAh = iA + 1, iAg = iA + 2;
    M[v * 18 + iqh] = 2; M[v * 18 + iqg] = -1;
    M[(3 + v) * 18 + iAh] = 2; M[(3 + v) * 18 + iAg] = -1;
    double fe = e_f(n, a, xe), A0e = e_A0(n, a, xe);
    double dPh = s * fe * sqrt(A0e) / 2 / pow(x[iAh], 1.5), dPf = s * fe * sqrt(A0e) / 2 / pow(x[iA], 1.5);
    if (v == 0) {
      M[8 * 18 + iAh] = dPh; M[9 * 18 + iAh] = dPh; M[10 * 18 + iA] = dPf; M[11 * 18 + iA] = dPf;
    } else {
      M[(7 + v) * 18 + iAh] = dPh; M[(9 + v) * 18 + iA] = dPf;
    }
    double fh = e_f(n, a, xj), A0h = e_A0(n, a, xj), df = e_dfdr(n, a, xj), dr = e_drdx(n, a, xj);
    double r = n->dt / J->dex[v], qg = x[iqg], Ag = x[iAg];
    M[(12 + v) * 18 + iq] = 1;
    M[(12 + v) * 18 + iqg] = s * r * 2 * qg / Ag + n->dt * rpi / n->db / n->Re / sqrt(Ag);
    M[(12 + v) * 18 + iAg] = s * r * (-(qg / Ag) * (qg / Ag) + fh / 2 * sqrt(A0h / Ag)) -
                             n->dt / 2 * (rpi / n->db / n->Re * qg / pow(Ag, 1.5) +
                                          (1 / sqrt(Ag) * (rpi * fh + sqrt(A0h) * df) - df) * dr);
    M[(15 + v) * 18 + iqg] = s * (v < 2 ? r : r_d1); /* an:662 */
    M[(15 + v) * 18 + iA] = 1;
  }
  M[6 * 18 + 0] = 1; M[6 * 18 + 3] = -1; M[6 * 18 + 6] = -1;
  M[7 * 18 + 1] = 1; M[7 * 18 + 4] = -1; M[7 * 18 + 7] = -1;
}

/* dgesv-like dense solve with partial pivoting; returns 1 on an exact zero pivot */
static int dense_solve(double* M, double* b, int n) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(M[i * n + k]) > fabs(M[p * n + k])) p = i;
    if (M[p * n + k] == 0.0) return 1;
    if (p != k) {
      for (int j = 0; j < n; ++j) { double t = M[k * n + j]; M[k * n + j] = M[p * n + j]; M[p * n + j] = t; }
      double t = b[k]; b[k] = b[p]; b[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      double l = M[i * n + k] / M[k * n + k];
      if (l == 0.0) continue;
      for (int j = k + 1; j < n; ++j) M[i * n + j] -= l * M[k * n + j];
      b[i] -= l * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= M[i * n + j] * b[j];
    b[i] = s / M[i * n + i];
  }
  return 0;
}

static int newton(const orc_net* n, const jn_t* J, double* x) { /* an:668-710 */
  double M[324], y[18];
  int solves = 0;
  for (int k = 0; k < n->jk_max; ++k) {
    jacobian(n, J, x, M);
    problem_function(n, J, x, y);
    double nrm = 0;
    for (int i = 0; i < 18; ++i) nrm += y[i] * y[i];
    if (sqrt(nrm) < n->jtol) break;
    if (dense_solve(M, y, 18)) {
      jacobian(n, J, x, M);
      problem_function(n, J, x, y);
      for (int i = 0; i < 18; ++i) M[i * 18 + i] += 1e-6;
      y[0] += 1e-6;
      if (dense_solve(M, y, 18)) break;
    }
    for (int i = 0; i < 18; ++i) x[i] -= y[i];
    ++solves;
  }
  return solves;
}

static void set_inner_bc(orc_net* n, int j) { /* an:713-764 */
  jn_t J;
  J.a[0] = n->jp[j]; J.a[1] = n->j1[j]; J.a[2] = n->j2[j];
  double M = 1e300;
  for (int v = 0; v < 3; ++v) {
    double xe = v == 0 ? n->L[J.a[v]] : 0.0, U[2];
    un_at(n, J.a[v], xe, U);
    double m = cfl_term(n, J.a[v], xe, U[0], U[1]);
    if (m < M) M = m;
  }
  double dex = (1 + n->margin) * n->dt / M;
  for (int v = 0; v < 3; ++v) J.dex[v] = dex;
  double* x = n->x + (size_t)j * 18;
  n->j_solves[j] = newton(n, &J, x);
  n->dex[J.a[0]] = dex; /* value the parent holds after set_bcs */
  n->bc[J.a[0] * 4 + 2] = x[9];  n->bc[J.a[0] * 4 + 3] = x[0];
  n->bc[J.a[1] * 4 + 0] = x[12]; n->bc[J.a[1] * 4 + 1] = x[3];
  n->bc[J.a[2] * 4 + 0] = x[15]; n->bc[J.a[2] * 4 + 1] = x[6];
}

/* ---- per-artery weak form (ar:191-230) + [3P] assembly and Newton ----------------- */
static double F2(double A, double q, double f, double A0) { return pow(q, 2) / A + f * sqrt(A0 * A); }
static double S2(const orc_net* n, double A, double q, double f, double A0, double df, double dr) {
  double rpi = sqrt(M_PI);
  return -2 * rpi / n->db / n->Re * q / sqrt(A) + (2 * sqrt(A) * (rpi * f + sqrt(A0) * df) - A * df) * dr;
}

#define KL 3
#define KU 3
#define BW (2 * KL + KU + 1)
#define MB(i, j) Mb[(size_t)(i)*BW + ((j) - (i) + KL)]

/* residual R[2*np] (vertex-major: A_j, q_j) and, if Mb, the banded Jacobian */
static void assemble(const orc_net* n, int a, const double* UA, const double* Uq, double* R, double* Mb) {
  int np = n->np, ne = n->Nx;
  double h = n->dx[a], dt = n->dt, th = n->theta, rpi = sqrt(M_PI), Kr = 2 * rpi / n->db / n->Re;
  const double* An = n->A + (size_t)a * np;
  const double* qn = n->q + (size_t)a * np;
  memset(R, 0, 2 * np * sizeof(double));
  if (Mb) memset(Mb, 0, (size_t)2 * np * BW * sizeof(double));
  for (int e = 0; e < ne; ++e) {
    double qx = (Uq[e + 1] - Uq[e]) / h, qnx = (qn[e + 1] - qn[e]) / h;
    for (int g = 0; g < 3; ++g) {
      size_t ci = ((size_t)a * ne + e) * 3 + g;
      double f = n->cf[ci], A0 = n->cA0[ci], df = n->cdf[ci], dr = n->cdr[ci];
      double phi[2] = {1.0 - GXI[g], GXI[g]}, dphi[2] = {-1.0 / h, 1.0 / h}, w = GW[g] * h;
      double A = UA[e] * phi[0] + UA[e + 1] * phi[1], Ae = A + DOLFIN_EPS;
      double q = Uq[e] * phi[0] + Uq[e + 1] * phi[1];
      double Aold = An[e] * phi[0] + An[e + 1] * phi[1], qold = qn[e] * phi[0] + qn[e + 1] * phi[1];
      double Fm = th * F2(Ae, q, f, A0) + (1 - th) * F2(Aold, qold, f, A0);
      double Sm = th * S2(n, Ae, q, f, A0, df, dr) + (1 - th) * S2(n, Aold, qold, f, A0, df, dr);
      double dFdA = 0, dFdq = 0, dSdA = 0, dSdq = 0;
      if (Mb) {
        dFdA = -pow(q, 2) / pow(Ae, 2) + f * A0 / (2 * sqrt(A0 * Ae));
        dFdq = 2 * q / Ae;
        dSdA = Kr / 2 * q / (Ae * sqrt(Ae)) + ((rpi * f + sqrt(A0) * df) / sqrt(Ae) - df) * dr;
        dSdq = -Kr / sqrt(Ae);
      }
      for (int i = 0; i < 2; ++i) {
        int ri = 2 * (e + i);
        R[ri] += w * ((A - Aold) * phi[i] + dt * (th * qx + (1 - th) * qnx) * phi[i]);
        R[ri + 1] += w * ((q - qold) * phi[i] - dt * Fm * dphi[i] - dt * Sm * phi[i]);
        if (!Mb) continue;
        for (int j = 0; j < 2; ++j) {
          int cj = 2 * (e + j);
          MB(ri, cj) += w * phi[i] * phi[j];
          MB(ri, cj + 1) += w * dt * th * dphi[j] * phi[i];
          MB(ri + 1, cj) += -w * dt * th * (dFdA * phi[j] * dphi[i] + dSdA * phi[j] * phi[i]);
          MB(ri + 1, cj + 1) += w * (phi[i] * phi[j] - dt * th * (dFdq * phi[j] * dphi[i] + dSdq * phi[j] * phi[i]));
        }
      }
    }
  }
  /* "+F2*v2*ds" at both ends (ar:197-198, 206-207) */
  for (int k = 0; k < 2; ++k) {
    int v = k == 0 ? 0 : ne;
    double x = k == 0 ? 0.0 : h * ne;
    double f = e_f(n, a, x), A0 = e_A0(n, a, x), Ae = UA[v] + DOLFIN_EPS, q = Uq[v];
    R[2 * v + 1] += dt * (th * F2(Ae, q, f, A0) + (1 - th) * F2(An[v], qn[v], f, A0));
    if (Mb) {
      MB(2 * v + 1, 2 * v) += dt * th * (-pow(q, 2) / pow(Ae, 2) + f * A0 / (2 * sqrt(A0 * Ae)));
      MB(2 * v + 1, 2 * v + 1) += dt * th * 2 * q / Ae;
    }
  }
  /* [3P] DirichletBC rows (ar:171-189) */
  int root = (a == 0), leaf = n->art_leaf[a] >= 0;
  int rows[4], nr = 0;
  double gv[4];
  if (!root) { rows[nr] = 0; gv[nr++] = n->bc[a * 4 + 0]; }
  rows[nr] = 1; gv[nr++] = n->bc[a * 4 + 1];
  rows[nr] = 2 * ne; gv[nr++] = n->bc[a * 4 + 2];
  if (!leaf) { rows[nr] = 2 * ne + 1; gv[nr++] = n->bc[a * 4 + 3]; }
  for (int k = 0; k < nr; ++k) {
    int r = rows[k];
    R[r] = ((r & 1) ? Uq[r / 2] : UA[r / 2]) - gv[k];
    if (Mb) {
      for (int c = 0; c < BW; ++c) Mb[(size_t)r * BW + c] = 0.0;
      MB(r, r) = 1.0;
    }
  }
}

/* banded LU with partial pivoting (LAPACK dgbsv recipe), b <- solution */
static void banded_solve(double* Mb, double* b, int m) {
  for (int k = 0; k < m; ++k) {
    int last = k + KL < m - 1 ? k + KL : m - 1;
    int p = k;
    for (int i = k + 1; i <= last; ++i)
      if (fabs(MB(i, k)) > fabs(MB(p, k))) p = i;
    int jmax = k + KU + KL < m - 1 ? k + KU + KL : m - 1;
    if (p != k) {
      for (int j = k; j <= jmax; ++j) { double t = MB(k, j); MB(k, j) = MB(p, j); MB(p, j) = t; }
      double t = b[k]; b[k] = b[p]; b[p] = t;
    }
    double piv = MB(k, k);
    for (int i = k + 1; i <= last; ++i) {
      double l = MB(i, k) / piv;
      if (l == 0.0) continue;
      for (int j = k + 1; j <= jmax; ++j) MB(i, j) -= l * MB(k, j);
      b[i] -= l * b[k];
    }
  }
  for (int i = m - 1; i >= 0; --i) {
    int jmax = i + KU + KL < m - 1 ? i + KU + KL : m - 1;
    double s = b[i];
    for (int j = i + 1; j <= jmax; ++j) s -= MB(i, j) * b[j];
    b[i] = s / MB(i, i);
  }
}

static double norm2(const double* v, int m) {
  double s = 0;
  for (int i = 0; i < m; ++i) s += v[i] * v[i];
  return sqrt(s);
}

/* Artery.solve (ar:233-244) + [3P] NewtonSolver; then update_solution.  Work arrays by caller. */
static void artery_solve(orc_net* n, int a, double* R, double* Mb) {
  int np = n->np;
  double* UA = n->UA + (size_t)a * np;
  double* Uq = n->Uq + (size_t)a * np;
  if (a == 0) { /* root's q_in already in bc */ }
  assemble(n, a, UA, Uq, R, NULL);
  double r0 = norm2(R, 2 * np), r = r0;
  int it = 0;
  while (!(r / r0 < n->rtol || r < n->atol)) {
    if (it >= n->max_it || !(r == r)) { n->failed = 1; break; }
    assemble(n, a, UA, Uq, R, Mb);
    banded_solve(Mb, R, 2 * np);
    for (int i = 0; i < np; ++i) { UA[i] -= R[2 * i]; Uq[i] -= R[2 * i + 1]; }
    ++it;
    assemble(n, a, UA, Uq, R, NULL);
    r = norm2(R, 2 * np);
  }
  n->art_its[a] = it;
}

/* ---- public API (ctypes) ------------------------------------------------------------- */
orc_net* orc_create(int na, int nj, int nl, int Nx, int Nt, double dt, double theta, const int* jp, const int* j1,
                    const int* j2, const int* leaf, double k1, double k2, double k3, double rho, double Re, double db,
                    double p0, const double* Ru, const double* Rd, const double* L, const double* q0,
                    const double* R1, const double* R2, const double* CT, const double* q_ins) {
  orc_net* n = (orc_net*)calloc(1, sizeof(orc_net));
  n->na = na; n->nj = nj; n->nl = nl; n->Nx = Nx; n->np = Nx + 1; n->Nt = Nt;
  n->dt = dt; n->theta = theta; n->k1 = k1; n->k2 = k2; n->k3 = k3; n->rho = rho; n->Re = Re; n->db = db; n->p0 = p0;
  n->rtol = 1e-9; n->atol = 1e-10; n->max_it = 1000; n->jk_max = 30; n->jtol = 1e-10; n->pk_max = 100;
  n->ptol = 1e-12; n->margin = 0.05;
#define DUP(dst, src, cnt, T) do { n->dst = (T*)malloc(((cnt) ? (cnt) : 1) * sizeof(T)); memcpy(n->dst, src, (cnt) * sizeof(T)); } while (0)
  DUP(jp, jp, nj, int); DUP(j1, j1, nj, int); DUP(j2, j2, nj, int); DUP(leaf, leaf, nl, int);
  DUP(Ru, Ru, na, double); DUP(Rd, Rd, na, double); DUP(L, L, na, double); DUP(q0, q0, na, double);
  DUP(R1, R1, nl, double); DUP(R2, R2, nl, double); DUP(CT, CT, nl, double); DUP(q_ins, q_ins, Nt, double);
  n->art_leaf = (int*)malloc(na * sizeof(int));
  for (int a = 0; a < na; ++a) n->art_leaf[a] = -1;
  for (int l = 0; l < nl; ++l) n->art_leaf[leaf[l]] = l;
  int np = n->np;
  n->dx = (double*)malloc(na * sizeof(double));
  n->A = (double*)malloc((size_t)na * np * sizeof(double)); n->q = (double*)malloc((size_t)na * np * sizeof(double));
  n->UA = (double*)malloc((size_t)na * np * sizeof(double)); n->Uq = (double*)malloc((size_t)na * np * sizeof(double));
  n->bc = (double*)malloc(na * 4 * sizeof(double)); n->dex = (double*)malloc(na * sizeof(double));
  n->x = (double*)malloc(((size_t)nj * 18 + 1) * sizeof(double));
  size_t nc = (size_t)na * Nx * 3;
  n->cf = (double*)malloc(nc * sizeof(double)); n->cA0 = (double*)malloc(nc * sizeof(double));
  n->cdf = (double*)malloc(nc * sizeof(double)); n->cdr = (double*)malloc(nc * sizeof(double));
  n->art_its = (int*)calloc(na, sizeof(int)); n->j_solves = (int*)calloc(nj + 1, sizeof(int));
  n->p_its = (int*)calloc(nl + 1, sizeof(int));
  for (int a = 0; a < na; ++a) {
    double h = n->dx[a] = L[a] / Nx;
    n->dex[a] = h;
    for (int j = 0; j < np; ++j) { /* Un = (A0(x_j), q0)  ar:154-158 */
      n->A[(size_t)a * np + j] = n->UA[(size_t)a * np + j] = e_A0(n, a, h * j);
      n->q[(size_t)a * np + j] = n->Uq[(size_t)a * np + j] = q0[a];
    }
    n->bc[a * 4 + 0] = e_A0(n, a, 0); n->bc[a * 4 + 1] = q0[a];
    n->bc[a * 4 + 2] = e_A0(n, a, L[a]); n->bc[a * 4 + 3] = q0[a];
    for (int e = 0; e < Nx; ++e) { /* [3P] P2 interpolants at the Gauss points */
      double xl = h * e, xr = h * (e + 1), xm = 0.5 * (xl + xr);
      double (*fn[4])(const orc_net*, int, double) = {e_f, e_A0, e_dfdr, e_drdx};
      double* dst[4] = {n->cf, n->cA0, n->cdf, n->cdr};
      for (int c = 0; c < 4; ++c) {
        double vl = fn[c](n, a, xl), vr = fn[c](n, a, xr), vm = fn[c](n, a, xm);
        for (int g = 0; g < 3; ++g) {
          double xi = GXI[g];
          dst[c][((size_t)a * Nx + e) * 3 + g] =
              vl * ((1 - xi) * (1 - 2 * xi)) + vr * (xi * (2 * xi - 1)) + vm * (4 * xi * (1 - xi));
        }
      }
    }
  }
  for (int j = 0; j < nj; ++j) { /* initial_x  an:434-461 */
    int ia[3] = {jp[j], j1[j], j2[j]};
    for (int v = 0; v < 3; ++v)
      for (int k = 0; k < 3; ++k) {
        n->x[j * 18 + 3 * v + k] = q0[ia[v]];
        n->x[j * 18 + 9 + 3 * v + k] = e_A0(n, ia[v], v == 0 ? L[ia[v]] : 0.0);
      }
  }
  return n;
}

void orc_destroy(orc_net* n) {
  if (!n) return;
  free(n->jp); free(n->j1); free(n->j2); free(n->leaf); free(n->art_leaf);
  free(n->Ru); free(n->Rd); free(n->L); free(n->dx); free(n->q0); free(n->R1); free(n->R2); free(n->CT);
  free(n->q_ins); free(n->A); free(n->q); free(n->UA); free(n->Uq); free(n->bc); free(n->dex); free(n->x);
  free(n->cf); free(n->cA0); free(n->cdf); free(n->cdr); free(n->art_its); free(n->j_solves); free(n->p_its);
  free(n);
}

/* loop body an:916-944 for global steps [g0, g0+count); returns the failed flag.
 * log_art/log_j (may be NULL): per-step counters [count][na] / [count][nj]. */
int orc_step(orc_net* n, long long g0, long long count, int nthreads, int* log_art, int* log_j) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int np = n->np;
#pragma omp parallel
  {
    double* R = (double*)malloc(2 * np * sizeof(double));
    double* Mb = (double*)malloc((size_t)2 * np * BW * sizeof(double));
    for (long long g = g0; g < g0 + count; ++g) {
      int step = (int)(g % n->Nt);
#pragma omp single
      n->bc[1] = n->q_ins[(step + 1) % n->Nt]; /* an:777, 916 */
#pragma omp for schedule(dynamic, 1)
      for (int t = 0; t < n->nj + n->nl; ++t) {
        if (t < n->nj) set_inner_bc(n, t);
        else n->bc[n->leaf[t - n->nj] * 4 + 2] = windkessel(n, t - n->nj);
      }
#pragma omp for schedule(dynamic, 1)
      for (int a = 0; a < n->na; ++a) artery_solve(n, a, R, Mb);
#pragma omp single
      {
        memcpy(n->A, n->UA, (size_t)n->na * np * sizeof(double)); /* update_solution  ar:247-251 */
        memcpy(n->q, n->Uq, (size_t)n->na * np * sizeof(double));
        for (int a = 0; a < n->na; ++a) n->tot[0] += n->art_its[a];
        for (int j = 0; j < n->nj; ++j) n->tot[1] += n->j_solves[j];
        for (int l = 0; l < n->nl; ++l) n->tot[2] += n->p_its[l];
        if (log_art) memcpy(log_art + (g - g0) * n->na, n->art_its, n->na * sizeof(int));
        if (log_j) memcpy(log_j + (g - g0) * n->nj, n->j_solves, n->nj * sizeof(int));
      }
    }
    free(R); free(Mb);
  }
  return n->failed;
}

void orc_get_state(const orc_net* n, double* A, double* q) {
  memcpy(A, n->A, (size_t)n->na * n->np * sizeof(double));
  memcpy(q, n->q, (size_t)n->na * n->np * sizeof(double));
}
void orc_set_state(orc_net* n, const double* A, const double* q) {
  size_t b = (size_t)n->na * n->np * sizeof(double);
  memcpy(n->A, A, b); memcpy(n->q, q, b); memcpy(n->UA, A, b); memcpy(n->Uq, q, b);
}
void orc_set_bc(orc_net* n, const double* bc) { memcpy(n->bc, bc, (size_t)n->na * 4 * sizeof(double)); }
/* the pending iterate U of every artery (== Un between steps) */
void orc_get_pending(const orc_net* n, double* A, double* q) {
  memcpy(A, n->UA, (size_t)n->na * n->np * sizeof(double));
  memcpy(q, n->Uq, (size_t)n->na * n->np * sizeof(double));
}
/* residual R[2*np] and Jacobian as 2x2 blocks J[3][np][2][2] (lower, diagonal, upper) of artery a
 * at the iterate (UA, Uq) with the current Un and BC holders, Dirichlet rows applied */
void orc_artery_assemble(const orc_net* n, int a, const double* UA, const double* Uq, double* R, double* J) {
  int np = n->np;
  double* Mb = (double*)malloc((size_t)2 * np * BW * sizeof(double));
  assemble(n, a, UA, Uq, R, Mb);
  memset(J, 0, (size_t)3 * np * 4 * sizeof(double));
  for (int j = 0; j < np; ++j)
    for (int b = 0; b < 3; ++b) {
      int jc = j + b - 1;
      if (jc < 0 || jc >= np) continue;
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) J[(((size_t)b * np + j) * 2 + r) * 2 + c] = MB(2 * j + r, 2 * jc + c);
    }
  free(Mb);
}
/* Artery.solve() of one artery with at most max_it linear solves (no update_solution); returns the count */
int orc_artery_solve(orc_net* n, int a, int max_it) {
  int np = n->np, keep = n->max_it, failed = n->failed;
  double* R = (double*)malloc(2 * np * sizeof(double));
  double* Mb = (double*)malloc((size_t)2 * np * BW * sizeof(double));
  n->max_it = max_it;
  artery_solve(n, a, R, Mb);
  n->max_it = keep; n->failed = failed;
  free(R); free(Mb);
  return n->art_its[a];
}
void orc_get_x(const orc_net* n, double* x) { memcpy(x, n->x, (size_t)n->nj * 18 * sizeof(double)); }
void orc_get_bc(const orc_net* n, double* bc) { memcpy(bc, n->bc, (size_t)n->na * 4 * sizeof(double)); }
void orc_get_totals(const orc_net* n, long long* t) { memcpy(t, n->tot, 3 * sizeof(long long)); }
void orc_get_pressure(const orc_net* n, double* p) { /* ar:254-261 */
  for (int a = 0; a < n->na; ++a)
    for (int j = 0; j < n->np; ++j) {
      double x = n->dx[a] * j;
      p[(size_t)a * n->np + j] = n->p0 + e_f(n, a, x) * (1 - sqrt(e_A0(n, a, x) / n->A[(size_t)a * n->np + j]));
    }
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
