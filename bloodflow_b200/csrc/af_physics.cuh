// af_physics.cuh -- per-thread FP64 math of the artery.fe network solve.
//
// Everything here is a small inline function operating on registers / caller
// supplied arrays; the kernels in af_kernels.cu decide which thread runs which
// piece.  References are to /root/reference/arteryfe (an = artery_network.py,
// ar = artery.py); [3P] marks FEniCS 2017.2.0 semantics.
//
// The functions are AF_HD so that tests/host_harness can compile this header
// with g++ and unit-test the formulas in the CPU-only build container.  The
// product library (libarteryfe_b200.so) only ever executes them on the device.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define AF_HD __host__ __device__ __forceinline__
#else
#define AF_HD inline
#endif

namespace af {

// Branch-free reciprocal and reciprocal square root: the hardware seed
// (MUFU.RCP64H / MUFU.RSQ64H, relative error ~2^-23) refined by ONE cubically
// convergent step in straight-line FMAs (error after the step ~2^-69, i.e. the
// result is the rounding of the last FMA: < 1 ulp, checked against 1/x and
// 1/sqrt(x) by tests/test_gpu_api.py through af_eval_math).  The library versions
// guard every call with a slow-path branch for denormal / huge arguments, which cuts
// the assembly of a cell into a dozen small basic blocks and lengthens the serial
// chain of the Windkessel iteration; areas and 2x2 block determinants are O(1)
// numbers, and 0 / negative / non-finite arguments still come out as inf / NaN and
// are caught by the non-finite tests.  Host builds (tests/host_harness) use the plain
// expressions.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double af_rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);   // 1 - x y
  e = fma(e, e, e);             // e + e^2
  return fma(y, e, y);          // y (1 + e + e^2) = 1/x (1 - e^3)
}
__device__ __forceinline__ double af_rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = x * y;
  double e = fma(-t, y, 1.0);                   // 1 - x y^2
  double p = fma(0.375, e, 0.5) * e;            // e/2 + 3 e^2/8
  return fma(y, p, y);                          // y (1 - e)^-1/2 up to (5/16) e^3
}
#else
inline double af_rcp_fast(double x) { return 1.0 / x; }
inline double af_rsqrt_fast(double x) { return 1.0 / sqrt(x); }
#endif
// a / b and sqrt(x > 0) of the boundary-condition code (the serial chain of the lone
// warp that owns a junction): through the branch-free functions on the device (<= 2 ulp),
// the IEEE operations on the host.
#ifdef __CUDA_ARCH__
#define AF_DIV(a, b) ((a) * af_rcp_fast(b))
__device__ __forceinline__ double af_sqrt_pos(double x) { return x * af_rsqrt_fast(x); }
#else
#define AF_DIV(a, b) ((a) / (b))
inline double af_sqrt_pos(double x) { return sqrt(x); }
#endif

constexpr double kPi = 3.14159265358979323846;
constexpr double kSqrtPi = 1.7724538509055159;   // == (double)sqrt((double)pi)
constexpr double kDolfinEps = 3.0e-16;           // [3P] DOLFIN_EPS
// [3P] 3-point Gauss-Legendre on [0,1] (quadrature_degree = 4, ar:9)
constexpr double kXi0 = 0.11270166537925831148;  // 1/2 - sqrt(3/5)/2
constexpr double kXi1 = 0.5;
constexpr double kXi2 = 0.88729833462074168852;
constexpr double kW0 = 5.0 / 18.0, kW1 = 8.0 / 18.0, kW2 = 5.0 / 18.0;

}  // namespace af
// status bits (per network); same values as include/arteryfe_b200.h
#ifndef AF_FLAG_ARTERY_NOCONV
#define AF_FLAG_ARTERY_NOCONV 1u
#define AF_FLAG_JUNCTION_KMAX 2u
#define AF_FLAG_PICARD_KMAX 4u
#define AF_FLAG_NONFINITE 8u
#define AF_FLAG_SINGULAR 16u
#define AF_FLAG_PICARD_CYCLE 32u
#endif
namespace af {

// ---------------------------------------------------------------------------
// vessel constants and the analytic Expressions  (ar:110-118)
// ---------------------------------------------------------------------------
struct Vessel {
  double Ru, Rd, L, dx;
  double k1, k2, k3;
  double lr_over_L;     // log(Rd/Ru)/L                      (ar:117-118)
  double rho;           // dimensional, used by CFL_term     (ar:322)
  double p0;
  double src_k;         // -2*sqrt(pi)/db/Re                 (an:322)
  double rpi_dbRe;      // sqrt(pi)/db/Re                    (an:639)
  int32_t tapered;      // Rd != Ru
  // Expression values cached for untapered vessels and for the two ends
  double f0, A00, fL, A0L;
  double dfdr0;         // dfdr at x = 0 (the constant value of an untapered vessel)
};

struct Coef {
  double f, A0, dfdr, drdx;
};

// Exact evaluation (used at set-up and for tapered vessels).
AF_HD Coef coef_exact(const Vessel& v, double x) {
  Coef c;
  // r0 = Ru*pow(Rd/Ru, x/L); pow(1, y) == 1 exactly, so the untapered branch
  // only skips work.
  double r0 = v.tapered ? v.Ru * pow(v.Rd / v.Ru, x / v.L) : v.Ru;
  double e = exp(v.k2 * r0);
  c.A0 = kPi * (r0 * r0);
  c.f = 4.0 / 3.0 * (v.k1 * e + v.k3);
  c.dfdr = 4.0 / 3.0 * v.k1 * v.k2 * e;
  c.drdx = v.lr_over_L * r0;
  return c;
}

// Hot-path evaluation: an untapered vessel (Rd == Ru) has x-independent
// Expressions (drdx == log(1)/L*r0 == 0 exactly), cached at set-up.
AF_HD Coef coef_at(const Vessel& v, double x) {
  if (v.tapered) return coef_exact(v, x);
  Coef c;
  c.f = v.f0; c.A0 = v.A00; c.dfdr = v.dfdr0; c.drdx = 0.0;
  return c;
}

// P1 point evaluation of (A, q) on the uniform mesh x_j = j*dx  ([3P] a15)
AF_HD void interp_state(const double* A, const double* q, int Nx, double dx, double x,
                        double& Ax, double& qx) {
  const double inv_dx = af_rcp_fast(dx);
  int j = (int)floor(x * inv_dx);
  j = j < 0 ? 0 : (j > Nx - 1 ? Nx - 1 : j);
  double xi = (x - j * dx) * inv_dx;
  Ax = A[j] * (1.0 - xi) + A[j + 1] * xi;
  qx = q[j] * (1.0 - xi) + q[j + 1] * xi;
}

// ---------------------------------------------------------------------------
// boundary-condition physics  (an:282-358)
// ---------------------------------------------------------------------------
// flux: F = (q, q^2 + f*sqrt(A0*A))   -- no "/A" (an:300)
AF_HD double bc_flux2(const Coef& c, double A, double q) { return q * q + c.f * af_sqrt_pos(c.A0 * A); }

// source: S2 (an:322-325)
AF_HD double bc_source2(const Vessel& v, const Coef& c, double A, double q) {
  double isA = af_rsqrt_fast(A), sA = A * isA;
  return v.src_k * q * isA + (2.0 * sA * (kSqrtPi * c.f + af_sqrt_pos(c.A0) * c.dfdr) - A * c.dfdr) * c.drdx;
}

// compute_U_half (an:329-358): note dt/(x1-x0) and dt/4
AF_HD void u_half(const Vessel& v, double dt, double x0, double x1, double A0s, double q0s, double A1s,
                  double q1s, double& Ah, double& qh) {
  Coef c0 = coef_at(v, x0), c1 = coef_at(v, x1);
  double F0 = bc_flux2(c0, A0s, q0s), F1 = bc_flux2(c1, A1s, q1s);
  double S0 = bc_source2(v, c0, A0s, q0s), S1 = bc_source2(v, c1, A1s, q1s);
  double r = AF_DIV(dt, x1 - x0);
  Ah = (A0s + A1s) / 2.0 - r * (q1s - q0s);
  qh = (q0s + q1s) / 2.0 - r * (F1 - F0) + dt / 4.0 * (S0 + S1);
}

// CFL_term (ar:304-325)
AF_HD double cfl_term(const Vessel& v, const Coef& c, double A, double q) {
  const double iA = af_rcp_fast(A);
  double w = af_sqrt_pos(AF_DIV(c.f / 2.0, v.rho) * af_sqrt_pos(c.A0 * iA));
  double u = q * iA;
  double a = fabs(u + w), b = fabs(u - w);
  return af_rcp_fast(a > b ? a : b);
}

AF_HD double outlet_pressure(const Vessel& v, double A) {   // ar:285-301
  return v.p0 + v.fL * (1.0 - af_sqrt_pos(v.A0L) * af_rsqrt_fast(A));
}

// ---------------------------------------------------------------------------
// Windkessel outlet (an:361-422)
// ---------------------------------------------------------------------------
struct WkResult {
  double A_out, dex;
  int its;
  int cycle;     // 1: the iteration fell into a round-off limit cycle and was cut short exactly
};

// The Picard iteration of an:405-422 given the outlet state (A0s, q0s) = Un(L),
// the Lax-Wendroff value qm1 at L-dex and the step dex.
AF_HD WkResult windkessel_picard(const Vessel& v, double dt, double dex, double A0s, double q0s, double qm1,
                                 double R1, double R2, double CT, int k_max, double tol) {
  WkResult r;
  r.cycle = 0;
  double pn = outlet_pressure(v, A0s);
  // Fixed-point map of an:411-420,
  //   qm0 = q + (p-pn)/R1 + dt/(R1 R2 C) pn - dt (R1+R2)/(R1 R2 C) q,
  //   Am0 = A - dt/dex (qm0 - qm1),   p = p0 + f(L) (1 - sqrt(A0(L)/Am0)),
  // with the p-independent terms folded:  Am0 = K2 - K3 p,  p = K4 - K5 rsqrt(Am0).
  // Two FMAs and one reciprocal square root per iteration on the serial chain;
  // rounding differs from the reference's expression order by a few ulp.
  // the three reciprocals are independent (the reference divides in sequence)
  const double i_rrc = af_rcp_fast(R1 * R2 * CT), i_r1 = af_rcp_fast(R1);
  const double c_pn = dt * i_rrc * pn;
  const double c_q = dt * (R1 + R2) * i_rrc * q0s;
  const double dtdex = AF_DIV(dt, dex);
  const double K3 = dtdex * i_r1;
  const double K2 = A0s - dtdex * (q0s + c_pn - c_q - qm1) + K3 * pn;
  const double K5 = v.fL * af_sqrt_pos(v.A0L);
  const double K4 = v.p0 + v.fL;
  double p = pn, p1 = pn, p2 = pn, p3 = pn;   // p_k, p_{k-1}, p_{k-2}, p_{k-3}
  double Am0 = A0s, Am1 = A0s, Am2 = A0s;     // Am_k, Am_{k-1}, Am_{k-2}
  int k = 0;
  while (k < k_max) {
    p3 = p2; p2 = p1; p1 = p;
    Am2 = Am1; Am1 = Am0;
    Am0 = K2 - K3 * p1;
    p = K4 - K5 * af_rsqrt_fast(Am0);
    ++k;
    if (fabs(p - p1) < tol) break;
    // The map is deterministic: once an iterate repeats (round-off limit
    // cycle of period 2 or 3, 10-40 % of the calls) the remaining iterations
    // are known exactly, so jump to what iteration k_max would return.
    int m = (k >= 2 && p == p2) ? 2 : ((k >= 3 && p == p3) ? 3 : 0);
    if (m) {
      int j = k_max - k;
      if (j > 0) {
        int jj = (j - 1) % m + 1;              // Am_{k+j} == Am_{k-m+jj}
        int back = m - jj;                     // 0: Am_k, 1: Am_{k-1}, 2: Am_{k-2}
        Am0 = back == 0 ? Am0 : (back == 1 ? Am1 : Am2);
      }
      k = k_max;
      r.cycle = 1;
      break;
    }
  }
  r.A_out = Am0;
  r.dex = dex;
  r.its = k;
  return r;
}

AF_HD WkResult windkessel(const Vessel& v, const double* A, const double* q, int Nx, double dt, double R1,
                          double R2, double CT, int k_max, double tol, double margin) {
  double L = v.L;
  Coef cL;
  cL.f = v.fL; cL.A0 = v.A0L;
  double AL, qL;
  interp_state(A, q, Nx, v.dx, L, AL, qL);
  double dex = AF_DIV((1.0 + margin) * dt, cfl_term(v, cL, AL, qL));       // adjust_dex (ar:349-365)
  double x2 = L - 2.0 * dex, x1 = L - dex, x0 = L;
  double x21 = L - 1.5 * dex, x10 = L - 0.5 * dex;
  double A2, q2, A1, q1, A0s, q0s;
  interp_state(A, q, Nx, v.dx, x2, A2, q2);
  interp_state(A, q, Nx, v.dx, x1, A1, q1);
  interp_state(A, q, Nx, v.dx, x0, A0s, q0s);
  double Ah21, qh21, Ah10, qh10;
  u_half(v, dt, x2, x1, A2, q2, A1, q1, Ah21, qh21);
  u_half(v, dt, x1, x0, A1, q1, A0s, q0s, Ah10, qh10);
  Coef c21 = coef_at(v, x21), c10 = coef_at(v, x10);
  double F21 = bc_flux2(c21, Ah21, qh21), S21 = bc_source2(v, c21, Ah21, qh21);
  double F10 = bc_flux2(c10, Ah10, qh10), S10 = bc_source2(v, c10, Ah10, qh10);
  double qm1 = q1 - AF_DIV(dt, dex) * (F10 - F21) + dt / 2.0 * (S10 + S21);
  return windkessel_picard(v, dt, dex, A0s, q0s, qm1, R1, R2, CT, k_max, tol);
}


// ---------------------------------------------------------------------------
// Bifurcation: 18 unknowns (an:476-710).  Unknown layout per vessel v
// (0 = parent outlet, 1/2 = daughter inlets):
//   x[3v] = q^{n+1}, x[3v+1] = q^{n+1/2}, x[3v+2] = q ghost^{n+1/2},
//   x[9+3v], x[10+3v], x[11+3v] = the same for A.
// ---------------------------------------------------------------------------
struct JunctionCtx {
  double dt;
  double dtdex[3];      // dt/dex of each vessel
  double Ae[3], qe[3];  // Un at the junction end
  double Ah[3], qh[3];  // interior half-step state
  double Fh2[3], Sh2[3];  // flux/source of it at the interior half point
  Coef cg[3];           // Expressions at the ghost half point  L+dex/2 | -dex/2   (an:503-508)
  Coef cj[3];           // Expressions at the Jacobian point    L+dex   | -dex     (an:589-597)
  double fe[3], A0e[3];   // Expressions at the junction end
  double src_k[3], rpi_dbRe[3];
};

// One vessel's share of the context.  `A`,`q` are that vessel's Un arrays.
AF_HD void junction_prepare_vessel(JunctionCtx& c, int v, const Vessel& ves, const double* A, const double* q,
                                   int Nx, double dt, double dex) {
  double L = ves.L;
  double xe = v == 0 ? L : 0.0;
  double xi = v == 0 ? L - dex : dex;
  double xg = v == 0 ? L + dex / 2.0 : -dex / 2.0;
  double xh = v == 0 ? L - dex / 2.0 : dex / 2.0;
  double xj = v == 0 ? L + dex : -dex;
  double Ae, qe, Ai, qi;
  interp_state(A, q, Nx, ves.dx, xe, Ae, qe);
  interp_state(A, q, Nx, ves.dx, xi, Ai, qi);
  double Ah, qh;
  if (v == 0)
    u_half(ves, dt, xi, xe, Ai, qi, Ae, qe, Ah, qh);
  else
    u_half(ves, dt, xe, xi, Ae, qe, Ai, qi, Ah, qh);
  Coef ch = coef_at(ves, xh);
  c.dt = dt;
  c.dtdex[v] = dt / dex;
  c.Ae[v] = Ae; c.qe[v] = qe;
  c.Ah[v] = Ah; c.qh[v] = qh;
  c.Fh2[v] = bc_flux2(ch, Ah, qh);
  c.Sh2[v] = bc_source2(ves, ch, Ah, qh);
  c.cg[v] = coef_at(ves, xg);
  c.cj[v] = coef_at(ves, xj);
  c.fe[v] = v == 0 ? ves.fL : ves.f0;
  c.A0e[v] = v == 0 ? ves.A0L : ves.A00;
  c.src_k[v] = ves.src_k;
  c.rpi_dbRe[v] = ves.rpi_dbRe;
}

// problem_function (an:476-562)
AF_HD void junction_residual(const JunctionCtx& c, const double* x, double* y) {
  double ph[3], pf[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    double s = v == 0 ? 1.0 : -1.0;
    double Ag = x[11 + 3 * v], qg = x[2 + 3 * v];
    // ghost flux/source (an:503-508); only the Expression coefficients are
    // needed from the Vessel, so the source is written out here
    double Fg2 = qg * qg + c.cg[v].f * sqrt(c.cg[v].A0 * Ag);
    double sA = sqrt(Ag);
    double Sg2 = c.src_k[v] * qg / sA +
                 (2.0 * sA * (kSqrtPi * c.cg[v].f + sqrt(c.cg[v].A0) * c.cg[v].dfdr) - Ag * c.cg[v].dfdr) *
                     c.cg[v].drdx;
    y[v] = 2.0 * x[3 * v + 1] - c.qh[v] - x[3 * v + 2];                       // eq (20)
    y[3 + v] = 2.0 * x[10 + 3 * v] - c.Ah[v] - x[11 + 3 * v];                 // eq (21)
    y[12 + v] = x[3 * v] - c.qe[v] + s * c.dtdex[v] * (Fg2 - c.Fh2[v]) - c.dt / 2.0 * (Sg2 + c.Sh2[v]);  // (26)
    y[15 + v] = x[9 + 3 * v] - c.Ae[v] + s * c.dtdex[v] * (qg - c.qh[v]);     // eq (27)
    ph[v] = c.fe[v] * (1.0 - sqrt(c.A0e[v] / x[10 + 3 * v]));
    pf[v] = c.fe[v] * (1.0 - sqrt(c.A0e[v] / x[9 + 3 * v]));
  }
  y[6] = x[0] - x[3] - x[6];                                                  // eq (22)
  y[7] = x[1] - x[4] - x[7];
  y[8] = ph[0] - ph[1];                                                       // eq (23), p0 omitted
  y[9] = ph[0] - ph[2];
  y[10] = pf[0] - pf[1];
  y[11] = pf[0] - pf[2];
}

// jacobian (an:565-665): hand-coded, 41 non-zeros, quirks kept.
// J is row-major with leading dimension ld.
AF_HD void junction_jacobian(const JunctionCtx& c, const double* x, double* J, int ld) {
  for (int i = 0; i < 18; ++i)
    for (int j = 0; j < 18; ++j) J[i * ld + j] = 0.0;
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    double s = v == 0 ? 1.0 : -1.0;
    int iq = 3 * v, iqh = 3 * v + 1, iqg = 3 * v + 2;
    int iA = 9 + 3 * v, iAh = 10 + 3 * v, iAg = 11 + 3 * v;
    J[v * ld + iqh] = 2.0;
    J[v * ld + iqg] = -1.0;
    J[(3 + v) * ld + iAh] = 2.0;
    J[(3 + v) * ld + iAg] = -1.0;
    double sg = v == 0 ? 1.0 : -1.0;
    double k = c.fe[v] * sqrt(c.A0e[v]) / 2.0;
    double dPh = sg * k / (x[iAh] * sqrt(x[iAh]));
    double dPf = sg * k / (x[iA] * sqrt(x[iA]));
    if (v == 0) {
      J[8 * ld + iAh] = dPh;
      J[9 * ld + iAh] = dPh;
      J[10 * ld + iA] = dPf;
      J[11 * ld + iA] = dPf;
    } else {
      J[(7 + v) * ld + iAh] = dPh;
      J[(9 + v) * ld + iA] = dPf;
    }
    double qg = x[iqg], Ag = x[iAg];
    double sA = sqrt(Ag);
    const Coef& cj = c.cj[v];
    J[(12 + v) * ld + iq] = 1.0;
    J[(12 + v) * ld + iqg] = s * c.dtdex[v] * 2.0 * qg / Ag + c.dt * c.rpi_dbRe[v] / sA;
    double u = qg / Ag;
    J[(12 + v) * ld + iAg] =
        s * c.dtdex[v] * (-(u * u) + cj.f / 2.0 * sqrt(cj.A0 / Ag)) -
        c.dt / 2.0 * (c.rpi_dbRe[v] * qg / (Ag * sA) +
                      (1.0 / sA * (kSqrtPi * cj.f + sqrt(cj.A0) * cj.dfdr) - cj.dfdr) * cj.drdx);
    // row 17 uses d1's dt/dex (an:662)
    J[(15 + v) * ld + iqg] = s * (v < 2 ? c.dtdex[v] : c.dtdex[1]);
    J[(15 + v) * ld + iA] = 1.0;
  }
  J[6 * ld + 0] = 1.0; J[6 * ld + 3] = -1.0; J[6 * ld + 6] = -1.0;
  J[7 * ld + 1] = 1.0; J[7 * ld + 4] = -1.0; J[7 * ld + 7] = -1.0;
}

// Dense LU with partial pivoting, the LAPACK dgesv recipe numpy.linalg.solve
// runs (an:702).  Solves J dx = b in place (b <- dx).  Returns 1 if an exact
// zero pivot is met (numpy raises LinAlgError there).
AF_HD int lu_solve18(double* J, int ld, double* b) {
  const int n = 18;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(J[k * ld + k]);
    for (int i = k + 1; i < n; ++i) {
      double a = fabs(J[i * ld + k]);
      if (a > best) { best = a; p = i; }
    }
    if (best == 0.0) return 1;
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        double t = J[k * ld + j]; J[k * ld + j] = J[p * ld + j]; J[p * ld + j] = t;
      }
      double t = b[k]; b[k] = b[p]; b[p] = t;
    }
    double inv = 1.0 / J[k * ld + k];
    for (int i = k + 1; i < n; ++i) {
      double l = J[i * ld + k] * inv;
      if (l != 0.0) {
        for (int j = k + 1; j < n; ++j) J[i * ld + j] -= l * J[k * ld + j];
        b[i] -= l * b[k];
      }
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= J[i * ld + j] * b[j];
    b[i] = s / J[i * ld + i];
  }
  return 0;
}

// 3x3 dense solve with partial pivoting, b <- solution; returns 1 on a zero pivot.
AF_HD int solve3(double M[9], double b[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
    double best = fabs(M[k * 3 + k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      double a = fabs(M[i * 3 + k]);
      if (a > best) { best = a; p = i; }
    }
    if (best == 0.0) return 1;
    if (p != k) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { double t = M[k * 3 + j]; M[k * 3 + j] = M[p * 3 + j]; M[p * 3 + j] = t; }
      double t = b[k]; b[k] = b[p]; b[p] = t;
    }
    double inv = 1.0 / M[k * 3 + k];
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      double l = M[i * 3 + k] * inv;
#pragma unroll
      for (int j = k + 1; j < 3; ++j) M[i * 3 + j] -= l * M[k * 3 + j];
      b[i] -= l * b[k];
    }
  }
  b[2] = b[2] / M[8];
  b[1] = (b[1] - M[5] * b[2]) / M[4];
  b[0] = (b[0] - M[1] * b[1] - M[2] * b[2]) / M[0];
  return 0;
}

// "Arrow" 3x3 system  [d0 d1 d2; m1 e1 0; m2 0 e2] x = b  (one dense row, two
// rows with two entries): eliminate x1, x2 through their own rows.  The two
// reciprocals are independent, so the serial chain is two divisions long.
// Returns 1 if a pivot vanishes.
AF_HD int solve_arrow3(double d0, double d1, double d2, double m1, double e1, double m2, double e2, double b[3]) {
  if (e1 == 0.0 || e2 == 0.0) return 1;
  const double i1 = af_rcp_fast(e1), i2 = af_rcp_fast(e2);
  const double den = d0 - d1 * (m1 * i1) - d2 * (m2 * i2);
  if (den == 0.0) return 1;
  const double x0 = (b[0] - d1 * (b[1] * i1) - d2 * (b[2] * i2)) * af_rcp_fast(den);
  b[1] = (b[1] - m1 * x0) * i1;
  b[2] = (b[2] - m2 * x0) * i2;
  b[0] = x0;
  return 0;
}

// One Newton iteration's data: residual y[18] and the 15 state-dependent
// Jacobian entries (the other 26 non-zeros are the constants 2, -1, +-1 and
// +-dt/dex).  Same formulas as junction_residual/junction_jacobian.
struct JunctionLin {
  double dPh[3], dPf[3];   // |d pressure / dA| at A^{n+1/2}, A^{n+1}   (rows 8-11)
  double gq[3], gA[3];     // J[12+v, qg_v], J[12+v, Ag_v]
};

AF_HD void junction_linearise(const JunctionCtx& c, const double* x, double* y, JunctionLin& l) {
  double ph[3], pf[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    double s = v == 0 ? 1.0 : -1.0;
    double Ag = x[11 + 3 * v], qg = x[2 + 3 * v];
    const Coef& cg = c.cg[v];
    const Coef& cj = c.cj[v];
    double sA = sqrt(Ag);
    double Fg2 = qg * qg + cg.f * sqrt(cg.A0 * Ag);
    double Sg2 = c.src_k[v] * qg / sA + (2.0 * sA * (kSqrtPi * cg.f + sqrt(cg.A0) * cg.dfdr) - Ag * cg.dfdr) * cg.drdx;
    y[v] = 2.0 * x[3 * v + 1] - c.qh[v] - x[3 * v + 2];
    y[3 + v] = 2.0 * x[10 + 3 * v] - c.Ah[v] - x[11 + 3 * v];
    y[12 + v] = x[3 * v] - c.qe[v] + s * c.dtdex[v] * (Fg2 - c.Fh2[v]) - c.dt / 2.0 * (Sg2 + c.Sh2[v]);
    y[15 + v] = x[9 + 3 * v] - c.Ae[v] + s * c.dtdex[v] * (qg - c.qh[v]);
    double Ah = x[10 + 3 * v], A1 = x[9 + 3 * v];
    double sAh = sqrt(Ah), sA1 = sqrt(A1), sA0e = sqrt(c.A0e[v]);
    ph[v] = c.fe[v] * (1.0 - sqrt(c.A0e[v] / Ah));
    pf[v] = c.fe[v] * (1.0 - sqrt(c.A0e[v] / A1));
    double k = c.fe[v] * sA0e / 2.0;
    l.dPh[v] = k / (Ah * sAh);
    l.dPf[v] = k / (A1 * sA1);
    double u = qg / Ag;
    l.gq[v] = s * c.dtdex[v] * 2.0 * qg / Ag + c.dt * c.rpi_dbRe[v] / sA;
    l.gA[v] = s * c.dtdex[v] * (-(u * u) + cj.f / 2.0 * sqrt(cj.A0 / Ag)) -
              c.dt / 2.0 * (c.rpi_dbRe[v] * qg / (Ag * sA) +
                            (1.0 / sA * (kSqrtPi * cj.f + sqrt(cj.A0) * cj.dfdr) - cj.dfdr) * cj.drdx);
  }
  y[6] = x[0] - x[3] - x[6];
  y[7] = x[1] - x[4] - x[7];
  y[8] = ph[0] - ph[1];
  y[9] = ph[0] - ph[2];
  y[10] = pf[0] - pf[1];
  y[11] = pf[0] - pf[2];
}

// Solve J d = y using the structure of the 18x18 matrix (an:601-663): rows
// 0-5 and 12-17 each define one unknown in terms of the ghost values
// (qg_v, Ag_v); substituting them into rows 6-11 leaves a 3x3 system for the
// three dqg (rows 7, 10, 11) and then one for the three dAg (rows 8, 9, 6).
// Exact-arithmetic equal to the dense solve numpy does; iteration counts were
// checked identical.  y <- d.  Returns 1 if a 3x3 pivot is zero.
AF_HD int junction_solve_structured(const JunctionCtx& c, const JunctionLin& l, double* y) {
  const double sig[3] = {1.0, -1.0, -1.0};
  // J[15+v, qg_v] = s*dt/dex (row 17 with d1's ratio, an:662)
  const double cq[3] = {c.dtdex[0], -c.dtdex[1], -c.dtdex[1]};
  double b[3];
  // rows 7, 10, 11 -> dqg
  b[0] = y[7] - (y[0] - y[1] - y[2]) / 2.0;
  b[1] = y[10] - l.dPf[0] * y[15] + l.dPf[1] * y[16];
  b[2] = y[11] - l.dPf[0] * y[15] + l.dPf[2] * y[17];
  if (solve_arrow3(0.5, -0.5, -0.5, -l.dPf[0] * cq[0], l.dPf[1] * cq[1], -l.dPf[0] * cq[0], l.dPf[2] * cq[2], b))
    return 1;
  double dqg[3] = {b[0], b[1], b[2]};
  // rows 6, 8, 9 -> dAg
  double acc = 0.0;
#pragma unroll
  for (int v = 0; v < 3; ++v) acc += sig[v] * (y[12 + v] - l.gq[v] * dqg[v]);
  b[0] = y[6] - acc;
  b[1] = y[8] - l.dPh[0] * y[3] / 2.0 + l.dPh[1] * y[4] / 2.0;
  b[2] = y[9] - l.dPh[0] * y[3] / 2.0 + l.dPh[2] * y[5] / 2.0;
  if (solve_arrow3(-sig[0] * l.gA[0], -sig[1] * l.gA[1], -sig[2] * l.gA[2], l.dPh[0] / 2.0, -l.dPh[1] / 2.0,
                   l.dPh[0] / 2.0, -l.dPh[2] / 2.0, b))
    return 1;
  double d[18];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    double dAg = b[v];
    d[3 * v + 2] = dqg[v];
    d[11 + 3 * v] = dAg;
    d[3 * v + 1] = (y[v] + dqg[v]) / 2.0;
    d[10 + 3 * v] = (y[3 + v] + dAg) / 2.0;
    d[9 + 3 * v] = y[15 + v] - cq[v] * dqg[v];
    d[3 * v] = y[12 + v] - l.gq[v] * dqg[v] - l.gA[v] * dAg;
  }
#pragma unroll
  for (int i = 0; i < 18; ++i) y[i] = d[i];
  return 0;
}

// Dense path, only for the reference's singular branch (an:703-708).
AF_HD int junction_dense_perturbed(const JunctionCtx& c, const double* x, double* y) {
  double J[18 * 18];
  junction_jacobian(c, x, J, 18);
  junction_residual(c, x, y);
  for (int i = 0; i < 18; ++i) J[i * 18 + i] += 1.0e-6;
  y[0] += 1.0e-6;
  return lu_solve18(J, 18, y);
}

// newton (an:668-710).  Returns the number of linear solves; *flags gets
// AF_FLAG_* bits.
AF_HD int junction_newton(const JunctionCtx& c, double* x, int k_max, double tol, uint32_t* flags) {
  double y[18];
  JunctionLin lin;
  int solves = 0;
  int k = 0;
  for (; k < k_max; ++k) {
    junction_linearise(c, x, y, lin);
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 18; ++i) nrm += y[i] * y[i];
    nrm = sqrt(nrm);
    if (nrm < tol) break;
    if (junction_solve_structured(c, lin, y)) {
      *flags |= AF_FLAG_SINGULAR;
      if (junction_dense_perturbed(c, x, y)) break;
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) x[i] -= y[i];
    ++solves;
  }
  if (k == k_max) *flags |= AF_FLAG_JUNCTION_KMAX;
  return solves;
}

// ---------------------------------------------------------------------------
// Hot-path form of the same bifurcation system, organised per vessel so that
// three lanes of a warp can work on the three vessels of one junction.  All
// sqrt/div of the ghost state are derived from reciprocal square roots (three
// per vessel and Newton iteration instead of 22 div/sqrt), and everything that
// does not depend on the iterate x is folded into JVessel once per time step.
// ---------------------------------------------------------------------------
struct JVessel {
  double s;                       // +1 parent outlet, -1 daughter inlet
  double dt, dtdex, hdt;          // dt, dt/dex, dt/2
  double qe, Ae, qh, Ah, Fh2, Sh2;
  double gF, gS1, gS2;            // ghost point:  Fg2 = qg^2 + gF sqrt(Ag),  Sg2 = src_k qg/sqrt(Ag) + gS1 sqrt(Ag) - gS2 Ag
  double jF, jS1, jS2;            // Jacobian point (an:589-597)
  double src_k, rk;               // -2 sqrt(pi)/(db Re), sqrt(pi)/(db Re)
  double fe, sA0e, kP;            // f, sqrt(A0) at the junction end; fe*sA0e/2
};

struct JVesselLin {
  double y0, y3, y12, y15;        // rows v, 3+v, 12+v, 15+v of problem_function
  double ph, pf;                  // pressure terms of rows 8-11
  double dPh, dPf, gq, gA;        // state-dependent Jacobian entries
};

AF_HD void jvessel_prepare(JVessel& j, int v, const Vessel& ves, const double* A, const double* q, int Nx,
                           double dt, double dex) {
  const double L = ves.L;
  const double xe = v == 0 ? L : 0.0;
  const double xi = v == 0 ? L - dex : dex;
  const double xg = v == 0 ? L + dex / 2.0 : -dex / 2.0;
  const double xh = v == 0 ? L - dex / 2.0 : dex / 2.0;
  const double xj = v == 0 ? L + dex : -dex;
  double Ae, qe, Ai, qi, Ah, qh;
  interp_state(A, q, Nx, ves.dx, xe, Ae, qe);
  interp_state(A, q, Nx, ves.dx, xi, Ai, qi);
  if (v == 0)
    u_half(ves, dt, xi, xe, Ai, qi, Ae, qe, Ah, qh);
  else
    u_half(ves, dt, xe, xi, Ae, qe, Ai, qi, Ah, qh);
  const Coef ch = coef_at(ves, xh), cg = coef_at(ves, xg), cj = coef_at(ves, xj);
  j.s = v == 0 ? 1.0 : -1.0;
  j.dt = dt; j.dtdex = AF_DIV(dt, dex); j.hdt = dt / 2.0;
  j.qe = qe; j.Ae = Ae; j.qh = qh; j.Ah = Ah;
  j.Fh2 = bc_flux2(ch, Ah, qh);
  j.Sh2 = bc_source2(ves, ch, Ah, qh);
  const double sg = af_sqrt_pos(cg.A0), sj = af_sqrt_pos(cj.A0);
  j.gF = cg.f * sg;
  j.gS1 = 2.0 * (kSqrtPi * cg.f + sg * cg.dfdr) * cg.drdx;
  j.gS2 = cg.dfdr * cg.drdx;
  j.jF = cj.f / 2.0 * sj;
  j.jS1 = (kSqrtPi * cj.f + sj * cj.dfdr) * cj.drdx;
  j.jS2 = cj.dfdr * cj.drdx;
  j.src_k = ves.src_k; j.rk = ves.rpi_dbRe;
  j.fe = v == 0 ? ves.fL : ves.f0;
  j.sA0e = af_sqrt_pos(v == 0 ? ves.A0L : ves.A00);
  j.kP = j.fe * j.sA0e / 2.0;
}

// xv = the six unknowns of this vessel: q^{n+1}, q^{n+1/2}, q ghost, A^{n+1}, A^{n+1/2}, A ghost
AF_HD void jvessel_linearise(const JVessel& j, const double xv[6], JVesselLin& o) {
  const double q1 = xv[0], qh = xv[1], qg = xv[2], A1 = xv[3], Ah = xv[4], Ag = xv[5];
  const double y = af_rsqrt_fast(Ag), sA = Ag * y, iA = y * y;
  const double Fg2 = qg * qg + j.gF * sA;
  const double Sg2 = j.src_k * qg * y + j.gS1 * sA - j.gS2 * Ag;
  o.y0 = 2.0 * qh - j.qh - qg;
  o.y3 = 2.0 * Ah - j.Ah - Ag;
  o.y12 = q1 - j.qe + j.s * j.dtdex * (Fg2 - j.Fh2) - j.hdt * (Sg2 + j.Sh2);
  o.y15 = A1 - j.Ae + j.s * j.dtdex * (qg - j.qh);
  const double yh = af_rsqrt_fast(Ah), y1 = af_rsqrt_fast(A1);
  o.ph = j.fe * (1.0 - j.sA0e * yh);
  o.pf = j.fe * (1.0 - j.sA0e * y1);
  o.dPh = j.kP * (yh * yh * yh);
  o.dPf = j.kP * (y1 * y1 * y1);
  const double u = qg * iA;
  o.gq = j.s * j.dtdex * 2.0 * u + j.dt * j.rk * y;
  o.gA = j.s * j.dtdex * (-(u * u) + j.jF * y) - j.hdt * (j.rk * u * y + (j.jS1 * y - j.jS2));
}

// Residual vector y[18] from the three vessels' shares.
AF_HD void junction_gather(const JVesselLin l[3], const double* x, double* y) {
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    y[v] = l[v].y0; y[3 + v] = l[v].y3; y[12 + v] = l[v].y12; y[15 + v] = l[v].y15;
  }
  y[6] = x[0] - x[3] - x[6];
  y[7] = x[1] - x[4] - x[7];
  y[8] = l[0].ph - l[1].ph;
  y[9] = l[0].ph - l[2].ph;
  y[10] = l[0].pf - l[1].pf;
  y[11] = l[0].pf - l[2].pf;
}

// One Newton update x -= J^-1 y from the gathered shares (structured solve).
// Returns 0 ok, 1 if a 3x3 pivot vanished (caller takes the dense path).
AF_HD int junction_update(const JVesselLin l[3], double dtdex_p, double dtdex_d1, double* y) {
  JunctionCtx c;                       // only dtdex is read by the structured solve
  c.dtdex[0] = dtdex_p; c.dtdex[1] = dtdex_d1; c.dtdex[2] = dtdex_d1;
  JunctionLin lin;
#pragma unroll
  for (int v = 0; v < 3; ++v) { lin.dPh[v] = l[v].dPh; lin.dPf[v] = l[v].dPf; lin.gq[v] = l[v].gq; lin.gA[v] = l[v].gA; }
  return junction_solve_structured(c, lin, y);
}

// Serial Newton over the three vessels (one thread per junction: ensembles).
// Returns the number of linear solves, or -1 if a 3x3 pivot vanished (the caller
// then re-runs the verbatim dense path, which owns the reference's singular branch).
AF_HD int junction_newton_fast(const JVessel jv[3], double* x, int k_max, double tol, uint32_t* flags) {
  double y[18];
  JVesselLin l[3];
  int solves = 0, k = 0;
  for (; k < k_max; ++k) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const double xv[6] = {x[3 * v], x[3 * v + 1], x[3 * v + 2], x[9 + 3 * v], x[10 + 3 * v], x[11 + 3 * v]};
      jvessel_linearise(jv[v], xv, l[v]);
    }
    junction_gather(l, x, y);
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;      // same three partial sums as the warp version
#pragma unroll
    for (int i = 0; i < 6; ++i) { n0 = fma(y[i], y[i], n0); n1 = fma(y[6 + i], y[6 + i], n1); n2 = fma(y[12 + i], y[12 + i], n2); }
    const double nrm = (n0 + n1) + n2;
    if (nrm < tol * tol) break;          // ||y|| < tol (an:684) without the square root on the serial chain
    if (junction_update(l, jv[0].dtdex, jv[1].dtdex, y)) return -1;
#pragma unroll
    for (int i = 0; i < 18; ++i) x[i] -= y[i];
    ++solves;
  }
  if (k == k_max) *flags |= AF_FLAG_JUNCTION_KMAX;
  return solves;
}

// adjust_bifurcation_step (an:713-737): M of one vessel end
AF_HD double junction_cfl(const Vessel& ves, int v, const double* A, const double* q, int Nx) {
  Coef ce;
  ce.f = v == 0 ? ves.fL : ves.f0;
  ce.A0 = v == 0 ? ves.A0L : ves.A00;
  double Ae, qe;
  interp_state(A, q, Nx, ves.dx, v == 0 ? ves.L : 0.0, Ae, qe);
  return cfl_term(ves, ce, Ae, qe);
}

// ---------------------------------------------------------------------------
// Per-artery theta-scheme weak form (ar:191-230) + [3P] assembly, one cell.
// ---------------------------------------------------------------------------
struct ArteryConst {
  double h, dt, theta;
  double Kr;            // 2*sqrt(pi)/(db*Re)
};

// Coefficients at the three Gauss points of one cell, derived from the P2
// interpolants f_h, A0_h, dfdr_h, drdx_h ([3P]):
//   c1 = f_h*sqrt(A0_h), t2 = (sqrt(pi) f_h + sqrt(A0_h) dfdr_h) drdx_h, t3 = dfdr_h*drdx_h
struct CellCoef {
  double c1[3], t2[3], t3[3];
};

AF_HD void p2_to_gauss(double vl, double vr, double vm, double out[3]) {
  const double xi[3] = {kXi0, kXi1, kXi2};
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    double n0 = (1.0 - xi[g]) * (1.0 - 2.0 * xi[g]);
    double n1 = xi[g] * (2.0 * xi[g] - 1.0);
    double nm = 4.0 * xi[g] * (1.0 - xi[g]);
    out[g] = vl * n0 + vr * n1 + vm * nm;
  }
}

AF_HD CellCoef cell_coef(const Vessel& v, double xl, double xr) {
  Coef cl = coef_exact(v, xl), cr = coef_exact(v, xr), cm = coef_exact(v, 0.5 * (xl + xr));
  double f[3], A0[3], df[3], dr[3];
  p2_to_gauss(cl.f, cr.f, cm.f, f);
  p2_to_gauss(cl.A0, cr.A0, cm.A0, A0);
  p2_to_gauss(cl.dfdr, cr.dfdr, cm.dfdr, df);
  p2_to_gauss(cl.drdx, cr.drdx, cm.drdx, dr);
  CellCoef c;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    double sA0 = sqrt(A0[g]);
    c.c1[g] = f[g] * sA0;
    c.t2[g] = (kSqrtPi * f[g] + sA0 * df[g]) * dr[g];
    c.t3[g] = df[g] * dr[g];
  }
  return c;
}

// F2 = q^2/A + f sqrt(A0 A), S2 (ar:197-222) and their derivatives at one point.
// One reciprocal square root y = A^-1/2 gives sqrt(A) = A*y and 1/A = y*y.
struct PointTerms {
  double F2, S2, dF2dA, dF2dq, dS2dA, dS2dq;
};

template <bool JAC>
AF_HD PointTerms point_terms(double A, double q, double c1, double t2, double t3, double Kr) {
  PointTerms t;
  const double isA = af_rsqrt_fast(A);
  const double sA = A * isA;
  const double qi = q * (isA * isA);
  t.F2 = fma(q, qi, c1 * sA);
  t.S2 = fma(-Kr * q, isA, fma(2.0 * t2, sA, -A * t3));
  if (JAC) {
    t.dF2dA = fma(-qi, qi, (0.5 * c1) * isA);
    t.dF2dq = 2.0 * qi;
    t.dS2dA = fma(isA, fma(0.5 * Kr, qi, t2), -t3);
    t.dS2dq = -Kr * isA;
  }
  return t;
}

// Flux + source integrals of one cell at state (A, q) with 3-point Gauss:
//   G_i = int (F2 phi_i' + S2 phi_i) dx      (i = left, right vertex)
// and, if JAC, the state-dependent q-row Jacobian sums WITHOUT the factor
// -dt*theta and without the mass matrix:
//   JA[i][j] = int (dF2dA phi_j phi_i' + dS2dA phi_j phi_i),  Jq[i][j] likewise with d/dq.
// The new iterate carries DOLFIN_EPS on A (ar:197-201, 215-218).
struct CellInt {
  double Gl, Gr;
  double JA[4], Jq[4];   // [ll, lr, rl, rr]
};

template <bool JAC>
AF_HD CellInt cell_integrals(double h, double Kr, const CellCoef& cc, double Al, double Ar, double ql, double qr) {
  // quadrature weight times the P1 shape-function products, folded at compile time so
  // that every Gauss point adds one FMA to each of the 13 sums
  const double xi[3] = {kXi0, kXi1, kXi2};
  const double wq[3] = {kW0, kW1, kW2};
  CellInt o;
  double sF = 0.0, sS0 = 0.0, sS1 = 0.0;
  double a0 = 0.0, a1 = 0.0, b00 = 0.0, b01 = 0.0, b11 = 0.0;
  double c0 = 0.0, c1 = 0.0, d00 = 0.0, d01 = 0.0, d11 = 0.0;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const double p0 = 1.0 - xi[g], p1 = xi[g], w = wq[g];
    const double k0 = w * p0, k1 = w * p1, k00 = w * p0 * p0, k01 = w * p0 * p1, k11 = w * p1 * p1;
    double A = Al * p0 + Ar * p1 + kDolfinEps, q = ql * p0 + qr * p1;
    PointTerms t = point_terms<JAC>(A, q, cc.c1[g], cc.t2[g], cc.t3[g], Kr);
    sF = fma(w, t.F2, sF);
    sS0 = fma(k0, t.S2, sS0);
    sS1 = fma(k1, t.S2, sS1);
    if (JAC) {
      a0 = fma(k0, t.dF2dA, a0); a1 = fma(k1, t.dF2dA, a1);
      b00 = fma(k00, t.dS2dA, b00); b01 = fma(k01, t.dS2dA, b01); b11 = fma(k11, t.dS2dA, b11);
      c0 = fma(k0, t.dF2dq, c0); c1 = fma(k1, t.dF2dq, c1);
      d00 = fma(k00, t.dS2dq, d00); d01 = fma(k01, t.dS2dq, d01); d11 = fma(k11, t.dS2dq, d11);
    }
  }
  // phi_l' = -1/h, phi_r' = +1/h; cell measure h
  o.Gl = -sF + h * sS0;
  o.Gr = sF + h * sS1;
  if (JAC) {
    o.JA[0] = -a0 + h * b00; o.JA[1] = -a1 + h * b01; o.JA[2] = a0 + h * b01; o.JA[3] = a1 + h * b11;
    o.Jq[0] = -c0 + h * d00; o.Jq[1] = -c1 + h * d01; o.Jq[2] = c0 + h * d01; o.Jq[3] = c1 + h * d11;
  }
  return o;
}

// Exterior-facet term "+F2*v2*ds" on the q row of an end vertex (ar:197-198,
// 206-207; no outward normal): F2 and its derivatives at the vertex.
// c1v = f(x_v)*sqrt(A0(x_v)).
template <bool JAC>
AF_HD PointTerms facet_terms(double c1v, double A, double q) {
  return point_terms<JAC>(A + kDolfinEps, q, c1v, 0.0, 0.0, 0.0);
}

// One vertex row of the Newton system from the integrals of its two cells.
//   left  = integrals of cell i-1 (this vertex is its right vertex), unused at i == 0
//   own   = integrals of cell i   (this vertex is its left vertex),  unused at i == ne
// U values of vertices i-1, i, i+1 (um/up unused at the ends).  Closed forms for
// the mass matrix (h/6 [1 4 1]) and the q_x term (+-1/2).
// The residual is split as R = I(U) + E(Un)  (SURVEY 9.2):
//   I = M U + dt*theta*X(U),   E = -M Un + dt*(1-theta)*X(Un),
//   X_A = int q_x phi_i,  X_q = -G_i + F2|facet.
struct VertexX {
  double mA, mq, XA, Xq;
};

AF_HD VertexX vertex_x(int i, int ne, double h, double Am, double qm, double A0, double q0, double Ap, double qp,
                       double G_left_r, double G_own_l, double facetF2) {
  VertexX o;
  const double h6 = h / 6.0;
  if (i == 0) {
    o.mA = h6 * (2.0 * A0 + Ap); o.mq = h6 * (2.0 * q0 + qp);
    o.XA = 0.5 * (qp - q0);
    o.Xq = -G_own_l + facetF2;
  } else if (i == ne) {
    o.mA = h6 * (Am + 2.0 * A0); o.mq = h6 * (qm + 2.0 * q0);
    o.XA = 0.5 * (q0 - qm);
    o.Xq = -G_left_r + facetF2;
  } else {
    o.mA = h6 * (Am + 4.0 * A0 + Ap); o.mq = h6 * (qm + 4.0 * q0 + qp);
    o.XA = 0.5 * (qp - qm);
    o.Xq = -(G_left_r + G_own_l);
  }
  return o;
}

struct VertexRow {
  double lo[4], dg[4], up[4];
};

// Jacobian row blocks d R[(i, r)] / d U[(i-1 | i | i+1, c)], before Dirichlet rows.
AF_HD VertexRow vertex_jacobian(int i, int ne, double h, double dth, const double* JA_left, const double* Jq_left,
                                const double* JA_own, const double* Jq_own, double facet_dA, double facet_dq) {
  VertexRow r;
  const double h6 = h / 6.0;
  double jA = 0.0, jq = 0.0;
  if (i > 0) {
    r.lo[0] = h6; r.lo[1] = -0.5 * dth;
    r.lo[2] = -dth * JA_left[2]; r.lo[3] = h6 - dth * Jq_left[2];
    jA += JA_left[3]; jq += Jq_left[3];
  } else {
    r.lo[0] = r.lo[1] = r.lo[2] = r.lo[3] = 0.0;
  }
  if (i < ne) {
    r.up[0] = h6; r.up[1] = 0.5 * dth;
    r.up[2] = -dth * JA_own[1]; r.up[3] = h6 - dth * Jq_own[1];
    jA += JA_own[0]; jq += Jq_own[0];
  } else {
    r.up[0] = r.up[1] = r.up[2] = r.up[3] = 0.0;
  }
  const double mdiag = (i == 0 || i == ne) ? 2.0 * h6 : 4.0 * h6;
  r.dg[0] = mdiag;
  r.dg[1] = (i == 0) ? -0.5 * dth : (i == ne ? 0.5 * dth : 0.0);
  r.dg[2] = -dth * jA + dth * facet_dA;
  r.dg[3] = mdiag - dth * jq + dth * facet_dq;
  return r;
}

// Everything a thread needs to turn cell integrals into the Newton row of a vertex.
struct ArteryCtx {
  int ne, np;
  double h, Kr, dth, dte;     // cell size, 2 sqrt(pi)/(db Re), dt*theta, dt*(1-theta)
  double c1_0, c1_L;          // f*sqrt(A0) at x = 0 and x = L (facet terms)
  int root, leaf;
  double gAin, gqin, gAout, gqout;   // Dirichlet data (ar:171-189)
};

struct Row {
  double lo[4], dg[4], up[4], r[2];
};

AF_HD void row_identity(Row& w) {
#pragma unroll
  for (int c = 0; c < 4; ++c) { w.lo[c] = 0.0; w.up[c] = 0.0; }
  w.dg[0] = 1.0; w.dg[1] = 0.0; w.dg[2] = 0.0; w.dg[3] = 1.0;
  w.r[0] = 0.0; w.r[1] = 0.0;
}

// Newton row of vertex i (residual + 2x2 blocks, Dirichlet rows applied) from
// the right part of cell i-1 (Gr, JA[2..3], Jq[2..3] in `left`) and the left
// part of cell i (Gl, JA[0..1], Jq[0..1] in `own`).  (Am,qm), (Ai,qi), (Ap,qp): the
// iterate U at vertices i-1, i, i+1 (anything where the vertex does not exist).
// On the first pass (U == Un) the explicit part E of the residual is produced,
// afterwards it is consumed.  Vertices i >= np are padding (identity rows).
template <bool JAC>
AF_HD void build_row(const ArteryCtx& c, int i, bool first_pass, double Am, double qm, double Ai, double qi,
                     double Ap, double qp, const CellInt& left, const CellInt& own, double& EA, double& Eq,
                     Row& row) {
  if (i >= c.np) {
    row_identity(row);
    return;
  }
  const int ne = c.ne;
  const bool end = (i == 0 || i == ne);
  double fF2 = 0.0, fdA = 0.0, fdq = 0.0;
  if (end) {
    PointTerms ft = facet_terms<JAC>(i == 0 ? c.c1_0 : c.c1_L, Ai, qi);
    fF2 = ft.F2;
    if (JAC) { fdA = ft.dF2dA; fdq = ft.dF2dq; }
  }
  VertexX vx = vertex_x(i, ne, c.h, Am, qm, Ai, qi, Ap, qp, left.Gr, own.Gl, fF2);
  if (first_pass) {
    EA = -vx.mA + c.dte * vx.XA;
    Eq = -vx.mq + c.dte * vx.Xq;
  }
  row.r[0] = vx.mA + c.dth * vx.XA + EA;
  row.r[1] = vx.mq + c.dth * vx.Xq + Eq;
  if (JAC) {
    VertexRow vr = vertex_jacobian(i, ne, c.h, c.dth, left.JA, left.Jq, own.JA, own.Jq, fdA, fdq);
#pragma unroll
    for (int k = 0; k < 4; ++k) { row.lo[k] = vr.lo[k]; row.dg[k] = vr.dg[k]; row.up[k] = vr.up[k]; }
  }
  if (end) {
    // [3P] DirichletBC rows: residual = U - g, Jacobian row = identity
    const bool fixA = (i == 0) ? !c.root : true;
    const bool fixq = (i == 0) ? true : !c.leaf;
    const double gA = (i == 0) ? c.gAin : c.gAout, gq = (i == 0) ? c.gqin : c.gqout;
    if (fixA) {
      row.r[0] = Ai - gA;
      row.dg[0] = 1.0; row.dg[1] = 0.0;
      row.lo[0] = row.lo[1] = row.up[0] = row.up[1] = 0.0;
    }
    if (fixq) {
      row.r[1] = qi - gq;
      row.dg[2] = 0.0; row.dg[3] = 1.0;
      row.lo[2] = row.lo[3] = row.up[2] = row.up[3] = 0.0;
    }
  }
}

AF_HD void cell_zero(CellInt& c) {
  c.Gl = 0.0; c.Gr = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { c.JA[k] = 0.0; c.Jq[k] = 0.0; }
}

// ---------------------------------------------------------------------------
// Register-resident part of the block cyclic reduction.  A thread owns four
// consecutive vertices (a, b, c, d = 4t .. 4t+3).  Level 0 eliminates a and c,
// level 1 eliminates b; what remains is one row per thread (vertex d) coupled
// to the d of the neighbouring threads -- the reduced system that goes through
// shared memory.  An eliminated vertex is kept as  x = g - P x_left - Q x_right.
// ---------------------------------------------------------------------------
struct Elim {
  double P[4], Q[4], g[2];
};

AF_HD void mul2(const double a[4], const double b[4], double o[4]) {
  o[0] = a[0] * b[0] + a[1] * b[2];
  o[1] = a[0] * b[1] + a[1] * b[3];
  o[2] = a[2] * b[0] + a[3] * b[2];
  o[3] = a[2] * b[1] + a[3] * b[3];
}

AF_HD void inv2v(const double d[4], double o[4]) {
  double r = af_rcp_fast(d[0] * d[3] - d[1] * d[2]);
  o[0] = d[3] * r; o[1] = -d[1] * r; o[2] = -d[2] * r; o[3] = d[0] * r;
}

AF_HD Elim make_elim(const Row& w) {
  Elim e;
  double inv[4];
  inv2v(w.dg, inv);
  mul2(inv, w.lo, e.P);
  mul2(inv, w.up, e.Q);
  e.g[0] = inv[0] * w.r[0] + inv[1] * w.r[1];
  e.g[1] = inv[2] * w.r[0] + inv[3] * w.r[1];
  return e;
}

// w.dg -= a * b (2x2 blocks) and w.r -= a * g, as FMA chains: two instructions per
// entry instead of multiply, multiply-add and subtract
AF_HD void sub_mul2(double d[4], const double a[4], const double b[4]) {
  d[0] = fma(-a[1], b[2], fma(-a[0], b[0], d[0]));
  d[1] = fma(-a[1], b[3], fma(-a[0], b[1], d[1]));
  d[2] = fma(-a[3], b[2], fma(-a[2], b[0], d[2]));
  d[3] = fma(-a[3], b[3], fma(-a[2], b[1], d[3]));
}

AF_HD void sub_mulv(double r[2], const double a[4], const double g[2]) {
  r[0] = fma(-a[1], g[1], fma(-a[0], g[0], r[0]));
  r[1] = fma(-a[3], g[1], fma(-a[2], g[0], r[1]));
}

// o = -(a * b)
AF_HD void neg_mul2(const double a[4], const double b[4], double o[4]) {
  o[0] = fma(-a[1], b[2], -a[0] * b[0]);
  o[1] = fma(-a[1], b[3], -a[0] * b[1]);
  o[2] = fma(-a[3], b[2], -a[2] * b[0]);
  o[3] = fma(-a[3], b[3], -a[2] * b[1]);
}

// eliminate the LEFT neighbour of row w (coupled through w.lo)
// `couple` = false skips the new coupling block (the last step of a reduction, after
// which only the diagonal block and the right-hand side are read).
AF_HD void apply_left(Row& w, const Elim& e, bool couple = true) {
  sub_mul2(w.dg, w.lo, e.Q);
  sub_mulv(w.r, w.lo, e.g);
  if (couple) {
    double m[4];
    neg_mul2(w.lo, e.P, m);
    w.lo[0] = m[0]; w.lo[1] = m[1]; w.lo[2] = m[2]; w.lo[3] = m[3];
  }
}

// eliminate the RIGHT neighbour of row w (coupled through w.up)
AF_HD void apply_right(Row& w, const Elim& e, bool couple = true) {
  sub_mul2(w.dg, w.up, e.P);
  sub_mulv(w.r, w.up, e.g);
  if (couple) {
    double m[4];
    neg_mul2(w.up, e.Q, m);
    w.up[0] = m[0]; w.up[1] = m[1]; w.up[2] = m[2]; w.up[3] = m[3];
  }
}

AF_HD void back_sub(const Elim& e, const double xl[2], const double xr[2], double x[2]) {
  x[0] = fma(-e.Q[1], xr[1], fma(-e.Q[0], xr[0], fma(-e.P[1], xl[1], fma(-e.P[0], xl[0], e.g[0]))));
  x[1] = fma(-e.Q[3], xr[1], fma(-e.Q[2], xr[0], fma(-e.P[3], xl[1], fma(-e.P[2], xl[0], e.g[1]))));
}

}  // namespace af
