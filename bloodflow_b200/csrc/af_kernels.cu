// af_kernels.cu -- sm_100a kernels and the C ABI (include/arteryfe_b200.h) of
// the artery.fe per-timestep network solve.
//
// Data layout in HBM (all FP64, SoA over vertices):
//   A[2][nb*na*S], q[2][nb*na*S]   Un, ping-pong (cur / 1-cur), S = vertices padded to 32
//   ves[nb*na]                     per-artery constants (struct Vessel)
//   bc[nb*na*4]                    A_in, q_in, A_out, q_out   (the reference's BC holders)
//   dex[nb*na], jx[nb*nj*18]       boundary step, junction unknowns (warm start)
//   cellcoef[nt][9][S]             Gauss-point coefficients, tapered arteries only
//   f_art[nb*na], f_out[nb*na]     release/acquire flags (completed steps per artery / junction)
//   mbox[2][nb*na*4]               Dirichlet mailbox junction -> artery (data = flag), by step parity
// Ensembles: one launch of k_boundary (all junctions + leaves of all networks) and one of
// k_artery (one CTA per artery) make a time step, chained by programmatic dependent launch
// in both directions; large ensembles are cut into member shares that run this chain on
// their own forked streams.  Single networks: the persistent pair k_artery_p / k_junction_p
// advances a chunk of time steps per launch, ordered by the flags and the mailbox alone
// (or, AF_NO_PERSIST=1, the same two per-step kernels with those flags).  af_net_step picks.
// Host side: the C ABI, a pinned ring that streams snapshots out during the loop,
// and an exact-size cache of released device / pinned blocks.  See DESIGN.md.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/arteryfe_b200.h"
#include "af_physics.cuh"

using namespace af;

#define AF_NO_FAIL 0xffffffffffffffffull

// ---- release / acquire flags in global memory (the semaphore pattern: one thread of the producer
// releases after a CTA / warp barrier, one thread of the consumer acquires before one).  In the PTX
// memory model a release is "fence.acq_rel ; st.relaxed" and an acquire "ld.relaxed ; fence.acq_rel";
// the fences are spelled out so that a producer with several flags pays for ONE (the junction releases
// three artery ends), a consumer spins on plain relaxed loads, and nobody issues the heavier
// sequentially-consistent fence of __threadfence() (ncu: `membar` was the 4th stall reason of k_boundary).
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ---- Dirichlet mailbox (single networks with the dataflow flags).  The boundary -> artery hand-over
// carries at most four doubles per artery, so the DATA is the flag: a junction / leaf task stores each
// value with one relaxed 8-byte store the moment it is known (no fence, no separate flag, nothing else
// written first), and lane k < 4 of the artery's first warp polls "its" slot until it no longer holds
// the sentinel -- a NaN payload no arithmetic produces -- takes the value and puts the sentinel back.
// Two slot sets alternate with the parity of the time step: the slot of step g is written again in
// step g + 2, by a task that has acquired the f_art flag of the artery that reset it (release / acquire
// chain through step g + 1), so the reset precedes that store in the coherence order of the location.
// Measured on the order-8 tree (scripts/timeline.py): fence + three flag stores + the artery's acquire
// and its re-load of nd.bc were 3.5 us of the 25 us junction <-> root-artery cycle that sets the step time.
#define AF_MBOX_EMPTY 0xfff8deadbeef0001ull
__device__ __forceinline__ void mbox_put(double* slot, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(slot), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long mbox_peek(const double* slot) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
  return v;
}
#ifndef AF_POLL_SLEEP_NS
#define AF_POLL_SLEEP_NS 100
#endif
#ifndef AF_POLL_SLEEP_NS_JUNCTION
#define AF_POLL_SLEEP_NS_JUNCTION 0
#endif
// Spin until *p >= want.  A failed run (fail_step set) ends every wait -- the producer may have left
// early; so does an absurdly long wait (~1 s), which is reported as AF_FLAG_SYNC_TIMEOUT instead of
// hanging the device.
#define AF_FLAG_SYNC_TIMEOUT 64u
template <bool SLEEP = false>
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned want, const unsigned long long* fail_step,
                                          uint32_t* status) {
  for (unsigned spins = 0; (int)(ld_relaxed_u32(p) - want) < 0; ++spins) {
    if ((spins & 63u) == 63u) {
      if (*(volatile const unsigned long long*)fail_step != AF_NO_FAIL) return;
      if (spins > (1u << (SLEEP ? 22 : 23))) { atomicOr(status, AF_FLAG_SYNC_TIMEOUT); return; }
    }
    if (SLEEP) __nanosleep(AF_POLL_SLEEP_NS);     // persistent kernels: a waiting warp must not disturb its neighbours
  }
  fence_acq_rel_gpu();
}

// The same wait for a whole warp, every lane on its own flag, with a warp-UNIFORM exit (vote): the lanes of a
// junction warp leave the loop together, so the warp is converged for the shuffles that follow.
__device__ __forceinline__ void wait_flag_warp(const unsigned* p, unsigned want, const unsigned long long* fail_step,
                                               uint32_t* status) {
  for (unsigned spins = 0;; ++spins) {
    const bool ok = (int)(ld_relaxed_u32(p) - want) >= 0;
    if (__all_sync(0xffffffffu, ok)) break;
    if ((spins & 63u) == 63u) {
      if (*(volatile const unsigned long long*)fail_step != AF_NO_FAIL) break;       // same address in every lane: uniform
      if (spins > (1u << 22)) {
        if ((threadIdx.x & 31) == 0) atomicOr(status, AF_FLAG_SYNC_TIMEOUT);
        break;
      }
    }
    if (AF_POLL_SLEEP_NS_JUNCTION) __nanosleep(AF_POLL_SLEEP_NS_JUNCTION);
  }
  fence_acq_rel_gpu();
}

#ifdef AF_TIMELINE
__device__ __forceinline__ unsigned long long af_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define AF_TL(nd, gstep, task, slot)                                                                              \
  do {                                                                                                           \
    if ((nd).tl && (gstep) >= 0)                                                                                 \
      (nd).tl[(((gstep) % (nd).tl_steps) * (size_t)((nd).na + (nd).nj + (nd).nl) + (task)) * 8 + (slot)] = af_globaltimer(); \
  } while (0)
#else
#define AF_TL(nd, gstep, task, slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------
// device view
// ---------------------------------------------------------------------------
struct NetDev {
  int nb, na, nj, nl, Nx, np, S, Nt;
  int root_artery;
  double dt, theta;
  const int* j_p;
  const int* j_d1;
  const int* j_d2;
  const int* leaf_art;
  const int* art_leaf;   // [na] leaf slot or -1
  Vessel* ves;
  double* A[2];
  double* q[2];
  double* bc;
  double* dex;
  double* jx;            // [nb*nj*18] junction unknowns (warm start), current
  double* jx_next;       // second buffer: written by the fused step kernel, then swapped
  const int* art_jin;    // [na] junction in which the artery is a daughter (-1: root)
  const int* art_role;   // [na] 1 / 2: first / second daughter of art_jin
  const int* art_jout;   // [na] junction of which the artery is the parent (-1: leaf)
  const double* wk;      // [nb*nl*3]  R1, R2, CT
  const double* q_ins;
  int q_ins_per_net;
  const int* coef_slot;  // [nb*na] index into cellcoef or -1
  double* cellcoef;
  int* art_its;
  int* j_solves;
  int* picard_its;
  unsigned long long* totals;  // [3]
  uint32_t* flags;             // [nb]
  unsigned long long* fail_step;  // [1] first global step at which an artery Newton failed (AF_NO_FAIL: none)
  int stop_on_fail;            // single networks: every later launch returns at once (DOLFIN raises, ar:244)
  // Dataflow between the kernels of the time loop (single networks, T >= 64): instead of waiting for the
  // whole previous grid, a task waits for exactly the arteries / junctions it depends on.
  //   f_art[ba] = number of time steps artery ba has completed (g + 1 after step g): junction / leaf tasks wait for it
  //   f_out[ba] = g + 1 once the junction whose PARENT vessel is ba has finished step g (its warm start x is
  //               complete for its next instance; per-step launches only)
  //   The Dirichlet data travels the other way through the mailbox below.
  unsigned* f_art;
  unsigned* f_out;
  double* mbox;                   // [2][nb*na*4] Dirichlet mailbox, slot set = parity of the time step
  const int* art_perm;            // persistent artery kernel: artery of CTA blockIdx.x (nullptr: identity)
  int dataflow;
#ifdef AF_TIMELINE
  // debug builds only (make timeline): %globaltimer stamps of the last tl_steps steps,
  // tl[((g % tl_steps) * (na + nj + nl) + task) * 8 + slot]; tasks = arteries, junctions, leaves
  unsigned long long* tl;
  int tl_steps;
#endif
  // solver constants
  double newton_rtol, newton_atol;
  int newton_max_it, junction_k_max, picard_k_max;
  double junction_tol, picard_tol, cfl_margin;
  // per-step counter log
  int* log_art;
  int* log_j;
  int* log_p;
};

// ---------------------------------------------------------------------------
// set-up: derived vessel constants, initial state, x0, coefficient tables
// (ar:67-189, an:258-279, 434-473)
// ---------------------------------------------------------------------------
__global__ void k_setup_vessels(NetDev nd, const double* Ru, const double* Rd, const double* L, const double* k1,
                                const double* k2, const double* k3, const double* rho, const double* Re,
                                const double* db, const double* p0) {
  int ba = blockIdx.x * blockDim.x + threadIdx.x;
  if (ba >= nd.nb * nd.na) return;
  int b = ba / nd.na;
  Vessel v;
  v.Ru = Ru[ba]; v.Rd = Rd[ba]; v.L = L[ba];
  v.dx = v.L / nd.Nx;
  v.k1 = k1[b]; v.k2 = k2[b]; v.k3 = k3[b];
  v.lr_over_L = log(v.Rd / v.Ru) / v.L;
  v.rho = rho[b];
  v.p0 = p0[b];
  v.src_k = -2.0 * kSqrtPi / db[b] / Re[b];
  v.rpi_dbRe = kSqrtPi / db[b] / Re[b];
  v.tapered = (v.Rd != v.Ru) ? 1 : 0;
  Coef c0 = coef_exact(v, 0.0), cL = coef_exact(v, v.L);
  v.f0 = c0.f; v.A00 = c0.A0; v.fL = cL.f; v.A0L = cL.A0; v.dfdr0 = c0.dfdr;
  nd.ves[ba] = v;
}

__global__ void k_setup_state(NetDev nd, const double* q0) {
  int ba = blockIdx.x;
  const Vessel v = nd.ves[ba];
  size_t off = (size_t)ba * nd.S;
  for (int j = threadIdx.x; j < nd.S; j += blockDim.x) {
    double A = 0.0, q = 0.0;
    if (j < nd.np) {
      A = coef_exact(v, v.dx * j).A0;  // Un = (A0(x_j), q0)  (ar:154-157)
      q = q0[ba];
    }
    nd.A[0][off + j] = A; nd.q[0][off + j] = q;
    nd.A[1][off + j] = A; nd.q[1][off + j] = q;
  }
  if (threadIdx.x == 0) {
    // BC holders (ar:171-187)
    nd.bc[ba * 4 + 0] = v.A00;
    nd.bc[ba * 4 + 1] = q0[ba];
    nd.bc[ba * 4 + 2] = v.A0L;
    nd.bc[ba * 4 + 3] = q0[ba];
    nd.dex[ba] = v.dx;                 // ar:98
  }
  int slot = nd.coef_slot[ba];
  if (slot >= 0) {
    double* tab = nd.cellcoef + (size_t)slot * 9 * nd.S;
    for (int e = threadIdx.x; e < nd.Nx; e += blockDim.x) {
      CellCoef cc = cell_coef(v, v.dx * e, v.dx * (e + 1));
      for (int g = 0; g < 3; ++g) {
        tab[(0 + g) * nd.S + e] = cc.c1[g];
        tab[(3 + g) * nd.S + e] = cc.t2[g];
        tab[(6 + g) * nd.S + e] = cc.t3[g];
      }
    }
  }
}

__global__ void k_setup_x(NetDev nd, const double* q0) {
  // initial_x / define_x (an:434-473)
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nd.nb * nd.nj) return;
  int b = t / nd.nj, j = t % nd.nj;
  int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
  double* x = nd.jx + (size_t)t * 18;
  for (int v = 0; v < 3; ++v) {
    int ba = b * nd.na + ia[v];
    const Vessel& ves = nd.ves[ba];
    double Ae = v == 0 ? ves.A0L : ves.A00;
    for (int k = 0; k < 3; ++k) {
      x[3 * v + k] = q0[ba];
      x[9 + 3 * v + k] = Ae;
    }
  }
}

// ---------------------------------------------------------------------------
// boundary update: every junction (18-unknown Newton) and every leaf
// (Windkessel Picard) of every network.  `lanes` threads per task, lane 0 works
// (lanes = 32 gives every task its own warp for single networks; 1 packs
// ensembles densely).                                    (an:740-786)
// ---------------------------------------------------------------------------
// common tail of a junction task: write x, the Dirichlet data and the counters
__device__ __forceinline__ void junction_finish(const NetDev& nd, int b, int j, const int* ia, const double* x,
                                                double dex, int solves, uint32_t fl, int log_slot,
                                                long long flag_step = -1) {
  // set_inner_bc (an:762-764)
  int bp = b * nd.na + ia[0], b1 = b * nd.na + ia[1], b2 = b * nd.na + ia[2];
  if (flag_step >= 0) {
    // dataflow: the six Dirichlet values go straight into the mailbox slots of the three artery ends,
    // before anything else is written (the artery CTAs copy them into nd.bc)
    double* mb = nd.mbox + (size_t)(flag_step & 1) * nd.nb * nd.na * 4;
    mbox_put(&mb[bp * 4 + 2], x[9]);  mbox_put(&mb[bp * 4 + 3], x[0]);
    mbox_put(&mb[b1 * 4 + 0], x[12]); mbox_put(&mb[b1 * 4 + 1], x[3]);
    mbox_put(&mb[b2 * 4 + 0], x[15]); mbox_put(&mb[b2 * 4 + 1], x[6]);
    AF_TL(nd, flag_step, nd.na + j, 4);
  } else {
    nd.bc[bp * 4 + 2] = x[9];  nd.bc[bp * 4 + 3] = x[0];
    nd.bc[b1 * 4 + 0] = x[12]; nd.bc[b1 * 4 + 1] = x[3];
    nd.bc[b2 * 4 + 0] = x[15]; nd.bc[b2 * 4 + 1] = x[6];
  }
  double* xg = nd.jx + ((size_t)b * nd.nj + j) * 18;
  for (int i = 0; i < 18; ++i) xg[i] = x[i];
  nd.dex[bp] = dex;    // the value a non-leaf vessel holds after set_bcs
  // the junction's own flag: its warm start x is complete for its next instance (which acquires it)
  if (flag_step >= 0) st_release_u32(&nd.f_out[bp], (unsigned)(flag_step + 1));
  // bookkeeping nobody waits for: after the release, off the critical path of the step
  nd.j_solves[b * nd.nj + j] = solves;
  if (log_slot >= 0) nd.log_j[(size_t)log_slot * nd.nb * nd.nj + b * nd.nj + j] = solves;
  atomicAdd(&nd.totals[1], (unsigned long long)solves);
  if (fl) atomicOr(&nd.flags[b], fl);
}

// verbatim dense path (reference formulas, 18x18 LU): owns the singular branch an:703-708
__device__ __noinline__ int junction_slow_path(const NetDev& nd, int cur, int b, const int* ia, double dex, double* x,
                                               uint32_t* fl) {
  JunctionCtx c;
  for (int v = 0; v < 3; ++v) {
    int ba = b * nd.na + ia[v];
    size_t off = (size_t)ba * nd.S;
    junction_prepare_vessel(c, v, nd.ves[ba], nd.A[cur] + off, nd.q[cur] + off, nd.Nx, nd.dt, dex);
  }
  return junction_newton(c, x, nd.junction_k_max, nd.junction_tol, fl);
}

// one thread per junction (ensembles: 32 junctions per warp)
__device__ __forceinline__ void junction_task(const NetDev& nd, int cur, int b, int j, int log_slot) {
  int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
  const double* A = nd.A[cur];
  const double* q = nd.q[cur];
  // adjust_bifurcation_step (an:713-737)
  double M = 1.0e300;
  for (int v = 0; v < 3; ++v) {
    int ba = b * nd.na + ia[v];
    size_t off = (size_t)ba * nd.S;
    double m = junction_cfl(nd.ves[ba], v, A + off, q + off, nd.Nx);
    M = m < M ? m : M;
  }
  double dex = AF_DIV((1.0 + nd.cfl_margin) * nd.dt, M);
  JVessel jv[3];
  for (int v = 0; v < 3; ++v) {
    int ba = b * nd.na + ia[v];
    size_t off = (size_t)ba * nd.S;
    jvessel_prepare(jv[v], v, nd.ves[ba], A + off, q + off, nd.Nx, nd.dt, dex);
  }
  double x[18], x0[18];
  const double* xg = nd.jx + ((size_t)b * nd.nj + j) * 18;
  for (int i = 0; i < 18; ++i) x0[i] = x[i] = xg[i];
  uint32_t fl = 0;
  int solves = junction_newton_fast(jv, x, nd.junction_k_max, nd.junction_tol, &fl);
  if (solves < 0) {
    for (int i = 0; i < 18; ++i) x[i] = x0[i];
    fl = 0;
    solves = junction_slow_path(nd, cur, b, ia, dex, x, &fl);
  }
  junction_finish(nd, b, j, ia, x, dex, solves, fl, log_slot);
}

// one warp per junction (single networks): lane v < 3 owns vessel v; the 3x3
// solves are done redundantly by every lane, so no result has to travel back.
// On return every lane holds x; dex, solves, fl are warp-uniform.
// `ia` (the three vessels of the junction) and `ves` (the constants of this lane's
// vessel, ia[lane % 3]) are loaded by the caller: they do not depend on the state,
// so the boundary kernel fetches them before its grid-dependency wait.
__device__ __forceinline__ void junction_solve_warp(const NetDev& nd, int cur, int b, int j, const int* ia,
                                                    const Vessel& ves, const double* xg, double* x, double& dex,
                                                    int& solves, uint32_t& fl, bool& singular_out,
                                                    long long tl_g = -1) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int v = lane % 3;
  const int ba = b * nd.na + (v == 0 ? ia[0] : (v == 1 ? ia[1] : ia[2]));
  const size_t off = (size_t)ba * nd.S;
  const double* A = nd.A[cur] + off;
  const double* q = nd.q[cur] + off;
  // adjust_bifurcation_step (an:713-737): min over the three vessels
  double M = junction_cfl(ves, v, A, q, nd.Nx);
  {
    double m0 = __shfl_sync(FULL, M, 0), m1 = __shfl_sync(FULL, M, 1), m2 = __shfl_sync(FULL, M, 2);
    M = 1.0e300;
    M = m0 < M ? m0 : M; M = m1 < M ? m1 : M; M = m2 < M ? m2 : M;
  }
  dex = AF_DIV((1.0 + nd.cfl_margin) * nd.dt, M);
  JVessel jv;
  jvessel_prepare(jv, v, ves, A, q, nd.Nx, nd.dt, dex);
  const double dtdex_p = __shfl_sync(FULL, jv.dtdex, 0), dtdex_d1 = __shfl_sync(FULL, jv.dtdex, 1);
  for (int i = 0; i < 18; ++i) x[i] = xg[i];
  fl = 0;
  solves = 0;
  int k = 0;
  bool singular = false;
  if (lane == 0) AF_TL(nd, tl_g, nd.na + j, 2);
  for (; k < nd.junction_k_max; ++k) {
#ifdef AF_TIMELINE
    if (lane == 0 && (k == 1 || k == 2)) AF_TL(nd, tl_g, nd.na + j, 4 + k);
#endif
    // this lane's six unknowns, picked by selects: indexing x[] with the run-time v would
    // put the 18 unknowns into local memory for the whole Newton loop
#define AF_PICK3(a, b, c) (v == 0 ? (a) : (v == 1 ? (b) : (c)))
    const double xv[6] = {AF_PICK3(x[0], x[3], x[6]),   AF_PICK3(x[1], x[4], x[7]),    AF_PICK3(x[2], x[5], x[8]),
                          AF_PICK3(x[9], x[12], x[15]), AF_PICK3(x[10], x[13], x[16]), AF_PICK3(x[11], x[14], x[17])};
#undef AF_PICK3
    JVesselLin mine, l[3];
    jvessel_linearise(jv, xv, mine);
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      l[w].y0 = __shfl_sync(FULL, mine.y0, w); l[w].y3 = __shfl_sync(FULL, mine.y3, w);
      l[w].y12 = __shfl_sync(FULL, mine.y12, w); l[w].y15 = __shfl_sync(FULL, mine.y15, w);
      l[w].ph = __shfl_sync(FULL, mine.ph, w); l[w].pf = __shfl_sync(FULL, mine.pf, w);
      l[w].dPh = __shfl_sync(FULL, mine.dPh, w); l[w].dPf = __shfl_sync(FULL, mine.dPf, w);
      l[w].gq = __shfl_sync(FULL, mine.gq, w); l[w].gA = __shfl_sync(FULL, mine.gA, w);
    }
    double y[18];
    junction_gather(l, x, y);
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;      // three partial sums: a third of the serial FMA chain
#pragma unroll
    for (int i = 0; i < 6; ++i) { n0 = fma(y[i], y[i], n0); n1 = fma(y[6 + i], y[6 + i], n1); n2 = fma(y[12 + i], y[12 + i], n2); }
    const double nrm = (n0 + n1) + n2;
    if (nrm < nd.junction_tol * nd.junction_tol) break;      // ||y|| < tol (an:684), no sqrt on the chain
    if (junction_update(l, dtdex_p, dtdex_d1, y)) { singular = true; break; }
#pragma unroll
    for (int i = 0; i < 18; ++i) x[i] -= y[i];
    ++solves;
  }
  if (k == nd.junction_k_max) fl |= AF_FLAG_JUNCTION_KMAX;
  if (lane == 0) AF_TL(nd, tl_g, nd.na + j, 3);
#ifdef AF_TIMELINE
  if (lane == 0 && nd.tl && tl_g >= 0)
    nd.tl[((tl_g % nd.tl_steps) * (size_t)(nd.na + nd.nj + nd.nl) + nd.na + j) * 8 + 7] = (unsigned long long)k;
#endif
  // a vanished 3x3 pivot (warp-uniform): the caller re-runs the step on the verbatim dense path, which owns the
  // reference's singular branch -- on arrays of its own, so that x above never becomes addressable (it would
  // live in local memory for the whole Newton loop) and is not merged with them through memory afterwards
  singular_out = singular;
}

__device__ __forceinline__ void junction_task_warp(const NetDev& nd, int cur, int b, int j, const int* ia,
                                                   const Vessel& ves, int log_slot, long long flag_step = -1) {
  double x[18], dex;
  int solves;
  uint32_t fl;
  const double* xg = nd.jx + ((size_t)b * nd.nj + j) * 18;
  bool singular;
  junction_solve_warp(nd, cur, b, j, ia, ves, xg, x, dex, solves, fl, singular, flag_step);
  // (if-blocks, not an early return: inside the step loop of k_junction_p the warp must be converged
  // again afterwards)
  if (!singular) {
    if ((threadIdx.x & 31) == 0) junction_finish(nd, b, j, ia, x, dex, solves, fl, log_slot, flag_step);
  } else {
    double xs[18];
    uint32_t fls = 0;
    for (int i = 0; i < 18; ++i) xs[i] = xg[i];
    const int s2 = junction_slow_path(nd, cur, b, ia, dex, xs, &fls);
    if ((threadIdx.x & 31) == 0) junction_finish(nd, b, j, ia, xs, dex, s2, fls, log_slot, flag_step);
  }
  __syncwarp();
}

__device__ __forceinline__ void leaf_task(const NetDev& nd, int cur, int b, int l, int log_slot) {
  int ba = b * nd.na + nd.leaf_art[l];
  size_t off = (size_t)ba * nd.S;
  const double* w = nd.wk + ((size_t)b * nd.nl + l) * 3;
  WkResult r = windkessel(nd.ves[ba], nd.A[cur] + off, nd.q[cur] + off, nd.Nx, nd.dt, w[0], w[1], w[2],
                          nd.picard_k_max, nd.picard_tol, nd.cfl_margin);
  nd.bc[ba * 4 + 2] = r.A_out;       // A_out (an:786)
  nd.dex[ba] = r.dex;
  nd.picard_its[b * nd.nl + l] = r.its;
  if (log_slot >= 0) nd.log_p[(size_t)log_slot * nd.nb * nd.nl + b * nd.nl + l] = r.its;
  atomicAdd(&nd.totals[2], (unsigned long long)r.its);
  if (r.its >= nd.picard_k_max) atomicOr(&nd.flags[b], (uint32_t)(r.cycle ? AF_FLAG_PICARD_CYCLE : AF_FLAG_PICARD_KMAX));
}

// one warp per leaf (single networks): lanes 0..2 take the three stencil points
// L-2dex, L-dex, L of windkessel() (an:383-390), lanes 0..1 the two half steps
// (an:393-398); the Picard loop is inherently serial.
// P1 interpolation of the state (interp_state) with the vertex index mapped: TS = 0, vertex i at [i] (Un in
// global memory); TS > 0, vertex i at [(i & 3) * TS + (i >> 2)] (the iterate of a TS-thread artery CTA in
// shared memory, see vix).  Same arithmetic, same order.
template <int TS>
__device__ __forceinline__ void interp_state_mapped(const double* A, const double* q, int Nx, double dx, double x,
                                                    double& Ax, double& qx) {
  const double inv_dx = af_rcp_fast(dx);
  int j = (int)floor(x * inv_dx);
  j = j < 0 ? 0 : (j > Nx - 1 ? Nx - 1 : j);
  double xi = (x - j * dx) * inv_dx;
  const int j0 = TS ? (j & 3) * TS + (j >> 2) : j, j1 = TS ? ((j + 1) & 3) * TS + ((j + 1) >> 2) : j + 1;
  Ax = A[j0] * (1.0 - xi) + A[j1] * xi;
  qx = q[j0] * (1.0 - xi) + q[j1] * xi;
}

// `A`, `q`: Un of the leaf artery -- in global memory (TS = 0: nd.A[cur] + ba * S) or, inside the persistent
// artery kernel, the CTA's own iterate in shared memory (TS = its thread count), which holds the same values
// two L2 round trips closer.
template <int TS = 0>
__device__ __forceinline__ WkResult leaf_solve_warp(const NetDev& nd, const double* A, const double* q, int b, int l,
                                                    int ba, const Vessel& ves, int log_slot, long long flag_step = -1) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int k = lane % 3;
  const double L = ves.L, dt = nd.dt;
  // adjust_dex (ar:349-365)
  Coef cL;
  cL.f = ves.fL; cL.A0 = ves.A0L;
  double AL, qL;
  interp_state_mapped<TS>(A, q, nd.Nx, ves.dx, L, AL, qL);
  const double dex = AF_DIV((1.0 + nd.cfl_margin) * dt, cfl_term(ves, cL, AL, qL));
  // stencil point of this lane and its flux / source (an:385-390)
  const double xk = L - (double)(2 - k) * dex;
  double Ak, qk;
  interp_state_mapped<TS>(A, q, nd.Nx, ves.dx, xk, Ak, qk);
  const Coef ck = coef_at(ves, xk);
  const double Fk = bc_flux2(ck, Ak, qk), Sk = bc_source2(ves, ck, Ak, qk);
  // half step between this point and the next one (compute_U_half, an:393-394)
  const double xn = __shfl_down_sync(FULL, xk, 1), An = __shfl_down_sync(FULL, Ak, 1);
  const double qn = __shfl_down_sync(FULL, qk, 1), Fn = __shfl_down_sync(FULL, Fk, 1);
  const double Sn = __shfl_down_sync(FULL, Sk, 1);
  const double rr = AF_DIV(dt, xn - xk);
  const double Ah = (Ak + An) / 2.0 - rr * (qn - qk);
  const double qh = (qk + qn) / 2.0 - rr * (Fn - Fk) + dt / 4.0 * (Sk + Sn);
  const double xh = L - (1.5 - (double)k) * dex;       // L-1.5dex, L-0.5dex
  const Coef ch = coef_at(ves, xh);
  const double Fh = bc_flux2(ch, Ah, qh), Sh = bc_source2(ves, ch, Ah, qh);
  // value at t_(n+1) in L-dex (an:401-403)
  const double F21 = __shfl_sync(FULL, Fh, 0), S21 = __shfl_sync(FULL, Sh, 0);
  const double F10 = __shfl_sync(FULL, Fh, 1), S10 = __shfl_sync(FULL, Sh, 1);
  const double q1 = __shfl_sync(FULL, qk, 1);
  const double A0s = __shfl_sync(FULL, Ak, 2), q0s = __shfl_sync(FULL, qk, 2);
  const double qm1 = q1 - AF_DIV(dt, dex) * (F10 - F21) + dt / 2.0 * (S10 + S21);
  WkResult r;
  r.A_out = 0.0; r.dex = dex; r.its = 0; r.cycle = 0;
  // (an if-block, not an early return of the other lanes: the warp goes on together afterwards, see junction_task_warp)
  if (lane == 0) {
    AF_TL(nd, flag_step, nd.na + nd.nj + l, 2);
    const double* w = nd.wk + ((size_t)b * nd.nl + l) * 3;
    r = windkessel_picard(ves, dt, dex, A0s, q0s, qm1, w[0], w[1], w[2], nd.picard_k_max, nd.picard_tol);
    AF_TL(nd, flag_step, nd.na + nd.nj + l, 3);
    if (flag_step >= 0)      // dataflow: A_out (an:786) into the mailbox slot of the outlet (see junction_finish)
      mbox_put(&nd.mbox[(size_t)(flag_step & 1) * nd.nb * nd.na * 4 + ba * 4 + 2], r.A_out);
    else
      nd.bc[ba * 4 + 2] = r.A_out;       // A_out (an:786)
    nd.dex[ba] = r.dex;
    AF_TL(nd, flag_step, nd.na + nd.nj + l, 4);
    nd.picard_its[b * nd.nl + l] = r.its;
    if (log_slot >= 0) nd.log_p[(size_t)log_slot * nd.nb * nd.nl + b * nd.nl + l] = r.its;
    atomicAdd(&nd.totals[2], (unsigned long long)r.its);
    if (r.its >= nd.picard_k_max) atomicOr(&nd.flags[b], (uint32_t)(r.cycle ? AF_FLAG_PICARD_CYCLE : AF_FLAG_PICARD_KMAX));
  }
  __syncwarp();
  return r;
}



// WARP = true: one warp per task (single networks); false: one thread per task
// (ensembles).  Two instantiations keep each kernel's code -- and the instruction
// cache footprint of the lone warp that runs a task -- small.
template <bool WARP>
__global__ void k_boundary(NetDev nd, int cur, int log_slot, int b0, int nbl, long long gstep) {
  // networks b0 .. b0+nbl-1 (the whole batch, or the share of one sub-stream)
  long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n_j = (long long)nbl * nd.nj, n_l = (long long)nbl * nd.nl;
  if (WARP) {
    // Topology and vessel constants do not depend on the state: fetch them first,
    // they arrive while this kernel -- which may have been made resident under
    // the artery kernel of the previous step (programmatic dependent launch) --
    // waits for that kernel's Un.  ONLY constants may be read up here: the artery
    // kernel triggers at its very start, possibly while the boundary kernel of the
    // previous step is still running, so two generations of this kernel can be
    // resident together and jx / bc / dex of the previous step may not be final yet
    // (measured: fetching the Newton warm start here broke bitwise reproducibility).
    long long task = gt >> 5;
    const int v = (threadIdx.x & 31) % 3;
    int b = 0, j = 0, l = 0, ba = 0, ia[3] = {0, 0, 0};
    const bool is_junction = task < n_j, is_leaf = !is_junction && task < n_j + n_l;
    if (is_junction) {
      b = b0 + (int)(task / nd.nj); j = (int)(task % nd.nj);
      ia[0] = nd.j_p[j]; ia[1] = nd.j_d1[j]; ia[2] = nd.j_d2[j];
      ba = b * nd.na + (v == 0 ? ia[0] : (v == 1 ? ia[1] : ia[2]));
    } else if (is_leaf) {
      task -= n_j;
      b = b0 + (int)(task / nd.nl); l = (int)(task % nd.nl);
      ba = b * nd.na + nd.leaf_art[l];
    }
    const Vessel ves = nd.ves[ba];
    const bool dataflow = nd.dataflow && gstep >= 0;
#ifdef AF_TIMELINE
    const int tl_task = is_junction ? nd.na + j : nd.na + nd.nj + l;
    if ((is_junction || is_leaf) && (threadIdx.x & 31) == 0) AF_TL(nd, gstep, tl_task, 0);
#endif
    if (dataflow) {
      // Dataflow: no grid-wide wait.  The artery kernel of this step may become resident right away
      // (its CTAs wait for their own flags); this task waits for the arteries whose Un it reads --
      // lane v of a junction for vessel v, a leaf for its artery -- to have completed step gstep - 1.
      cudaTriggerProgrammaticLaunchCompletion();
      if (is_junction || is_leaf) {
        if ((threadIdx.x & 31) < (is_junction ? 3 : 1)) wait_flag(&nd.f_art[ba], (unsigned)gstep, nd.fail_step, &nd.flags[b]);
        // ... and, a junction, for the warm start x its own previous instance left (f_out of its parent vessel)
        if (is_junction && (threadIdx.x & 31) == 3)
          wait_flag(&nd.f_out[b * nd.na + ia[0]], (unsigned)gstep, nd.fail_step, &nd.flags[b]);
        __syncwarp();
#ifdef AF_TIMELINE
        if ((threadIdx.x & 31) == 0) AF_TL(nd, gstep, tl_task, 1);
#endif
      }
    } else {
      cudaGridDependencySynchronize();
      // only now may the artery kernel of THIS step start its prologue: it reads the
      // Un that the previous artery kernel has just finished writing
      cudaTriggerProgrammaticLaunchCompletion();
    }
    const long long flag_step = dataflow ? gstep : -1;
    if (nd.stop_on_fail && *(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL) {
      // an artery solve failed: the run is over -- but nobody may be left waiting for this task
      // (the artery CTAs polling their mailbox slots watch fail_step themselves)
      if (dataflow && (threadIdx.x & 31) == 0 && is_junction)
        st_release_u32(&nd.f_out[b * nd.na + ia[0]], (unsigned)(gstep + 1));
      return;
    }
    if (is_junction)
      junction_task_warp(nd, cur, b, j, ia, ves, log_slot, flag_step);
    else if (is_leaf)
      leaf_solve_warp(nd, nd.A[cur] + (size_t)ba * nd.S, nd.q[cur] + (size_t)ba * nd.S, b, l, ba, ves, log_slot, flag_step);
  } else {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    long long task = gt;
    if (task < n_j) {
      junction_task(nd, cur, b0 + (int)(task / nd.nj), (int)(task % nd.nj), log_slot);
    } else if (task < n_j + n_l) {
      task -= n_j;
      leaf_task(nd, cur, b0 + (int)(task / nd.nl), (int)(task % nd.nl), log_slot);
    }
  }
}

// ---------------------------------------------------------------------------
// per-artery Newton on the theta-scheme weak form: one CTA per artery, every
// thread owns FOUR consecutive vertices.                  (ar:233-251 + [3P])
//
// Per Newton pass a thread (1) integrates its four cells, (2) builds its four
// rows in registers, (3) after the residual test eliminates three of them in
// registers (levels 0 and 1 of the cyclic reduction; the eliminated vertices
// are parked in shared memory as x = g - P x_l - Q x_r), (4) solves the T-row
// reduced system by parallel cyclic reduction with the row kept in registers,
// and (5) back-substitutes its own vertices.  Shared memory, in doubles (T threads):
//   U(A,q) 8T | PCR exchange 10T | parked a,b,c 30T   = 48T
// i.e. 96 KB for the 1001-vertex arteries of the order-8 tree (T = 256), two
// CTAs per SM, so all 255 arteries are co-resident.
// ---------------------------------------------------------------------------
#define AF_ARTERY_SMEM_PER_THREAD 48
#ifndef AF_T32_BLOCKS
#define AF_T32_BLOCKS 16     // resident one-warp CTAs per SM the T = 32 kernel is compiled for
#endif

// CTA-wide barrier; a one-warp CTA (T = 32, the ensemble kernel) only needs the
// warp-level one.
template <int T>
__device__ __forceinline__ void cta_sync() {
  if (T == 32)
    __syncwarp();
  else
    __syncthreads();
}

template <int T>
__device__ __forceinline__ double block_sum(double v, double* red) {
  // fixed-order: lanes by xor-tree, warps in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (T == 32) {
    __syncwarp();          // orders the shared-memory writes of the pass, like the barriers below
    return v;
  }
  int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < T / 32; ++k) s += red[k];
  return s;
}

__device__ __forceinline__ void park(double* base, int T, int t, const Elim& e) {
#pragma unroll
  for (int c = 0; c < 4; ++c) { base[c * T + t] = e.P[c]; base[(4 + c) * T + t] = e.Q[c]; }
  base[8 * T + t] = e.g[0];
  base[9 * T + t] = e.g[1];
}

__device__ __forceinline__ Elim unpark(const double* base, int T, int t) {
  Elim e;
#pragma unroll
  for (int c = 0; c < 4; ++c) { e.P[c] = base[c * T + t]; e.Q[c] = base[(4 + c) * T + t]; }
  e.g[0] = base[8 * T + t];
  e.g[1] = base[9 * T + t];
  return e;
}

// The iterate U lives in shared memory with the four vertices of a thread in four
// planes: vertex i at [(i & 3) * T + (i >> 2)].  A warp reading "its" vertex k (or
// the neighbouring thread's) touches 32 consecutive doubles: no bank conflicts
// (the natural order, stride 4 doubles per thread, is a 4-way conflict on every
// access -- 0.96 M conflicts per launch in profiles/r01_final_ncu_full_summary.txt).
template <int T>
__device__ __forceinline__ int vix(int i) { return (i & 3) * T + (i >> 2); }

template <int T>
__device__ __forceinline__ void row_of(const ArteryCtx& c, int i, bool first, const double* sA, const double* sq,
                                       const CellInt& left, const CellInt& own, double& EA, double& Eq, Row& row) {
  // a neighbour that does not exist is replaced by the vertex itself (build_row ignores it)
  const int jm = vix<T>(i > 0 ? i - 1 : i), j0 = vix<T>(i), jp = vix<T>(i < 4 * T - 1 ? i + 1 : i);
  build_row<true>(c, i, first, sA[jm], sq[jm], sA[j0], sq[j0], sA[jp], sq[jp], left, own, EA, Eq, row);
}

template <int T, bool JAC>
__device__ __forceinline__ CellInt artery_cell(const ArteryCtx& c, const CellCoef& ccu, const double* tab, int S,
                                               const double* sA, const double* sq, int i) {
  CellInt ci;
  cell_zero(ci);
  if (i < c.ne) {
    CellCoef cc = ccu;
    if (tab) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        cc.c1[g] = tab[(0 + g) * S + i];
        cc.t2[g] = tab[(3 + g) * S + i];
        cc.t3[g] = tab[(6 + g) * S + i];
      }
    }
    const int jl = vix<T>(i), jr = vix<T>(i + 1);
    ci = cell_integrals<JAC>(c.h, c.Kr, cc, sA[jl], sA[jr], sq[jl], sq[jr]);
  }
  return ci;
}

// An assembly pass of a thread has two halves separated by one CTA barrier.
// (1) cells_to_smem: the flux/source/Jacobian integrals of its four cells, which
// depend on the iterate U only (not on the Dirichlet data), go to shared memory:
// cell k of thread t into slot t of region k (pa, pb, pc, lo; 10 doubles each).
template <int T>
__device__ __forceinline__ void cells_to_smem(const ArteryCtx& c, const CellCoef& ccu, const double* tab, int S,
                                              const double* sA, const double* sq, double* pa, double* pb,
                                              double* pc, double* lo, int tid = threadIdx.x) {
  // (`tid`: the thread whose four cells are integrated -- normally the caller itself; the cell integrals depend
  // on the iterate in shared memory only, so any thread can stand in for another, see artery_newton)
  const int i0 = 4 * tid;
  if (T == 32) {
    // one-warp (ensemble) kernel: ONE copy of the cell routine, executed four times.
    // With 16 independent one-warp CTAs per SM the instruction cache, not ILP, is
    // what the cells compete for (profiles/r01_c5_ncu_full_summary.txt).
    // regions are contiguous: lo | pa | pb | pc, cell k -> region (k + 1) % 4
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      CellInt ci = artery_cell<T, true>(c, ccu, tab, S, sA, sq, i0 + k);
      double* r = lo + ((k + 1) & 3) * 10 * T;
      r[0 * T + tid] = ci.Gl; r[1 * T + tid] = ci.Gr;
#pragma unroll
      for (int m = 0; m < 4; ++m) { r[(2 + m) * T + tid] = ci.JA[m]; r[(6 + m) * T + tid] = ci.Jq[m]; }
    }
    return;
  }
  double* reg[4] = {pa, pb, pc, lo};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    CellInt ci = artery_cell<T, true>(c, ccu, tab, S, sA, sq, i0 + k);
    double* r = reg[k];
    r[0 * T + tid] = ci.Gl; r[1 * T + tid] = ci.Gr;
#pragma unroll
    for (int m = 0; m < 4; ++m) { r[(2 + m) * T + tid] = ci.JA[m]; r[(6 + m) * T + tid] = ci.Jq[m]; }
  }
}

__device__ __forceinline__ void load_cell_left_part(const double* r, int T, int t, CellInt& ci) {
  ci.Gl = r[0 * T + t];
  ci.JA[0] = r[2 * T + t]; ci.JA[1] = r[3 * T + t];
  ci.Jq[0] = r[6 * T + t]; ci.Jq[1] = r[7 * T + t];
}

__device__ __forceinline__ void load_cell_right_part(const double* r, int T, int t, CellInt& ci) {
  ci.Gr = r[1 * T + t];
  ci.JA[2] = r[4 * T + t]; ci.JA[3] = r[5 * T + t];
  ci.Jq[2] = r[8 * T + t]; ci.Jq[3] = r[9 * T + t];
}

// (2) rows_from_smem: the four vertex rows, streamed so that at most two rows are
// live: a and c are turned into their eliminated form (x = g - P x_l - Q x_r,
// parked in shared memory, over the slot of the cell just consumed) as soon as
// they are complete and applied to b and d (level 0 of the cyclic reduction, own
// part).  This is done before the convergence test is known; on the converged
// pass it is discarded.  Returns the thread's share of ||R||^2.
template <int T>
__device__ __forceinline__ double rows_from_smem(const ArteryCtx& c, const double* sA, const double* sq, double* pa,
                                                 double* pb, double* pc, const double* lo, bool first, double* EA,
                                                 double* Eq, Row& rb, Row& rd) {
  const int tid = threadIdx.x, i0 = 4 * tid;
  CellInt left, own;
  cell_zero(left);
  cell_zero(own);
  double part = 0.0;
  if (tid > 0) load_cell_right_part(lo, T, tid - 1, left);     // last cell of the previous thread
  load_cell_left_part(pa, T, tid, own);
  {
    Row ra;
    row_of<T>(c, i0, first, sA, sq, left, own, EA[0], Eq[0], ra);
    part += ra.r[0] * ra.r[0] + ra.r[1] * ra.r[1];
    load_cell_right_part(pa, T, tid, left);
    load_cell_left_part(pb, T, tid, own);
    row_of<T>(c, i0 + 1, first, sA, sq, left, own, EA[1], Eq[1], rb);
    part += rb.r[0] * rb.r[0] + rb.r[1] * rb.r[1];
    Elim ea = make_elim(ra);
    park(pa, T, tid, ea);
    apply_left(rb, ea);
  }
  {
    Row rc;
    load_cell_right_part(pb, T, tid, left);
    load_cell_left_part(pc, T, tid, own);
    row_of<T>(c, i0 + 2, first, sA, sq, left, own, EA[2], Eq[2], rc);
    part += rc.r[0] * rc.r[0] + rc.r[1] * rc.r[1];
    load_cell_right_part(pc, T, tid, left);
    load_cell_left_part(lo, T, tid, own);
    row_of<T>(c, i0 + 3, first, sA, sq, left, own, EA[3], Eq[3], rd);
    part += rd.r[0] * rd.r[0] + rd.r[1] * rd.r[1];
    Elim ec = make_elim(rc);
    park(pc, T, tid, ec);
    apply_right(rb, ec);
    apply_left(rd, ec);
  }
  return part;
}

// The convergence check after a Newton update needs the residual only (DOLFIN tests ||R|| after every
// update, [3P] NewtonSolver).  Lean pass: the flux / source sums of the four own cells stay in registers,
// one double per thread (the right-vertex sum of its last cell) crosses to the next thread, no Jacobian,
// nothing parked.  Returns the thread's share of ||R||^2.  `ex` is a T-double exchange buffer.
template <int T>
__device__ __forceinline__ double residual_pass(const ArteryCtx& c, const CellCoef& ccu, const double* tab, int S,
                                                const double* sA, const double* sq, double* ex, double* EA,
                                                double* Eq) {
  const int tid = threadIdx.x, i0 = 4 * tid;
  if constexpr (T == 32) {
    // one-warp (ensemble) kernel: ONE copy of the cell routine and ONE of the row routine, each executed
    // four times (16 independent one-warp CTAs per SM compete for the instruction cache, see cells_to_smem);
    // the eight flux / source sums of the thread go through the exchange buffer: ex[(2k) T + t] = Gl,
    // ex[(2k+1) T + t] = Gr of cell k
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const CellInt ci = artery_cell<T, false>(c, ccu, tab, S, sA, sq, i0 + k);
      ex[(2 * k) * T + tid] = ci.Gl;
      ex[(2 * k + 1) * T + tid] = ci.Gr;
    }
    cta_sync<T>();
    double part = 0.0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      const int jm = vix<T>(i > 0 ? i - 1 : i), j0 = vix<T>(i), jp = vix<T>(i < 4 * T - 1 ? i + 1 : i);
      CellInt lf, own;
      lf.Gr = k > 0 ? ex[(2 * k - 1) * T + tid] : (tid > 0 ? ex[7 * T + tid - 1] : 0.0);
      own.Gl = ex[(2 * k) * T + tid];
      double ea = k == 0 ? EA[0] : (k == 1 ? EA[1] : (k == 2 ? EA[2] : EA[3]));
      double eq = k == 0 ? Eq[0] : (k == 1 ? Eq[1] : (k == 2 ? Eq[2] : Eq[3]));
      Row row;
      build_row<false>(c, i, false, sA[jm], sq[jm], sA[j0], sq[j0], sA[jp], sq[jp], lf, own, ea, eq, row);
      part += row.r[0] * row.r[0] + row.r[1] * row.r[1];
    }
    return part;
  } else {
    CellInt cell[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cell[k] = artery_cell<T, false>(c, ccu, tab, S, sA, sq, i0 + k);
    ex[tid] = cell[3].Gr;
    cta_sync<T>();
    CellInt left;
    left.Gr = tid > 0 ? ex[tid - 1] : 0.0;
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      const int jm = vix<T>(i > 0 ? i - 1 : i), j0 = vix<T>(i), jp = vix<T>(i < 4 * T - 1 ? i + 1 : i);
      Row row;
      const CellInt& lf = k == 0 ? left : cell[k > 0 ? k - 1 : 0];      // the cell to the left of vertex i
      build_row<false>(c, i, false, sA[jm], sq[jm], sA[j0], sq[j0], sA[jp], sq[jp], lf, cell[k], EA[k], Eq[k], row);
      part += row.r[0] * row.r[0] + row.r[1] * row.r[1];
    }
    return part;
  }
}

// The Newton iteration of one artery on the iterate U in shared memory ([3P] NewtonSolver on the form of
// ar:191-230; one CTA, four vertices per thread).  `dirichlet_hook` runs once, after the cell integrals of
// the first pass (which need no Dirichlet data): it waits for / fetches the Dirichlet values into `c` and
// must leave the CTA converged (the barrier that follows it publishes the cells).  `expect` = linear solves
// this artery needed in the previous time step.  Returns the number of linear solves; `fl` gets AF_FLAG_* bits.
template <int T, class Hook>
__device__ __forceinline__ int artery_newton(const NetDev& nd, ArteryCtx& c, const CellCoef& ccu, const double* tab,
                                             int S, double* sA, double* sq, double* Lo, double* pa, double* pb,
                                             double* pc, double* red, int expect, uint32_t& fl, Hook&& dirichlet_hook,
                                             long long tl_g, int a, bool wk_split = false) {
  const int tid = threadIdx.x;
  double EA[4], Eq[4];          // explicit part of the residual of the own vertices
  int it = 0;
  double r_first = 0.0;
  fl = 0;
  // How many linear solves this artery needed in the previous time step: the convergence test after
  // solve number `expect` (and later ones) is done by the lean residual-only pass at the bottom of the
  // loop -- it will most likely say "converged", and a full pass there would assemble a Jacobian and
  // eliminate half the rows for nothing.  Before that the test is part of the full pass that the next
  // solve needs anyway.  The prediction only selects which code computes ||R||; the iterates, the tests
  // and the count are those of [3P] NewtonSolver either way.
  bool test_here = true;
  for (;;) {
    const bool first = (it == 0);
    // ---- full pass: cells -> shared memory, then rows (residual + Jacobian) and level 0 of the reduction
    Row rb, rd;
    if (T > 32 && first && wk_split) {
      // Leaf artery of the persistent kernel: its second warp has just run the Windkessel of this step (set_bcs,
      // an:780-786) while the others integrated their cells; the first warp, which would only wait now,
      // integrates the second warp's cells as well.  Same arithmetic on the same inputs, another thread.
      const int w = tid >> 5;
#pragma unroll 1
      for (int rep = 0; rep < (w == 0 ? 2 : (w == 1 ? 0 : 1)); ++rep)
        cells_to_smem<T>(c, ccu, tab, S, sA, sq, pa, pb, pc, Lo, tid + 32 * rep);
    } else {
      cells_to_smem<T>(c, ccu, tab, S, sA, sq, pa, pb, pc, Lo);
    }
    if (first) dirichlet_hook();
    cta_sync<T>();
    double part = rows_from_smem<T>(c, sA, sq, pa, pb, pc, Lo, first, EA, Eq, rb, rd);
    const double r2 = block_sum<T>(part, red);    // its two barriers also publish pa/pc
#ifdef AF_TIMELINE
    if (tid == 0 && first) AF_TL(nd, tl_g, a, 4);
#endif
    if (test_here) {
      // ---- [3P] NewtonSolver::converged: r/r0 < rtol or r < atol ----------
      const double r = r2 > 0.0 ? af_sqrt_pos(r2) : r2;     // branch-free sqrt; 0 and NaN pass through
      if (first) r_first = r;
      if (r < nd.newton_rtol * r_first || r < nd.newton_atol) break;
      if (!(r == r) || r > 1.0e300) { fl = AF_FLAG_NONFINITE | AF_FLAG_ARTERY_NOCONV; break; }
      if (it >= nd.newton_max_it) { fl = AF_FLAG_ARTERY_NOCONV; break; }
    }
    // ---- cyclic reduction, rest of level 0 and level 1 in registers -----------
    if (tid + 1 < T) {
      Elim en = unpark(pa, T, tid + 1);      // vertex a of the next thread
      apply_right(rd, en);
    }
    {
      Elim eb = make_elim(rb);
      park(pb, T, tid, eb);
      apply_left(rd, eb);
    }
    cta_sync<T>();
    if (tid + 1 < T) {
      Elim en = unpark(pb, T, tid + 1);      // vertex b of the next thread
      apply_right(rd, en);
    }
    // ---- the reduced system (one row per thread): parallel cyclic reduction ---
    // The row stays in registers.  In step s = 1, 2, 4, ... every thread turns
    // its row into x_t = g - P x_{t-s} - Q x_{t+s}, publishes (P, Q, g) and
    // eliminates both neighbours at distance s; after log2(T) steps the rows are
    // decoupled.  All warps stay busy and there is no backward sweep: on this
    // latency-bound kernel the shorter serial chain beats the work-optimal
    // cyclic reduction (profiles/README.md).
    double xd[2], xl[2] = {0.0, 0.0};
    {
      double* buf = Lo;                       // 10T doubles of the 14T scratch
#pragma unroll 1
      for (int sdist = 1; sdist < T; sdist *= 2) {
        park(buf, T, tid, make_elim(rd));
        cta_sync<T>();
        const bool more = 2 * sdist < T;       // the last step leaves rows that are never coupled again
        if (tid >= sdist) apply_left(rd, unpark(buf, T, tid - sdist), more);
        if (tid + sdist < T) apply_right(rd, unpark(buf, T, tid + sdist), more);
        cta_sync<T>();
      }
      double inv[4];
      inv2v(rd.dg, inv);
      xd[0] = inv[0] * rd.r[0] + inv[1] * rd.r[1];
      xd[1] = inv[2] * rd.r[0] + inv[3] * rd.r[1];
      buf[tid] = xd[0]; buf[T + tid] = xd[1];
      cta_sync<T>();
      if (tid > 0) { xl[0] = buf[tid - 1]; xl[1] = buf[T + tid - 1]; }
    }
    // ---- back-substitution of the own vertices, U -= dU ----------------------
    {
      double xa[2], xb[2], xc[2];
      back_sub(unpark(pb, T, tid), xl, xd, xb);
      back_sub(unpark(pa, T, tid), xl, xb, xa);
      back_sub(unpark(pc, T, tid), xb, xd, xc);
      sA[tid] -= xa[0]; sq[tid] -= xa[1];
      sA[T + tid] -= xb[0]; sq[T + tid] -= xb[1];
      sA[2 * T + tid] -= xc[0]; sq[2 * T + tid] -= xc[1];
      sA[3 * T + tid] -= xd[0]; sq[3 * T + tid] -= xd[1];
    }
    cta_sync<T>();
    ++it;
#ifdef AF_TIMELINE
    if (tid == 0) AF_TL(nd, tl_g, a, 5);
#endif
    // ---- convergence test after the update ([3P] NewtonSolver::converged), residual only ----------
    test_here = it < expect;          // too early to expect convergence: the next full pass tests
    if (!test_here) {
      const double part2 = residual_pass<T>(c, ccu, tab, S, sA, sq, Lo, EA, Eq);
      const double q2 = block_sum<T>(part2, red);
      const double r = q2 > 0.0 ? af_sqrt_pos(q2) : q2;
      if (r < nd.newton_rtol * r_first || r < nd.newton_atol) break;
      if (!(r == r) || r > 1.0e300) { fl = AF_FLAG_NONFINITE | AF_FLAG_ARTERY_NOCONV; break; }
      if (it >= nd.newton_max_it) { fl = AF_FLAG_ARTERY_NOCONV; break; }
    }
  }
  return it;
}

template <int T, bool FUSED>
__global__ void __launch_bounds__(T, (T == 32 ? AF_T32_BLOCKS : (T <= 256 ? 512 / T : 1)))
k_artery(NetDev nd, int cur, int ba0, int inlet_index, int log_slot, long long gstep) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  __shared__ double bcv[4];
  constexpr int NV = 4 * T;
  const int np = nd.np, ne = nd.Nx, S = nd.S;
  const int tid = threadIdx.x;
  double* sA = sm;
  double* sq = sA + NV;
  double* Lo = sq + NV;       // 10T: exchange buffer of the reduced-system PCR
  double* pa = Lo + 10 * T;   // parked vertex a
  double* pb = pa + 10 * T;
  double* pc = pb + 10 * T;

  // the next boundary kernel may be made resident early; it waits for this grid to finish
  cudaTriggerProgrammaticLaunchCompletion();
  // A failed artery solve of an earlier step ends a single-network run (DOLFIN raises there,
  // ar:244).  The previous artery kernel has completed before this one can start (the boundary
  // kernel between them waits for it before it triggers), so its fail_step is visible here.
  // (not in the one-warp kernel, whose register budget it would break: there the boundary kernel's
  // early return and the host's polling end the run)
  const int ba = blockIdx.x + ba0;
  const int b = ba / nd.na, a = ba % nd.na;
  const bool dataflow = T > 32 && !FUSED && nd.dataflow && gstep >= 0;
#ifdef AF_TIMELINE
  const long long tl_g = (T > 32 && !FUSED && nd.nb == 1) ? gstep : -1;
  if (tid == 0) AF_TL(nd, tl_g, a, 0);
#endif
  if (T > 32 && nd.stop_on_fail && *(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL) {
    if (dataflow && tid == 0) st_release_u32(&nd.f_art[ba], (unsigned)(gstep + 1));   // nobody may be left waiting
    return;
  }
  if (dataflow) {
    // this artery's own previous step (another CTA of the previous artery kernel) must be complete before
    // its Un is read; in the plain scheme the boundary kernel's grid-wide wait guarantees that
    if (tid == 0) wait_flag(&nd.f_art[ba], (unsigned)gstep, nd.fail_step, &nd.flags[b]);
    __syncthreads();
  }
#ifdef AF_TIMELINE
  if (tid == 0) AF_TL(nd, tl_g, a, 1);
#endif
  const Vessel v = nd.ves[ba];
  const size_t off = (size_t)ba * S;
  {
    // U <- Un (ar:158; the Newton initial guess is the previous solution)
    const double* gA = nd.A[cur] + off;
    const double* gq = nd.q[cur] + off;
    // each thread fetches its own four vertices (two 16-byte loads per field: a warp
    // reads 1 KB contiguously); the rows of an artery start 256-byte aligned
    const int i0 = 4 * tid;
    double va[4] = {1.0, 1.0, 1.0, 1.0}, vq[4] = {0.0, 0.0, 0.0, 0.0};     // padding vertices
    if (i0 + 3 < np) {
      const double2 a01 = *reinterpret_cast<const double2*>(gA + i0), a23 = *reinterpret_cast<const double2*>(gA + i0 + 2);
      const double2 q01 = *reinterpret_cast<const double2*>(gq + i0), q23 = *reinterpret_cast<const double2*>(gq + i0 + 2);
      va[0] = a01.x; va[1] = a01.y; va[2] = a23.x; va[3] = a23.y;
      vq[0] = q01.x; vq[1] = q01.y; vq[2] = q23.x; vq[3] = q23.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k < np) { va[k] = gA[i0 + k]; vq[k] = gq[i0 + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sA[k * T + tid] = va[k]; sq[k * T + tid] = vq[k]; }
  }
  ArteryCtx c;
  c.ne = ne; c.np = np;
  c.h = v.dx; c.Kr = -v.src_k;
  c.dth = nd.dt * nd.theta; c.dte = nd.dt * (1.0 - nd.theta);
  c.c1_0 = v.f0 * sqrt(v.A00); c.c1_L = v.fL * sqrt(v.A0L);
  c.root = (a == nd.root_artery); c.leaf = nd.art_leaf[a] >= 0;
  if (FUSED) {
    // set_bcs for THIS artery inside its own CTA (an:767-786): warp 0 solves the
    // junction at the inlet (the artery is a daughter there), warp 1 the junction
    // at the outlet (the artery is its parent) or the Windkessel of a leaf.  A
    // junction is therefore solved by the three CTAs of its vessels from the same
    // inputs (bitwise identical results); its parent's CTA owns the bookkeeping
    // (x -> jx_next, dex, counters).  All of it reads only Un of the `cur` buffers.
    const int w = tid >> 5, lane = tid & 31;
    if (w == 0) {
      if (c.root) {
        if (lane == 0) {
          bcv[0] = nd.bc[ba * 4 + 0];
          bcv[1] = inlet_index >= 0 ? nd.q_ins[(nd.q_ins_per_net ? (size_t)b * nd.Nt : 0) + inlet_index]
                                    : nd.bc[ba * 4 + 1];
        }
      } else {
        const int j = nd.art_jin[a];
        double x[18], dex;
        int solves;
        uint32_t jfl;
        const int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
        const Vessel jves = nd.ves[b * nd.na + ia[lane % 3]];
        bool sing;
        junction_solve_warp(nd, cur, b, j, ia, jves, nd.jx + ((size_t)b * nd.nj + j) * 18, x, dex, solves, jfl, sing);
        if (sing) {             // verbatim dense path (reference's singular branch)
          const double* xg0 = nd.jx + ((size_t)b * nd.nj + j) * 18;
          for (int i = 0; i < 18; ++i) x[i] = xg0[i];
          jfl = 0;
          solves = junction_slow_path(nd, cur, b, ia, dex, x, &jfl);
        }
        if (lane == 0) {
          const bool first_daughter = nd.art_role[a] == 1;
          bcv[0] = first_daughter ? x[12] : x[15];
          bcv[1] = first_daughter ? x[3] : x[6];
        }
      }
    } else if (w == 1) {
      if (c.leaf) {
        WkResult r = leaf_solve_warp(nd, nd.A[cur] + off, nd.q[cur] + off, b, nd.art_leaf[a], ba, v, log_slot);   // writes bc/dex/counters itself
        if (lane == 0) { bcv[2] = r.A_out; bcv[3] = nd.bc[ba * 4 + 3]; }
      } else {
        const int j = nd.art_jout[a];
        double x[18], dex;
        int solves;
        uint32_t jfl;
        const int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
        const Vessel jves = nd.ves[b * nd.na + ia[lane % 3]];
        bool sing;
        junction_solve_warp(nd, cur, b, j, ia, jves, nd.jx + ((size_t)b * nd.nj + j) * 18, x, dex, solves, jfl, sing);
        if (sing) {             // verbatim dense path (reference's singular branch)
          const double* xg0 = nd.jx + ((size_t)b * nd.nj + j) * 18;
          for (int i = 0; i < 18; ++i) x[i] = xg0[i];
          jfl = 0;
          solves = junction_slow_path(nd, cur, b, ia, dex, x, &jfl);
        }
        if (lane == 0) {
          bcv[2] = x[9]; bcv[3] = x[0];
          double* xo = nd.jx_next + ((size_t)b * nd.nj + j) * 18;
          for (int i = 0; i < 18; ++i) xo[i] = x[i];
          nd.dex[ba] = dex;
          nd.j_solves[b * nd.nj + j] = solves;
          if (log_slot >= 0) nd.log_j[(size_t)log_slot * nd.nb * nd.nj + b * nd.nj + j] = solves;
          atomicAdd(&nd.totals[1], (unsigned long long)solves);
          if (jfl) atomicOr(&nd.flags[b], jfl);
        }
      }
    }
    cta_sync<T>();
    c.gAin = bcv[0]; c.gqin = bcv[1]; c.gAout = bcv[2]; c.gqout = bcv[3];
    if (tid == 0) {
      nd.bc[ba * 4 + 0] = bcv[0]; nd.bc[ba * 4 + 1] = bcv[1];
      nd.bc[ba * 4 + 2] = bcv[2]; nd.bc[ba * 4 + 3] = bcv[3];
    }
  }
  CellCoef ccu;
#pragma unroll
  for (int g = 0; g < 3; ++g) { ccu.c1[g] = c.c1_0; ccu.t2[g] = 0.0; ccu.t3[g] = 0.0; }
  const int slot = nd.coef_slot[ba];
  const double* tab = slot >= 0 ? nd.cellcoef + (size_t)slot * 9 * S : nullptr;
  cta_sync<T>();

  // How many linear solves this artery needed in the previous time step (see artery_newton)
  const int prev_its = nd.art_its[ba];
  uint32_t fl = 0;
#ifndef AF_TIMELINE
  const long long tl_g = -1;
#endif
  const int it = artery_newton<T>(nd, c, ccu, tab, S, sA, sq, Lo, pa, pb, pc, red, prev_its > 1 ? prev_its : 1, fl, [&]() {
    if (!FUSED) {
      // Everything up to here (vessel constants, U <- Un, the cell integrals of the
      // first pass) only needs the previous artery kernel's output and has been
      // running under the boundary kernel of this step (programmatic dependent
      // launch); the Dirichlet data it writes is needed from here on.
#ifdef AF_TIMELINE
      if (tid == 0) AF_TL(nd, tl_g, a, 2);
#endif
      if (dataflow) {
        // the Dirichlet data of exactly this artery's two ends, from the mailbox: lane k < 4 polls slot k
        // (A_in, q_in, A_out, q_out; ar:171-189 says which of them this artery has), takes the value, puts
        // the sentinel back and keeps the holder array nd.bc up to date
        if (tid < 4) {
          double val;
          if (tid < 2 ? !c.root : (tid == 2 || !c.leaf)) {
            double* mb_slot = nd.mbox + (size_t)(gstep & 1) * nd.nb * nd.na * 4 + ba * 4 + tid;
            unsigned long long u = mbox_peek(mb_slot);
            for (unsigned spins = 0; u == AF_MBOX_EMPTY; ++spins) {
              if ((spins & 63u) == 63u) {
                if (*(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL) break;
                if (spins > (1u << 23)) { atomicOr(&nd.flags[b], AF_FLAG_SYNC_TIMEOUT); break; }
              }
              u = mbox_peek(mb_slot);
            }
            val = __longlong_as_double((long long)u);
            mbox_put(mb_slot, __longlong_as_double((long long)AF_MBOX_EMPTY));
            nd.bc[ba * 4 + tid] = val;
          } else {
            val = nd.bc[ba * 4 + tid];
          }
          bcv[tid] = val;
        }
        __syncthreads();
      } else {
        cudaGridDependencySynchronize();
      }
#ifdef AF_TIMELINE
      if (tid == 0) AF_TL(nd, tl_g, a, 3);
#endif
      // Dirichlet values (ar:171-189); the root reads the inlet table (an:777, 916)
      if (dataflow) {
        c.gAin = bcv[0]; c.gqin = bcv[1]; c.gAout = bcv[2]; c.gqout = bcv[3];
      } else {
        c.gAin = nd.bc[ba * 4 + 0]; c.gqin = nd.bc[ba * 4 + 1];
        c.gAout = nd.bc[ba * 4 + 2]; c.gqout = nd.bc[ba * 4 + 3];
      }
      if (c.root && inlet_index >= 0) {
        c.gqin = nd.q_ins[(nd.q_ins_per_net ? (size_t)b * nd.Nt : 0) + inlet_index];
        if (tid == 0) nd.bc[ba * 4 + 1] = c.gqin;      // arteries[0].q_in = q_in (an:777): the holder keeps it
      }
    }
  }, tl_g, a);
#ifdef AF_TIMELINE
  if (tid == 0) AF_TL(nd, tl_g, a, 6);
#endif
  // update_solution (ar:247-251): Un <- U, into the other buffer
  {
    double* gA = nd.A[cur ^ 1] + off;
    double* gq = nd.q[cur ^ 1] + off;
    const int i0 = 4 * tid;
    if (i0 + 3 < np) {
      *reinterpret_cast<double2*>(gA + i0) = make_double2(sA[tid], sA[T + tid]);
      *reinterpret_cast<double2*>(gA + i0 + 2) = make_double2(sA[2 * T + tid], sA[3 * T + tid]);
      *reinterpret_cast<double2*>(gq + i0) = make_double2(sq[tid], sq[T + tid]);
      *reinterpret_cast<double2*>(gq + i0 + 2) = make_double2(sq[2 * T + tid], sq[3 * T + tid]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k < np) { gA[i0 + k] = sA[k * T + tid]; gq[i0 + k] = sq[k * T + tid]; }
    }
  }
  if (dataflow) __syncthreads();        // every thread's part of Un is written before thread 0 releases it
  if (tid == 0) {
    nd.art_its[ba] = it;
    // the release is cumulative over what the other threads wrote before the barrier above
    if (dataflow) st_release_u32(&nd.f_art[ba], (unsigned)(gstep + 1));
#ifdef AF_TIMELINE
    AF_TL(nd, tl_g, a, 7);
#endif
    // bookkeeping nobody waits for: after the release, off the critical path of the step
    if (log_slot >= 0) nd.log_art[(size_t)log_slot * nd.nb * nd.na + ba] = it;
    atomicAdd(&nd.totals[0], (unsigned long long)it);
    if (fl) {
      atomicOr(&nd.flags[b], fl);
      if (fl & AF_FLAG_ARTERY_NOCONV) atomicMin(nd.fail_step, (unsigned long long)(gstep >= 0 ? gstep : 0));
    }
  }
}

// ---------------------------------------------------------------------------
// Persistent time loop of a single network: ONE launch of k_artery_p (a CTA per artery) and one of
// k_junction_p (a warp per junction) advance a whole chunk of time steps.  Everything that ordered the
// per-step launches is already in global memory -- f_art[artery] counts the steps an artery has completed,
// the Dirichlet mailbox carries junction results -- so the kernels simply loop.  What it buys
// (scripts/timeline.py on the per-step launches): the CTA of step n+1 of an artery could only start once
// the boundary kernel of step n+1 was completely resident, which in turn needed SMs drained of step-n artery
// CTAs -- the root artery re-started 6.4 us after it had released its previous step, and that re-start, not the
// junction solve (done after 7.2 us), set the 22.4 us period of the root <-> first-junction cycle.  Here
// the iterate stays in shared memory between steps (Un is still written out: the junctions and the dumps read
// it), vessel constants stay in registers, and a leaf artery runs its own Windkessel (its second warp, while
// the others integrate cells) -- with the leaf tasks in a kernel of their own the two kernels would not fit
// the register files together (255 x 32 K + 255 warps x 8 K > 148 x 64 K).  All CTAs of both kernels must be
// co-resident (af_net_create checks; otherwise the per-step launches are used): an artery spins on its
// mailbox, a junction on its arteries' flags; a failed run or a ~1 s wait ends every spin.
// ---------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, (T <= 256 ? 512 / T : 1))
k_artery_p(NetDev nd, int cur, long long g0, int n_steps, int log_slot0, int n_log) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  __shared__ double bcv[4];
  __shared__ double qin_s;
  __shared__ int stop_s;
  constexpr int NV = 4 * T;
  const int np = nd.np, ne = nd.Nx, S = nd.S;
  const int tid = threadIdx.x;
  double* sA = sm;
  double* sq = sA + NV;
  double* Lo = sq + NV;
  double* pa = Lo + 10 * T;
  double* pb = pa + 10 * T;
  double* pc = pb + 10 * T;
  // the junction kernel of this chunk follows in the stream as a programmatic dependent launch
  cudaTriggerProgrammaticLaunchCompletion();
  const int ba = nd.art_perm ? nd.art_perm[blockIdx.x] : blockIdx.x, b = 0, a = ba;
  const Vessel v = nd.ves[ba];
  const size_t off = (size_t)ba * S;
  {
    // U <- Un (ar:158), once per chunk: afterwards the converged iterate of a step IS the next initial guess
    const double* gA = nd.A[cur] + off;
    const double* gq = nd.q[cur] + off;
    const int i0 = 4 * tid;
    double va[4] = {1.0, 1.0, 1.0, 1.0}, vq[4] = {0.0, 0.0, 0.0, 0.0};     // padding vertices
    if (i0 + 3 < np) {
      const double2 a01 = *reinterpret_cast<const double2*>(gA + i0), a23 = *reinterpret_cast<const double2*>(gA + i0 + 2);
      const double2 q01 = *reinterpret_cast<const double2*>(gq + i0), q23 = *reinterpret_cast<const double2*>(gq + i0 + 2);
      va[0] = a01.x; va[1] = a01.y; va[2] = a23.x; va[3] = a23.y;
      vq[0] = q01.x; vq[1] = q01.y; vq[2] = q23.x; vq[3] = q23.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k < np) { va[k] = gA[i0 + k]; vq[k] = gq[i0 + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sA[k * T + tid] = va[k]; sq[k * T + tid] = vq[k]; }
  }
  ArteryCtx c;
  c.ne = ne; c.np = np;
  c.h = v.dx; c.Kr = -v.src_k;
  c.dth = nd.dt * nd.theta; c.dte = nd.dt * (1.0 - nd.theta);
  c.c1_0 = v.f0 * sqrt(v.A00); c.c1_L = v.fL * sqrt(v.A0L);
  c.root = (a == nd.root_artery); c.leaf = nd.art_leaf[a] >= 0;
  CellCoef ccu;
#pragma unroll
  for (int g = 0; g < 3; ++g) { ccu.c1[g] = c.c1_0; ccu.t2[g] = 0.0; ccu.t3[g] = 0.0; }
  const int slot = nd.coef_slot[ba];
  const double* tab = slot >= 0 ? nd.cellcoef + (size_t)slot * 9 * S : nullptr;
  int prev_its = nd.art_its[ba];
  // holder values that no junction / leaf delivers keep what they were (root: A_in; leaf: q_out)
  if (tid < 4) bcv[tid] = nd.bc[ba * 4 + tid];
  for (int s = 0; s < n_steps; ++s, cur ^= 1) {
    const long long g = g0 + s;
    // a failed artery solve ends the run at its step (DOLFIN raises there, ar:244)
    if (tid == 0) stop_s = nd.stop_on_fail && *(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL;
    __syncthreads();
    if (stop_s) break;
    const int log_slot = s < n_log ? log_slot0 + s : -1;
    const int inlet_index = (int)((g % nd.Nt + 1) % nd.Nt);                 // an:916
    if (tid == 0) { AF_TL(nd, g, a, 0); AF_TL(nd, g, a, 1); }
    if (c.root && tid == 0) {
      // the inlet sample of this step, fetched asynchronously into shared memory (consumed after the cells)
      const double* src = nd.q_ins + (nd.q_ins_per_net ? (size_t)b * nd.Nt : 0) + inlet_index;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&qin_s);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
    if (c.leaf && (tid >> 5) == 1) {
      // set_bcs of a leaf (an:780-786): its Windkessel needs nothing but this artery's own Un, which is the
      // iterate this CTA holds in shared memory (U == Un between two steps)
      WkResult r = leaf_solve_warp<T>(nd, sA, sq, b, nd.art_leaf[a], ba, v, log_slot, -1);
      if ((tid & 31) == 0) bcv[2] = r.A_out;
    }
    uint32_t fl = 0;
    const int it = artery_newton<T>(nd, c, ccu, tab, S, sA, sq, Lo, pa, pb, pc, red, prev_its > 1 ? prev_its : 1, fl, [&]() {
      if (tid == 0) AF_TL(nd, g, a, 2);
      if (tid < 4 && (tid < 2 ? !c.root : !c.leaf)) {
        // Dirichlet values from the mailbox (see k_artery)
        double* mb_slot = nd.mbox + (size_t)(g & 1) * nd.nb * nd.na * 4 + ba * 4 + tid;
        unsigned long long u = mbox_peek(mb_slot);
        for (unsigned spins = 0; u == AF_MBOX_EMPTY; ++spins) {
          if ((spins & 63u) == 63u) {
            if (*(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL) break;
            if (spins > (1u << 22)) { atomicOr(&nd.flags[b], AF_FLAG_SYNC_TIMEOUT); break; }
          }
          // a waiting CTA must be quiet: with all 255 artery CTAs resident and polling back to back the junction
          // warps next to them ran 10x slower (scripts/timeline.py: 22 us instead of 2.2 us per junction Newton)
          __nanosleep(AF_POLL_SLEEP_NS);
          u = mbox_peek(mb_slot);
        }
        const double val = __longlong_as_double((long long)u);
        mbox_put(mb_slot, __longlong_as_double((long long)AF_MBOX_EMPTY));
        nd.bc[ba * 4 + tid] = val;
        bcv[tid] = val;
      }
      if (c.root && tid == 0) asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      if (tid == 0) AF_TL(nd, g, a, 3);
      c.gAin = bcv[0]; c.gqin = bcv[1]; c.gAout = bcv[2]; c.gqout = bcv[3];
      if (c.root) {
        c.gqin = qin_s;
        if (tid == 0) nd.bc[ba * 4 + 1] = c.gqin;      // arteries[0].q_in = q_in (an:777): the holder keeps it
      }
    }, g, a, c.leaf);
    if (tid == 0) AF_TL(nd, g, a, 6);
    // update_solution (ar:247-251): Un <- U, into the other buffer
    {
      double* gA = nd.A[cur ^ 1] + off;
      double* gq = nd.q[cur ^ 1] + off;
      const int i0 = 4 * tid;
      if (i0 + 3 < np) {
        *reinterpret_cast<double2*>(gA + i0) = make_double2(sA[tid], sA[T + tid]);
        *reinterpret_cast<double2*>(gA + i0 + 2) = make_double2(sA[2 * T + tid], sA[3 * T + tid]);
        *reinterpret_cast<double2*>(gq + i0) = make_double2(sq[tid], sq[T + tid]);
        *reinterpret_cast<double2*>(gq + i0 + 2) = make_double2(sq[2 * T + tid], sq[3 * T + tid]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i0 + k < np) { gA[i0 + k] = sA[k * T + tid]; gq[i0 + k] = sq[k * T + tid]; }
      }
    }
    __syncthreads();        // every thread's part of Un is written before thread 0 releases it
    if (tid == 0) {
      nd.art_its[ba] = it;
      st_release_u32(&nd.f_art[ba], (unsigned)(g + 1));
      AF_TL(nd, g, a, 7);
      if (log_slot >= 0) nd.log_art[(size_t)log_slot * nd.nb * nd.na + ba] = it;
      atomicAdd(&nd.totals[0], (unsigned long long)it);
      if (fl) {
        atomicOr(&nd.flags[b], fl);
        if (fl & AF_FLAG_ARTERY_NOCONV) atomicMin(nd.fail_step, (unsigned long long)g);
      }
    }
    prev_its = it;
  }
}

// one warp per junction, looping over the steps of the chunk (see k_artery_p)
__global__ void __launch_bounds__(128) k_junction_p(NetDev nd, int cur, long long g0, int n_steps, int log_slot0, int n_log) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= nd.nj) return;
  const int b = 0, v = lane % 3;
  const int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
  const int ba = v == 0 ? ia[0] : (v == 1 ? ia[1] : ia[2]);
  for (int s = 0; s < n_steps; ++s, cur ^= 1) {
    const long long g = g0 + s;
    if (nd.stop_on_fail && *(volatile const unsigned long long*)nd.fail_step != AF_NO_FAIL) break;
    if (lane == 0) AF_TL(nd, g, nd.na + j, 0);
    // the vessel constants are fetched again every step, before the wait: kept in registers across the
    // loop they cost 270 B of spills inside the Newton iteration
    const Vessel ves = nd.ves[ba];
    // the three vessels must have completed step g - 1 (lane v waits for vessel v); the barrier also orders
    // this warp's own warm start x, written by lane 0 at the end of the previous step
    // (every lane polls the flag of "its" vessel lane % 3: no divergent region around the wait)
    wait_flag_warp(&nd.f_art[ba], (unsigned)g, nd.fail_step, &nd.flags[b]);
    __syncwarp();
    if (lane == 0) AF_TL(nd, g, nd.na + j, 1);
    junction_task_warp(nd, cur, b, j, ia, ves, s < n_log ? log_slot0 + s : -1, g);
    __syncwarp();
  }
}

__global__ void k_fill_u64(unsigned long long* p, int n, unsigned long long v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_set_flags(unsigned* p, int n, unsigned v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------
// snapshots (an:924-938) and update_pressure (ar:254-261)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double node_pressure(const Vessel& v, int j, double A) {
  double f, A0;
  if (v.tapered) {
    Coef c = coef_at(v, v.dx * j);
    f = c.f; A0 = c.A0;
  } else {
    f = v.f0; A0 = v.A00;
  }
  return v.p0 + f * (1.0 - sqrt(A0 / A));
}

__global__ void k_dump(NetDev nd, int cur, int fields, double* flow, double* area, double* pressure) {
  int ba = blockIdx.x;
  const Vessel v = nd.ves[ba];
  size_t off = (size_t)ba * nd.S, o = (size_t)ba * nd.np;
  for (int j = threadIdx.x; j < nd.np; j += blockDim.x) {
    double A = nd.A[cur][off + j];
    if (fields & AF_STORE_FLOW) flow[o + j] = nd.q[cur][off + j];
    if (fields & AF_STORE_AREA) area[o + j] = A;
    if (fields & AF_STORE_PRESSURE) pressure[o + j] = node_pressure(v, j, A);
  }
}

__global__ void k_dump_outlet(NetDev nd, int cur, double* out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nd.nb * nd.nl) return;
  int b = t / nd.nl, l = t % nd.nl;
  int ba = b * nd.na + nd.leaf_art[l];
  out[t] = outlet_pressure(nd.ves[ba], nd.A[cur][(size_t)ba * nd.S + nd.Nx]);
}

// ---------------------------------------------------------------------------
// ensemble statistics: moments of the stored leaf-outlet pressures over the members
// ("statistics kernel + one collective")
// op [n_snap][nb][nl] -> part [P][5][n_snap*nl] (count, sum, sum of squares, min, max), P = gridDim.y
// member slices; k_moments_fold adds the P partials in index order (deterministic).
// ---------------------------------------------------------------------------
__global__ void k_outlet_moments(const double* op, int nb, int nl, int n_out, double* part) {
  extern __shared__ double sh[];                 // [4][blockDim.x]
  const int s = blockIdx.x, P = gridDim.y, pi = blockIdx.y;
  const int B = blockDim.x;                      // a multiple of nl: thread t always sees leaf t % nl
  const int m0 = (int)((long long)nb * pi / P), m1 = (int)((long long)nb * (pi + 1) / P);
  const double* src = op + ((size_t)s * nb + m0) * nl;
  const size_t cnt = (size_t)(m1 - m0) * nl;
  double sum = 0.0, ssq = 0.0, lo = 1.0e300, hi = -1.0e300;
  for (size_t i = threadIdx.x; i < cnt; i += B) {
    const double v = src[i];
    sum += v; ssq = fma(v, v, ssq);
    lo = v < lo ? v : lo; hi = v > hi ? v : hi;
  }
  sh[threadIdx.x] = sum; sh[B + threadIdx.x] = ssq; sh[2 * B + threadIdx.x] = lo; sh[3 * B + threadIdx.x] = hi;
  __syncthreads();
  if (threadIdx.x < nl) {
    const int l = threadIdx.x;
    sum = 0.0; ssq = 0.0; lo = 1.0e300; hi = -1.0e300;
    for (int t = l; t < B; t += nl) {            // fixed order
      sum += sh[t]; ssq += sh[B + t];
      lo = sh[2 * B + t] < lo ? sh[2 * B + t] : lo; hi = sh[3 * B + t] > hi ? sh[3 * B + t] : hi;
    }
    double* o = part + (size_t)pi * 5 * n_out + (size_t)s * nl + l;
    o[0] = (double)(m1 - m0); o[n_out] = sum; o[2 * (size_t)n_out] = ssq; o[3 * (size_t)n_out] = lo; o[4 * (size_t)n_out] = hi;
  }
}

// parts [P][5][n] -> out [5][n]; finalize = 0: (count, sum, ssq, min, max), 1: (count, mean, variance, min, max)
__global__ void k_moments_fold(const double* parts, int P, int n, int finalize, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double c = 0.0, sum = 0.0, ssq = 0.0, lo = 1.0e300, hi = -1.0e300;
  for (int p = 0; p < P; ++p) {
    const double* q = parts + (size_t)p * 5 * n + i;
    c += q[0]; sum += q[n]; ssq += q[2 * (size_t)n];
    lo = q[3 * (size_t)n] < lo ? q[3 * (size_t)n] : lo; hi = q[4 * (size_t)n] > hi ? q[4 * (size_t)n] : hi;
  }
  if (finalize) {
    const double mean = sum / c, var = ssq / c - mean * mean;
    sum = mean; ssq = var > 0.0 ? var : 0.0;
  }
  out[i] = c; out[n + i] = sum; out[2 * (size_t)n + i] = ssq; out[3 * (size_t)n + i] = lo; out[4 * (size_t)n + i] = hi;
}

// state <-> dense host layout
__global__ void k_pack(NetDev nd, int cur, double* A, double* q) {
  int ba = blockIdx.x;
  for (int j = threadIdx.x; j < nd.np; j += blockDim.x) {
    A[(size_t)ba * nd.np + j] = nd.A[cur][(size_t)ba * nd.S + j];
    q[(size_t)ba * nd.np + j] = nd.q[cur][(size_t)ba * nd.S + j];
  }
}

__global__ void k_unpack(NetDev nd, int cur, const double* A, const double* q) {
  int ba = blockIdx.x;
  for (int j = threadIdx.x; j < nd.np; j += blockDim.x) {
    nd.A[cur][(size_t)ba * nd.S + j] = A[(size_t)ba * nd.np + j];
    nd.q[cur][(size_t)ba * nd.S + j] = q[(size_t)ba * nd.np + j];
  }
}

// ---------------------------------------------------------------------------
// one-thread evaluation kernels behind the fine-grained ABI entry points
// ---------------------------------------------------------------------------
// mode 0: flux/source at points x with states U -> out[n][3]; 1: u_half (x = x0, x1; U = U0, U1) -> out[2];
// 2: CFL_term at x[0] with U[0..1] -> out[0]
// af_eval_math: the kernel-side elementary functions, for their accuracy test
__global__ void k_eval_math(int which, int n, const double* x, double* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = which == 0 ? af_rcp_fast(x[i]) : af_rsqrt_fast(x[i]);
}

__global__ void k_eval_bc(NetDev nd, int ba, int mode, int n, const double* x, const double* U, double* out) {
  const Vessel v = nd.ves[ba];
  if (mode == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      Coef c = coef_at(v, x[i]);
      out[3 * i] = U[2 * i + 1];
      out[3 * i + 1] = bc_flux2(c, U[2 * i], U[2 * i + 1]);
      out[3 * i + 2] = bc_source2(v, c, U[2 * i], U[2 * i + 1]);
    }
  } else if (threadIdx.x == 0) {
    if (mode == 1) {
      u_half(v, nd.dt, x[0], x[1], U[0], U[1], U[2], U[3], out[0], out[1]);
    } else {
      out[0] = cfl_term(v, coef_at(v, x[0]), U[0], U[1]);
    }
  }
}

__global__ void k_eval_state(NetDev nd, int cur, int ba, int n, const double* x, double* out) {
  size_t off = (size_t)ba * nd.S;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    interp_state(nd.A[cur] + off, nd.q[cur] + off, nd.Nx, nd.ves[ba].dx, x[i], out[2 * i], out[2 * i + 1]);
}

__global__ void k_eval_coef(NetDev nd, int ba, int n, const double* x, double* out) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    Coef c = coef_exact(nd.ves[ba], x[i]);
    out[4 * i] = c.f; out[4 * i + 1] = c.A0; out[4 * i + 2] = c.dfdr; out[4 * i + 3] = c.drdx;
  }
}

// mode 0: cfl dex -> out[0]; 1: residual+jacobian at x (out = y[18], J[324]);
// 2: newton (out = x[18], solves)
__global__ void k_junction_debug(NetDev nd, int cur, int b, int j, int mode, const double* dex3, const double* xin,
                                 int k_max, double tol, double* out) {
  int ia[3] = {nd.j_p[j], nd.j_d1[j], nd.j_d2[j]};
  const double* A = nd.A[cur];
  const double* q = nd.q[cur];
  if (mode == 0) {
    double M = 1.0e300;
    for (int v = 0; v < 3; ++v) {
      int ba = b * nd.na + ia[v];
      size_t off = (size_t)ba * nd.S;
      double m = junction_cfl(nd.ves[ba], v, A + off, q + off, nd.Nx);
      M = m < M ? m : M;
    }
    out[0] = (1.0 + nd.cfl_margin) * nd.dt / M;
    out[1] = M;
    return;
  }
  JunctionCtx c;
  for (int v = 0; v < 3; ++v) {
    int ba = b * nd.na + ia[v];
    size_t off = (size_t)ba * nd.S;
    junction_prepare_vessel(c, v, nd.ves[ba], A + off, q + off, nd.Nx, nd.dt, dex3[v]);
  }
  double x[18];
  for (int i = 0; i < 18; ++i) x[i] = xin[i];
  if (mode == 1) {
    junction_residual(c, x, out);
    junction_jacobian(c, x, out + 18, 18);
  } else {
    uint32_t fl = 0;
    JVessel jv[3];
    for (int v = 0; v < 3; ++v) {
      int ba = b * nd.na + ia[v];
      size_t off = (size_t)ba * nd.S;
      jvessel_prepare(jv[v], v, nd.ves[ba], A + off, q + off, nd.Nx, nd.dt, dex3[v]);
    }
    double xs[18];
    for (int i = 0; i < 18; ++i) xs[i] = x[i];
    int solves = junction_newton_fast(jv, x, k_max, tol, &fl);      // the time loop's solver
    if (solves < 0) {
      for (int i = 0; i < 18; ++i) x[i] = xs[i];
      fl = 0;
      solves = junction_newton(c, x, k_max, tol, &fl);              // verbatim dense path
    }
    for (int i = 0; i < 18; ++i) out[i] = x[i];
    out[18] = (double)solves;
    out[19] = (double)fl;
    if (fl) atomicOr(&nd.flags[b], fl);       // e.g. AF_FLAG_SINGULAR: the reference prints 'Singular' (an:704)
  }
}

__global__ void k_windkessel_debug(NetDev nd, int cur, int b, int l, int k_max, double tol, double* out) {
  int ba = b * nd.na + nd.leaf_art[l];
  size_t off = (size_t)ba * nd.S;
  const double* w = nd.wk + ((size_t)b * nd.nl + l) * 3;
  WkResult r = windkessel(nd.ves[ba], nd.A[cur] + off, nd.q[cur] + off, nd.Nx, nd.dt, w[0], w[1], w[2], k_max, tol,
                          nd.cfl_margin);
  out[0] = r.A_out; out[1] = r.dex; out[2] = (double)r.its;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define AF_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      return fail(e_ == cudaErrorMemoryAllocation ? AF_ENOMEM : AF_ECUDA,                      \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    }                                                                                          \
  } while (0)

// ---- block cache -------------------------------------------------------------------
// Device and pinned-host blocks released by a handle are kept (up to a cap) and handed
// to the next request of exactly the same size.  Parameter sweeps and ensembles create
// and destroy same-shaped networks in a loop; cudaMalloc / cudaFree / cudaHostAlloc of
// the state and the snapshot store cost as much as ~1000 time steps of the order-8 tree
// (scripts/e2e_breakdown.py).  AF_CACHE_MB caps the cached device bytes (default 1024,
// 0 disables both caches); af_release_cached_memory() empties them.
struct BlockCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> dev;    // (device, bytes) -> block
  std::multimap<size_t, void*> host;                   // pinned
  size_t dev_bytes = 0, host_bytes = 0;
  size_t dev_cap = (size_t)1024 << 20, host_cap = (size_t)256 << 20;
  BlockCache() {
    if (const char* env = getenv("AF_CACHE_MB")) {
      long long mb = atoll(env);
      dev_cap = mb > 0 ? (size_t)mb << 20 : 0;
      if (mb <= 0) host_cap = 0;
    }
  }
  void* take_dev(int device, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = dev.find({device, bytes});
    if (it == dev.end()) return nullptr;
    void* p = it->second;
    dev.erase(it);
    dev_bytes -= bytes;
    return p;
  }
  bool give_dev(int device, size_t bytes, void* p) {     // false: not kept, caller frees
    std::lock_guard<std::mutex> lk(mu);
    if (dev_bytes + bytes > dev_cap) return false;
    dev.insert({{device, bytes}, p});
    dev_bytes += bytes;
    return true;
  }
  void* take_host(size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = host.find(bytes);
    if (it == host.end()) return nullptr;
    void* p = it->second;
    host.erase(it);
    host_bytes -= bytes;
    return p;
  }
  bool give_host(size_t bytes, void* p) {
    std::lock_guard<std::mutex> lk(mu);
    if (host_bytes + bytes > host_cap) return false;
    host.insert({bytes, p});
    host_bytes += bytes;
    return true;
  }
};
static BlockCache& cache() {
  static BlockCache* c = new BlockCache();     // never destroyed: no CUDA calls at process exit
  return *c;
}

// Pinned host buffers handed to callers as result arrays (af_net_set_store_host_pinned): they outlive
// the handle and go back to the host cache through af_host_release.
static std::mutex g_pinned_mu;
static std::map<void*, size_t> g_pinned;

static void* pinned_take(size_t bytes) {
  void* p = cache().take_host(bytes);
  if (!p && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned[p] = bytes;
  return p;
}

extern "C" int af_host_release(void* p) {
  if (!p) return AF_OK;
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    auto it = g_pinned.find(p);
    if (it == g_pinned.end()) return AF_EINVAL;
    bytes = it->second;
    g_pinned.erase(it);
  }
  if (!cache().give_host(bytes, p)) cudaFreeHost(p);
  return AF_OK;
}

struct af_net {
  NetDev nd;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int cur = 0;
  std::vector<std::pair<void*, size_t>> allocs;   // device blocks and their sizes in bytes
  // store
  std::vector<uint8_t> mask;
  int first_cycle = 0, max_snap = 0, fields = 0, n_snap = 0, snap_cap = 0;
  double *s_flow = nullptr, *s_area = nullptr, *s_press = nullptr, *s_outlet = nullptr;
  std::vector<int64_t> snap_step;
  std::vector<cudaEvent_t> snap_event;   // recorded after the dump kernels of each snapshot
  cudaStream_t copy_stream = nullptr;    // snapshots are copied out while the time loop keeps running
  // host destinations registered with af_net_set_store_host: af_net_step streams every snapshot
  // through a small pinned ring (device -> pinned on copy_stream, pinned -> user buffer by the
  // enqueueing host thread, which runs ahead of the GPU anyway)
  double *h_flow = nullptr, *h_area = nullptr, *h_press = nullptr, *h_outlet = nullptr;
  static const int kRing = 4;
  double* ring = nullptr;                // pinned, kRing slots of slot_doubles
  size_t slot_doubles = 0;
  cudaEvent_t ring_done[kRing] = {nullptr, nullptr, nullptr, nullptr};
  int ring_snap[kRing] = {-1, -1, -1, -1};   // snapshot held by each slot (-1: free)
  bool direct = false;                   // h_* are pinned buffers of the library: snapshots are copied straight into them
  int n_streamed = 0;                    // snapshots whose device->pinned copy has been enqueued
  int n_delivered = 0;                   // snapshots that have arrived in the user buffers
  // log
  int log_cap = 0, log_n = 0;
  // scratch for the fine-grained calls
  double* scratch = nullptr;      // device, 4096 doubles
  int artery_threads = 256;
  size_t artery_smem = 0;
  int boundary_lanes = 32;
  bool pdl = true;                 // programmatic dependent launch of k_artery (AF_NO_PDL=1 disables)
  bool fused = false;              // af_net_step: set_bcs inside the artery CTAs (single networks)
  bool persist = false;            // af_net_step: persistent kernels over chunks of steps (k_artery_p / k_junction_p)
  static const int kChunk = 64;    // steps per launch pair of the persistent path (failure polling granularity)
  int boundary_block = 64;         // threads per CTA of k_boundary (AF_BOUNDARY_BLOCK overrides)
  // Ensembles: the member range is cut into n_sub contiguous shares, each stepped on its own
  // stream.  Members are independent, so the shares drift freely inside one af_net_step call
  // and the boundary kernel of one share (latency / memory-bound, few warps) fills the machine
  // under the artery kernel of another (FP64 / issue-bound) and under its tail wave.  The
  // shares fork from and join the handle's stream at the ends of the call and around snapshots,
  // so to the caller everything stays ordered on that one stream.
  static const int kMaxSub = 4;
  int n_sub = 1;
  cudaStream_t sub[kMaxSub] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSub] = {nullptr, nullptr, nullptr, nullptr};
  // failure polling (single networks): every kPoll steps the device's fail_step is copied to a pinned
  // word; once it shows a failed artery solve af_net_step stops enqueuing (the launches already queued
  // return at once, see k_artery)
  static const int kPoll = 64;
  static const size_t kFailBytes = 64;
  unsigned long long* fail_h = nullptr;   // pinned (from / back to the host block cache)
  cudaEvent_t fail_ev = nullptr;
  bool fail_poll_pending = false;
  bool failed = false;
  double* moments = nullptr;      // [5][n_snap*nl] outlet-pressure moments over the members (+ partials)
  size_t moments_cap = 0;
  std::vector<uint8_t> pending;   // per artery: U computed by af_artery_solve, not yet assigned to Un
  int n_pending = 0;

  // Set-up arena: while af_net_create runs, small blocks (topology, parameters, holders, counters) are
  // carved out of ONE device block, and what is uploaded into them is first written to a pinned mirror of
  // it; arena_flush() then moves zeros and data with a single copy.  A network description is ~40 small
  // arrays: one allocation + one copy instead of ~100 driver calls (0.3 ms of every ArteryNetwork(param)).
  static const size_t kArenaBytes = (size_t)1 << 20, kArenaBlockMax = (size_t)128 << 10;
  char* arena_d = nullptr;
  char* arena_h = nullptr;        // pinned mirror, only alive during af_net_create
  size_t arena_off = 0;
  bool arena_open = false;
  int arena_begin() {
    if (cache().host_cap == 0) return AF_OK;     // caches disabled (AF_CACHE_MB=0): no pinned mirror per network
    void* q = nullptr;
    {
      double* blk = nullptr;
      int rc = dalloc(&blk, kArenaBytes / sizeof(double), /*clear=*/false);
      if (rc) return rc;
      q = blk;
    }
    arena_h = (char*)cache().take_host(kArenaBytes);
    if (!arena_h && cudaHostAlloc((void**)&arena_h, kArenaBytes, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      arena_h = nullptr;
      return AF_OK;                // no mirror: every block takes the ordinary path
    }
    arena_d = (char*)q;
    arena_off = 0;
    arena_open = true;
    return AF_OK;
  }
  // one copy for everything carved so far (zeros included); no further block comes from the arena
  cudaError_t arena_flush() {
    if (!arena_open) return cudaSuccess;
    arena_open = false;
    return arena_off ? cudaMemcpyAsync(arena_d, arena_h, arena_off, cudaMemcpyHostToDevice, stream) : cudaSuccess;
  }
  void arena_release_mirror() {    // after the stream has been synchronised
    if (arena_h && !cache().give_host(kArenaBytes, arena_h)) cudaFreeHost(arena_h);
    arena_h = nullptr;
    arena_open = false;
  }
  bool in_arena(const void* p) const {
    return arena_d && (const char*)p >= arena_d && (const char*)p < arena_d + kArenaBytes;
  }

  template <typename Tp>
  int dalloc(Tp** p, size_t n, bool clear = true) {
    const size_t bytes = (((n ? n : 1) * sizeof(Tp)) + 511) & ~(size_t)511;
    if (arena_open && bytes <= kArenaBlockMax && arena_off + bytes <= kArenaBytes) {
      memset(arena_h + arena_off, 0, bytes);
      *p = (Tp*)(arena_d + arena_off);
      arena_off += bytes;
      return AF_OK;
    }
    void* q = cache().take_dev(device, bytes);
    if (q) {
      // a recycled block holds the previous owner's data: clear it (a memset at HBM speed is
      // far cheaper than the cudaMalloc it replaces)
      cudaError_t e = clear ? cudaMemsetAsync(q, 0, bytes, stream) : cudaSuccess;
      if (e != cudaSuccess) {
        cudaFree(q);
        return fail(AF_ECUDA, std::string("cudaMemsetAsync of a recycled block: ") + cudaGetErrorString(e));
      }
    } else {
      cudaError_t e = cudaMalloc(&q, bytes);
      if (e == cudaErrorMemoryAllocation) {           // give the cached blocks back and retry once
        cudaGetLastError();
        af_release_cached_memory();
        e = cudaMalloc(&q, bytes);
      }
      if (e != cudaSuccess) return fail(AF_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    allocs.push_back({q, bytes});
    *p = (Tp*)q;
    return AF_OK;
  }
  // the caller guarantees that no enqueued work still uses the block
  static void release_block(int device, void* q, size_t bytes) {
    if (!cache().give_dev(device, bytes, q)) cudaFree(q);
  }
  template <typename Tp>
  void dfree(Tp** p) {
    if (!*p) return;
    for (auto it = allocs.begin(); it != allocs.end(); ++it)
      if (it->first == (void*)*p) {
        release_block(device, it->first, it->second);
        allocs.erase(it);
        break;
      }
    *p = nullptr;
  }
  size_t ring_bytes = 0;
  int ring_alloc(size_t bytes) {
    ring = (double*)cache().take_host(bytes);
    if (!ring) {
      cudaError_t e = cudaHostAlloc((void**)&ring, bytes, cudaHostAllocDefault);
      if (e != cudaSuccess) { ring = nullptr; return fail(AF_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    }
    ring_bytes = bytes;
    return AF_OK;
  }
  void ring_free() {          // the caller guarantees that no copy into the ring is in flight
    if (!ring) return;
    if (!cache().give_host(ring_bytes, ring)) cudaFreeHost(ring);
    ring = nullptr;
    ring_bytes = 0;
  }
  template <typename Tp>
  int upload(Tp** p, const Tp* h, size_t n) {
    int rc = dalloc(p, n);
    if (rc) return rc;
    if (arena_open && in_arena(*p)) {
      if (n) memcpy(arena_h + ((char*)*p - arena_d), h, n * sizeof(Tp));
      return AF_OK;
    }
    AF_CUDA(cudaMemcpyAsync(*p, h, n * sizeof(Tp), cudaMemcpyHostToDevice, stream));
    return AF_OK;
  }
};

extern "C" int af_version(void) { return AF_ABI_VERSION; }
extern "C" const char* af_last_error(void) { return g_err.c_str(); }

extern "C" int af_release_cached_memory(void) {
  BlockCache& c = cache();
  std::lock_guard<std::mutex> lk(c.mu);
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& kv : c.dev) {
    cudaSetDevice(kv.first.first);
    cudaFree(kv.second);
  }
  cudaSetDevice(prev);
  for (auto& kv : c.host) cudaFreeHost(kv.second);
  c.dev.clear(); c.host.clear();
  c.dev_bytes = c.host_bytes = 0;
  return AF_OK;
}

extern "C" void af_desc_defaults(af_desc* d) {
  memset(d, 0, sizeof(*d));
  d->abi_version = AF_ABI_VERSION;
  d->n_networks = 1;
  d->newton_rtol = 1e-9;
  d->newton_atol = 1e-10;
  d->newton_max_it = 1000;
  d->junction_k_max = 30;
  d->junction_tol = 1e-10;
  d->picard_k_max = 100;
  d->picard_tol = 1e-12;
  d->cfl_margin = 0.05;
}

template <int T>
static cudaError_t launch_artery(const af_net* n, int ba0, int count, int inlet_index, int log_slot, bool fused,
                                 cudaStream_t stream, long long gstep) {
  NetDev nd = n->nd;
  if (gstep < 0) nd.stop_on_fail = 0;      // single-phase entry points always run
  if (fused && T >= 64) {
    k_artery<T, (T >= 64)><<<count, T, n->artery_smem, stream>>>(nd, n->cur, ba0, inlet_index, log_slot, gstep);
    return cudaGetLastError();
  }
  // programmatic stream serialisation: may start while the preceding kernel of the
  // stream (k_boundary, which triggers early) is still running; see k_artery
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)count);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = n->artery_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = n->pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_artery<T, false>, nd, n->cur, ba0, inlet_index, log_slot, gstep);
}

// persistent path: can all CTAs of k_artery_p and k_junction_p be resident together?  Registers and shared
// memory per SM from the kernels' attributes; the artery CTAs are placed first (the hardware may spread them
// over the SMs or fill SMs up -- both orders are tried), the junction CTAs must fit into what is left.
template <int T>
static bool persistent_fits(const af_net* n) {
  // (kernel attributes and device properties are queried once per process and device: cudaGetDeviceProperties
  // alone costs more than a millisecond, and networks are created in loops)
  struct Info { bool ok; cudaFuncAttributes fa, fj; cudaDeviceProp prop; };
  static std::mutex mu;
  static std::map<int, Info> cache_;
  Info info;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache_.find(n->device);
    if (it == cache_.end()) {
      Info q;
      q.ok = cudaFuncGetAttributes(&q.fa, k_artery_p<T>) == cudaSuccess &&
             cudaFuncGetAttributes(&q.fj, k_junction_p) == cudaSuccess &&
             cudaGetDeviceProperties(&q.prop, n->device) == cudaSuccess;
      if (!q.ok) cudaGetLastError();
      it = cache_.insert({n->device, q}).first;
    }
    info = it->second;
  }
  if (!info.ok) return false;
  const cudaFuncAttributes &fa = info.fa, &fj = info.fj;
  const cudaDeviceProp& prop = info.prop;
  auto warp_regs = [](int r) { return ((r * 32 + 255) / 256) * 256; };       // allocation unit: 256 registers per warp
  const long long regs_a = (long long)warp_regs(fa.numRegs) * (T / 32), regs_j = (long long)warp_regs(fj.numRegs) * 4;
  const long long smem_a = (long long)n->artery_smem + fa.sharedSizeBytes + 1024, smem_j = fj.sharedSizeBytes + 1024;
  const long long R = prop.regsPerMultiprocessor, M = prop.sharedMemPerMultiprocessor;
  const int nsm = prop.multiProcessorCount, thr = prop.maxThreadsPerMultiProcessor;
  const int na = n->nd.na, ncj = (n->nd.nj + 3) / 4;
  int per_sm = (int)std::min(R / regs_a, M / smem_a);
  per_sm = std::min(per_sm, thr / T);
  if (per_sm < 1 || (long long)per_sm * nsm < na) return false;
  for (int order = 0; order < 2; ++order) {
    std::vector<int> cnt(nsm, 0);
    if (order == 0) for (int i = 0; i < na; ++i) cnt[i % nsm]++;                       // breadth first
    else for (int i = 0, left = na; i < nsm && left > 0; ++i) { cnt[i] = std::min(per_sm, left); left -= cnt[i]; }
    long long room = 0;
    for (int i = 0; i < nsm; ++i) {
      const long long rj = (R - cnt[i] * regs_a) / regs_j, mj = (M - cnt[i] * smem_a) / std::max(smem_j, 1LL);
      const long long tj = (thr - cnt[i] * T) / 128;
      room += std::max(0LL, std::min(rj, std::min(mj, tj)));
    }
    if (room < ncj) return false;
  }
  return true;
}

template <int T>
static cudaError_t launch_persistent(const af_net* n, long long g0, int len, int log_slot0, int n_log) {
  NetDev nd = n->nd;
  k_artery_p<T><<<nd.na, T, n->artery_smem, n->stream>>>(nd, n->cur, g0, len, log_slot0, n_log);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || nd.nj == 0) return e;
  // the junction kernel becomes resident beside the artery kernel (which triggers at its first instruction)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((nd.nj + 3) / 4));
  cfg.blockDim = dim3(128);
  cfg.stream = n->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_junction_p, nd, n->cur, g0, len, log_slot0, n_log);
}

static cudaError_t run_persistent(const af_net* n, long long g0, int len, int log_slot0, int n_log) {
  switch (n->artery_threads) {
    case 64: return launch_persistent<64>(n, g0, len, log_slot0, n_log);
    case 128: return launch_persistent<128>(n, g0, len, log_slot0, n_log);
    case 256: return launch_persistent<256>(n, g0, len, log_slot0, n_log);
    default: return launch_persistent<512>(n, g0, len, log_slot0, n_log);
  }
}

template <int T>
static cudaError_t prepare_artery(af_net* n) {
  cudaError_t e = cudaFuncSetAttribute(k_artery<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)n->artery_smem);
  if (e == cudaSuccess && T >= 64)
    e = cudaFuncSetAttribute(k_artery<T, (T >= 64)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)n->artery_smem);
  if (e == cudaSuccess && T >= 64 && n->nd.dataflow) {
    e = cudaFuncSetAttribute(k_artery_p<(T >= 64 ? T : 64)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)n->artery_smem);
    // persistent kernels over chunks of steps (AF_NO_PERSIST=1: one launch pair per step); every leaf must be
    // a vessel the artery kernel knows as one, and all CTAs of both kernels must fit the device together
    n->persist = e == cudaSuccess && n->pdl && persistent_fits<(T >= 64 ? T : 64)>(n);
    if (const char* env = getenv("AF_NO_PERSIST")) if (atoi(env)) n->persist = false;
    // Kernel profilers (Nsight Compute) run one kernel at a time: the persistent pair, whose kernels wait for each
    // other, would only end on its time-outs.  Under a profiler the per-step launches are used (same device code).
    if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("CUDA_INJECTION64_PATH")) n->persist = false;
    if (n->persist) {
      // CTAs are handed to the SMs round-robin: with na CTAs on nsm SMs (nsm < na <= 2 nsm) the CTAs
      // na - nsm .. nsm - 1 keep an SM to themselves.  Those go to the upper generations of the tree (arteries are
      // numbered generation by generation): the root <-> first-junction cycle sets the period of the whole
      // network, and a CTA that shares neither the FP64 pipe nor the shared-memory bandwidth runs its Newton
      // passes faster (AF_NO_SOLO=1 keeps the identity mapping).
      int nsm = 0;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, n->device);
      const int na = n->nd.na;
      const char* env = getenv("AF_NO_SOLO");
      if (nsm > 0 && na > nsm && na <= 2 * nsm && !(env && atoi(env))) {
        std::vector<int> perm(na);
        const int lo = na - nsm, hi = nsm;          // solo CTAs: [lo, hi)
        int next_upper = 0, next_rest = hi - lo;
        for (int bi = 0; bi < na; ++bi) perm[bi] = (bi >= lo && bi < hi) ? next_upper++ : next_rest++;
        int* dperm = nullptr;
        if (n->upload(&dperm, perm.data(), (size_t)na) == AF_OK) n->nd.art_perm = dperm;
      }
      // Both kernels run once, empty, while nothing else is resident: the first launch of a kernel may have
      // to load its code (lazy module loading) and to grow the device's local-memory pool for its stack frame
      // -- the driver does both only on an idle context, which a junction kernel launched NEXT TO a spinning
      // artery kernel would wait for forever (measured: the arteries' mailbox polls ran into their time-out).
      k_artery_p<(T >= 64 ? T : 64)><<<1, (T >= 64 ? T : 64), n->artery_smem, n->stream>>>(n->nd, 0, 0, 0, -1, 0);
      k_junction_p<<<1, 128, 0, n->stream>>>(n->nd, 0, 0, 0, -1, 0);
      e = cudaGetLastError();
    }
  }
  return e;
}

static cudaError_t run_artery(const af_net* n, int ba0, int count, int inlet_index, int log_slot,
                              bool fused = false, cudaStream_t stream = nullptr, long long gstep = -1) {
  if (!stream) stream = n->stream;
  switch (n->artery_threads) {
    case 32: return launch_artery<32>(n, ba0, count, inlet_index, log_slot, fused, stream, gstep);
    case 64: return launch_artery<64>(n, ba0, count, inlet_index, log_slot, fused, stream, gstep);
    case 128: return launch_artery<128>(n, ba0, count, inlet_index, log_slot, fused, stream, gstep);
    case 256: return launch_artery<256>(n, ba0, count, inlet_index, log_slot, fused, stream, gstep);
    default: return launch_artery<512>(n, ba0, count, inlet_index, log_slot, fused, stream, gstep);
  }
}

static cudaError_t run_boundary(const af_net* n, int log_slot, int b0 = 0, int nbl = -1, cudaStream_t stream = nullptr,
                                bool in_loop = false, long long gstep = -1, bool allow_early = true) {
  NetDev nd = n->nd;
  if (!in_loop) nd.stop_on_fail = 0;
  if (nbl < 0) nbl = n->nd.nb;
  if (!stream) stream = n->stream;
  long long tasks = (long long)nbl * (n->nd.nj + n->nd.nl);
  long long threads = tasks * n->boundary_lanes;
  int block = n->boundary_block;
  long long grid = (threads + block - 1) / block;
  if (grid == 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(block);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (n->pdl && allow_early) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (n->boundary_lanes == 32)
    return cudaLaunchKernelEx(&cfg, k_boundary<true>, nd, n->cur, log_slot, b0, nbl, gstep);
  return cudaLaunchKernelEx(&cfg, k_boundary<false>, nd, n->cur, log_slot, b0, nbl, gstep);
}

extern "C" int af_net_create(const af_desc* d, af_net** out) {
  if (!d || !out) return fail(AF_EINVAL, "null argument");
  *out = nullptr;
  if (d->abi_version != AF_ABI_VERSION) return fail(AF_EINVAL, "af_desc.abi_version mismatch");
  if (d->n_networks < 1 || d->n_arteries < 1 || d->Nx < 2 || d->Nt < 1 || d->n_junctions < 0 || d->n_leaves < 0)
    return fail(AF_EINVAL, "bad sizes in af_desc");
  // one CTA per artery, four vertices per thread, 48 doubles of shared memory per thread:
  // 512 threads (196 KB) is the largest CTA that fits the 227 KB of an SM
  if (d->Nx + 1 > 2048) return fail(AF_EINVAL, "Nx+1 > 2048 vertices per artery is not supported");
  if (!(d->dt > 0.0)) return fail(AF_EINVAL, "af_desc.dt must be positive");
  if (d->root_artery >= d->n_arteries) return fail(AF_EINVAL, "root_artery out of range");
  if (!d->k1 || !d->k2 || !d->k3 || !d->rho || !d->Re || !d->db || !d->p0 || !d->Ru || !d->Rd || !d->L || !d->q0 ||
      !d->q_ins || (d->n_leaves && (!d->R1 || !d->R2 || !d->CT || !d->leaf_artery)) ||
      (d->n_junctions && (!d->junction_parent || !d->junction_d1 || !d->junction_d2)))
    return fail(AF_EINVAL, "null array in af_desc");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(AF_ENODEVICE, "no CUDA device: this library has no CPU path");
  if (d->device < 0 || d->device >= ndev) return fail(AF_EINVAL, "bad device ordinal");
  AF_CUDA(cudaSetDevice(d->device));

  af_net* n = new (std::nothrow) af_net();
  if (!n) return fail(AF_ENOMEM, "host allocation failed");
  n->device = d->device;
  if (d->stream) {
    n->stream = (cudaStream_t)d->stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete n; return fail(AF_ECUDA, cudaGetErrorString(e)); }
    n->own_stream = true;
  }
  NetDev& nd = n->nd;
  memset(&nd, 0, sizeof(nd));
  nd.nb = d->n_networks; nd.na = d->n_arteries; nd.nj = d->n_junctions; nd.nl = d->n_leaves;
  nd.Nx = d->Nx; nd.np = d->Nx + 1; nd.S = (nd.np + 31) & ~31; nd.Nt = d->Nt;
  nd.dt = d->dt; nd.theta = d->theta;
  nd.newton_rtol = d->newton_rtol; nd.newton_atol = d->newton_atol; nd.newton_max_it = d->newton_max_it;
  nd.junction_k_max = d->junction_k_max; nd.junction_tol = d->junction_tol;
  nd.picard_k_max = d->picard_k_max; nd.picard_tol = d->picard_tol; nd.cfl_margin = d->cfl_margin;
  nd.q_ins_per_net = d->q_ins_per_network;
  nd.root_artery = d->root_artery;

  const size_t nba = (size_t)nd.nb * nd.na;
  int rc = AF_OK;
#define AF_TRY(x) do { rc = (x); if (rc) { af_net_destroy(n); return rc; } } while (0)
  AF_TRY(n->arena_begin());
  // topology
  std::vector<int> art_leaf(nd.na, -1);
  for (int l = 0; l < nd.nl; ++l) {
    int a = d->leaf_artery[l];
    if (a < 0 || a >= nd.na) { af_net_destroy(n); return fail(AF_EINVAL, "leaf_artery out of range"); }
    if (art_leaf[a] >= 0) { af_net_destroy(n); return fail(AF_EINVAL, "leaf_artery lists an artery twice"); }
    art_leaf[a] = l;
  }
  for (int j = 0; j < nd.nj; ++j) {
    int ia[3] = {d->junction_parent[j], d->junction_d1[j], d->junction_d2[j]};
    for (int v = 0; v < 3; ++v)
      if (ia[v] < 0 || ia[v] >= nd.na) { af_net_destroy(n); return fail(AF_EINVAL, "junction index out of range"); }
  }
  std::vector<int> art_jin(nd.na, -1), art_role(nd.na, 0), art_jout(nd.na, -1);
  for (int j = 0; j < nd.nj; ++j) {
    const int p = d->junction_parent[j], d1 = d->junction_d1[j], d2 = d->junction_d2[j];
    // every end of an artery carries exactly one boundary condition per step (an:767-786)
    if (p == d1 || p == d2 || d1 == d2 || art_jout[p] >= 0 || art_jin[d1] >= 0 || art_jin[d2] >= 0 ||
        art_leaf[p] >= 0 || d1 == nd.root_artery || d2 == nd.root_artery) {
      af_net_destroy(n);
      return fail(AF_EINVAL, "inconsistent topology: an artery end is claimed by more than one junction/outlet/inlet");
    }
    art_jout[p] = j;
    art_jin[d1] = j; art_role[d1] = 1;
    art_jin[d2] = j; art_role[d2] = 2;
  }
  int *dp, *dd1, *dd2, *dla, *dal, *djin, *drole, *djout;
  AF_TRY(n->upload(&dp, d->junction_parent, nd.nj));
  AF_TRY(n->upload(&dd1, d->junction_d1, nd.nj));
  AF_TRY(n->upload(&dd2, d->junction_d2, nd.nj));
  AF_TRY(n->upload(&dla, d->leaf_artery, nd.nl));
  AF_TRY(n->upload(&dal, art_leaf.data(), nd.na));
  AF_TRY(n->upload(&djin, art_jin.data(), nd.na));
  AF_TRY(n->upload(&drole, art_role.data(), nd.na));
  AF_TRY(n->upload(&djout, art_jout.data(), nd.na));
  nd.j_p = dp; nd.j_d1 = dd1; nd.j_d2 = dd2; nd.leaf_art = dla; nd.art_leaf = dal;
  nd.art_jin = djin; nd.art_role = drole; nd.art_jout = djout;
  // parameters
  double *k1, *k2, *k3, *rho, *Re, *db, *p0, *Ru, *Rd, *L, *q0, *qins;
  AF_TRY(n->upload(&k1, d->k1, nd.nb)); AF_TRY(n->upload(&k2, d->k2, nd.nb)); AF_TRY(n->upload(&k3, d->k3, nd.nb));
  AF_TRY(n->upload(&rho, d->rho, nd.nb)); AF_TRY(n->upload(&Re, d->Re, nd.nb)); AF_TRY(n->upload(&db, d->db, nd.nb));
  AF_TRY(n->upload(&p0, d->p0, nd.nb));
  AF_TRY(n->upload(&Ru, d->Ru, nba)); AF_TRY(n->upload(&Rd, d->Rd, nba)); AF_TRY(n->upload(&L, d->L, nba));
  AF_TRY(n->upload(&q0, d->q0, nba));
  AF_TRY(n->upload(&qins, d->q_ins, (size_t)nd.Nt * (nd.q_ins_per_net ? nd.nb : 1)));
  nd.q_ins = qins;
  std::vector<double> wk((size_t)nd.nb * nd.nl * 3);
  for (size_t i = 0; i < (size_t)nd.nb * nd.nl; ++i) {
    wk[3 * i] = d->R1[i]; wk[3 * i + 1] = d->R2[i]; wk[3 * i + 2] = d->CT[i];
  }
  double* dwk;
  AF_TRY(n->upload(&dwk, wk.data(), wk.size()));
  nd.wk = dwk;
  // coefficient tables only for tapered arteries
  std::vector<int> slot(nba, -1);
  int n_tab = 0;
  for (size_t i = 0; i < nba; ++i)
    if (d->Rd[i] != d->Ru[i]) slot[i] = n_tab++;
  int* dslot;
  AF_TRY(n->upload(&dslot, slot.data(), nba));
  nd.coef_slot = dslot;
  AF_TRY(n->dalloc(&nd.cellcoef, (size_t)n_tab * 9 * nd.S));
  // state
  AF_TRY(n->dalloc(&nd.ves, nba));
  for (int k = 0; k < 2; ++k) {
    AF_TRY(n->dalloc(&nd.A[k], nba * nd.S));
    AF_TRY(n->dalloc(&nd.q[k], nba * nd.S));
  }
  AF_TRY(n->dalloc(&nd.bc, nba * 4));
  AF_TRY(n->dalloc(&nd.dex, nba));
  AF_TRY(n->dalloc(&nd.jx, (size_t)nd.nb * nd.nj * 18));
  AF_TRY(n->dalloc(&nd.jx_next, (size_t)nd.nb * nd.nj * 18));
  AF_TRY(n->dalloc(&nd.art_its, nba));
  AF_TRY(n->dalloc(&nd.j_solves, (size_t)nd.nb * nd.nj));
  AF_TRY(n->dalloc(&nd.picard_its, (size_t)nd.nb * nd.nl));
  AF_TRY(n->dalloc(&nd.totals, 3));
  AF_TRY(n->dalloc(&nd.flags, nd.nb));
  AF_TRY(n->dalloc(&nd.fail_step, 1));
  AF_TRY(n->dalloc(&nd.f_art, nba));
  AF_TRY(n->dalloc(&nd.f_out, nba));
  AF_TRY(n->dalloc(&nd.mbox, nd.nb == 1 ? 2 * nba * 4 : 1));
  AF_TRY(n->dalloc(&n->scratch, 4096));
  {
    cudaError_t ea = n->arena_flush();         // before the memsets below: the mirror holds zeros there
    if (ea != cudaSuccess) { af_net_destroy(n); return fail(AF_ECUDA, std::string("set-up copy: ") + cudaGetErrorString(ea)); }
  }
  cudaMemsetAsync(nd.art_its, 0, nba * sizeof(int), n->stream);
  cudaMemsetAsync(nd.j_solves, 0, (size_t)nd.nb * nd.nj * sizeof(int) + 0, n->stream);
  cudaMemsetAsync(nd.picard_its, 0, (size_t)nd.nb * nd.nl * sizeof(int) + 0, n->stream);
  cudaMemsetAsync(nd.totals, 0, 3 * sizeof(unsigned long long), n->stream);
  cudaMemsetAsync(nd.flags, 0, nd.nb * sizeof(uint32_t), n->stream);
  cudaMemsetAsync(nd.fail_step, 0xff, sizeof(unsigned long long), n->stream);      // AF_NO_FAIL
  nd.stop_on_fail = (nd.nb == 1) ? 1 : 0;      // ensemble members are independent: a failed one is only flagged
  if (nd.stop_on_fail) {
    cudaError_t ef = cudaSuccess;
    n->fail_h = (unsigned long long*)cache().take_host(af_net::kFailBytes);
    if (!n->fail_h) ef = cudaHostAlloc((void**)&n->fail_h, af_net::kFailBytes, cudaHostAllocDefault);
    if (ef == cudaSuccess) ef = cudaEventCreateWithFlags(&n->fail_ev, cudaEventDisableTiming);
    if (ef != cudaSuccess) { af_net_destroy(n); return fail(AF_ECUDA, std::string("failure polling: ") + cudaGetErrorString(ef)); }
    *n->fail_h = AF_NO_FAIL;
  }

  n->pending.assign(nba, 0);
  // launch configuration
  // four vertices per thread, power-of-two thread count (the reduced system has T rows)
  n->artery_threads = 32;
  while (4 * n->artery_threads < nd.np) n->artery_threads *= 2;
  n->artery_smem = (size_t)AF_ARTERY_SMEM_PER_THREAD * n->artery_threads * sizeof(double);
  n->boundary_lanes = ((long long)nd.nb * (nd.nj + nd.nl) >= 148LL * 64) ? 1 : 32;
  n->boundary_block = 64;
  // fused step (one launch per time step) needs two warps per artery for the two ends and is meant for
  // single networks; every non-root artery must hang on a junction and every non-leaf must own one
  n->fused = n->boundary_lanes == 32 && n->artery_threads >= 64 && nd.root_artery == 0;
  // measured on the order-8 tree: 66.9 us/step fused vs 59.3 us two kernels (the boundary code spills under
  // the artery kernel's 128-register cap), so the fused step is opt-in (AF_FUSE=1)
  { const char* env = getenv("AF_FUSE"); if (!(env && atoi(env))) n->fused = false; }
  for (int a = 0; a < nd.na && n->fused; ++a)
    if ((a != 0 && art_jin[a] < 0) || (art_leaf[a] < 0 && art_jout[a] < 0)) n->fused = false;
  if (const char* env = getenv("AF_NO_FUSE")) if (atoi(env)) n->fused = false;
  if (const char* env = getenv("AF_NO_PDL")) if (atoi(env)) n->pdl = false;
  // dataflow flags instead of grid-wide waits: single networks on the warp-per-task boundary kernel and a
  // multi-warp artery kernel (AF_NO_DATAFLOW=1: the plain programmatic-dependent-launch chain)
  nd.dataflow = (nd.nb == 1 && n->boundary_lanes == 32 && n->artery_threads >= 64 && !n->fused) ? 1 : 0;
  if (const char* env = getenv("AF_NO_DATAFLOW")) if (atoi(env)) nd.dataflow = 0;
  // measured on the order-8 tree, whole cycle: with the flags alone 64 threads 23.84 us per step, 256: 23.60;
  // with the Dirichlet mailbox 64: 23.21, 128: 23.08, 256: 23.35 (same box); without the flags the CTA size
  // has no effect (25.5 us)
  if (nd.dataflow) n->boundary_block = 128;
  // sub-streams for ensembles: every share must still fill the machine (148 SMs x 16 one-warp CTAs) twice over
  if (n->boundary_lanes == 1) {
    const long long ctas = (long long)nd.nb * nd.na, fill = 2LL * 148 * AF_T32_BLOCKS;
    n->n_sub = ctas >= 4 * fill ? 4 : (ctas >= 2 * fill ? 2 : 1);
    if (const char* env = getenv("AF_SUBSTREAMS")) {
      int v = atoi(env);
      if (v >= 1 && v <= af_net::kMaxSub) n->n_sub = v;
    }
    if (n->n_sub > nd.nb) n->n_sub = nd.nb;
  }
  if (n->n_sub > 1) {
    cudaError_t es = cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming);
    for (int k = 0; k < n->n_sub && es == cudaSuccess; ++k) {
      es = cudaStreamCreateWithFlags(&n->sub[k], cudaStreamNonBlocking);
      if (es == cudaSuccess) es = cudaEventCreateWithFlags(&n->ev_join[k], cudaEventDisableTiming);
    }
    if (es != cudaSuccess) { af_net_destroy(n); return fail(AF_ECUDA, std::string("sub-streams: ") + cudaGetErrorString(es)); }
  }
#ifdef AF_TIMELINE
  nd.tl = nullptr; nd.tl_steps = 0;
  if (const char* env = getenv("AF_TIMELINE_STEPS")) {
    int v = atoi(env);
    if (v > 0 && nd.nb == 1) {
      nd.tl_steps = v;
      AF_TRY(n->dalloc(&nd.tl, (size_t)v * (nd.na + nd.nj + nd.nl) * 8));
    }
  }
#endif
  if (const char* env = getenv("AF_BOUNDARY_BLOCK")) {
    int v = atoi(env);
    if (v >= 32 && v <= 256 && v % 32 == 0) n->boundary_block = v;
  }
  cudaError_t e;
  switch (n->artery_threads) {
    case 32: e = prepare_artery<32>(n); break;
    case 64: e = prepare_artery<64>(n); break;
    case 128: e = prepare_artery<128>(n); break;
    case 256: e = prepare_artery<256>(n); break;
    default: e = prepare_artery<512>(n); break;
  }
  if (e != cudaSuccess) { af_net_destroy(n); return fail(AF_ECUDA, std::string("artery kernel smem: ") + cudaGetErrorString(e)); }

  k_setup_vessels<<<(unsigned)((nba + 127) / 128), 128, 0, n->stream>>>(nd, Ru, Rd, L, k1, k2, k3, rho, Re, db, p0);
  k_setup_state<<<(unsigned)nba, 128, 0, n->stream>>>(nd, q0);
  if (nd.nj) k_setup_x<<<(unsigned)(((size_t)nd.nb * nd.nj + 127) / 128), 128, 0, n->stream>>>(nd, q0);
  e = cudaStreamSynchronize(n->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { af_net_destroy(n); return fail(AF_ECUDA, std::string("set-up kernels: ") + cudaGetErrorString(e)); }
  n->arena_release_mirror();
#undef AF_TRY
  *out = n;
  return AF_OK;
}

extern "C" void af_net_destroy(af_net* n) {
  if (!n) return;
  cudaSetDevice(n->device);
  if (n->stream) cudaStreamSynchronize(n->stream);
  if (n->copy_stream) cudaStreamSynchronize(n->copy_stream);
  for (int k = 0; k < af_net::kMaxSub; ++k) {
    if (n->sub[k]) { cudaStreamSynchronize(n->sub[k]); cudaStreamDestroy(n->sub[k]); }
    if (n->ev_join[k]) cudaEventDestroy(n->ev_join[k]);
  }
  if (n->ev_fork) cudaEventDestroy(n->ev_fork);
  if (n->fail_ev) cudaEventDestroy(n->fail_ev);
  if (n->fail_h && !cache().give_host(af_net::kFailBytes, n->fail_h)) cudaFreeHost(n->fail_h);
  n->arena_release_mirror();
  for (auto& blk : n->allocs) af_net::release_block(n->device, blk.first, blk.second);
  for (cudaEvent_t ev : n->snap_event) cudaEventDestroy(ev);
  for (int k = 0; k < af_net::kRing; ++k) if (n->ring_done[k]) cudaEventDestroy(n->ring_done[k]);
  n->ring_free();
  if (n->copy_stream) cudaStreamDestroy(n->copy_stream);
  if (n->own_stream && n->stream) cudaStreamDestroy(n->stream);
  delete n;
}

static int check(af_net* n) {
  if (!n) return fail(AF_EINVAL, "null handle");
  AF_CUDA(cudaSetDevice(n->device));
  return AF_OK;
}

#ifdef AF_TIMELINE
// debug builds only: the %globaltimer stamps, [tl_steps][na + nj + nl][8]
extern "C" int af_net_debug_timeline(af_net* n, unsigned long long* out, int64_t count) {
  int rc = check(n);
  if (rc) return rc;
  if (!n->nd.tl) return fail(AF_ESTATE, "no timeline (AF_TIMELINE_STEPS)");
  AF_CUDA(cudaStreamSynchronize(n->stream));
  AF_CUDA(cudaMemcpy(out, n->nd.tl, (size_t)count * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return AF_OK;
}
#endif

extern "C" int af_net_schedule(af_net* n, int32_t* kind) {
  if (!n || !kind) return fail(AF_EINVAL, "null argument");
  *kind = n->fused ? AF_SCHEDULE_FUSED
                   : (n->persist && n->nd.dataflow ? AF_SCHEDULE_PERSISTENT
                                                   : (n->nd.dataflow ? AF_SCHEDULE_DATAFLOW : AF_SCHEDULE_PER_STEP));
  return AF_OK;
}

extern "C" int af_net_sync(af_net* n) {
  int rc = check(n);
  if (rc) return rc;
  AF_CUDA(cudaStreamSynchronize(n->stream));
  AF_CUDA(cudaGetLastError());
  return AF_OK;
}

// ---- streaming snapshots to host buffers -----------------------------------------------
static size_t snapshot_doubles(const af_net* n) {
  const NetDev& nd = n->nd;
  size_t per = (size_t)nd.nb * nd.na * nd.np, pl = (size_t)nd.nb * nd.nl, d = 0;
  if (n->fields & AF_STORE_FLOW) d += per;
  if (n->fields & AF_STORE_AREA) d += per;
  if (n->fields & AF_STORE_PRESSURE) d += per;
  if (n->fields & AF_STORE_OUTLET_PRESSURE) d += pl;
  return d;
}

// pinned slot -> user buffers (host memcpy); frees the slot
static void deliver_slot(af_net* n, int slot) {
  const NetDev& nd = n->nd;
  const size_t per = (size_t)nd.nb * nd.na * nd.np, pl = (size_t)nd.nb * nd.nl;
  const int i = n->ring_snap[slot];
  const double* src = n->ring + (size_t)slot * n->slot_doubles;
  if (n->fields & AF_STORE_FLOW) { if (n->h_flow) memcpy(n->h_flow + per * i, src, per * sizeof(double)); src += per; }
  if (n->fields & AF_STORE_AREA) { if (n->h_area) memcpy(n->h_area + per * i, src, per * sizeof(double)); src += per; }
  if (n->fields & AF_STORE_PRESSURE) { if (n->h_press) memcpy(n->h_press + per * i, src, per * sizeof(double)); src += per; }
  if (n->fields & AF_STORE_OUTLET_PRESSURE) { if (n->h_outlet) memcpy(n->h_outlet + pl * i, src, pl * sizeof(double)); }
  n->ring_snap[slot] = -1;
  n->n_delivered++;
}

// Deliver finished copies; with `block` wait for the oldest one (ring full / final drain).
static int drain_ring(af_net* n, bool block_oldest, bool all) {
  for (;;) {
    int oldest = -1;
    for (int k = 0; k < af_net::kRing; ++k)
      if (n->ring_snap[k] >= 0 && (oldest < 0 || n->ring_snap[k] < n->ring_snap[oldest])) oldest = k;
    if (oldest < 0) return AF_OK;
    cudaError_t q = cudaEventQuery(n->ring_done[oldest]);
    if (q == cudaErrorNotReady) {
      if (!block_oldest && !all) return AF_OK;
      AF_CUDA(cudaEventSynchronize(n->ring_done[oldest]));
    } else if (q != cudaSuccess) {
      return fail(AF_ECUDA, cudaGetErrorString(q));
    }
    deliver_slot(n, oldest);
    block_oldest = false;
  }
}

// enqueue device -> pinned copies of snapshot i (gated on its event) into a free ring slot
static int stream_snapshot(af_net* n, int i) {
  const NetDev& nd = n->nd;
  if (n->direct) {
    const size_t per = (size_t)nd.nb * nd.na * nd.np, pl = (size_t)nd.nb * nd.nl;
    AF_CUDA(cudaStreamWaitEvent(n->copy_stream, n->snap_event[i], 0));
    if ((n->fields & AF_STORE_FLOW) && n->h_flow) AF_CUDA(cudaMemcpyAsync(n->h_flow + per * i, n->s_flow + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if ((n->fields & AF_STORE_AREA) && n->h_area) AF_CUDA(cudaMemcpyAsync(n->h_area + per * i, n->s_area + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if ((n->fields & AF_STORE_PRESSURE) && n->h_press) AF_CUDA(cudaMemcpyAsync(n->h_press + per * i, n->s_press + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if ((n->fields & AF_STORE_OUTLET_PRESSURE) && n->h_outlet) AF_CUDA(cudaMemcpyAsync(n->h_outlet + pl * i, n->s_outlet + pl * i, pl * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    n->n_streamed = n->n_delivered = i + 1;
    return AF_OK;
  }
  int slot = -1;
  for (int k = 0; k < af_net::kRing; ++k)
    if (n->ring_snap[k] < 0) { slot = k; break; }
  if (slot < 0) {
    int rc = drain_ring(n, true, false);
    if (rc) return rc;
    for (int k = 0; k < af_net::kRing; ++k)
      if (n->ring_snap[k] < 0) { slot = k; break; }
  }
  const size_t per = (size_t)nd.nb * nd.na * nd.np, pl = (size_t)nd.nb * nd.nl;
  double* dst = n->ring + (size_t)slot * n->slot_doubles;
  AF_CUDA(cudaStreamWaitEvent(n->copy_stream, n->snap_event[i], 0));
  if (n->fields & AF_STORE_FLOW) { AF_CUDA(cudaMemcpyAsync(dst, n->s_flow + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream)); dst += per; }
  if (n->fields & AF_STORE_AREA) { AF_CUDA(cudaMemcpyAsync(dst, n->s_area + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream)); dst += per; }
  if (n->fields & AF_STORE_PRESSURE) { AF_CUDA(cudaMemcpyAsync(dst, n->s_press + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream)); dst += per; }
  if (n->fields & AF_STORE_OUTLET_PRESSURE) { AF_CUDA(cudaMemcpyAsync(dst, n->s_outlet + pl * i, pl * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream)); }
  AF_CUDA(cudaEventRecord(n->ring_done[slot], n->copy_stream));
  n->ring_snap[slot] = i;
  n->n_streamed = i + 1;
  return AF_OK;
}

static int take_snapshot(af_net* n, int64_t g) {
  if (n->n_snap >= n->max_snap) return AF_OK;
  const NetDev& nd = n->nd;
  size_t per = (size_t)nd.nb * nd.na * nd.np;
  if (n->fields & (AF_STORE_FLOW | AF_STORE_AREA | AF_STORE_PRESSURE)) {
    k_dump<<<nd.nb * nd.na, 128, 0, n->stream>>>(nd, n->cur, n->fields,
                                                  n->s_flow ? n->s_flow + per * n->n_snap : nullptr,
                                                  n->s_area ? n->s_area + per * n->n_snap : nullptr,
                                                  n->s_press ? n->s_press + per * n->n_snap : nullptr);
    AF_CUDA(cudaGetLastError());
  }
  if (n->fields & AF_STORE_OUTLET_PRESSURE) {
    size_t pl = (size_t)nd.nb * nd.nl;
    k_dump_outlet<<<(unsigned)((pl + 127) / 128), 128, 0, n->stream>>>(nd, n->cur, n->s_outlet + pl * n->n_snap);
    AF_CUDA(cudaGetLastError());
  }
  if ((int)n->snap_event.size() <= n->n_snap) {
    cudaEvent_t ev;
    AF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    n->snap_event.push_back(ev);
  }
  AF_CUDA(cudaEventRecord(n->snap_event[n->n_snap], n->stream));
  n->snap_step.push_back(g);
  n->n_snap++;
  if (n->ring || n->direct) return stream_snapshot(n, n->n_snap - 1);
  return AF_OK;
}

// sub-streams: everything enqueued on the handle's stream so far happens before the shares go on ...
static int fork_subs(af_net* n) {
  AF_CUDA(cudaEventRecord(n->ev_fork, n->stream));
  for (int k = 0; k < n->n_sub; ++k) AF_CUDA(cudaStreamWaitEvent(n->sub[k], n->ev_fork, 0));
  return AF_OK;
}
// ... and the handle's stream continues after all of them
static int join_subs(af_net* n) {
  for (int k = 0; k < n->n_sub; ++k) {
    AF_CUDA(cudaEventRecord(n->ev_join[k], n->sub[k]));
    AF_CUDA(cudaStreamWaitEvent(n->stream, n->ev_join[k], 0));
  }
  return AF_OK;
}

// One persistent kernel pair per device at a time (per process): two pairs at once could each hold part of the
// SMs and wait for CTAs that never become resident.  af_net_step takes the lease for the work it enqueues and
// marks the end of that work with an event; a call from ANOTHER handle that finds the event not yet reached
// simply uses the per-step launches (same results), e.g. two networks stepped concurrently from two threads.
// (An event, not a host function at the end of the work: cudaLaunchHostFunc costs ~0.25 ms of stream time.)
struct PersistLease {
  std::mutex mu;
  const af_net* owner = nullptr;     // only ever compared
  cudaEvent_t done = nullptr;        // end of the owner's enqueued persistent work
  bool acquire(const af_net* n) {
    std::lock_guard<std::mutex> lk(mu);
    if (!done && cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      done = nullptr;
      return false;
    }
    if (owner && owner != n) {
      if (cudaEventQuery(done) != cudaSuccess) { cudaGetLastError(); return false; }
      owner = nullptr;
    }
    owner = n;
    return true;
  }
  void mark_end(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(mu);
    if (cudaEventRecord(done, s) != cudaSuccess) cudaGetLastError();
  }
};
static PersistLease g_persist_lease[64];

extern "C" int af_net_step(af_net* n, int64_t n_first, int64_t n_steps) {
  int rc = check(n);
  if (rc) return rc;
  if (n_first < 0 || n_steps < 0) return fail(AF_EINVAL, "negative step range");
  if (n->n_pending) return fail(AF_ESTATE, "arteries with a pending solve: call af_artery_update first");
  const int Nt = n->nd.Nt;
  if (n->nd.dataflow && n_steps > 0) {
    // the flags count time steps: everything up to step n_first - 1 is complete, nothing of step n_first is
    const int nba = n->nd.nb * n->nd.na;
    k_set_flags<<<(nba + 127) / 128, 128, 0, n->stream>>>(n->nd.f_art, nba, (unsigned)n_first);
    k_set_flags<<<(nba + 127) / 128, 128, 0, n->stream>>>(n->nd.f_out, nba, (unsigned)n_first);
    k_fill_u64<<<(8 * nba + 127) / 128, 128, 0, n->stream>>>((unsigned long long*)n->nd.mbox, 8 * nba, AF_MBOX_EMPTY);
    AF_CUDA(cudaGetLastError());
  }
  PersistLease* lease = nullptr;
  if (n->persist && n->nd.dataflow && n_steps > 0 && n->device >= 0 && n->device < 64 &&
      g_persist_lease[n->device].acquire(n))
    lease = &g_persist_lease[n->device];
  if (lease) {
    // the lease ends where the stream has run everything enqueued below (also on an error return)
    struct Release {
      PersistLease* l; cudaStream_t s;
      ~Release() { l->mark_end(s); }
    } release{lease, n->stream};
    // chunks of up to kChunk steps; a chunk ends before a step whose Un is dumped (an:920-938: Un of t_n is
    // stored BEFORE the solve of step n), so snapshots stay ordinary stream-ordered work between two chunks
    const int64_t end = n_first + n_steps;
    int64_t g = n_first;
    while (g < end) {
      if (n->fail_h && (g % af_net::kPoll) == 0) {        // same cadence as the per-step path below
        if (n->fail_poll_pending && cudaEventQuery(n->fail_ev) == cudaSuccess) {
          n->fail_poll_pending = false;
          if (*n->fail_h != AF_NO_FAIL) n->failed = true;
        }
        if (n->failed) break;
        if (!n->fail_poll_pending) {
          AF_CUDA(cudaMemcpyAsync(n->fail_h, n->nd.fail_step, sizeof(unsigned long long), cudaMemcpyDeviceToHost, n->stream));
          AF_CUDA(cudaEventRecord(n->fail_ev, n->stream));
          n->fail_poll_pending = true;
        }
      }
      auto dump_due = [&](int64_t gg) {
        return n->max_snap && gg / Nt >= n->first_cycle && n->mask[(int)(gg % Nt)] && n->n_snap < n->max_snap;
      };
      if (dump_due(g)) {
        rc = take_snapshot(n, g);
        if (rc) return rc;
      }
      if (n->ring && n->n_delivered < n->n_streamed) {
        rc = drain_ring(n, false, false);
        if (rc) return rc;
      }
      int len = 1;                           // chunks end at multiples of kChunk (= kPoll: the failure poll above)
      while ((g + len) % af_net::kChunk != 0 && g + len < end && !dump_due(g + len)) ++len;
      const int n_log = std::max(0, std::min(len, n->log_cap - n->log_n));
      const int log_slot0 = n_log > 0 ? n->log_n : -1;
      n->log_n += n_log;
      AF_CUDA(run_persistent(n, g, len, log_slot0, n_log));
      n->cur ^= (len & 1);
      g += len;
    }
    return AF_OK;
  }
  const bool split = n->n_sub > 1 && n_steps > 0;
  if (split && (rc = fork_subs(n))) return rc;
  for (int64_t g = n_first; g < n_first + n_steps; ++g) {
    int step = (int)(g % Nt);
    int64_t cycle = g / Nt;
    if (n->fail_h && (g % af_net::kPoll) == 0) {
      if (n->fail_poll_pending && cudaEventQuery(n->fail_ev) == cudaSuccess) {
        n->fail_poll_pending = false;
        if (*n->fail_h != AF_NO_FAIL) n->failed = true;
      }
      if (n->failed) break;                  // what is already enqueued drains without doing anything
      if (!n->fail_poll_pending) {
        AF_CUDA(cudaMemcpyAsync(n->fail_h, n->nd.fail_step, sizeof(unsigned long long), cudaMemcpyDeviceToHost, n->stream));
        AF_CUDA(cudaEventRecord(n->fail_ev, n->stream));
        n->fail_poll_pending = true;
      }
    }
    if (n->max_snap && cycle >= n->first_cycle && n->mask[step] && n->n_snap < n->max_snap) {
      // the dump kernels read every member: gather the shares, dump on the handle's stream, release them
      if (split && (rc = join_subs(n))) return rc;
      rc = take_snapshot(n, g);
      if (rc) return rc;
      if (split && (rc = fork_subs(n))) return rc;
    }
    if (n->ring && n->n_delivered < n->n_streamed && (g & 7) == 0) {
      rc = drain_ring(n, false, false);
      if (rc) return rc;
    }
    int log_slot = (n->log_n < n->log_cap) ? n->log_n++ : -1;
    int inlet = (step + 1) % Nt;                       // an:916
    if (n->fused) {
      AF_CUDA(run_artery(n, 0, n->nd.nb * n->nd.na, inlet, log_slot, true, nullptr, g));
      std::swap(n->nd.jx, n->nd.jx_next);
    } else if (split) {
      for (int k = 0; k < n->n_sub; ++k) {
        const int b0 = (int)((long long)n->nd.nb * k / n->n_sub), b1 = (int)((long long)n->nd.nb * (k + 1) / n->n_sub);
        AF_CUDA(run_boundary(n, log_slot, b0, b1 - b0, n->sub[k], true, g));
        AF_CUDA(run_artery(n, b0 * n->nd.na, (b1 - b0) * n->nd.na, inlet, log_slot, false, n->sub[k], g));
      }
    } else {
      // dataflow: the first boundary kernel of a call starts in plain stream order, after the kernels that
      // reset the flags have completed (a dependent launch that never calls cudaGridDependencySynchronize
      // has no visibility guarantee for what its predecessor wrote)
      AF_CUDA(run_boundary(n, log_slot, 0, -1, nullptr, true, g, !(n->nd.dataflow && g == n_first)));
      AF_CUDA(run_artery(n, 0, n->nd.nb * n->nd.na, inlet, log_slot, false, nullptr, g));
    }
    n->cur ^= 1;
  }
  if (split && (rc = join_subs(n))) return rc;
  return AF_OK;
}

extern "C" int af_phase_boundary(af_net* n) {
  int rc = check(n);
  if (rc) return rc;
  AF_CUDA(run_boundary(n, -1));
  return AF_OK;
}

extern "C" int af_phase_arteries(af_net* n, int32_t inlet_index) {
  int rc = check(n);
  if (rc) return rc;
  if (inlet_index < -1 || inlet_index >= n->nd.Nt) return fail(AF_EINVAL, "inlet_index out of range");
  if (n->n_pending) return fail(AF_ESTATE, "arteries with a pending solve: call af_artery_update first");
  AF_CUDA(run_artery(n, 0, n->nd.nb * n->nd.na, inlet_index, -1));
  n->cur ^= 1;
  return AF_OK;
}

extern "C" int af_artery_solve(af_net* n, int32_t network, int32_t artery) {
  int rc = check(n);
  if (rc) return rc;
  if (network < 0 || network >= n->nd.nb || artery < 0 || artery >= n->nd.na)
    return fail(AF_EINVAL, "network/artery index out of range");
  int ba = network * n->nd.na + artery;
  AF_CUDA(run_artery(n, ba, 1, -1, -1));      // U -> the other buffer, Un untouched
  if (!n->pending[ba]) { n->pending[ba] = 1; n->n_pending++; }
  return AF_OK;
}

extern "C" int af_artery_update(af_net* n, int32_t network, int32_t artery) {
  int rc = check(n);
  if (rc) return rc;
  if (network < 0 || network >= n->nd.nb || artery < 0 || artery >= n->nd.na)
    return fail(AF_EINVAL, "network/artery index out of range");
  int ba = network * n->nd.na + artery;
  if (!n->pending[ba]) return AF_OK;          // U == Un already
  size_t off = (size_t)ba * n->nd.S, bytes = (size_t)n->nd.S * sizeof(double);
  AF_CUDA(cudaMemcpyAsync(n->nd.A[n->cur] + off, n->nd.A[n->cur ^ 1] + off, bytes, cudaMemcpyDeviceToDevice, n->stream));
  AF_CUDA(cudaMemcpyAsync(n->nd.q[n->cur] + off, n->nd.q[n->cur ^ 1] + off, bytes, cudaMemcpyDeviceToDevice, n->stream));
  n->pending[ba] = 0;
  n->n_pending--;
  return AF_OK;
}

// member slices per snapshot of the moments reduction (>= 1024 members each)
static int moments_slices(int nb) {
  const int P = nb / 1024;
  return P < 1 ? 1 : (P > 32 ? 32 : P);
}

extern "C" int af_net_set_store(af_net* n, const uint8_t* step_mask, int32_t first_cycle, int32_t max_snapshots,
                                int32_t fields) {
  int rc = check(n);
  if (rc) return rc;
  if (!step_mask || max_snapshots < 1 || !fields) return fail(AF_EINVAL, "bad store configuration");
  const NetDev& nd = n->nd;
  // (re)configuration: buffers are reused when they are large enough
  if (fields != n->fields || max_snapshots > n->snap_cap) {
    size_t per = (size_t)nd.nb * nd.na * nd.np;
    // nothing may still be reading the old buffers (enqueued dumps, snapshot downloads)
    AF_CUDA(cudaStreamSynchronize(n->stream));
    if (n->copy_stream) AF_CUDA(cudaStreamSynchronize(n->copy_stream));
    for (double** b : {&n->s_flow, &n->s_area, &n->s_press, &n->s_outlet}) n->dfree(b);
    n->snap_cap = 0;
    if (fields & AF_STORE_FLOW) { rc = n->dalloc(&n->s_flow, per * max_snapshots); if (rc) return rc; }
    if (fields & AF_STORE_AREA) { rc = n->dalloc(&n->s_area, per * max_snapshots); if (rc) return rc; }
    if (fields & AF_STORE_PRESSURE) { rc = n->dalloc(&n->s_press, per * max_snapshots); if (rc) return rc; }
    if (fields & AF_STORE_OUTLET_PRESSURE) {
      rc = n->dalloc(&n->s_outlet, (size_t)nd.nb * nd.nl * max_snapshots);
      if (rc) return rc;
      // the moments of the stored outlet pressures (af_net_outlet_moments_d) are reduced into this buffer:
      // allocated here, so that the statistics call after the loop only launches kernels
      const size_t need = (size_t)(moments_slices(nd.nb) + 1) * 5 * max_snapshots * nd.nl;
      if (need > n->moments_cap) {
        n->dfree(&n->moments);
        n->moments_cap = 0;
        if ((rc = n->dalloc(&n->moments, need))) return rc;
        n->moments_cap = need;
      }
    }
    n->snap_cap = max_snapshots;
  }
  n->mask.assign(step_mask, step_mask + nd.Nt);
  n->first_cycle = first_cycle;
  n->fields = fields;
  n->max_snap = max_snapshots;
  n->n_snap = 0;
  n->snap_step.clear();
  n->n_streamed = n->n_delivered = 0;
  n->h_flow = n->h_area = n->h_press = n->h_outlet = nullptr;
  if (n->direct && n->copy_stream) AF_CUDA(cudaStreamSynchronize(n->copy_stream));
  n->direct = false;
  if (n->ring) {
    if (n->copy_stream) AF_CUDA(cudaStreamSynchronize(n->copy_stream));   // undelivered copies are dropped
    n->ring_free();
  }
  for (int k = 0; k < af_net::kRing; ++k) n->ring_snap[k] = -1;
  return AF_OK;
}

extern "C" int af_net_set_store_host(af_net* n, double* flow, double* area, double* pressure,
                                     double* outlet_pressure) {
  int rc = check(n);
  if (rc) return rc;
  if (!n->max_snap) return fail(AF_ESTATE, "af_net_set_store first");
  if (n->n_snap) return fail(AF_ESTATE, "snapshots already taken");
  if (!n->copy_stream) AF_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
  n->h_flow = flow; n->h_area = area; n->h_press = pressure; n->h_outlet = outlet_pressure;
  n->slot_doubles = snapshot_doubles(n);
  if (n->ring) {
    AF_CUDA(cudaStreamSynchronize(n->copy_stream));
    n->ring_free();
  }
  if ((rc = n->ring_alloc(n->slot_doubles * af_net::kRing * sizeof(double)))) return rc;
  for (int k = 0; k < af_net::kRing; ++k) {
    if (!n->ring_done[k]) AF_CUDA(cudaEventCreateWithFlags(&n->ring_done[k], cudaEventDisableTiming));
    n->ring_snap[k] = -1;
  }
  n->n_streamed = n->n_delivered = 0;
  return AF_OK;
}

extern "C" int af_net_set_store_host_pinned(af_net* n, double** flow, double** area, double** pressure,
                                            double** outlet_pressure) {
  int rc = check(n);
  if (rc) return rc;
  if (!n->max_snap) return fail(AF_ESTATE, "af_net_set_store first");
  if (n->n_snap) return fail(AF_ESTATE, "snapshots already taken");
  if (!n->copy_stream) AF_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
  if (n->ring) {
    AF_CUDA(cudaStreamSynchronize(n->copy_stream));
    n->ring_free();
  }
  const NetDev& nd = n->nd;
  const size_t per = (size_t)nd.nb * nd.na * nd.np * n->max_snap * sizeof(double);
  const size_t pl = (size_t)nd.nb * nd.nl * n->max_snap * sizeof(double);
  struct { double** out; int bit; size_t bytes; double** dst; } f[4] = {
      {flow, AF_STORE_FLOW, per, &n->h_flow}, {area, AF_STORE_AREA, per, &n->h_area},
      {pressure, AF_STORE_PRESSURE, per, &n->h_press}, {outlet_pressure, AF_STORE_OUTLET_PRESSURE, pl, &n->h_outlet}};
  n->h_flow = n->h_area = n->h_press = n->h_outlet = nullptr;
  for (auto& e : f) {
    if (e.out) *e.out = nullptr;
    if (!(n->fields & e.bit) || !e.out) continue;
    void* p = pinned_take(e.bytes);
    if (!p) {
      for (auto& u : f) if (u.out && *u.out) { af_host_release(*u.out); *u.out = nullptr; *u.dst = nullptr; }
      return fail(AF_ENOMEM, "cudaHostAlloc of the result arrays failed");
    }
    *e.out = (double*)p;
    *e.dst = (double*)p;
  }
  n->direct = true;
  n->n_streamed = n->n_delivered = 0;
  return AF_OK;
}

extern "C" int af_net_snapshot_count(af_net* n, int32_t* count) {
  if (!n || !count) return fail(AF_EINVAL, "null argument");
  *count = n->n_snap;
  return AF_OK;
}

extern "C" int af_net_fetch_snapshots(af_net* n, double* flow, double* area, double* pressure,
                                      double* outlet_pressure, int64_t* step) {
  int rc = check(n);
  if (rc) return rc;
  const NetDev& nd = n->nd;
  if (flow && !n->s_flow) return fail(AF_ESTATE, "flow not stored");
  if (area && !n->s_area) return fail(AF_ESTATE, "area not stored");
  if (pressure && !n->s_press) return fail(AF_ESTATE, "pressure not stored");
  if (outlet_pressure && !n->s_outlet) return fail(AF_ESTATE, "outlet pressure not stored");
  if (n->direct) {
    if ((flow && flow != n->h_flow) || (area && area != n->h_area) || (pressure && pressure != n->h_press) ||
        (outlet_pressure && outlet_pressure != n->h_outlet))
      return fail(AF_EINVAL, "pinned store: pass the buffers obtained from af_net_set_store_host_pinned (or NULL)");
    AF_CUDA(cudaStreamSynchronize(n->copy_stream));
    if (step) for (int i = 0; i < n->n_snap; ++i) step[i] = n->snap_step[i];
    return AF_OK;
  }
  if (n->ring) {
    // streaming mode: everything is already on its way to the registered host buffers
    if ((flow && flow != n->h_flow) || (area && area != n->h_area) || (pressure && pressure != n->h_press) ||
        (outlet_pressure && outlet_pressure != n->h_outlet))
      return fail(AF_EINVAL, "streaming store: pass the buffers registered with af_net_set_store_host (or NULL)");
    rc = drain_ring(n, false, true);
    if (rc) return rc;
    if (step) for (int i = 0; i < n->n_snap; ++i) step[i] = n->snap_step[i];
    return AF_OK;
  }
  if (!n->copy_stream) AF_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
  // Snapshot by snapshot on a second stream, each copy gated by the event recorded
  // after its dump kernels: the downloads overlap the rest of the enqueued time loop.
  const size_t per = (size_t)nd.nb * nd.na * nd.np, pl = (size_t)nd.nb * nd.nl;
  for (int i = 0; i < n->n_snap; ++i) {
    AF_CUDA(cudaStreamWaitEvent(n->copy_stream, n->snap_event[i], 0));
    if (flow) AF_CUDA(cudaMemcpyAsync(flow + per * i, n->s_flow + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if (area) AF_CUDA(cudaMemcpyAsync(area + per * i, n->s_area + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if (pressure) AF_CUDA(cudaMemcpyAsync(pressure + per * i, n->s_press + per * i, per * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
    if (outlet_pressure)
      AF_CUDA(cudaMemcpyAsync(outlet_pressure + pl * i, n->s_outlet + pl * i, pl * sizeof(double), cudaMemcpyDeviceToHost, n->copy_stream));
  }
  AF_CUDA(cudaStreamSynchronize(n->copy_stream));
  if (step) for (int i = 0; i < n->n_snap; ++i) step[i] = n->snap_step[i];
  return AF_OK;
}

extern "C" int af_net_outlet_pressure_d(af_net* n, double** ptr_d) {
  if (!n || !ptr_d) return fail(AF_EINVAL, "null argument");
  if (!n->s_outlet) return fail(AF_ESTATE, "outlet pressure not stored");
  *ptr_d = n->s_outlet;
  return AF_OK;
}

extern "C" int af_net_outlet_moments_d(af_net* n, double** moments_d, int32_t* n_snapshots) {
  int rc = check(n);
  if (rc) return rc;
  if (!moments_d) return fail(AF_EINVAL, "null argument");
  if (!n->s_outlet) return fail(AF_ESTATE, "outlet pressure not stored");
  const NetDev& nd = n->nd;
  if (nd.nl < 1 || nd.nl > 1024) return fail(AF_EINVAL, "1..1024 leaves supported");
  const int ns = n->n_snap, n_out = ns * nd.nl;
  if (n_snapshots) *n_snapshots = ns;
  if (ns == 0) { *moments_d = nullptr; return AF_OK; }
  const int P = moments_slices(nd.nb);
  const size_t need = (size_t)(P + 1) * 5 * n_out;
  if (need > n->moments_cap) {
    AF_CUDA(cudaStreamSynchronize(n->stream));
    n->dfree(&n->moments);
    n->moments_cap = 0;
    if ((rc = n->dalloc(&n->moments, need))) return rc;
    n->moments_cap = need;
  }
  const int B = nd.nl * (256 / nd.nl > 0 ? 256 / nd.nl : 1);
  double* parts = n->moments + (size_t)5 * n_out;
  k_outlet_moments<<<dim3(ns, P), B, 4 * B * sizeof(double), n->stream>>>(n->s_outlet, nd.nb, nd.nl, n_out, parts);
  AF_CUDA(cudaGetLastError());
  k_moments_fold<<<(n_out + 127) / 128, 128, 0, n->stream>>>(parts, P, n_out, 0, n->moments);
  AF_CUDA(cudaGetLastError());
  *moments_d = n->moments;
  return AF_OK;
}

extern "C" int af_moments_combine_d(int32_t device, void* stream, const double* parts_d, int32_t n_parts, int32_t n,
                                    double* out_d) {
  if (!parts_d || !out_d || n_parts < 1 || n < 1) return fail(AF_EINVAL, "bad argument");
  AF_CUDA(cudaSetDevice(device));
  k_moments_fold<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(parts_d, n_parts, n, 1, out_d);
  AF_CUDA(cudaGetLastError());
  return AF_OK;
}

// dense <-> padded copies go through a temporary device buffer
static int with_dense(af_net* n, double* hA, double* hq, const double* cA, const double* cq, bool to_host) {
  const NetDev& nd = n->nd;
  size_t cnt = (size_t)nd.nb * nd.na * nd.np;
  double *dA = nullptr, *dq = nullptr;       // from the handle's block cache, returned below
  int rc = n->dalloc(&dA, cnt);
  if (rc) return rc;
  if ((rc = n->dalloc(&dq, cnt))) { n->dfree(&dA); return rc; }
  if (to_host) {
    k_pack<<<nd.nb * nd.na, 128, 0, n->stream>>>(nd, n->cur, dA, dq);
    cudaMemcpyAsync(hA, dA, cnt * sizeof(double), cudaMemcpyDeviceToHost, n->stream);
    cudaMemcpyAsync(hq, dq, cnt * sizeof(double), cudaMemcpyDeviceToHost, n->stream);
  } else {
    cudaMemcpyAsync(dA, cA, cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream);
    cudaMemcpyAsync(dq, cq, cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream);
    k_unpack<<<nd.nb * nd.na, 128, 0, n->stream>>>(nd, n->cur, dA, dq);
    // the Newton initial guess U is Un (ar:158), nothing else to reset
  }
  cudaError_t e = cudaStreamSynchronize(n->stream);
  n->dfree(&dA); n->dfree(&dq);
  if (e != cudaSuccess) return fail(AF_ECUDA, cudaGetErrorString(e));
  return AF_OK;
}

extern "C" int af_net_get_state(af_net* n, double* A, double* q) {
  int rc = check(n);
  if (rc) return rc;
  if (!A || !q) return fail(AF_EINVAL, "null argument");
  return with_dense(n, A, q, nullptr, nullptr, true);
}

extern "C" int af_net_set_state(af_net* n, const double* A, const double* q) {
  int rc = check(n);
  if (rc) return rc;
  if (!A || !q) return fail(AF_EINVAL, "null argument");
  std::fill(n->pending.begin(), n->pending.end(), (uint8_t)0);   // U.assign(Un) (ar:158)
  n->n_pending = 0;
  return with_dense(n, nullptr, nullptr, A, q, false);
}

extern "C" int af_net_get_pressure(af_net* n, double* p) {
  int rc = check(n);
  if (rc) return rc;
  if (!p) return fail(AF_EINVAL, "null argument");
  const NetDev& nd = n->nd;
  size_t cnt = (size_t)nd.nb * nd.na * nd.np;
  double* dp = nullptr;
  if ((rc = n->dalloc(&dp, cnt))) return rc;
  k_dump<<<nd.nb * nd.na, 128, 0, n->stream>>>(nd, n->cur, AF_STORE_PRESSURE, nullptr, nullptr, dp);
  cudaMemcpyAsync(p, dp, cnt * sizeof(double), cudaMemcpyDeviceToHost, n->stream);
  cudaError_t e = cudaStreamSynchronize(n->stream);
  n->dfree(&dp);
  if (e != cudaSuccess) return fail(AF_ECUDA, cudaGetErrorString(e));
  return AF_OK;
}

template <typename Tp>
static int d2h(af_net* n, Tp* host, const Tp* dev, size_t cnt) {
  if (!host) return AF_OK;
  AF_CUDA(cudaMemcpyAsync(host, dev, cnt * sizeof(Tp), cudaMemcpyDeviceToHost, n->stream));
  AF_CUDA(cudaStreamSynchronize(n->stream));
  return AF_OK;
}

template <typename Tp>
static int h2d(af_net* n, Tp* dev, const Tp* host, size_t cnt) {
  if (!host) return fail(AF_EINVAL, "null argument");
  AF_CUDA(cudaMemcpyAsync(dev, host, cnt * sizeof(Tp), cudaMemcpyHostToDevice, n->stream));
  AF_CUDA(cudaStreamSynchronize(n->stream));
  return AF_OK;
}

extern "C" int af_net_get_bc(af_net* n, double* bc) {
  int rc = check(n);
  return rc ? rc : d2h(n, bc, n->nd.bc, (size_t)n->nd.nb * n->nd.na * 4);
}
extern "C" int af_net_set_bc(af_net* n, const double* bc) {
  int rc = check(n);
  return rc ? rc : h2d(n, n->nd.bc, bc, (size_t)n->nd.nb * n->nd.na * 4);
}
extern "C" int af_net_get_dex(af_net* n, double* dex) {
  int rc = check(n);
  return rc ? rc : d2h(n, dex, n->nd.dex, (size_t)n->nd.nb * n->nd.na);
}
extern "C" int af_net_set_dex(af_net* n, const double* dex) {
  int rc = check(n);
  return rc ? rc : h2d(n, n->nd.dex, dex, (size_t)n->nd.nb * n->nd.na);
}
extern "C" int af_net_get_junction_x(af_net* n, double* x) {
  int rc = check(n);
  return rc ? rc : d2h(n, x, n->nd.jx, (size_t)n->nd.nb * n->nd.nj * 18);
}
extern "C" int af_net_set_junction_x(af_net* n, const double* x) {
  int rc = check(n);
  return rc ? rc : h2d(n, n->nd.jx, x, (size_t)n->nd.nb * n->nd.nj * 18);
}
extern "C" int af_net_get_counters(af_net* n, int32_t* artery_its, int32_t* junction_solves, int32_t* picard_its) {
  int rc = check(n);
  if (rc) return rc;
  const NetDev& nd = n->nd;
  if ((rc = d2h(n, artery_its, nd.art_its, (size_t)nd.nb * nd.na))) return rc;
  if ((rc = d2h(n, junction_solves, nd.j_solves, (size_t)nd.nb * nd.nj))) return rc;
  return d2h(n, picard_its, nd.picard_its, (size_t)nd.nb * nd.nl);
}
extern "C" int af_net_get_totals(af_net* n, int64_t totals[3]) {
  int rc = check(n);
  return rc ? rc : d2h(n, (unsigned long long*)totals, n->nd.totals, 3);
}
extern "C" int af_net_get_flags(af_net* n, uint32_t* flags) {
  int rc = check(n);
  return rc ? rc : d2h(n, flags, n->nd.flags, (size_t)n->nd.nb);
}

extern "C" int af_net_clear_flags(af_net* n) {
  int rc = check(n);
  if (rc) return rc;
  AF_CUDA(cudaStreamSynchronize(n->stream));
  AF_CUDA(cudaMemsetAsync(n->nd.flags, 0, n->nd.nb * sizeof(uint32_t), n->stream));
  AF_CUDA(cudaMemsetAsync(n->nd.fail_step, 0xff, sizeof(unsigned long long), n->stream));
  if (n->fail_h) *n->fail_h = AF_NO_FAIL;
  n->failed = false;
  n->fail_poll_pending = false;
  return AF_OK;
}

extern "C" int af_net_failed_step(af_net* n, int64_t* step) {
  int rc = check(n);
  if (rc) return rc;
  if (!step) return fail(AF_EINVAL, "null argument");
  unsigned long long v = AF_NO_FAIL;
  if ((rc = d2h(n, &v, n->nd.fail_step, 1))) return rc;
  *step = v == AF_NO_FAIL ? -1 : (int64_t)v;
  return AF_OK;
}

extern "C" int af_net_set_counter_log(af_net* n, int32_t capacity) {
  int rc = check(n);
  if (rc) return rc;
  if (n->log_cap) return fail(AF_ESTATE, "counter log already configured");
  if (capacity < 1) return fail(AF_EINVAL, "capacity < 1");
  NetDev& nd = n->nd;
  if ((rc = n->dalloc(&nd.log_art, (size_t)capacity * nd.nb * nd.na))) return rc;
  if ((rc = n->dalloc(&nd.log_j, (size_t)capacity * nd.nb * nd.nj))) return rc;
  if ((rc = n->dalloc(&nd.log_p, (size_t)capacity * nd.nb * nd.nl))) return rc;
  n->log_cap = capacity;
  n->log_n = 0;
  return AF_OK;
}

extern "C" int af_net_fetch_counter_log(af_net* n, int32_t* n_logged, int32_t* artery_its, int32_t* junction_solves,
                                        int32_t* picard_its) {
  int rc = check(n);
  if (rc) return rc;
  const NetDev& nd = n->nd;
  if (n_logged) *n_logged = n->log_n;
  if ((rc = d2h(n, artery_its, nd.log_art, (size_t)n->log_n * nd.nb * nd.na))) return rc;
  if ((rc = d2h(n, junction_solves, nd.log_j, (size_t)n->log_n * nd.nb * nd.nj))) return rc;
  return d2h(n, picard_its, nd.log_p, (size_t)n->log_n * nd.nb * nd.nl);
}

// ---- fine-grained evaluations ------------------------------------------------
static int check_ba(af_net* n, int network, int artery) {
  if (network < 0 || network >= n->nd.nb || artery < 0 || artery >= n->nd.na)
    return fail(AF_EINVAL, "network/artery index out of range");
  return AF_OK;
}

extern "C" int af_eval_state(af_net* n, int32_t network, int32_t artery, int32_t which, int32_t cnt, const double* x,
                             double* out) {
  int rc = check(n);
  if (rc) return rc;
  if ((rc = check_ba(n, network, artery))) return rc;
  if (cnt < 0 || cnt > 1024 || !x || !out) return fail(AF_EINVAL, "bad point list (max 1024)");
  int ba = network * n->nd.na + artery;
  int buf = (which == 1 && n->pending[ba]) ? (n->cur ^ 1) : n->cur;   // U == Un unless a solve is pending
  AF_CUDA(cudaMemcpyAsync(n->scratch, x, cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  k_eval_state<<<1, 64, 0, n->stream>>>(n->nd, buf, ba, cnt, n->scratch, n->scratch + 1024);
  AF_CUDA(cudaGetLastError());
  return d2h(n, out, n->scratch + 1024, (size_t)cnt * 2);
}

extern "C" int af_eval_coefficients(af_net* n, int32_t network, int32_t artery, int32_t cnt, const double* x,
                                    double* out) {
  int rc = check(n);
  if (rc) return rc;
  if ((rc = check_ba(n, network, artery))) return rc;
  if (cnt < 0 || cnt > 512 || !x || !out) return fail(AF_EINVAL, "bad point list (max 512)");
  AF_CUDA(cudaMemcpyAsync(n->scratch, x, cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  k_eval_coef<<<1, 64, 0, n->stream>>>(n->nd, network * n->nd.na + artery, cnt, n->scratch, n->scratch + 1024);
  AF_CUDA(cudaGetLastError());
  return d2h(n, out, n->scratch + 1024, (size_t)cnt * 4);
}

extern "C" int af_eval_flux_source(af_net* n, int32_t network, int32_t artery, int32_t cnt, const double* x,
                                   const double* U, double* out) {
  int rc = check(n);
  if (rc) return rc;
  if ((rc = check_ba(n, network, artery))) return rc;
  if (cnt < 0 || cnt > 256 || !x || !U || !out) return fail(AF_EINVAL, "bad point list (max 256)");
  double* s = n->scratch;
  AF_CUDA(cudaMemcpyAsync(s, x, cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  AF_CUDA(cudaMemcpyAsync(s + 256, U, 2 * cnt * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  k_eval_bc<<<1, 64, 0, n->stream>>>(n->nd, network * n->nd.na + artery, 0, cnt, s, s + 256, s + 1024);
  AF_CUDA(cudaGetLastError());
  return d2h(n, out, s + 1024, (size_t)cnt * 3);
}

extern "C" int af_eval_u_half(af_net* n, int32_t network, int32_t artery, double x0, double x1, const double U0[2],
                              const double U1[2], double out[2]) {
  int rc = check(n);
  if (rc) return rc;
  if ((rc = check_ba(n, network, artery))) return rc;
  if (!U0 || !U1 || !out) return fail(AF_EINVAL, "null argument");
  double h[6] = {x0, x1, U0[0], U0[1], U1[0], U1[1]};
  double* s = n->scratch;
  AF_CUDA(cudaMemcpyAsync(s, h, sizeof(h), cudaMemcpyHostToDevice, n->stream));
  k_eval_bc<<<1, 32, 0, n->stream>>>(n->nd, network * n->nd.na + artery, 1, 1, s, s + 2, s + 1024);
  AF_CUDA(cudaGetLastError());
  return d2h(n, out, s + 1024, 2);
}

extern "C" int af_eval_cfl_term(af_net* n, int32_t network, int32_t artery, double x, double A, double q, double* M) {
  int rc = check(n);
  if (rc) return rc;
  if ((rc = check_ba(n, network, artery))) return rc;
  if (!M) return fail(AF_EINVAL, "null argument");
  double h[3] = {x, A, q};
  double* s = n->scratch;
  AF_CUDA(cudaMemcpyAsync(s, h, sizeof(h), cudaMemcpyHostToDevice, n->stream));
  k_eval_bc<<<1, 32, 0, n->stream>>>(n->nd, network * n->nd.na + artery, 2, 1, s, s + 1, s + 1024);
  AF_CUDA(cudaGetLastError());
  return d2h(n, M, s + 1024, 1);
}

extern "C" int af_eval_math(int32_t device, int32_t which, int32_t cnt, const double* x, double* out) {
  if (!x || !out || cnt < 0 || which < 0 || which > 1) return fail(AF_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(AF_ENODEVICE, "no CUDA device");
  AF_CUDA(cudaSetDevice(device));
  if (cnt == 0) return AF_OK;
  double* d = nullptr;
  AF_CUDA(cudaMalloc(&d, 2 * (size_t)cnt * sizeof(double)));
  cudaMemcpy(d, x, cnt * sizeof(double), cudaMemcpyHostToDevice);
  k_eval_math<<<148, 256>>>(which, cnt, d, d + cnt);
  cudaError_t e = cudaMemcpy(out, d + cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(AF_ECUDA, cudaGetErrorString(e));
  return AF_OK;
}

static int junction_call(af_net* n, int network, int junction, int mode, const double* dex3, const double* x,
                         int k_max, double tol, double* host_out, size_t n_out) {
  int rc = check(n);
  if (rc) return rc;
  if (network < 0 || network >= n->nd.nb || junction < 0 || junction >= n->nd.nj)
    return fail(AF_EINVAL, "network/junction index out of range");
  double* s = n->scratch;
  if (dex3) AF_CUDA(cudaMemcpyAsync(s, dex3, 3 * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  if (x) AF_CUDA(cudaMemcpyAsync(s + 8, x, 18 * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  k_junction_debug<<<1, 1, 0, n->stream>>>(n->nd, n->cur, network, junction, mode, s, s + 8, k_max, tol, s + 64);
  AF_CUDA(cudaGetLastError());
  return d2h(n, host_out, s + 64, n_out);
}

extern "C" int af_junction_cfl_dex(af_net* n, int32_t network, int32_t junction, double* dex) {
  if (!dex) return fail(AF_EINVAL, "null argument");
  return junction_call(n, network, junction, 0, nullptr, nullptr, 0, 0.0, dex, 1);
}

extern "C" int af_junction_cfl_min(af_net* n, int32_t network, int32_t junction, double* M) {
  if (!M) return fail(AF_EINVAL, "null argument");
  double buf[2];
  int rc = junction_call(n, network, junction, 0, nullptr, nullptr, 0, 0.0, buf, 2);
  if (rc) return rc;
  *M = buf[1];
  return AF_OK;
}

extern "C" int af_junction_eval(af_net* n, int32_t network, int32_t junction, const double dex[3], const double* x,
                                double* y, double* J) {
  if (!dex || !x) return fail(AF_EINVAL, "null argument");
  double buf[18 + 324];
  int rc = junction_call(n, network, junction, 1, dex, x, 0, 0.0, buf, 18 + 324);
  if (rc) return rc;
  if (y) memcpy(y, buf, 18 * sizeof(double));
  if (J) memcpy(J, buf + 18, 324 * sizeof(double));
  return AF_OK;
}

extern "C" int af_junction_newton(af_net* n, int32_t network, int32_t junction, const double dex[3], double* x,
                                  int32_t k_max, double tol, int32_t* solves) {
  if (!dex || !x) return fail(AF_EINVAL, "null argument");
  double buf[20];
  int rc = junction_call(n, network, junction, 2, dex, x, k_max, tol, buf, 20);
  if (rc) return rc;
  memcpy(x, buf, 18 * sizeof(double));
  if (solves) *solves = (int32_t)buf[18];
  return AF_OK;
}

extern "C" int af_windkessel(af_net* n, int32_t network, int32_t leaf, int32_t k_max, double tol, double* A_out,
                             double* dex, int32_t* its) {
  int rc = check(n);
  if (rc) return rc;
  if (network < 0 || network >= n->nd.nb || leaf < 0 || leaf >= n->nd.nl)
    return fail(AF_EINVAL, "network/leaf index out of range");
  k_windkessel_debug<<<1, 1, 0, n->stream>>>(n->nd, n->cur, network, leaf, k_max, tol, n->scratch);
  AF_CUDA(cudaGetLastError());
  double buf[3];
  if ((rc = d2h(n, buf, n->scratch, 3))) return rc;
  if (A_out) *A_out = buf[0];
  if (dex) *dex = buf[1];
  if (its) *its = (int32_t)buf[2];
  return AF_OK;
}

// ---- FP64 peak microbenchmark (roofline denominator) --------------------------------
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
  double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int af_measure_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return fail(AF_EINVAL, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(AF_ENODEVICE, "no CUDA device");
  AF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AF_CUDA(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  double* d = nullptr;
  AF_CUDA(cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_peak<<<blocks, threads>>>(d, 1024);        // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma_peak<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(d); return fail(AF_ECUDA, cudaGetErrorString(e)); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return AF_OK;
}
