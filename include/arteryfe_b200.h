/* arteryfe_b200.h -- C ABI of the B200-native artery.fe network solve.
 *
 * The reference (KVSlab/bloodflow, package `arteryfe`) has no FFI of its own:
 * its hot path is Python that drives FEniCS through SWIG.  This header is the
 * drop-in boundary a maintainer binds instead (ctypes stub in INTEGRATION.md).
 * Every entry point names the reference code it replaces (paths relative to
 * the reference tree, `an` = arteryfe/artery_network.py, `ar` = arteryfe/artery.py).
 *
 * Conventions
 *   - plain C, no CUDA/torch types; all floating point is IEEE double (FP64)
 *   - all pointers are HOST pointers unless the name ends in _d
 *   - arrays are C-contiguous; "artery" indices are COMPACT (the existing
 *     arteries of the reference's heap-indexed tree, ascending)
 *   - functions return 0 on success, a negative AF_E* code on a usage/runtime
 *     error (text via af_last_error()); they never throw and never exit
 *   - one handle <-> one CUDA device + one stream; calls on one handle are
 *     not thread-safe, different handles are independent
 *   - there is no CPU implementation behind this ABI: without a CUDA device
 *     af_net_create fails with AF_ENODEVICE
 */
#ifndef ARTERYFE_B200_H
#define ARTERYFE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AF_ABI_VERSION 1

#define AF_OK 0
#define AF_EINVAL (-1)     /* bad argument / descriptor */
#define AF_ENODEVICE (-2)  /* no usable CUDA device */
#define AF_ECUDA (-3)      /* CUDA runtime error, see af_last_error() */
#define AF_ENOMEM (-4)
#define AF_ESTATE (-5)     /* call not valid in the handle's current state */

/* numerical status bits, per network (af_net_get_flags) */
#define AF_FLAG_ARTERY_NOCONV 1u  /* artery Newton hit max_it: DOLFIN raises RuntimeError (ar:244) */
#define AF_FLAG_JUNCTION_KMAX 2u  /* junction Newton ran all k_max iterations (an:693, silent in the reference) */
#define AF_FLAG_PICARD_KMAX 4u    /* Windkessel Picard ran all k_max iterations (an:411, silent) */
#define AF_FLAG_NONFINITE 8u      /* NaN/Inf in a residual norm */
#define AF_FLAG_SINGULAR 16u      /* singular 18x18: the perturbation path an:703-708 was taken */
#define AF_FLAG_SYNC_TIMEOUT 64u  /* a dataflow wait between kernels of the time loop gave up (never expected) */
#define AF_FLAG_PICARD_CYCLE 32u  /* Windkessel Picard ended in a round-off limit cycle: the reference runs all k_max
                                     iterations there (an:411); the kernel returns the same value without them */

/* what af_net_step stores at dump steps */
#define AF_STORE_FLOW 1
#define AF_STORE_AREA 2
#define AF_STORE_PRESSURE 4
#define AF_STORE_OUTLET_PRESSURE 8 /* leaf-outlet pressure only (ensemble statistics) */

typedef struct af_net af_net;

/* Description of a batch of `n_networks` networks that share one topology
 * (an ensemble; 1 for a single network).  Everything is nondimensional, i.e.
 * what ArteryNetwork.__init__ holds after nondimensionalise_parameters()
 * (an:56-130, utils.py:47-104). */
typedef struct af_desc {
  int32_t abi_version;  /* AF_ABI_VERSION */
  int32_t n_networks;
  int32_t n_arteries;   /* existing arteries per network */
  int32_t n_junctions;  /* == len(range_parent_arteries) */
  int32_t n_leaves;     /* == len(range_leaf_arteries) */
  int32_t Nx;           /* cells per artery (geo['Nx']); Nx+1 vertices */
  int32_t Nt;           /* steps per cardiac cycle (geo['Nt']) */
  int32_t q_ins_per_network; /* 0: q_ins is [Nt], shared; 1: [n_networks][Nt] */
  double theta;         /* sol['theta'] */
  double dt;            /* T/Nt (an:83) */

  /* topology, shared by the batch; artery `root_artery` (0) is the root */
  const int32_t* junction_parent; /* [n_junctions] compact artery index */
  const int32_t* junction_d1;
  const int32_t* junction_d2;
  const int32_t* leaf_artery;     /* [n_leaves], order of range_leaf_arteries */

  /* per network: [n_networks] */
  const double* k1;
  const double* k2;
  const double* k3;
  const double* rho; /* dimensional rho, used by CFL_term (ar:322) */
  const double* Re;
  const double* db;  /* sqrt(nu*T/2/pi) (ar:100) */
  const double* p0;

  /* per network and artery: [n_networks][n_arteries] */
  const double* Ru;
  const double* Rd; /* after the leaf rule Rd==1 -> R_term (ar:108-109) */
  const double* L;
  const double* q0; /* initial flow (an:268-277) */

  /* per network and leaf: [n_networks][n_leaves]; already multiplied by the
   * Windkessel factors of an:408-409 (0.3, 0.01 at HEAD) */
  const double* R1;
  const double* R2;
  const double* CT;

  const double* q_ins; /* inlet table (an:80-82), see q_ins_per_network */

  /* solver constants; af_desc_defaults() fills the reference's values */
  double newton_rtol;      /* 1e-9  [3P] NewtonSolver relative_tolerance */
  double newton_atol;      /* 1e-10 [3P] absolute_tolerance */
  int32_t newton_max_it;   /* 1000  (ar:243) */
  int32_t junction_k_max;  /* 30    (an:668) */
  double junction_tol;     /* 1e-10 (an:668) */
  int32_t picard_k_max;    /* 100   (an:361) */
  double picard_tol;       /* 1e-12 (an:361) */
  double cfl_margin;       /* 0.05  (an:713, ar:349) */

  int32_t device;          /* CUDA device ordinal */
  int32_t root_artery;     /* compact index of the root (inlet: q only, ar:172-174); 0 by default,
                              -1: no root, every inlet takes (A, q) (a stand-alone non-root Artery) */
  void* stream;            /* cudaStream_t to enqueue on, or NULL for an own stream */
} af_desc;

int af_version(void);
const char* af_last_error(void);

/* Fill the solver constants (and zero everything else). */
void af_desc_defaults(af_desc* d);

/* Replaces ArteryNetwork.__init__/define_geometry/define_solution/define_x
 * (an:56-130, 244-279, 434-473; ar:67-189): builds the device-resident state
 * Un = (A0(x_j), q0), BC holders, junction unknowns x and coefficient tables. */
int af_net_create(const af_desc* d, af_net** out);
void af_net_destroy(af_net* net);

/* Device and pinned-host blocks of destroyed / reconfigured handles are cached
 * (exact-size reuse, AF_CACHE_MB caps the device bytes, default 1024, 0 disables; recycled blocks are cleared)
 * so that loops over ArteryNetwork(param).solve() (the reference's way to run a
 * parameter study) do not pay cudaMalloc/cudaHostAlloc every time.  This call
 * returns everything cached to the driver. */
int af_release_cached_memory(void);

/* Replaces the time loop body (an:908-946) for global steps
 * [n_first, n_first + n_steps): step g has n_cycle = g / Nt, n = g % Nt; it
 * applies set_bcs(q_ins[(n+1) % Nt]) (an:916), solves and updates every
 * artery (an:941-944).  Snapshots configured with af_net_set_store are taken
 * at the start of a step (state at t_n, an:924-938).  Enqueues only; no host
 * synchronisation. */
int af_net_step(af_net* net, int64_t n_first, int64_t n_steps);

/* Failure at the failing step (DOLFIN raises RuntimeError inside Artery.solve, ar:244): when the Newton
 * solve of an artery of a SINGLE network does not converge (or goes non-finite) at global step g, every
 * later launch of the loop returns at once and af_net_step stops enqueuing as soon as its polling (every
 * 64 steps, no synchronisation) sees it.  af_net_failed_step waits for the stream and returns g, or -1.
 * Ensemble members are independent: a failed member is flagged (af_net_get_flags), the others go on. */
int af_net_failed_step(af_net* net, int64_t* step);
/* Reset the status bits and the failed step (they are sticky otherwise). */
int af_net_clear_flags(af_net* net);
/* How af_net_step schedules the loop body of artery_network.py:908-946 on this handle (chosen at creation from
 * the network's shape, the device and the AF_* environment knobs; results are bitwise the same for all):
 *   AF_SCHEDULE_PER_STEP      two launches per time step, ordered by grid-wide waits (ensembles; AF_NO_DATAFLOW=1)
 *   AF_SCHEDULE_DATAFLOW      two launches per time step, ordered by per-artery flags and the Dirichlet mailbox
 *   AF_SCHEDULE_PERSISTENT    one launch pair per chunk of up to 64 time steps (single networks whose kernels fit
 *                             the device together)
 *   AF_SCHEDULE_FUSED         one launch per time step, boundary update inside the artery CTAs (AF_FUSE=1) */
#define AF_SCHEDULE_PER_STEP 0
#define AF_SCHEDULE_DATAFLOW 1
#define AF_SCHEDULE_PERSISTENT 2
#define AF_SCHEDULE_FUSED 3
int af_net_schedule(af_net* net, int32_t* kind);

/* Wait for everything enqueued on the handle's stream. */
int af_net_sync(af_net* net);

/* Dump rule (an:924-926): a snapshot is taken at step g when
 * g / Nt >= first_cycle and step_mask[g % Nt] != 0.  `step_mask` is [Nt]
 * (the caller evaluates the reference's float modulo).  `fields` is a mask of
 * AF_STORE_*.  Capacity `max_snapshots`; further dump steps are dropped. */
int af_net_set_store(af_net* net, const uint8_t* step_mask, int32_t first_cycle, int32_t max_snapshots,
                     int32_t fields);
/* Optional, after af_net_set_store: register HOST destinations (layouts as in
 * af_net_fetch_snapshots; NULL to skip a field).  af_net_step then streams every
 * snapshot out while the loop runs (device -> pinned ring on a second stream ->
 * these buffers); af_net_fetch_snapshots with the same pointers only waits for the
 * last ones.  The buffers must stay valid until then. */
int af_net_set_store_host(af_net* net, double* flow, double* area, double* pressure, double* outlet_pressure);
/* The same with result arrays OWNED BY THE LIBRARY: pinned host buffers sized for the whole store
 * ([max_snapshots][...] per configured field, NULL for the others) are returned, and every snapshot is
 * copied straight into them (no staging ring, no host-side copy).  Meant for small stores; the buffers
 * outlive the handle and must be returned with af_host_release once the caller is done with them. */
int af_net_set_store_host_pinned(af_net* net, double** flow, double** area, double** pressure,
                                 double** outlet_pressure);
int af_host_release(void* buffer);
int af_net_snapshot_count(af_net* net, int32_t* count);
/* Copy stored fields out; each of flow/area/pressure is
 * [count][n_networks][n_arteries][Nx+1] (NULL to skip), outlet_pressure is
 * [count][n_networks][n_leaves], step is [count] (global step index).
 * Returns when the snapshots taken so far have arrived; each download is gated
 * on its own snapshot only, so calling this right after af_net_step overlaps
 * the copies with the rest of the enqueued time loop. */
int af_net_fetch_snapshots(af_net* net, double* flow, double* area, double* pressure, double* outlet_pressure,
                           int64_t* step);
/* Device pointer of the outlet-pressure store [max_snapshots][n_networks][n_leaves]
 * (for on-device ensemble statistics / NCCL reductions without a host copy). */
int af_net_outlet_pressure_d(af_net* net, double** ptr_d);

/* Ensemble statistics on the device: moments of the stored outlet pressures over the members
 * of THIS handle, [5][n_snapshots*n_leaves] = count, sum, sum of squares, min, max (enqueued on the handle's
 * stream; the buffer belongs to the handle).  Ranks exchange these packed buffers with ONE collective. */
int af_net_outlet_moments_d(af_net* net, double** moments_d, int32_t* n_snapshots);
/* Combine n_parts packed moment buffers ([n_parts][5][n], e.g. the all-gathered buffers of all ranks) into
 * [5][n] = count, mean, variance, min, max on `device`, enqueued on `stream` (a cudaStream_t or NULL). */
int af_moments_combine_d(int32_t device, void* stream, const double* parts_d, int32_t n_parts, int32_t n,
                         double* out_d);

/* Un of every artery (ar:154-158): A, q are [n_networks][n_arteries][Nx+1]. */
int af_net_get_state(af_net* net, double* A, double* q);
int af_net_set_state(af_net* net, const double* A, const double* q);
/* update_pressure (ar:254-261): p0 + f(x_j)(1 - sqrt(A0(x_j)/A_j)) of the current Un */
int af_net_get_pressure(af_net* net, double* p);
/* BC holders (ar:368-415): [n_networks][n_arteries][4] = A_in, q_in, A_out, q_out */
int af_net_get_bc(af_net* net, double* bc);
int af_net_set_bc(af_net* net, const double* bc);
/* dex of every artery after the last boundary update: [n_networks][n_arteries] */
int af_net_get_dex(af_net* net, double* dex);
int af_net_set_dex(af_net* net, const double* dex);
/* junction unknowns ArteryNetwork.x (an:279): [n_networks][n_junctions][18] */
int af_net_get_junction_x(af_net* net, double* x);
int af_net_set_junction_x(af_net* net, const double* x);
/* iteration counts of the LAST step: artery Newton linear solves
 * [n_networks][n_arteries], junction Newton solves [n_networks][n_junctions],
 * Picard iterations [n_networks][n_leaves] (any may be NULL) */
int af_net_get_counters(af_net* net, int32_t* artery_its, int32_t* junction_solves, int32_t* picard_its);
/* running totals over all steps so far: {artery solves, junction solves, picard its} */
int af_net_get_totals(af_net* net, int64_t totals[3]);
/* status bits, [n_networks] */
int af_net_get_flags(af_net* net, uint32_t* flags);

/* Optional per-step log of the three counters for parity tests: records steps
 * while fewer than `capacity` have been logged. */
int af_net_set_counter_log(af_net* net, int32_t capacity);
int af_net_fetch_counter_log(af_net* net, int32_t* n_logged, int32_t* artery_its, int32_t* junction_solves,
                             int32_t* picard_its);

/* Single-phase entry points (parity tests; also what ArteryNetwork.set_bcs and
 * Artery.solve/update_solution map to).  `inlet_index` selects q_ins[inlet_index]. */
int af_phase_boundary(af_net* net);                      /* set_bcs (junctions + leaves)  an:779-786 */
/* Artery.solve + update_solution for every artery (ar:233-251).  inlet_index
 * >= 0: the root's q_in is q_ins[inlet_index]; -1: the root's BC holder. */
int af_phase_arteries(af_net* net, int32_t inlet_index);
/* Artery.solve() of ONE artery (ar:233-244): U is computed from Un and the BC
 * holders and kept pending; af_artery_update is update_solution() (ar:247-251). */
int af_artery_solve(af_net* net, int32_t network, int32_t artery);
int af_artery_update(af_net* net, int32_t network, int32_t artery);

/* Fine-grained device evaluations of the boundary physics on the CURRENT
 * state (each launches a one-thread kernel; for tests and the Python facade).
 * `network` selects the ensemble member. */
/* Un(x) (which = 0) or the pending U(x) (which = 1) of one artery at n points
 * (P1 interpolation, an:380,385,511-513): out [n][2] = A, q */
int af_eval_state(af_net* net, int32_t network, int32_t artery, int32_t which, int32_t n, const double* x,
                  double* out);
/* Expressions f, A0, dfdr, drdx of one artery at n points (ar:110-118): out [n][4] */
int af_eval_coefficients(af_net* net, int32_t network, int32_t artery, int32_t n, const double* x, double* out);
/* flux / source (an:282-326) at n points x with states U [n][2]: out [n][3] = F1, F2, S2 */
int af_eval_flux_source(af_net* net, int32_t network, int32_t artery, int32_t n, const double* x, const double* U,
                        double* out);
/* compute_U_half (an:329-358): out[2] */
int af_eval_u_half(af_net* net, int32_t network, int32_t artery, double x0, double x1, const double U0[2],
                   const double U1[2], double out[2]);
/* CFL_term (ar:304-325) at point x with state (A, q) */
int af_eval_cfl_term(af_net* net, int32_t network, int32_t artery, double x, double A, double q, double* M);
/* The kernels' branch-free reciprocal (which = 0) and reciprocal square root
 * (which = 1) at n points: validation of their accuracy against 1/x, 1/sqrt(x)
 * (the reference uses the IEEE operations of numpy / the form compiler). */
int af_eval_math(int32_t device, int32_t which, int32_t n, const double* x, double* out);
/* adjust_bifurcation_step (an:713-737): the common dex of one junction */
int af_junction_cfl_dex(af_net* net, int32_t network, int32_t junction, double* dex);
/* the same junction's min over the three CFL terms M (an:731-736); dex = (1 + margin) * dt / M */
int af_junction_cfl_min(af_net* net, int32_t network, int32_t junction, double* M);
/* problem_function / jacobian (an:476-665) at x[18] with the given dex of
 * (parent, d1, d2): y[18], J[18*18] row-major (either may be NULL) */
int af_junction_eval(af_net* net, int32_t network, int32_t junction, const double dex[3], const double* x,
                     double* y, double* J);
/* newton (an:668-710): x[18] in/out, returns the number of linear solves */
int af_junction_newton(af_net* net, int32_t network, int32_t junction, const double dex[3], double* x,
                       int32_t k_max, double tol, int32_t* solves);
/* windkessel / compute_A_out (an:361-422) of one leaf: A_out, the dex it chose, Picard iterations */
int af_windkessel(af_net* net, int32_t network, int32_t leaf, int32_t k_max, double tol, double* A_out,
                  double* dex, int32_t* its);

/* Measurement aid (no reference counterpart): sustained FP64 FMA throughput of
 * the device in TFLOP/s (2 flop per DFMA), the denominator of the FP64 roofline. */
int af_measure_fp64_peak(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ARTERYFE_B200_H */
