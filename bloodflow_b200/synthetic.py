"""Synthetic symmetric trees and UQ ensembles of the shapes BASELINE.json names
(SURVEY.md 8(d) recipe).  Produces objects with the ``.param/.geo/.solution``
dicts of a ParamParser, i.e. exactly what a .cfg file of the reference would
hold, so they go through the same ArteryNetwork constructor.

theta = 0.6 by default: with the shipped configs' 0.55 the fine-grid trees (C3, C4) and the extreme members
of the C5 ensemble amplify round-off into grid-scale noise (a 1e-13 perturbation reaches 1e-2 within 320
steps, 11 of 65 536 C5 members end non-finite); at 0.6 the same shapes are stable over 4 cycles
(scripts/stability_probe.py, profiles/r02_stability_*.jsonl).

Recipe: ``build_geometry`` tree with alpha, L; untapered leaves
(R_term = Ru0*alpha**(order-1)); per leaf R_tot = p0/(mean_flow/n_leaves),
R1_eff = 2*Z0 (Picard gain 0.5), R2_eff = R_tot - R1_eff, CT = tau/R2_eff; the
cfg values are divided by HEAD's factors 0.3 / 0.01 so HEAD lands on them.
"""
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INLET = os.path.join(ROOT, 'data', 'example_inlet.csv')

PHYS = dict(rc=1.0, qc=10.0, k1=2.0e7, k2=-22.53, k3=8.65e5, rho=1.06, nu=0.046, p0=119990.131579)


def leaf_windkessel(order, Ru0=0.37, alpha=0.8, mean_flow=5.5, tau=0.0186, phys=PHYS):
    """(R1, R2, CT) cfg values (dimensional, before HEAD's factors) of one leaf."""
    r = Ru0 * alpha ** (order - 1)
    n_leaves = 2 ** (order - 1)
    f = 4.0 / 3.0 * (phys['k1'] * np.exp(phys['k2'] * r) + phys['k3'])
    Z0 = phys['rho'] * np.sqrt(f / (2 * phys['rho'])) / (np.pi * r * r)
    R_tot = phys['p0'] / (mean_flow / n_leaves)
    R1_eff = 2 * Z0
    R2_eff = R_tot - R1_eff
    return R1_eff / 0.3, R2_eff / 0.01, tau / R2_eff


def leaf_radius(Ru0, alpha, order):
    """Upstream radius of a last-generation vessel exactly as build_geometry
    (an:195-217) produces it: the repeated product alpha*Ru_parent*1.0."""
    r = float(Ru0)
    for _ in range(order - 1):
        r = alpha * r * 1.0
    return r


def tree_param(order, Nx, Nt, N_cycles=1, Nt_store=None, N_cycles_store=1, Ru0=0.37, alpha=0.8, L=20.0,
               theta=0.6, output_location='output/synthetic', inlet=INLET, store_area=1, store_pressure=1):
    """ParamParser-like description of a symmetric tree of ``order`` levels."""
    N = 2 ** order - 1
    n_leaves = 2 ** (order - 1)
    R1, R2, CT = leaf_windkessel(order, Ru0, alpha)
    mult = np.ones(N)
    mult[0] = Ru0
    param = dict(PHYS)
    # "untapered leaves: R_term = leaf Ru"  has to hold bitwise, or the leaves count as
    # (1e-16-)tapered vessels and take the pow/exp + coefficient-table path: generate R_term by the
    # same repeated product build_geometry uses for the leaf radii (an:195-217)
    r_leaf = leaf_radius(Ru0, alpha, order)
    param.update(order=order, Ru=mult.copy(), Rd=mult.copy(), alpha=alpha, L=L,
                 R_term=r_leaf, p_term=6000.0,
                 R1=np.full(n_leaves, R1), R2=np.full(n_leaves, R2), CT=np.full(n_leaves, CT))
    geo = dict(Nx=Nx, Nt=Nt, N_cycles=N_cycles)
    sol = dict(inlet_flow_location=inlet, output_location=output_location, theta=theta,
               Nt_store=Nt_store if Nt_store is not None else min(Nt, 100), N_cycles_store=N_cycles_store,
               store_area=store_area, store_pressure=store_pressure)
    return types.SimpleNamespace(param=param, geo=geo, solution=sol, fparams=None)


# ---- UQ ensemble perturbation model ------------------------------------
# Member ``i`` draws from numpy.random.Generator(PCG64(SEED + i)), in this order:
#   s ~ U(0.9, 1.1) common factor on k1, k3; r ~ U(0.9, 1.1) factor on the root radius;
#   R1, R2, CT ~ U(0.8, 1.2) per leaf.   (No device code here: bench.py's CPU reference arm uses it.)
SEED = 20260000


def member_factors(i, n_leaves, seed=SEED):
    rng = np.random.Generator(np.random.PCG64(seed + int(i)))
    s = rng.uniform(0.9, 1.1)
    r = rng.uniform(0.9, 1.1)
    return s, r, rng.uniform(0.8, 1.2, n_leaves), rng.uniform(0.8, 1.2, n_leaves), rng.uniform(0.8, 1.2, n_leaves)


def perturb_param(param, i, seed=SEED):
    """The member-``i`` version of a ParamParser-like description (a tree built
    from ``alpha``): what a user would feed to ArteryNetwork for that member."""
    import copy
    p = copy.deepcopy(param)
    n_l = len(np.atleast_1d(p.param['R1']))
    s, r, f1, f2, fc = member_factors(i, n_l, seed)
    p.param['k1'] = p.param['k1'] * s
    p.param['k3'] = p.param['k3'] * s
    p.param['Ru'] = np.array(p.param['Ru'], dtype=float)
    p.param['Rd'] = np.array(p.param['Rd'], dtype=float)
    nominal_leaf = leaf_radius(p.param['Ru'][0], p.param['alpha'], p.param['order']) if 'alpha' in p.param else None
    untapered_leaves = nominal_leaf is not None and p.param['R_term'] == nominal_leaf
    p.param['Ru'][0] *= r
    p.param['Rd'][0] *= r
    # leaves that are untapered in the nominal tree (R_term == leaf Ru bitwise) stay so
    p.param['R_term'] = (leaf_radius(p.param['Ru'][0], p.param['alpha'], p.param['order']) if untapered_leaves
                         else p.param['R_term'] * r)
    p.param['R1'] = np.atleast_1d(p.param['R1']) * f1
    p.param['R2'] = np.atleast_1d(p.param['R2']) * f2
    p.param['CT'] = np.atleast_1d(p.param['CT']) * fc
    return p


def write_cfg(p, path):
    """Write a tree_param() description as a reference-format .cfg file."""
    def fmt(v):
        if np.ndim(v):
            return ','.join(repr(float(x)) for x in v)
        return repr(v) if not isinstance(v, str) else v
    with open(path, 'w') as fh:
        for section, d in (('Parameters', p.param), ('Geometry', p.geo), ('Solution', p.solution)):
            fh.write('[%s]\n' % section)
            for k, v in d.items():
                fh.write('%s = %s\n' % (k, fmt(v)))
            fh.write('\n')
