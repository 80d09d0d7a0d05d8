"""UQ ensembles: many parameter-perturbed copies of one network topology,
solved as ONE batched device network per GPU (BASELINE.json config 5).

The reference has no ensemble driver (it is serial); an ensemble
there is a Python loop over ``ArteryNetwork(param).solve()``.  Here the members
are the batch dimension of the C ABI (``af_desc.n_networks``): same topology,
per-member stiffness / radii / Windkessel parameters.  Networks are independent,
so GPUs own disjoint slices of the member index range and nothing is
communicated during the time loop; the only collective is the reduction of the
outlet-pressure moments at the end (NCCL through torch.distributed).

Perturbation model: member ``i`` draws from
``numpy.random.Generator(PCG64(seed + i))``, in this order,
  s  ~ U(0.9, 1.1)      common factor on k1, k3 (stiffness)
  r  ~ U(0.9, 1.1)      factor on the root radius (all radii and lengths of a
                        geometric-series tree scale with it)
  R1 ~ U(0.8, 1.2)^nl,  R2 ~ U(0.8, 1.2)^nl,  CT ~ U(0.8, 1.2)^nl   per leaf.
Member 0 of ``first_index == 0`` is NOT special: use ``perturb=False`` for the
nominal network.
"""
import numpy as np

from . import _ffi
from .engine import DeviceNetwork
from .synthetic import SEED, leaf_radius, member_factors, perturb_param     # noqa: F401  (re-exported)

def stacked_description(param, indices, seed=SEED, wk_scale=None):
    """Batched description of the members ``indices`` without building every
    member through the facade: the nominal network is set up once (host only)
    and the perturbations are applied to its flat description.  Exact for
    geometric-series trees (``alpha`` geometry): radii scale with the root
    radius, lengths are L*Ru (second daughters 0.01 shorter), the initial flow
    split depends on area ratios only."""
    import copy
    from .arteryfe.artery_network import ArteryNetwork
    if 'alpha' not in param.param:
        raise ValueError("ensembles are generated for 'alpha' (geometric-series) trees")
    nominal = ArteryNetwork(copy.deepcopy(param), wk_scale=wk_scale, _host_only=True)
    d = nominal._description
    nb, na, nl = len(indices), len(d['Ru']), len(d['leaves'])
    rc = nominal.nondim['rc']
    out = {k: d[k] for k in ('Nx', 'Nt', 'theta', 'dt', 'junctions', 'leaves', 'k2', 'rho', 'Re', 'db', 'p0', 'q_ins')}
    k1, k3 = np.empty(nb), np.empty(nb)
    Ru, Rd, L = np.empty((nb, na)), np.empty((nb, na)), np.empty((nb, na))
    R1, R2, CT = np.empty((nb, nl)), np.empty((nb, nl)), np.empty((nb, nl))
    # nominal lengths are L*Ru (- 0.01/rc for second daughters): separate the two parts
    short = np.zeros(na)
    for k, i in enumerate(nominal._idx):
        if i > 0 and i % 2 == 0:
            short[k] = 0.01 / rc
    for m, i in enumerate(indices):
        s, r, f1, f2, fc = member_factors(i, nl, seed)
        k1[m], k3[m] = d['k1'] * s, d['k3'] * s
        Ru[m], Rd[m] = d['Ru'] * r, d['Rd'] * r
        L[m] = (d['L'] + short) * r - short
        R1[m], R2[m], CT[m] = d['R1'] * f1, d['R2'] * f2, d['CT'] * fc
    out.update(k1=k1, k3=k3, Ru=Ru, Rd=Rd, L=L, q0=np.broadcast_to(d['q0'], (nb, na)).copy(), R1=R1, R2=R2, CT=CT)
    return out, nominal


class EnsembleNetwork(object):
    """``n_networks`` perturbed members ``first_index ... first_index+n-1`` on
    one device.  ``solve(n_cycles)`` runs them all device-resident and stores
    the leaf-outlet pressures at the dump steps of the reference's rule."""

    def __init__(self, param, n_networks, first_index=0, seed=SEED, device=0, stream=None, wk_scale=None):
        self.indices = np.arange(first_index, first_index + n_networks)
        desc, self.nominal = stacked_description(param, self.indices, seed, wk_scale)
        self.geo, self.sol = self.nominal.geo, self.nominal.sol
        self.dev = DeviceNetwork(
            desc['Nx'], desc['Nt'], desc['theta'], desc['dt'], desc['junctions'], desc['leaves'],
            k1=desc['k1'], k2=desc['k2'], k3=desc['k3'], rho=desc['rho'], Re=desc['Re'], db=desc['db'], p0=desc['p0'],
            Ru=desc['Ru'], Rd=desc['Rd'], L=desc['L'], q0=desc['q0'], R1=desc['R1'], R2=desc['R2'], CT=desc['CT'],
            q_ins=desc['q_ins'], device=device, stream=stream)
        self.description = desc
        self.n_networks = n_networks
        self.vertices = n_networks * self.dev.na * self.dev.np

    def configure_store(self, n_cycles, n_cycles_store=1):
        mask = self.nominal.store_mask()
        self.dev.set_store(mask, n_cycles - n_cycles_store, max(1, int(mask.sum()) * n_cycles_store),
                           _ffi.AF_STORE_OUTLET_PRESSURE)

    def run(self, n_cycles=None, n_cycles_store=1):
        """Enqueue the whole run; the stored outlet pressures stay on the device
        (``device_moments`` / ``dev.fetch_snapshots``)."""
        n_cycles = self.geo['N_cycles'] if n_cycles is None else n_cycles
        self.configure_store(n_cycles, n_cycles_store)
        self.dev.step(0, n_cycles * self.geo['Nt'])

    def solve(self, n_cycles=None, n_cycles_store=1):
        """Returns the stored outlet pressures [snapshot, member, leaf]."""
        self.run(n_cycles, n_cycles_store)
        self.dev.sync()
        return self.dev.fetch_snapshots()['outlet_pressure']

    def flagged_members(self, mask=_ffi.AF_FLAG_ARTERY_NOCONV | _ffi.AF_FLAG_NONFINITE):
        """Number of members whose status word has one of the ``mask`` bits."""
        return int(np.count_nonzero(self.dev.get_flags() & np.uint32(mask)))

    def flags(self):
        return self.dev.get_flags()


# ---- ensemble statistics -------------------------------------------------------------
# Packed moments: a [5, n] tensor, rows = count, sum, sum of squares, min, max over members;
# n = n_snapshots * n_leaves.  On the GPU the reduction over the members and the combination of the
# ranks' buffers are kernels of the library (af_net_outlet_moments_d / af_moments_combine_d) and the
# ranks exchange their packed buffers with ONE collective (all_gather).  The torch expressions below
# do the same on CPU tensors; they exist for the gloo tests of the host logic only.
def device_moments(dev, device):
    """Packed moments [5, n_snapshots, n_leaves] of the members of ``dev`` (a DeviceNetwork), computed
    on the device by the library; a zero-copy torch view of the library's buffer."""
    ptr, n_snap = dev.outlet_moments_device()
    return device_tensor(ptr, (5, n_snap, dev.nl), device)


def host_moments(op):
    """The same packed moments from a CPU tensor op [snapshot, member, leaf] (tests)."""
    import torch
    n = torch.full(op.shape[::2], float(op.shape[1]), dtype=op.dtype, device=op.device)
    return torch.stack([n, op.sum(dim=1), (op * op).sum(dim=1), op.min(dim=1).values, op.max(dim=1).values])


def combine_moments(packed, group=None):
    """ONE collective: all-gather the ranks' packed moments (NCCL for CUDA tensors, gloo for CPU
    tensors), then fold them.  Returns {count, mean, var, min, max}, each shaped like packed[0]."""
    import torch
    import torch.distributed as dist
    packed = packed.contiguous()
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        flat = torch.empty(world * packed.numel(), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(flat, packed.reshape(-1), group=group)
        parts = flat.view((world,) + tuple(packed.shape))
    else:
        parts = packed[None]
    n = packed[0].numel()
    if packed.is_cuda:
        out = torch.empty_like(packed)
        stream = torch.cuda.current_stream(packed.device).cuda_stream
        _ffi.check(_ffi.lib.af_moments_combine_d(packed.device.index, stream, parts.data_ptr(), world, n,
                                                 out.data_ptr()))
    else:
        cnt, s, ss = parts[:, 0].sum(0), parts[:, 1].sum(0), parts[:, 2].sum(0)
        mean = s / cnt
        out = torch.stack([cnt, mean, (ss / cnt - mean * mean).clamp_min(0.0),
                           parts[:, 3].min(0).values, parts[:, 4].max(0).values])
    return {"count": out[0], "mean": out[1], "var": out[2], "min": out[3], "max": out[4]}


def shard(n_total, rank, world):
    """Contiguous slice [lo, hi) of the member range owned by ``rank``."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def device_tensor(ptr, shape, device):
    """Zero-copy torch view of FP64 device memory owned by the C library."""
    import torch

    class _Wrap(object):
        pass

    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(w, device=torch.device('cuda', device))
