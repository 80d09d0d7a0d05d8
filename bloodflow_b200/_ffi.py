"""ctypes binding of include/arteryfe_b200.h (the drop-in C ABI).

Loading is strict: if ``csrc/libarteryfe_b200.so`` is missing or does not
export a declared symbol this module raises -- the product has no other path.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# the in-tree build; ARTERYFE_B200_LIB points at a library installed elsewhere
LIB_PATH = os.environ.get('ARTERYFE_B200_LIB') or os.path.join(_HERE, 'csrc', 'libarteryfe_b200.so')

AF_ABI_VERSION = 1
AF_OK, AF_EINVAL, AF_ENODEVICE, AF_ECUDA, AF_ENOMEM, AF_ESTATE = 0, -1, -2, -3, -4, -5
AF_FLAG_ARTERY_NOCONV, AF_FLAG_JUNCTION_KMAX, AF_FLAG_PICARD_KMAX = 1, 2, 4
AF_FLAG_NONFINITE, AF_FLAG_SINGULAR, AF_FLAG_PICARD_CYCLE, AF_FLAG_SYNC_TIMEOUT = 8, 16, 32, 64
AF_STORE_FLOW, AF_STORE_AREA, AF_STORE_PRESSURE, AF_STORE_OUTLET_PRESSURE = 1, 2, 4, 8
AF_SCHEDULE_PER_STEP, AF_SCHEDULE_DATAFLOW, AF_SCHEDULE_PERSISTENT, AF_SCHEDULE_FUSED = 0, 1, 2, 3

c_int, c_i32, c_i64, c_dbl = ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
P_dbl = ctypes.POINTER(ctypes.c_double)
P_i32 = ctypes.POINTER(ctypes.c_int32)
P_i64 = ctypes.POINTER(ctypes.c_int64)
P_u32 = ctypes.POINTER(ctypes.c_uint32)
P_u8 = ctypes.POINTER(ctypes.c_uint8)
Handle = ctypes.c_void_p


class AfDesc(ctypes.Structure):
    """struct af_desc (include/arteryfe_b200.h)."""
    _fields_ = [
        ('abi_version', c_i32), ('n_networks', c_i32), ('n_arteries', c_i32), ('n_junctions', c_i32),
        ('n_leaves', c_i32), ('Nx', c_i32), ('Nt', c_i32), ('q_ins_per_network', c_i32),
        ('theta', c_dbl), ('dt', c_dbl),
        ('junction_parent', P_i32), ('junction_d1', P_i32), ('junction_d2', P_i32), ('leaf_artery', P_i32),
        ('k1', P_dbl), ('k2', P_dbl), ('k3', P_dbl), ('rho', P_dbl), ('Re', P_dbl), ('db', P_dbl), ('p0', P_dbl),
        ('Ru', P_dbl), ('Rd', P_dbl), ('L', P_dbl), ('q0', P_dbl),
        ('R1', P_dbl), ('R2', P_dbl), ('CT', P_dbl),
        ('q_ins', P_dbl),
        ('newton_rtol', c_dbl), ('newton_atol', c_dbl), ('newton_max_it', c_i32), ('junction_k_max', c_i32),
        ('junction_tol', c_dbl), ('picard_k_max', c_i32), ('picard_tol', c_dbl), ('cfl_margin', c_dbl),
        ('device', c_i32), ('root_artery', c_i32), ('stream', ctypes.c_void_p),
    ]


# name -> (restype, argtypes); exactly the declarations of the header
PROTOTYPES = {
    'af_version': (c_int, []),
    'af_last_error': (ctypes.c_char_p, []),
    'af_desc_defaults': (None, [ctypes.POINTER(AfDesc)]),
    'af_net_create': (c_int, [ctypes.POINTER(AfDesc), ctypes.POINTER(Handle)]),
    'af_net_destroy': (None, [Handle]),
    'af_release_cached_memory': (c_int, []),
    'af_net_step': (c_int, [Handle, c_i64, c_i64]),
    'af_net_sync': (c_int, [Handle]),
    'af_net_failed_step': (c_int, [Handle, P_i64]),
    'af_net_schedule': (c_int, [Handle, P_i32]),
    'af_net_clear_flags': (c_int, [Handle]),
    'af_net_outlet_moments_d': (c_int, [Handle, ctypes.POINTER(ctypes.c_void_p), P_i32]),
    'af_moments_combine_d': (c_int, [c_i32, ctypes.c_void_p, ctypes.c_void_p, c_i32, c_i32, ctypes.c_void_p]),
    'af_junction_cfl_min': (c_int, [Handle, c_i32, c_i32, P_dbl]),
    'af_net_set_store': (c_int, [Handle, P_u8, c_i32, c_i32, c_i32]),
    'af_net_set_store_host': (c_int, [Handle, P_dbl, P_dbl, P_dbl, P_dbl]),
    'af_net_set_store_host_pinned': (c_int, [Handle, ctypes.POINTER(P_dbl), ctypes.POINTER(P_dbl),
                                             ctypes.POINTER(P_dbl), ctypes.POINTER(P_dbl)]),
    'af_host_release': (c_int, [ctypes.c_void_p]),
    'af_net_snapshot_count': (c_int, [Handle, P_i32]),
    'af_net_fetch_snapshots': (c_int, [Handle, P_dbl, P_dbl, P_dbl, P_dbl, P_i64]),
    'af_net_outlet_pressure_d': (c_int, [Handle, ctypes.POINTER(ctypes.c_void_p)]),
    'af_net_get_state': (c_int, [Handle, P_dbl, P_dbl]),
    'af_net_set_state': (c_int, [Handle, P_dbl, P_dbl]),
    'af_net_get_pressure': (c_int, [Handle, P_dbl]),
    'af_net_get_bc': (c_int, [Handle, P_dbl]),
    'af_net_set_bc': (c_int, [Handle, P_dbl]),
    'af_net_get_dex': (c_int, [Handle, P_dbl]),
    'af_net_set_dex': (c_int, [Handle, P_dbl]),
    'af_net_get_junction_x': (c_int, [Handle, P_dbl]),
    'af_net_set_junction_x': (c_int, [Handle, P_dbl]),
    'af_net_get_counters': (c_int, [Handle, P_i32, P_i32, P_i32]),
    'af_net_get_totals': (c_int, [Handle, P_i64]),
    'af_net_get_flags': (c_int, [Handle, P_u32]),
    'af_net_set_counter_log': (c_int, [Handle, c_i32]),
    'af_net_fetch_counter_log': (c_int, [Handle, P_i32, P_i32, P_i32, P_i32]),
    'af_phase_boundary': (c_int, [Handle]),
    'af_phase_arteries': (c_int, [Handle, c_i32]),
    'af_artery_solve': (c_int, [Handle, c_i32, c_i32]),
    'af_artery_update': (c_int, [Handle, c_i32, c_i32]),
    'af_eval_state': (c_int, [Handle, c_i32, c_i32, c_i32, c_i32, P_dbl, P_dbl]),
    'af_eval_coefficients': (c_int, [Handle, c_i32, c_i32, c_i32, P_dbl, P_dbl]),
    'af_eval_flux_source': (c_int, [Handle, c_i32, c_i32, c_i32, P_dbl, P_dbl, P_dbl]),
    'af_eval_u_half': (c_int, [Handle, c_i32, c_i32, c_dbl, c_dbl, P_dbl, P_dbl, P_dbl]),
    'af_eval_cfl_term': (c_int, [Handle, c_i32, c_i32, c_dbl, c_dbl, c_dbl, P_dbl]),
    'af_eval_math': (c_int, [c_i32, c_i32, c_i32, P_dbl, P_dbl]),
    'af_junction_cfl_dex': (c_int, [Handle, c_i32, c_i32, P_dbl]),
    'af_junction_eval': (c_int, [Handle, c_i32, c_i32, P_dbl, P_dbl, P_dbl, P_dbl]),
    'af_junction_newton': (c_int, [Handle, c_i32, c_i32, P_dbl, P_dbl, c_i32, c_dbl, P_i32]),
    'af_windkessel': (c_int, [Handle, c_i32, c_i32, c_i32, c_dbl, P_dbl, P_dbl, P_i32]),
    'af_measure_fp64_peak': (c_int, [c_i32, P_dbl]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "bloodflow_b200: %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C bloodflow_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.af_version() != AF_ABI_VERSION:
        raise ImportError("libarteryfe_b200.so ABI version %d != %d" % (lib.af_version(), AF_ABI_VERSION))
    return lib


lib = _load()


class AfError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, "arteryfe_b200 error %d: %s" % (code, text))
        self.code = code


def check(rc):
    if rc != AF_OK:
        raise AfError(rc, lib.af_last_error().decode('utf-8', 'replace'))


def dptr(a):
    """double* of a C-contiguous float64 array."""
    assert a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(P_dbl)


def iptr(a):
    assert a.dtype == np.int32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(P_i32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
