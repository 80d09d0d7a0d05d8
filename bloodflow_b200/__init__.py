"""bloodflow_b200 -- B200-native (sm_100a, FP64) implementation of artery.fe's
per-timestep network solve behind the reference's ``arteryfe`` API.

    import bloodflow_b200.arteryfe as af      # or: import arteryfe as af (repo-root shim)
    an = af.ArteryNetwork(af.ParamParser('config/demo_arterybranch.cfg'))
    an.solve()

The compute path is the hand-written CUDA library ``csrc/libarteryfe_b200.so``
(C ABI in ``include/arteryfe_b200.h``), reached through ctypes.  There is no
CPU fallback: importing ``bloodflow_b200._ffi`` without the built library, or
creating a network without a CUDA device, raises.
"""
__version__ = "0.1.0"


def release_cached_memory():
    """Return the device / pinned blocks cached by the library to the driver
    (af_release_cached_memory; the analogue of torch.cuda.empty_cache)."""
    from . import _ffi
    _ffi.check(_ffi.lib.af_release_cached_memory())
