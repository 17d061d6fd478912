"""wmf_heavy_b200 -- B200-native (sm_100a) drop-in for the raster-drainage hot path of WMF
(uranzea/WMF_heavy): D8 flow accumulation + basin delineation, and the SHIA cell update behind
``SimuBasin.run_shia``.

    cu_py.basics / models_py.core   wmf_py-signature operators  (numpy in, numpy out)
    cu / models                     f2py-signature namespaces with module-level state
    wmfv2                           Basin / SimuBasin / run_shia with the reference signatures
    ops                             device-tensor operators over the C ABI (include/wmf_b200.h)

Everything computes in ``libwmfb200.so`` (hand-written CUDA); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["ops", "cu", "models", "cu_py", "models_py", "wmfv2"]
__version__ = "0.1.0"
