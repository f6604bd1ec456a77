"""mp2p_icp_b200 — B200-native (sm_100a) implementation of the mp2p_icp Matcher+Solver hot path.

The product is the C-ABI shared library ``libmp2p_b200.so`` (sources in ``csrc/``, header
``include/mp2p_b200.h``) plus the C++ host mirror of the reference's Matcher/Solver plugin
interface (``host/``). This Python package is only the ctypes binding used by the tests, the
benchmark and multi-process (torch.distributed) drivers. There is no CPU fallback: importing works
anywhere, every compute call needs the built library and a CUDA device.
"""
from .capi import (  # noqa: F401
    AdaptiveParams,
    PAIR_PT2LN,
    PAIR_PT2PL,
    PAIR_PT2PT,
    Cloud,
    Context,
    GNParams,
    HornParams,
    InlierRatioParams,
    Map,
    Mp2pError,
    Pt2LnParams,
    Pt2PlParams,
    Pt2PtParams,
    library_path,
    load_library,
    pack_bits,
)
