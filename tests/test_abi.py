"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mp2p_b200.h declares; struct layouts match the reference's 36/72-byte records; without a
GPU the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mp2p_icp_b200 as b200
from mp2p_icp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mp2p_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mp2p_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = b200.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in include/mp2p_b200.h but not exported"
    assert sorted(capi.EXPORTS) == declared


def test_record_layouts():
    assert capi.PAIR_PT2PT.itemsize == 36  # mrpt::tfest::TMatchingPair
    assert capi.PAIR_PT2PL.itemsize == 72  # mp2p_icp::point_plane_pair_t
    assert capi.PAIR_PT2PL.fields["local"][1] == 56
    assert C.sizeof(capi._Pt2PtParams) == 40 and C.sizeof(capi._GNParams) == 56


def test_pack_bits():
    f = np.zeros(70, np.uint8)
    f[[0, 31, 32, 69]] = 1
    w = capi.pack_bits(f)
    assert list(w) == [0x80000001, 0x1, 1 << 5]


def test_host_side_packet_functions_need_no_gpu():
    """gn_step_from_packet / horn_finish are pure host math on the reduced accumulators."""
    from oracle import oracle_py as orc

    rng = np.random.default_rng(0)
    J = rng.normal(size=(40, 6))
    r = rng.normal(size=40) * 0.01
    H, g = J.T @ J, J.T @ r
    pk = np.zeros(32)
    pk[:21] = H[np.triu_indices(6)]
    pk[21:27] = g
    pk[27] = r @ r
    T0 = orc.pose_from_xyzypr(1, 2, 3, 0.1, 0.2, 0.3)
    T1, conv = capi.gn_step_from_packet(pk, capi.GNParams(), T0)
    exp = orc.compose(T0, orc.se3_exp(-np.linalg.solve(H, g)))
    assert np.abs(T1 - exp).max() < 1e-12 and not conv

    # Horn finish from exact sums
    A = rng.uniform(0, 10, (50, 3))
    gt = orc.pose_from_xyzypr(0.3, -0.2, 0.1, 0.05, -0.02, 0.03)
    B = (A - gt[:, 3]) @ gt[:, :3]
    sums = np.zeros(32)
    sums[0:3], sums[3:6], sums[6] = B.sum(0), A.sum(0), 50
    cl, cg = B.mean(0), A.mean(0)
    mom = np.zeros(32)
    mom[:9] = ((B - cl).T @ (A - cg)).reshape(-1)
    mom[9] = 50
    ok, T = capi.horn_finish(sums, mom)
    assert ok and np.abs(T - gt).max() < 1e-9


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.Mp2pError, match="no CPU fallback"):
        b200.Context(0)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mp2p_icp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", txt, flags=re.M) or "liboracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_adaptive_threshold_is_host_math_and_matches_the_oracle_statement():
    """mp2p_b200_adaptive_threshold (the library's restatement of MRPT's CHistogram normalisation and
    confidenceIntervalsFromHistogram, Matcher_Adaptive.cpp:188-214) needs no GPU; it must agree with the
    numpy statement of the same steps used to pin the oracle (tests/test_oracle_golden.py)."""
    import ctypes as C

    from tests.test_oracle_golden import _adaptive_threshold_numpy

    L = capi.load_library()
    rng = np.random.default_rng(5)
    for ci, mind in [(0.8, 0.1), (0.75, 0.01), (0.5, 0.0), (0.95, 0.3)]:
        e = np.abs(rng.normal(0, 0.2, 5000)).astype(np.float32) ** 2
        lo, hi = float(e.min()), float(e.max())
        inv = 49.0 / (hi - lo)
        hist = np.bincount((inv * (e.astype(np.float64) - lo)).astype(np.int64), minlength=50)[:50].astype(np.uint64)
        ci_high, thr = C.c_double(0), C.c_double(0)
        rc = L.mp2p_b200_adaptive_threshold(hist.ctypes.data_as(C.c_void_p), C.c_double(lo), C.c_double(hi), C.c_uint64(len(e)), C.c_double(ci), C.c_double(mind), C.byref(ci_high), C.byref(thr))
        exp_thr, exp_hi = _adaptive_threshold_numpy(e, ci, mind)
        assert rc == 0 and abs(ci_high.value - exp_hi) <= 1e-12 * max(1.0, exp_hi) and abs(thr.value - exp_thr) <= 1e-12 * max(1.0, exp_thr)
    # the reference throws: no sample, or max == min
    z = np.zeros(50, np.uint64)
    assert L.mp2p_b200_adaptive_threshold(z.ctypes.data_as(C.c_void_p), C.c_double(0), C.c_double(0), C.c_uint64(0), C.c_double(0.8), C.c_double(0.1), C.byref(ci_high), C.byref(thr)) != 0
    assert L.mp2p_b200_adaptive_threshold(z.ctypes.data_as(C.c_void_p), C.c_double(1.0), C.c_double(1.0), C.c_uint64(7), C.c_double(0.8), C.c_double(0.1), C.byref(ci_high), C.byref(thr)) != 0
