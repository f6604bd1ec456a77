"""Restatement of the caller of the hot path — the ICP::align loop — for tests and the bench.

Follows mp2p_icp/src/ICP.cpp:108-308 (reference): iterate run_matchers -> run_solvers, terminate on
NoPairings / SolverError / Stalled (min of the 1-step and 2-step SE(3) log increments below
minAbsStep_trans / minAbsStep_rot, Parameters.h:42-52) or maxIterations. Matcher and solver are
callables so the same loop drives the CPU oracle and the CUDA product.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from oracle import oracle_py as orc


@dataclass
class IcpParams:  # mp2p_icp/include/mp2p_icp/Parameters.h:42-52
    maxIterations: int = 40
    minAbsStep_trans: float = 5e-4
    minAbsStep_rot: float = 1e-4


@dataclass
class IcpResult:
    pose: np.ndarray
    nIterations: int
    terminationReason: str
    finalPairings: object


def align(match_fn, solve_fn, init_guess, p: IcpParams = None) -> IcpResult:
    """match_fn(pose, iteration) -> pairings or None;  solve_fn(pairings, guess_pose, iteration) -> (ok, pose)."""
    p = p or IcpParams()
    cur = np.array(init_guess, dtype=np.float64).reshape(3, 4)
    prev, prev2 = cur.copy(), None
    reason, pairings, it = "MaxIterations", None, 0
    for it in range(p.maxIterations):
        pairings = match_fn(cur, it)  # ICP.cpp:143
        if pairings is None or _n(pairings) == 0:
            reason = "NoPairings"  # ICP.cpp:148
            break
        ok, new = solve_fn(pairings, cur, it)  # ICP.cpp:170
        if not ok:
            reason = "SolverError"
            break
        cur = np.array(new, dtype=np.float64).reshape(3, 4)
        d = orc.se3_log(orc.inverse_compose(cur, prev))  # ICP.cpp:203-206
        dxyz, drot = np.linalg.norm(d[:3]), np.linalg.norm(d[3:])
        if prev2 is not None:  # ICP.cpp:208-215
            d2 = orc.se3_log(orc.inverse_compose(cur, prev2))
            dxyz, drot = min(dxyz, np.linalg.norm(d2[:3])), min(drot, np.linalg.norm(d2[3:]))
        if dxyz < p.minAbsStep_trans and drot < p.minAbsStep_rot:  # ICP.cpp:228-229
            reason = "Stalled"
            it += 1  # the reference leaves nIterations un-incremented on break; we report passes run
            break
        prev2, prev = prev, cur.copy()
    else:
        it = p.maxIterations
    return IcpResult(cur, it, reason, pairings)


def _n(pairings):
    if isinstance(pairings, tuple):
        return sum(len(x) for x in pairings if x is not None)
    return len(pairings)


def load_xyz_gz(path):
    import gzip

    with gzip.open(path, "rt") as f:
        a = np.loadtxt(f, dtype=np.float32)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2])
