"""The deterministic simplex of oracle/fjsp_lp.c reaches the LP optimum: its max-min
rate t* equals scipy/HiGHS's on fluid models built from random instances."""
import ctypes

import numpy as np
import pytest
from scipy.optimize import linprog
from scipy.sparse import csc_matrix

import oracle_py
from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance


def fluid_lp(inst, n_rj, prec):
    """Canonical LP of DESIGN.md for unprocessed counts n_rj and precedence flags."""
    M = inst.machine_count
    kt = inst.kind_task_tuple
    KT = len(kt)
    prec_row, nprec = {}, 0
    for q, (r, j) in enumerate(kt):
        if j + 1 < inst.ntask[r] and prec[q]:
            prec_row[q] = M + KT + nprec
            nprec += 1
    colptr, rowidx, vals = [0], [], []
    for q, (r, j) in enumerate(kt):
        for m in range(M):
            if m not in inst.time_rjm[(r, j)]:
                continue
            rate = 1.0 / inst.time_rjm[(r, j)][m]
            rowidx += [m, M + q]
            vals += [1.0, -(rate / n_rj[q])]
            if j > 0 and (q - 1) in prec_row:
                rowidx.append(prec_row[q - 1]); vals.append(rate)
            if q in prec_row:
                rowidx.append(prec_row[q]); vals.append(-rate)
            colptr.append(len(rowidx))
    rowidx += [M + q for q in range(KT)]
    vals += [1.0] * KT
    colptr.append(len(rowidx))
    nrow, ncol = M + KT + nprec, len(colptr) - 1
    b = np.zeros(nrow); b[:M] = 1.0
    return (ncol, nrow, np.array(colptr, np.int32), np.array(rowidx, np.int32), np.array(vals), b)


@pytest.mark.parametrize("seed", range(12))
def test_simplex_reaches_highs_optimum(seed):
    rng = np.random.default_rng(seed)
    inst = FJSPInstance.generate(seed, 1.0, int(rng.integers(3, 21)), 1,
                                 "DA3C" if seed % 2 else "HMPSAC", scale=0.3)
    KT = len(inst.kind_task_tuple)
    n_rj = rng.integers(1, 60, KT).astype(float)
    prec = rng.random(KT) < 0.7
    ncol, nrow, colptr, rowidx, vals, b = fluid_lp(inst, n_rj, prec)
    z = np.zeros(ncol)
    iters = ctypes.c_int(0)
    rc = oracle_py.lib().fjsp_lp_solve_sparse(ncol, nrow, colptr.ctypes.data, rowidx.ctypes.data, vals.ctypes.data,
                                              b.ctypes.data, ncol - 1, z.ctypes.data, ctypes.byref(iters))
    assert rc == 0
    A = csc_matrix((vals, rowidx, colptr), shape=(nrow, ncol))
    c = np.zeros(ncol); c[-1] = -1.0
    res = linprog(c, A_ub=A, b_ub=b, bounds=[(0, None)] * ncol, method="highs")
    assert res.status == 0
    assert z[-1] == pytest.approx(res.x[-1], rel=1e-9)
    assert (A @ z <= b + 1e-9).all() and (z >= 0).all()
    assert (z[:-1] > 0).sum() <= nrow          # a vertex: at most nrow basic variables
