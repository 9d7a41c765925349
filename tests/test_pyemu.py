"""oracle/pyemu.h against the live interpreter: CPython's float sum() and the slot
order of small-int sets (what breaks ties in the reference's machine_select)."""
import ctypes
import sys

import numpy as np
import pytest

import oracle_py


@pytest.fixture(scope="module")
def L():
    lib = oracle_py.lib()
    lib.fjsp_pysum.restype = ctypes.c_double
    return lib


@pytest.mark.skipif(sys.version_info < (3, 12), reason="compensated sum() is CPython >= 3.12")
def test_pysum_matches_builtin_sum(L):
    rng = np.random.default_rng(0)
    for trial in range(4000):
        n = int(rng.integers(1, 200))
        x = [rng.normal(size=n), rng.uniform(0, 1, n) ** 2,
             rng.normal(size=n) * 10.0 ** rng.integers(-8, 8, n),
             np.round(rng.normal(size=n) * 5)][trial % 4]
        x = np.ascontiguousarray(x, np.float64)
        assert L.fjsp_pysum(x.ctypes.data, n, 1) == sum(float(v) for v in x)
        acc = 0.0
        for v in x:
            acc += float(v)
        assert L.fjsp_pysum(x.ctypes.data, n, 0) == acc


def test_set_order_matches_cpython(L):
    rng = np.random.default_rng(1)
    out = np.zeros(32, np.int32)
    for trial in range(20000):
        M = int(rng.integers(1, 33))
        n = int(rng.integers(1, M + 1))
        seq = np.ascontiguousarray(rng.permutation(M)[:n], np.int32)
        k = L.fjsp_pyset_order(seq.ctypes.data, n, out.ctypes.data)
        assert list(out[:k]) == list(set(int(v) for v in seq))


def test_intersection_list_matches_cpython(L):
    rng = np.random.default_rng(2)
    out = np.zeros(32, np.int32)
    for trial in range(30000):
        M = int(rng.integers(1, 33))
        idle = [m for m in range(M) if rng.random() < rng.choice([0.2, 0.5, 0.9])]
        n = int(rng.integers(1, M + 1))
        tup = [int(v) for v in rng.permutation(M)[:n]]
        a = np.ascontiguousarray(idle, np.int32)
        b = np.ascontiguousarray(tup, np.int32)
        k = L.fjsp_pyset_intersection_list(a.ctypes.data, len(idle), b.ctypes.data, n, out.ctypes.data)
        assert list(out[:k]) == list(set(idle) & set(tup))
