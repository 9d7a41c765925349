"""TEST INFRASTRUCTURE ONLY -- a stand-in for `docplex.mp.model.Model`.

The reference's fluid model (environments/class_FJSP.py:256-290,
class_MODFJSP.py:251-285) is solved by IBM CPLEX through `docplex`, a closed
third-party dependency that is absent from /root/reference and from this image.
This shim implements exactly the slice of the docplex modelling API that
`fluid_model()` touches, so that the UNMODIFIED reference environments can be
imported and stepped here to generate golden vectors (tests/golden/).

The LP is handed to one of two backends:

  * "oracle" (default): the deterministic simplex of oracle/fjsp_lp.c (the same
    pivoting specification the CUDA kernel implements), through ctypes;
  * "highs": scipy.optimize.linprog(method="highs"), used only to check that the
    oracle simplex reaches the LP optimum (the max-min objective value).

Canonical LP handed to the backends (see DESIGN.md "fluid LP specification"):
  columns : X variables sorted by (key[1], key[0]) = ((r, j), m), then t
  rows    : machine capacity rows in the order they were added (m ascending),
            then one "rate/n >= t" row per argument of model.min() in argument
            order, then the precedence rows in the order they were added
  all rows are written  a.z <= b  with b >= 0, all variables >= 0.
The returned value dict is keyed in the iteration order of the dict passed to
get_value_dict() (what real docplex does), which for the reference is the CPython
set-iteration order of `var_list`.
"""
import ctypes
import os

import numpy as np

BACKEND = os.environ.get("FJSP_SHIM_LP_BACKEND", "oracle")
LAST_LP = {}  # the last canonical LP and its solution, for tests
STATS = {"solves": 0, "iterations": 0, "max_iterations": 0}


class LinExpr:
    __slots__ = ("terms", "const")

    def __init__(self, terms=None, const=0.0):
        self.terms = terms if terms is not None else {}
        self.const = const

    @staticmethod
    def lift(v):
        if isinstance(v, LinExpr):
            return v
        if isinstance(v, Var):
            return LinExpr({v: 1.0})
        return LinExpr({}, float(v))

    def __add__(self, other):
        o = LinExpr.lift(other)
        terms = dict(self.terms)
        for v, c in o.terms.items():
            terms[v] = terms.get(v, 0.0) + c
        return LinExpr(terms, self.const + o.const)

    __radd__ = __add__

    def __sub__(self, other):
        o = LinExpr.lift(other)
        terms = dict(self.terms)
        for v, c in o.terms.items():
            terms[v] = terms.get(v, 0.0) - c
        return LinExpr(terms, self.const - o.const)

    def __mul__(self, k):
        k = float(k)
        return LinExpr({v: c * k for v, c in self.terms.items()}, self.const * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        k = float(k)
        return LinExpr({v: c / k for v, c in self.terms.items()}, self.const / k)

    def __ge__(self, other):
        return Constraint(LinExpr.lift(other) - self)  # other - self <= 0

    def __le__(self, other):
        return Constraint(self - LinExpr.lift(other))  # self - other <= 0


class Var:
    __slots__ = ("key", "lb", "ub")

    def __init__(self, key, lb, ub):
        self.key, self.lb, self.ub = key, lb, ub

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __mul__(self, k):
        return LinExpr({self: float(k)})

    __rmul__ = __mul__

    def __add__(self, other):
        return LinExpr.lift(self) + other

    __radd__ = __add__

    def __sub__(self, other):
        return LinExpr.lift(self) - other

    def __le__(self, other):
        return LinExpr.lift(self) <= other

    def __ge__(self, other):
        return LinExpr.lift(self) >= other


class Constraint:
    """expr <= 0 with expr = sum(terms) + const, i.e. sum(terms) <= -const."""
    __slots__ = ("expr",)

    def __init__(self, expr):
        self.expr = expr


class MinExpr:
    def __init__(self, args):
        self.args = [LinExpr.lift(a) for a in args]


class Solution:
    def __init__(self, values, objective):
        self._values = values
        self.objective_value = objective

    def get_value_dict(self, var_dict):
        return {k: self._values[v] for k, v in var_dict.items()}

    def get_value(self, v):
        return self._values[v]


_LIB = None


def _oracle_lib():
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.normpath(os.path.join(here, "..", "..", "..", "_build", "libfjsp_oracle.so"))
        _LIB = ctypes.CDLL(path)
        _LIB.fjsp_lp_solve_sparse.restype = ctypes.c_int
        _LIB.fjsp_lp_solve_sparse.argtypes = [
            ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


class Model:
    def __init__(self, name=None, **kw):
        self.name = name
        self.vars = []
        self.constraints = []
        self.objective = None

    def continuous_var_dict(self, keys, lb=0, ub=None, name=None):
        out = {}
        for k in keys:  # iteration order of `keys` is preserved, as in docplex
            v = Var(k, lb, ub)
            self.vars.append(v)
            out[k] = v
        return out

    def sum(self, args):
        acc = LinExpr()
        for a in args:
            acc = acc + a
        return acc

    def min(self, *args):
        if len(args) == 1 and not isinstance(args[0], (LinExpr, Var)):
            args = list(args[0])
        return MinExpr(args)

    def maximize(self, expr):
        self.objective = expr

    def add_constraints(self, cts):
        for c in cts:
            self.constraints.append(c)

    def add_constraint(self, c):
        self.constraints.append(c)

    # ------------------------------------------------------------------ solve
    def canonical(self):
        assert isinstance(self.objective, MinExpr), "shim only supports maximize(min(...))"
        cols = sorted(self.vars, key=lambda v: (v.key[1], v.key[0]))
        col_of = {v: i for i, v in enumerate(cols)}
        nx = len(cols)
        t_col = nx
        rows = []  # list of (dict col->coef, rhs)
        cap = [c for c in self.constraints if c.expr.const != 0.0]
        prec = [c for c in self.constraints if c.expr.const == 0.0]
        for c in cap:
            rows.append(({col_of[v]: a for v, a in c.expr.terms.items()}, -c.expr.const))
        for e in self.objective.args:  # t - e <= 0
            r = {col_of[v]: -a for v, a in e.terms.items()}
            r[t_col] = 1.0
            assert e.const == 0.0
            rows.append((r, 0.0))
        for c in prec:
            rows.append(({col_of[v]: a for v, a in c.expr.terms.items()}, 0.0))
        ncol, nrow = nx + 1, len(rows)
        # CSC, row indices ascending inside a column
        percol = [[] for _ in range(ncol)]
        for i, (r, _) in enumerate(rows):
            for j, a in r.items():
                percol[j].append((i, a))
        colptr = np.zeros(ncol + 1, dtype=np.int32)
        rowidx, vals = [], []
        for j in range(ncol):
            percol[j].sort()
            colptr[j + 1] = colptr[j] + len(percol[j])
            rowidx += [i for i, _ in percol[j]]
            vals += [a for _, a in percol[j]]
        b = np.array([rhs for _, rhs in rows], dtype=np.float64)
        return dict(cols=cols, ncol=ncol, nrow=nrow, colptr=colptr,
                    rowidx=np.array(rowidx, dtype=np.int32),
                    vals=np.array(vals, dtype=np.float64), b=b, t_col=t_col)

    def solve(self, **kw):
        lp = self.canonical()
        ncol, nrow = lp["ncol"], lp["nrow"]
        if BACKEND == "highs":
            from scipy.optimize import linprog
            from scipy.sparse import csc_matrix
            A = csc_matrix((lp["vals"], lp["rowidx"], lp["colptr"]), shape=(nrow, ncol))
            c = np.zeros(ncol)
            c[lp["t_col"]] = -1.0
            res = linprog(c, A_ub=A, b_ub=lp["b"], bounds=[(0, None)] * ncol, method="highs")
            assert res.status == 0, res.message
            z = res.x
            iters = int(res.nit)
        else:
            lib = _oracle_lib()
            z = np.zeros(ncol, dtype=np.float64)
            iters_c = ctypes.c_int(0)
            rc = lib.fjsp_lp_solve_sparse(
                ncol, nrow, lp["colptr"].ctypes.data, lp["rowidx"].ctypes.data,
                lp["vals"].ctypes.data, lp["b"].ctypes.data, lp["t_col"],
                z.ctypes.data, ctypes.byref(iters_c))
            assert rc == 0, "oracle simplex failed rc=%d" % rc
            iters = iters_c.value
        STATS["solves"] += 1
        STATS["iterations"] += iters
        STATS["max_iterations"] = max(STATS["max_iterations"], iters)
        LAST_LP.clear()
        LAST_LP.update(lp)
        LAST_LP["z"] = z.copy()
        values = {v: float(z[i]) for i, v in enumerate(lp["cols"])}
        return Solution(values, float(z[lp["t_col"]]))
