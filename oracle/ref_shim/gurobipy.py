"""Stand-in for the ``gurobipy`` module.  TEST INFRASTRUCTURE ONLY (oracle).

Gurobi is a closed-source, licensed third-party dependency of the reference
(``/root/reference/src/MaCroDNA/macrodna.py:2-3``; no version pin in the repo,
logs show 9.5.1 and 10.0.2) and is not installed in this image.  This module
provides exactly the API surface the reference's ``ilp`` touches
(``macrodna.py:33-82``): ``Model`` with ``addVar / addConstr / update /
setObjective / optimize / objVal / IsMIP / status / Params.MIPFocus``, ``Var.x``,
``quicksum``, ``LinExpr`` and ``GRB.{BINARY, MAXIMIZE, Status.INFEASIBLE}``.

``Model.optimize()`` rebuilds the |R| x N cost matrix from the objective's
``x[i,j]`` coefficients and solves the rectangular assignment exactly with
``scipy.optimize.linear_sum_assignment(maximize=True)``; the constraints the
reference adds (row sums <= 1, column sums <= 1, total == n_min) are those of
that assignment problem, and are asserted rather than interpreted.
"""
import re

import numpy as np
from scipy.optimize import linear_sum_assignment

_NAME = re.compile(r"x\[(\d+),(\d+)\]")


class _Status:
    OPTIMAL = 2
    INFEASIBLE = 3


class GRB:
    BINARY = "B"
    MAXIMIZE = -1
    MINIMIZE = 1
    Status = _Status
    OPTIMAL = 2
    INFEASIBLE = 3


class LinExpr:
    """Sparse linear expression: list of (coefficient, Var)."""

    __slots__ = ("terms",)

    def __init__(self, terms=None):
        self.terms = terms if terms is not None else []

    def __iadd__(self, other):
        if isinstance(other, LinExpr):
            self.terms.extend(other.terms)
        elif isinstance(other, Var):
            self.terms.append((1.0, other))
        elif other == 0:
            pass
        else:
            raise TypeError(other)
        return self

    def __add__(self, other):
        out = LinExpr(list(self.terms))
        out += other
        return out

    __radd__ = __add__

    def __le__(self, rhs):
        return ("<=", self, rhs)

    def __eq__(self, rhs):  # noqa: D105 - constraint builder, like gurobipy
        return ("==", self, rhs)

    __hash__ = None


class Var:
    __slots__ = ("i", "j", "x", "vtype")

    def __init__(self, i, j, vtype):
        self.i, self.j, self.vtype, self.x = i, j, vtype, 0.0

    def __mul__(self, coeff):
        return LinExpr([(float(coeff), self)])

    __rmul__ = __mul__


def quicksum(items):
    out = LinExpr()
    for it in items:
        out += it
    return out


class _Params:
    MIPFocus = 0


class Model:
    def __init__(self, name=""):
        self.name = name
        self._vars = []
        self._constrs = []
        self._obj = None
        self._sense = GRB.MAXIMIZE
        self.Params = _Params()
        self.IsMIP = 1
        self.status = GRB.Status.OPTIMAL
        self.objVal = float("nan")

    def addVar(self, vtype=GRB.BINARY, name=""):
        m = _NAME.fullmatch(name)
        if m is None:
            raise ValueError("stand-in expects variables named x[i,j]: %r" % name)
        v = Var(int(m.group(1)), int(m.group(2)), vtype)
        self._vars.append(v)
        return v

    def addConstr(self, constr, name=""):
        self._constrs.append((name, constr))

    def update(self):
        pass

    def setObjective(self, expr, sense):
        self._obj, self._sense = expr, sense

    def optimize(self):
        assert self._sense == GRB.MAXIMIZE
        nr = 1 + max(v.i for v in self._vars)
        nc = 1 + max(v.j for v in self._vars)
        assert len(self._vars) == nr * nc
        # the reference's constraint system: nr row<=1, nc col<=1, one global == n_min
        assert len(self._constrs) == nr + nc + 1
        op, _, rhs = self._constrs[-1][1]
        assert op == "==" and rhs == min(nr, nc)
        cost = np.zeros((nr, nc))
        for coeff, v in self._obj.terms:
            cost[v.i, v.j] += coeff
        if not np.isfinite(cost).all():
            raise ValueError("matrix contains invalid numeric entries")
        r, c = linear_sum_assignment(cost, maximize=True)
        for v in self._vars:
            v.x = 0.0
        by_ij = {(v.i, v.j): v for v in self._vars}
        for i, j in zip(r, c):
            by_ij[(int(i), int(j))].x = 1.0
        self.objVal = float(cost[r, c].sum())
        self.status = GRB.Status.OPTIMAL


__all__ = ["GRB", "LinExpr", "Model", "Var", "quicksum"]
