"""Evaluate a parsed model's expression DAG with Python floats.

Host-side check of the parser / DAG (constant folding, integer-power lowering,
operator association): with IEEE doubles and glibc's libm behind :mod:`math`, the
result must equal the reference's numba cfunc bit for bit.  Used by the CPU
tests; the product never evaluates right-hand sides on the host.
"""
from __future__ import annotations

import math

from .ir import CALL1, CALL2, int_pow, np_mod
from .parse import ParsedModel


def evaluate(pm: ParsedModel, t: float, y, p):
    """Return (dy[ns], p_after[np]) of one right-hand-side evaluation."""
    dag = pm.dag
    roots = [pm.dy[c] for c in sorted(pm.dy)] + [pm.out[c] for c in sorted(pm.out)]
    val = {}
    for nid in dag.reachable(roots):
        n = dag.nodes[nid]
        a = [val[c] for c in n.args]
        op = n.op
        if op in ("const", "iconst"):
            v = float(n.val)
        elif op == "param":
            v = float(p[n.val])
        elif op == "state":
            v = float(y[n.val])
        elif op == "time":
            v = float(t)
        elif op == "add":
            v = a[0] + a[1]
        elif op == "sub":
            v = a[0] - a[1]
        elif op == "mul":
            v = a[0] * a[1]
        elif op == "div":
            v = a[0] / a[1]
        elif op == "neg":
            v = -a[0]
        elif op == "exp":
            v = math.exp(a[0])
        elif op == "log":
            v = math.log(a[0])
        elif op == "sqrt":
            v = math.sqrt(a[0])
        elif op == "pow":
            v = math.pow(a[0], a[1])
        elif op == "powi":
            v = int_pow(a[0], int(n.val))
        elif op == "mod":
            v = np_mod(a[0], a[1])
        elif op in ("lt", "le", "gt", "ge", "eq", "ne"):
            v = float({"lt": a[0] < a[1], "le": a[0] <= a[1], "gt": a[0] > a[1], "ge": a[0] >= a[1],
                       "eq": a[0] == a[1], "ne": a[0] != a[1]}[op])
        elif op == "call1":
            v = float(CALL1[n.val](a[0]))
        elif op == "call2":
            v = float(CALL2[n.val](a[0], a[1]))
        elif op == "select":
            v = a[1] if a[0] != 0.0 else a[2]
        else:
            raise AssertionError(op)
        val[nid] = v
    dy = [val[pm.dy[c]] for c in sorted(pm.dy)]
    p_after = [float(x) for x in p]
    for c, nid in pm.out.items():
        p_after[c] = val[nid]
    return dy, p_after
