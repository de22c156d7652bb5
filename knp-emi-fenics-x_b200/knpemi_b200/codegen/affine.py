"""Collapse chains of affine operations on one node into a single fused multiply-add.

Gotran-style right-hand sides rescale the membrane potential step by step,
``u = 1e3*(V + 65e-3)``, ``25. - u``, ``(25. - u)/10.``, ``0.1e3*(25. - u)``
(reference mm_hh.py:160-186): four dependent operations before the first rate is touched,
several of them feeding more than one rate.  Every such value is ``a*V + b`` with constant
(or parameter-only) ``a`` and ``b``, so each one that is consumed by non-affine arithmetic can
be produced directly from ``V`` by one FMA: fewer instructions, and the rates start from
independent one-instruction values instead of a shared dependent chain.

Numerics: ``fma(a, V, b)`` rounds the exact ``a*V + b`` once, so the collapsed value is at
least as accurate as the chain it replaces (which rounds at every link); it is *not* bit-equal
to it.  Quantities derived from the same chain value -- numerator ``0.1e3*(25 - u)`` and
exponent ``(25 - u)/10`` of ``x/(exp(x) - 1)`` -- become separately rounded, each to half an
ulp of itself, which keeps their ratio at rounding level also where both pass through zero.

Off by default (``EmitOptions(collapse_affine=True)`` / ``KNPEMI_COLLAPSE_AFFINE=1``): the
pass is checked against the interpreter on the CPU (tests/test_codegen.py) but its parity and
gain on the device have not been measured yet.
"""
from __future__ import annotations

import math

from .fuse_exp import _Analysis, substitute
from .ir import S
from .parse import ParsedModel

_AFFINE_OPS = ("add", "sub", "mul", "div", "neg")


def collapse_affine(pm: ParsedModel) -> tuple[ParsedModel, list]:
    """Returns the rewritten model and a report: one (node, atom, slope, chain length) per
    collapsed value."""
    dag = pm.dag
    roots = [pm.dy[c] for c in sorted(pm.dy)] + [pm.out[c] for c in sorted(pm.out)]
    order = dag.reachable(roots)
    an = _Analysis(dag)

    parents: dict[int, list] = {}
    for nid in order:
        for c in dag.nodes[nid].args:
            parents.setdefault(c, []).append(nid)

    # instructions of the chain between the atom and a node as nvcc would emit it: a
    # multiplication (or negation) that only feeds an addition contracts with it into one FMA
    cost: dict[int, int] = {}

    def chain_cost(nid: int, atom: int) -> int:
        if nid == atom or S not in dag.deps[nid]:
            return 0
        c = cost.get(nid)
        if c is not None:
            return c
        node = dag.nodes[nid]
        if node.op in ("add", "sub"):
            c = 0
            for arg in node.args:
                inner = dag.nodes[arg]
                if S in dag.deps[arg] and arg != atom and inner.op in ("mul", "neg") \
                        and len(parents.get(arg, [])) == 1:
                    c = max(c, max(chain_cost(x, atom) for x in inner.args))     # rides in the FMA
                else:
                    c = max(c, chain_cost(arg, atom))
            c += 1
        else:
            c = 1 + max(chain_cost(x, atom) for x in node.args)
        cost[nid] = c
        return c

    def is_affine(nid: int):
        if dag.nodes[nid].op not in _AFFINE_OPS or S not in dag.deps[nid]:
            return None
        a = an.of(nid)
        if a.atom is None or a.atom == nid or not math.isfinite(a.slope):
            return None
        return a

    replace: dict[int, int] = {}
    report = []
    for nid in order:
        a = is_affine(nid)
        if a is None:
            continue
        # only the outermost affine value of a chain: something non-affine (or a root) reads it
        outer = nid in roots or any(is_affine(p) is None or an.of(p).atom != a.atom
                                    for p in parents.get(nid, []))
        if not outer:
            continue
        n_ops = chain_cost(nid, a.atom)
        if n_ops < 2:
            continue
        if a.slope == 1.0:
            v = a.atom
        elif a.slope == -1.0:
            v = dag.unary("neg", a.atom)
        else:
            v = dag.binary("mul", a.atom, dag.const(a.slope))
        if a.off is not None and not (dag.is_const(a.off) and dag.fvalue(a.off) == 0.0):
            if a.slope == -1.0:
                v = dag.binary("sub", a.off, a.atom)
            else:
                v = dag.binary("add", v, a.off)
        if nid in dag.names:
            dag.names.setdefault(v, dag.names[nid])
        replace[nid] = v
        report.append({"node": nid, "atom": dag.names.get(a.atom, f"node{a.atom}"), "slope": a.slope,
                       "operations_replaced": n_ops})
    if not replace:
        return pm, report
    return substitute(pm, replace), report
