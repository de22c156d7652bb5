"""Relaxation form of first-order gate kinetics.

Hodgkin-Huxley style models write every gating variable as

    dx/dt = alpha*(1 - x) - beta*x          (reference mm_hh.py:196-197, 206, 212-213)

three dependent fp64 instructions after alpha and beta are known (``1 - x``, a product, an
FMA).  The algebraically identical ``alpha - x*(alpha + beta)`` takes two (an addition and an
FMA) and depends on ``x`` only in the last one.  Both forms are sums of two products rounded at
every step, accurate to an ulp of ``max(alpha, beta)``; they differ in the last bit, not by an
amplified error, so the rewrite sits at the same level as the contraction of ``a*b + c`` into an
FMA that nvcc already applies to the reference's expressions.  ``EmitOptions(math="fast")``
only; ``relax_gates=False`` / ``KNPEMI_RELAX_GATES=0`` keeps the written form.
"""
from __future__ import annotations

from .fuse_exp import substitute
from .ir import S
from .parse import ParsedModel


def relax_gates(pm: ParsedModel) -> tuple[ParsedModel, list]:
    """Rewrite ``a*(1 - x) - b*x`` (factors in either order) to ``a - x*(a + b)``.
    Returns the rewritten model and the list of (node, x) it changed."""
    dag = pm.dag
    roots = [pm.dy[c] for c in sorted(pm.dy)] + [pm.out[c] for c in sorted(pm.out)]
    replace, report = {}, []

    def one_minus(nid: int):
        """x if nid is `1 - x`, else None."""
        n = dag.nodes[nid]
        if n.op == "sub" and dag.is_const(n.args[0]) and dag.fvalue(n.args[0]) == 1.0:
            return n.args[1]
        return None

    for nid in dag.reachable(roots):
        n = dag.nodes[nid]
        if n.op != "sub" or S not in dag.deps[nid]:
            continue
        left, right = dag.nodes[n.args[0]], dag.nodes[n.args[1]]
        if left.op != "mul" or right.op != "mul":
            continue
        for a, om in ((left.args[0], left.args[1]), (left.args[1], left.args[0])):
            x = one_minus(om)
            if x is None:
                continue
            for b, xx in ((right.args[0], right.args[1]), (right.args[1], right.args[0])):
                if xx == x and a != x and b != x:
                    total = dag.binary("add", a, b)
                    new = dag.binary("sub", a, dag.binary("mul", x, total))
                    if nid in dag.names:
                        dag.names.setdefault(new, dag.names[nid])
                    replace[nid] = new
                    report.append((nid, x))
                    break
            if nid in replace:
                break
    if not replace:
        return pm, []
    return substitute(pm, replace), report
