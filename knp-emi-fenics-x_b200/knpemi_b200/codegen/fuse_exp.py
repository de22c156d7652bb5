"""Share exponentials between the gating-rate expressions of a right-hand side.

Hodgkin-Huxley style rates (reference mm_hh.py:160-186) are exponentials of affine
functions of the same state, ``exp((25 - u)/10)``, ``exp(-u/18)``, ``exp(-u/20)``,
``exp((30 - u)/10)``, ``exp((10 - u)/10)``, ``exp(-u/80)`` with ``u = 1e3*(V + 65e-3)``:
six ``exp`` at 16 FP64 instructions each out of the right-hand side's ~160.  Because

    exp(k*a*x + b) = exp(a*x)**k * exp(b)

exponentials of the same node can share work; ``exp(b)`` folds to a literal when ``b`` is
constant and is hoisted out of the sub-step loop when it depends on parameters only.
Two rewrites, in this order, per node ``x``:

shift   An exponential whose value is subtracted from something -- the ``exp(x) - 1`` of a
        rate with a removable singularity, ``x/(exp(x) - 1)`` -- stays exactly as written,
        because there the error of exp is amplified by 1/|x|.  Every other exponential of
        the same slope becomes that value times ``exp(b_i - b_0)``: one multiplication,
        about 2 ulp.  (HH: ``exp((30 - u)/10)`` and ``exp((10 - u)/10)`` from
        ``exp((25 - u)/10)``.)
chain   The remaining exponentials, none of which feeds a difference, become integer powers
        of one exponential of the smallest commensurate slope, by a multiplication chain.
        (HH: ``exp(-u/80)``, ``exp(-u/20)``, ``exp(-u/18)`` are powers 9, 36, 40 of
        ``exp(-u/720)``: 7 multiplications.)  A k-th power carries k times the relative
        error of its base plus the chain's roundings: up to ~2 * ``max_power`` * 1.1e-16;
        measured 7e-15 for the HH rates (tests/test_codegen.py).

HH: six exp -> two exp + 12 multiplications.  An earlier version that also derived the
``exp(x) - 1`` exponentials from the chain missed the path's parity bar (2.8e-10 against
1e-10 on 20 000 DOFs x 10 steps: a DOF crossing V = -40 mV within 2e-9 V of a stage point
sees the chain's 1e-14 amplified to 6e-8 on alpha_m), hence the first rule.  One visible
difference remains: *exactly on* the singular point of a shifted rate (V = -55 mV for
alpha_n) the reference evaluates 0/0 = NaN and fails its ``assert success``; the shifted
``exp`` is not exactly 1 there, so the quotient is 0/tiny = 0 for that one stage.

Range: with the smallest slope as base and positive powers, every intermediate of a chain
lies between 1 and the largest original exponential of the group, so a chain overflows or
underflows only where an original ``exp(k*a*x)`` does; a constant shift is folded out only
for |b| <= 40.  Exponentials whose offset depends on parameters (``exp((V - E_K)/k)``) are left
alone unless ``param_offsets`` is set (``EmitOptions(fuse_exp_param_offsets=True)``): the
hoisted factor ``exp(b)`` has no bound the generator could check, and none of the reference's
models needs it.

``EmitOptions(fuse_exp=False)`` (or ``KNPEMI_FUSE_EXP=0``) and the "libm" build leave every
``exp`` as written.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction

from .ir import S, T, Dag, Node
from .parse import ParsedModel

EXP_COST = 16           # FP64-pipe instructions of kem::exp (csrc/kem_math.cuh)
MAX_OFFSET = 40.0       # |b| of a constant shift exp(b) that is still folded out


@dataclass(frozen=True)
class _Affine:
    atom: int | None    # node the expression is affine in (None: no state/time dependence)
    slope: float
    off: int | None     # node id of the offset (constant or parameter-only), None = 0


class _Analysis:
    def __init__(self, dag: Dag):
        self.dag = dag
        self.memo: dict[int, _Affine] = {}

    def _add(self, a, b, sub=False):
        dag = self.dag
        if b is None:
            return a
        if a is None:
            return dag.unary("neg", b) if sub else b
        return dag.binary("sub" if sub else "add", a, b)

    def _scale(self, a, c: float):
        return None if a is None else self.dag.binary("mul", a, self.dag.const(c))

    def of(self, nid: int) -> _Affine:
        r = self.memo.get(nid)
        if r is None:
            r = self.memo[nid] = self._of(nid)
        return r

    def _of(self, nid: int) -> _Affine:
        dag = self.dag
        node, deps = dag.nodes[nid], dag.deps[nid]
        if S not in deps and T not in deps:
            return _Affine(None, 0.0, nid)
        atom = _Affine(nid, 1.0, None)
        if T in deps and S not in deps:
            return atom
        op = node.op
        if op == "neg":
            a = self.of(node.args[0])
            return _Affine(a.atom, -a.slope, self._add(None, a.off, sub=True))
        if op in ("add", "sub"):
            l, r = self.of(node.args[0]), self.of(node.args[1])
            if l.atom is not None and r.atom is not None and l.atom != r.atom:
                return atom
            slope = l.slope - r.slope if op == "sub" else l.slope + r.slope
            if slope == 0.0:
                return atom
            return _Affine(l.atom if l.atom is not None else r.atom, slope,
                           self._add(l.off, r.off, sub=(op == "sub")))
        if op == "mul":
            for x, c in ((node.args[0], node.args[1]), (node.args[1], node.args[0])):
                if dag.is_const(c):
                    a, cv = self.of(x), dag.fvalue(c)
                    if cv == 0.0 or a.atom is None:
                        return atom
                    return _Affine(a.atom, a.slope * cv, self._scale(a.off, cv))
            return atom
        if op == "div" and dag.is_const(node.args[1]) and dag.fvalue(node.args[1]) != 0.0:
            a, cv = self.of(node.args[0]), dag.fvalue(node.args[1])
            if a.atom is None:
                return atom
            return _Affine(a.atom, a.slope / cv, self._scale(a.off, 1.0 / cv))
        return atom


def _commensurate(slopes, max_power: int, max_den: int = 64):
    """(base, [k_i]) with slopes[i] ~= k_i * base, k_i positive integers <= max_power, or None."""
    amin = min(slopes, key=abs)
    fr = []
    for s in slopes:
        q = s / amin
        f = Fraction(q).limit_denominator(max_den)
        if q <= 0 or abs(float(f) - q) > 1e-12 * abs(q):
            return None
        fr.append(f)
    den = 1
    for f in fr:
        den = den * f.denominator // math.gcd(den, f.denominator)
    ks = [int(f * den) for f in fr]
    if max(ks) > max_power:
        return None
    return amin / den, ks


def _chain(targets):
    """Multiplication chain reaching every power in ``targets`` from 1: list of (k, a, b)
    meaning x**k = x**a * x**b, in an order where a and b are already available."""
    have, steps = {1}, []

    def get(k):
        if k in have:
            return
        for a in sorted(have, reverse=True):
            if k - a in have:
                steps.append((k, a, k - a))
                have.add(k)
                return
        if k % 2 == 0:
            get(k // 2)
            steps.append((k, k // 2, k // 2))
        else:
            get(k - 1)
            steps.append((k, k - 1, 1))
        have.add(k)

    for k in sorted(set(targets)):
        get(k)
    return steps


def fuse_exponentials(pm: ParsedModel, max_power: int = 96,
                      param_offsets: bool = False) -> tuple[ParsedModel, list]:
    """Rewrite ``pm`` so that exponentials of commensurate affine functions of one node share
    one ``exp``.  Returns the rewritten model and a report (one dict per fused group)."""
    dag = pm.dag
    roots = [pm.dy[c] for c in sorted(pm.dy)] + [pm.out[c] for c in sorted(pm.out)]
    an = _Analysis(dag)
    by_atom: dict[int, list] = {}
    order = dag.reachable(roots)
    for nid in order:
        node = dag.nodes[nid]
        if node.op != "exp" or S not in dag.deps[nid]:
            continue
        a = an.of(node.args[0])
        if a.atom is None or a.atom == node.args[0] or a.slope == 0.0 or not math.isfinite(a.slope):
            continue
        if a.off is not None and dag.is_const(a.off) and abs(dag.fvalue(a.off)) > MAX_OFFSET:
            continue
        if a.off is not None and not dag.is_const(a.off) and not param_offsets:
            continue          # exp(b) of a parameter-dependent b has no bound known here
        by_atom.setdefault(a.atom, []).append((nid, a))

    # who reads each node: an exponential whose value is subtracted from something (the
    # `exp(x) - 1` of a removable singularity) must keep the accuracy of a directly evaluated exp
    parents: dict[int, list] = {}
    for nid in order:
        for c in dag.nodes[nid].args:
            parents.setdefault(c, []).append(nid)

    def sensitive(nid: int) -> bool:
        for p in parents.get(nid, []):
            pn = dag.nodes[p]
            if pn.op in ("mul", "div", "neg"):
                continue
            if pn.op == "add":
                other = pn.args[1] if pn.args[0] == nid else pn.args[0]
                if dag.is_const(other) and dag.fvalue(other) >= 0.0:
                    continue
            return True
        return nid in roots

    def nonzero(off) -> bool:
        return off is not None and not (dag.is_const(off) and dag.fvalue(off) == 0.0)

    replace: dict[int, int] = {}
    shifted: dict[int, tuple] = {}      # value * exp(offset) node -> (value, exp(offset))
    report = []
    for atom, members in by_atom.items():
        members.sort(key=lambda m: (abs(m[1].slope), m[0]))
        # 1. exponentials that feed a difference stay as written; exponentials of the same slope
        #    become that value times exp(offset difference): one multiplication, about 2 ulp
        rest = []
        bases = [m for m in members if sensitive(m[0])]
        kept: list = []
        for m in bases:
            if not any(abs(m[1].slope - b[1].slope) <= 1e-12 * abs(m[1].slope) for b in kept):
                kept.append(m)
        for m in members:
            if m in kept:
                continue
            b = next((b for b in kept if abs(m[1].slope - b[1].slope) <= 1e-12 * abs(m[1].slope)), None)
            if b is None:
                if not sensitive(m[0]):
                    rest.append(m)
                continue
            diff = an._add(m[1].off, b[1].off, sub=True)
            if diff is not None and dag.is_const(diff) and abs(dag.fvalue(diff)) > MAX_OFFSET:
                continue
            v = b[0]
            if nonzero(diff):
                e = dag.unary("exp", diff)
                v = dag.binary("mul", b[0], e)
                shifted[v] = (b[0], e)
            replace[m[0]] = v
            report.append({"atom": dag.names.get(atom, f"node{atom}"), "kind": "shift", "of": b[0],
                           "exps_replaced": 1, "multiplications": int(nonzero(diff)),
                           "nodes": [(m[0], v)]})
        # 2. the others: powers of one exponential of the smallest commensurate slope
        groups: list[list] = []
        for m in rest:
            for g in groups:
                if _commensurate([x[1].slope for x in g] + [m[1].slope], max_power):
                    g.append(m)
                    break
            else:
                groups.append([m])
        for g in groups:
            base, ks = _commensurate([x[1].slope for x in g], max_power)
            steps = _chain(ks)
            shifts = sum(1 for _, a in g if nonzero(a.off))
            n_distinct = len({(a.slope, a.off) for _, a in g})
            if EXP_COST + 1 + len(steps) + shifts >= (EXP_COST + 1) * n_distinct:
                continue
            pw = {1: dag.unary("exp", dag.binary("mul", atom, dag.const(base)))}
            for k, a, b in steps:
                pw[k] = dag.binary("mul", pw[a], pw[b])
            for (nid, a), k in zip(g, ks):
                v = pw[k]
                if nonzero(a.off):
                    e = dag.unary("exp", a.off)
                    v = dag.binary("mul", v, e)
                    shifted[v] = (pw[k], e)
                replace[nid] = v
            report.append({"atom": dag.names.get(atom, f"node{atom}"), "kind": "chain", "base_slope": base,
                           "powers": sorted(ks), "exps_replaced": len(g),
                           "multiplications": len(steps) + shifts,
                           "nodes": [(nid, replace[nid]) for nid, _ in g]})
    if not replace:
        return pm, report
    return substitute(pm, replace, shifted), report


def substitute(pm: ParsedModel, replace: dict, shifted: dict | None = None) -> ParsedModel:
    """``pm`` with every node of ``replace`` exchanged for its image and everything above those
    nodes rebuilt (nodes are immutable and hash-consed).  ``shifted`` maps a node of the form
    value * factor to (value, factor): a constant or parameter-only scale of such a node joins
    the factor instead of costing a second multiplication."""
    dag = pm.dag
    shifted = shifted or {}
    memo: dict[int, int] = {}

    def rebuild(root: int) -> int:
        stack = [root]
        while stack:
            nid = stack[-1]
            if nid in memo:
                stack.pop()
                continue
            if nid in replace and replace[nid] != nid:
                # the image is rebuilt too: it may sit on top of other replaced nodes
                image = replace[nid]
                if image not in memo:
                    stack.append(image)
                    continue
                memo[nid] = memo[image]
                stack.pop()
                continue
            node = dag.nodes[nid]
            todo = [c for c in node.args if c not in memo]
            if todo:
                stack.extend(todo)
                continue
            args = tuple(memo[c] for c in node.args)
            folded = None
            if node.op == "mul":
                # c * (x**k * exp(b))  ->  x**k * (c * exp(b))
                for x, c in ((args[0], args[1]), (args[1], args[0])):
                    if x in shifted and S not in dag.deps[c] and T not in dag.deps[c]:
                        folded = dag.binary("mul", shifted[x][0], dag.binary("mul", shifted[x][1], c))
                        break
            if folded is not None:
                if nid in dag.names:
                    dag.names.setdefault(folded, dag.names[nid])
                memo[nid] = folded
            elif args == node.args:
                memo[nid] = nid
            else:
                deps = frozenset().union(*(dag.deps[c] for c in args))
                new = dag._intern(Node(node.op, args, node.val), deps)
                if nid in dag.names:
                    dag.names.setdefault(new, dag.names[nid])
                memo[nid] = new
            stack.pop()
        return memo[root]

    dy = {c: rebuild(n) for c, n in pm.dy.items()}
    out = {c: rebuild(n) for c, n in pm.out.items()}
    return ParsedModel(dag=dag, dy=dy, out=out, source_file=pm.source_file, lineno=pm.lineno)
