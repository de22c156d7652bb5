"""Expression DAG for membrane-model right-hand sides.

The input language is the straight-line subset the reference's Gotran-style
``rhs_numba`` bodies use (e.g.
examples/idealized_geometries/mm_hh.py:139-227): assignments only; reads of
``states[c]``, ``parameters[c]`` and ``t``; ``+ - * / **`` and unary minus;
``math.exp/log/pow/sqrt``, ``np.exp/log/sqrt/mod``; one comparison used as a
0/1 factor; writes to ``values[c]`` and ``parameters[c]``.

Nodes are hash-consed, so textually repeated sub-expressions (the reference
repeats ``1.0e3*(states[3] + 65.0e-3)`` six times) become one node.  Every node
carries its dependency class, a subset of {"P" parameters, "S" states,
"T" time}; the emitter places a node by that class:

    {}        folded here, with Python floats (IEEE double, glibc libm -- the
              same arithmetic numba/LLVM folds the reference's constants with)
    {P}       hoisted: once per DOF per PDE step, before the sub-step loop
    {T}       evaluated on the host per stage time (glibc), passed as a table
    otherwise evaluated inside the sub-step loop (device)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

P, S, T = "P", "S", "T"

BINOPS = ("add", "sub", "mul", "div", "pow", "mod", "lt", "le", "gt", "ge", "eq", "ne")
UNOPS = ("neg", "exp", "log", "sqrt")

#: further libm functions a model may call (beyond what the reference's models use); they go
#: to CUDA's libm on the device and to Python's math module when constants are folded
CALL1 = {"fabs": math.fabs, "tanh": math.tanh, "sinh": math.sinh, "cosh": math.cosh,
         "sin": math.sin, "cos": math.cos, "tan": math.tan, "atan": math.atan,
         "asin": math.asin, "acos": math.acos, "floor": math.floor, "ceil": math.ceil,
         "log10": math.log10, "log1p": math.log1p, "expm1": math.expm1, "erf": math.erf}
CALL2 = {"fmin": min, "fmax": max, "atan2": math.atan2, "copysign": math.copysign,
         "fmod": math.fmod}


class ModelSourceError(ValueError):
    """The model's right-hand side uses a construct outside the supported subset."""


@dataclass(frozen=True)
class Node:
    op: str                      # const | iconst | param | state | time | <BINOPS> | <UNOPS> | powi
    args: tuple = ()             # child node ids
    val: object = None           # float (const), int (iconst, param/state index, powi exponent)


def np_mod(a: float, b: float) -> float:
    """numpy.mod on doubles: fmod with the result taking the divisor's sign."""
    r = math.fmod(a, b)
    if r != 0.0:
        if (b < 0.0) != (r < 0.0):
            r += b
    else:
        r = math.copysign(0.0, b)
    return r


def int_pow(x: float, n: int) -> float:
    """numba's integer power: exponentiation by squaring, in this order."""
    if n < 0:
        return 1.0 / int_pow(x, -n)
    r = None
    a = x
    while n:
        if n & 1:
            r = a if r is None else r * a
        n >>= 1
        if n:
            a = a * a
    return 1.0 if r is None else r


@dataclass
class Dag:
    nodes: list = field(default_factory=list)
    index: dict = field(default_factory=dict)
    deps: list = field(default_factory=list)
    names: dict = field(default_factory=dict)   # node id -> preferred variable name

    # ------------------------------------------------------------ construction
    def _intern(self, node: Node, deps: frozenset) -> int:
        key = (node.op, node.args, repr(node.val) if node.op == "const" else node.val)
        nid = self.index.get(key)
        if nid is None:
            nid = len(self.nodes)
            self.nodes.append(node)
            self.deps.append(deps)
            self.index[key] = nid
        return nid

    def const(self, v: float) -> int:
        return self._intern(Node("const", (), float(v)), frozenset())

    def iconst(self, v: int) -> int:
        return self._intern(Node("iconst", (), int(v)), frozenset())

    def param(self, c: int) -> int:
        return self._intern(Node("param", (), int(c)), frozenset(P))

    def state(self, c: int) -> int:
        return self._intern(Node("state", (), int(c)), frozenset(S))

    def time(self) -> int:
        return self._intern(Node("time"), frozenset(T))

    def is_const(self, nid: int) -> bool:
        return self.nodes[nid].op in ("const", "iconst")

    def value(self, nid: int):
        return self.nodes[nid].val

    def fvalue(self, nid: int) -> float:
        return float(self.nodes[nid].val)

    def is_int(self, nid: int) -> bool:
        return self.nodes[nid].op == "iconst"

    def unary(self, op: str, a: int) -> int:
        if self.is_const(a):
            if op == "neg":
                v = self.value(a)
                return self.iconst(-v) if self.is_int(a) else self.const(-v)
            x = self.fvalue(a)
            try:
                return self.const({"exp": math.exp, "log": math.log, "sqrt": math.sqrt}[op](x))
            except (ValueError, OverflowError) as e:
                raise ModelSourceError(f"constant {op}({x}) is not finite: {e}")
        return self._intern(Node(op, (a,)), self.deps[a])

    def call1(self, name: str, a: int) -> int:
        if self.is_const(a):
            try:
                return self.const(float(CALL1[name](self.fvalue(a))))
            except (ValueError, OverflowError) as e:
                raise ModelSourceError(f"constant {name}({self.fvalue(a)}) failed: {e}")
        return self._intern(Node("call1", (a,), name), self.deps[a])

    def call2(self, name: str, a: int, b: int) -> int:
        if self.is_const(a) and self.is_const(b):
            return self.const(float(CALL2[name](self.fvalue(a), self.fvalue(b))))
        return self._intern(Node("call2", (a, b), name), self.deps[a] | self.deps[b])

    def select(self, cond: int, a: int, b: int) -> int:
        """`a if cond else b` with cond a 0/1 (or any zero / non-zero) value."""
        if self.is_const(cond):
            return a if self.fvalue(cond) != 0.0 else b
        return self._intern(Node("select", (cond, a, b)), self.deps[cond] | self.deps[a] | self.deps[b])

    def binary(self, op: str, a: int, b: int) -> int:
        if self.is_const(a) and self.is_const(b):
            return self._fold_binary(op, a, b)
        if op == "pow":
            # float ** <integer literal>  and  math.pow(float, <integer literal>)
            # are lowered by numba to repeated squaring, not to libm pow
            if self.is_int(b):
                return self._intern(Node("powi", (a,), int(self.value(b))), self.deps[a])
        return self._intern(Node(op, (a, b)), self.deps[a] | self.deps[b])

    def _fold_binary(self, op: str, a: int, b: int) -> int:
        both_int = self.is_int(a) and self.is_int(b)
        x, y = self.value(a), self.value(b)
        try:
            if op == "add":
                r = x + y
            elif op == "sub":
                r = x - y
            elif op == "mul":
                r = x * y
            elif op == "div":
                r = x / y
                both_int = False
            elif op == "pow":
                if self.is_int(b) and not self.is_int(a):
                    r = int_pow(float(x), int(y))
                elif both_int and y >= 0:
                    r = x ** y
                else:
                    r = math.pow(float(x), float(y))
                    both_int = False
            elif op == "mod":
                r = np_mod(float(x), float(y))
                both_int = False
            elif op in ("lt", "le", "gt", "ge", "eq", "ne"):
                r = {"lt": x < y, "le": x <= y, "gt": x > y, "ge": x >= y, "eq": x == y, "ne": x != y}[op]
                return self.iconst(int(r))
            else:
                raise ModelSourceError(f"cannot fold {op}")
        except (ZeroDivisionError, OverflowError, ValueError) as e:
            raise ModelSourceError(f"constant expression {x} {op} {y} failed: {e}")
        return self.iconst(r) if both_int else self.const(r)

    # --------------------------------------------------------------- traversal
    def reachable(self, roots) -> list:
        """Node ids reachable from ``roots`` in topological (children first) order."""
        seen, order = set(), []
        stack = [(r, False) for r in reversed(list(roots))]
        while stack:
            nid, done = stack.pop()
            if done:
                order.append(nid)
                continue
            if nid in seen:
                continue
            seen.add(nid)
            stack.append((nid, True))
            for c in reversed(self.nodes[nid].args):
                if c not in seen:
                    stack.append((c, False))
        return order

    def canonical(self, roots) -> tuple:
        """Structure-only fingerprint of the sub-DAG under ``roots`` (no names)."""
        memo = {}

        def rec(nid):
            if nid in memo:
                return memo[nid]
            n = self.nodes[nid]
            v = repr(n.val) if n.op == "const" else n.val
            memo[nid] = (n.op, v, tuple(rec(c) for c in n.args))
            return memo[nid]

        import sys
        lim = sys.getrecursionlimit()
        sys.setrecursionlimit(max(lim, 10000))
        try:
            return tuple(rec(r) for r in roots)
        finally:
            sys.setrecursionlimit(lim)
