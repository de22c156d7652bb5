"""Expression DAG -> CUDA C++ source of one model library.

The emitted translation unit defines ``struct Model`` with

* ``hoist(p, q)``      parameter-only nodes, once per DOF per PDE step (device)
* ``deriv(y, dy, q, ts)``   state derivatives, straight-line fp64 (device)
* ``outputs(y, o, q, ts)``  the parameter slots the reference RHS writes as a side
  effect -- the channel currents ``I_ch_*`` (mm_hh.py:220-225)            (device)
* ``tonly(t, ts)``     time-only nodes, evaluated with glibc on the host

and instantiates the generic fused kernel of ``csrc/kem_kernel.cuh`` for it.
Every arithmetic node is emitted as exactly one C++ operation on named
temporaries in the association the Python source has, constants as hex-float
literals, so the only differences to the reference's compiled cfunc are FMA
contraction by nvcc and CUDA's libm (both <= 1 ulp per operation).
"""
from __future__ import annotations

import hashlib
import math
import os
from dataclasses import dataclass

from .ir import P, S, T, Dag, ModelSourceError
from .parse import ParsedModel

CODEGEN_VERSION = "13"


@dataclass
class EmitOptions:
    default_block: int = 128
    #: "fast": exp, log, sqrt, x**1.5 and division through the branch-free routines of
    #:         csrc/kem_math.cuh; x/const -> x*(1/const); const/x -> const*rcp(x);
    #:         x/param -> x*rcp(param) with the reciprocal hoisted out of the loop.
    #:         Every replaced operation stays within 1 ulp of the exact result
    #:         (x**1.5 = x*sqrt(x): 1.3 ulp; CUDA's pow is specified to 2 ulp).
    #: "libm": every operation exactly as written, CUDA libm exp and IEEE division
    #:         (the triage build: differs from the oracle only by FMA contraction
    #:         and CUDA-vs-glibc libm).
    math: str = "fast"
    #: right-hand sides with at most this many in-loop operations get the four RK4 stages inlined
    #: (measured on B200 with the branch-free math: +5 % HH, +12 % glial, +3.5 % calibration; the
    #: cap only keeps very large user models from outgrowing the instruction cache)
    unroll_below: int = int(os.environ.get("KNPEMI_UNROLL_BELOW", "600"))
    #: "fast" only: exponentials of affine functions of one state whose slopes are integer
    #: multiples of a common base share one exp and a multiplication chain (codegen/fuse_exp.py;
    #: HH: six exp -> one exp + 8 multiplications).  Costs up to ~max_exp_power * 1.5e-16 relative
    #: accuracy on the rewritten rates, against the path's parity bar of 1e-10.
    fuse_exp: bool = os.environ.get("KNPEMI_FUSE_EXP", "1") != "0"
    max_exp_power: int = 96
    #: also share exponentials whose offset depends on parameters (hoisted exp(offset); unbounded)
    fuse_exp_param_offsets: bool = False
    #: "fast" only: a/b -> a*rcp(b) (1.5 ulp, 4 FP64 instructions) instead of the residual-corrected
    #: kem::div (0.5 ulp, 6 instructions) for in-loop divisions by state-dependent values; nothing in
    #: the path amplifies a quotient's last bit (the removable singularities amplify the error of
    #: exp(x) - 1, which both forms receive unchanged)
    exact_div: bool = os.environ.get("KNPEMI_EXACT_DIV", "0") == "1"
    #: "fast" only: a*(1 - x) - b*x -> a - x*(a + b) (codegen/relax.py)
    relax_gates: bool = os.environ.get("KNPEMI_RELAX_GATES", "1") != "0"
    #: "fast" only, experimental (off): chains of affine operations on one node collapse into one
    #: FMA of that node (codegen/affine.py); checked on the CPU, not yet measured on the device
    collapse_affine: bool = os.environ.get("KNPEMI_COLLAPSE_AFFINE", "0") == "1"


@dataclass
class EmittedModel:
    name: str
    source: str
    source_hash: str
    ns: int
    np: int
    out_cols: list
    used_cols: list
    n_tslots: int
    n_hoisted: int
    stats: dict


def _lit(v: float) -> str:
    if math.isnan(v) or math.isinf(v):
        raise ModelSourceError("non-finite constant in model expression")
    s = float(v).hex()
    return f"({s})" if s.startswith("-") else s


class _Emitter:
    def __init__(self, pm: ParsedModel, ns: int, np_: int, opts: EmitOptions):
        self.pm, self.dag, self.ns, self.np, self.opts = pm, pm.dag, ns, np_, opts
        for c in pm.dy:
            if not 0 <= c < ns:
                raise ModelSourceError(f"values[{c}] outside the {ns} states of the model")
        missing = sorted(set(range(ns)) - set(pm.dy))
        if missing:
            raise ModelSourceError(f"right-hand side never assigns values{missing}")
        for c in pm.out:
            if not 0 <= c < np_:
                raise ModelSourceError(f"parameters[{c}] outside the {np_} parameters of the model")
        self.out_cols = sorted(pm.out)
        if opts.math not in ("fast", "libm"):
            raise ValueError(f"unknown math mode {opts.math!r}")
        self.fast = opts.math == "fast"
        self.rcp_hoist = []      # parameter-only divisors of in-loop divisions
        self.const_table = []    # device literals that are not 32-bit immediates

    def dev_lit(self, v: float) -> str:
        """Device-side constant.  A double whose low mantissa word is zero (1.0, 25.0,
        4000.0, ...) is a free 32-bit immediate of DFMA/DADD/DMUL; every other literal
        would be rebuilt with two UMOV per use inside the sub-step loop, so those go to
        a __constant__ table that ptxas loads into uniform registers once."""
        import struct
        v = float(v)
        if math.isnan(v) or math.isinf(v):
            raise ModelSourceError("non-finite constant in model expression")
        if struct.unpack("<Q", struct.pack("<d", v))[0] & 0xFFFFFFFF == 0:
            return _lit(v)
        key = v.hex()
        for k, (kk, _) in enumerate(self.const_table):
            if kk == key:
                return f"KC[{k}]"
        self.const_table.append((key, v))
        return f"KC[{len(self.const_table) - 1}]"

    # ------------------------------------------------------------------ naming
    def nm(self, nid: int) -> str:
        return f"v{nid}"

    def comment(self, nid: int) -> str:
        name = self.dag.names.get(nid)
        return f"  // {name}" if name else ""

    # ------------------------------------------------------- where a node lives
    def klass(self, nid: int) -> str:
        d = self.dag.deps[nid]
        if not d:
            return "const"
        if d == frozenset(P):
            return "hoist"
        if d == frozenset(T):
            return "time"
        return "dyn"

    def operand(self, nid: int, ctx: str) -> str:
        """C++ text for reading node ``nid`` from code section ``ctx``."""
        n = self.dag.nodes[nid]
        if n.op in ("const", "iconst"):
            return _lit(float(n.val)) if ctx == "time" else self.dev_lit(float(n.val))
        k = self.klass(nid)
        if ctx == "dyn":
            if n.op == "state":
                return f"y[{n.val}]"
            if k == "hoist":
                return f"q.{self.nm(nid)}"
            if k == "time":
                return f"ts[{self.tslot[nid]}]"
            return self.nm(nid)
        if ctx == "hoist":
            if n.op == "param":
                return f"p[{n.val}]"
            return self.nm(nid)
        if ctx in ("time", "time_dev"):
            if n.op == "time":
                return "t"
            return self.nm(nid)
        raise AssertionError(ctx)

    def rhs_text(self, nid: int, ctx: str) -> str:
        n = self.dag.nodes[nid]
        a = [self.operand(c, ctx) for c in n.args]
        op = n.op
        if op == "add":
            return f"{a[0]} + {a[1]}"
        if op == "sub":
            return f"{a[0]} - {a[1]}"
        if op == "mul":
            return f"{a[0]} * {a[1]}"
        if op == "div":
            num, den = n.args
            if self.dag.is_const(den) and self.dag.fvalue(den) == 0.0:
                raise ModelSourceError("division by the literal 0 in the model's right-hand side")
            # The branch-free kem::rcp / kem::div trade the IEEE special cases (zero, infinite,
            # denormal divisors give NaN) for speed inside the sub-step loop.  The parameter-only
            # section runs once per DOF per PDE step: there the plain IEEE division costs nothing
            # measurable and x/0 stays +-inf, c/(1+inf) stays 0 like in the reference's cfunc.
            if self.fast and ctx == "dyn":
                if self.dag.is_const(den):
                    return f"{a[0]} * {self.dev_lit(1.0 / self.dag.fvalue(den))}"
                if den in self.rcp_hoist:
                    return f"{a[0]} * q.r{den}"
                if self.dag.is_const(num):
                    if self.dag.fvalue(num) == 1.0:
                        return f"kem::rcp({a[1]})"
                    return f"{a[0]} * kem::rcp({a[1]})"
                if self.opts.exact_div:
                    return f"kem::div({a[0]}, {a[1]})"
                return f"{a[0]} * kem::rcp({a[1]})"
            return f"{a[0]} / {a[1]}"
        if op == "neg":
            return f"-{a[0]}"
        if op in ("exp", "log", "sqrt") and self.fast and ctx != "time":
            return f"kem::{op}({a[0]})"
        if op in ("exp", "log", "sqrt"):
            return f"{op}({a[0]})"
        if op == "pow":
            if self.fast and ctx != "time" and self.dag.is_const(n.args[1]):
                e = self.dag.fvalue(n.args[1])
                if e == 1.5:
                    return f"kem::pow15({a[0]})"
                if e == 0.5:
                    return f"kem::sqrt({a[0]})"
            return f"pow({a[0]}, {a[1]})"
        if op == "mod":
            return f"kem_npmod({a[0]}, {a[1]})" if ctx != "time" else f"kem_npmod_host({a[0]}, {a[1]})"
        if op in ("lt", "le", "gt", "ge", "eq", "ne"):
            sym = {"lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eq": "==", "ne": "!="}[op]
            return f"(double)({a[0]} {sym} {a[1]})"
        if op == "call1":
            return f"{n.val}({a[0]})"
        if op == "call2":
            return f"{n.val}({a[0]}, {a[1]})"
        if op == "select":
            return f"(({a[0]}) != 0.0 ? ({a[1]}) : ({a[2]}))"
        raise AssertionError(op)

    def powi_lines(self, nid: int, ctx: str, indent: str) -> list:
        """numba's exponentiation by squaring, one multiply per line."""
        n = self.dag.nodes[nid]
        e = int(n.val)
        x = self.operand(n.args[0], ctx)
        lines = []
        neg = e < 0
        e = abs(e)
        if e == 0:
            return [f"{indent}const double {self.nm(nid)} = 1.0;"]
        r, a, k = None, x, 0
        while e:
            if e & 1:
                if r is None:
                    r = a
                else:
                    t = f"{self.nm(nid)}_r{k}"
                    lines.append(f"{indent}const double {t} = {r} * {a};")
                    r = t
            e >>= 1
            if e:
                t = f"{self.nm(nid)}_s{k}"
                lines.append(f"{indent}const double {t} = {a} * {a};")
                a = t
            k += 1
        final = f"1.0 / {r}" if neg else r
        lines.append(f"{indent}const double {self.nm(nid)} = {final};{self.comment(nid)}")
        return lines

    def section(self, order, ctx: str, indent: str = "        ") -> list:
        lines = []
        for nid in order:
            n = self.dag.nodes[nid]
            if n.op in ("const", "iconst", "param", "state", "time"):
                continue
            if self.klass(nid) != ("time" if ctx == "time_dev" else ctx):
                continue
            if n.op == "powi":
                lines += self.powi_lines(nid, ctx, indent)
            else:
                lines.append(f"{indent}const double {self.nm(nid)} = {self.rhs_text(nid, ctx)};"
                             f"{self.comment(nid)}")
        return lines

    # -------------------------------------------------------------------- emit
    def emit(self, name: str) -> EmittedModel:
        dag, pm = self.dag, self.pm
        dy_roots = [pm.dy[c] for c in range(self.ns)]
        out_roots = [pm.out[c] for c in self.out_cols]
        order_dy = dag.reachable(dy_roots)
        order_out = dag.reachable(out_roots)
        order_all = dag.reachable(dy_roots + out_roots)

        used_cols = sorted({dag.nodes[n].val for n in order_all if dag.nodes[n].op == "param"})
        for c in used_cols:
            if not 0 <= c < self.np:
                raise ModelSourceError(f"parameters[{c}] outside the {self.np} parameters of the model")

        # frontier: hoisted / time nodes a device "dyn" node (or an output root) reads directly
        hoist_front, time_front = [], []

        def note(nid):
            k = self.klass(nid)
            if k == "hoist" and nid not in hoist_front:
                hoist_front.append(nid)
            elif k == "time" and nid not in time_front:
                time_front.append(nid)

        for nid in order_all:
            if self.klass(nid) == "dyn":
                node = dag.nodes[nid]
                if self.fast and node.op == "div" and self.klass(node.args[1]) == "hoist":
                    # x / <parameter-only>  ->  x * rcp, reciprocal taken once per PDE step
                    if node.args[1] not in self.rcp_hoist:
                        self.rcp_hoist.append(node.args[1])
                    note(node.args[0])
                    continue
                for c in node.args:
                    note(c)
        for r in dy_roots + out_roots:
            note(r)
        hoist_front.sort()
        time_front.sort()
        self.tslot = {nid: k for k, nid in enumerate(time_front)}

        self.rcp_hoist.sort()
        order_hoist = dag.reachable(hoist_front + self.rcp_hoist)
        order_time = dag.reachable(time_front)

        L = []
        w = L.append
        w(f"// GENERATED by knpemi_b200.codegen v{CODEGEN_VERSION} -- do not edit.")
        shown = "".join(ch if (ch.isalnum() or ch in "._-/") else "?" for ch in str(pm.source_file))
        w(f"// model {name!r} from {shown}:{pm.lineno}")
        w(f"// options: default_block={self.opts.default_block} math={self.opts.math}"
          f" fuse_exp={int(self.fast and self.opts.fuse_exp)}"
          f" exact_div={int(self.opts.exact_div)} relax_gates={int(self.fast and self.opts.relax_gates)}"
          + (" collapse_affine=1" if self.fast and self.opts.collapse_affine else ""))
        w('#include <math.h>')
        w('#include "kem_math.cuh"')
        w('#include "kem_kernel.cuh"')
        w("")
        w("namespace {")
        w("@@CONST_TABLE@@")
        w("struct Model {")
        w(f"    static constexpr int NS = {self.ns}, NP = {self.np}, NOUT = {len(self.out_cols)}, "
          f"NT = {len(time_front)};")
        w(f"    static constexpr int DEFAULT_BLOCK = {self.opts.default_block};")
        w("    static constexpr bool USES_LOG = @@USES_LOG@@;   // load the log table into shared memory")
        n_loop_ops = sum(1 for nid in order_dy if self.klass(nid) == "dyn"
                         and dag.nodes[nid].op not in ("const", "iconst", "param", "state", "time"))
        w(f"    static constexpr int STAGE_UNROLL = {4 if n_loop_ops <= self.opts.unroll_below else 1};"
          f"  // {n_loop_ops} in-loop operations")
        cond = " || ".join(f"c == {c}" for c in used_cols) or "false"
        w(f"    __host__ __device__ static constexpr bool used(int c) {{ return {cond}; }}")
        w("")
        w("    // parameter-only values handed from hoist() to the sub-step loop")
        w("    struct H {")
        for nid in hoist_front:
            w(f"        double {self.nm(nid)};{self.comment(nid)}")
        for nid in self.rcp_hoist:
            w(f"        double r{nid};  // 1 / {self.dag.names.get(nid, self.nm(nid))}")
        if not hoist_front and not self.rcp_hoist:
            w("        double unused;")
        w("    };")
        w("")
        w("    static __device__ __forceinline__ void hoist(const double (&p)[NP > 0 ? NP : 1], H &q)")
        w("    {")
        L.extend(self.section(order_hoist, "hoist"))
        for nid in hoist_front:
            w(f"        q.{self.nm(nid)} = {self.operand(nid, 'hoist')};")
        for nid in self.rcp_hoist:
            w(f"        q.r{nid} = 1.0 / {self.operand(nid, 'hoist')};")
        if not hoist_front and not self.rcp_hoist:
            w("        q.unused = 0.0; (void)p;")
        w("    }")
        w("")
        w("    static __device__ __forceinline__ void deriv(const double (&y)[NS], double (&dy)[NS],")
        w("                                                 const H &q, const double *__restrict__ ts)")
        w("    {")
        w("        (void)q; (void)ts;")
        L.extend(self.section(order_dy, "dyn"))
        for c in range(self.ns):
            w(f"        dy[{c}] = {self.operand(pm.dy[c], 'dyn')};")
        w("    }")
        w("")
        w("    static __device__ __forceinline__ void outputs(const double (&y)[NS],")
        w("                                                   double (&o)[NOUT > 0 ? NOUT : 1],")
        w("                                                   const H &q, const double *__restrict__ ts)")
        w("    {")
        w("        (void)y; (void)q; (void)ts; (void)o;")
        L.extend(self.section(order_out, "dyn"))
        for k, c in enumerate(self.out_cols):
            w(f"        o[{k}] = {self.operand(pm.out[c], 'dyn')};  // parameters[{c}]")
        w("    }")
        w("")
        w("    // device: the same time-only factors, for schemes whose stage times are not known on")
        w("    // the host (error-controlled stepping)")
        w("    static __device__ __forceinline__ void tonly_dev(double t, double *ts)")
        w("    {")
        w("        (void)t; (void)ts;")
        L.extend(self.section(order_time, "time_dev"))
        for nid in time_front:
            w(f"        ts[{self.tslot[nid]}] = {self.operand(nid, 'time_dev')};")
        w("    }")
        w("")
        w("    // host: time-only factors at stage time t (glibc libm, like the reference cfunc)")
        w("    static void tonly(double t, double *ts)")
        w("    {")
        w("        (void)t; (void)ts;")
        L.extend(self.section(order_time, "time"))
        for nid in time_front:
            w(f"        ts[{self.tslot[nid]}] = {self.operand(nid, 'time')};")
        w("    }")
        w("};")
        w("")
        oc = ", ".join(map(str, self.out_cols)) or "0"
        uc = ", ".join(map(str, used_cols)) or "0"
        w(f"const int OUT_COLS[] = {{{oc}}};")
        w(f"const int USED_COLS[] = {{{uc}}};")
        # output slots assigned a literal: nothing has to fetch them from the device
        const_out = [(c, dag.fvalue(pm.out[c])) for c in self.out_cols if dag.is_const(pm.out[c])]
        cc = ", ".join(str(c) for c, _ in const_out) or "0"
        cv = ", ".join(_lit(v) for _, v in const_out) or "0.0"
        w(f"const int CONST_OUT_COLS[] = {{{cc}}};")
        w(f"const double CONST_OUT_VALS[] = {{{cv}}};")
        w("}  // namespace")
        w("")
        if self.const_table:
            rows = ",\n".join(f"    {_lit(v)}  /* [{k}] {v!r} */" for k, (_, v) in enumerate(self.const_table))
            table = ("// literals that are not 32-bit immediates (see _Emitter.dev_lit)\n"
                     f"__constant__ double KC[{len(self.const_table)}] = {{\n{rows}}};\n")
        else:
            table = ""
        body = "\n".join(L).replace("@@CONST_TABLE@@", table)
        body = body.replace("@@USES_LOG@@", "true" if "kem::log(" in body else "false")
        # hash of the code proper: the leading "// ..." lines (generator version, options) are
        # left out so that a generator release that emits the same code keeps the same hash
        code = "\n".join(ln for ln in body.split("\n") if not ln.startswith("// GENERATED by"))
        h = hashlib.sha256(code.encode()).hexdigest()[:16]
        body += (f'\nKEM_DEFINE_MODEL(Model, "{name}", "{h}", OUT_COLS, USED_COLS, {len(used_cols)}, '
                 f'{len(const_out)}, CONST_OUT_COLS, CONST_OUT_VALS)\n')

        def count(order, ctx):
            c = {}
            for nid in order:
                n = dag.nodes[nid]
                if n.op in ("const", "iconst", "param", "state", "time") or self.klass(nid) != ctx:
                    continue
                key = n.op
                if n.op == "powi":
                    e = abs(int(n.val))
                    c["mul"] = c.get("mul", 0) + max(e.bit_length() - 1, 0) + max(bin(e).count("1") - 1, 0)
                    continue
                c[key] = c.get(key, 0) + 1
            return c

        stats = {"deriv": count(order_dy, "dyn"), "outputs": count(order_out, "dyn"),
                 "hoist": count(order_hoist, "hoist"), "time": count(order_time, "time")}
        return EmittedModel(name=name, source=body, source_hash=h, ns=self.ns, np=self.np,
                            out_cols=self.out_cols, used_cols=used_cols, n_tslots=len(time_front),
                            n_hoisted=len(hoist_front), stats=stats)


def emit_model(pm: ParsedModel, name: str, ns: int, np_: int, opts: EmitOptions | None = None) -> EmittedModel:
    opts = opts or EmitOptions()
    fused = []
    if opts.math == "fast" and opts.fuse_exp:
        from .fuse_exp import fuse_exponentials
        pm, fused = fuse_exponentials(pm, opts.max_exp_power, opts.fuse_exp_param_offsets)
    relaxed = []
    if opts.math == "fast" and opts.relax_gates:
        from .relax import relax_gates
        pm, relaxed = relax_gates(pm)
    collapsed = []
    if opts.math == "fast" and opts.collapse_affine:
        from .affine import collapse_affine
        pm, collapsed = collapse_affine(pm)
    em = _Emitter(pm, ns, np_, opts).emit(name)
    em.stats["fused_exp"] = fused
    em.stats["collapsed_affine"] = collapsed
    em.stats["relaxed_gates"] = relaxed
    return em
