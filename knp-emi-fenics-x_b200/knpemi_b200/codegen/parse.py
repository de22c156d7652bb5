"""Python source of a model's ``rhs_numba``  ->  expression DAG.

Reads the module *file* (``ode.__file__``) with :mod:`ast`; nothing is executed,
so neither numba nor numbalsoda is needed to generate a kernel.  The function
looked for is ``rhs_numba(t, states, values, parameters)`` -- the name and the
argument order of the plugin protocol (reference mm_hh.py:138-139).
"""
from __future__ import annotations

import ast
from dataclasses import dataclass

from .ir import CALL1, CALL2, Dag, ModelSourceError

_CALLS = {
    ("math", "exp"): "exp", ("np", "exp"): "exp", ("numpy", "exp"): "exp",
    ("math", "log"): "log", ("np", "log"): "log", ("numpy", "log"): "log",
    ("math", "sqrt"): "sqrt", ("np", "sqrt"): "sqrt", ("numpy", "sqrt"): "sqrt",
    ("math", "pow"): "pow", ("np", "power"): "pow", ("numpy", "power"): "pow",
    ("np", "mod"): "mod", ("numpy", "mod"): "mod", ("math", "fmod"): None,
}
_BIN = {ast.Add: "add", ast.Sub: "sub", ast.Mult: "mul", ast.Div: "div", ast.Pow: "pow",
        ast.Mod: "mod"}
_CMP = {ast.Lt: "lt", ast.LtE: "le", ast.Gt: "gt", ast.GtE: "ge", ast.Eq: "eq", ast.NotEq: "ne"}
# spellings of the extra libm functions: math.*, np.*, and the builtins abs/min/max
_ALIASES = {"abs": "fabs", "absolute": "fabs", "arctan": "atan", "arcsin": "asin", "arccos": "acos",
            "arctan2": "atan2", "minimum": "fmin", "maximum": "fmax", "min": "fmin", "max": "fmax"}


@dataclass
class ParsedModel:
    dag: Dag
    dy: dict            # state column -> node id      (values[c] = ...)
    out: dict           # parameter column -> node id  (parameters[c] = ...)
    source_file: str
    lineno: int


def _is_values_none_guard(stmt: ast.If, values_name: str) -> bool:
    t = stmt.test
    return (isinstance(t, ast.Compare) and isinstance(t.left, ast.Name) and t.left.id == values_name
            and len(t.ops) == 1 and isinstance(t.ops[0], (ast.Is, ast.IsNot))
            and isinstance(t.comparators[0], ast.Constant) and t.comparators[0].value is None)


def find_rhs(tree: ast.Module, func_name: str = "rhs_numba") -> ast.FunctionDef:
    found = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func_name]
    if not found:
        raise ModelSourceError(f"no top-level function {func_name!r} in model source")
    return found[-1]      # later definitions shadow earlier ones, as at import


def parse_model_source(source: str, filename: str = "<model>", func_name: str = "rhs_numba") -> ParsedModel:
    fn = find_rhs(ast.parse(source, filename), func_name)
    argnames = [a.arg for a in fn.args.args]
    if len(argnames) != 4:
        raise ModelSourceError(f"{func_name} must take (t, states, values, parameters)")
    a_t, a_y, a_dy, a_p = argnames

    dag = Dag()
    env: dict[str, int] = {}
    dy: dict[int, int] = {}
    out: dict[int, int] = {}

    def err(node, msg):
        return ModelSourceError(f"{filename}:{getattr(node, 'lineno', '?')}: {msg}")

    def const_index(node) -> int:
        s = node.slice
        if isinstance(s, ast.Constant) and isinstance(s.value, int) and not isinstance(s.value, bool):
            return s.value
        raise err(node, "subscripts must be integer literals")

    def expr(node) -> int:
        if isinstance(node, ast.Constant):
            v = node.value
            if isinstance(v, bool):
                return dag.iconst(int(v))
            if isinstance(v, int):
                return dag.iconst(v)
            if isinstance(v, float):
                return dag.const(v)
            raise err(node, f"unsupported literal {v!r}")
        if isinstance(node, ast.Name):
            if node.id == a_t:
                return dag.time()
            if node.id in env:
                return env[node.id]
            raise err(node, f"name {node.id!r} is not defined before use")
        if isinstance(node, ast.Subscript):
            if not isinstance(node.value, ast.Name):
                raise err(node, "unsupported subscript target")
            c = const_index(node)
            if node.value.id == a_y:
                return dag.state(c)
            if node.value.id == a_p:
                if c in out:
                    raise err(node, f"parameters[{c}] is read after the right-hand side wrote it")
                return dag.param(c)
            raise err(node, f"cannot read {node.value.id}[{c}]")
        if isinstance(node, ast.UnaryOp):
            if isinstance(node.op, ast.USub):
                return simplify_neg(expr(node.operand))
            if isinstance(node.op, ast.UAdd):
                return expr(node.operand)
            raise err(node, "unsupported unary operator")
        if isinstance(node, ast.BinOp):
            op = _BIN.get(type(node.op))
            if op is None:
                raise err(node, f"unsupported operator {type(node.op).__name__}")
            return simplify_bin(op, expr(node.left), expr(node.right))
        if isinstance(node, ast.Compare):
            if len(node.ops) != 1 or type(node.ops[0]) not in _CMP:
                raise err(node, "only single <, <=, >, >=, ==, != comparisons are supported")
            return dag.binary(_CMP[type(node.ops[0])], expr(node.left), expr(node.comparators[0]))
        if isinstance(node, ast.IfExp):
            return dag.select(expr(node.test), expr(node.body), expr(node.orelse))
        if isinstance(node, ast.BoolOp):
            vals = [dag.binary("ne", expr(v), dag.iconst(0)) for v in node.values]
            acc = vals[0]
            for v in vals[1:]:
                if isinstance(node.op, ast.And):
                    acc = dag.binary("mul", acc, v)                       # both 0/1
                else:
                    acc = dag.binary("gt", dag.binary("add", acc, v), dag.iconst(0))
            return acc
        if isinstance(node, ast.Call):
            f = node.func
            key = None
            if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name):
                key = (f.value.id, f.attr)
            elif isinstance(f, ast.Name):
                key = ("math", f.id)
            op = _CALLS.get(key)
            if op is None and key is not None and key[0] in ("math", "np", "numpy") and not node.keywords:
                name = _ALIASES.get(key[1], key[1])
                args = [expr(a) for a in node.args]
                if name in CALL1 and len(args) == 1:
                    return dag.call1(name, args[0])
                if name in CALL2 and len(args) == 2:
                    return dag.call2(name, args[0], args[1])
                if name == "where" and len(args) == 3:
                    return dag.select(args[0], args[1], args[2])
            if op is None or node.keywords:
                raise err(node, f"unsupported call {ast.unparse(f)}")
            args = [expr(a) for a in node.args]
            if op in ("exp", "log", "sqrt"):
                if len(args) != 1:
                    raise err(node, f"{op} takes one argument")
                return dag.unary(op, args[0])
            if len(args) != 2:
                raise err(node, f"{op} takes two arguments")
            return dag.binary(op, args[0], args[1])
        raise err(node, f"unsupported expression {type(node).__name__}")

    def simplify_neg(a: int) -> int:
        return dag.unary("neg", a)

    def simplify_bin(op: str, a: int, b: int) -> int:
        # x*1, 1*x and x-0.0 are exact identities in IEEE arithmetic
        if op == "mul":
            if dag.is_const(a) and dag.fvalue(a) == 1.0:
                return b
            if dag.is_const(b) and dag.fvalue(b) == 1.0:
                return a
        if op == "sub" and dag.is_const(b) and dag.fvalue(b) == 0.0 and not dag.is_const(a):
            import math
            if math.copysign(1.0, dag.fvalue(b)) > 0:
                return a
        return dag.binary(op, a, b)

    for stmt in fn.body:
        if isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant) \
                and isinstance(stmt.value.value, str):
            continue          # docstring, or a triple-quoted block used as a comment
        if isinstance(stmt, (ast.Pass, ast.Assert, ast.Import, ast.ImportFrom)):
            continue          # Gotran emits `assert(len(states) == 4)` guards
        if isinstance(stmt, ast.Return):
            break
        if isinstance(stmt, ast.If) and _is_values_none_guard(stmt, a_dy):
            continue          # Gotran's `if values is None: values = np.zeros(...) else: assert ...`
        if isinstance(stmt, ast.AugAssign) and isinstance(stmt.target, ast.Name) \
                and type(stmt.op) in _BIN and stmt.target.id in env:
            env[stmt.target.id] = simplify_bin(_BIN[type(stmt.op)], env[stmt.target.id], expr(stmt.value))
            continue
        if not isinstance(stmt, ast.Assign) or len(stmt.targets) != 1:
            raise err(stmt, f"only simple assignments are supported, got {type(stmt).__name__}")
        # `m, h, n, V = states` / `g_Na, g_K = parameters` (Gotran's unpacking of the arguments)
        if isinstance(stmt.targets[0], (ast.Tuple, ast.List)) and isinstance(stmt.value, ast.Name) \
                and stmt.value.id in (a_y, a_p) \
                and all(isinstance(e, ast.Name) for e in stmt.targets[0].elts):
            for c, e in enumerate(stmt.targets[0].elts):
                nid = dag.state(c) if stmt.value.id == a_y else dag.param(c)
                env[e.id] = nid
                dag.names.setdefault(nid, e.id)
            continue
        target = stmt.targets[0]
        nid = expr(stmt.value)
        if isinstance(target, ast.Name):
            if target.id in argnames:
                raise err(stmt, f"cannot rebind argument {target.id!r}")
            env[target.id] = nid
            dag.names.setdefault(nid, target.id)
        elif isinstance(target, ast.Subscript) and isinstance(target.value, ast.Name):
            c = const_index(target)
            if target.value.id == a_dy:
                dy[c] = nid
            elif target.value.id == a_p:
                out[c] = nid
            else:
                raise err(stmt, f"cannot assign to {target.value.id}[{c}]")
        else:
            raise err(stmt, "unsupported assignment target")

    if not dy:
        raise ModelSourceError(f"{filename}: {func_name} assigns no values[...]")
    return ParsedModel(dag=dag, dy=dy, out=out, source_file=filename, lineno=fn.lineno)
