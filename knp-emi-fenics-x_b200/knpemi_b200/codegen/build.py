"""Model module -> generated CUDA source -> compiled model library (cached).

``model_library(ode)`` is what ``MembraneModel`` calls where the reference
takes ``ode.rhs_numba.address`` (src/knpemi/odeSolver.py:96).  The library is
compiled in-tree with ``nvcc -gencode arch=compute_100a,code=sm_100a`` into
``knpemi_b200/_generated/`` and keyed by the hash of the generated source, so a
given model is compiled once.  The six builtin models are compiled ahead of
time by ``__graft_entry__.build()``.

There is no fallback: if the library is not cached and nvcc is not available,
``model_library`` raises.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile

from .emit import CODEGEN_VERSION, EmitOptions, EmittedModel, emit_model
from .ir import ModelSourceError
from .parse import parse_model_source

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PROJECT = os.path.dirname(_PKG)
CSRC_DIR = os.path.join(_PROJECT, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(_PROJECT), "include")
GENERATED_DIR = os.path.join(_PKG, "_generated")
LIB_DIR = os.path.join(_PROJECT, "lib")
RUNTIME_LIB = os.path.join(LIB_DIR, "libknpemi_b200.so")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "--shared", "-Xcompiler", "-fPIC"]


class BuildError(RuntimeError):
    pass


def find_nvcc() -> str:
    for cand in (os.environ.get("KNPEMI_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise BuildError("nvcc not found: a generated model library cannot be compiled "
                     "(set KNPEMI_NVCC, or pre-build with __graft_entry__.build())")


def _run(cmd, what):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise BuildError(f"{what} failed ({' '.join(cmd)}):\n{r.stdout}\n{r.stderr}")
    return r.stdout + r.stderr


def build_runtime(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/kem_runtime.cu -> lib/libknpemi_b200.so (sm_100a)."""
    src = os.path.join(CSRC_DIR, "kem_runtime.cu")
    deps = [src, os.path.join(CSRC_DIR, "kem_model_api.h"), os.path.join(CSRC_DIR, "kem_copy_pool.h"),
            os.path.join(INCLUDE_DIR, "knpemi_b200.h")]
    key = _digest(*[_read(d) for d in deps], " ".join(NVCC_ARCH + NVCC_COMMON))
    stamp = RUNTIME_LIB + ".key"
    if not force and os.path.exists(RUNTIME_LIB) and os.path.exists(stamp) \
            and _read(stamp).decode().strip() == key:
        return RUNTIME_LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc(), *NVCC_ARCH, *NVCC_COMMON, "-Xptxas", "-v", "-I", INCLUDE_DIR,
           "-o", RUNTIME_LIB, src, "-ldl"]
    out = _run(cmd, "runtime build")
    with open(stamp, "w") as f:
        f.write(key)
    if verbose:
        print(out)
    return RUNTIME_LIB


def model_name(ode) -> str:
    return getattr(ode, "__name__", "model").rsplit(".", 1)[-1]


def model_source_file(ode) -> str:
    path = getattr(ode, "__file__", None)
    if not path or not os.path.exists(path):
        raise ModelSourceError(f"model module {ode!r} has no readable __file__; the RHS->CUDA "
                               "generator needs the Python source of rhs_numba")
    return path


def generate(ode, opts: EmitOptions | None = None) -> EmittedModel:
    """Parse ``ode``'s source file and emit the CUDA translation unit."""
    path = model_source_file(ode)
    with open(path, "r") as f:
        source = f.read()
    ns = len(ode.init_state_values())
    np_ = len(ode.init_parameter_values())
    pm = parse_model_source(source, filename=os.path.basename(path))
    return emit_model(pm, model_name(ode), ns, np_, opts)


def generate_from_source(source: str, name: str, ns: int, np_: int, opts: EmitOptions | None = None,
                         filename: str = "<string>") -> EmittedModel:
    return emit_model(parse_model_source(source, filename=filename), name, ns, np_, opts)


def _digest(*parts) -> str:
    import hashlib
    h = hashlib.sha256()
    for part in parts:
        h.update(part if isinstance(part, bytes) else str(part).encode())
        h.update(b"\0")
    return h.hexdigest()


def _read(path) -> bytes:
    with open(path, "rb") as f:
        return f.read()


def _build_key(em: EmittedModel, extra_flags=()) -> str:
    """Content key of a model library: generated source + kernel headers + flags.

    Content-addressed (no mtimes), so libraries pre-built by build() are reused
    on any box the tree is copied to."""
    return _digest(em.source, _read(os.path.join(CSRC_DIR, "kem_kernel.cuh")),
                   _read(os.path.join(CSRC_DIR, "kem_model_api.h")),
                   _read(os.path.join(CSRC_DIR, "kem_math.cuh")),
                   " ".join(NVCC_ARCH + NVCC_COMMON), " ".join(extra_flags))[:16]


def library_path(em: EmittedModel, extra_flags=()) -> str:
    return os.path.join(GENERATED_DIR, f"libkem_{em.name}_{_build_key(em, extra_flags)}.so")


def compile_model(em: EmittedModel, extra_flags=(), force: bool = False, keep_source: bool = True,
                  verbose: bool = False) -> str:
    """Compile an emitted model into its cached shared library; returns the path."""
    out = library_path(em, extra_flags)
    if not force and os.path.exists(out):
        return out
    os.makedirs(GENERATED_DIR, exist_ok=True)
    cu = out[:-3] + ".cu"
    with open(cu, "w") as f:
        f.write(em.source)
    fd, tmp = tempfile.mkstemp(suffix=".so", dir=GENERATED_DIR)
    os.close(fd)
    try:
        cmd = [find_nvcc(), *NVCC_ARCH, *NVCC_COMMON, "-Xcompiler", "-fvisibility=hidden",
               "-Xptxas", "-v", "-I", CSRC_DIR, *extra_flags, "-o", tmp, cu]
        log = _run(cmd, f"model {em.name!r} build")
        with open(out[:-3] + ".ptxas.log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    if verbose:
        print(log)
    if not keep_source:
        os.unlink(cu)
    return out


def model_library(ode, opts: EmitOptions | None = None, extra_flags=()) -> tuple[str, EmittedModel]:
    """Generated + compiled library for a model module (cached by source hash)."""
    em = generate(ode, opts)
    return compile_model(em, extra_flags=extra_flags), em


__all__ = ["BuildError", "CODEGEN_VERSION", "EmitOptions", "build_runtime", "compile_model",
           "generate", "generate_from_source", "library_path", "model_library", "RUNTIME_LIB"]
