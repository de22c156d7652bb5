"""Import shim for the absent third-party package ``numbalsoda``.

TEST INFRASTRUCTURE ONLY.  The reference model modules do
``from numbalsoda import lsoda_sig`` (e.g. examples/idealized_geometries/mm_hh.py:133);
numbalsoda itself (an un-pinned dependency, pyproject.toml:14) is not installed
in this image.  This shim supplies the one symbol the model modules need so the
reference ``mm_*.py`` files import verbatim from /root/reference when golden
vectors are generated (tests/golden/make_golden.py).
"""
from numba import types

lsoda_sig = types.void(types.double,
                       types.CPointer(types.double),
                       types.CPointer(types.double),
                       types.CPointer(types.double))


def lsoda(funcptr, u0, t_eval, data=None, rtol=1.0e-8, atol=1.0e-10):
    """Stand-in for ``numbalsoda.lsoda`` with the call signature of src/knpemi/odeSolver.py:116-120.

    numbalsoda is absent and un-pinned, so the integrator itself cannot be
    reproduced; this shim integrates the row with the normative fixed-step scheme O1 (RK4 x 25 +
    one evaluation at the end point, oracle/knpemi_oracle.c:step_row) through the cfunc address
    it is given -- the reference's own compiled right-hand side.  ``data`` is mutated in place
    by the right-hand side exactly like numbalsoda would let it.  Returns ``(usol, success)``."""
    import numpy as np
    from oracle import cpu_oracle
    t0, t1 = float(t_eval[0]), float(t_eval[-1])
    y = np.array(u0, dtype=np.float64).reshape(1, -1)
    p = data if data is not None else np.zeros(1)
    assert p.dtype == np.float64 and p.flags.c_contiguous
    bad = cpu_oracle.step_fn(funcptr, y, p.reshape(1, -1), t0, t1 - t0, 25, 1)
    return np.vstack([np.asarray(u0, dtype=np.float64), y[0]]), bad == 0
