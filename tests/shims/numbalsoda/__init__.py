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
