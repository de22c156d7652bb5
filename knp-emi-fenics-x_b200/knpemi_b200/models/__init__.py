"""Builtin membrane model modules (plugin protocol of reference odeSolver.py:8-49).

Each module is a from-scratch restatement of one reference model module; the
CUDA kernels for these six are generated and compiled ahead of time by
``__graft_entry__.build()``.  Any other module following the same protocol
(e.g. the reference's own ``mm_*.py`` files) is accepted by ``MembraneModel``
and compiled on first use by ``knpemi_b200.codegen``.
"""
from . import calibration, glial_bench, glial_tissue, hh_ideal, hh_test, hh_tissue

BUILTIN = {
    "hh_ideal": hh_ideal,
    "hh_tissue": hh_tissue,
    "glial_tissue": glial_tissue,
    "glial_bench": glial_bench,
    "calibration": calibration,
    "hh_test": hh_test,
}

#: reference file each builtin restates (relative to the reference root)
REFERENCE_FILE = {
    "hh_ideal": "examples/idealized_geometries/mm_hh.py",
    "hh_tissue": "examples/local_astrocyte_depolarization/mm_hh.py",
    "glial_tissue": "examples/local_astrocyte_depolarization/mm_glial.py",
    "glial_bench": "examples/benchmark/mm_glial.py",
    "calibration": "examples/calibrate_initial_conditions/mm_calibration.py",
    "hh_test": "tests/mm_test_ode.py",
}
