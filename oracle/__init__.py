"""CPU oracle for the membrane-ODE stage -- TEST INFRASTRUCTURE, not product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``knpemi_b200`` never
does: the product path has no CPU fallback.
"""
