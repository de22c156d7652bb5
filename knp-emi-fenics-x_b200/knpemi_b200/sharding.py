"""Contiguous DOF ranges over devices / ranks.

Membrane DOFs are independent (each row of the reference's loop touches only its own
``states[row]`` / ``parameters[row]``, odeSolver.py:107-122), so the stage shards by
range with no collective.  ``dof_ranges`` is the same rule ``kem_create`` applies to the
devices of one handle: ``ceil(n / parts)`` DOFs per part, the remainder on the last.
"""
from __future__ import annotations


def dof_ranges(n: int, parts: int) -> list[tuple[int, int]]:
    if parts < 1:
        raise ValueError("parts must be >= 1")
    if n < 0:
        raise ValueError("n must be >= 0")
    per = (n + parts - 1) // parts
    out = []
    for k in range(parts):
        begin = min(k * per, n)
        out.append((begin, min(begin + per, n)))
    return out


def rank_range(n: int, rank: int, world: int) -> tuple[int, int]:
    return dof_ranges(n, world)[rank]
