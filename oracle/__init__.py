"""CPU oracle for the morsi hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package (see morsi_oracle.c).
"""
from .oracle import (OPS, Oracle, Reference, build_all, oracle, reference,  # noqa: F401
                     have_reference)
