"""
ORACLE -- test infrastructure, not product code.

A CPU restatement of the COSMOMAP2 map-making hot path (NumPy/SciPy + a plain-C twin of the
reference's weave loops), used as the parity checker and as the timed CPU baseline.  Only
tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import it.
See oracle/operators.py for the pinning statement.
"""
from . import linop_min as lp  # noqa: F401
from .operators import *  # noqa: F401,F403
from .operators import dgemm, norm2, scalprod  # noqa: F401
from .krylov import (arnoldi, build_hess, build_Z, run_krypy_arnoldi,  # noqa: F401
                     find_ritz_eigenvalues, krypy_arnoldi, krypy_ritz)
