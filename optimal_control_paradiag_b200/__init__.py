"""B200-native ParaDiag block-circulant preconditioner for the 1-D wave optimal-control
all-at-once system: a drop-in for the ``DiagFFTPC`` hot path of
Molin-Han/Optimal_Control_ParaDiag (``Code/Control_Wave_PC.py``).

The arithmetic lives in ``libparadiag.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/paradiag.h``); this package is the thin Python host that mirrors the
reference's petsc4py / Firedrake python-PC surface.  There is no CPU fallback: importing
works anywhere, but every compute call needs the built library and a CUDA device.
"""
from ._lib import LibraryNotBuilt, ParaDiagError, load_library, library_path  # noqa: F401
from .handle import ParaDiagHandle  # noqa: F401
from .pc import DiagFFTPC, PCBase  # noqa: F401
from .problem import Optimal_Control_Wave_Equation, default_parameters  # noqa: F401

__all__ = [
    "DiagFFTPC", "PCBase", "ParaDiagHandle", "Optimal_Control_Wave_Equation", "default_parameters",
    "load_library", "library_path", "LibraryNotBuilt", "ParaDiagError",
]
