"""vpic_b200 — B200-native particle-advance engine behind lanl/vpic's own API.

Python here is host-side plumbing only: ctypes bindings to the C-ABI library
(libvpic_b200.so, hand-written CUDA for sm_100a) and a thin mirror of the
reference's operator interface used by tests and bench.py.
"""
from . import abi  # noqa: F401
