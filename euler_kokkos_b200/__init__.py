"""euler_kokkos_b200 -- B200-native MUSCL-Hancock time step (Euler + CT-MHD) of pkestene/euler_kokkos.

The product is the C-ABI shared library ``lib/libek_b200.so`` (include/ek_b200.h) plus the C++ host
driver that mirrors the reference's SolverFactory / ConfigMap / SolverBase API (csrc/host).  This
Python package is only a thin ctypes binding used by bench.py and the tests; it contains no
numerics and no CPU fallback: importing :mod:`euler_kokkos_b200.capi` raises if the CUDA library
has not been built.
"""
__version__ = "0.1.0"
