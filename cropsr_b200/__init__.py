"""cropsr_b200 -- CROPSR's Cas9 gRNA candidate scan + Rule-Set-1 scoring on B200.

Python host (ingest, emission plan, CSV) over hand-written sm_100a CUDA kernels
reached through a C ABI (include/cropsr_b200.h, libcropsr_b200.so, ctypes).
Importing the package does not load the CUDA library; ``cropsr_b200.engine``
does, and raises if it has not been built.  There is no CPU fallback.
"""
__version__ = "0.1.0"
