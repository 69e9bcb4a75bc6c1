"""dbb_b200 -- B200-native (sm_100a) implementation of DBB's data-parallel hot path.

Modules mirror the reference's names (``genomicSignatures``, ``tdNeighbours``, ``gcNeighbours``,
``coverageNeighbours``, ``distributions``) so reference call sites work unchanged.  All compute
goes through ``libdbbcuda.so`` (C ABI, include/dbb_cuda.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
