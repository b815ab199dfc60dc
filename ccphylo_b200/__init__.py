"""ccphylo_b200 -- B200 (sm_100a) implementation of ccphylo's `dist` hot path.

The product is the C-ABI shared library built from ``csrc/`` (declared in
``include/ccphylo_gpu.h``); ``api`` is its thin ctypes mirror and ``synth``
generates the synthetic KMA-consensus workloads of SURVEY.md section 8(d).
"""
from . import api, synth  # noqa: F401

__all__ = ["api", "synth"]
