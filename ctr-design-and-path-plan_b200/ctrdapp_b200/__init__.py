"""ctrdapp_b200 -- host-side mirror of the ctrdapp.model / ctrdapp.collision /
ctrdapp.heuristic interfaces (ConorMesser/ctr-design-and-path-plan) over
libctrb200.so, the hand-written sm_100a CUDA hot path.

There is no CPU fallback: importing works anywhere (so the ABI can be
inspected), but every compute call needs a CUDA device and raises
``CtrbError`` without one.
"""
from ._lib import CtrbError, Context, default_context, lib, lib_path  # noqa: F401

__all__ = ["CtrbError", "Context", "default_context", "lib", "lib_path"]
