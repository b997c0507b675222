"""sci_b200 -- B200-native backend for flomat's dense FP64 hot path (times / expm / LU solves / elementwise).

The product is sci_b200/lib/libflomat_cuda.so (hand-written CUDA for sm_100a behind the C ABI of
include/flomat_cuda.h); this package is the host-side mirror of the reference's Racket API over that ABI.
"""
from . import _lib
from ._lib import FlomatCudaError, init, load
from .flomat import *  # noqa: F401,F403
from .flomat import Flomat, Pivots

__all__ = ["Flomat", "Pivots", "FlomatCudaError", "init", "load"]
