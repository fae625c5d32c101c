"""stark_rings_b200: B200-native (sm_100a) hot path of NethermindEth/stark-rings.

Batched CRT / ICRT, NTT-form slot-wise multiplication, the fused ring multiplication and the ring
matrix x vector product for the Goldilocks, BabyBear and Starknet-prime ring models, computed by
hand-written CUDA kernels behind a C ABI (include/stark_rings_cuda.h, libstarkrings_cuda.so).
This package is the thin host-side mirror of the reference's operator surface; it contains no
CPU implementation of the arithmetic.
"""
from .errors import AlgebraError, DifferentLengths, LengthPanic, StarkRingsError  # noqa: F401
from .rings import (  # noqa: F401
    CONFIGS, CRT, ICRT, BabyBearRingConfig, Context, GoldilocksRingConfig, RingConfig, RqNTT, RqPoly,
    StarkRingConfig, default_context,
)
from .linalg import Matrix, SparseMatrix  # noqa: F401

__version__ = "0.1.0"
