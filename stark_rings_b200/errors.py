"""Error types mirroring the reference's conventions."""


class StarkRingsError(RuntimeError):
    """CUDA / argument errors reported by libstarkrings_cuda.so."""


class LengthPanic(AssertionError):
    """Wrong slice length handed to crt/icrt: the reference panics (assert_eq!(len, D),
    goldilocks/ntt.rs:136,241; babybear/ntt.rs:144,239; stark_prime/ntt.rs:122,246)."""


class AlgebraError(Exception):
    """linear_algebra/src/error.rs:3-8."""


class DifferentLengths(AlgebraError):
    """AlgebraError::DifferentLengths(usize, usize): 'Unexpected different lengths: {0} and {1}'."""

    def __init__(self, a, b):
        super().__init__("Unexpected different lengths: %d and %d" % (a, b))
        self.lengths = (a, b)
