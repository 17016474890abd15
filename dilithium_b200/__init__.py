"""dilithium_b200 - B200-native batched Dilithium polynomial-arithmetic engine.

The product is `libdilithium_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/dilithium_b200.h).  This package is the thin Python host mirror used by the tests and
bench; it never computes anything itself and fails loudly when the library is missing."""
from .engine import (INTT_OUTPUT, LEVEL_DIMS, N, NTT_INPUT, Q, RHO_PER_ITEM, RHO_SHARED, DilithiumError, Engine, Pool, SignKey, VerifyKey)

__all__ = ["Engine", "Pool", "SignKey", "VerifyKey", "DilithiumError", "Q", "N", "LEVEL_DIMS", "RHO_SHARED", "RHO_PER_ITEM", "NTT_INPUT", "INTT_OUTPUT"]
