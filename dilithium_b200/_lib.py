"""ctypes loader for libdilithium_b200.so (the C ABI in include/dilithium_b200.h).

The library is built in-tree by `make -C dilithium_b200/csrc` (see __graft_entry__.build()).
Loading fails loudly when it is missing: there is no Python or CPU fallback for any
operation of the engine."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdilithium_b200.so")
SHIM_PATH = os.path.join(_HERE, "libdilithium_b200_shim.so")

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_void = ctypes.c_void_p
c_size = ctypes.c_size_t
c_int = ctypes.c_int
c_uint = ctypes.c_uint

# name -> (restype, argtypes); every symbol include/dilithium_b200.h declares
SIGNATURES = {
    "dil_engine_create": (c_int, [ctypes.POINTER(c_void), c_int]),
    "dil_engine_destroy": (c_int, [c_void]),
    "dil_status_string": (ctypes.c_char_p, [c_int]),
    "dil_last_error": (ctypes.c_char_p, [c_void]),
    "dil_engine_device": (c_int, [c_void]),
    "dil_engine_sm_count": (c_int, [c_void]),
    "dil_level_dims": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "dil_engine_launch_count": (ctypes.c_uint64, [c_void]),
    "dil_ntt_dev": (c_int, [c_void, c_void, c_void, c_size, c_void]),
    "dil_invntt_dev": (c_int, [c_void, c_void, c_void, c_size, c_void]),
    "dil_pointwise_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void]),
    "dil_pointwise_acc_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void]),
    "dil_add_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void]),
    "dil_sub_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void]),
    "dil_matvec_dev": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size, c_void]),
    "dil_expand_a_dev": (c_int, [c_void, c_void, c_void, c_size, c_int, c_int, c_void]),
    "dil_matvec_expand_dev": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size, c_uint, c_void]),
    "dil_signcore_dev": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size, c_void]),
    "dil_ntt_host": (c_int, [c_void, c_void, c_size]),
    "dil_invntt_host": (c_int, [c_void, c_void, c_size]),
    "dil_pointwise_host": (c_int, [c_void, c_void, c_void, c_void, c_size]),
    "dil_pointwise_acc_host": (c_int, [c_void, c_void, c_void, c_void, c_size]),
    "dil_add_host": (c_int, [c_void, c_void, c_void, c_void, c_size]),
    "dil_sub_host": (c_int, [c_void, c_void, c_void, c_void, c_size]),
    "dil_matvec_host": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size]),
    "dil_expand_a_host": (c_int, [c_void, c_void, c_void, c_size, c_int, c_int]),
    "dil_matvec_expand_host": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size, c_uint]),
    "dil_signcore_host": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size]),
    "dil_sign_sizes": (c_int, [c_int, ctypes.POINTER(c_size), ctypes.POINTER(c_size)]),
    "dil_sign_key_create": (c_int, [c_void, ctypes.POINTER(c_void), c_int, c_void, c_void, c_void, c_void, c_void, c_void]),
    "dil_sign_key_destroy": (c_int, [c_void, c_void]),
    "dil_sign_batch_host": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_sign_batch_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void, c_void]),
    "dil_sign_batch_dev_begin": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void, c_void]),
    "dil_sign_batch_host_begin": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_sign_batch_finish": (c_int, [c_void, c_void]),
    "dil_sign_multi_host": (c_int, [c_void, c_int] + [c_void] * 8 + [c_size] + [c_void] * 4),
    "dil_sign_multi_dev": (c_int, [c_void, c_int] + [c_void] * 8 + [c_size] + [c_void] * 5),
    "dil_sign_last_rounds": (ctypes.c_uint32, [c_void]),
    "dil_sign_last_slots": (ctypes.c_uint64, [c_void]),
    "dil_sign_key_set_tuning": (c_int, [c_void, c_void]),
    "dil_sign_set_profile": (c_int, [c_void, c_int]),
    "dil_sign_get_profile": (c_int, [c_void, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "dil_verify_key_create": (c_int, [c_void, ctypes.POINTER(c_void), c_int, c_void, c_void]),
    "dil_verify_key_destroy": (c_int, [c_void, c_void]),
    "dil_verify_batch_host": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_verify_batch_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void, c_void]),
    "dil_verify_multi_host": (c_int, [c_void, c_int, c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_verify_multi_dev": (c_int, [c_void, c_int, c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void, c_void]),
    "dil_keygen_batch_host": (c_int, [c_void, c_int, c_void, c_size] + [c_void] * 7),
    "dil_keygen_batch_dev": (c_int, [c_void, c_int, c_void, c_size] + [c_void] * 8),
    "dil_pool_create": (c_int, [ctypes.POINTER(c_void), c_void, c_int]),
    "dil_pool_destroy": (c_int, [c_void]),
    "dil_pool_size": (c_int, [c_void]),
    "dil_pool_engine": (c_void, [c_void, c_int]),
    "dil_pool_sign_key_create": (c_int, [c_void, ctypes.POINTER(c_void), c_int, c_void, c_void, c_void, c_void, c_void, c_void]),
    "dil_pool_sign_key_destroy": (c_int, [c_void, c_void]),
    "dil_pool_sign_batch_host": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_pool_sign_batch_host_begin": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void, c_void, c_void, c_void]),
    "dil_pool_sign_batch_finish": (c_int, [c_void, c_void]),
    "dil_diag_item_rows_threshold": (c_int, [c_size]),
    "dil_diag_keccak_dev": (c_int, [c_void, c_void, c_uint, c_uint, c_void]),
    "dil_invntt_tomont_dev": (c_int, [c_void, c_void, c_void, c_size, c_void]),
    "dil_poly_pointwise_dev": (c_int, [c_void, c_void, c_void, c_void, c_size, c_void]),
    "dil_polyvec_matrix_pointwise_dev": (c_int, [c_void, c_void, c_void, c_void, c_int, c_int, c_size, c_void]),
}

_lib = None


def load():
    """Load the engine library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA engine first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C dilithium_b200/csrc)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
