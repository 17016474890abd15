"""Error behaviour of the C ABI on a live engine: bad arguments are refused with DIL_ERR_ARG and
never crash or silently compute; zero-sized batches are no-ops (the reference's functions are
`void`; the engine adds a status channel, SURVEY.md §8b "Errors")."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ARG = -3


@pytest.fixture(scope="module")
def env():
    import torch
    import dilithium_b200 as d
    eng = d.Engine(0)
    return eng, eng._lib, eng._h, torch


def test_null_and_misaligned_pointers(env):
    eng, lib, h, torch = env
    x = torch.zeros((4, 256), dtype=torch.int32, device="cuda")
    P = ctypes.c_void_p
    assert lib.dil_ntt_dev(h, None, P(x.data_ptr()), 4, None) == ARG
    assert lib.dil_ntt_dev(h, P(x.data_ptr()), None, 4, None) == ARG
    assert lib.dil_ntt_dev(h, P(x.data_ptr() + 4), P(x.data_ptr()), 3, None) == ARG      # not 16-byte aligned
    assert lib.dil_invntt_dev(None, P(x.data_ptr()), P(x.data_ptr()), 4, None) == ARG     # no engine
    assert lib.dil_pointwise_dev(h, P(x.data_ptr()), P(x.data_ptr()), None, 4, None) == ARG
    assert lib.dil_ntt_host(h, None, 4) == ARG
    assert lib.dil_ntt_dev(h, P(x.data_ptr()), P(x.data_ptr()), 0, None) == 0             # empty batch: no-op
    assert lib.dil_ntt_host(h, None, 0) == 0


def test_bad_dims_levels_flags(env):
    eng, lib, h, torch = env
    x = torch.zeros((8, 256), dtype=torch.int32, device="cuda")
    rho = torch.zeros(32, dtype=torch.uint8, device="cuda")
    P = ctypes.c_void_p
    assert lib.dil_matvec_dev(h, P(x.data_ptr()), P(x.data_ptr()), P(x.data_ptr()), 9, 4, 1, None) == ARG
    assert lib.dil_matvec_dev(h, P(x.data_ptr()), P(x.data_ptr()), P(x.data_ptr()), 4, 0, 1, None) == ARG
    assert lib.dil_signcore_dev(h, P(x.data_ptr()), P(x.data_ptr()), P(x.data_ptr()), 4, 5, 1, None) == ARG   # not a level shape
    assert lib.dil_matvec_expand_dev(h, P(x.data_ptr()), P(rho.data_ptr()), P(x.data_ptr()), 4, 4, 1, 8, None) == ARG  # unknown flag
    z, hb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.dil_sign_sizes(4, ctypes.byref(z), ctypes.byref(hb)) == ARG
    assert lib.dil_sign_sizes(2, ctypes.byref(z), ctypes.byref(hb)) == 0 and (z.value, hb.value) == (2304, 84)
    assert lib.dil_sign_sizes(3, ctypes.byref(z), ctypes.byref(hb)) == 0 and (z.value, hb.value) == (3200, 61)
    assert lib.dil_sign_sizes(5, ctypes.byref(z), ctypes.byref(hb)) == 0 and (z.value, hb.value) == (4480, 83)
    k = ctypes.c_void_p()
    buf = np.zeros(4096, dtype=np.uint8)
    q = buf.ctypes.data_as(P)
    assert lib.dil_sign_key_create(h, ctypes.byref(k), 4, q, q, q, q, q, q) == ARG
    assert lib.dil_sign_key_create(h, ctypes.byref(k), 2, q, q, q, None, q, q) == ARG
    assert lib.dil_verify_key_create(h, ctypes.byref(k), 7, q, q) == ARG
    assert lib.dil_keygen_batch_host(h, 2, None, 1, q, q, q, q, q, q, q) == ARG
    assert lib.dil_keygen_batch_host(h, 2, q, 0, q, q, q, q, q, q, q) == 0


def test_engine_bad_device_and_launch_counter(env):
    eng, lib, h, torch = env
    e2 = ctypes.c_void_p()
    assert lib.dil_engine_create(ctypes.byref(e2), 4096) == ARG
    assert lib.dil_engine_create(None, 0) == ARG
    before = eng.launch_count
    x = torch.zeros((4, 256), dtype=torch.int32, device="cuda")
    eng.ntt(x)
    eng.invntt(x)
    assert eng.launch_count == before + 2
    assert eng.sm_count == torch.cuda.get_device_properties(0).multi_processor_count


def test_in_place_and_aliasing(env):
    eng, lib, h, torch = env
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randint(0, 8380417, (100, 256), dtype=torch.int32, device="cuda", generator=g)
    b = torch.randint(0, 8380417, (100, 256), dtype=torch.int32, device="cuda", generator=g)
    ref = eng.pointwise_barrett(a, b)
    a2 = a.clone()
    eng.pointwise_barrett(a2, b, c=a2)       # c == a aliasing, as ntt2x2_test.cpp:102 uses it
    assert torch.equal(a2, ref)
    n1 = eng.ntt(a)
    a3 = a.clone()
    eng.ntt(a3, out=a3)
    assert torch.equal(a3, n1)
    eng.invntt(a3, out=a3)
    assert torch.equal(a3, a)
