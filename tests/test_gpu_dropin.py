"""Drop-in proof on the GPU: the reference's own, unmodified test main
(dilithium-256/reference_code/ref_test_ntt_ntt2x2.cpp, 100 000 forward + 100 000 inverse
differential checks) linked against libdilithium_b200_shim.so runs every ntt()/ntt2x2_ref()/
invntt()/invntt2x2_ref() call on the B200 and must print OK/OK.  The binary is built in the
build container by oracle/Makefile (the GPU box has no reference tree) and travels with the repo."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_test_main_runs_on_engine():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_test_on_engine")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_test_on_engine not built (reference tree absent at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(":OK") == 2, out.stdout


def test_shim_symbols_batch_of_one(oracle):
    """Call the reference-mangled symbols directly (as a C++ caller would after linking the shim)."""
    from dilithium_b200 import _lib
    shim = ctypes.CDLL(_lib.SHIM_PATH)
    rng = np.random.default_rng(5)
    a = rng.integers(0, 8380417, size=256).astype(np.int32)
    b = rng.integers(0, 8380417, size=256).astype(np.int32)
    P = ctypes.POINTER(ctypes.c_int32)
    for sym, ref in (("_Z3nttPi", oracle.ntt), ("_Z10ntt2x2_refPi", oracle.ntt), ("_Z6invnttPi", oracle.invntt),
                     ("_Z13invntt2x2_refPi", oracle.invntt), ("_Z13invntt_tomontPi", oracle.invntt)):
        x = a.copy()
        getattr(shim, sym)(x.ctypes.data_as(P))
        assert np.array_equal(x, ref(a)), sym
    c = a.copy()  # c == a aliasing, as ntt2x2_test.cpp:102 uses it
    shim._Z17pointwise_barrettPiPKiS1_(c.ctypes.data_as(P), c.ctypes.data_as(P), b.ctypes.data_as(P))
    assert np.array_equal(c, oracle.pointwise(a, b)[0])
    a_hat = rng.integers(0, 8380417, size=(16, 256)).astype(np.int32)
    v = rng.integers(0, 8380417, size=(4, 256)).astype(np.int32)
    w = np.empty((4, 256), dtype=np.int32)
    shim._Z24polyvec_matrix_pointwisePiPKiS1_ii(w.ctypes.data_as(P), a_hat.ctypes.data_as(P), v.ctypes.data_as(P), 4, 4)
    assert np.array_equal(w, oracle.matvec(a_hat, v, 4, 4)[0])
