"""Pin the CPU oracle (oracle/*.c) to the reference: golden vectors produced by the
reference's own compiled C++ (tools/make_golden.py -> tests/golden/ntt_golden.npz), the
reference's twiddle table / ROM image, hashlib for SHAKE, and - when oracle/_ref is
present - a live differential run mirroring ref_test_ntt_ntt2x2.cpp:44-92."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

Q = ol.Q


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ntt_golden.npz"))


def test_zetas_match_reference_table_and_rom(oracle, g):
    z = oracle.zetas()
    assert np.array_equal(z, g["zetas_barrett"])          # consts.cpp:64-97
    assert np.array_equal(z.astype(np.int64) % Q, g["zetas_rom"])  # zetas.txt (RTL ROM image)
    assert z[0] == 0 and z[1] == -3572223 and z[128] == 1753


@pytest.mark.parametrize("name", ["ntt", "invntt", "ntt2x2", "invntt2x2"])
def test_transforms_match_reference_outputs(oracle, g, name):
    out = getattr(oracle, name)(g["x"])
    assert out.min() >= 0 and out.max() < Q                 # canonical contract
    assert np.array_equal(out, g[name])


def test_radix2x2_equals_radix2(g):
    # the reference's own differential claim (ref_test_ntt_ntt2x2.cpp:31-42)
    assert np.array_equal(g["ntt"], g["ntt2x2"])
    assert np.array_equal(g["invntt"], g["invntt2x2"])


def test_pointwise_add_sub(oracle, g):
    assert np.array_equal(oracle.pointwise(g["x"], g["y"]), g["pointwise"])
    x, y = g["x"].astype(np.int64), g["y"].astype(np.int64)
    assert np.array_equal(oracle.addsub(g["x"], g["y"]), ((x + y) % Q).astype(np.int32))
    assert np.array_equal(oracle.addsub(g["x"], g["y"], sub=True), ((x - y) % Q).astype(np.int32))


def test_roundtrip_and_negacyclic_product(oracle):
    rng = np.random.default_rng(7)
    a = rng.integers(0, Q, size=(8, 256)).astype(np.int32)
    b = rng.integers(0, Q, size=(8, 256)).astype(np.int32)
    assert np.array_equal(oracle.invntt(oracle.ntt(a)), a)
    prod = oracle.invntt(oracle.pointwise(oracle.ntt(a), oracle.ntt(b)))
    # schoolbook negacyclic product of row 0
    a0, b0 = a[0].astype(object), b[0].astype(object)
    c = [0] * 256
    for i in range(256):
        for j in range(256):
            k = i + j
            if k < 256:
                c[k] += a0[i] * b0[j]
            else:
                c[k - 256] -= a0[i] * b0[j]
    assert [int(x) % Q for x in c] == prod[0].tolist()


def test_ntt_is_evaluation_at_odd_powers(oracle):
    # ntt(a)[i] = a(zeta^(2*brv8(i)+1))
    rng = np.random.default_rng(3)
    a = rng.integers(0, Q, size=256).astype(np.int32)
    out = oracle.ntt(a)
    for i in (0, 1, 77, 255):
        r = pow(1753, 2 * int(f"{i:08b}"[::-1], 2) + 1, Q)
        acc = 0
        for c in reversed(a.tolist()):
            acc = (acc * r + c) % Q
        assert acc == out[i]


@pytest.mark.parametrize("n", [0, 1, 33, 135, 136, 137, 167, 168, 169, 1000])
def test_shake_against_hashlib(oracle, n):
    data = (bytes(range(256)) * 4)[:n]
    assert oracle.shake_bytes(data, 500, 256) == hashlib.shake_256(data).digest(500)
    assert oracle.shake_bytes(data, 500, 128) == hashlib.shake_128(data).digest(500)


def test_expand_a_against_hashlib(oracle):
    rho = np.arange(32, dtype=np.uint8)
    a = oracle.expand_a(rho, 4, 4)
    for (i, j) in [(0, 0), (1, 3), (3, 2)]:
        stream = hashlib.shake_128(rho.tobytes() + bytes([j, i])).digest(168 * 8)
        vals = []
        for p in range(0, len(stream), 3):
            t = stream[p] | (stream[p + 1] << 8) | ((stream[p + 2] & 0x7F) << 16)
            if t < Q:
                vals.append(t)
        assert vals[:256] == a[i * 4 + j].tolist()


def test_live_differential_vs_compiled_reference(oracle):
    ref = ol.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(11)
    x = rng.integers(0, Q, size=(2000, 256)).astype(np.int32)
    y = rng.integers(-Q + 1, Q, size=(2000, 256)).astype(np.int32)
    assert np.array_equal(ref.zetas(), oracle.zetas())
    for xs in (x, y):
        assert np.array_equal(oracle.ntt(xs), ref.run("ref_ntt_batch", xs))
        assert np.array_equal(oracle.invntt(xs), ref.run("ref_invntt_batch", xs))
        assert np.array_equal(oracle.ntt2x2(xs[:200]), ref.run("ref_ntt2x2_batch", xs[:200]))
        assert np.array_equal(oracle.invntt2x2(xs[:200]), ref.run("ref_invntt2x2_batch", xs[:200]))
    assert np.array_equal(oracle.pointwise(x, y), ref.pointwise(x, y))
    a_hat = rng.integers(0, Q, size=(16, 256)).astype(np.int32)
    w, _ = oracle.signcore(a_hat, x[:64 * 4], 4, 4)
    assert np.array_equal(w, ref.signcore(a_hat, x[:64 * 4], 4, 4))
