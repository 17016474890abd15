"""Pin the oracle end-to-end to the reference's KAT/ vectors (all 100 per level, levels
2/3/5; fixtures in tests/golden/kat_L*.npz made by tools/make_golden.py):
  keygen chain (rho,s1,s2)->(t1,t0)  pins ExpandA + NTT + MULT-ACC + INTT + ADD
  sign  (rho,K,tr,M,s1,s2,t0)->(z,h,c)  pins A*y, all INTTs and the rejection logic (§8c-iii)
  verify (rho,c,z,t1,h,M)               pins A*z - c*t1*2^13 (§8c-ii)
"""
import numpy as np
import pytest

import oracle_lib as ol

LEVELS = (2, 3, 5)
EXPECTED_ATTEMPTS = {2: (4.21, 17), 3: (4.23, 19), 5: (4.40, 24)}


@pytest.mark.parametrize("level", LEVELS)
def test_keygen_chain_all_kats(oracle, level):
    K = ol.kat(level)
    for i in range(100):
        t1, t0 = oracle.keygen_chain(level, K["rho"][i], K["s1"][i], K["s2"][i])
        assert np.array_equal(t1, K["t1"][i]) and np.array_equal(t0, K["t0"][i]), (level, i)


@pytest.mark.parametrize("level", LEVELS)
def test_full_keygen_all_kats(oracle, level):
    K = ol.kat(level)
    for i in range(100):
        kg = oracle.keygen(level, K["z"][i])
        for f in ("rho", "k", "tr", "s1", "s2", "t1", "t0"):
            assert np.array_equal(kg[f], K[f][i]), (level, i, f)


@pytest.mark.parametrize("level", LEVELS)
def test_sign_all_kats(oracle, level):
    K = ol.kat(level)
    attempts = []
    for i in range(100):
        zp, hp, c, a = oracle.sign(level, K["rho"][i], K["k"][i], K["tr"][i], K["s1"][i], K["s2"][i], K["t0"][i],
                                   K["msgs"][i])
        assert np.array_equal(zp, K["zs"][i]) and np.array_equal(hp, K["h"][i]) and np.array_equal(c, K["c"][i]), (level, i)
        attempts.append(a)
    mean, mx = EXPECTED_ATTEMPTS[level]
    assert abs(np.mean(attempts) - mean) < 0.01 and max(attempts) == mx


@pytest.mark.parametrize("level", LEVELS)
def test_verify_all_kats(oracle, level):
    K = ol.kat(level)
    for i in range(100):
        assert oracle.verify(level, K["rho"][i], K["t1"][i], K["msgs"][i], K["zs"][i], K["h"][i], K["c"][i]) == 0
    bad = K["zs"][0].copy()
    bad[7] ^= 0x10
    assert oracle.verify(level, K["rho"][0], K["t1"][0], K["msgs"][0], bad, K["h"][0], K["c"][0]) == 1
    assert oracle.verify(level, K["rho"][0], K["t1"][0], K["msgs"][1], K["zs"][0], K["h"][0], K["c"][0]) == 1


def test_expand_a_block_count_on_kats(oracle):
    # 5 rate-168 blocks always sufficed on the KAT keys (max consumed 780 B)
    mx = 0
    for level in LEVELS:
        K = ol.kat(level)
        P = ol.PARAMS[level]
        for i in range(0, 100, 10):
            for r in range(P["k"]):
                for c in range(P["l"]):
                    mx = max(mx, oracle.expand_a_blocks(K["rho"][i], r, c))
    assert mx == 5
