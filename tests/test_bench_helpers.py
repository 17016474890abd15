"""Host-only checks of bench.py's bookkeeping: algorithmic bytes per slot, the Keccak compute roofline and the
committed ncu-derived traffic table all cover the kernel classes the signing pipeline reports."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_class_tables_cover_profile_classes():
    b = load_bench()
    classes = {"expand_mask", "signcore", "challenge", "tail", "resolve"}
    for level in (2, 3, 5):
        cb = b.class_bytes(level)
        assert classes <= set(cb), (level, set(cb))
        k, l = b.LEVEL_DIMS[level]
        assert cb["expand_mask"] == 66 + l * 1024
        assert cb["signcore"] == (l + k) * 1024 + k * b.LEVEL_EXTRA[level]["w1"]
        for c in classes:
            assert b.class_kernel_name(c, level)
    assert "sparse" in b.class_kernel_name("tail", 2) and "sparse" not in b.class_kernel_name("tail", 3)


def test_keccak_roofline_counts():
    b = load_bench()
    # level 2: 4 polynomials x 5 blocks; challenge absorbs 64 + 768 bytes (7 permutations) + SampleInBall (1)
    r = b.keccak_roofline("expand_mask", 2, 65536, 0.333, 4.27, "a measurement")
    assert r["permutations_per_slot"] == 20 and 0.5 < r["frac"] < 1.0 and r["peak_source"] == "a measurement"
    assert "frac" not in b.keccak_roofline("expand_mask", 2, 65536, 0.333)      # no peak measured -> no fraction claimed
    assert b.keccak_roofline("challenge", 2, 65536, 0.165)["permutations_per_slot"] == 8
    assert b.keccak_roofline("challenge", 5, 65536, 0.2)["permutations_per_slot"] == (64 + 8 * 128) // 136 + 2
    assert b.keccak_roofline("tail", 2, 65536, 0.2) is None


def test_limiter_comes_from_the_committed_ncu_table():
    b = load_bench()
    lim = b.ncu_limiter("expand_mask")
    assert lim["top_pipe"] == "alu" and lim["pct_busy"]["alu"] > 90 and "profiles/" in lim["source"]
    assert b.ncu_limiter("signcore")["top_pipe"] == "fmaheavy"
    assert b.ncu_limiter("no such class") is None


def test_keccak_sass_mix_of_the_shipped_library():
    """The ALU-pipe model of the Keccak peak counts the round body's instructions in the shipped .so (skipped without cuobjdump)."""
    import shutil
    import pytest
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not installed")
    b = load_bench()
    mix = b.keccak_sass_mix()
    assert mix is not None and 170 <= mix["instructions_per_round"] <= 210 and mix["alu_pipe"] >= 160


def test_committed_traffic_table():
    b = load_bench()
    prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel.json")))
    for cls in ("expand_mask", "signcore", "challenge", "tail"):
        assert prof[cls]["slots_per_launch"] == 65536 and prof[cls]["dram_bytes_per_launch"] > 0
        assert b.ncu_traffic(cls) == prof[cls]["dram_bytes_per_launch"]
        assert os.path.exists(os.path.join(ROOT, prof[cls]["source"].split(" ")[0]))
