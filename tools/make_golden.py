#!/usr/bin/env python
"""Generate tests/golden/* from the reference tree (run in the build container only).

  kat_L{2,3,5}.npz  - the reference's KAT/ vectors (KAT/*.txt, 100 per level), hex -> bytes.
                      Field names follow the KAT file stems: rho, k, tr,
                      z (= keygen seed xi), s1, s2, t0, t1, zs (= packed signature z), h, c, mlen.
  kat_msgs.npz      - the 100 messages (first mlen bytes of each m_2.txt line; the m_* files
                      are identical across levels).
  ntt_golden.npz    - seeded inputs + outputs of the reference's own compiled C++
                      (oracle/_ref/libdilref.so: ntt / invntt / pointwise_barrett / ntt2x2_ref /
                      invntt2x2_ref), canonicalised mod Q, plus zetas_barrett[] and zetas.txt.

Nothing here is read at run time on the GPU box; the .npz files are the fixtures.
"""
import ctypes, os, sys
import numpy as np

REF = os.environ.get("DIL_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
Q = 8380417


def hexlines(path):
    with open(path) as f:
        return [bytes.fromhex(line.strip()) for line in f if line.strip()]


def kat():
    for lvl in (2, 3, 5):
        d = {}
        for stem in ("rho", "k", "tr", "z", "s1", "s2", "t0", "t1", "zs", "h", "c"):
            rows = hexlines(f"{REF}/KAT/{stem}_{lvl}.txt")
            assert len(rows) == 100 and len({len(r) for r in rows}) == 1, stem
            d[stem] = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(100, -1)
        d["mlen"] = np.array([int.from_bytes(r, "big") for r in hexlines(f"{REF}/KAT/mlen_{lvl}.txt")], dtype=np.int32)
        np.savez_compressed(os.path.join(OUT, f"kat_L{lvl}.npz"), **d)
    msgs = hexlines(f"{REF}/KAT/m_2.txt")
    for lvl in (3, 5):
        assert msgs == hexlines(f"{REF}/KAT/m_{lvl}.txt")
    mlen = [int.from_bytes(r, "big") for r in hexlines(f"{REF}/KAT/mlen_2.txt")]
    blob = b"".join(m[:n] for m, n in zip(msgs, mlen))
    np.savez_compressed(os.path.join(OUT, "kat_msgs.npz"), blob=np.frombuffer(blob, dtype=np.uint8),
                        mlen=np.array(mlen, dtype=np.int32))


def ntt_golden():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdilref.so"))
    lib.ref_zetas.restype = ctypes.POINTER(ctypes.c_int32)
    zetas = np.ctypeslib.as_array(lib.ref_zetas(), shape=(256,)).copy()
    rom = np.array([int(x, 16) for x in open(f"{REF}/zetas.txt").read().split()], dtype=np.int64)
    rng = np.random.default_rng(0x44494C32)
    n = 96
    x = rng.integers(0, Q, size=(n, 256), dtype=np.int64)
    # edge cases: zeros, all Q-1, unit impulses, negative representatives in (-Q,0)
    x[0] = 0
    x[1] = Q - 1
    for i in range(8):
        x[2 + i] = 0
        x[2 + i, [0, 1, 127, 128, 129, 200, 254, 255][i]] = 1
    x[10:20] -= Q
    x[10:20][x[10:20] == -Q] = 0
    x = x.astype(np.int32)
    y = rng.integers(0, Q, size=(n, 256), dtype=np.int64).astype(np.int32)
    P = ctypes.POINTER(ctypes.c_int32)

    def run(fn, a):
        a = np.ascontiguousarray(a.copy())
        getattr(lib, fn)(a.ctypes.data_as(P), ctypes.c_size_t(a.shape[0]), 1)
        return (a.astype(np.int64) % Q).astype(np.int32)

    out = dict(zetas_barrett=zetas, zetas_rom=rom, x=x, y=y,
               ntt=run("ref_ntt_batch", x), invntt=run("ref_invntt_batch", x),
               ntt2x2=run("ref_ntt2x2_batch", x), invntt2x2=run("ref_invntt2x2_batch", x))
    c = np.zeros_like(x)
    lib.ref_pointwise_batch(c.ctypes.data_as(P), x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_size_t(n), 1)
    out["pointwise"] = (c.astype(np.int64) % Q).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "ntt_golden.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    kat()
    ntt_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
