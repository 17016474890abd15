"""ctypes binding of the CPU oracle (oracle/liboracle.so) and, when present, of the
compiled reference (oracle/_ref/libdilref.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
Q = 8380417
N = 256
I32P = ctypes.POINTER(ctypes.c_int32)
U8P = ctypes.POINTER(ctypes.c_uint8)

PARAMS = {  # level: k, l, eta, tau, gamma1_bits, gamma2, beta, omega
    2: dict(k=4, l=4, eta=2, tau=39, gamma1_bits=17, gamma2=(Q - 1) // 88, beta=78, omega=80),
    3: dict(k=6, l=5, eta=4, tau=49, gamma1_bits=19, gamma2=(Q - 1) // 32, beta=196, omega=55),
    5: dict(k=8, l=7, eta=2, tau=60, gamma1_bits=19, gamma2=(Q - 1) // 32, beta=120, omega=75),
}


def _i32(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(I32P)


def _u8(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(U8P)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_zetas.restype = I32P
        lib.orc_expand_a_poly.restype = ctypes.c_int

    def zetas(self):
        return np.ctypeslib.as_array(self.lib.orc_zetas(), shape=(N,)).copy()

    def _inplace(self, fn, a, threads=None):
        a = np.ascontiguousarray(a, dtype=np.int32).copy().reshape(-1, N)
        if threads is None:
            getattr(self.lib, fn)(_i32(a), ctypes.c_size_t(a.shape[0]))
        else:
            getattr(self.lib, fn)(_i32(a), ctypes.c_size_t(a.shape[0]), ctypes.c_int(threads))
        return a

    def ntt(self, a, threads=None):
        return self._inplace("orc_ntt_batch" if threads is None else "orc_ntt_batch_mt", a, threads).reshape(np.shape(a))

    def invntt(self, a, threads=None):
        return self._inplace("orc_invntt_batch" if threads is None else "orc_invntt_batch_mt", a, threads).reshape(np.shape(a))

    def _single(self, fn, a):
        a = np.ascontiguousarray(a, dtype=np.int32).copy().reshape(-1, N)
        for row in a:
            getattr(self.lib, fn)(_i32(row))
        return a.reshape(np.shape(a))

    def ntt2x2(self, a):
        return self._single("orc_ntt2x2", a)

    def invntt2x2(self, a):
        return self._single("orc_invntt2x2", a)

    def pointwise(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, N)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, N)
        c = np.empty_like(a)
        self.lib.orc_pointwise_batch(_i32(c), _i32(a), _i32(b), ctypes.c_size_t(a.shape[0]))
        return c

    def addsub(self, a, b, sub=False):
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, N)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, N)
        c = np.empty_like(a)
        fn = self.lib.orc_sub if sub else self.lib.orc_add
        for i in range(a.shape[0]):
            fn(_i32(c[i]), _i32(a[i]), _i32(b[i]))
        return c

    def expand_a(self, rho, k, l):
        rho = np.ascontiguousarray(rho, dtype=np.uint8)
        out = np.empty((k * l, N), dtype=np.int32)
        self.lib.orc_expand_a(_i32(out), _u8(rho), k, l)
        return out

    def expand_a_blocks(self, rho, i, j):
        rho = np.ascontiguousarray(rho, dtype=np.uint8)
        out = np.empty(N, dtype=np.int32)
        return self.lib.orc_expand_a_poly(_i32(out), _u8(rho), i, j)

    def matvec(self, a_hat, v, k, l, threads=1):
        v = np.ascontiguousarray(v, dtype=np.int32).reshape(-1, l, N)
        a_hat = np.ascontiguousarray(a_hat, dtype=np.int32)
        w = np.empty((v.shape[0], k, N), dtype=np.int32)
        self.lib.orc_matvec_batch_mt(_i32(w), _i32(a_hat), _i32(v), k, l, ctypes.c_size_t(v.shape[0]), threads)
        return w

    def matvec_expand(self, rho, v, k, l, ntt_in=False, invntt_out=False):
        rho = np.ascontiguousarray(rho, dtype=np.uint8).reshape(-1, 32)
        v = np.ascontiguousarray(v, dtype=np.int32).reshape(-1, l, N)
        w = np.empty((v.shape[0], k, N), dtype=np.int32)
        self.lib.orc_matvec_expand_batch(_i32(w), _u8(rho), ctypes.c_size_t(rho.shape[0]), _i32(v), k, l,
                                         ctypes.c_size_t(v.shape[0]), int(ntt_in), int(invntt_out))
        return w

    def signcore(self, a_hat, y, k, l, threads=1):
        """cfg2 core: w = INTT(A_hat * NTT(y)).  Returns (w, y_hat)."""
        y = np.ascontiguousarray(y, dtype=np.int32).copy().reshape(-1, l, N)
        a_hat = np.ascontiguousarray(a_hat, dtype=np.int32)
        w = np.empty((y.shape[0], k, N), dtype=np.int32)
        self.lib.orc_signcore_batch_mt(_i32(w), _i32(y), _i32(a_hat), k, l, ctypes.c_size_t(y.shape[0]), threads)
        return w, y

    def time_signcore(self, a_hat, y, k, l, threads=1, steps=1, warmup=0):
        return _time_signcore(self.lib.orc_signcore_batch_mt, a_hat, y, k, l, threads, steps, warmup)

    def sign_batch(self, level, K, i, msgs, threads=1):
        return _sign_batch(self.lib, level, K, i, msgs, threads)

    def shake_bytes(self, data: bytes, outlen: int, bits=256) -> bytes:
        buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
        out = np.empty(outlen, dtype=np.uint8)
        fn = self.lib.orc_shake256 if bits == 256 else self.lib.orc_shake128
        fn(_u8(out), ctypes.c_size_t(outlen), _u8(buf), ctypes.c_size_t(len(data)))
        return out.tobytes()

    # ---- scheme level ----
    def keygen_chain(self, level, rho, s1p, s2p):
        P = PARAMS[level]
        t1p = np.empty(P["k"] * 320, dtype=np.uint8)
        t0p = np.empty(P["k"] * 416, dtype=np.uint8)
        rc = self.lib.orc_keygen_chain(level, _u8(np.ascontiguousarray(rho)), _u8(np.ascontiguousarray(s1p)),
                                       _u8(np.ascontiguousarray(s2p)), _u8(t1p), _u8(t0p))
        assert rc == 0
        return t1p, t0p

    def keygen(self, level, xi):
        P = PARAMS[level]
        sb = 96 if P["eta"] == 2 else 128
        rho, key, tr = (np.empty(32, dtype=np.uint8) for _ in range(3))
        s1p = np.empty(P["l"] * sb, dtype=np.uint8)
        s2p = np.empty(P["k"] * sb, dtype=np.uint8)
        t1p = np.empty(P["k"] * 320, dtype=np.uint8)
        t0p = np.empty(P["k"] * 416, dtype=np.uint8)
        rc = self.lib.orc_keygen(level, _u8(np.ascontiguousarray(xi)), _u8(rho), _u8(key), _u8(tr), _u8(s1p),
                                 _u8(s2p), _u8(t1p), _u8(t0p))
        assert rc == 0
        return dict(rho=rho, k=key, tr=tr, s1=s1p, s2=s2p, t1=t1p, t0=t0p)

    def sign(self, level, rho, key, tr, s1p, s2p, t0p, msg: bytes):
        P = PARAMS[level]
        zb = N * (P["gamma1_bits"] + 1) // 8
        zp = np.empty(P["l"] * zb, dtype=np.uint8)
        hp = np.empty(P["omega"] + P["k"], dtype=np.uint8)
        c = np.empty(32, dtype=np.uint8)
        m = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, np.uint8)
        c8 = np.ascontiguousarray
        attempts = self.lib.orc_sign(level, _u8(c8(rho)), _u8(c8(key)), _u8(c8(tr)), _u8(c8(s1p)), _u8(c8(s2p)),
                                     _u8(c8(t0p)), _u8(m), ctypes.c_size_t(len(msg)), _u8(zp), _u8(hp), _u8(c))
        assert attempts > 0
        return zp, hp, c, attempts

    def verify(self, level, rho, t1p, msg: bytes, zp, hp, c):
        m = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, np.uint8)
        c8 = np.ascontiguousarray
        return self.lib.orc_verify(level, _u8(c8(rho)), _u8(c8(t1p)), _u8(m), ctypes.c_size_t(len(msg)),
                                   _u8(c8(zp)), _u8(c8(hp)), _u8(c8(c)))


def _sign_batch(lib, level, K, i, msgs, threads, attempts=True):
    """Threaded CPU batch sign with KAT key i.  Returns (z, h, c, attempts, seconds)."""
    import time
    P = PARAMS[level]
    n = len(msgs)
    zb = P["l"] * N * (P["gamma1_bits"] + 1) // 8
    hb = P["omega"] + P["k"]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    z, h, c = np.empty((n, zb), np.uint8), np.empty((n, hb), np.uint8), np.empty((n, 32), np.uint8)
    att = np.zeros(n, dtype=np.uint32)
    c8 = np.ascontiguousarray
    g = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    args = [level] + [g(c8(K[f][i])) for f in ("rho", "k", "tr", "s1", "s2", "t0")] + [g(blob), g(off), ctypes.c_size_t(n), g(z), g(h), g(c), g(att), threads]
    t0 = time.perf_counter()
    lib.orc_sign_batch_mt(*args)
    return z, h, c, att, time.perf_counter() - t0


def _time_signcore(cfn, a_hat, y, k, l, threads, steps, warmup):
    """Wall-clock seconds per call of a C sign-core batch driver (buffers prepared outside the
    timed region; the transforms are data-independent so y is transformed in place repeatedly)."""
    import time
    y = np.ascontiguousarray(y, dtype=np.int32).copy().reshape(-1, l, N)
    a_hat = np.ascontiguousarray(a_hat, dtype=np.int32)
    w = np.empty((y.shape[0], k, N), dtype=np.int32)
    args = (_i32(w), _i32(y), _i32(a_hat), k, l, ctypes.c_size_t(y.shape[0]), threads)
    for _ in range(warmup):
        cfn(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        cfn(*args)
    return (time.perf_counter() - t0) / steps


class Ref:
    """The reference's own compiled C++ (oracle/_ref/libdilref.so)."""

    def time_signcore(self, a_hat, y, k, l, threads=1, steps=1, warmup=0):
        return _time_signcore(self.lib.ref_signcore_batch, a_hat, y, k, l, threads, steps, warmup)

    def sign_batch(self, level, K, i, msgs, threads=1):
        """Full sign on host cores: reference ntt/invntt/pointwise_barrett for all polynomial
        arithmetic (ref_bridge.cpp) + the oracle's scheme glue."""
        return _sign_batch(self.lib, level, K, i, msgs, threads)

    def __init__(self, lib):
        self.lib = lib
        lib.ref_zetas.restype = I32P

    def zetas(self):
        return np.ctypeslib.as_array(self.lib.ref_zetas(), shape=(N,)).copy()

    def run(self, fn, a, threads=1, canon=True):
        a = np.ascontiguousarray(a, dtype=np.int32).copy().reshape(-1, N)
        getattr(self.lib, fn)(_i32(a), ctypes.c_size_t(a.shape[0]), threads)
        return (a.astype(np.int64) % Q).astype(np.int32) if canon else a

    def pointwise(self, a, b, threads=1):
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, N)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, N)
        c = np.empty_like(a)
        self.lib.ref_pointwise_batch(_i32(c), _i32(a), _i32(b), ctypes.c_size_t(a.shape[0]), threads)
        return (c.astype(np.int64) % Q).astype(np.int32)

    def signcore(self, a_hat, y, k, l, threads=1):
        y = np.ascontiguousarray(y, dtype=np.int32).copy().reshape(-1, l, N)
        a_hat = np.ascontiguousarray(a_hat, dtype=np.int32)
        w = np.empty((y.shape[0], k, N), dtype=np.int32)
        self.lib.ref_signcore_batch(_i32(w), _i32(y), _i32(a_hat), k, l, ctypes.c_size_t(y.shape[0]), threads)
        return (w.astype(np.int64) % Q).astype(np.int32)


def build():
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", ORC_DIR, "-s"], check=True, capture_output=True)


def load():
    path = os.path.join(ORC_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    return Oracle(ctypes.CDLL(path))


def load_ref():
    path = os.path.join(ORC_DIR, "_ref", "libdilref.so")
    if not os.path.exists(path):
        return None
    return Ref(ctypes.CDLL(path))


def kat(level):
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", f"kat_L{level}.npz")))
    m = np.load(os.path.join(ROOT, "tests", "golden", "kat_msgs.npz"))
    blob, mlen = m["blob"].tobytes(), m["mlen"]
    off = np.concatenate([[0], np.cumsum(mlen)])
    d["msgs"] = [blob[off[i]:off[i + 1]] for i in range(len(mlen))]
    return d
