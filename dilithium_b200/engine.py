"""Host-side mirror of the reference's polynomial-arithmetic interface on top of the C ABI.

Names follow the reference (`dilithium-256/reference_code/ref_ntt.h:30-36`, `ref_ntt2x2.h:31-33`)
plus the north_star aliases:

    Engine.ntt / invntt / pointwise_barrett / ntt2x2_ref / invntt2x2_ref
    Engine.invntt_tomont / poly_pointwise / polyvec_matrix_pointwise
    Engine.pointwise_acc / add / sub            (operation_module modes 2/3/4, butterfly.v:144-164)
    Engine.expand_a / matvec_expand / signcore  (ExpandA-fused mat-vec, combined_top.v:1850-1933)

Every method accepts either
  * numpy int32 arrays (host memory): routed through the `dil_*_host` entry points, which copy
    to the device, run the CUDA kernels and copy back; a new array is returned; or
  * torch CUDA int32 tensors: routed through the `dil_*_dev` entry points on torch's current
    stream, no copies, asynchronous; the result tensor is returned (in-place when `out` is
    the input).
PyTorch is only the owner of device memory/streams here; all arithmetic happens in
libdilithium_b200.so.  There is no CPU implementation behind this class.
"""
import ctypes

import numpy as np

from . import _lib

Q = 8380417
N = 256
LEVEL_DIMS = {2: (4, 4), 3: (6, 5), 5: (8, 7)}  # level -> (k, l), combined_top.v:520-551
RHO_SHARED, RHO_PER_ITEM, NTT_INPUT, INTT_OUTPUT = 0, 1, 2, 4


class DilithiumError(RuntimeError):
    pass


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.dil_engine_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise DilithiumError(f"dil_engine_create(device={device}) failed: "
                                 f"{self._lib.dil_status_string(rc).decode()} (no CPU fallback exists)")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dil_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def _check(self, rc, what):
        if rc != 0:
            raise DilithiumError(f"{what}: {self._lib.dil_status_string(rc).decode()}: "
                                 f"{self._lib.dil_last_error(self._h).decode()}")

    @property
    def sm_count(self):
        return self._lib.dil_engine_sm_count(self._h)

    @property
    def launch_count(self):
        return int(self._lib.dil_engine_launch_count(self._h))

    @staticmethod
    def _np(a, dtype=np.int32):
        a = np.ascontiguousarray(a, dtype=dtype)
        return a

    @staticmethod
    def _stream():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _tcheck(t, dtype=None):
        import torch
        dtype = dtype or torch.int32
        if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
            raise DilithiumError(f"expected a contiguous CUDA {dtype} tensor")
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _npolys(a):
        n = a.size if isinstance(a, np.ndarray) else a.numel()
        if n % N:
            raise DilithiumError("array size is not a multiple of 256 coefficients")
        return n // N

    # ---- transforms ----
    def _unary(self, name, a, out=None):
        if _is_torch(a):
            out = a.new_empty(a.shape) if out is None else out
            rc = getattr(self._lib, f"dil_{name}_dev")(self._h, self._tcheck(out), self._tcheck(a), self._npolys(a), self._stream())
            self._check(rc, f"dil_{name}_dev")
            return out
        buf = self._np(a).copy()
        rc = getattr(self._lib, f"dil_{name}_host")(self._h, buf.ctypes.data_as(ctypes.c_void_p), self._npolys(buf))
        self._check(rc, f"dil_{name}_host")
        return buf

    def ntt(self, a, out=None):
        """ref_ntt.h:30  void ntt(data_t a[256]) - batched, canonical output."""
        return self._unary("ntt", a, out)

    def invntt(self, a, out=None):
        """ref_ntt.h:36  void invntt(data_t a[256]) - batched, includes 256^-1, canonical output."""
        return self._unary("invntt", a, out)

    ntt2x2_ref = ntt          # ref_ntt2x2.h:31 (same map; the engine's schedule is radix-2x2-style)
    invntt2x2_ref = invntt    # ref_ntt2x2.h:33
    invntt_tomont = invntt    # north_star alias (plain domain)

    # ---- coefficient-wise ----
    def _binary(self, name, a, b, c=None):
        if _is_torch(a):
            c = a.new_empty(a.shape) if c is None else c
            rc = getattr(self._lib, f"dil_{name}_dev")(self._h, self._tcheck(c), self._tcheck(a), self._tcheck(b), self._npolys(a), self._stream())
            self._check(rc, f"dil_{name}_dev")
            return c
        a, b = self._np(a), self._np(b)
        if a.shape != b.shape:
            raise DilithiumError("shape mismatch")
        c = np.empty_like(a) if c is None else self._np(c).copy()
        rc = getattr(self._lib, f"dil_{name}_host")(self._h, c.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                                                    b.ctypes.data_as(ctypes.c_void_p), self._npolys(a))
        self._check(rc, f"dil_{name}_host")
        return c

    def pointwise_barrett(self, a, b, c=None):
        """ref_ntt.h:32-34  c = a o b mod Q."""
        return self._binary("pointwise", a, b, c)

    poly_pointwise = pointwise_barrett

    def pointwise_acc(self, c, a, b):
        """butterfly.v:144-150 MULT mode: c + a o b."""
        return self._binary("pointwise_acc", a, b, c)

    def add(self, a, b, c=None):
        return self._binary("add", a, b, c)

    def sub(self, a, b, c=None):
        return self._binary("sub", a, b, c)

    # ---- mat-vec family ----
    def polyvec_matrix_pointwise(self, a_hat, v, k, l, w=None):
        """MULT_MODE loop nest (combined_top.v:921-958): w[b,i] = sum_j a_hat[i*l+j] o v[b,j]."""
        if _is_torch(v):
            batch = self._npolys(v) // l
            w = v.new_empty((batch, k, N)) if w is None else w
            rc = self._lib.dil_matvec_dev(self._h, self._tcheck(w), self._tcheck(a_hat), self._tcheck(v), k, l, batch, self._stream())
            self._check(rc, "dil_matvec_dev")
            return w
        a_hat, v = self._np(a_hat), self._np(v)
        batch = self._npolys(v) // l
        w = np.empty((batch, k, N), dtype=np.int32)
        rc = self._lib.dil_matvec_host(self._h, w.ctypes.data_as(ctypes.c_void_p), a_hat.ctypes.data_as(ctypes.c_void_p),
                                       v.ctypes.data_as(ctypes.c_void_p), k, l, batch)
        self._check(rc, "dil_matvec_host")
        return w

    matvec = polyvec_matrix_pointwise

    def expand_a(self, rho, k, l):
        """gen_a_ext / sampler_a_ext / rejection_a: rho[n,32] -> a_hat[n, k*l, 256]."""
        if _is_torch(rho):
            import torch
            n = rho.numel() // 32
            a = torch.empty((n, k * l, N), dtype=torch.int32, device=rho.device)
            rc = self._lib.dil_expand_a_dev(self._h, self._tcheck(a), self._tcheck(rho, torch.uint8), n, k, l, self._stream())
            self._check(rc, "dil_expand_a_dev")
            return a
        rho = self._np(rho, np.uint8).reshape(-1, 32)
        a = np.empty((rho.shape[0], k * l, N), dtype=np.int32)
        rc = self._lib.dil_expand_a_host(self._h, a.ctypes.data_as(ctypes.c_void_p), rho.ctypes.data_as(ctypes.c_void_p), rho.shape[0], k, l)
        self._check(rc, "dil_expand_a_host")
        return a

    def matvec_expand(self, rho, v, k, l, per_item=False, ntt_input=False, intt_output=False, w=None):
        """w[b] = [INTT](ExpandA(rho) * [NTT] v[b]) with A generated on chip (never in HBM)."""
        flags = (RHO_PER_ITEM if per_item else 0) | (NTT_INPUT if ntt_input else 0) | (INTT_OUTPUT if intt_output else 0)
        if _is_torch(v):
            import torch
            batch = self._npolys(v) // l
            w = v.new_empty((batch, k, N)) if w is None else w
            rc = self._lib.dil_matvec_expand_dev(self._h, self._tcheck(w), self._tcheck(rho, torch.uint8), self._tcheck(v), k, l, batch, flags, self._stream())
            self._check(rc, "dil_matvec_expand_dev")
            return w
        rho, v = self._np(rho, np.uint8), self._np(v)
        batch = self._npolys(v) // l
        if rho.size != (32 * batch if per_item else 32):
            raise DilithiumError("rho has the wrong size for this mode")
        w = np.empty((batch, k, N), dtype=np.int32)
        rc = self._lib.dil_matvec_expand_host(self._h, w.ctypes.data_as(ctypes.c_void_p), rho.ctypes.data_as(ctypes.c_void_p),
                                              v.ctypes.data_as(ctypes.c_void_p), k, l, batch, flags)
        self._check(rc, "dil_matvec_expand_host")
        return w

    def signcore(self, a_hat, y, k, l, w=None):
        """cfg2 core in one kernel: w[b] = INTT(a_hat * NTT(y[b]))  (NTT_Y -> MULT_A_Y -> NTTI_W)."""
        if _is_torch(y):
            batch = self._npolys(y) // l
            w = y.new_empty((batch, k, N)) if w is None else w
            rc = self._lib.dil_signcore_dev(self._h, self._tcheck(w), self._tcheck(a_hat), self._tcheck(y), k, l, batch, self._stream())
            self._check(rc, "dil_signcore_dev")
            return w
        a_hat, y = self._np(a_hat), self._np(y)
        batch = self._npolys(y) // l
        w = np.empty((batch, k, N), dtype=np.int32)
        rc = self._lib.dil_signcore_host(self._h, w.ctypes.data_as(ctypes.c_void_p), a_hat.ctypes.data_as(ctypes.c_void_p),
                                         y.ctypes.data_as(ctypes.c_void_p), k, l, batch)
        self._check(rc, "dil_signcore_host")
        return w

    def sign_multi(self, level, rho, key, tr, s1_packed, s2_packed, t0_packed, msgs):
        """Sign n messages, each under its own secret key (n-record arrays, bit-packed as the KAT files): the I/O of
        rtl_tb/tb_sign_top.v:171-284.  Returns (z, h, ctilde, attempts)."""
        n = len(msgs)
        k, l = LEVEL_DIMS[level]
        zb = l * (576 if level == 2 else 640)
        hb = {2: 84, 3: 61, 5: 83}[level]
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        keys = [np.ascontiguousarray(x, dtype=np.uint8) for x in (rho, key, tr, s1_packed, s2_packed, t0_packed)]
        z = np.empty((n, zb), np.uint8); h = np.empty((n, hb), np.uint8); c = np.empty((n, 32), np.uint8)
        att = np.zeros(n, np.uint32)
        P = ctypes.c_void_p
        rc = self._lib.dil_sign_multi_host(self._h, int(level), *[a.ctypes.data_as(P) for a in keys], blob.ctypes.data_as(P),
                                           off.ctypes.data_as(P), n, z.ctypes.data_as(P), h.ctypes.data_as(P), c.ctypes.data_as(P),
                                           att.ctypes.data_as(P))
        self._check(rc, "dil_sign_multi_host")
        return z, h, c, att

    def verify_multi(self, level, rho, t1_packed, msgs, z, h, ctilde):
        """Verify n signatures, each under its own public key (rho[i], t1[i]).  Returns ok[n] (1 = accept)."""
        n = len(msgs)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        arrs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (rho, t1_packed)]
        sig = [np.ascontiguousarray(x, dtype=np.uint8) for x in (z, h, ctilde)]
        ok = np.zeros(n, dtype=np.uint8)
        P = ctypes.c_void_p
        rc = self._lib.dil_verify_multi_host(self._h, int(level), arrs[0].ctypes.data_as(P), arrs[1].ctypes.data_as(P),
                                             blob.ctypes.data_as(P), off.ctypes.data_as(P), n, sig[0].ctypes.data_as(P),
                                             sig[1].ctypes.data_as(P), sig[2].ctypes.data_as(P), ok.ctypes.data_as(P))
        self._check(rc, "dil_verify_multi_host")
        return ok

    def verify_multi_dev(self, level, d_rho, d_t1, d_msgs, d_offsets, n, d_z, d_h, d_c, d_ok):
        """Device-resident per-key verification (torch CUDA uint8 tensors) on torch's current stream."""
        P = ctypes.c_void_p
        rc = self._lib.dil_verify_multi_dev(self._h, int(level), P(d_rho.data_ptr()), P(d_t1.data_ptr()), P(d_msgs.data_ptr()),
                                            P(d_offsets.data_ptr()), n, P(d_z.data_ptr()), P(d_h.data_ptr()), P(d_c.data_ptr()),
                                            P(d_ok.data_ptr()), self._stream())
        self._check(rc, "dil_verify_multi_dev")

    def keccak_rate(self, ctas_per_sm=1, perms=2000, repeats=3):
        """Measured pure Keccak-f[1600] rate of this GPU in permutations/s (dil_diag_keccak_dev, CUDA-event timed):
        the ALU-pipe speed of light of the engine's hash kernels."""
        import torch
        threads = self.sm_count * ctas_per_sm * 128
        out = torch.empty(threads, dtype=torch.int64, device=f"cuda:{self.device}")
        best = 0.0
        for _ in range(repeats + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self._check(self._lib.dil_diag_keccak_dev(self._h, ctypes.c_void_p(out.data_ptr()), ctas_per_sm, perms, self._stream()),
                        "dil_diag_keccak_dev")
            b.record()
            b.synchronize()
            best = max(best, threads * perms / (a.elapsed_time(b) * 1e-3))
        return best

    def keygen(self, level, seeds):
        """Batched key generation from 32-byte seeds xi (host path).  Returns a dict of uint8 arrays with the
        KAT field names: rho, k, tr, s1, s2, t1, t0 (bit-packed as the reference's KAT files)."""
        seeds = self._np(seeds, np.uint8).reshape(-1, 32)
        n = seeds.shape[0]
        k, l = LEVEL_DIMS[level]
        sb = 128 if level == 3 else 96
        out = dict(rho=np.empty((n, 32), np.uint8), k=np.empty((n, 32), np.uint8), tr=np.empty((n, 32), np.uint8),
                   s1=np.empty((n, l * sb), np.uint8), s2=np.empty((n, k * sb), np.uint8),
                   t1=np.empty((n, k * 320), np.uint8), t0=np.empty((n, k * 416), np.uint8))
        P = ctypes.c_void_p
        rc = self._lib.dil_keygen_batch_host(self._h, int(level), seeds.ctypes.data_as(P), n,
                                             *[out[f].ctypes.data_as(P) for f in ("rho", "k", "tr", "s1", "s2", "t1", "t0")])
        self._check(rc, "dil_keygen_batch_host")
        return out


def _keygen_dev(self, level, d_seeds):
    """Device-resident key generation (dil_keygen_batch_dev): d_seeds = CUDA uint8 tensor [n, 32]; returns a dict of CUDA
    uint8 tensors with the KAT field names, produced on torch's current stream."""
    import torch
    n = d_seeds.numel() // 32
    k, l = LEVEL_DIMS[level]
    sb = 128 if level == 3 else 96
    dev = d_seeds.device
    out = {f: torch.empty((n, w), dtype=torch.uint8, device=dev) for f, w in
           (("rho", 32), ("k", 32), ("tr", 32), ("s1", l * sb), ("s2", k * sb), ("t1", k * 320), ("t0", k * 416))}
    P = ctypes.c_void_p
    rc = self._lib.dil_keygen_batch_dev(self._h, int(level), P(d_seeds.data_ptr()), n,
                                        *[P(out[f].data_ptr()) for f in ("rho", "k", "tr", "s1", "s2", "t1", "t0")], self._stream())
    self._check(rc, "dil_keygen_batch_dev")
    return out


Engine.keygen_dev = _keygen_dev


class SignKey:
    """Expanded signing key on the device (ExpandA + NTT of s1, s2, t0 done once; LOAD_RHO / NTT_S1 /
    NTT_S2 / NTT_T0 of combined_top.v:1560-1767).  Inputs are bit-packed exactly as the reference's
    KAT files / rtl_tb/tb_sign_top.v:171-284 feed them."""

    def __init__(self, engine, level, rho, key, tr, s1_packed, s2_packed, t0_packed):
        self.engine, self.level = engine, int(level)
        lib = engine._lib
        zb, hb = ctypes.c_size_t(), ctypes.c_size_t()
        engine._check(lib.dil_sign_sizes(self.level, ctypes.byref(zb), ctypes.byref(hb)), "dil_sign_sizes")
        self.z_bytes, self.h_bytes = zb.value, hb.value
        bufs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (rho, key, tr, s1_packed, s2_packed, t0_packed)]
        h = ctypes.c_void_p()
        rc = lib.dil_sign_key_create(engine._h, ctypes.byref(h), self.level, *[b.ctypes.data_as(ctypes.c_void_p) for b in bufs])
        engine._check(rc, "dil_sign_key_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.engine, "_h", None):
            self.engine._lib.dil_sign_key_destroy(self.engine._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_rounds(self):
        return int(self.engine._lib.dil_sign_last_rounds(self._h))

    @property
    def last_slots(self):
        """Signing attempts (slots) the last batch executed, speculative ones included."""
        return int(self.engine._lib.dil_sign_last_slots(self._h))

    class _Tuning(ctypes.Structure):   # dil_sign_tuning
        _fields_ = [("spec_target", ctypes.c_uint32), ("spec_max", ctypes.c_uint32), ("dev_chunk", ctypes.c_size_t),
                    ("host_chunk", ctypes.c_size_t), ("host_copy_path", ctypes.c_int), ("fused_mask", ctypes.c_int),
                    ("mask_producers", ctypes.c_int)]

    def set_tuning(self, spec_target=0, spec_max=0, dev_chunk=0, host_chunk=0, host_copy_path=False, fused_mask=False,
                   mask_producers=0):
        """Scheduler knobs of this key (dil_sign_key_set_tuning); zeros restore the defaults.  Results never change."""
        t = self._Tuning(int(spec_target), int(spec_max), int(dev_chunk), int(host_chunk), int(bool(host_copy_path)),
                         int(bool(fused_mask)), int(mask_producers))
        self.engine._check(self.engine._lib.dil_sign_key_set_tuning(self._h, ctypes.byref(t)), "dil_sign_key_set_tuning")

    PROFILE_CLASSES = ("init", "expand_mask", "signcore", "pack_w1", "challenge", "tail", "resolve")

    def set_profile(self, on=True):
        self.engine._check(self.engine._lib.dil_sign_set_profile(self._h, int(on)), "dil_sign_set_profile")

    def get_profile(self):
        """{class: (device ms, slots processed)} of the last batch signed with profiling on."""
        ms = (ctypes.c_double * 8)()
        units = (ctypes.c_uint64 * 8)()
        self.engine._check(self.engine._lib.dil_sign_get_profile(self._h, ms, units), "dil_sign_get_profile")
        return {name: (ms[i], int(units[i])) for i, name in enumerate(self.PROFILE_CLASSES)}

    def sign(self, msgs, pinned=False):
        """Sign a list of byte strings (host path, dil_sign_batch_host).
        Returns (z[n, z_bytes], h[n, h_bytes], ctilde[n, 32], attempts[n]).
        pinned=True returns the outputs in pinned host memory, which lets the library stream finished
        signatures out round by round instead of copying whole chunks afterwards."""
        n = len(msgs)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        if pinned:
            import torch
            self._pinned = [torch.empty(shape, dtype=dt).pin_memory() for shape, dt in
                            (((max(n, 1), self.z_bytes), torch.uint8), ((max(n, 1), self.h_bytes), torch.uint8),
                             ((max(n, 1), 32), torch.uint8), ((max(n, 1),), torch.int32))]
            z, h, c = (t.numpy()[:n] for t in self._pinned[:3])
            att = self._pinned[3].numpy()[:n].view(np.uint32)
            att[:] = 0
        else:
            z = np.empty((n, self.z_bytes), dtype=np.uint8)
            h = np.empty((n, self.h_bytes), dtype=np.uint8)
            c = np.empty((n, 32), dtype=np.uint8)
            att = np.zeros(n, dtype=np.uint32)
        P = ctypes.c_void_p
        rc = self.engine._lib.dil_sign_batch_host(self.engine._h, self._h, blob.ctypes.data_as(P), off.ctypes.data_as(P), n,
                                                  z.ctypes.data_as(P), h.ctypes.data_as(P), c.ctypes.data_as(P), att.ctypes.data_as(P))
        self.engine._check(rc, "dil_sign_batch_host")
        return z, h, c, att

    def sign_dev(self, d_msgs, d_offsets, n, d_z, d_h, d_c, d_att):
        """Device-resident variant (torch CUDA uint8 / int64-as-uint64 / int32 tensors) on torch's current stream."""
        rc = self.engine._lib.dil_sign_batch_dev(self.engine._h, self._h, ctypes.c_void_p(d_msgs.data_ptr()),
                                                 ctypes.c_void_p(d_offsets.data_ptr()), n, ctypes.c_void_p(d_z.data_ptr()),
                                                 ctypes.c_void_p(d_h.data_ptr()), ctypes.c_void_p(d_c.data_ptr()),
                                                 ctypes.c_void_p(d_att.data_ptr()), self.engine._stream())
        self.engine._check(rc, "dil_sign_batch_dev")


def _sign_dev_begin(self, d_msgs, d_offsets, n, d_z, d_h, d_c, d_att):
    """Asynchronous device-resident signing (dil_sign_batch_dev_begin): enqueues the batch on torch's current stream and returns;
    finish() completes it.  One batch per key handle at a time."""
    rc = self.engine._lib.dil_sign_batch_dev_begin(self.engine._h, self._h, ctypes.c_void_p(d_msgs.data_ptr()),
                                                   ctypes.c_void_p(d_offsets.data_ptr()), n, ctypes.c_void_p(d_z.data_ptr()),
                                                   ctypes.c_void_p(d_h.data_ptr()), ctypes.c_void_p(d_c.data_ptr()),
                                                   ctypes.c_void_p(d_att.data_ptr()), self.engine._stream())
    self.engine._check(rc, "dil_sign_batch_dev_begin")


def _sign_host_begin(self, msgs_t, offsets_t, n, z_t, h_t, c_t, att_t):
    """Asynchronous host-pointer signing (dil_sign_batch_host_begin): pinned torch tensors (or anything with data_ptr()) for the
    messages, offsets and outputs; they must stay alive until finish()."""
    P = ctypes.c_void_p
    rc = self.engine._lib.dil_sign_batch_host_begin(self.engine._h, self._h, P(msgs_t.data_ptr()), P(offsets_t.data_ptr()), n,
                                                    P(z_t.data_ptr()), P(h_t.data_ptr()), P(c_t.data_ptr()), P(att_t.data_ptr()))
    self.engine._check(rc, "dil_sign_batch_host_begin")


def _sign_finish(self):
    """Complete the batch begun with sign_dev_begin / sign_host_begin (dil_sign_batch_finish)."""
    self.engine._check(self.engine._lib.dil_sign_batch_finish(self.engine._h, self._h), "dil_sign_batch_finish")


SignKey.sign_dev_begin = _sign_dev_begin
SignKey.sign_host_begin = _sign_host_begin
SignKey.finish = _sign_finish


class VerifyKey:
    """Expanded public key on the device (rho, t1 as in the KAT files / tb_verify_top.v:144-240)."""

    def __init__(self, engine, level, rho, t1_packed):
        self.engine, self.level = engine, int(level)
        lib = engine._lib
        zb, hb = ctypes.c_size_t(), ctypes.c_size_t()
        engine._check(lib.dil_sign_sizes(self.level, ctypes.byref(zb), ctypes.byref(hb)), "dil_sign_sizes")
        self.z_bytes, self.h_bytes = zb.value, hb.value
        bufs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (rho, t1_packed)]
        h = ctypes.c_void_p()
        engine._check(lib.dil_verify_key_create(engine._h, ctypes.byref(h), self.level, *[b.ctypes.data_as(ctypes.c_void_p) for b in bufs]),
                      "dil_verify_key_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.engine, "_h", None):
            self.engine._lib.dil_verify_key_destroy(self.engine._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def verify(self, msgs, z, h, ctilde):
        """Host path: returns ok[n] (uint8, 1 = accept)."""
        n = len(msgs)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        z, h, c = (np.ascontiguousarray(x, dtype=np.uint8) for x in (z, h, ctilde))
        assert z.size == n * self.z_bytes and h.size == n * self.h_bytes and c.size == n * 32
        ok = np.zeros(n, dtype=np.uint8)
        P = ctypes.c_void_p
        rc = self.engine._lib.dil_verify_batch_host(self.engine._h, self._h, blob.ctypes.data_as(P), off.ctypes.data_as(P), n,
                                                    z.ctypes.data_as(P), h.ctypes.data_as(P), c.ctypes.data_as(P), ok.ctypes.data_as(P))
        self.engine._check(rc, "dil_verify_batch_host")
        return ok

    def verify_dev(self, d_msgs, d_offsets, n, d_z, d_h, d_c, d_ok):
        rc = self.engine._lib.dil_verify_batch_dev(self.engine._h, self._h, ctypes.c_void_p(d_msgs.data_ptr()),
                                                   ctypes.c_void_p(d_offsets.data_ptr()), n, ctypes.c_void_p(d_z.data_ptr()),
                                                   ctypes.c_void_p(d_h.data_ptr()), ctypes.c_void_p(d_c.data_ptr()),
                                                   ctypes.c_void_p(d_ok.data_ptr()), self.engine._stream())
        self.engine._check(rc, "dil_verify_batch_dev")


class Pool:
    """Several GPUs in one process (dil_pool_*): one engine per device, batches split into contiguous shards."""

    def __init__(self, devices=None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._lib.dil_pool_create(ctypes.byref(h), arr, len(devices))
        else:
            rc = self._lib.dil_pool_create(ctypes.byref(h), None, 0)
        if rc != 0:
            raise DilithiumError(f"dil_pool_create failed: {self._lib.dil_status_string(rc).decode()}")
        self._h = h
        self._key = None

    @property
    def size(self):
        return self._lib.dil_pool_size(self._h)

    def load_key(self, level, rho, key, tr, s1_packed, s2_packed, t0_packed):
        self.drop_key()
        bufs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (rho, key, tr, s1_packed, s2_packed, t0_packed)]
        k = ctypes.c_void_p()
        rc = self._lib.dil_pool_sign_key_create(self._h, ctypes.byref(k), int(level), *[b.ctypes.data_as(ctypes.c_void_p) for b in bufs])
        if rc != 0:
            raise DilithiumError(f"dil_pool_sign_key_create failed: {self._lib.dil_status_string(rc).decode()}")
        self._key, self.level = k, int(level)

    def drop_key(self):
        if self._key:
            self._lib.dil_pool_sign_key_destroy(self._h, self._key)
            self._key = None

    def sign_into(self, msgs_ptr, offsets_ptr, n, z_ptr, h_ptr, c_ptr, att_ptr):
        """Raw-pointer form (host memory; pinned portable output buffers are streamed to)."""
        P = ctypes.c_void_p
        rc = self._lib.dil_pool_sign_batch_host(self._h, self._key, P(msgs_ptr), P(offsets_ptr), n, P(z_ptr), P(h_ptr), P(c_ptr), P(att_ptr))
        if rc != 0:
            raise DilithiumError(f"dil_pool_sign_batch_host failed: {self._lib.dil_status_string(rc).decode()}")

    def sign_into_begin(self, msgs_ptr, offsets_ptr, n, z_ptr, h_ptr, c_ptr, att_ptr):
        """Asynchronous form (dil_pool_sign_batch_host_begin): pinned portable output buffers; finish() completes the batch."""
        P = ctypes.c_void_p
        rc = self._lib.dil_pool_sign_batch_host_begin(self._h, self._key, P(msgs_ptr), P(offsets_ptr), n, P(z_ptr), P(h_ptr), P(c_ptr), P(att_ptr))
        if rc != 0:
            raise DilithiumError(f"dil_pool_sign_batch_host_begin failed: {self._lib.dil_status_string(rc).decode()}")

    def finish(self):
        rc = self._lib.dil_pool_sign_batch_finish(self._h, self._key)
        if rc != 0:
            raise DilithiumError(f"dil_pool_sign_batch_finish failed: {self._lib.dil_status_string(rc).decode()}")

    def sign(self, msgs):
        n = len(msgs)
        k, l = LEVEL_DIMS[self.level]
        zb = l * (576 if self.level == 2 else 640)
        hb = {2: 84, 3: 61, 5: 83}[self.level]
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs])
        blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        z = np.empty((n, zb), np.uint8); h = np.empty((n, hb), np.uint8); c = np.empty((n, 32), np.uint8)
        att = np.zeros(n, np.uint32)
        self.sign_into(blob.ctypes.data, off.ctypes.data, n, z.ctypes.data, h.ctypes.data, c.ctypes.data, att.ctypes.data)
        return z, h, c, att

    def close(self):
        if getattr(self, "_h", None):
            self.drop_key()
            self._lib.dil_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
