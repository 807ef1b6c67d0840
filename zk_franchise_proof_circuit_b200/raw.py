"""Raw kernel entry points (BASELINE.json config 5: BN254 G1/G2 MSM and Fr NTT on caller data)."""
import ctypes
import numpy as np
from . import _native


def field_op(field: str, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Element-wise op on Montgomery residues; a, b: uint32[n, 8]."""
    ops = {"mul": 0, "add": 1, "sub": 2, "inv": 3, "sqr": 4, "to_mont": 5, "from_mont": 6, "neg": 7}
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.zeros_like(a)
    _native.check(_native.lib().zkb_raw_field_op(0 if field == "fq" else 1, ops[op], a.ctypes.data, b.ctypes.data,
                                                 out.ctypes.data, a.shape[0]))
    return out


def bench_modmul(field="fq", iters=4096, blocks_per_sm=8):
    r, ms = ctypes.c_double(), ctypes.c_double()
    _native.check(_native.lib().zkb_bench_modmul(0 if field == "fq" else 1, iters, blocks_per_sm,
                                                 ctypes.byref(r), ctypes.byref(ms)))
    return r.value, ms.value


def ntt(values: np.ndarray, inverse=False, timing=False):
    """values: uint8[nvec, 2^k, 32] or [2^k, 32] canonical LE; natural order in and out."""
    v = np.ascontiguousarray(values, dtype=np.uint8).copy()
    shp = v.shape
    nvec = 1 if v.ndim == 2 else shp[0]
    n = shp[-2]
    logn = n.bit_length() - 1
    assert 1 << logn == n
    ms = ctypes.c_float()
    _native.check(_native.lib().zkb_raw_ntt(v.ctypes.data, logn, nvec, 1 if inverse else 0, ctypes.byref(ms)))
    return (v, ms.value) if timing else v


def coset_ntt(values: np.ndarray):
    """NTT(coset_shift(iNTT(values))) - the prover's odd-coset evaluation (snarkjs groth16 prove)."""
    v = np.ascontiguousarray(values, dtype=np.uint8).copy()
    nvec = 1 if v.ndim == 2 else v.shape[0]
    n = v.shape[-2]
    logn = n.bit_length() - 1
    _native.check(_native.lib().zkb_raw_coset_ntt(v.ctypes.data, logn, nvec))
    return v


def _msm(fn, psize, bases, scalars, timing):
    b = np.ascontiguousarray(bases, dtype=np.uint8)
    s = np.ascontiguousarray(scalars, dtype=np.uint8)
    n = b.shape[0]
    nbatch = 1 if s.ndim == 2 else s.shape[0]
    assert s.shape[-2] == n
    out = np.zeros((nbatch, psize), dtype=np.uint8)
    kms, tms = ctypes.c_float(), ctypes.c_float()
    _native.check(fn(b.ctypes.data, n, s.ctypes.data, nbatch, out.ctypes.data,
                     ctypes.byref(kms) if timing else None, ctypes.byref(tms) if timing else None))
    return (out, kms.value, tms.value) if timing else out


def msm_g1(bases, scalars, timing=False, pair_tree=False, levels=0, group=0, variable_base=False):
    """bases: uint8[n, 64] canonical affine (zeros = infinity); scalars: uint8[(nbatch,) n, 32].
    pair_tree: sum the bucket lists with the batched-affine pair tree (zkb_raw_msm_g1_ex flag 1)."""
    if not pair_tree and not variable_base:
        return _msm(_native.lib().zkb_raw_msm_g1, 64, bases, scalars, timing)
    fn = _native.lib().zkb_raw_msm_g1_ex
    fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                   ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_uint32]
    flags = (2 if variable_base else 1) | (levels << 8) | (group << 16)
    return _msm(lambda *a: fn(*a, flags), 64, bases, scalars, timing)


def msm_g2(bases, scalars, timing=False):
    return _msm(_native.lib().zkb_raw_msm_g2, 128, bases, scalars, timing)


class MsmSession:
    """One rank's share of a 2^logn-point synthetic G1 MSM split by point range (include/zkcensus_b200.h,
    zkb_msm_session_*).  Ranks > 0 push their partial sum into rank 0's exchange buffer over NVLink; rank 0 combines."""

    def __init__(self, logn, rank=0, nranks=1, device=0, seed=1, window=16, variable_base=False):
        """variable_base: no precomputed window table (the bases are used as they are): one bucket set per window and a
        Horner combination at the end - the MSM for bases that are not key material."""
        self.h = ctypes.c_void_p()
        self.rank, self.nranks = rank, nranks
        _native.check(_native.lib().zkb_msm_session_create(device, logn, rank, nranks, seed,
                                                           window | (0x100 if variable_base else 0), ctypes.byref(self.h)))
        info = (ctypes.c_uint64 * 6)()
        _native.check(_native.lib().zkb_msm_session_info(self.h, info))
        self.points, self.sub_size, self.subs, self.window = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        self.table_ms, self.gen_ms = info[4] / 1e3, info[5] / 1e3

    def export_handle(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _native.check(_native.lib().zkb_msm_session_export(self.h, buf))
        return buf.raw

    def attach(self, handle: bytes):
        _native.check(_native.lib().zkb_msm_session_attach(self.h, ctypes.create_string_buffer(handle, 64)))

    def attach_local(self, root: "MsmSession"):
        _native.check(_native.lib().zkb_msm_session_attach_local(self.h, root.h))

    def run(self, wait=False):
        ms = ctypes.c_float()
        _native.check(_native.lib().zkb_msm_session_run(self.h, ctypes.byref(ms) if wait else None))
        return ms.value if wait else None

    def combine(self, nslots=None):
        """rank 0: (affine canonical point bytes, device ms from the start of this rank's run())."""
        out = ctypes.create_string_buffer(64)
        ms = ctypes.c_float()
        _native.check(_native.lib().zkb_msm_session_combine(self.h, self.nranks if nslots is None else nslots, out,
                                                            ctypes.byref(ms)))
        return out.raw, ms.value

    def madds(self) -> int:
        v = ctypes.c_uint64()
        _native.check(_native.lib().zkb_msm_session_madds(self.h, ctypes.byref(v)))
        return int(v.value)

    def read(self):
        b = np.zeros((self.points, 64), dtype=np.uint8)
        s = np.zeros((self.points, 32), dtype=np.uint8)
        _native.check(_native.lib().zkb_msm_session_read(self.h, b.ctypes.data, s.ctypes.data))
        return b, s

    def close(self):
        if self.h:
            _native.lib().zkb_msm_session_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ntt_bench(logn, nvec=1, iters=5, device=0):
    """(ms per inverse DIF transform incl. coset scale, ms per forward DIT transform), data resident."""
    a, b = ctypes.c_float(), ctypes.c_float()
    _native.check(_native.lib().zkb_ntt_bench(device, logn, nvec, iters, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


class NttDist:
    """One rank of the 4-step multi-GPU NTT (include/zkcensus_b200.h, zkb_ntt_dist_*)."""

    def __init__(self, logn, rank=0, nranks=1, device=0, seed=1):
        self.h = ctypes.c_void_p()
        self.logn, self.rank, self.nranks = logn, rank, nranks
        self.per = (1 << logn) // nranks
        _native.check(_native.lib().zkb_ntt_dist_create(device, logn, rank, nranks, seed, ctypes.byref(self.h)))

    def export_handle(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _native.check(_native.lib().zkb_ntt_dist_export(self.h, buf))
        return buf.raw

    def attach(self, peer_rank, handle: bytes):
        _native.check(_native.lib().zkb_ntt_dist_attach(self.h, peer_rank, ctypes.create_string_buffer(handle, 64)))

    def attach_local(self, peer_rank, peer: "NttDist"):
        _native.check(_native.lib().zkb_ntt_dist_attach_local(self.h, peer_rank, peer.h))

    def fill(self):
        _native.check(_native.lib().zkb_ntt_dist_fill(self.h))

    def run(self):
        _native.check(_native.lib().zkb_ntt_dist_run(self.h))

    def sync(self):
        """(total, column transforms, wait + P2P exchange, row transforms) in ms, device time."""
        ms = (ctypes.c_float * 4)()
        _native.check(_native.lib().zkb_ntt_dist_sync(self.h, ms))
        return tuple(float(x) for x in ms)

    def read(self, which):
        out = np.zeros((self.per, 32), dtype=np.uint8)
        _native.check(_native.lib().zkb_ntt_dist_read(self.h, which, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            _native.lib().zkb_ntt_dist_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ntt_dif_forward(values: np.ndarray):
    """Single-GPU forward transform, natural order in, bit-reversed order out (the decimation-in-frequency plan)."""
    v = np.ascontiguousarray(values, dtype=np.uint8).copy()
    logn = v.shape[0].bit_length() - 1
    _native.check(_native.lib().zkb_raw_ntt_dif_forward(v.ctypes.data, logn))
    return v
