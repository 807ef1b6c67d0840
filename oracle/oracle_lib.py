"""ctypes front-end of oracle/liboracle.so + snarkjs JSON <-> binary helpers (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
JSON layouts follow the reference fixtures `artifacts/zkCensus/dev/160/{proof,signals,
verification_key}.json`.
"""
import ctypes
import json
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(os.path.join(HERE, "liboracle.so"))
        vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        L.orc_setup.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, u64]
        L.orc_zkey_info.argtypes = [vp, sz, vp]
        L.orc_prove.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.orc_prove_shortcut.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.orc_h_scalars.argtypes = [vp, sz, vp, vp, vp]
        L.orc_verify.argtypes = [vp, vp, i32, vp, vp]
        L.orc_verify_many.argtypes = [vp, vp, i32, vp, vp, i32, vp]
        L.orc_alphabeta12.argtypes = [vp, vp, vp]
        L.orc_ntt.argtypes = [vp, i32, i32]
        L.orc_msm_g1.argtypes = [vp, vp, sz, vp]
        L.orc_msm_g2.argtypes = [vp, vp, sz, vp]
        L.orc_zkey_points.argtypes = [vp, sz, i32, vp, sz]
        L.orc_zkey_points.restype = ctypes.c_long
        L.orc_g1_mul_gen.argtypes = [vp, vp]
        L.orc_check_closed_form.argtypes = [ctypes.c_char_p, u64, vp, vp, vp, vp]
        L.orc_set_threads.argtypes = [i32]
        L.orc_chain_artifacts.argtypes = [ctypes.c_uint32, u64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, vp, i32]
        _lib = L
    return _lib


def le32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def from_le(b) -> int:
    return int.from_bytes(bytes(b), "little")


# ---- JSON <-> binary ---------------------------------------------------------------------------

def g1_bin(p):
    if p[2] == "0":
        return bytes(64)
    return le32(p[0]) + le32(p[1])


def g2_bin(p):
    if p[2][0] == "0" and p[2][1] == "0":
        return bytes(128)
    return le32(p[0][0]) + le32(p[0][1]) + le32(p[1][0]) + le32(p[1][1])


def proof_bin(proof: dict) -> bytes:
    """proof.json dict -> 256 bytes A.x A.y | B.x0 B.x1 B.y0 B.y1 | C.x C.y"""
    return g1_bin(proof["pi_a"]) + g2_bin(proof["pi_b"]) + g1_bin(proof["pi_c"])


def proof_json(b: bytes, snarkjs=False) -> dict:
    v = [str(from_le(b[i:i + 32])) for i in range(0, 256, 32)]
    d = {"pi_a": [v[0], v[1], "1"], "pi_b": [[v[2], v[3]], [v[4], v[5]], ["1", "0"]],
         "pi_c": [v[6], v[7], "1"]}
    if snarkjs:
        d["protocol"] = "groth16"
        d["curve"] = "bn128"
    return d


def vkey_bin(vk: dict):
    """-> (vk 448 bytes, ic bytes, nPublic)"""
    b = g1_bin(vk["vk_alpha_1"]) + g2_bin(vk["vk_beta_2"]) + g2_bin(vk["vk_gamma_2"]) + g2_bin(vk["vk_delta_2"])
    ic = b"".join(g1_bin(p) for p in vk["IC"])
    return b, ic, int(vk["nPublic"])


def pub_bin(pub) -> bytes:
    return b"".join(le32(int(x)) for x in pub)


def _buf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


def verify(vk: dict, pub, proof: dict) -> bool:
    vkb, ic, n = vkey_bin(vk)
    assert len(pub) == n
    return lib().orc_verify(_buf(vkb), _buf(ic), n, _buf(pub_bin(pub)), _buf(proof_bin(proof))) == 1


def verify_many(vk: dict, pubs: bytes, proofs: bytes, n: int):
    vkb, ic, npub = vkey_bin(vk)
    ok = np.zeros(n, dtype=np.uint8)
    lib().orc_verify_many(_buf(vkb), _buf(ic), npub, _buf(pubs), _buf(proofs), n, ok.ctypes.data)
    return ok


def alphabeta12(vk: dict):
    out = (ctypes.c_uint8 * 384)()
    lib().orc_alphabeta12(_buf(g1_bin(vk["vk_alpha_1"])), _buf(g2_bin(vk["vk_beta_2"])), out)
    v = [str(from_le(bytes(out)[i:i + 32])) for i in range(0, 384, 32)]
    return [[[v[(i * 3 + j) * 2], v[(i * 3 + j) * 2 + 1]] for j in range(3)] for i in range(2)]


def chain_artifacts(links: int, seed: int, out_dir: str, check=True):
    """Synthetic Poseidon-shaped chain circuit (BASELINE.json configs[3]) with `links` permutations: writes
    proving_key.zkey, verification_key.json, witness.wtns into out_dir; returns (nWires, nConstraints, domain)."""
    os.makedirs(out_dir, exist_ok=True)
    info = np.zeros(3, dtype=np.uint32)
    rc = lib().orc_chain_artifacts(links, seed, os.path.join(out_dir, "proving_key.zkey").encode(),
                                   os.path.join(out_dir, "verification_key.json").encode(),
                                   os.path.join(out_dir, "witness.wtns").encode(), info.ctypes.data, 1 if check else 0)
    if rc:
        raise RuntimeError(f"orc_chain_artifacts failed ({rc})")
    return tuple(int(x) for x in info)


class ZKeyRef:
    """Holds zkey bytes alive at a fixed address (the oracle caches its parse by pointer)."""

    def __init__(self, data: bytes):
        self.arr = np.frombuffer(data, dtype=np.uint8).copy()
        info = np.zeros(4, dtype=np.uint32)
        rc = lib().orc_zkey_info(self.arr.ctypes.data, self.arr.size, info.ctypes.data)
        if rc:
            raise ValueError("bad zkey")
        self.n_vars, self.n_public, self.domain, self.n_coefs = (int(x) for x in info)

    def prove(self, wtns: np.ndarray, r: int, s: int, partials=False):
        """wtns: uint8[nVars,32] -> 256-byte proof (and 384-byte MSM partials)."""
        w = np.ascontiguousarray(wtns, dtype=np.uint8)
        out = np.zeros(256, dtype=np.uint8)
        part = np.zeros(384, dtype=np.uint8)
        rc = lib().orc_prove(self.arr.ctypes.data, self.arr.size, w.ctypes.data, _buf(le32(r)), _buf(le32(s)),
                             out.ctypes.data, part.ctypes.data if partials else None)
        if rc:
            raise RuntimeError("orc_prove failed")
        return (out.tobytes(), part.tobytes()) if partials else out.tobytes()

    def prove_shortcut(self, wtns: np.ndarray, tmpl: np.ndarray, r: int, s: int):
        """Same proof as prove(), computed with the proof-independent-wire shortcut: the four witness
        MSMs run over (wtns - tmpl) only, the template sums are cached per (key, template)."""
        w = np.ascontiguousarray(wtns, dtype=np.uint8)
        t = np.ascontiguousarray(tmpl, dtype=np.uint8)
        out = np.zeros(256, dtype=np.uint8)
        rc = lib().orc_prove_shortcut(self.arr.ctypes.data, self.arr.size, w.ctypes.data, t.ctypes.data, _buf(le32(r)),
                                      _buf(le32(s)), out.ctypes.data)
        if rc:
            raise RuntimeError("orc_prove_shortcut failed")
        return out.tobytes()

    def h_scalars(self, wtns: np.ndarray, abc=False):
        w = np.ascontiguousarray(wtns, dtype=np.uint8)
        h = np.zeros((self.domain, 32), dtype=np.uint8)
        a = np.zeros((3, self.domain, 32), dtype=np.uint8) if abc else None
        lib().orc_h_scalars(self.arr.ctypes.data, self.arr.size, w.ctypes.data, h.ctypes.data,
                            a.ctypes.data if abc else None)
        return (h, a) if abc else h

    def points(self, which: int, max_points=None):
        size = 128 if which == 7 else 64
        n = {5: self.n_vars, 6: self.n_vars, 7: self.n_vars, 8: self.n_vars - self.n_public - 1,
             9: self.domain}[which]
        if max_points is not None:
            n = min(n, max_points)
        out = np.zeros((n, size), dtype=np.uint8)
        lib().orc_zkey_points(self.arr.ctypes.data, self.arr.size, which, out.ctypes.data, n)
        return out


def ntt(values: np.ndarray, inverse=False) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.uint8).copy()
    n = v.shape[0]
    logn = n.bit_length() - 1
    assert 1 << logn == n
    lib().orc_ntt(v.ctypes.data, logn, 1 if inverse else 0)
    return v


def msm_g1(bases: np.ndarray, scalars: np.ndarray) -> bytes:
    b = np.ascontiguousarray(bases, dtype=np.uint8)
    s = np.ascontiguousarray(scalars, dtype=np.uint8)
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_msm_g1(b.ctypes.data, s.ctypes.data, b.shape[0], out.ctypes.data)
    return out.tobytes()


def msm_g2(bases: np.ndarray, scalars: np.ndarray) -> bytes:
    b = np.ascontiguousarray(bases, dtype=np.uint8)
    s = np.ascontiguousarray(scalars, dtype=np.uint8)
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_msm_g2(b.ctypes.data, s.ctypes.data, b.shape[0], out.ctypes.data)
    return out.tobytes()


def setup(r1cs_path, zkey_path, vkey_path, seed=0x5EED):
    rc = lib().orc_setup(r1cs_path.encode(), zkey_path.encode(), vkey_path.encode(), seed)
    if rc:
        raise RuntimeError(f"orc_setup failed: {rc}")


def check_closed_form(r1cs_path, seed, wtns: np.ndarray, r: int, s: int, proof256: bytes) -> int:
    w = np.ascontiguousarray(wtns, dtype=np.uint8)
    return lib().orc_check_closed_form(r1cs_path.encode(), seed, w.ctypes.data, _buf(le32(r)), _buf(le32(s)),
                                       _buf(proof256))
