"""ctypes wrapper around oracle/_ref/libcensus_wasm.so (TEST INFRASTRUCTURE, not product).

`libcensus_wasm.so` is the reference's own witness calculator (`circuit.wasm`) transpiled to C by
`oracle/wasm2c.py`; this module feeds it an inputs.json the way go-rapidsnark/witness does
(`zk_census_test.go:89`): keys flattened in order, FNV-1a-64 name hash, decimal strings -> 8 LE
32-bit words.
"""
import ctypes
import json
import os
import numpy as np
from wasm_tools import fnv1a64

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617

_lib = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "libcensus_wasm.so")) and \
        os.path.exists(os.path.join(REF_DIR, "census_wasm.mem"))


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(REF_DIR, "libcensus_wasm.so"))
        _lib.wc_load.argtypes = [ctypes.c_char_p]
        _lib.wc_witness.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int]
        rc = _lib.wc_load(os.path.join(REF_DIR, "census_wasm.mem").encode())
        if rc != 0:
            raise RuntimeError(f"wc_load failed: {rc}")
    return _lib


def flatten_inputs(inputs: dict):
    """[(name_hash, index, value)] in JSON key order, arrays flattened (circom_runtime order)."""
    out = []
    for k, v in inputs.items():
        h = fnv1a64(k)
        vs = v if isinstance(v, list) else [v]
        for i, x in enumerate(vs):
            out.append((h, i, int(x) % R_MOD))
    return out


def witness(inputs: dict, sanity=True):
    """Returns (code, witness) with witness = np.uint8[nWires, 32] (LE, canonical form)."""
    L = lib()
    flat = flatten_inputs(inputs)
    n = len(flat)
    hashes = np.array([f[0] for f in flat], dtype=np.uint64)
    idx = np.array([f[1] for f in flat], dtype=np.uint32)
    vals = np.frombuffer(b"".join(f[2].to_bytes(32, "little") for f in flat), dtype=np.uint32).copy()
    nw = L.wc_witness_size()
    out = np.zeros((nw, 32), dtype=np.uint8)
    code = L.wc_witness(n, hashes.ctypes.data, idx.ctypes.data, vals.ctypes.data, out.ctypes.data,
                        1 if sanity else 0)
    return code, out


class RefWasm:
    """A second transpiled witness calculator beside the census one (oracle/_ref/<sub>/lib<name>.so + census_wasm.mem):
    the SMTVerifier program of oracle/make_smt_wasm.py, built by `make ref_smt`."""

    def __init__(self, sub="smt", libname="libsmt_wasm.so"):
        self.dir = os.path.join(REF_DIR, sub)
        self.path = os.path.join(self.dir, libname)
        self._lib = None

    def available(self):
        return os.path.exists(self.path) and os.path.exists(os.path.join(self.dir, "census_wasm.mem"))

    def lib(self):
        if self._lib is None:
            L = ctypes.CDLL(self.path)
            L.wc_load.argtypes = [ctypes.c_char_p]
            L.wc_witness.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            rc = L.wc_load(os.path.join(self.dir, "census_wasm.mem").encode())
            if rc != 0:
                raise RuntimeError(f"wc_load failed: {rc}")
            self._lib = L
        return self._lib

    def witness(self, inputs: dict, sanity=True):
        L = self.lib()
        flat = flatten_inputs(inputs)
        n = len(flat)
        hashes = np.array([f[0] for f in flat], dtype=np.uint64)
        idx = np.array([f[1] for f in flat], dtype=np.uint32)
        vals = np.frombuffer(b"".join(f[2].to_bytes(32, "little") for f in flat), dtype=np.uint32).copy()
        nw = L.wc_witness_size()
        out = np.zeros((nw, 32), dtype=np.uint8)
        code = L.wc_witness(n, hashes.ctypes.data, idx.ctypes.data, vals.ctypes.data, out.ctypes.data, 1 if sanity else 0)
        return code, out


if __name__ == "__main__":
    import hashlib
    import sys
    import time
    p = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/artifacts/zkCensus/dev/160/inputs_example.json"
    inp = json.load(open(p))
    t = time.time()
    code, w = witness(inp)
    print("code", code, "time", time.time() - t)
    print("sha256", hashlib.sha256(w.tobytes()).hexdigest())
    for i in list(range(0, 13)) + [333, 334, 335, 336]:
        print(i, int.from_bytes(w[i].tobytes(), "little"))
