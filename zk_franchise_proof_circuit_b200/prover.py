"""Host-side mirror of the reference's prover interface, on top of the C ABI (libzkcensus_b200.so).

Mirrors, name for name:
  * Go  `go.vocdoni.io/dvote/crypto/zk/prover` as used by `zk_census_test.go:89-122`:
        Prove(zkey, wasm, inputs) -> Proof ; Proof.Bytes() -> (proofJSON, pubSignalsJSON) ;
        ParseProof(proofData, pubSignals) -> Proof
  * JS  `snarkjs.groth16.fullProve(inputs, wasmFile, zkeyFile)` as used by `ts_inputs/src/example.ts:358-362`
        -> {"proof": {..., "protocol": "groth16", "curve": "bn128"}, "publicSignals": [...]}

All proving happens on the GPU inside the shared library; this file only moves bytes.  A loaded circuit
(device-resident key tables) is cached per (zkey, wasm) digest, so repeated Prove calls pay the key
upload once - the reference re-parses the zkey on every call.
"""
import ctypes
import hashlib
import json
import os
import numpy as np

from . import _native
from ._native import NativeError  # noqa: F401  (re-export)

OK, ERROR, SHORT_BUFFER, INVALID_WITNESS_LENGTH, ASSERT_FAILED, UNSUPPORTED_CIRCUIT, INVALID_PROOF = range(7)

_vp, _sz, _i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_bound = False


def _lib():
    global _bound
    L = _native.lib()
    if not _bound:
        L.zkb_ctx_create.argtypes = [_i32, ctypes.POINTER(_vp)]
        L.zkb_ctx_destroy.argtypes = [_vp]
        L.zkb_ctx_stream.argtypes = [_vp]
        L.zkb_ctx_stream.restype = _vp
        L.zkb_load_circuit.argtypes = [_vp, _vp, _sz, _vp, _sz, ctypes.POINTER(_vp)]
        L.zkb_load_circuit_ex.argtypes = [_vp, _vp, _sz, _vp, _sz, ctypes.c_uint32, ctypes.POINTER(_vp)]
        L.zkb_verify_batch_bin.argtypes = [ctypes.c_char_p, _sz, _i32, _vp, _vp, _vp]
        L.zkb_load_circuit_shard.argtypes = [_vp, _vp, _sz, _i32, _i32, ctypes.POINTER(_vp)]
        L.zkb_shard_export.argtypes = [_vp, _vp]
        L.zkb_shard_attach.argtypes = [_vp, _vp]
        L.zkb_shard_attach_local.argtypes = [_vp, _vp]
        L.zkb_shard_export_witness.argtypes = [_vp, _vp]
        L.zkb_shard_attach_witness.argtypes = [_vp, _i32, _vp]
        L.zkb_shard_attach_witness_local.argtypes = [_vp, _i32, _vp]
        L.zkb_circuit_destroy.argtypes = [_vp]
        L.zkb_circuit_info.argtypes = [_vp, _vp]
        L.zkb_set_blinding.argtypes = [_vp, _vp, _vp]
        L.zkb_batch_set_inputs.argtypes = [_vp, _i32, _vp]
        L.zkb_batch_prove_resident.argtypes = [_vp, _i32, _vp]
        L.zkb_batch_get_results.argtypes = [_vp, _i32, _vp, _vp, _vp]
        L.zkb_batch_get_witness.argtypes = [_vp, _i32, _i32, _vp]
        L.zkb_debug_partials.argtypes = [_vp, _vp, _vp]
        L.zkb_poseidon_hash.argtypes = [_vp, _i32, _i32, _vp, _vp]
        L.zkb_launch_count.restype = ctypes.c_uint64
        L.zkb_verify.argtypes = [ctypes.c_char_p, _sz, ctypes.c_char_p, _sz, ctypes.c_char_p, _sz]
        L.zkb_verify_batch.argtypes = [ctypes.c_char_p, _sz, _i32, _vp, _vp, _vp, _vp, _vp]
        L.zkb_work_counters.argtypes = [_vp, _vp]
        L.zkb_fullprove_batch.argtypes = [_vp, _i32, _vp, _vp, _vp, _sz, _vp, _sz, _vp]
        L.zkb_fullprove.argtypes = [_vp, ctypes.c_char_p, _sz, _vp, ctypes.POINTER(_sz), _vp, ctypes.POINTER(_sz),
                                    _vp, _sz]
        L.zkb_witness.argtypes = [_vp, ctypes.c_char_p, _sz, _vp, ctypes.POINTER(_sz)]
        L.zkb_prove_wtns.argtypes = [_vp, _vp, _sz, _vp, ctypes.POINTER(_sz), _vp, ctypes.POINTER(_sz)]
        L.zkb_prove_wtns_stages.argtypes = [_vp, _vp, _sz, _vp, ctypes.POINTER(_sz), _vp, ctypes.POINTER(_sz), _vp]
        L.groth16_prover.argtypes = [_vp, ctypes.c_ulong, _vp, ctypes.c_ulong, _vp, ctypes.POINTER(ctypes.c_ulong),
                                     _vp, ctypes.POINTER(ctypes.c_ulong), _vp, ctypes.c_ulong]
        _bound = True
    return L


INPUT_ORDER = ["electionId", "nullifier", "voteHash", "sikRoot", "censusRoot", "voteWeight", "availableWeight",
               "address", "password", "signature", "censusSiblings", "sikSiblings"]
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class Context:
    """One GPU (one process per GPU in multi-GPU runs: pick device = LOCAL_RANK)."""

    def __init__(self, device=0):
        self.h = _vp()
        _native.check(_lib().zkb_ctx_create(device, ctypes.byref(self.h)))
        self.device = device

    @property
    def stream(self):
        """Raw cudaStream_t of the pipeline (wrap with torch.cuda.ExternalStream to time with torch events)."""
        return _lib().zkb_ctx_stream(self.h)

    def close(self):
        if self.h:
            _lib().zkb_ctx_destroy(self.h)
            self.h = _vp()


class Circuit:
    """A proving key + witness calculator loaded onto the GPU (zkb_load_circuit)."""

    def __init__(self, ctx: Context, zkey: bytes, wasm: bytes = None, shard=None, dense=False, generic=False):
        """shard = (rank, nranks): keep only this rank's point ranges of the key (zkb_load_circuit_shard).
        dense = True: ZKB_LOAD_DENSE, the proof-independent-wire shortcut is off (measurement aid).
        generic = True: ZKB_LOAD_GENERIC_WITNESS, witness by the program extracted from the wasm (any circom circuit)."""
        self.ctx = ctx
        self.dense = dense
        self.h = _vp()
        self.shard = shard
        zb = (ctypes.c_char * len(zkey)).from_buffer_copy(zkey)
        wb = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm) if wasm else None
        if shard:
            _native.check(_lib().zkb_load_circuit_shard(ctx.h, ctypes.addressof(zb), len(zkey), shard[0], shard[1],
                                                        ctypes.byref(self.h)))
        else:
            _native.check(_lib().zkb_load_circuit_ex(ctx.h, ctypes.addressof(zb), len(zkey),
                                                     ctypes.addressof(wb) if wasm else None, len(wasm) if wasm else 0,
                                                     (1 if dense else 0) | (2 if generic else 0), ctypes.byref(self.h)))
        info = np.zeros(8, dtype=np.uint32)
        _lib().zkb_circuit_info(self.h, info.ctypes.data)
        self.n_vars, self.n_public, self.domain, self.n_inputs, self.n_levels1 = (int(x) for x in info[:5])

    def shard_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _native.check(_lib().zkb_shard_export(self.h, buf))
        return buf.raw

    def shard_attach(self, handle: bytes):
        _native.check(_lib().zkb_shard_attach(self.h, ctypes.create_string_buffer(handle, 64)))

    def shard_attach_local(self, root: "Circuit"):
        _native.check(_lib().zkb_shard_attach_local(self.h, root.h))

    def shard_export_witness(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _native.check(_lib().zkb_shard_export_witness(self.h, buf))
        return buf.raw

    def shard_attach_witness(self, peer_rank: int, handle: bytes):
        _native.check(_lib().zkb_shard_attach_witness(self.h, peer_rank, ctypes.create_string_buffer(handle, 64)))

    def shard_attach_witness_local(self, peer_rank: int, peer: "Circuit"):
        _native.check(_lib().zkb_shard_attach_witness_local(self.h, peer_rank, peer.h))

    def close(self):
        if self.h:
            _lib().zkb_circuit_destroy(self.h)
            self.h = _vp()

    def set_blinding(self, r=None, s=None):
        """Pin the Groth16 blinding scalars (tests); None restores random r, s."""
        if r is None or s is None:
            _native.check(_lib().zkb_set_blinding(self.h, None, None))
        else:
            rb = (ctypes.c_uint8 * 32).from_buffer_copy(int(r).to_bytes(32, "little"))
            sb = (ctypes.c_uint8 * 32).from_buffer_copy(int(s).to_bytes(32, "little"))
            _native.check(_lib().zkb_set_blinding(self.h, rb, sb))

    def _public_cap(self):
        # public.json: one decimal string of up to 77 digits per public signal, quotes, separators
        return max(2048, 96 * self.n_public + 64)

    # ---- reference-shaped calls (JSON in, JSON out; host<->device copies inside) --------------
    def fullprove_batch(self, inputs_json):
        """inputs_json: list of bytes/str documents.  Returns (proofs, publics, status) lists."""
        n = len(inputs_json)
        docs = [d if isinstance(d, bytes) else d.encode() for d in inputs_json]
        arr = (ctypes.c_char_p * n)(*docs)
        lens = (ctypes.c_size_t * n)(*[len(d) for d in docs])
        ps, qs = 1024, self._public_cap()
        pbuf = ctypes.create_string_buffer(ps * n)
        qbuf = ctypes.create_string_buffer(qs * n)
        status = (ctypes.c_int * n)()
        _native.check(_lib().zkb_fullprove_batch(self.h, n, arr, lens, pbuf, ps, qbuf, qs, status))
        praw, qraw = pbuf.raw, qbuf.raw          # one copy each (.raw copies the whole buffer every time it is read)
        proofs = [praw[i * ps:praw.index(b"\0", i * ps)] for i in range(n)]
        pubs = [qraw[i * qs:qraw.index(b"\0", i * qs)] for i in range(n)]
        return proofs, pubs, list(status)

    def fullprove(self, inputs_json):
        doc = inputs_json if isinstance(inputs_json, bytes) else inputs_json.encode()
        qcap = self._public_cap()
        pbuf, qbuf, ebuf = ctypes.create_string_buffer(1024), ctypes.create_string_buffer(qcap), ctypes.create_string_buffer(256)
        pn, qn = ctypes.c_size_t(1024), ctypes.c_size_t(qcap)
        rc = _lib().zkb_fullprove(self.h, doc, len(doc), pbuf, ctypes.byref(pn), qbuf, ctypes.byref(qn), ebuf, 256)
        if rc:
            raise NativeError(rc, ebuf.value.decode(errors="replace"))
        return pbuf.raw[:pn.value], qbuf.raw[:qn.value]

    def witness(self, inputs_json) -> bytes:
        """.wtns file contents for one inputs.json (what the reference's wasm calculator produces)."""
        doc = inputs_json if isinstance(inputs_json, bytes) else inputs_json.encode()
        n = ctypes.c_size_t(0)
        _lib().zkb_witness(self.h, doc, len(doc), None, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        _native.check(_lib().zkb_witness(self.h, doc, len(doc), buf, ctypes.byref(n)))
        return buf.raw[:n.value]

    def prove_wtns(self, wtns: bytes, stages=False):
        """Groth16 from a .wtns (go-rapidsnark Groth16ProverRaw).  stages=True also returns the 8 device stage times.
        wtns=None proves again from the witness already resident on the device."""
        qcap = self._public_cap()
        pbuf, qbuf = ctypes.create_string_buffer(1024), ctypes.create_string_buffer(qcap)
        pn, qn = ctypes.c_size_t(1024), ctypes.c_size_t(qcap)
        # the bytes object is handed over as it is (no copy: round 1 copied the 134 MB witness of the 2^22-constraint
        # circuit into a fresh ctypes array on every call, ~60 ms of the measured wall time)
        wb = ctypes.cast(ctypes.c_char_p(wtns), _vp) if wtns is not None else None
        st = np.zeros(8, dtype=np.float32)
        _native.check(_lib().zkb_prove_wtns_stages(self.h, wb, len(wtns) if wtns is not None else 0, pbuf,
                                                   ctypes.byref(pn), qbuf, ctypes.byref(qn),
                                                   st.ctypes.data if stages else None))
        out = (pbuf.raw[:pn.value], qbuf.raw[:qn.value])
        return out + (st,) if stages else out

    def poseidon(self, rows):
        """Batched Poseidon on the GPU: rows = list of equal-length tuples of ints (arity 2..4) -> list of ints."""
        n = len(rows)
        if n == 0:
            return []
        arity = len(rows[0])
        buf = np.frombuffer(b"".join(int(x % R_MOD).to_bytes(32, "little") for r in rows for x in r), dtype=np.uint8)
        out = np.zeros((n, 32), dtype=np.uint8)
        _native.check(_lib().zkb_poseidon_hash(self.h, arity, n, buf.ctypes.data, out.ctypes.data))
        return [int.from_bytes(out[i].tobytes(), "little") for i in range(n)]

    def census_tree(self, leaves: dict, n_levels=160):
        """arbo-style Poseidon sparse Merkle tree of {key: value} built by the library (zkb_census_tree: structure in
        C++, hashing on the GPU).  Returns (root, {key: [siblings, without the zero padding]})."""
        keys = list(leaves)
        n = len(keys)
        kb = np.frombuffer(b"".join(int(k).to_bytes(32, "little") for k in keys), dtype=np.uint8)
        vb = np.frombuffer(b"".join(int(leaves[k] % R_MOD).to_bytes(32, "little") for k in keys), dtype=np.uint8)
        root = np.zeros(32, dtype=np.uint8)
        sib = np.zeros((n, n_levels + 1, 32), dtype=np.uint8)
        L = _lib()
        L.zkb_census_tree.argtypes = [_vp, _i32, _vp, _vp, _i32, _vp, _vp]
        _native.check(L.zkb_census_tree(self.h, n, kb.ctypes.data, vb.ctypes.data, n_levels, root.ctypes.data, sib.ctypes.data))
        out = {}
        for i, k in enumerate(keys):
            row = [int.from_bytes(sib[i, j].tobytes(), "little") for j in range(n_levels + 1)]
            while row and row[-1] == 0:
                row.pop()
            out[k] = row
        return int.from_bytes(root.tobytes(), "little"), out

    # ---- resident path (inputs already in HBM; used for device-only timing) ---------------------
    def set_inputs(self, packed: np.ndarray):
        """packed: uint8[n, n_inputs, 32] canonical values in main-signal order (see pack_inputs)."""
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        assert p.shape[1:] == (self.n_inputs, 32)
        _native.check(_lib().zkb_batch_set_inputs(self.h, p.shape[0], p.ctypes.data))
        self._resident = p.shape[0]

    def prove_resident(self, n=None, stages=False):
        n = n or self._resident
        st = np.zeros(8, dtype=np.float32)
        _native.check(_lib().zkb_batch_prove_resident(self.h, n, st.ctypes.data if stages else None))
        return st if stages else None

    def work_counters(self):
        """Executed MSM work of the last chunk, per proof (exact counts from the device)."""
        o = np.zeros(8, dtype=np.uint64)
        _native.check(_lib().zkb_work_counters(self.h, o.ctypes.data))
        m = int(o[4])
        return {"g1_madds_per_proof": float(o[0]) / m, "g2_madds_per_proof": float(o[1]) / m,
                "witness_digit_entries_per_proof": float(o[2]) / m, "h_digit_entries_per_proof": float(o[3]) / m,
                "chunk": int(o[5]), "g1_affine_adds_per_proof": float(o[6]) / m,
                "g1_inversions_per_proof": float(o[7]) / m}

    def get_results(self, n=None):
        n = n or self._resident
        proofs = np.zeros((n, 256), dtype=np.uint8)
        pubs = np.zeros((n, self.n_public, 32), dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        _native.check(_lib().zkb_batch_get_results(self.h, n, proofs.ctypes.data, pubs.ctypes.data, status.ctypes.data))
        return proofs, pubs, status

    def debug_partials(self, with_h=False):
        """MSM partial sums (384 B) and H scalars of the first proof of the last chunk (parity debugging)."""
        out = np.zeros(384, dtype=np.uint8)
        h = np.zeros((self.domain, 32), dtype=np.uint8) if with_h else None
        _native.check(_lib().zkb_debug_partials(self.h, out.ctypes.data, h.ctypes.data if with_h else None))
        return (out.tobytes(), h) if with_h else out.tobytes()

    def get_witness(self, first=0, n=1):
        w = np.zeros((n, self.n_vars, 32), dtype=np.uint8)
        _native.check(_lib().zkb_batch_get_witness(self.h, first, n, w.ctypes.data))
        return w


def launch_count() -> int:
    """Kernels of this library launched by the proving pipeline since the process started."""
    return int(_lib().zkb_launch_count())


def pack_inputs(inputs: dict) -> np.ndarray:
    """inputs.json dict -> uint8[n_inputs, 32] canonical values in the circuit's main-signal order."""
    vals = []
    for k in INPUT_ORDER:
        v = inputs[k]
        vals += [int(x) % R_MOD for x in (v if isinstance(v, list) else [v])]
    return np.frombuffer(b"".join(x.to_bytes(32, "little") for x in vals), dtype=np.uint8).reshape(-1, 32).copy()


# ---- Go-shaped API (go.vocdoni.io/dvote/crypto/zk/prover) -----------------------------------------

class Proof:
    """prover.Proof{Data{A "pi_a", B "pi_b", C "pi_c"}, PubSignals}."""

    def __init__(self, data: dict, pub_signals: list):
        self.data = data
        self.pub_signals = pub_signals

    def verify(self, vkey: bytes) -> None:
        """(*Proof).Verify(vkey []byte) error  -  zk_census_test.go:122.  Raises NativeError(INVALID_PROOF) when the
        proof does not verify; the pairing check runs on the GPU (zkb_verify)."""
        pj, sj = self.bytes()
        verify(vkey, sj, pj)

    def bytes(self):
        """(*Proof).Bytes(): compact JSON of pi_a/pi_b/pi_c only, and of the public signals (zk_census_test.go:93)."""
        d = {k: self.data[k] for k in ("pi_a", "pi_b", "pi_c")}
        return (json.dumps(d, separators=(",", ":")).encode(), json.dumps(self.pub_signals, separators=(",", ":")).encode())


_ctx = {}
_circuits = {}


def _context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ctx:
        _ctx[device] = Context(device)
    return _ctx[device]


def load(zkey: bytes, wasm: bytes = None, device=None, dense=False, generic=False) -> Circuit:
    """Cached zkb_load_circuit."""
    key = (hashlib.sha256(zkey).digest(), hashlib.sha256(wasm).digest() if wasm else None, device, dense, generic)
    if key not in _circuits:
        _circuits[key] = Circuit(_context(device), zkey, wasm, dense=dense, generic=generic)
    return _circuits[key]


def load_shard(zkey: bytes, rank: int, nranks: int, device=None) -> Circuit:
    """One rank's share of a key sharded by point range over nranks GPUs (not cached)."""
    return Circuit(_context(device), zkey, None, shard=(rank, nranks))


def prove(zkey: bytes, wasm: bytes, inputs: bytes) -> Proof:
    """prover.Prove(zkey, wasm, inputs []byte) (*Proof, error)  -  zk_census_test.go:89"""
    c = load(zkey, wasm)
    pj, sj = c.fullprove(inputs)
    return Proof(json.loads(pj), json.loads(sj))


def parse_proof(proof_data: bytes, pub_signals: bytes) -> Proof:
    """prover.ParseProof(proofData, pubSignals []byte) (*Proof, error)  -  zk_census_test.go:118"""
    d = json.loads(proof_data)
    for k in ("pi_a", "pi_b", "pi_c"):
        if k not in d:
            raise ValueError(f"proof: missing {k}")
    return Proof(d, json.loads(pub_signals))


def verify(vkey: bytes, public_signals: bytes, proof: bytes) -> None:
    """Groth16 verification of JSON documents on the GPU; raises NativeError (code INVALID_PROOF) if invalid."""
    as_b = lambda x: x if isinstance(x, bytes) else (x.encode() if isinstance(x, str) else json.dumps(x).encode())
    vk, ps, pf = as_b(vkey), as_b(public_signals), as_b(proof)
    _native.check(_lib().zkb_verify(vk, len(vk), ps, len(ps), pf, len(pf)))


def verify_batch(vkey: bytes, public_signals: list, proofs: list) -> list:
    """ok flags (1/0) for n (public signals, proof) JSON pairs under one verification key (zkb_verify_batch)."""
    n = len(proofs)
    as_b = lambda x: x if isinstance(x, bytes) else (x.encode() if isinstance(x, str) else json.dumps(x).encode())
    vk = as_b(vkey)
    ps, pf = [as_b(x) for x in public_signals], [as_b(x) for x in proofs]
    pa, fa = (ctypes.c_char_p * n)(*ps), (ctypes.c_char_p * n)(*pf)
    pl, fl = (ctypes.c_size_t * n)(*[len(x) for x in ps]), (ctypes.c_size_t * n)(*[len(x) for x in pf])
    ok = (ctypes.c_int * n)()
    _native.check(_lib().zkb_verify_batch(vk, len(vk), n, pa, pl, fa, fl, ok))
    return list(ok)


def verify_batch_bin(vkey: bytes, publics: np.ndarray, proofs256: np.ndarray) -> np.ndarray:
    """ok flags for binary results as Circuit.get_results returns them (zkb_verify_batch_bin)."""
    vk = vkey if isinstance(vkey, bytes) else (vkey.encode() if isinstance(vkey, str) else json.dumps(vkey).encode())
    pr = np.ascontiguousarray(proofs256, dtype=np.uint8)
    pb = np.ascontiguousarray(publics, dtype=np.uint8)
    n = pr.shape[0]
    ok = np.zeros(n, dtype=np.int32)
    _native.check(_lib().zkb_verify_batch_bin(vk, len(vk), n, pb.ctypes.data, pr.ctypes.data, ok.ctypes.data))
    return ok


def export_vkey(zkey: bytes) -> bytes:
    """`snarkjs zkey export verificationkey` (circuit/circuit-compiler.sh:128-134): verification_key.json of a proving
    key, the exact text snarkjs writes (zkb_export_vkey; host arithmetic, no GPU needed)."""
    L = _native.lib()
    L.zkb_export_vkey.argtypes = [_vp, _sz, _vp, ctypes.POINTER(_sz)]
    zb = (ctypes.c_char * len(zkey)).from_buffer_copy(zkey)
    n = _sz(0)
    L.zkb_export_vkey(ctypes.addressof(zb), len(zkey), None, ctypes.byref(n))
    buf = ctypes.create_string_buffer(n.value)
    _native.check(L.zkb_export_vkey(ctypes.addressof(zb), len(zkey), buf, ctypes.byref(n)))
    return buf.raw[:n.value]


# ---- snarkjs-shaped API -------------------------------------------------------------------------------

class groth16:  # noqa: N801  (mirrors `import { groth16 } from "snarkjs"`)
    @staticmethod
    def verify(vkey, public_signals, proof) -> bool:
        """groth16.verify(vKey, publicSignals, proof) -> boolean"""
        try:
            verify(vkey, public_signals, proof)
            return True
        except NativeError as e:
            if e.code == INVALID_PROOF:
                return False
            raise

    @staticmethod
    def full_prove(inputs, wasm_file, zkey_file):
        """groth16.fullProve(input, wasmFile, zkeyFile) -> {proof, publicSignals}  -  ts_inputs/src/example.ts:358"""
        zkey = open(zkey_file, "rb").read()
        wasm = open(wasm_file, "rb").read()
        doc = inputs if isinstance(inputs, (bytes, str)) else json.dumps(inputs)
        c = load(zkey, wasm)
        pj, sj = c.fullprove(doc)
        proof = json.loads(pj)
        proof["protocol"] = "groth16"
        proof["curve"] = "bn128"
        return {"proof": proof, "publicSignals": json.loads(sj)}
