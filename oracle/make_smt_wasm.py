#!/usr/bin/env python3
"""A SECOND circom program for the generic witness path (TEST INFRASTRUCTURE).

No circom compiler exists in the image and the reference ships one wasm, so this derives a different witness calculator
from it by binary surgery: the main component becomes the census circuit's `SMTVerifier(160)` sub-template (circomlib
smtverifier.circom, `circuit/census.circom:79-103`) instead of `ZkFranchiseProofCircuit`:

  * `init` creates SMTVerifier_160 at signal 1 and expects 169 input values instead of 334 (same-length LEB patches)
  * `setInputSignal` runs SMTVerifier_160_run when the last input has arrived
  * `getInputSize` returns 169
  * the input-name hashmap holds the template's nine inputs (enabled, root, siblings[161], oldKey, oldValue, isOld0,
    key, value, fnc) at signals 1..169
  * the witness map is the identity: wire w = signal w for the first 82,754 signals (no circom -O2 simplification)

Everything else - the field runtime, every template's code, the Poseidon constants - is the reference's, untouched.  The
result is a genuine circom-runtime program with a different main, different inputs and a different witness layout:
what `zkb_load_circuit` sees when it is handed a wasm that is not census.circom.

    python oracle/make_smt_wasm.py out.wasm          # write the patched wasm
"""
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import wasm_tools as W    # noqa: E402

INPUTS = [("enabled", 1), ("root", 1), ("siblings", 161), ("oldKey", 1), ("oldValue", 1), ("isOld0", 1), ("key", 1),
          ("value", 1), ("fnc", 1)]
N_IN = sum(s for _, s in INPUTS)


def _leb_u_padded(v, nbytes):
    out = bytearray()
    for i in range(nbytes):
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if i + 1 < nbytes else 0))
    assert v == 0
    return bytes(out)


def _leb_s_padded(v, nbytes):
    assert v >= 0 and v < (1 << (7 * nbytes - 1))
    return _leb_u_padded(v, nbytes)          # non-negative and the top payload bit stays clear: same encoding


def _leb_len(b, p):
    n = 1
    while b[p + n - 1] & 0x80:
        n += 1
    return n


def patch(wasm: bytes) -> bytes:
    mod = W.Module(wasm)
    b = bytearray(wasm)
    by_name = {f.name: f for f in mod.funcs}
    create, run = by_name["SMTVerifier_160_create"].idx, by_name["SMTVerifier_160_run"].idx
    main_create = by_name["ZkFranchiseProofCircuit_234_create"].idx
    main_run = by_name["ZkFranchiseProofCircuit_234_run"].idx

    def patch_const(fname, old, new):
        hits = 0
        for pos, op, imm in mod.instrs(by_name[fname]):
            if op == 0x41 and imm[0] == old:
                n = _leb_len(b, pos + 1)
                b[pos + 1:pos + 1 + n] = _leb_s_padded(new, n)
                hits += 1
        assert hits == 1, (fname, old, hits)

    def patch_call(fname, old, new):
        hits = 0
        for pos, op, imm in mod.instrs(by_name[fname]):
            if op == 0x10 and imm[0] == old:
                n = _leb_len(b, pos + 1)
                b[pos + 1:pos + 1 + n] = _leb_u_padded(new, n)
                hits += 1
        assert hits == 1, (fname, old, hits)

    n_in_old = W.getter_const(mod, "getInputSize")
    patch_const("init", n_in_old, N_IN)
    patch_const("getInputSize", n_in_old, N_IN)
    patch_call("init", main_create, create)
    patch_call("setInputSignal", main_run, run)
    # data segments: find them in the file by content
    n_wires = W.getter_const(mod, "getWitnessSize")
    for off, data in mod.data:
        at = wasm.find(bytes(data))
        if len(data) == 4096:                                   # input-name hashmap
            assert wasm.count(bytes(data)) == 1
            hm = bytearray(4096)
            pos = 1
            for name, size in INPUTS:
                h = W.fnv1a64(name)
                slot = h & 0xFF
                while hm[16 * slot:16 * slot + 8] != bytes(8) or hm[16 * slot + 12:16 * slot + 16] != bytes(4):
                    slot = (slot + 1) & 0xFF
                hm[16 * slot:16 * slot + 16] = struct.pack("<QII", h, pos, size)
                pos += size
            b[at:at + 4096] = hm
        elif len(data) == 4 * n_wires:                          # witness -> signal map: identity
            assert wasm.count(bytes(data)) == 1
            b[at:at + len(data)] = b"".join(struct.pack("<I", i) for i in range(n_wires))
    return bytes(b)


def smt_inputs(voter: dict) -> dict:
    """inputs of the patched program that make it check a census voter's Merkle proof (census.circom:79-89 wiring)"""
    return {"enabled": "1", "root": voter["censusRoot"], "siblings": list(voter["censusSiblings"]), "oldKey": "0",
            "oldValue": "0", "isOld0": "0", "key": voter["address"], "value": voter["availableWeight"], "fnc": "0"}


if __name__ == "__main__":
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(HERE), "artifacts", "zkCensus", "dev", "160", "circuit.wasm")
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/smt_verifier.wasm"
    open(out, "wb").write(patch(open(src, "rb").read()))
    print(out)
