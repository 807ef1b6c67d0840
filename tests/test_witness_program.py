"""CPU tests of the generic witness path: csrc/wasm_symexec.cc extracts a straight-line witness
program from the reference's circuit.wasm by symbolic execution; evaluated with csrc/tape_ops.cuh (the code the GPU
lanes run, built for the host here) it must reproduce the reference wasm's witness bit for bit, including the
circuit asserts (exception code 4)."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H

INPUT_ORDER = ["electionId", "nullifier", "voteHash", "sikRoot", "censusRoot", "voteWeight", "availableWeight",
               "address", "password", "signature", "censusSiblings", "sikSiblings"]


@pytest.fixture(scope="module")
def tape(art_dir, tmp_path_factory):
    src = os.path.join(H.ROOT, "tests", "host_emul", "tape_host.cc")
    so = str(tmp_path_factory.mktemp("tape") / "libtape_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = ctypes.CDLL(so)
    L.tape_host_build.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    L.tape_host_input.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    L.tape_host_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    wasm = open(art_dir + "/circuit.wasm", "rb").read()
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    info = np.zeros(8, dtype=np.uint32)
    err = ctypes.create_string_buffer(512)
    assert L.tape_host_build(ctypes.addressof(buf), len(wasm), info.ctypes.data, err, 512) == 0, err.value
    return L, info


def _pack(L, inputs, n_inputs):
    out = np.zeros((n_inputs, 32), dtype=np.uint8)
    for name, v in inputs.items():
        size = ctypes.c_uint32(0)
        pos = L.tape_host_input(name.encode(), ctypes.byref(size))
        vals = v if isinstance(v, list) else [v]
        assert pos >= 0 and size.value == len(vals), name
        for k, x in enumerate(vals):
            out[pos + k] = np.frombuffer((int(x) % H.census_gen.P).to_bytes(32, "little"), dtype=np.uint8)
    return out


def _eval(L, info, inputs):
    packed = _pack(L, inputs, int(info[0]))
    w = np.zeros((int(info[1]), 32), dtype=np.uint8)
    return L.tape_host_eval(packed.ctypes.data, w.ctypes.data), w


def test_program_shape(tape):
    L, info = tape
    n_inputs, n_wires, n_slots, n_ops, n_levels, n_consts, n_asserts, n_selects = (int(x) for x in info)
    assert (n_inputs, n_wires) == (334, 82754)
    # 328 Poseidon permutations (~196 k products) + the rest; far fewer slots than values
    assert 250_000 < n_ops < 500_000 and n_slots < 100_000 and n_levels < n_ops
    assert n_asserts > 100 and n_selects == 654          # IsZero's `in != 0 ? 1/in : 0`: 322 + 5 per tree x 2
    # the flat input order is the signal order of the wasm's hashmap == the census circuit's main-signal order
    pos = 0
    for name in INPUT_ORDER:
        size = ctypes.c_uint32(0)
        assert L.tape_host_input(name.encode(), ctypes.byref(size)) == pos, name
        pos += size.value
    assert pos == 334


def test_generic_witness_fixture_bit_exact(tape):
    L, info = tape
    code, w = _eval(L, info, H.fixture_inputs())
    assert code == 0
    assert H.sha(w.tobytes()) == H.WITNESS_SHA256           # the reference wasm's witness (golden KAT)


def test_generic_witness_synthetic_and_deep_voters(tape):
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    L, info = tape
    for v in list(H.voters(3)) + H.deep_voters():
        code, w = _eval(L, info, v)
        rc, ref = RW.witness(v)
        assert code == 0 and rc == 0
        assert np.array_equal(w, ref)


def test_generic_witness_asserts_match_wasm(tape):
    """inputs the circuit rejects: the tape reports 4 exactly where the wasm calls exceptionHandler(4)"""
    import ref_witness as RW
    L, info = tape
    inp = H.fixture_inputs()
    for patch in ({"voteWeight": "11"}, {"nullifier": "5"}, {"censusRoot": "7"}, {"availableWeight": "4"}):
        bad = dict(inp)
        bad.update(patch)
        code, _ = _eval(L, info, bad)
        assert code == 4, patch
        if RW.available():
            assert RW.witness(bad)[0] == 4


# ---- a second circom program: SMTVerifier(160) as the main component (oracle/make_smt_wasm.py) -----------------------

@pytest.fixture(scope="module")
def smt_tape(art_dir, tmp_path_factory):
    import make_smt_wasm as SW
    src = os.path.join(H.ROOT, "tests", "host_emul", "tape_host.cc")
    so = str(tmp_path_factory.mktemp("tape_smt") / "libtape_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = ctypes.CDLL(so)
    L.tape_host_build.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    L.tape_host_input.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    L.tape_host_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    wasm = SW.patch(open(art_dir + "/circuit.wasm", "rb").read())
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    info = np.zeros(8, dtype=np.uint32)
    err = ctypes.create_string_buffer(512)
    assert L.tape_host_build(ctypes.addressof(buf), len(wasm), info.ctypes.data, err, 512) == 0, err.value
    return L, info, wasm


def test_second_program_is_a_different_circuit(smt_tape):
    L, info, wasm = smt_tape
    n_inputs, n_wires, n_slots, n_ops, n_levels, n_consts, n_asserts, n_selects = (int(x) for x in info)
    assert (n_inputs, n_wires) == (169, 82754)
    assert 150_000 < n_ops < 250_000 and n_selects == 326 and n_asserts > 500
    kat = json.load(open(H.GOLDEN + "/smt_verifier_kat.json"))
    assert H.sha(wasm) == kat["wasm_sha256"]
    size = ctypes.c_uint32(0)
    assert L.tape_host_input(b"siblings", ctypes.byref(size)) == 2 and size.value == 161
    assert L.tape_host_input(b"censusSiblings", ctypes.byref(size)) == -1


def test_second_program_witness_matches_native_run_and_kat(smt_tape):
    """The SMTVerifier program through the generic path == the same wasm transpiled to C and run natively
    (oracle/_ref/smt, `make ref_smt`) == the committed known answers; a wrong root raises exception 4 in both."""
    import make_smt_wasm as SW
    import ref_witness as RW
    L, info, _ = smt_tape
    kat = json.load(open(H.GOLDEN + "/smt_verifier_kat.json"))
    ref = RW.RefWasm()
    voters = [H.fixture_inputs()] + list(H.voters(2)) + [H.deep_voters()[0]]
    for i, v in enumerate(voters):
        inp = SW.smt_inputs(v)
        code, w = _eval(L, info, inp)
        assert code == 0
        assert H.sha(w.tobytes()) == kat["witness_sha256"][i], f"case {i}"
        if ref.available():
            rc, rw = ref.witness(inp)
            assert rc == 0 and np.array_equal(rw, w)
    bad = SW.smt_inputs(voters[0])
    bad["root"] = "7"
    assert _eval(L, info, bad)[0] == 4
    if ref.available():
        assert ref.witness(bad)[0] == 4
