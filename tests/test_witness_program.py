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


def test_every_tape_operation_matches_the_wasm_runtime(tape, art_dir):
    """tape_ops.cuh (what the GPU lanes execute) against the reference wasm's own Fr_* functions run by the interpreter,
    for every operation the extractor can emit - including those the census program never uses - on random and edge
    operands (0, 1, r-1, the sign boundary (r-1)/2, shift amounts around 254 and negative ones)."""
    L, _ = tape
    wasm = open(art_dir + "/circuit.wasm", "rb").read()
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    L.tape_host_rt_load.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.tape_host_rt_op.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.tape_host_apply.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    assert L.tape_host_rt_load(ctypes.addressof(buf), len(wasm)) == 0
    P = H.census_gen.P
    ops = [("Fr_add", 0, 0), ("Fr_sub", 1, 0), ("Fr_mul", 2, 0), ("Fr_neg", 3, 1), ("Fr_div", 4, 0), ("Fr_inv", 5, 1),
           ("Fr_eq", 6, 0), ("Fr_neq", 7, 0), ("Fr_lt", 8, 0), ("Fr_gt", 9, 0), ("Fr_leq", 10, 0), ("Fr_geq", 11, 0),
           ("Fr_land", 12, 0), ("Fr_lor", 13, 0), ("Fr_lnot", 14, 1), ("Fr_shr", 15, 0), ("Fr_shl", 16, 0),
           ("Fr_band", 17, 0), ("Fr_bor", 18, 0), ("Fr_bxor", 19, 0), ("Fr_bnot", 20, 1),
           ("Fr_idiv", 26, 0), ("Fr_mod", 27, 0), ("Fr_pow", 28, 0)]
    import random
    rnd = random.Random(7)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 253, (1 << 253) - 1, 12345]
    shifts = [0, 1, 31, 32, 63, 64, 200, 253, 254, 255, 300, P - 1, P - 5, P - 253, P - 254, P - 255]
    le = lambda x: (ctypes.c_uint8 * 32).from_buffer_copy(int(x).to_bytes(32, "little"))
    for name, op, unary in ops:
        pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(40)]
        if name in ("Fr_shr", "Fr_shl"):
            pairs = [(a, s) for a in edge + [rnd.randrange(P) for _ in range(6)] for s in shifts]
        if name in ("Fr_div", "Fr_inv", "Fr_pow"):
            pairs = [(a, b) for a, b in pairs[:60]]
        if name in ("Fr_idiv", "Fr_mod"):
            pairs += [(rnd.randrange(P), rnd.randrange(1, 1 << k)) for k in (1, 8, 32, 33, 64, 128, 200, 253)]
        for a, b in pairs:
            want, got = (ctypes.c_uint8 * 32)(), (ctypes.c_uint8 * 32)()
            if name in ("Fr_idiv", "Fr_mod") and b == 0:
                # the wasm traps on a zero divisor; the tape reports the proof as failed (status 4) instead
                assert L.tape_host_rt_op(name.encode(), unary, le(a), le(b), want) == 3
                assert L.tape_host_apply(op, le(a), le(b), le(0), got) == 4
                continue
            assert L.tape_host_rt_op(name.encode(), unary, le(a), le(b), want) == 0, name
            assert L.tape_host_apply(op, le(a), le(b), le(0), got) == 0
            if name == "Fr_idiv":
                assert int.from_bytes(bytes(want), "little") == a // b
            if name == "Fr_mod":
                assert int.from_bytes(bytes(want), "little") == a % b
            if name == "Fr_pow":
                assert int.from_bytes(bytes(want), "little") == pow(a, b, P)
            assert bytes(got) == bytes(want), (name, a, b, int.from_bytes(bytes(got), "little"), int.from_bytes(bytes(want), "little"))


def test_extractor_rejects_what_it_cannot_run(tape, art_dir):
    """not-a-wasm, a truncated wasm and a wasm without its name section are reported with a message (the product turns
    that into ZKB_UNSUPPORTED_CIRCUIT), never a crash or a silent fallback"""
    L, _ = tape
    wasm = open(art_dir + "/circuit.wasm", "rb").read()

    def build(b):
        buf = (ctypes.c_char * len(b)).from_buffer_copy(b)
        info = np.zeros(8, dtype=np.uint32)
        err = ctypes.create_string_buffer(512)
        return L.tape_host_build(ctypes.addressof(buf), len(b), info.ctypes.data, err, 512), err.value.decode()
    rc, msg = build(b"\0asm\x01\0\0\0" + bytes(32))
    assert rc == 1 and msg
    rc, msg = build(wasm[:len(wasm) // 2])
    assert rc == 1 and msg
    cut = wasm.rfind(b"\x04name")                 # the custom "name" section is the last one circom emits
    assert cut > 0
    rc, msg = build(wasm[:cut - 4])               # section header = id 0x00 + 3-byte size before the name string
    assert rc == 1 and "name" in msg
    # the harness still works afterwards
    rc, msg = build(wasm)
    assert rc == 0


# ---- third program: every runtime operation, nested conditionals, an assert (oracle/make_ops_wasm.py) ----------------

@pytest.fixture(scope="module")
def ops_tape(art_dir, tmp_path_factory):
    import make_ops_wasm as OW
    src = os.path.join(H.ROOT, "tests", "host_emul", "tape_host.cc")
    so = str(tmp_path_factory.mktemp("tape_ops") / "libtape_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = ctypes.CDLL(so)
    L.tape_host_build.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    L.tape_host_input.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    L.tape_host_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.tape_host_wasm_witness.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    wasm = OW.build(open(art_dir + "/circuit.wasm", "rb").read())
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    info = np.zeros(8, dtype=np.uint32)
    err = ctypes.create_string_buffer(512)
    assert L.tape_host_build(ctypes.addressof(buf), len(wasm), info.ctypes.data, err, 512) == 0, err.value
    return L, info, wasm, buf


def test_third_program_shape(ops_tape):
    L, info, wasm, _ = ops_tape
    n_inputs, n_wires, n_slots, n_ops, n_levels, n_consts, n_asserts, n_selects = (int(x) for x in info)
    assert (n_inputs, n_wires, n_asserts) == (3, 30, 1)
    assert n_selects >= 3                         # the nested pair + the one-armed conditional
    assert 30 <= n_ops <= 60 and n_levels <= 6
    kat = json.load(open(H.GOLDEN + "/ops_program_kat.json"))
    assert H.sha(wasm) == kat["wasm_sha256"]
    size = ctypes.c_uint32(0)
    assert L.tape_host_input(b"x", ctypes.byref(size)) == 0 and size.value == 3


def test_third_program_all_operations_match_model_and_wasm(ops_tape):
    """OpsTest three ways: (1) the Python model of the circom operators (the committed KAT), (2) the wasm itself run
    concretely by the interpreter - every Fr_* call executes the reference runtime's own code, (3) the extracted
    straight-line program with tape_ops.cuh semantics (what the GPU runs).  Integer division, remainder, power, every
    comparison / shift / bit operation, nested and one-armed conditionals, an assert."""
    L, info, wasm, buf = ops_tape
    kat = json.load(open(H.GOLDEN + "/ops_program_kat.json"))
    n_valid = 0
    for case in kat["cases"]:
        x = [int(v) for v in case["x"]]
        inp = np.zeros((3, 32), dtype=np.uint8)
        for i, v in enumerate(x):
            inp[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
        wt, ww = np.zeros((30, 32), dtype=np.uint8), np.zeros((30, 32), dtype=np.uint8)
        rc_tape = L.tape_host_eval(inp.ctypes.data, wt.ctypes.data)
        rc_wasm = L.tape_host_wasm_witness(ctypes.addressof(buf), len(wasm), inp.ctypes.data, ww.ctypes.data)
        assert rc_tape == case["status"], case["x"]
        if x[1] == 0:
            assert rc_wasm == 3                   # the wasm traps inside Fr_idiv; the tape reports a failed proof
            continue
        assert rc_wasm == case["status"], case["x"]
        if case["status"] == 0:
            want = np.stack([np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8) for v in case["witness"]])
            bad = [k for k in range(30) if not np.array_equal(ww[k], want[k])]
            assert not bad, ("wasm vs model", case["x"], bad)
            bad = [k for k in range(30) if not np.array_equal(wt[k], want[k])]
            assert not bad, ("tape vs model", case["x"], bad)
            n_valid += 1
    assert n_valid >= 20


def test_census_program_run_concretely_matches_the_fixture(tape, art_dir):
    """the interpreter itself, with no symbols: the reference wasm on the reference's inputs_example.json gives the
    committed witness (same sha256 as the native build of the wasm and the GPU kernels)"""
    L, info = tape
    L.tape_host_wasm_witness.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    wasm = open(art_dir + "/circuit.wasm", "rb").read()
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    packed = _pack(L, H.fixture_inputs(), int(info[0]))
    w = np.zeros((int(info[1]), 32), dtype=np.uint8)
    assert L.tape_host_wasm_witness(ctypes.addressof(buf), len(wasm), packed.ctypes.data, w.ctypes.data) == 0
    code, wt = _eval(L, info, H.fixture_inputs())
    assert code == 0 and np.array_equal(w, wt)


def test_signal_used_as_integer_is_rejected_with_the_reason(ops_tape, art_dir):
    """a program whose main component calls Fr_toInt on an input (a loop bound or array index that depends on a signal)
    has no straight-line form: the extractor reports it, the product would return ZKB_UNSUPPORTED_CIRCUIT"""
    import make_ops_wasm as OW
    L = ops_tape[0]
    wasm = OW.build(open(art_dir + "/circuit.wasm", "rb").read(), body="toint")
    buf = (ctypes.c_char * len(wasm)).from_buffer_copy(wasm)
    info = np.zeros(8, dtype=np.uint32)
    err = ctypes.create_string_buffer(512)
    assert L.tape_host_build(ctypes.addressof(buf), len(wasm), info.ctypes.data, err, 512) == 1
    assert "used as an integer" in err.value.decode()


def test_damaged_wasm_through_the_extractor_under_address_sanitizer(art_dir):
    """tools/fuzz_wasm_extractor.py, bounded: 3 child processes x 8 damaged copies of the small test program (truncated,
    bit flips in section headers, random bytes in type / function / element / code / data / name sections) through
    build_witness_program compiled with -fsanitize=address: a message or a program, never a memory error."""
    import importlib.util
    if not os.path.exists(os.path.join(H.ROOT, "artifacts", "opsTest", "dev", "1", "circuit.wasm")):
        pytest.skip("artifacts/opsTest not generated (run __graft_entry__.build())")
    spec = importlib.util.spec_from_file_location("fuzz_wasm_extractor", os.path.join(H.ROOT, "tools", "fuzz_wasm_extractor.py"))
    F = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(F)
    clean, failures = F.run(3, first_seed=132, n_mutants=8, timeout=600)
    assert not failures, failures[0][1][-1500:]
    assert clean == 3
