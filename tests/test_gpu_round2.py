"""GPU parity tests added in round 2: the full 1,024-proof configuration, the proof-independent-wire shortcut
switched off (ZKB_LOAD_DENSE), censuses of a chosen depth, the binary batch verifier, the hardened .wtns parser and the
later rapidsnark prover.h entry points.  Everything goes through the C ABI (ctypes)."""
import ctypes
import json
import os
import struct

import numpy as np
import pytest

import helpers as H
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def files(art_dir):
    return (open(art_dir + "/proving_key.zkey", "rb").read(), open(art_dir + "/circuit.wasm", "rb").read(),
            open(art_dir + "/verification_key.json", "rb").read())


@pytest.fixture(scope="module")
def circuit(files):
    from zk_franchise_proof_circuit_b200 import prover
    return prover.load(files[0], files[1])


def _ref_witness(inputs):
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    return RW.witness(inputs)


def test_config2_all_1024_proofs_verify_and_sample_matches_oracle(circuit, files):
    """BASELINE configs[1] at its full size: 1,024 proofs with r = 1234567, s = 7654321; ALL verified (GPU verifier and,
    for a slice, the CPU oracle's verifier), 6 spread over the four pipeline lanes compared limb for limb with the
    CPU oracle."""
    from zk_franchise_proof_circuit_b200 import prover
    vs = H.voters(1024)
    packed = np.stack([prover.pack_inputs(v) for v in vs])
    circuit.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        circuit.set_inputs(packed)
        circuit.prove_resident()
    finally:
        circuit.set_blinding(None, None)
    proofs, pubs, status = circuit.get_results()
    assert (status == 0).all()
    assert prover.verify_batch_bin(files[2], pubs, proofs).sum() == 1024
    assert O.verify_many(H.dev_vkey(), pubs[:32].tobytes(), proofs[:32].tobytes(), 32).all()
    for i in (0, 127, 128, 511, 640, 1023):
        code, w = _ref_witness(vs[i])
        assert code == 0
        assert proofs[i].tobytes() == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED), f"voter {i}"
    # tampering is caught by the binary verifier too
    bad = proofs[:4].copy()
    bad[1, 0] ^= 1
    bad[2, 255] = 0xFF                                   # coordinate >= q
    assert list(prover.verify_batch_bin(files[2], pubs[:4], bad)) == [1, 0, 0, 1]


def test_dense_circuit_matches_oracle(files):
    """ZKB_LOAD_DENSE: no template, all 161 levels hashed, A/B1/B2/C over all wires - same witness, same proof."""
    from zk_franchise_proof_circuit_b200 import prover
    c = prover.load(files[0], files[1], dense=True)
    vs = [H.fixture_inputs()] + list(H.voters(3)) + [H.deep_voters()[0]]
    c.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        proofs, pubs, status = c.fullprove_batch([json.dumps(v) for v in vs])
    finally:
        c.set_blinding(None, None)
    assert status == [0] * len(vs)
    for i, v in enumerate(vs):
        code, w = _ref_witness(v)
        assert np.array_equal(c.get_witness(i, 1)[0], w), f"witness {i}"
        assert O.proof_bin(json.loads(proofs[i])) == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED), f"proof {i}"
    c.close()
    prover._circuits.clear()


@pytest.mark.parametrize("depth", [1, 4, 40, 160])
def test_census_of_chosen_depth(circuit, files, depth):
    """census_tree.gen_census_depth (bench.py's depth sweep): inputs pass the reference wasm, the GPU witness is
    bit-exact, proofs == oracle and verify."""
    from zk_franchise_proof_circuit_b200 import prover, census_tree
    vs = census_tree.gen_census_depth(circuit, 3, depth, seed=11)
    assert all(census_tree.tree_depth(v) == depth for v in vs)
    circuit.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        proofs, pubs, status = circuit.fullprove_batch([json.dumps(v) for v in vs])
    finally:
        circuit.set_blinding(None, None)
    assert status == [0, 0, 0]
    assert prover.verify_batch(files[2], pubs, proofs) == [1, 1, 1]
    code, w = _ref_witness(vs[1])
    assert code == 0
    assert np.array_equal(circuit.get_witness(1, 1)[0], w)
    assert O.proof_bin(json.loads(proofs[1])) == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED)


def _wtns_bytes(w, n8=32, prime=None, data_len=None):
    prime = prime if prime is not None else census_r()
    body = w.tobytes()
    if data_len is not None:
        body = body[:data_len]
    return (b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, 40) + struct.pack("<I", n8) +
            prime.to_bytes(32, "little") + struct.pack("<I", w.shape[0]) + struct.pack("<IQ", 2, len(body)) + body)


def census_r():
    return 21888242871839275222246405745257275088548364400416034343698204186575808495617


def test_wtns_parser_rejects_malformed(circuit):
    """short data sections, wrong field size / prime and non-canonical values are errors, not reads past
    the caller's buffer."""
    from zk_franchise_proof_circuit_b200.prover import NativeError
    code, w = _ref_witness(H.fixture_inputs())
    good = _wtns_bytes(w)
    pj, sj = circuit.prove_wtns(good)
    assert json.loads(sj) == json.load(open(H.GOLDEN + "/signals.json"))
    for bad in (_wtns_bytes(w, data_len=w.size - 32), _wtns_bytes(w, n8=64), _wtns_bytes(w, prime=census_r() + 2)):
        with pytest.raises(NativeError):
            circuit.prove_wtns(bad)
    w2 = w.copy()
    w2[5] = np.frombuffer((census_r() + 1).to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(NativeError):
        circuit.prove_wtns(_wtns_bytes(w2))
    with pytest.raises(NativeError):
        circuit.prove_wtns(good[:200])


def test_rapidsnark_later_entry_points(files, tmp_path):
    """groth16_public_size_for_zkey_buf / _file, groth16_proof_size, groth16_prover_zkey_file: same answers as
    groth16_prover; the key cache tells two keys of equal size apart."""
    from zk_franchise_proof_circuit_b200 import _native, prover
    L = _native.lib()
    ul, vp = ctypes.c_ulong, ctypes.c_void_p
    L.groth16_proof_size.argtypes = [ctypes.POINTER(ul)]
    L.groth16_proof_size.restype = None
    L.groth16_public_size_for_zkey_buf.argtypes = [vp, ul, ctypes.POINTER(ul), vp, ul]
    L.groth16_public_size_for_zkey_file.argtypes = [ctypes.c_char_p, ctypes.POINTER(ul), vp, ul]
    L.groth16_prover_zkey_file.argtypes = [ctypes.c_char_p, vp, ul, vp, ctypes.POINTER(ul), vp, ctypes.POINTER(ul), vp, ul]
    prover._lib()
    zkey = files[0]
    zb = (ctypes.c_char * len(zkey)).from_buffer_copy(zkey)
    n = ul(0)
    L.groth16_proof_size(ctypes.byref(n))
    assert n.value >= 700
    err = ctypes.create_string_buffer(256)
    assert L.groth16_public_size_for_zkey_buf(ctypes.addressof(zb), len(zkey), ctypes.byref(n), err, 256) == 0
    assert n.value >= 8 * 80
    assert L.groth16_public_size_for_zkey_buf(ctypes.addressof(zb), 100, ctypes.byref(n), err, 256) == 1
    zpath = os.path.join(H.ART, "proving_key.zkey").encode()
    assert L.groth16_public_size_for_zkey_file(zpath, ctypes.byref(n), err, 256) == 0
    code, w = _ref_witness(H.fixture_inputs())
    wt = _wtns_bytes(w)
    wb = (ctypes.c_char * len(wt)).from_buffer_copy(wt)
    pbuf, qbuf = ctypes.create_string_buffer(2048), ctypes.create_string_buffer(2048)
    pn, qn = ul(2048), ul(2048)
    assert L.groth16_prover_zkey_file(zpath, ctypes.addressof(wb), len(wt), pbuf, ctypes.byref(pn), qbuf, ctypes.byref(qn), err, 256) == 0, err.value
    assert json.loads(qbuf.raw[:qn.value]) == json.load(open(H.GOLDEN + "/signals.json"))
    assert O.verify(H.dev_vkey(), json.loads(qbuf.raw[:qn.value]), json.loads(pbuf.raw[:pn.value]))
    # a second key of the SAME size that differs only deep inside a point section must not reuse the cached tables:
    # flip one bit in section 9 (H points) -> the proof of the same witness changes (and no longer verifies)
    z2 = bytearray(zkey)
    pos, off9 = 12, None
    for _ in range(struct.unpack_from("<I", zkey, 8)[0]):
        sid, sz = struct.unpack_from("<IQ", zkey, pos)
        if sid == 9:
            off9 = pos + 12
        pos += 12 + sz
    z2[off9 + 64 * 777 + 5] ^= 1
    z2b = (ctypes.c_char * len(z2)).from_buffer_copy(bytes(z2))
    outs = []
    for key in (zb, z2b, zb):
        buf = ctypes.create_string_buffer(2048)
        pn, qn = ul(2048), ul(2048)
        assert L.groth16_prover(ctypes.addressof(key), len(zkey), ctypes.addressof(wb), len(wt), buf, ctypes.byref(pn), qbuf,
                                ctypes.byref(qn), err, 256) == 0, err.value
        outs.append(json.loads(buf.raw[:pn.value]))
    pub = json.load(open(H.GOLDEN + "/signals.json"))
    assert O.verify(H.dev_vkey(), pub, outs[0]) and O.verify(H.dev_vkey(), pub, outs[2])
    assert not O.verify(H.dev_vkey(), pub, outs[1])


def test_verify_runs_on_the_current_device(files):
    """the verifier caches its device objects per GPU (no cross-device pointers)."""
    from zk_franchise_proof_circuit_b200 import prover, _native
    pf = open(H.GOLDEN + "/proof.json", "rb").read()
    pub = open(H.GOLDEN + "/signals.json", "rb").read()
    vk = open(H.GOLDEN + "/verification_key.json", "rb").read()
    prover.verify(vk, pub, pf)
    if _native.lib().zkb_device_count() >= 2:
        cuda = ctypes.CDLL("libcudart.so.12")
        cuda.cudaSetDevice(1)
        try:
            prover.verify(vk, pub, pf)
        finally:
            cuda.cudaSetDevice(0)
    prover.verify(vk, pub, pf)


# ---- generic witness path -------------------------------------------------------------------------

@pytest.fixture(scope="module")
def generic_circuit(files):
    """the census key with ZKB_LOAD_GENERIC_WITNESS: the witness program is extracted from circuit.wasm by symbolic
    execution at load time; nothing census-specific runs"""
    from zk_franchise_proof_circuit_b200 import prover
    return prover.load(files[0], files[1], generic=True)


def test_generic_witness_gpu_bit_exact(generic_circuit):
    c = generic_circuit
    assert (c.n_vars, c.n_inputs) == (82754, 334)
    inp = H.fixture_inputs()
    w = H.wtns_payload(c.witness(json.dumps(inp)), c.n_vars)
    assert H.sha(w.tobytes()) == H.WITNESS_SHA256                      # golden KAT of the reference wasm
    for v in list(H.voters(2)) + [H.deep_voters()[0]]:
        code, ref = _ref_witness(v)
        assert code == 0
        assert np.array_equal(H.wtns_payload(c.witness(json.dumps(v)), c.n_vars), ref)


def test_generic_witness_proofs_match_oracle_and_asserts(generic_circuit, files):
    from zk_franchise_proof_circuit_b200 import prover
    c = generic_circuit
    vs = list(H.voters(40))
    bad = dict(vs[3], voteWeight="11")                                   # fails LessEqThan: the wasm raises exception 4
    docs = [json.dumps(v) for v in vs[:3]] + [json.dumps(bad)] + [json.dumps(v) for v in vs[4:]]
    c.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        proofs, pubs, status = c.fullprove_batch(docs)
    finally:
        c.set_blinding(None, None)
    assert status == [0, 0, 0, 4] + [0] * 36
    ok = prover.verify_batch(files[2], [q for i, q in enumerate(pubs) if i != 3], [p for i, p in enumerate(proofs) if i != 3])
    assert ok == [1] * 39
    for i in (0, 17, 39):
        code, w = _ref_witness(vs[i])
        assert O.proof_bin(json.loads(proofs[i])) == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED), f"voter {i}"
    # unknown / missing input names are errors of that document only
    from zk_franchise_proof_circuit_b200.prover import NativeError
    with pytest.raises(NativeError):
        c.fullprove(json.dumps(dict(vs[0], extra="1")))
    miss = dict(vs[0])
    del miss["password"]
    with pytest.raises(NativeError):
        c.fullprove(json.dumps(miss))


def test_non_circom_wasm_is_reported_not_crashed(files):
    """a wasm that is not a circom witness calculator: ZKB_UNSUPPORTED_CIRCUIT with a message, no crash"""
    from zk_franchise_proof_circuit_b200 import prover
    from zk_franchise_proof_circuit_b200.prover import NativeError, UNSUPPORTED_CIRCUIT
    junk = b"\0asm\x01\0\0\0" + bytes(64)
    with pytest.raises(NativeError) as ei:
        prover.Circuit(prover._context(None), files[0], junk)
    assert ei.value.code in (UNSUPPORTED_CIRCUIT, 1)
    trunc = files[1][:200000]
    with pytest.raises(NativeError):
        prover.Circuit(prover._context(None), files[0], trunc)


def test_rlc_batch_verification(circuit, files):
    """n >= 8 proofs are checked with one random-linear-combination pairing product; a batch with an
    invalid proof falls back to per-proof checks and names exactly the bad ones."""
    from zk_franchise_proof_circuit_b200 import prover
    vs = H.voters(24)
    packed = np.stack([prover.pack_inputs(v) for v in vs])
    circuit.set_inputs(packed)
    circuit.prove_resident()
    proofs, pubs, status = circuit.get_results()
    assert (status == 0).all()
    assert list(prover.verify_batch_bin(files[2], pubs, proofs)) == [1] * 24
    bad = proofs.copy()
    bad[5, 64:128] = proofs[6, 64:128]                   # B.x of another proof: a valid curve point? no - x without its y
    bad[17] = proofs[16]                                 # a valid proof of ANOTHER statement
    pb = pubs.copy()
    pb[20, 7, 0] ^= 1                                    # voteWeight changed
    got = list(prover.verify_batch_bin(files[2], pb, bad))
    want = [1] * 24
    want[5] = want[17] = want[20] = 0
    assert got == want
    assert O.verify_many(H.dev_vkey(), pubs[:8].tobytes(), proofs[:8].tobytes(), 8).all()


def test_pair_tree_prover_path_matches_oracle(files, monkeypatch):
    """ZKB_AFFINE=1: the H MSM of a batch runs through the batched-affine pair tree (opt-in):
    same proofs, bit for bit, and the work counters show where the additions went."""
    from zk_franchise_proof_circuit_b200 import prover
    monkeypatch.setenv("ZKB_AFFINE", "1")
    c = prover.Circuit(prover._context(None), files[0], files[1])
    try:
        vs = H.voters(40)
        c.set_blinding(H.R_FIXED, H.S_FIXED)
        c.set_inputs(np.stack([prover.pack_inputs(v) for v in vs]))
        c.prove_resident()
        proofs, pubs, status = c.get_results()
        assert (status == 0).all()
        wc = c.work_counters()
        assert wc["g1_affine_adds_per_proof"] > 1.5e6 and wc["g1_inversions_per_proof"] > 0
        assert prover.verify_batch_bin(files[2], pubs, proofs).sum() == 40
        for i in (0, 21, 39):
            code, w = _ref_witness(vs[i])
            assert proofs[i].tobytes() == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED), f"voter {i}"
    finally:
        c.close()


def test_second_circuit_proves_end_to_end():
    """A wasm that is NOT census.circom (zk_census_test.go:27-34 takes other circuit names): the
    SMTVerifier(160) program of oracle/make_smt_wasm.py with its own dev key, in the reference's on-disk layout
    (artifacts/smtVerifier/dev/160).  zkb_load_circuit picks the generic witness path by itself; the witness equals the
    wasm run natively (committed KAT + oracle/_ref/smt), the proof equals the CPU oracle's and verifies under the key
    zkb_export_vkey writes; a wrong root is exception 4."""
    import make_smt_wasm as SW
    import ref_witness as RW
    from zk_franchise_proof_circuit_b200 import prover
    from zk_franchise_proof_circuit_b200.prover import NativeError
    d = os.path.join(H.ROOT, "artifacts", "smtVerifier", "dev", "160")
    if not os.path.exists(d + "/proving_key.zkey"):
        pytest.skip("artifacts/smtVerifier not generated (run __graft_entry__.build())")
    zkey, wasm = open(d + "/proving_key.zkey", "rb").read(), open(d + "/circuit.wasm", "rb").read()
    c = prover.load(zkey, wasm)                       # no flag: the wasm is simply not the census program
    try:
        assert (c.n_vars, c.n_public, c.n_inputs) == (82754, 2, 169)
        kat = json.load(open(H.GOLDEN + "/smt_verifier_kat.json"))
        voters = [H.fixture_inputs()] + list(H.voters(2)) + [H.deep_voters()[0]]
        docs = [json.dumps(SW.smt_inputs(v)) for v in voters]
        for i, doc in enumerate(docs):
            w = H.wtns_payload(c.witness(doc), c.n_vars)
            assert H.sha(w.tobytes()) == kat["witness_sha256"][i], f"case {i}"
        c.set_blinding(H.R_FIXED, H.S_FIXED)
        proofs, pubs, status = c.fullprove_batch(docs)
        c.set_blinding(None, None)
        assert status == [0] * 4
        vkey = prover.export_vkey(zkey)
        assert vkey == open(d + "/verification_key.json", "rb").read()
        assert prover.verify_batch(vkey, pubs, proofs) == [1] * 4
        assert json.loads(pubs[0]) == ["1", voters[0]["censusRoot"]]
        ref = RW.RefWasm()
        if ref.available():
            zk = O.ZKeyRef(zkey)
            for i in (0, 3):
                rc, w = ref.witness(SW.smt_inputs(voters[i]))
                assert rc == 0 and O.proof_bin(json.loads(proofs[i])) == zk.prove(w, H.R_FIXED, H.S_FIXED), f"proof {i}"
        bad = SW.smt_inputs(voters[0])
        bad["root"] = "7"
        with pytest.raises(NativeError) as ei:
            c.fullprove(json.dumps(bad))
        assert ei.value.code == 4
    finally:
        c.close()
        prover._circuits.clear()


@pytest.mark.gpu
def test_third_circuit_every_runtime_operation_on_the_gpu():
    """A SMALL non-census circuit (30 wires, artifacts/opsTest/dev/1, oracle/make_ops_wasm.py): integer division,
    remainder, power, all comparisons / shifts / bit operations, nested and one-armed conditionals and an assert run
    as an extracted program on the GPU.  Witness == the committed Python-model KAT (which the wasm, run concretely,
    also gives: tests/test_witness_program.py); proofs == CPU oracle, verify; a tampered public output does not;
    a failed assert and a zero divisor are status 4."""
    from zk_franchise_proof_circuit_b200 import prover
    d = os.path.join(H.ROOT, "artifacts", "opsTest", "dev", "1")
    if not os.path.exists(d + "/proving_key.zkey"):
        pytest.skip("artifacts/opsTest not generated (run __graft_entry__.build())")
    zkey, wasm = open(d + "/proving_key.zkey", "rb").read(), open(d + "/circuit.wasm", "rb").read()
    kat = json.load(open(H.GOLDEN + "/ops_program_kat.json"))
    assert H.sha(wasm) == kat["wasm_sha256"]
    c = prover.load(zkey, wasm)
    try:
        assert (c.n_vars, c.n_public, c.n_inputs) == (30, 26, 3)
        good = [k for k in kat["cases"] if k["status"] == 0]
        bad = [k for k in kat["cases"] if k["status"] != 0]
        docs = [json.dumps({"x": k["x"]}) for k in good]
        for k, doc in zip(good, docs):
            w = H.wtns_payload(c.witness(doc), c.n_vars)
            got = [int.from_bytes(w[i].tobytes(), "little") for i in range(c.n_vars)]
            assert got == [int(v) for v in k["witness"]], k["x"]
        c.set_blinding(H.R_FIXED, H.S_FIXED)
        proofs, pubs, status = c.fullprove_batch(docs + [json.dumps({"x": k["x"]}) for k in bad])
        c.set_blinding(None, None)
        assert status == [0] * len(good) + [4] * len(bad)
        vkey = open(d + "/verification_key.json", "rb").read()
        assert prover.export_vkey(zkey) == vkey
        assert prover.verify_batch(vkey, pubs[:len(good)], proofs[:len(good)]) == [1] * len(good)
        zk = O.ZKeyRef(zkey)
        for i in (0, 5, len(good) - 1):
            assert json.loads(pubs[i]) == good[i]["witness"][1:27]
            w = np.stack([np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8) for v in good[i]["witness"]])
            assert O.proof_bin(json.loads(proofs[i])) == zk.prove(w, H.R_FIXED, H.S_FIXED), f"proof {i}"
        pj, sj = c.fullprove(docs[1])                      # 26 public signals: public.json is > 2 KB
        assert json.loads(sj) == good[1]["witness"][1:27] and prover.verify_batch(vkey, [sj], [pj]) == [1]
        tampered = json.loads(pubs[0])
        tampered[25] = str(int(tampered[25]) + 1)
        assert prover.verify_batch(vkey, [json.dumps(tampered)], proofs[:1]) == [0]
    finally:
        c.close()
        prover._circuits.clear()
