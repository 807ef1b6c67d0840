"""CPU tests: the Groth16 oracle against the reference's verifier fixtures and its own closed form."""
import json
import os
import numpy as np
import pytest

import helpers as H
import oracle_lib as O


def test_reference_proof_verifies_under_reference_vkey():
    """proof.json + signals.json verify under verification_key.json; tampering is rejected."""
    vk = json.load(open(H.GOLDEN + "/verification_key.json"))
    pf = json.load(open(H.GOLDEN + "/proof.json"))
    pub = json.load(open(H.GOLDEN + "/signals.json"))
    assert O.verify(vk, pub, pf)
    bad = list(pub)
    bad[7] = "6"
    assert not O.verify(vk, bad, pf)
    pf2 = json.loads(json.dumps(pf))
    pf2["pi_a"][0], pf2["pi_c"][0] = pf["pi_c"][0], pf["pi_a"][0]
    assert not O.verify(vk, pub, pf2)


def test_vk_alphabeta_12_reproduced():
    vk = json.load(open(H.GOLDEN + "/verification_key.json"))
    assert O.alphabeta12(vk) == vk["vk_alphabeta_12"]


def test_dev_zkey_shape(art_dir):
    zk = H.zkey_ref()
    assert (zk.n_vars, zk.n_public, zk.domain) == (82754, 8, 131072)
    assert zk.n_coefs == 462889                                  # nnzA + nnzB + 9
    vk = H.dev_vkey()
    assert vk["nPublic"] == 8 and len(vk["IC"]) == 9
    assert O.alphabeta12(vk) == vk["vk_alphabeta_12"]


def test_oracle_prove_verifies_and_matches_closed_form(art_dir):
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    inp = H.fixture_inputs()
    code, w = RW.witness(inp)
    pf = H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED)
    pub = [str(O.from_le(w[i])) for i in range(1, 9)]
    assert O.verify(H.dev_vkey(), pub, O.proof_json(pf))
    r1cs = os.path.join(H.ART, "circuit.r1cs")
    if os.path.exists(r1cs):
        import make_dev_artifacts as MDA
        assert O.check_closed_form(r1cs, MDA.SETUP_SEED, w, H.R_FIXED, H.S_FIXED, pf) == 1
    # a different blinding gives a different, still valid proof
    pf2 = H.zkey_ref().prove(w, 99, 7)
    assert pf2 != pf and O.verify(H.dev_vkey(), pub, O.proof_json(pf2))


def test_oracle_shortcut_prove_equals_dense_prove(art_dir):
    """The "CPU with the same shortcut" leg of bench.py (SURVEY 8a W7 / BASELINE.md section 3): the four witness MSMs
    over (w - template) + cached template sums give the same proof as the dense prover, whatever the template."""
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    _, w = RW.witness(H.fixture_inputs())
    _, t = RW.witness(H.voters(2)[1])
    zk = H.zkey_ref()
    want = zk.prove(w, H.R_FIXED, H.S_FIXED)
    assert zk.prove_shortcut(w, t, H.R_FIXED, H.S_FIXED) == want
    assert zk.prove_shortcut(w, w, H.R_FIXED, H.S_FIXED) == want          # empty difference
    assert zk.prove_shortcut(w, np.zeros_like(w), H.R_FIXED, H.S_FIXED) == want   # empty template


def test_chain600_kat_pins_the_committed_artifacts():
    """tests/golden/chain600_kat.json (what bench.py --gpus N compares the sharded proof with) matches the key and
    witness build() generates into artifacts/chain600, and its proof verifies under that key's vkey."""
    import hashlib
    d = os.path.join(H.ROOT, "artifacts", "chain600")
    if not os.path.exists(os.path.join(d, "proving_key.zkey")):
        pytest.skip("artifacts/chain600 not generated (run __graft_entry__.build())")
    kat = json.load(open(H.GOLDEN + "/chain600_kat.json"))
    assert hashlib.sha256(open(os.path.join(d, "proving_key.zkey"), "rb").read()).hexdigest() == kat["zkey_sha256"]
    assert hashlib.sha256(open(os.path.join(d, "witness.wtns"), "rb").read()).hexdigest() == kat["wtns_sha256"]
    vk = json.load(open(os.path.join(d, "verification_key.json")))
    assert O.verify(vk, kat["public"], kat["proof"])


def test_oracle_ntt_and_msm_self_consistency():
    rng = np.random.default_rng(3)
    v = rng.integers(0, 256, size=(1 << 10, 32), dtype=np.uint8)
    v[:, 31] &= 0x0F
    assert np.array_equal(O.ntt(O.ntt(v), inverse=True), v)
    # NTT of delta_1 is the geometric sequence of w: spot check linearity against a direct evaluation
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    w = pow(5, (R - 1) >> 10, R)
    d = np.zeros((1 << 10, 32), dtype=np.uint8)
    d[1, 0] = 1
    out = O.ntt(d)
    assert [O.from_le(out[i]) for i in range(4)] == [pow(w, i, R) for i in range(4)]
    # MSM linearity: msm(P, a) + msm(P, b) == msm(P, a+b) checked through k*G
    import ctypes
    pts = np.zeros((8, 64), dtype=np.uint8)
    for i in range(8):
        buf = (ctypes.c_uint8 * 64)()
        O.lib().orc_g1_mul_gen(O._buf(O.le32(i + 2)), buf)
        pts[i] = np.frombuffer(bytes(buf), dtype=np.uint8)
    sc = [int(x) for x in rng.integers(1, 1 << 60, size=8)]
    s = np.frombuffer(b"".join(O.le32(x) for x in sc), dtype=np.uint8).reshape(8, 32)
    total = sum((i + 2) * k for i, k in enumerate(sc)) % R
    buf = (ctypes.c_uint8 * 64)()
    O.lib().orc_g1_mul_gen(O._buf(O.le32(total)), buf)
    assert O.msm_g1(pts, s) == bytes(buf)


def test_chain_circuit_golden_vector(tmp_path):
    """The oracle's chain-circuit generator, dev setup and prover reproduce the committed known answer."""
    import hashlib
    kat = json.load(open(H.GOLDEN + "/chain40_kat.json"))
    n_wires, n_cons, domain = O.chain_artifacts(kat["links"], kat["setup_seed"], str(tmp_path))
    assert (n_wires, n_cons, domain) == (kat["n_wires"], kat["n_constraints"], kat["domain"])
    zkey = open(tmp_path / "proving_key.zkey", "rb").read()
    wtns = open(tmp_path / "witness.wtns", "rb").read()
    assert hashlib.sha256(zkey).hexdigest() == kat["zkey_sha256"]
    assert hashlib.sha256(wtns).hexdigest() == kat["wtns_sha256"]
    w = H.wtns_payload(wtns, n_wires)
    proof = O.proof_json(O.ZKeyRef(zkey).prove(w, kat["r"], kat["s"]))
    assert proof == kat["proof"]
    assert O.verify(json.load(open(tmp_path / "verification_key.json")), kat["public"], proof)
