"""GPU parity tests of the census proving path through the C ABI / the reference-shaped host API.

Bar: the witness is bit-exact against the reference's circuit.wasm (oracle/_ref) and, with r,s pinned,
pi_a / pi_b / pi_c are bit-exact against the CPU oracle's restatement of snarkjs groth16.prove; every proof
verifies under the vkey of the dev setup."""
import json
import numpy as np
import pytest

import helpers as H
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def circuit(art_dir):
    from zk_franchise_proof_circuit_b200 import prover
    zkey = open(art_dir + "/proving_key.zkey", "rb").read()
    wasm = open(art_dir + "/circuit.wasm", "rb").read()
    c = prover.load(zkey, wasm)
    assert (c.n_vars, c.n_public, c.domain, c.n_inputs, c.n_levels1) == (82754, 8, 131072, 334, 161)
    return c


def _ref_witness(inputs):
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built")
    return RW.witness(inputs)


def test_witness_fixture_bit_exact(circuit):
    inp = H.fixture_inputs()
    wtns = circuit.witness(json.dumps(inp))
    w = H.wtns_payload(wtns, circuit.n_vars)
    assert H.sha(w.tobytes()) == H.WITNESS_SHA256                      # golden KAT (reference wasm)
    kat = json.load(open(H.GOLDEN + "/witness_kat.json"))
    for i, v in kat["wires"].items():
        assert O.from_le(w[int(i)]) == int(v)
    code, ref = _ref_witness(inp)
    assert code == 0 and np.array_equal(ref, w)
    # public signals == signals.json
    pub = json.load(open(H.GOLDEN + "/signals.json"))
    assert [str(O.from_le(w[i])) for i in range(1, 9)] == pub


def test_witness_synthetic_voters_bit_exact(circuit):
    from zk_franchise_proof_circuit_b200 import prover
    vs = H.voters(64)
    packed = np.stack([prover.pack_inputs(v) for v in vs])
    circuit.set_inputs(packed)
    circuit.prove_resident()
    _, _, status = circuit.get_results()
    assert (status == 0).all()
    for i in (0, 1, 31, 63):
        code, ref = _ref_witness(vs[i])
        assert code == 0
        assert np.array_equal(circuit.get_witness(i, 1)[0], ref), f"voter {i}"


def test_assert_failures_match_wasm(circuit):
    inp = H.fixture_inputs()
    for patch in ({"voteWeight": "11"}, {"nullifier": "5"}, {"censusRoot": "7"}):
        bad = dict(inp)
        bad.update(patch)
        proofs, pubs, status = circuit.fullprove_batch([json.dumps(bad), json.dumps(inp)])
        assert status == [4, 0], patch                                  # exceptionHandler(4); batch continues
        assert proofs[0] == b"" and proofs[1] != b""
    from zk_franchise_proof_circuit_b200.prover import NativeError
    with pytest.raises(NativeError) as ei:
        circuit.fullprove(json.dumps(dict(inp, voteWeight="11")))
    assert ei.value.code == 4
    missing = dict(inp)
    del missing["address"]
    with pytest.raises(NativeError):
        circuit.fullprove(json.dumps(missing))


def test_proof_fixture_bit_exact_with_pinned_blinding(circuit):
    inp = H.fixture_inputs()
    circuit.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        pj, sj = circuit.fullprove(json.dumps(inp))
    finally:
        circuit.set_blinding(None, None)
    proof, pub = json.loads(pj), json.loads(sj)
    assert set(proof) == {"pi_a", "pi_b", "pi_c"}
    assert pub == json.load(open(H.GOLDEN + "/signals.json"))
    code, w = _ref_witness(inp)
    exp, exp_part = H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED, partials=True)
    # stage-by-stage first, so that a mismatch names the stage: H scalars, then the five MSM partial sums
    part, h = circuit.debug_partials(with_h=True)
    assert np.array_equal(h, H.zkey_ref().h_scalars(w)), "H scalars (buildABC / NTT / join)"
    for name, lo, hi in (("pi_a'", 0, 64), ("pi_b1'", 64, 128), ("pi_b'", 128, 256), ("pi_c'", 256, 320), ("pi_h", 320, 384)):
        assert part[lo:hi] == exp_part[lo:hi], f"MSM partial {name}"
    got = O.proof_bin(proof)
    for name, lo, hi in (("pi_a", 0, 64), ("pi_b", 64, 192), ("pi_c", 192, 256)):
        assert got[lo:hi] == exp[lo:hi], f"proof element {name}"         # A, B, C limb for limb
    assert O.verify(H.dev_vkey(), pub, proof)


def test_batch_proofs_verify_and_match_oracle(circuit):
    vs = H.voters(64)
    docs = [json.dumps(v) for v in vs]
    circuit.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        proofs, pubs, status = circuit.fullprove_batch(docs)
    finally:
        circuit.set_blinding(None, None)
    assert status == [0] * 64
    vk = H.dev_vkey()
    pb = b"".join(O.proof_bin(json.loads(p)) for p in proofs)
    qb = b"".join(O.pub_bin(json.loads(q)) for q in pubs)
    ok = O.verify_many(vk, qb, pb, 64)
    assert ok.all()
    for i in (0, 17, 40, 63):                                            # bit-exact vs the CPU oracle
        code, w = _ref_witness(vs[i])
        assert pb[i * 256:(i + 1) * 256] == H.zkey_ref().prove(w, H.R_FIXED, H.S_FIXED), f"voter {i}"


def test_random_blinding_differs_but_verifies(circuit):
    inp = json.dumps(H.fixture_inputs())
    p1, s1 = circuit.fullprove(inp)
    p2, s2 = circuit.fullprove(inp)
    assert p1 != p2 and s1 == s2
    vk = H.dev_vkey()
    assert O.verify(vk, json.loads(s1), json.loads(p1)) and O.verify(vk, json.loads(s2), json.loads(p2))


def test_go_and_snarkjs_shaped_api(art_dir):
    from zk_franchise_proof_circuit_b200 import prover
    zkey = open(art_dir + "/proving_key.zkey", "rb").read()
    wasm = open(art_dir + "/circuit.wasm", "rb").read()
    inputs = open(H.GOLDEN + "/inputs_example.json", "rb").read()
    proof = prover.prove(zkey, wasm, inputs)                             # zk_census_test.go:89
    pdata, psig = proof.bytes()                                          # :93
    again = prover.parse_proof(pdata, psig)                              # :118
    assert O.verify(H.dev_vkey(), again.pub_signals, again.data)         # :122 (oracle verifier)
    assert b" " not in pdata and pdata.startswith(b'{"pi_a":["')
    res = prover.groth16.full_prove(json.loads(inputs), art_dir + "/circuit.wasm", art_dir + "/proving_key.zkey")
    assert res["proof"]["protocol"] == "groth16" and res["proof"]["curve"] == "bn128"
    assert res["publicSignals"] == json.load(open(H.GOLDEN + "/signals.json"))
    assert O.verify(H.dev_vkey(), res["publicSignals"], res["proof"])


def test_rapidsnark_abi_groth16_prover(circuit, art_dir):
    """groth16_prover(zkey, wtns, ...) - the symbol go-rapidsnark binds - with a .wtns produced by zkb_witness."""
    import ctypes
    from zk_franchise_proof_circuit_b200 import prover
    L = prover._lib()
    zkey = open(art_dir + "/proving_key.zkey", "rb").read()
    wtns = circuit.witness(json.dumps(H.fixture_inputs()))
    pbuf, qbuf, ebuf = ctypes.create_string_buffer(16), ctypes.create_string_buffer(2048), ctypes.create_string_buffer(256)
    pn, qn = ctypes.c_ulong(16), ctypes.c_ulong(2048)
    zb = (ctypes.c_char * len(zkey)).from_buffer_copy(zkey)
    wb = (ctypes.c_char * len(wtns)).from_buffer_copy(wtns)
    rc = L.groth16_prover(zb, len(zkey), wb, len(wtns), pbuf, ctypes.byref(pn), qbuf, ctypes.byref(qn), ebuf, 256)
    assert rc == 2 and pn.value > 16                                     # PROVER_ERROR_SHORT_BUFFER, size reported
    pbuf = ctypes.create_string_buffer(pn.value + 64)                    # random blinding: digits may differ
    pn, qn = ctypes.c_ulong(len(pbuf)), ctypes.c_ulong(2048)             # sizes are in/out, as in rapidsnark
    rc = L.groth16_prover(zb, len(zkey), wb, len(wtns), pbuf, ctypes.byref(pn), qbuf, ctypes.byref(qn), ebuf, 256)
    assert rc == 0, ebuf.value
    proof, pub = json.loads(pbuf.raw[:pn.value]), json.loads(qbuf.raw[:qn.value])
    assert O.verify(H.dev_vkey(), pub, proof)
    pn, qn = ctypes.c_ulong(len(pbuf)), ctypes.c_ulong(2048)
    rc = L.groth16_prover(zb, len(zkey), wb, len(wtns) - 32, pbuf, ctypes.byref(pn), qbuf, ctypes.byref(qn), ebuf, 256)
    assert rc == 1


def test_unsupported_inputs_are_errors_not_fallbacks(art_dir):
    from zk_franchise_proof_circuit_b200 import prover
    zkey = open(art_dir + "/proving_key.zkey", "rb").read()
    with pytest.raises(prover.NativeError):
        prover.Circuit(prover._context(), zkey[:1000], None)
    with pytest.raises(prover.NativeError) as ei:
        prover.Circuit(prover._context(), zkey, b"\0asm\x01\0\0\0")
    assert ei.value.code == prover.UNSUPPORTED_CIRCUIT


def test_gpu_census_generator_matches_oracle(circuit):
    """census_tree.gen_census (GPU Poseidon, one launch per tree level) == oracle/census_gen.py, field for field."""
    from zk_franchise_proof_circuit_b200 import census_tree
    import census_gen as G
    P = G.Poseidon(H.poseidon_tables())
    rows = [(1, 2), (0, 0), (2**200, 5)]
    assert circuit.poseidon(rows) == [P(list(r)) for r in rows]
    assert circuit.poseidon([(1, 2, 3), (0, 0, 1)]) == [P([1, 2, 3]), P([0, 0, 1])]
    assert circuit.poseidon([(1, 2, 3, 4)]) == [P([1, 2, 3, 4])]
    got = census_tree.gen_census(circuit, 64, seed=0xC0FFEE)
    assert got == H.voters(64)


def test_stage_timing_and_work_counters(circuit):
    from zk_franchise_proof_circuit_b200 import prover
    vs = H.voters(64)
    packed = np.stack([prover.pack_inputs(v) for v in vs[:32]])
    circuit.set_inputs(packed)
    n0 = prover.launch_count()
    st = circuit.prove_resident(32, stages=True)
    assert prover.launch_count() > n0
    assert st.shape == (8,) and (st > 0).all()
    wc = circuit.work_counters()
    # H is dense (131,072 full-width scalars x 16 windows, minus zero digits and one free add per bucket); the four
    # witness MSMs only see the wires that differ from the per-key template, far below 16 x 82,754
    # (the H MSM's additions are split between the batched-affine pair tree and the XYZZ tail)
    adds = wc["g1_madds_per_proof"] + wc["g1_affine_adds_per_proof"]
    assert 16 * 131072 * 0.9 < adds < 16 * 131072 + 3 * 16 * 82754 * 0.25
    assert wc["g1_affine_adds_per_proof"] == 0 or wc["g1_inversions_per_proof"] > 0
    assert 0 < wc["g2_madds_per_proof"] < 16 * 82754 * 0.25
    print("stage ms for 32 proofs:", [round(float(x), 2) for x in st], wc)


def test_gpu_verifier_reference_fixture_and_own_proofs(circuit):
    """zkb_verify: the reference's committed proof verifies under its committed vkey; tampering and
    out-of-range encodings are rejected; our own proofs verify under the dev vkey and agree with the oracle verifier."""
    from zk_franchise_proof_circuit_b200 import prover
    vk = open(H.GOLDEN + "/verification_key.json", "rb").read()
    pf = open(H.GOLDEN + "/proof.json", "rb").read()
    pub = open(H.GOLDEN + "/signals.json", "rb").read()
    proof = prover.parse_proof(pf, pub)                                  # zk_census_test.go:118
    proof.verify(vk)                                                     # :122 - no exception
    assert prover.groth16.verify(vk, json.loads(pub), json.loads(pf)) is True
    sig = json.loads(pub)
    bad = list(sig)
    bad[7] = "6"
    assert prover.groth16.verify(vk, bad, json.loads(pf)) is False
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    aliased = list(sig)
    aliased[7] = str(int(sig[7]) + R)                                    # same residue, must still be rejected
    assert prover.groth16.verify(vk, aliased, json.loads(pf)) is False
    swapped = json.loads(pf)
    swapped["pi_a"], swapped["pi_c"] = swapped["pi_c"], swapped["pi_a"]
    assert prover.groth16.verify(vk, sig, swapped) is False
    with pytest.raises(prover.NativeError) as ei:
        prover.parse_proof(json.dumps(swapped).encode(), pub).verify(vk)
    assert ei.value.code == prover.INVALID_PROOF
    # own proofs, batch
    vs = H.voters(64)[:16]
    proofs, pubs, status = circuit.fullprove_batch([json.dumps(v) for v in vs])
    dev_vk = json.dumps(H.dev_vkey()).encode()
    ok = prover.verify_batch(dev_vk, pubs, proofs)
    assert ok == [1] * 16
    assert prover.verify_batch(dev_vk, pubs[:2], [proofs[1], proofs[0]]) == [0, 0]
    assert prover.verify_batch(vk, pubs[:1], proofs[:1]) == [0]          # wrong key
    for p, q in zip(proofs[:4], pubs[:4]):
        assert O.verify(H.dev_vkey(), json.loads(q), json.loads(p))


@pytest.mark.parametrize("links", [40, 600])
def test_generic_circuit_chain_bit_exact(tmp_path, links):
    """A non-census key: the synthetic Poseidon-shaped chain circuit (BASELINE configs[3]), loaded without a wasm,
    proved from its .wtns; A/B/C == the CPU oracle with pinned r,s, and the proof verifies under the matching vkey
    on the GPU and on the CPU.  40 links = 17.5 k rows (one point range per MSM); 600 links = 262,803 wires, domain
    2^19: the witness MSMs run as 3 point ranges (the last one padded) and the H MSM as 4."""
    from zk_franchise_proof_circuit_b200 import prover
    n_wires, n_cons, domain = O.chain_artifacts(links, 7, str(tmp_path))
    assert n_cons == links * 438 + 1 and domain == (1 << 15 if links == 40 else 1 << 19)
    zkey = open(tmp_path / "proving_key.zkey", "rb").read()
    wtns = open(tmp_path / "witness.wtns", "rb").read()
    vkey = open(tmp_path / "verification_key.json", "rb").read()
    c = prover.load(zkey, None)
    assert c.n_vars == n_wires
    c.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        pj, sj = c.prove_wtns(wtns)
    finally:
        c.set_blinding(None, None)
    w = H.wtns_payload(wtns, n_wires)
    exp = O.ZKeyRef(zkey).prove(w, H.R_FIXED, H.S_FIXED)
    proof, pub = json.loads(pj), json.loads(sj)
    assert O.proof_bin(proof) == exp
    assert pub == [str(int.from_bytes(w[1].tobytes(), "little"))]
    if links == 40:                                  # committed known answer (tests/golden/chain40_kat.json)
        kat = json.load(open(H.GOLDEN + "/chain40_kat.json"))
        assert proof == kat["proof"] and pub == kat["public"]
    assert O.verify(json.loads(vkey), pub, proof)
    prover.verify(vkey, sj, pj)
    c.close()


def _sharded_proof(zkey, wtns, nranks, devices, slices=False):
    from zk_franchise_proof_circuit_b200 import prover
    cs = [prover.load_shard(zkey, r, nranks, device=devices[r % len(devices)]) for r in range(nranks)]
    for c in cs[1:]:
        c.shard_attach_local(cs[0])
    cs[0].set_blinding(H.R_FIXED, H.S_FIXED)
    if slices:
        # witness-slice exchange: the ranks wait for each other's slices on the device, so each rank needs its own
        # host thread (as in a multi-process run); ctypes releases the GIL during the call
        import threading
        for a in cs:
            for b in cs:
                if a is not b:
                    a.shard_attach_witness_local(b.shard[0], b)
        outs = []
        try:
            for _ in range(2):                 # two proofs: two epochs
                res = [None] * nranks

                def work(r):
                    try:
                        res[r] = cs[r].prove_wtns(wtns)
                    except Exception as e:      # noqa: BLE001
                        res[r] = e
                ts = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
                assert all(res[r] == (b"", b"") for r in range(1, nranks)), res[1:]
                assert not isinstance(res[0], Exception), res[0]
                outs.append(res[0])
            assert outs[0] == outs[1]
        finally:
            for c in cs:
                c.close()
        return outs[0]
    for c in cs[1:]:                       # single thread: the peers publish first, then rank 0 combines
        assert c.prove_wtns(wtns) == (b"", b"")
    pj, sj = cs[0].prove_wtns(wtns)
    # second proof: next epoch
    for c in cs[1:]:
        c.prove_wtns(wtns)
    assert cs[0].prove_wtns(wtns) == (pj, sj)
    for c in cs:
        c.close()
    return pj, sj


def test_sharded_key_matches_oracle(tmp_path):
    """262,803-wire chain circuit sharded over 2 and 3 ranks (contexts of one GPU): rank 0's proof == the CPU oracle
    with pinned r,s (the same bytes as the unsharded path), and it verifies."""
    from zk_franchise_proof_circuit_b200 import prover
    n_wires, _, _ = O.chain_artifacts(600, 7, str(tmp_path), check=False)
    zkey = open(tmp_path / "proving_key.zkey", "rb").read()
    wtns = open(tmp_path / "witness.wtns", "rb").read()
    vkey = open(tmp_path / "verification_key.json", "rb").read()
    exp = O.ZKeyRef(zkey).prove(H.wtns_payload(wtns, n_wires), H.R_FIXED, H.S_FIXED)
    for nranks in (2, 3):
        pj, sj = _sharded_proof(zkey, wtns, nranks, [0])
        assert O.proof_bin(json.loads(pj)) == exp
        prover.verify(vkey, sj, pj)
    # more ranks than 2^17-point ranges: the key is cut into finer ranges (8 ranks -> 2^15 points each)
    pj, sj = _sharded_proof(zkey, wtns, 8, [0])
    assert O.proof_bin(json.loads(pj)) == exp
    with pytest.raises(Exception):
        prover.load_shard(zkey, 0, 17)      # more than SHARD_MAX_RANKS


def test_sharded_key_two_gpus(tmp_path):
    from zk_franchise_proof_circuit_b200 import prover, _native
    if _native.lib().zkb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n_wires, _, _ = O.chain_artifacts(600, 7, str(tmp_path), check=False)
    zkey = open(tmp_path / "proving_key.zkey", "rb").read()
    wtns = open(tmp_path / "witness.wtns", "rb").read()
    exp = O.ZKeyRef(zkey).prove(H.wtns_payload(wtns, n_wires), H.R_FIXED, H.S_FIXED)
    pj, sj = _sharded_proof(zkey, wtns, 2, [0, 1])
    assert O.proof_bin(json.loads(pj)) == exp
    # witness slices over NVLink: every rank uploads only its slice of the .wtns and gathers the rest from its peer.
    # (Needs the ranks on different GPUs: they wait for each other with spinning kernels, and contexts of ONE device can
    # end up on the same hardware queue, where a spinning kernel blocks the kernel it waits for - DESIGN.md section 5.
    # With one process per GPU, as bench.py --gpus N runs it, that cannot happen.)
    pj, sj = _sharded_proof(zkey, wtns, 2, [0, 1], slices=True)
    assert O.proof_bin(json.loads(pj)) == exp


def test_concurrent_callers(art_dir):
    """cgo calls arrive on arbitrary OS threads (SURVEY 8b threading): 4 threads x 3 zkb_fullprove calls on one
    circuit handle; every proof verifies and the public signals are the fixture's."""
    import threading
    from zk_franchise_proof_circuit_b200 import prover
    c = prover.load(open(art_dir + "/proving_key.zkey", "rb").read(), open(art_dir + "/circuit.wasm", "rb").read())
    doc = open(H.GOLDEN + "/inputs_example.json", "rb").read()
    vkey = open(art_dir + "/verification_key.json", "rb").read()
    out, errs = [], []

    def work():
        try:
            for _ in range(3):
                out.append(c.fullprove(doc))
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs and len(out) == 12
    for pj, sj in out:
        assert json.loads(sj) == json.load(open(H.GOLDEN + "/signals.json"))
        prover.verify(vkey, sj, pj)
    assert len({pj for pj, _ in out}) == 12       # fresh blinding every time


def test_maximum_depth_census_bit_exact(circuit):
    """Edge of the circuit: a voter whose path has a non-zero sibling at all 160 levels (no level block equals the
    per-key template, the witness MSMs see ~80 k differing wires), next to depth-1 and depth-81 voters of the same
    census.  Witness == reference wasm, A/B/C == CPU oracle with pinned r,s, every proof verifies."""
    vs = H.deep_voters()
    assert sum(1 for x in vs[0]["censusSiblings"] if x != "0") == 160
    assert sum(1 for x in vs[1]["censusSiblings"] if x != "0") == 1
    circuit.set_blinding(H.R_FIXED, H.S_FIXED)
    try:
        proofs, pubs, status = circuit.fullprove_batch([json.dumps(v) for v in vs])
    finally:
        circuit.set_blinding(None, None)
    assert status == [0, 0, 0]
    for v, pj, sj in zip(vs, proofs, pubs):
        code, ref = _ref_witness(v)
        assert code == 0
        w = H.wtns_payload(circuit.witness(json.dumps(v)), circuit.n_vars)
        assert np.array_equal(w, ref)
        proof, pub = json.loads(pj), json.loads(sj)
        assert O.proof_bin(proof) == H.zkey_ref().prove(ref, H.R_FIXED, H.S_FIXED)
        assert O.verify(H.dev_vkey(), pub, proof)
    # one more siblings entry than the circuit allows -> the wasm's assert (code 4), not a proof
    bad = dict(vs[0])
    bad["censusSiblings"] = vs[0]["censusSiblings"][:160] + ["1"]
    assert circuit.fullprove_batch([json.dumps(bad)])[2] == [4]
    assert _ref_witness(bad)[0] == 4


def test_inputs_are_reduced_mod_r_like_circom_runtime(circuit):
    """circom_runtime normalises every input with Fr.e(): values >= r and negative values are taken mod r.  The same
    inputs.json with voteHash[0] + r and voteHash[1] - r must give the same witness and public signals."""
    inp = H.fixture_inputs()
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    alt = dict(inp)
    alt["voteHash"] = [str(int(inp["voteHash"][0]) + R), str(int(inp["voteHash"][1]) - R)]
    w0 = H.wtns_payload(circuit.witness(json.dumps(inp)), circuit.n_vars)
    w1 = H.wtns_payload(circuit.witness(json.dumps(alt)), circuit.n_vars)
    assert np.array_equal(w0, w1)
    _, sj = circuit.fullprove(json.dumps(alt))
    assert json.loads(sj) == json.load(open(H.GOLDEN + "/signals.json"))
    # bare JSON numbers are accepted as well as strings
    alt2 = dict(inp)
    alt2["voteWeight"] = int(inp["voteWeight"])
    assert np.array_equal(H.wtns_payload(circuit.witness(json.dumps(alt2)), circuit.n_vars), w0)
