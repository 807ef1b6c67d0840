"""CPU tests: the oracle's witness side against the reference's golden vectors."""
import ctypes
import json
import os
import subprocess
import numpy as np
import pytest

import helpers as H
import oracle_lib as O


def _ref():
    import ref_witness as RW
    if not RW.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; run __graft_entry__.build())")
    return RW


def test_reference_wasm_witness_kat(art_dir):
    """The reference's own witness calculator, run natively, reproduces the golden witness."""
    RW = _ref()
    code, w = RW.witness(H.fixture_inputs())
    assert code == 0
    assert H.sha(w.tobytes()) == H.WITNESS_SHA256
    pub = json.load(open(H.GOLDEN + "/signals.json"))
    assert [str(O.from_le(w[i])) for i in range(1, 9)] == pub
    kat = json.load(open(H.GOLDEN + "/witness_kat.json"))
    assert kat["sha256"] == H.WITNESS_SHA256
    for i, v in kat["wires"].items():
        assert O.from_le(w[int(i)]) == int(v)


def test_reference_wasm_assert_codes(art_dir):
    RW = _ref()
    inp = H.fixture_inputs()
    assert RW.witness(dict(inp, voteWeight="11"))[0] == 4
    assert RW.witness(dict(inp, nullifier="5"))[0] == 4
    assert RW.witness(dict(inp, voteWeight="10"))[0] == 0      # voteWeight == availableWeight is allowed


def test_python_model_matches_wasm(art_dir):
    """census_model.py (restatement of census.circom + circomlib) == wasm, all 82,754 wires."""
    import census_model as M
    RW = _ref()
    inp = H.fixture_inputs()
    w = M.witness(H.poseidon_tables(), H.witness_map(), inp)
    code, ref = RW.witness(inp)
    got = b"".join(x.to_bytes(32, "little") for x in w)
    assert got == ref.tobytes()
    with pytest.raises(M.AssertFailed):
        M.witness(H.poseidon_tables(), H.witness_map(), dict(inp, voteWeight="11"))


def test_poseidon_kats(art_dir):
    """circomlib KAT poseidon([1,2]) and the fixture's nullifier / roots from the wasm's constants."""
    import census_gen as G
    P = G.Poseidon(H.poseidon_tables())
    assert P([1, 2]) == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    inp = H.fixture_inputs()
    sig, pw, el = int(inp["signature"]), int(inp["password"]), [int(x) for x in inp["electionId"]]
    assert P([sig, pw, el[0], el[1]]) == int(inp["nullifier"])
    sik = P([int(inp["address"]), pw, sig])
    assert sik == 2763692874536778083565625622297672041243620578525557176716292711186292779392
    for root_key, sib_key, value in (("censusRoot", "censusSiblings", int(inp["availableWeight"])),
                                     ("sikRoot", "sikSiblings", sik)):
        key = int(inp["address"])
        sibs = [int(x) for x in inp[sib_key]]
        depth = max(i for i, s in enumerate(sibs) if s) + 1
        node = P([key, value, 1])
        for d in range(depth - 1, -1, -1):
            node = P([sibs[d], node]) if (key >> d) & 1 else P([node, sibs[d]])
        assert node == int(inp[root_key])


def test_synthetic_voters_pass_the_wasm(art_dir):
    RW = _ref()
    vs = H.voters(16)
    for v in (vs[0], vs[7], vs[15]):
        assert RW.witness(v)[0] == 0
    assert len({v["address"] for v in vs}) == 16
    assert all(len(v["censusSiblings"]) == 161 and v["censusSiblings"][-1] == "0" for v in vs)


def test_input_encoding_kat():
    """example.ts client values -> inputs_example.json fields: sha256 split + LE address."""
    import census_gen as G
    inp = H.fixture_inputs()
    el = G.bytes_to_arbo(bytes.fromhex(G.ELECTION_HEX))
    assert [str(x) for x in el] == inp["electionId"]
    assert [str(x) for x in G.bytes_to_arbo((10).to_bytes(1, "big"))] == inp["voteHash"]
    assert str(int.from_bytes(b"password123", "big")) == inp["password"]


def test_host_build_of_witness_program_matches_wasm(art_dir, tmp_path):
    """csrc/census_witness.cuh compiled for the host (its portable path) == wasm, dense and template modes."""
    RW = _ref()
    src = os.path.join(H.ROOT, "tests", "host_emul", "witness_host.cc")
    so = str(tmp_path / "libwitness_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    L = ctypes.CDLL(so)
    T, wm = H.poseidon_tables(), H.witness_map()
    R = O.lib and 21888242871839275222246405745257275088548364400416034343698204186575808495617
    consts, lens = [], []
    for t in (3, 4, 5):
        for k in "CSMP":
            consts += [(x << 256) % R for x in T[k][t]]
            lens.append(len(T[k][t]))
    cb = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in consts), dtype=np.uint32).copy()
    lens = np.array(lens, dtype=np.uint32)
    nsig = L.host_layout_signals(161)
    assert nsig == 272804
    s2w = np.full(nsig, -1, dtype=np.int32)
    s2w[np.array(wm)] = np.arange(len(wm), dtype=np.int32)
    from zk_franchise_proof_circuit_b200.prover import pack_inputs

    def run(inp, skip, template=None):
        out = np.zeros((len(wm), 32), dtype=np.uint8) if template is None else template.copy()
        fi = pack_inputs(inp)
        st = L.host_witness(cb.ctypes.data, lens.ctypes.data, s2w.ctypes.data, nsig, 161, fi.ctypes.data,
                            out.ctypes.data, skip)
        return st, out

    inp = H.fixture_inputs()
    code, ref = RW.witness(inp)
    st, dense = run(inp, 0)
    assert st == 0 and np.array_equal(dense, ref)
    zero = {k: (["0"] * len(v) if isinstance(v, list) else "0") for k, v in inp.items()}
    _, tmpl = run(zero, 0)
    st, fast = run(inp, 1, tmpl)
    assert st == 0 and np.array_equal(fast, ref)
    for v in H.voters(16)[:3]:
        code, ref = RW.witness(v)
        st, fast = run(v, 1, tmpl)
        assert st == 0 and np.array_equal(fast, ref)
    assert run(dict(inp, voteWeight="11"), 1, tmpl)[0] == 4
    assert run(dict(inp, sikRoot="1"), 1, tmpl)[0] == 4
