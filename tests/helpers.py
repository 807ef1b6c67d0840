"""Shared test helpers (oracle side).  Everything here is checker code, never product code."""
import functools
import hashlib
import json
import os
import numpy as np

import census_gen
import oracle_lib as O
import wasm_tools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
GOLDEN = os.path.join(ROOT, "tests", "golden")
WITNESS_SHA256 = "ebf5467e953a0427fa50c9a0b0521ac1c5c70684ef3603c75807c1bc4315e71b"
R_FIXED, S_FIXED = 1234567, 7654321


@functools.lru_cache(maxsize=None)
def wasm_module():
    return wasm_tools.Module(open(os.path.join(ART, "circuit.wasm"), "rb").read())


@functools.lru_cache(maxsize=None)
def poseidon_tables():
    return wasm_tools.poseidon_tables(wasm_module())


@functools.lru_cache(maxsize=None)
def witness_map():
    return wasm_tools.witness_map(wasm_module())[0]


def fixture_inputs():
    return json.load(open(os.path.join(GOLDEN, "inputs_example.json")))


@functools.lru_cache(maxsize=None)
def voters(n, seed=0xC0FFEE):
    return census_gen.gen_census(poseidon_tables(), n, seed=seed)


@functools.lru_cache(maxsize=None)
def zkey_ref():
    return O.ZKeyRef(open(os.path.join(ART, "proving_key.zkey"), "rb").read())


def dev_vkey():
    return json.load(open(os.path.join(ART, "verification_key.json")))


def wtns_payload(wtns: bytes, n_vars: int) -> np.ndarray:
    """strip the .wtns header -> uint8[n_vars, 32]"""
    assert wtns[:4] == b"wtns"
    return np.frombuffer(wtns[-n_vars * 32:], dtype=np.uint8).reshape(n_vars, 32)


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


@functools.lru_cache(maxsize=None)
def deep_voters():
    """Maximum-depth census (checker side): 160 leaves whose addresses are `base` with one bit flipped, plus `base`
    itself, so that base's Merkle path has a NON-ZERO sibling at every one of the 160 levels (the circuit's limit:
    siblings[160] must be zero, census.circom:79-103 / SMTLevIns).  Returns inputs for `base` (depth 160), for the
    leaf that splits off at level 0 (depth 1) and for the one that splits off at level 80 (depth 81)."""
    G = census_gen
    Hh = G.Poseidon(poseidon_tables())
    base = int.from_bytes(hashlib.sha256(b"deep-census").digest()[:20], "little")
    addrs = [base] + [base ^ (1 << j) for j in range(160)]
    password = int.from_bytes(b"password123", "big") % G.P
    election = G.bytes_to_arbo(bytes.fromhex(G.ELECTION_HEX))
    vote_hash = G.bytes_to_arbo((10).to_bytes(1, "big"))
    sigs = {a: int.from_bytes(hashlib.sha256(b"sig" + a.to_bytes(20, "little")).digest() * 2, "big") % G.P for a in addrs}
    siks = {a: Hh([a, password, sigs[a]]) for a in addrs}
    census = G.SMT(Hh, {a: 10 for a in addrs})
    siktree = G.SMT(Hh, siks)
    out = []
    for a in (base, base ^ 1, base ^ (1 << 80)):
        pad = lambda s: [str(x) for x in s] + ["0"] * (161 - len(s))
        out.append({
            "electionId": [str(election[0]), str(election[1])],
            "nullifier": str(Hh([sigs[a], password, election[0], election[1]])),
            "availableWeight": "10", "voteHash": [str(vote_hash[0]), str(vote_hash[1])],
            "sikRoot": str(siktree.root), "censusRoot": str(census.root), "address": str(a),
            "password": str(password), "signature": str(sigs[a]), "voteWeight": "5",
            "censusSiblings": pad(census.siblings(a)), "sikSiblings": pad(siktree.siblings(a))})
    return out
