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
