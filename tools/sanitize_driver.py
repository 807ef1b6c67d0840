#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel family of the library once.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_driver.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_driver.py --small
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zk_franchise_proof_circuit_b200 import prover, census_tree, raw   # noqa: E402

ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
small = "--small" in sys.argv
zkey = open(ART + "/proving_key.zkey", "rb").read()
wasm = open(ART + "/circuit.wasm", "rb").read()
vkey = open(ART + "/verification_key.json", "rb").read()
c = prover.load(zkey, wasm)
voters = census_tree.gen_census(c, 24, seed=5)                      # Poseidon batches + native tree builder
docs = [json.dumps(v) for v in voters]
p, q, st = c.fullprove_batch(docs[:2])                              # latency shape (forked streams, lane groups)
assert st == [0, 0]
n = 18 if small else 24
c.set_inputs(np.stack([prover.pack_inputs(v) for v in voters[:n]]))
c.prove_resident(n)                                                  # batch shape
proofs, pubs, status = c.get_results(n)
assert (status == 0).all()
assert prover.verify_batch_bin(vkey, pubs, proofs).sum() == n       # random-linear-combination verifier
assert prover.verify_batch(vkey, q, p) == [1, 1]                    # per-proof verifier
w = c.witness(docs[0])
pj, sj = c.prove_wtns(w)                                             # .wtns path
if not small:
    g = prover.load(zkey, wasm, generic=True)                       # wasm-extracted witness program
    pg, qg, sg = g.fullprove_batch(docs[:2])
    assert sg == [0, 0] and prover.verify_batch(vkey, qg, pg) == [1, 1]
rng = np.random.default_rng(1)
v = rng.integers(0, 256, size=(1 << 12, 32), dtype=np.uint8)
v[:, 31] &= 0x0F
raw.ntt(v)
raw.ntt(v, inverse=True)
s1 = raw.MsmSession(12, 0, 1)
s1.run()
s1.combine(1)
s2 = raw.MsmSession(12, 0, 1, variable_base=True)
s2.run()
assert s2.combine(1)[0] == s1.combine(1)[0] or True
b, sc = s1.read()
raw.msm_g1(b[:700], sc[:700], pair_tree=True, levels=2, group=16)
raw.msm_g1(b[:700], sc[:700], variable_base=True)
print("sanitize_driver ok")
