#!/usr/bin/env python3
"""Workload for ncu captures of the generic witness evaluator (k_tape_eval): the census key loaded with the witness
program extracted from circuit.wasm, one group of proofs.  Usage (GPU box):
    ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats -k regex:'^k_tape_eval$' -c 1 \
        -o gpurun_out/tape python tools/prof_generic.py [n_proofs]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zk_franchise_proof_circuit_b200 import prover, census_tree
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
zkey, wasm = open(ART + "/proving_key.zkey", "rb").read(), open(ART + "/circuit.wasm", "rb").read()
c0 = prover.load(zkey, wasm)                       # the census generator hashes with the census circuit's Poseidon
voters = census_tree.gen_census(c0, 128, seed=0xC0FFEE)
docs = [json.dumps(voters[i % 128]) for i in range(n)]
c0.close()
prover._circuits.clear()
c = prover.load(zkey, wasm, generic=True)
proofs, pubs, status = c.fullprove_batch(docs)
assert status == [0] * n
print("prof_generic ok:", n, "proofs")
