#!/usr/bin/env python3
"""Workload for ncu captures: one chunk (default 128 proofs) of the synthetic 1,024-voter census through the
resident proving path, twice.  Usage (on the GPU box, one GPU):
    ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass|k_build_abc' -c 12 \
        -o gpurun_out/prof python tools/prof_driver.py [n_proofs] [passes]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zk_franchise_proof_circuit_b200 import prover, census_tree
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = prover.load(open(ART + "/proving_key.zkey", "rb").read(), open(ART + "/circuit.wasm", "rb").read())
voters = census_tree.gen_census(c, 1024, seed=0xC0FFEE)
c.set_inputs(np.stack([prover.pack_inputs(voters[i % 1024]) for i in range(n)]))
for _ in range(passes):
    c.prove_resident(n)
_, _, status = c.get_results(n)
assert (status == 0).all()
print("prof_driver ok:", n, "proofs x", passes)
