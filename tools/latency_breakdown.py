#!/usr/bin/env python3
"""Stage times (CUDA events) of the proving pipeline for small batches - where single-proof latency goes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zk_franchise_proof_circuit_b200 import prover
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
c = prover.load(open(ART + "/proving_key.zkey", "rb").read(), open(ART + "/circuit.wasm", "rb").read())
inp = json.load(open(os.path.join(ROOT, "tests", "golden", "inputs_example.json")))
names = ("witness", "build_abc", "ntt_join", "msm_sort", "msm_acc_g1", "msm_acc_g2", "msm_reduce", "finalize")
for n in (1, 8, 64):
    c.set_inputs(np.stack([prover.pack_inputs(inp)] * n))
    c.prove_resident(n)
    st = c.prove_resident(n, stages=True)
    t0 = time.perf_counter(); c.prove_resident(n); dt = (time.perf_counter() - t0) * 1e3
    print(n, "proofs: wall", round(dt, 2), "ms;", {k: round(float(v), 2) for k, v in zip(names, st)})
