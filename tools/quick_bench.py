#!/usr/bin/env python3
"""Development aid: resident throughput (proofs/s) and serial stage times for the 1,024-proof census batch,
without the CPU baseline / e2e / latency legs of bench.py.  Env: ZKB_LANES, ZKB_CHUNK, ... as the library reads."""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zk_franchise_proof_circuit_b200 import prover, census_tree
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = prover.load(open(ART + "/proving_key.zkey", "rb").read(), open(ART + "/circuit.wasm", "rb").read())
voters = census_tree.gen_census(c, min(n, 1024), seed=0xC0FFEE)
c.set_inputs(np.stack([prover.pack_inputs(voters[i % len(voters)]) for i in range(n)]))
for _ in range(2):
    c.prove_resident(n)
t0 = time.perf_counter()
for _ in range(steps):
    c.prove_resident(n)
dt = (time.perf_counter() - t0) / steps
names = ("witness", "build_abc", "ntt_join", "msm_sort", "msm_acc_g1", "msm_acc_g2", "msm_reduce", "finalize")
st = c.prove_resident(n, stages=True)
_, _, status = c.get_results(n)
assert (status == 0).all()
print(f"lanes={os.environ.get('ZKB_LANES', 'dflt')} chunk={os.environ.get('ZKB_CHUNK', 'dflt')} "
      f"{n / dt:.1f} proofs/s ({dt * 1e3:.1f} ms/step); serial stages ms:",
      {k: round(float(v), 1) for k, v in zip(names, st)}, "sum", round(float(sum(st)), 1))
